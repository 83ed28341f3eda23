"""Column-strip pipeline of ONE long pair over the GPUs of a node (BASELINE config 4).

One process per GPU (torch.distributed is only the plumbing: it carries the 64-byte CUDA IPC
handles at setup and the 40-byte result records at the end).  On the data path every rank's
kernel stores its strip's right boundary column directly into the next rank's HBM over NVLink
and releases a system-scope flag -- see psa_align_long_strip_device in include/psa.h.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

STRIP_ALIGN = 256        # every strip but the last is a multiple of the tile width


def strip_ranges(n_total: int, world: int) -> List[Tuple[int, int]]:
    """Column ranges [c0, c1) per rank: equal shares rounded to 256 columns, the last takes the rest.
    Ranks whose share would be empty get (n_total, n_total)."""
    tiles = (n_total + STRIP_ALIGN - 1) // STRIP_ALIGN
    out, c = [], 0
    for r in range(world):
        share = tiles // world + (1 if r < tiles % world else 0)
        c1 = min(n_total, c + share * STRIP_ALIGN)
        if r == world - 1:
            c1 = n_total
        out.append((c, c1))
        c = c1
    return out


def merge_local_results(items: np.ndarray) -> np.ndarray:
    """Local mode: best (score desc, end_i asc, end_j asc) over the ranks' strip-local results."""
    order = sorted(range(len(items)), key=lambda k: (-int(items[k]["score"]), int(items[k]["end_i"]), int(items[k]["end_j"])))
    return items[order[0]]


class StripPipeline:
    """Per-rank state: own incoming buffer, mapped pointer to the next rank's incoming buffer."""

    def __init__(self, ctx, m_cap: int, rank: int, world: int):
        import torch.distributed as dist
        self.ctx, self.rank, self.world, self.m_cap = ctx, rank, world, m_cap
        self.xin, handle = ctx.xbuf_create(m_cap)
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, handle)
        else:
            handles = [handle]
        self.xout = ctx.xbuf_open(handles[rank + 1]) if rank + 1 < world else 0
        self.epoch = 0
        if world > 1:
            dist.barrier()

    def run(self, d_a: int, d_b_strip: int, m: int, c0: int, c1: int, n_total: int, d_item: int, mode: int, g: int = 1,
            h: int = 2, stream: int = 0):
        """Launches this rank's strip [c0, c1); asynchronous on `stream`.  All ranks call it the
        same number of times (the epoch is the call counter)."""
        self.epoch += 1
        if c1 <= c0:
            return          # more ranks than 256-column tiles: this rank owns no columns (its epoch still advances)
        first, last = (c0 == 0), (c1 == n_total)
        self.ctx.align_long_strip_device(d_a, d_b_strip, m, c1 - c0, c0, n_total, d_item, self.m_cap,
                                         0 if first else self.xin, 0 if last else self.xout, self.epoch, mode, g, h, stream)

    def close(self):
        if self.xout:
            self.ctx.xbuf_close(self.xout)
            self.xout = 0
        if self.xin:
            self.ctx.xbuf_destroy(self.xin)
            self.xin = 0
