"""ONE long pair over the GPUs of a node (BASELINE config 4): block-cyclic systolic panels.

One process per GPU (torch.distributed is only the plumbing: it carries the 64-byte CUDA IPC handles at setup, the
panel width and the 40-byte result records).  The matrix is cut into panels of `panel_strips` strips of
psa_long_strip_columns() = 256 columns; panel
q belongs to rank q mod world.  On the data path the last strip of a panel stores its boundary column -- 8 bytes per
row, validity tag in-band -- directly into the next rank's ring buffer over NVLink; no NCCL, no barrier between
calls.  See psa_align_long_cyclic_device in include/psa.h.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

STRIP_COLS = 256          # psa_long_strip_columns() of a default context (8 columns per lane)


def panel_owner_ranges(n_total: int, world: int, panel_strips: int, strip_cols: int = STRIP_COLS) -> List[List[Tuple[int, int]]]:
    """Column ranges [c0, c1) of the panels of every rank (panel q -> rank q mod world)."""
    pw = panel_strips * strip_cols
    out: List[List[Tuple[int, int]]] = [[] for _ in range(world)]
    q, c = 0, 0
    while c < n_total:
        c1 = min(n_total, c + pw)
        out[q % world].append((c, c1))
        c, q = c1, q + 1
    return out


def balanced_panel_strips(n_total: int, world: int, capacity: int, strip_cols: int = STRIP_COLS) -> int:
    """Panel width (in strips) that gives every rank the same number of equally wide panels: the smallest number of
    rounds whose panels fit the resident capacity."""
    strips = (n_total + strip_cols - 1) // strip_cols
    rounds = 1
    while (strips + rounds * world - 1) // (rounds * world) > capacity:
        rounds += 1
    return max(1, (strips + rounds * world - 1) // (rounds * world))


def merge_local_results(items: np.ndarray) -> np.ndarray:
    """Local mode: best (score desc, end_i asc, end_j asc) over the ranks' results."""
    order = sorted(range(len(items)), key=lambda k: (-int(items[k]["score"]), int(items[k]["end_i"]), int(items[k]["end_j"])))
    return items[order[0]]


def last_panel_rank(n_total: int, world: int, panel_strips: int, strip_cols: int = STRIP_COLS) -> int:
    """Global mode: the rank whose item holds T1/T2/T3[m][n]."""
    pw = panel_strips * strip_cols
    return ((n_total + pw - 1) // pw - 1) % world


class CyclicPanels:
    """Per-rank state: own incoming ring, mapped pointer to the next rank's incoming ring, agreed panel width."""

    def __init__(self, ctx, m_cap: int, rank: int, world: int):
        import torch
        import torch.distributed as dist
        self.ctx, self.rank, self.world, self.m_cap = ctx, rank, world, m_cap
        self.xin = self.xout = 0
        cap = ctx.long_panel_strips
        self.strip_cols = ctx.long_strip_columns
        if world > 1:
            t = torch.tensor([cap], dtype=torch.int64, device=f"cuda:{ctx.device}")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            cap = int(t.item())
            self.xin, handle = ctx.xbuf_create(m_cap)
            handles = [None] * world
            dist.all_gather_object(handles, handle)
            self.xout = ctx.xbuf_open(handles[(rank + 1) % world])
            dist.barrier()
        self.capacity = cap

    def panel_strips(self, n_total: int) -> int:
        return balanced_panel_strips(n_total, self.world, self.capacity, self.strip_cols)

    def run(self, d_a: int, d_b: int, m: int, n_total: int, d_item: int, mode: int, g: int = 1, h: int = 2, stream: int = 0):
        """Launches this rank's panels; asynchronous on `stream`.  Every rank calls it with the same arguments."""
        self.ctx.align_long_cyclic_device(d_a, d_b, m, n_total, self.rank, self.world, self.panel_strips(n_total), d_item,
                                          self.m_cap, self.xin, self.xout, mode, g, h, stream)

    def close(self):
        if self.xout:
            self.ctx.xbuf_close(self.xout)
            self.xout = 0
        if self.xin:
            self.ctx.xbuf_destroy(self.xin)
            self.xin = 0
