"""Multi-GPU paths.  CPU: the strip partition / merge logic.  GPU (needs >= 2 devices, skipped on a
1-GPU box): the NVLink column-strip pipeline of one long pair against the 1-GPU result and the
oracle, launched with torchrun (one process per GPU)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from cse305_parallel_sequence_alignment_b200 import multigpu
from cse305_parallel_sequence_alignment_b200.capi import ITEM_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_strip_ranges_cover_and_align():
    for n in (256, 1000, 100_000, 1_000_000, 999_999):
        for world in (1, 2, 4, 8):
            r = multigpu.strip_ranges(n, world)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            # every strip that has a non-empty right neighbour is a whole number of 256-column tiles
            assert all((r[k][1] - r[k][0]) % multigpu.STRIP_ALIGN == 0 for k in range(world - 1) if r[k + 1][1] > r[k + 1][0])
            if n >= world * multigpu.STRIP_ALIGN * 4:
                sizes = [c1 - c0 for c0, c1 in r]
                assert max(sizes) - min(sizes) < 2 * multigpu.STRIP_ALIGN


def test_merge_local_results_order():
    it = np.zeros(3, dtype=ITEM_DTYPE)
    it["score"] = [10, 12, 12]
    it["end_i"] = [5, 9, 7]
    it["end_j"] = [1, 2, 900]
    best = multigpu.merge_local_results(it)
    assert (best["score"], best["end_i"], best["end_j"]) == (12, 7, 900)


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
def test_strip_pipeline_matches_single_gpu_and_oracle(mode):
    n = _ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ, C4_LEN="30000", MODE=str(mode), REPS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + mode), os.path.join(ROOT, "tools", "strip_bench.py")]
    run = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr[-2000:]
    line = [x for x in run.stdout.splitlines() if x.startswith("{")][-1]
    out = json.loads(line)
    assert out["matches_1gpu"] and out["matches_oracle"], out
