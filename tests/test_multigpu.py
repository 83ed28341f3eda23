"""Multi-GPU paths.  CPU: the panel partition / merge logic.  GPU, any box: the multi-rank data path of the cyclic
panels (rings between ranks, lap tags, back-pressure, back-to-back calls) with 2 and 3 ranks sharing ONE device
against the oracle.  GPU, >= 2 devices (skipped on a 1-GPU box): the same over NVLink peer mappings against the
1-GPU result and the oracle, launched with torchrun (one process per GPU)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from cse305_parallel_sequence_alignment_b200 import multigpu
from cse305_parallel_sequence_alignment_b200.capi import ITEM_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_panels_cover_every_column_once_and_balance():
    for n in (100, 256, 1000, 100_000, 1_000_000, 999_999):
        for world in (1, 2, 3, 4, 8):
            for cap, sc in ((8, 256), (1184, 256), (1184, 128)):
                ps = multigpu.balanced_panel_strips(n, world, cap, sc)
                assert 1 <= ps <= cap
                per_rank = multigpu.panel_owner_ranges(n, world, ps, sc)
                flat = sorted(r for ranges in per_rank for r in ranges)
                assert flat[0][0] == 0 and flat[-1][1] == n
                assert all(flat[k][1] == flat[k + 1][0] for k in range(len(flat) - 1))
                assert all((c1 - c0) == ps * sc for c0, c1 in flat[:-1])
                # panel q belongs to rank q mod world, and nobody has more than one panel more than anybody else
                for q, rng in enumerate(flat):
                    assert rng in per_rank[q % world]
                counts = [len(x) for x in per_rank]
                assert max(counts) - min(counts) <= 1
                assert multigpu.last_panel_rank(n, world, ps, sc) == (len(flat) - 1) % world


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
def test_cyclic_panels_match_single_gpu_and_oracle(mode):
    n = _ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    # a small forced panel width (8 strips) makes every rank run many panels through the NVLink rings
    env = dict(os.environ, C4_LEN="30000", MODE=str(mode), REPS="2", PANEL_STRIPS="8")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + mode), os.path.join(ROOT, "tools", "strip_bench.py")]
    run = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr[-2000:]
    line = [x for x in run.stdout.splitlines() if x.startswith("{")][-1]
    out = json.loads(line)
    assert out["matches_1gpu"] and out["matches_oracle"], out


@pytest.mark.gpu
def test_cyclic_panels_ranks_sharing_one_gpu():
    """W ranks = W contexts + streams on device 0, the inter-rank rings plain device buffers (tools/cyclic_one_gpu.py):
    what a multi-GPU run executes minus the IPC mapping, so a 1-GPU box checks it too.  Run in a child process with a
    time limit, so that a ring protocol error shows up as a failure, not as a hang of the whole test run."""
    variants = [{"world": 2, "panel_strips": 8, "kc": 8, "mode": 1}, {"world": 3, "panel_strips": 8, "kc": 8, "mode": 0},
                {"world": 3, "panel_strips": 5, "kc": 4, "mode": 1}, {"world": 2, "panel_strips": 16, "kc": 4, "mode": 0}]
    env = dict(os.environ, M="9000", N="20000", CALLS="3", VARIANTS=json.dumps(variants))
    run = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "cyclic_one_gpu.py")], env=env, capture_output=True,
                         text=True, timeout=240)
    assert run.returncode == 0, run.stderr[-2000:]
    outs = [json.loads(x) for x in run.stdout.splitlines() if x.startswith("{")]
    assert len(outs) == len(variants)
    for out in outs:
        assert out["matches_oracle"] and out["single_rank_matches_oracle"], out
        assert min(out["panels_per_rank"]) >= 2, out            # every rank really chains several panels through its ring
