"""Config 2 on one GPU, device-resident: whole step, fill launches alone and score-only, for both traceback
flavours of the packed kernel (option pack_traceback: 0 = direction-code ring, 1 = checkpoints + tile
recompute), plus a bit-for-bit comparison of their results.  Prints JSON lines."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cse305_parallel_sequence_alignment_b200 as psa  # noqa: E402
from cse305_parallel_sequence_alignment_b200 import synth  # noqa: E402
from cse305_parallel_sequence_alignment_b200.capi import ITEM_DTYPE  # noqa: E402

N = int(os.environ.get("PAIRS", "1000000"))
L = int(os.environ.get("READ_LEN", "150"))
MODE = int(os.environ.get("MODE", "1"))
STEPS = int(os.environ.get("STEPS", "10"))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
A, B = synth.read_pair_batch(N, L, synth.SEED_C2)
off, ln = synth.fixed_length_layout(N, L)
dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
dOff, dLen = torch.from_numpy(off).cuda(), torch.from_numpy(ln).cuda()
stride = (2 * L + 15) // 16 + 1


def run(ctx, items, ops, tb):
    ctx.align_batch_device(dA.data_ptr(), dOff.data_ptr(), dLen.data_ptr(), dB.data_ptr(), dOff.data_ptr(), dLen.data_ptr(),
                           N, L, L, items.data_ptr(), ops.data_ptr() if tb else 0, stride if tb else 0, MODE, 1, 2, tb,
                           stream.cuda_stream)


def timed(ctx, items, ops, tb):
    for _ in range(3):
        run(ctx, items, ops, tb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(STEPS):
        run(ctx, items, ops, tb)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / STEPS


results = {}
FLAVOURS = [int(x) for x in os.environ.get("FLAVOURS", "0,1").split(",")]
for flavour in FLAVOURS:
    ctx = psa.Context(0)
    ctx.set_option("pack_traceback", flavour)
    for k, v in (json.loads(os.environ.get("OPTS", "{}"))).items():
        ctx.set_option(k, v)
    items = torch.zeros((N, 10), dtype=torch.int32, device="cuda")
    ops = torch.zeros((N, stride), dtype=torch.int32, device="cuda")
    ms = timed(ctx, items, ops, True)
    ctx.set_option("pack_skip_walk", 1)
    ms_fill = timed(ctx, items, ops, True)
    ctx.set_option("pack_skip_walk", 0)
    ms_score = timed(ctx, items, ops, False)
    run(ctx, items, ops, True)
    torch.cuda.synchronize()
    results[flavour] = (items.cpu().numpy().reshape(-1).view(ITEM_DTYPE), ops.cpu().numpy().view(np.uint32))
    cells = N * L * L
    print(json.dumps({"flavour": "checkpoint+recompute" if flavour == 1 else "code ring", "pairs": N, "len": L, "mode": MODE,
                      "step_ms": ms, "fill_only_ms": ms_fill, "score_only_ms": ms_score, "gcups": cells / ms / 1e6,
                      "whole_step_frac_of_18.5T": cells * (7 if MODE else 6) / 2 / (ms * 1e-3) / 18.5e12}), flush=True)
    ctx.close()
if len(results) < 2:
    sys.exit(0)
a, b = results[0], results[1]
same_items = all(np.array_equal(a[0][f], b[0][f]) for f in ("score", "end_i", "end_j", "start_i", "start_j", "aln_len", "t1", "t2", "t3", "end_state"))
words = (a[0]["aln_len"] + 15) // 16
mask = np.arange(a[1].shape[1])[None, :] < words[:, None]
print(json.dumps({"flavours_agree_items": bool(same_items), "flavours_agree_ops": bool(np.array_equal(a[1][mask], b[1][mask]))}), flush=True)
