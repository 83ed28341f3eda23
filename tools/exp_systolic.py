"""Systolic kernel probes on one GPU: time per row-step, residency and lane-width sweep.  Prints JSON lines."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cse305_parallel_sequence_alignment_b200 as psa  # noqa: E402
from cse305_parallel_sequence_alignment_b200 import synth  # noqa: E402

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
A, B = synth.mutated_pair(1_000_000, synth.SEED_C4)
dA, dB = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)


def time_long(ctx, m, n, mode=psa.LOCAL, reps=2):
    item = torch.zeros(10, dtype=torch.int32, device=dev)
    f = lambda: ctx.align_long_device(dA.data_ptr(), dB.data_ptr(), m, n, item.data_ptr(), 0, 0, mode, 1, 2, False, stream.cuda_stream)
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(stream); f(); e1.record(stream); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    it = item.cpu()
    return float(min(ts)), (int(it[3]), int(it[5]), int(it[6]))


# CONFIGS="kc,rb,warps_per_sm;..."  SIZES="mxn;..."  ROWBLOCK=0 skips the row-block reference run
CONFIGS = [tuple(int(x) for x in c.split(",")) for c in os.environ.get("CONFIGS", "4,4,8;4,4,16;8,4,8;4,2,8").split(";")]
SIZES = [tuple(int(x) for x in c.split("x")) for c in os.environ.get("SIZES", "").split(";") if c]
for kc, rb, wpsm in CONFIGS:
    ctx = psa.Context(0)
    ctx.set_option("long_systolic", 1)
    ctx.set_option("systolic_kc", kc)
    ctx.set_option("systolic_rb", rb)
    ctx.set_option("systolic_warps_per_sm", wpsm)
    for (m, n) in (SIZES or ((1_000_000, 32 * kc * 200), (1_000_000, 125_000), (1_000_000, 1_000_000))):
        ms, res = time_long(ctx, m, n)
        strips = (n + 32 * kc - 1) // (32 * kc)
        print(json.dumps({"probe": "systolic", "kc": kc, "rb": rb, "warps_per_sm": wpsm, "m": m, "n": n, "ms": ms, "gcups": m * n / ms / 1e6,
                          "strips": strips, "res": res}), flush=True)
    ctx.close()
if os.environ.get("ROWBLOCK", "1") == "0":
    sys.exit(0)
ctx = psa.Context(0)
ctx.set_option("long_systolic", 0)
ms, res = time_long(ctx, 1_000_000, 1_000_000)
print(json.dumps({"probe": "rowblock", "m": 1_000_000, "n": 1_000_000, "ms": ms, "gcups": 1e12 / ms / 1e6, "res": res}), flush=True)
