"""Round-2 design probes (one gpurun call): lone-warp tile latency of the long-pair geometries, and the
systolic (column-stationary) kernel at several residencies.  Prints JSON lines."""
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


def time_long(ctx, m, n, mode=psa.LOCAL, reps=3):
    A, B = synth.mutated_pair(max(m, n), synth.SEED_C4)
    dA, dB = torch.from_numpy(A[:m].copy()).to(dev), torch.from_numpy(B[:n].copy()).to(dev)
    item = torch.zeros(10, dtype=torch.int32, device=dev)
    f = lambda: ctx.align_long_device(dA.data_ptr(), dB.data_ptr(), m, n, item.data_ptr(), 0, 0, mode, 1, 2, False, stream.cuda_stream)
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(stream); f(); e1.record(stream); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), int(item.cpu()[3])


out = []
# (a) a single row block: one warp sweeps every tile of the pair -> time per tile of a LONE warp
for geo, R, K in ((0, 128, 8), (4, 128, 4), (6, 128, 24), (7, 256, 24)):
    ctx = psa.Context(0)
    ctx.set_option("long_geometry", geo)
    n = 200_000
    ms, sc = time_long(ctx, R, n)
    tiles = (n + 32 * K - 1) // (32 * K)
    rec = {"probe": "lone_warp_tile", "geo": geo, "R": R, "K": K, "ms": ms, "tiles": tiles, "us_per_tile": ms * 1e3 / tiles,
           "ns_per_step": ms * 1e6 / tiles / (R + 31)}
    print(json.dumps(rec), flush=True)
    # many row blocks, same width: the throughput regime
    ms, sc = time_long(ctx, 200_000, n)
    print(json.dumps({"probe": "square_200k", "geo": geo, "ms": ms, "gcups": 200_000 * n / ms / 1e6, "score": sc}), flush=True)
    ctx.close()

# (b) the systolic kernel at several residencies
for wpsm in (4, 8, 12, 16, 24):
    ctx = psa.Context(0)
    ctx.set_option("long_systolic", 1)
    ctx.set_option("systolic_warps_per_sm", wpsm)
    for (m, n) in ((200_000, 200_000), (1_000_000, 60_000)):
        ms, sc = time_long(ctx, m, n)
        print(json.dumps({"probe": "systolic", "warps_per_sm": wpsm, "m": m, "n": n, "ms": ms, "gcups": m * n / ms / 1e6,
                          "ns_per_row": ms * 1e6 / (m + 32 * ((n + 127) // 128)), "score": sc}), flush=True)
    ctx.close()
