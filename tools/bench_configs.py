"""Times every BASELINE config shape on one GPU (device-resident inputs, CUDA events) and writes
gpurun_out/configs.json.  Sizes of C4/C5 can be reduced with env vars for quick runs."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cse305_parallel_sequence_alignment_b200 as psa  # noqa: E402
from cse305_parallel_sequence_alignment_b200 import synth  # noqa: E402


def timed(fn, stream, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ctx = psa.Context(0)
for _k, _v in __import__('json').loads(os.environ.get('OPTS', '{}')).items():   # psa_ctx options, e.g. OPTS='{"long_geometry": 6}'
    ctx.set_option(_k, _v)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    out = {}

    def long_pair(name, L, mode, tb, seed, reps=3):
        A, B = synth.mutated_pair(L, seed)
        dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
        item = torch.zeros(10, dtype=torch.int32, device="cuda")
        words = (2 * L + 15) // 16 + 1
        ops = torch.zeros(words if tb else 1, dtype=torch.int32, device="cuda")
        ms = timed(lambda: ctx.align_long_device(dA.data_ptr(), dB.data_ptr(), L, L, item.data_ptr(), ops.data_ptr() if tb else 0,
                                                 words if tb else 0, mode, 1, 2, tb, st), stream, reps=reps)
        it = item.cpu().numpy()
        out[name] = {"ms": ms, "gcups": L * L / ms / 1e6, "score": int(it[3]), "aln_len": int(it[9])}
        print(name, out[name], flush=True)

    def batch(name, n, L, mode, tb, seed, reps=3):
        A, B = synth.read_pair_batch(n, L, seed)
        off, ln = synth.fixed_length_layout(n, L)
        dA, dB = torch.from_numpy(A.reshape(-1)).cuda(), torch.from_numpy(B.reshape(-1)).cuda()
        dOff, dLen = torch.from_numpy(off).cuda(), torch.from_numpy(ln).cuda()
        items = torch.zeros(n * 10, dtype=torch.int32, device="cuda")
        stride = (2 * L + 15) // 16 + 1
        ops = torch.zeros(n * stride if tb else 1, dtype=torch.int32, device="cuda")
        ms = timed(lambda: ctx.align_batch_device(dA.data_ptr(), dOff.data_ptr(), dLen.data_ptr(), dB.data_ptr(), dOff.data_ptr(),
                                                  dLen.data_ptr(), n, L, L, items.data_ptr(), ops.data_ptr() if tb else 0,
                                                  stride if tb else 0, mode, 1, 2, tb, st), stream, reps=reps)
        out[name] = {"ms": ms, "gcups": n * L * L / ms / 1e6, "pairs": n}
        print(name, out[name], flush=True)

    batch("C2_1M_150_local_tb", int(os.environ.get("C2_PAIRS", "1000000")), 150, psa.LOCAL, True, synth.SEED_C2)
    batch("C2_1M_150_local_score", int(os.environ.get("C2_PAIRS", "1000000")), 150, psa.LOCAL, False, synth.SEED_C2)
    long_pair("C3_10k_global_tb", 10000, psa.GLOBAL, True, synth.SEED_C3)
    long_pair("C3_10k_global_score", 10000, psa.GLOBAL, False, synth.SEED_C3)
    long_pair("C4_100k_local_score", 100000, psa.LOCAL, False, synth.SEED_C4)
    L4 = int(os.environ.get("C4_LEN", "1000000"))
    long_pair(f"C4_{L4}_local_score", L4, psa.LOCAL, False, synth.SEED_C4, reps=1)
    batch("C5_5k_local_score", int(os.environ.get("C5_PAIRS", "4096")), 5000, psa.LOCAL, False, synth.SEED_C5, reps=1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
