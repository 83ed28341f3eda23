"""One-off GPU probe: integer-pipe microbenchmark + a quick timing of the batch kernel.
Writes gpurun_out/probe.json.  Not part of the product or the tests."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cse305_parallel_sequence_alignment_b200 as psa  # noqa: E402

KINDS = {0: "VIADDMNMX.s32", 1: "VIADDMNMX.s16x2", 2: "VIMNMX3.s32", 3: "VIMNMX3.s16x2", 4: "IADD3", 5: "IMAD",
         6: "ALU+FMA 1:1", 7: "PRMT", 8: "IDP.4A", 9: "VIADD.16x2", 10: "LOP3", 11: "SHF", 12: "cell mix s16x2",
         13: "VIMNMX.s16x2"}


def main():
    out = {}
    ctx = psa.Context(0)
    peaks = {}
    for k, name in KINDS.items():
        v, ms = ctx.peak_int_ops(k)
        peaks[name] = {"Tlaneops_s": v / 1e12, "ms": ms}
        print(f"{name:18s} {v/1e12:8.2f} Tlane-ops/s  ({ms:.3f} ms)")
    out["peaks"] = peaks
    n = int(os.environ.get("PROBE_PAIRS", "262144"))
    rng = np.random.default_rng(1)
    A = rng.integers(0, 4, size=(n, 150), dtype=np.uint8)
    B = A.copy()
    mut = rng.random((n, 150)) < 0.06
    B[mut] = (B[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) % 4
    B[1::2] = rng.integers(0, 4, size=(len(B[1::2]), 150), dtype=np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    dA = torch.from_numpy(lut[A].reshape(-1)).cuda()
    dB = torch.from_numpy(lut[B].reshape(-1)).cuda()
    off = torch.arange(n, dtype=torch.int64, device="cuda") * 150
    ln = torch.full((n,), 150, dtype=torch.int32, device="cuda")
    items = torch.zeros(n * 10, dtype=torch.int32, device="cuda")
    stride = 20
    ops = torch.zeros(n * stride, dtype=torch.int32, device="cuda")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)      # events below are recorded on the stream the kernels run on
    st = stream.cuda_stream
    for mode, tb in [(psa.LOCAL, True), (psa.LOCAL, False), (psa.GLOBAL, True), (psa.GLOBAL, False)]:
        def run():
            ctx.align_batch_device(dA.data_ptr(), off.data_ptr(), ln.data_ptr(), dB.data_ptr(), off.data_ptr(),
                                   ln.data_ptr(), n, 150, 150, items.data_ptr(), ops.data_ptr(), stride, mode, 1, 2,
                                   tb, st)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        gcups = n * 150 * 150 / (ms * 1e-3) / 1e9
        key = f"short_{'local' if mode else 'global'}_{'tb' if tb else 'score'}"
        out[key] = {"ms": ms, "gcups": gcups, "pairs": n}
        print(f"{key:22s} {ms:8.3f} ms  {gcups:9.1f} GCUPS")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
