"""Bandwidth of psa_similarity_batch_device: 1M pairs of 150 bp and 64 pairs of 4 Mbp."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cse305_parallel_sequence_alignment_b200 as psa
ctx = psa.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
for n, L in ((1_000_000, 150), (64, 4_000_000)):
    A = torch.randint(0, 4, (n * L,), dtype=torch.uint8, device="cuda")
    B = torch.randint(0, 4, (n * L,), dtype=torch.uint8, device="cuda")
    off = (torch.arange(n, dtype=torch.int64, device="cuda") * L)
    ln = torch.full((n,), L, dtype=torch.int32, device="cuda")
    out = torch.zeros(n, dtype=torch.float64, device="cuda")
    go = lambda: ctx.similarity_batch_device(A.data_ptr(), off.data_ptr(), ln.data_ptr(), B.data_ptr(), off.data_ptr(),
                                             ln.data_ptr(), n, L, out.data_ptr(), stream.cuda_stream)
    for _ in range(3): go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10): go()
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    ref = (A.view(n, L) == B.view(n, L)).sum(1).double() / L
    print(f"{n} x {L}: {ms:.3f} ms, {2 * n * L / ms / 1e6:.0f} GB/s algorithmic, max |diff| {float((out - ref).abs().max()):.1e}")
