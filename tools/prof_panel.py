import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cse305_parallel_sequence_alignment_b200 as psa
from cse305_parallel_sequence_alignment_b200 import synth
ctx = psa.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
rng = np.random.default_rng(1)
MODE = int(os.environ.get("MODE", "0"))
shapes = [(100000, 128), (100000, 256), (100000, 1024), (100000, 8192), (20000, 128 * 512)]
for (m, n) in shapes:
    A = synth.ACGT[rng.integers(0, 4, m, dtype=np.uint8)]; B = synth.ACGT[rng.integers(0, 4, n, dtype=np.uint8)]
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    item = torch.zeros(10, dtype=torch.int32, device="cuda")
    def run():
        ctx.align_long_device(dA.data_ptr(), dB.data_ptr(), m, n, item.data_ptr(), 0, 0, MODE, 1, 2, False, stream.cuda_stream)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); run(); e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    strips = (n + 127) // 128
    print(f"mode={MODE} m={m} n={n} strips={strips}: {ms:.3f} ms; ns per (m + 64*strips) step: {ms*1e6/(m + 64*strips):.0f}; GCUPS {m*n/ms/1e6:.1f}", flush=True)
