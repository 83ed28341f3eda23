"""Times the device-resident config-2 pass (1M pairs of 150x150, local + traceback); used with the
debug environment switches of the packed path to split fill and traceback time."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cse305_parallel_sequence_alignment_b200 as psa
from cse305_parallel_sequence_alignment_b200 import synth
N = int(os.environ.get("PAIRS", "1000000")); L = 150
ctx = psa.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
A, B = synth.read_pair_batch(N, L, synth.SEED_C2)
off, ln = synth.fixed_length_layout(N, L)
dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
dOff, dLen = torch.from_numpy(off).cuda(), torch.from_numpy(ln).cuda()
stride = (2 * L + 15) // 16 + 1
items = torch.zeros((N, 10), dtype=torch.int32, device="cuda")
ops = torch.zeros((N, stride), dtype=torch.int32, device="cuda")
tb = os.environ.get("TB", "1") == "1"
def go():
    ctx.align_batch_device(dA.data_ptr(), dOff.data_ptr(), dLen.data_ptr(), dB.data_ptr(), dOff.data_ptr(), dLen.data_ptr(),
                           N, L, L, items.data_ptr(), ops.data_ptr() if tb else 0, stride if tb else 0, psa.LOCAL, 1, 2, tb,
                           stream.cuda_stream)
for _ in range(3): go()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(10): go()
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"pairs {N} tb {tb} env {dict((k, v) for k, v in os.environ.items() if k.startswith('PSA_'))}: {ms:.3f} ms  {N * L * L / ms / 1e6:.0f} GCUPS")
