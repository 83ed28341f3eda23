"""Small config-5 run for ncu: 2400 pairs of 5k x 5k local score."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cse305_parallel_sequence_alignment_b200 as psa
from cse305_parallel_sequence_alignment_b200 import synth
n, L = int(os.environ.get("C5_PAIRS", "4736")), int(os.environ.get("C5_LEN", "5000"))
ctx = psa.Context(0)
for _k, _v in __import__('json').loads(os.environ.get('OPTS', '{}')).items():   # psa_ctx options, e.g. OPTS='{"long_geometry": 6}'
    ctx.set_option(_k, _v)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
A, B = synth.read_pair_batch(n, L, synth.SEED_C5)
off, ln = synth.fixed_length_layout(n, L)
dA, dB = torch.from_numpy(A.reshape(-1)).cuda(), torch.from_numpy(B.reshape(-1)).cuda()
dOff, dLen = torch.from_numpy(off).cuda(), torch.from_numpy(ln).cuda()
items = torch.zeros(n * 10, dtype=torch.int32, device="cuda")
def run():
    ctx.align_batch_device(dA.data_ptr(), dOff.data_ptr(), dLen.data_ptr(), dB.data_ptr(), dOff.data_ptr(), dLen.data_ptr(),
                           n, L, L, items.data_ptr(), 0, 0, psa.LOCAL, 1, 2, False, stream.cuda_stream)
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); run(); e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"C5 {n} pairs x {L}^2: {ms:.2f} ms, {n*L*L/ms/1e6:.1f} GCUPS")
