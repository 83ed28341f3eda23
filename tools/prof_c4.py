import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cse305_parallel_sequence_alignment_b200 as psa
from cse305_parallel_sequence_alignment_b200 import synth
L = int(os.environ.get("C4_LEN", "100000"))
ctx = psa.Context(0)
for _k, _v in __import__('json').loads(os.environ.get('OPTS', '{}')).items():   # psa_ctx options, e.g. OPTS='{"long_geometry": 6}'
    ctx.set_option(_k, _v)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
A, B = synth.mutated_pair(L, synth.SEED_C4)
dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
item = torch.zeros(10, dtype=torch.int32, device="cuda")
def run():
    ctx.align_long_device(dA.data_ptr(), dB.data_ptr(), L, L, item.data_ptr(), 0, 0, psa.LOCAL, 1, 2, False, stream.cuda_stream)
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); run(); e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"C4 {L}^2 opts={os.environ.get('OPTS','{}')}: {ms:.2f} ms, {L*L/ms/1e6:.1f} GCUPS score={int(item.cpu()[3])}")
