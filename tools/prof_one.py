import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cse305_parallel_sequence_alignment_b200 as psa
from cse305_parallel_sequence_alignment_b200 import synth
ctx = psa.Context(0)
for _k, _v in __import__('json').loads(os.environ.get('OPTS', '{}')).items():   # psa_ctx options, e.g. OPTS='{"long_geometry": 6}'
    ctx.set_option(_k, _v)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
rng = np.random.default_rng(1)
m, n, MODE = int(os.environ.get("M", "256")), int(os.environ.get("N", "100000")), int(os.environ.get("MODE", "1"))
A = synth.ACGT[rng.integers(0, 4, m, dtype=np.uint8)]; B = synth.ACGT[rng.integers(0, 4, n, dtype=np.uint8)]
dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
item = torch.zeros(10, dtype=torch.int32, device="cuda")
for _ in range(2):
    ctx.align_long_device(dA.data_ptr(), dB.data_ptr(), m, n, item.data_ptr(), 0, 0, MODE, 1, 2, False, stream.cuda_stream)
torch.cuda.synchronize()
print("done", item.cpu()[:5])

