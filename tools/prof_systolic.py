"""One systolic launch for ncu: M x N local score, options from OPTS (JSON)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cse305_parallel_sequence_alignment_b200 as psa
from cse305_parallel_sequence_alignment_b200 import synth
M, N = int(os.environ.get("M", "200000")), int(os.environ.get("N", "25600"))
ctx = psa.Context(0)
ctx.set_option("long_systolic", 1)
for k, v in json.loads(os.environ.get("OPTS", "{}")).items():
    ctx.set_option(k, v)
A, B = synth.mutated_pair(max(M, N), synth.SEED_C4)
dA, dB = torch.from_numpy(A[:M].copy()).cuda(), torch.from_numpy(B[:N].copy()).cuda()
item = torch.zeros(10, dtype=torch.int32, device="cuda")
for _ in range(2):
    ctx.align_long_device(dA.data_ptr(), dB.data_ptr(), M, N, item.data_ptr(), 0, 0, psa.LOCAL, 1, 2, False, 0)
torch.cuda.synchronize()
print(item.cpu()[3])
