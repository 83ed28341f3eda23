"""Config 3: one 10 kbp x 10 kbp mutated pair, global alignment with checkpointed traceback."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cse305_parallel_sequence_alignment_b200 as psa
from cse305_parallel_sequence_alignment_b200 import synth
L = int(os.environ.get("C3_LEN", "10000"))
MODE = int(os.environ.get("MODE", "0"))
ctx = psa.Context(0)
for _k, _v in __import__('json').loads(os.environ.get('OPTS', '{}')).items():   # psa_ctx options, e.g. OPTS='{"long_geometry": 6}'
    ctx.set_option(_k, _v)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
A, B = synth.mutated_pair(L, synth.SEED_C3)
dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
item = torch.zeros(10, dtype=torch.int32, device="cuda")
words = (2 * L + 15) // 16 + 1
ops = torch.zeros(words, dtype=torch.int32, device="cuda")
def run(tb):
    ctx.align_long_device(dA.data_ptr(), dB.data_ptr(), L, L, item.data_ptr(), ops.data_ptr() if tb else 0, words if tb else 0,
                          MODE, 1, 2, tb, stream.cuda_stream)
for tb in (True, False):
    for _ in range(2): run(tb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5): run(tb)
    e1.record(stream); torch.cuda.synchronize()
    it = item.cpu().numpy()
    print(f"C3 {L}^2 mode={MODE} traceback={tb} opts={os.environ.get('OPTS','{}')}: "
          f"{e0.elapsed_time(e1) / 5:.3f} ms  score={it[3]} aln_len={it[9]}")
