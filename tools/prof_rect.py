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
MODE = int(os.environ.get('MODE','1'))
for (m, n) in [(128, 100000), (256, 100000), (1024, 100000), (10000, 10000), (30000, 30000)]:
    A = synth.ACGT[rng.integers(0, 4, m, dtype=np.uint8)]; B = synth.ACGT[rng.integers(0, 4, n, dtype=np.uint8)]
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    item = torch.zeros(10, dtype=torch.int32, device="cuda")
    def run():
        ctx.align_long_device(dA.data_ptr(), dB.data_ptr(), m, n, item.data_ptr(), 0, 0, MODE, 1, 2, False, stream.cuda_stream)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); run(); e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    NB, S = (m + 127) // 128, (n + 255) // 256
    print(f"mode={MODE} m={m} n={n} NB={NB} S={S}: {ms:.3f} ms  per (NB+S-1) tile: {ms*1000/(NB+S-1):.1f} us  GCUPS {m*n/ms/1e6:.1f}")
