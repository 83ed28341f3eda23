"""Golden values for BASELINE config 4 (1 Mbp x 1 Mbp, seed 20250004): the linear-space oracle
(oracle/gotoh_oracle.c: orc_score_linear, pinned to the compiled reference on <= 20 kbp squares by
tests/test_oracle_golden.py) run ONCE on prefixes of the full-size pair -- SURVEY 8c "C4 specifics".
About 1e10 cells per 100 kbp prefix and mode: minutes of CPU, which is why the values are committed
instead of recomputed by the tests.

Usage:  python tests/golden/make_golden_c4.py        -> tests/golden/c4_prefix.json
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from cse305_parallel_sequence_alignment_b200 import synth  # noqa: E402

PREFIXES = [20_000, 100_000]
# rectangular prefixes exercise strip splits that do not line up with the row blocks
RECTS = [(30_000, 100_000), (100_000, 50_000)]


def main():
    po.build()
    A, B = synth.mutated_pair(1_000_000, synth.SEED_C4)
    out = {"seed": synth.SEED_C4, "length": 1_000_000, "g": 1, "h": 2, "cases": []}
    shapes = [(L, L) for L in PREFIXES] + RECTS
    for (m, n) in shapes:
        a, b = A[:m].tobytes(), B[:n].tobytes()
        for mode, name in ((po.LOCAL, "local"), (po.GLOBAL, "global")):
            t0 = time.time()
            r = po.score_linear(a, b, 1, 2, mode=mode)
            rec = {"m": m, "n": n, "mode": name, "score": int(r.score), "t1": int(r.t1), "t2": int(r.t2), "t3": int(r.t3),
                   "end_state": int(r.end_state), "end_i": int(r.end_i), "end_j": int(r.end_j),
                   "oracle_seconds": round(time.time() - t0, 1)}
            print(rec, flush=True)
            out["cases"].append(rec)
    json.dump(out, open(os.path.join(HERE, "c4_prefix.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
