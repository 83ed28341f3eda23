"""CPU checks of the synthetic workloads and of the fixture that pins config 4 at scale: the generator is
deterministic (the committed goldens were produced from exactly these bytes, and the GPU box regenerates them instead
of shipping 2 MB of sequence), has the stated shape, and the linear-space oracle reproduces the committed 20 kbp
prefix values of tests/golden/c4_prefix.json from them."""
import hashlib
import json
import os

import numpy as np

from cse305_parallel_sequence_alignment_b200 import synth
from oracle import pyoracle as po
from tests.helpers import GOLDEN


def _md5(*arrays):
    return hashlib.md5(b"".join(np.ascontiguousarray(x).tobytes() for x in arrays)).hexdigest()


def test_generators_are_pinned():
    assert _md5(*synth.read_pair_batch(1000, 150, synth.SEED_C2)) == "8e981d1dd51ae21c5c89c1deddf6f146"
    assert _md5(*synth.mutated_pair(10_000, synth.SEED_C3)) == "7d131c645dc33aa916dacde16ed7aba0"
    assert _md5(*synth.mutated_pair(1_000_000, synth.SEED_C4)) == "7835ebd506bdcdbc10836e382f44127c"
    assert _md5(*synth.read_pair_batch(64, 5000, synth.SEED_C5)) == "82034952dc8a84db677e9f3748ec760a"


def test_shapes_and_mutation_rates():
    A, B = synth.read_pair_batch(2000, 150, 11)
    assert A.shape == B.shape == (2000, 150) and A.dtype == np.uint8
    assert set(np.unique(A)) <= set(b"ACGT") and set(np.unique(B)) <= set(b"ACGT")
    assert 0.22 < (A[1::2] == B[1::2]).mean() < 0.28          # independent reads agree by chance
    off, ln = synth.fixed_length_layout(200, 150)
    res = po.score_batch(np.ascontiguousarray(A[:200].reshape(-1)), off, ln, np.ascontiguousarray(B[:200].reshape(-1)), off, ln, 1, 2, po.LOCAL)
    sc = np.array([r.score for r in res])
    assert sc[0::2].mean() > 115 and sc[0::2].min() > 90     # mutated copies align over (nearly) their whole length
    assert sc[1::2].max() < 80 and sc[1::2].mean() < 60      # unrelated reads: a mismatch costs 0, so chance matches add up to ~50
    # a long mutated copy aligns with ~5 % mismatches and ~2 % gaps: its global score stays near 0.9 per base
    a, b = synth.mutated_pair(3000, 5)
    r = po.score_linear(a.tobytes(), b.tobytes(), 1, 2, mode=po.GLOBAL)
    assert 0.80 * 3000 < r.score < 0.97 * 3000
    off, ln = synth.fixed_length_layout(5, 150)
    assert off.tolist() == [0, 150, 300, 450, 600] and ln.tolist() == [150] * 5 and off.dtype == np.int64 and ln.dtype == np.int32


def test_c4_prefix_golden_reproduced_by_the_oracle():
    gold = json.load(open(os.path.join(GOLDEN, "c4_prefix.json")))
    assert gold["seed"] == synth.SEED_C4 and gold["length"] == 1_000_000
    A, B = synth.mutated_pair(gold["length"], gold["seed"])
    done = 0
    for case in gold["cases"]:
        if case["m"] * case["n"] > 4e8:          # the 100 kbp cases took a CPU-minute each once; the 20 kbp ones run here
            continue
        mode = po.LOCAL if case["mode"] == "local" else po.GLOBAL
        r = po.score_linear(A[:case["m"]].tobytes(), B[:case["n"]].tobytes(), gold["g"], gold["h"], mode=mode)
        assert (r.score, r.end_i, r.end_j) == (case["score"], case["end_i"], case["end_j"]), case
        if mode == po.GLOBAL:
            assert (r.t1, r.t2, r.t3, r.end_state) == (case["t1"], case["t2"], case["t3"], case["end_state"]), case
        done += 1
    assert done >= 2
