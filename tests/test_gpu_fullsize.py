"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle cannot run
there in seconds): two independent GPU implementations must agree on everything, results must be
symmetric under swapping the sequences, and sampled members are checked against the oracle."""
import os

import numpy as np
import pytest
import torch

import cse305_parallel_sequence_alignment_b200 as psa
from cse305_parallel_sequence_alignment_b200 import synth
from cse305_parallel_sequence_alignment_b200.capi import ITEM_DTYPE
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = psa.Context(0)
    yield c
    c.close()


def _device_batch(ctx, A, B, mode, tb, stream):
    n, L = A.shape
    off, ln = synth.fixed_length_layout(n, L)
    dA, dB = torch.from_numpy(A.reshape(-1)).cuda(), torch.from_numpy(B.reshape(-1)).cuda()
    dOff, dLen = torch.from_numpy(off).cuda(), torch.from_numpy(ln).cuda()
    items = torch.zeros(n * 10, dtype=torch.int32, device="cuda")
    stride = (2 * L + 15) // 16 + 1
    ops = torch.zeros(n * stride if tb else 1, dtype=torch.int32, device="cuda")
    ctx.align_batch_device(dA.data_ptr(), dOff.data_ptr(), dLen.data_ptr(), dB.data_ptr(), dOff.data_ptr(), dLen.data_ptr(),
                           n, L, L, items.data_ptr(), ops.data_ptr() if tb else 0, stride if tb else 0, mode, 1, 2, tb,
                           stream.cuda_stream)
    torch.cuda.synchronize()
    return items.cpu().numpy().view(ITEM_DTYPE), (ops.cpu().numpy().view(np.uint32).reshape(n, stride) if tb else None)


def test_config2_full_1M_packed_equals_generic(ctx):
    """All 1 000 000 pairs of config 2: the packed .S16x2 pipeline and the generic int32 kernel
    (different arithmetic, different direction codes, different traceback walkers) agree on every
    score, end cell, start cell, length and op word; a sample is checked against the oracle."""
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    n = 1_000_000
    A, B = synth.read_pair_batch(n, 150, synth.SEED_C2)
    it_p, ops_p = _device_batch(ctx, A, B, psa.LOCAL, True, stream)
    ctx.set_option("pack", 0)
    try:
        it_g, ops_g = _device_batch(ctx, A, B, psa.LOCAL, True, stream)
    finally:
        ctx.set_option("pack", 1)
    # the checkpoint + tile-recompute traceback (no code stream): a third, independent route to the same answers
    ctx.set_option("pack_traceback", 1)
    try:
        it_c, ops_c = _device_batch(ctx, A, B, psa.LOCAL, True, stream)
    finally:
        ctx.set_option("pack_traceback", 0)
    for f in ("score", "end_i", "end_j", "start_i", "start_j", "aln_len"):
        assert np.array_equal(it_p[f], it_g[f]), f
        assert np.array_equal(it_p[f], it_c[f]), f
    # compare only the words that carry ops (the tail of a pair's stride is unspecified)
    words = (it_p["aln_len"] + 15) // 16
    mask = np.arange(ops_p.shape[1])[None, :] < words[:, None]
    assert np.array_equal(ops_p[mask], ops_g[mask])
    assert np.array_equal(ops_p[mask], ops_c[mask])
    # invariants over the whole batch
    assert (it_p["score"] >= 0).all() and (it_p["score"] <= 150).all()
    assert ((it_p["aln_len"] == 0) == (it_p["score"] == 0)).all()
    assert (it_p["end_i"] - it_p["start_i"] + 1 <= it_p["aln_len"]).all()
    assert (it_p["score"][0::2].mean() > 110) and (it_p["score"][1::2].mean() < 70)    # mutated copies vs random (mismatch costs 0)
    for k in range(0, n, 49999):
        w = po.align(A[k].tobytes(), B[k].tobytes(), 1, 2, mode=po.LOCAL)
        assert (it_p[k]["score"], it_p[k]["end_i"], it_p[k]["end_j"], it_p[k]["start_i"], it_p[k]["start_j"]) == \
               (w.score, w.end_i, w.end_j, w.start_i, w.start_j)
        assert psa.unpack_ops(ops_p[k], int(it_p[k]["aln_len"])) == w.ops


def test_config5_shape_packed_equals_int32(ctx):
    """Config 5 pair size (5 kbp x 5 kbp), 1 536 pairs: packed strip kernel == int32 tile kernel."""
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    A, B = synth.read_pair_batch(1536, 5000, synth.SEED_C5)
    for mode in (psa.LOCAL, psa.GLOBAL):
        it_p, _ = _device_batch(ctx, A, B, mode, False, stream)
        ctx.set_option("pack", 0)
        try:
            it_g, _ = _device_batch(ctx, A, B, mode, False, stream)
        finally:
            ctx.set_option("pack", 1)
        for f in ("score", "end_i", "end_j", "t1", "t2", "t3", "end_state"):
            assert np.array_equal(it_p[f], it_g[f]), (mode, f)


def _c4_variants():
    """Every kernel variant that can run one long pair score-only: the row-block tile geometries
    (option long_geometry) and the column-stationary systolic kernel (option long_systolic)."""
    out = [("auto", {})]
    for geo in (0, 4, 6, 7):
        out.append((f"geometry{geo}", {"long_geometry": geo}))
    out.append(("systolic", {"long_systolic": 1}))
    return out


def _long_item(c, dA, dB, m, n, mode, stream):
    item = torch.zeros(10, dtype=torch.int32, device="cuda")
    c.align_long_device(dA.data_ptr(), dB.data_ptr(), m, n, item.data_ptr(), 0, 0, mode, 1, 2, False, stream.cuda_stream)
    torch.cuda.synchronize()
    return item.cpu().numpy().view(ITEM_DTYPE)[0]


def test_config4_prefixes_pinned_to_linear_oracle():
    """SURVEY 8c 'C4 specifics': prefixes of the seed-20250004 1 Mbp pair (100 kbp x 100 kbp and two
    rectangles) against the committed values of the linear-space oracle (tests/golden/c4_prefix.json,
    generated by tests/golden/make_golden_c4.py; the oracle itself is pinned to the compiled reference on
    <= 20 kbp squares) -- score, end cell (local) and corner values / end state (global), for EVERY
    kernel variant that can run the pair."""
    import json
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c4_prefix.json")))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    A, B = synth.mutated_pair(gold["length"], gold["seed"])
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    for name, opts in _c4_variants():
        c = psa.Context(0)
        for k, v in opts.items():
            c.set_option(k, v)
        for case in gold["cases"]:
            mode = psa.LOCAL if case["mode"] == "local" else psa.GLOBAL
            it = _long_item(c, dA, dB, case["m"], case["n"], mode, stream)
            if mode == psa.LOCAL:
                assert (int(it["score"]), int(it["end_i"]), int(it["end_j"])) == (case["score"], case["end_i"], case["end_j"]), (name, case)
            else:
                assert (int(it["t1"]), int(it["t2"]), int(it["t3"]), int(it["end_state"])) == \
                       (case["t1"], case["t2"], case["t3"], case["end_state"]), (name, case)
        c.close()


def test_config4_full_1Mbp_variants_agree(ctx):
    """Config 4 at full size (10^6 x 10^6, 10^12 cells): score AND end cell are identical between the
    default kernel, a different tile geometry and the systolic kernel (three different decompositions of the
    same matrix), invariant under swapping the two sequences, and consistent with the pinned 100 kbp prefix."""
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    L = 1_000_000
    A, B = synth.mutated_pair(L, synth.SEED_C4)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    ref = _long_item(ctx, dA, dB, L, L, psa.LOCAL, stream)
    key = (int(ref["score"]), int(ref["end_i"]), int(ref["end_j"]))
    assert key[0] > 800_000
    for name, opts in (("geometry6", {"long_geometry": 6}), ("systolic", {"long_systolic": 1})):
        c = psa.Context(0)
        for k, v in opts.items():
            c.set_option(k, v)
        it = _long_item(c, dA, dB, L, L, psa.LOCAL, stream)
        assert (int(it["score"]), int(it["end_i"]), int(it["end_j"])) == key, name
        c.close()
    swapped = _long_item(ctx, dB, dA, L, L, psa.LOCAL, stream)
    assert int(swapped["score"]) == key[0]
