"""Thread-safety of the boundary under the reference's own calling pattern: hardware_concurrency() host
threads, each calling main_alignment_function on its own pairs (/root/reference/test_functions/testing.cpp:
145-152, 269-276, 352-358).  Here: many host threads, one psa_ctx each, every thread mixing single pairs of
very different sizes (different shared-memory footprints of the SAME kernel instantiation -- the round-1
race), packed batches with ragged maxima, long pairs with checkpointed traceback and score-only long pairs,
repeated; every result is compared with the oracle."""
import random
import threading

import numpy as np
import pytest

import cse305_parallel_sequence_alignment_b200 as psa
from oracle import pyoracle as po
from tests.helpers import mutated_copy, py_random_pair, random_dna

pytestmark = pytest.mark.gpu

N_THREADS = 32
REPS = 6


def _make_jobs(seed):
    """One thread's job list: (kind, payload, expected)."""
    rnd = random.Random(seed)
    rng = np.random.default_rng(seed)
    jobs = []
    # single pairs, global + local, sizes chosen so that the SAME short-kernel instantiation (K = ceil(n/32))
    # is launched with very different dynamic shared-memory sizes by different threads
    for max_m, max_n in ((10, 30), (120, 200), (500, 200), (1500, 30), (60, 250)):
        a, b = py_random_pair(rnd, max_m, max_n, rnd.choice([b"ACGT", b"ACGTN", b"AC"]))
        b = b[:256]
        mode = rnd.choice([psa.GLOBAL, psa.LOCAL])
        jobs.append(("pair", (a, b, mode), po.align(a, b, 1, 2, mode=mode)))
    # a packed batch (>= 64 pairs) whose maxima differ per thread
    mm, nn = rnd.choice([(40, 48), (96, 128), (150, 150), (300, 150), (200, 256)])
    As = [random_dna(rng, int(rng.integers(max(1, mm // 2), mm + 1))) for _ in range(80)]
    Bs = [mutated_copy(rng, x, int(rng.integers(max(1, nn // 2), nn + 1))) if k % 2 == 0 else
          random_dna(rng, int(rng.integers(max(1, nn // 2), nn + 1))) for k, x in enumerate(As)]
    mode = rnd.choice([psa.GLOBAL, psa.LOCAL])
    jobs.append(("batch", (As, Bs, mode), [po.align(x, y, 1, 2, mode=mode) for x, y in zip(As, Bs)]))
    # one long pair with checkpointed traceback and one score-only
    m, n = int(rng.integers(300, 900)), int(rng.integers(900, 1600))
    a = random_dna(rng, m)
    b = mutated_copy(rng, a, n)
    jobs.append(("pair", (a, b, psa.GLOBAL), po.align(a, b, 1, 2, mode=psa.GLOBAL)))
    jobs.append(("score", (a, b, psa.LOCAL), po.score_linear(a, b, 1, 2, mode=psa.LOCAL)))
    return jobs


def _run_job(ctx, job):
    kind, payload, want = job
    if kind == "pair":
        a, b, mode = payload
        got = ctx.align_pair(a, b, mode, 1, 2)
        assert got.score == want.score and got.ops == want.ops, (len(a), len(b), mode)
        assert (got.row_a, got.row_b) == (want.row_a, want.row_b)
        assert (got.end_i, got.end_j, got.start_i, got.start_j) == (want.end_i, want.end_j, want.start_i, want.start_j)
        if mode == psa.GLOBAL:
            assert (got.t1, got.t2, got.t3, got.end_state) == (want.t1, want.t2, want.t3, want.end_state)
    elif kind == "score":
        a, b, mode = payload
        got = ctx.align_pair(a, b, mode, 1, 2, traceback=False)
        assert (got.score, got.end_i, got.end_j) == (want.score, want.end_i, want.end_j)
    else:
        As, Bs, mode = payload
        ba, oa, la = psa.pack_pairs(As)
        bb, ob, lb = psa.pack_pairs(Bs)
        items, ops = ctx.align_batch(ba, oa, la, bb, ob, lb, mode, 1, 2, traceback=True)
        for k, w in enumerate(want):
            it = items[k]
            assert it["score"] == w.score and (it["end_i"], it["end_j"]) == (w.end_i, w.end_j), k
            assert psa.unpack_ops(ops[k], int(it["aln_len"])) == w.ops, k


def test_many_host_threads_mixed_shapes_repeated():
    all_jobs = [_make_jobs(1000 + t) for t in range(N_THREADS)]
    errors = []
    start = threading.Barrier(N_THREADS)

    def worker(t):
        try:
            ctx = psa.Context(0)
            start.wait()
            rnd = random.Random(t)
            for rep in range(REPS):
                order = list(range(len(all_jobs[t])))
                rnd.shuffle(order)                 # threads hit the kernels in different orders every repetition
                for k in order:
                    _run_job(ctx, all_jobs[t][k])
            ctx.close()
        except BaseException as e:  # noqa: BLE001 -- report every failure of every thread
            errors.append((t, repr(e)))
            try:
                start.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(N_THREADS)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors[:4]


def test_one_context_two_streams_are_chained():
    """Asynchronous device calls of ONE context on two different user streams share its scratch; the library
    orders them (psa_common.cuh: psa_stream_enter/leave), so both results are right."""
    import torch
    from cse305_parallel_sequence_alignment_b200 import synth
    from cse305_parallel_sequence_alignment_b200.capi import ITEM_DTYPE
    ctx = psa.Context(0)
    dev = torch.device("cuda", 0)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    n = 4096
    outs = []
    for seed, st in ((1, s1), (2, s2), (3, s1), (4, s2)):
        A, B = synth.read_pair_batch(n, 150, seed)
        off, ln = synth.fixed_length_layout(n, 150)
        with torch.cuda.stream(st):
            dA, dB = torch.from_numpy(A.reshape(-1).copy()).to(dev), torch.from_numpy(B.reshape(-1).copy()).to(dev)
            dOff, dLen = torch.from_numpy(off).to(dev), torch.from_numpy(ln).to(dev)
            items = torch.zeros(n * 10, dtype=torch.int32, device=dev)
            ops = torch.zeros(n * 20, dtype=torch.int32, device=dev)
            st.synchronize()
            ctx.align_batch_device(dA.data_ptr(), dOff.data_ptr(), dLen.data_ptr(), dB.data_ptr(), dOff.data_ptr(),
                                   dLen.data_ptr(), n, 150, 150, items.data_ptr(), ops.data_ptr(), 20, psa.LOCAL, 1, 2,
                                   True, st.cuda_stream)
        outs.append((A, B, items, ops, (dA, dB, dOff, dLen)))
    torch.cuda.synchronize()
    for A, B, items, ops, _ in outs:
        it = items.cpu().numpy().view(ITEM_DTYPE)
        ow = ops.cpu().numpy().view(np.uint32).reshape(n, 20)
        for k in range(0, n, 257):
            w = po.align(A[k].tobytes(), B[k].tobytes(), 1, 2, mode=po.LOCAL)
            assert (it[k]["score"], it[k]["end_i"], it[k]["end_j"]) == (w.score, w.end_i, w.end_j)
            assert psa.unpack_ops(ow[k], int(it[k]["aln_len"])) == w.ops
    ctx.close()
