"""CPU proof of the claim the GPU tracebacks rest on ("Traceback without the matrix"):
find_alignment's first-equality predecessor order (subproblem_alignment.cpp:147-169) is a function of a few bits
per SOURCE cell, so the kernels never need T1/T2/T3 again once a cell has left its code behind.

Both encodings are restated here in numpy from full tables -- the compiled reference's own Subproblem tables when
oracle/_ref is built, else the oracle's -- and walked exactly as the kernels walk them; the result must be the
reference's / the oracle's alignment:
  * the 4-bit code of the int32 kernels (psa_short.cu, psa_long.cu): d1, z2, e3;
  * the 5-bit code of the packed .S16x2 kernel (psa_pack.cu): min(H-T1,1), min(H-E,3), min(H-F,3) for h <= 2, its
    next-state table (pack_tb_lut), the linear form 11 H - max(T1,H-1) - 2 max(E,H-3) - 8 max(F,H-3) the fill
    computes it with, and the local-mode floor test by carrying the running score instead of reading T1.
No GPU, no libpsa: this pins the derivation, the -m gpu tests pin the kernels."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.helpers import mutated_copy, random_dna

CASES = [(1, 2), (2, 1), (1, 0), (0, 2), (1, 1), (3, 2)]


def _pairs(seed, count, alphabet):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(count):
        m = int(rng.integers(1, 60))
        n = int(rng.integers(m, m + 25))
        if alphabet == "dna":
            a = random_dna(rng, m)
            b = mutated_copy(rng, a, n) if k % 2 else random_dna(rng, n)
        else:                                                   # two letters: ties everywhere
            a = bytes(rng.choice(list(b"AC"), size=m).tolist())
            b = bytes(rng.choice(list(b"AC"), size=n).tolist())
        out.append((a, b))
    return out


def _tables(a, b, g, h, mode):
    """(T1, E, F) as int64 [m+1, n+1] plus the expected forward ops and end (state, i, j)."""
    want, t = po.align(a, b, g, h, mode=mode, want_tables=True)
    if mode == po.GLOBAL and po.have_ref():
        corner, end_state, nodes, tr = po.ref_subproblem(a, b, g, h, 1, want_tables=True)
        assert np.array_equal(tr, t)                           # oracle tables == the reference's own tables
        assert bytes(int(x) for x in nodes[:, 2]) == want.ops
        assert end_state == want.end_state
    return t.astype(np.int64), want


def _match(a, b):
    return (np.frombuffer(a, dtype=np.uint8)[:, None] == np.frombuffer(b, dtype=np.uint8)[None, :]).astype(np.int64)


def _walk(next_state, stop_at, state, i, j):
    """The kernels' walk: emit the state of (i, j), step to its source cell, look the next state up there."""
    ops = []
    while i > 0 and j > 0:
        ops.append(state)
        si, sj = (i if state == 2 else i - 1), (j if state == 3 else j - 1)
        if stop_at(state, i, j, si, sj):
            break
        if si == 0 or sj == 0:                                 # the border node find_alignment drops (cpp:170)
            break
        state = next_state(state, si, sj)
        i, j = si, sj
    return bytes(ops[::-1])


@pytest.mark.parametrize("alphabet", ["dna", "two"])
@pytest.mark.parametrize("g,h", CASES)
def test_4bit_codes_replay_find_alignment(alphabet, g, h):
    for a, b in _pairs(100 * g + h, 40, alphabet):
        for mode in (po.GLOBAL, po.LOCAL):
            (T1, E, F), want = _tables(a, b, g, h, mode)
            H = np.maximum(T1, np.maximum(E, F))
            d1 = np.where(T1 == H, 1, np.where(E >= F, 2, 3))
            z2 = np.where(d1 == 1, E + h > H, E + h >= H)
            e3 = F + h > H
            if mode == po.LOCAL:
                d1 = np.where(H <= 0, 0, d1)                   # the 0 floor (kernels keep H = 0 on the borders)

            def nxt(state, si, sj):
                if state == 1:
                    return int(d1[si, sj])
                if state == 2:
                    return (2 if z2[si, sj] else 1) if d1[si, sj] == 1 else (2 if z2[si, sj] else 3)
                return 3 if e3[si, sj] else int(d1[si, sj])

            def stop(state, i, j, si, sj):
                return mode == po.LOCAL and state == 1 and (si == 0 or sj == 0 or d1[si, sj] == 0)

            if mode == po.LOCAL and want.score == 0:
                assert want.ops == b""
                continue
            got = _walk(nxt, stop, want.end_state if mode == po.GLOBAL else 1, int(want.end_i), int(want.end_j))
            assert got == want.ops, (a, b, g, h, mode)


@pytest.mark.parametrize("alphabet", ["dna", "two"])
@pytest.mark.parametrize("g,h", [c for c in CASES if c[1] <= 2])
def test_5bit_packed_codes_lut_and_linear_form(alphabet, g, h):
    # pack_tb_lut (psa_pack.cu): next state for (current state, 5-bit code of the source cell)
    lut = np.zeros((3, 32), dtype=np.int64)
    for code in range(32):
        ma, mb, mc = code & 1, (code >> 1) & 3, (code >> 3) & 3
        d1 = 1 if ma == 0 else (2 if mb == 0 else 3)
        z2 = (mb < h) if d1 == 1 else (mb <= h)
        lut[0, code] = d1
        lut[1, code] = (2 if z2 else 1) if d1 == 1 else (2 if z2 else 3)
        lut[2, code] = 3 if mc < h else d1
    for a, b in _pairs(7 + 10 * g + h, 40, alphabet):
        f = _match(a, b)
        for mode in (po.GLOBAL, po.LOCAL):
            (T1, E, F), want = _tables(a, b, g, h, mode)
            if mode == po.LOCAL:                               # the packed kernel's tables: H = max(0, H_spec)
                H = np.maximum(0, np.maximum(T1, np.maximum(E, F)))
            else:
                H = np.maximum(T1, np.maximum(E, F))
            code = np.minimum(H - T1, 1) + 2 * np.minimum(H - E, 3) + 8 * np.minimum(H - F, 3)
            linear = 11 * H - np.maximum(T1, H - 1) - 2 * np.maximum(E, H - 3) - 8 * np.maximum(F, H - 3)
            assert np.array_equal(code[1:, 1:], linear[1:, 1:])   # the form the fill kernel evaluates (3 VIADDMNMX + 4 IMAD)
            assert code[1:, 1:].min() >= 0 and code[1:, 1:].max() < 32
            if mode == po.LOCAL and want.score == 0:
                continue
            run = {"v": int(want.score)}                       # local: the value of the current state, carried along

            def nxt(state, si, sj):
                nx = int(lut[state - 1, code[si, sj]])
                if mode == po.LOCAL:                           # value of the next state at the source cell
                    if state == 2:
                        run["v"] += g if nx == 2 else g + h
                    elif state == 3:
                        run["v"] += g if nx == 3 else g + h
                return nx

            def stop(state, i, j, si, sj):
                if mode != po.LOCAL or state != 1:
                    return False
                run["v"] -= int(f[i - 1, j - 1])               # H of the source cell = T1 - f
                return run["v"] == 0                           # T1[i][j] == f: the path starts here

            got = _walk(nxt, stop, want.end_state if mode == po.GLOBAL else 1, int(want.end_i), int(want.end_j))
            assert got == want.ops, (a, b, g, h, mode)
            if mode == po.LOCAL:
                assert run["v"] == 0                           # the carried score reaches the floor exactly at the start cell
