"""CPU proofs of the algebraic rewrites the kernels use instead of the reference's literal recurrence
(subproblem_alignment.cpp:229-249, :396-398), each restated in plain Python / numpy on small pairs and compared
with the oracle's full tables (which tests/test_oracle_golden.py pins to the compiled reference):

  1. the one-instruction-per-recurrence form of the systolic and TF-form tile kernels (psa_systolic.cu sys_step,
     psa_tile.cuh): F = max(F' - g, H' - go), TF = max(F, T1) - go, E[j] = max(E[j-1] - g, TF[j-1]),
     H - go = max(E - go, TF)  -- E's own contribution to H can be dropped from its successor because h >= 0;
  2. local mode without a floor instruction: H = 0 on the borders and no max(0, .) in T1
     gives T1 identical to the spec everywhere and H = max(0, H_spec), hence the same score and end cell;
  3. the partition finder's combine rule (partial.cpp:101-108; psa_crossing_kernel): on EVERY row r the best
     crossing max_j max(Hf + Hr, T3f + T3r + h) of a forward and a reverse table equals the global optimum.
No GPU, no libpsa."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.helpers import mutated_copy, random_dna

NEG = -(1 << 29)              # the kernels' sentinel: only ever loses comparisons, never wraps


def _pairs(seed, count, lo=1, hi=48):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(count):
        m = int(rng.integers(lo, hi))
        n = int(rng.integers(lo, hi))
        a = random_dna(rng, m) if k % 3 else bytes(rng.choice(list(b"AC"), size=m).tolist())
        b = mutated_copy(rng, a, n) if k % 2 else (random_dna(rng, n) if k % 3 else bytes(rng.choice(list(b"AC"), size=n).tolist()))
        out.append((a, b))
    return out


def _finite(x):
    """Map everything that stands for -infinity (the oracle's sentinel, ours, and sentinel - penalties) to one value."""
    x = np.asarray(x, dtype=np.int64)
    return np.where(x < NEG // 2, NEG, x)


@pytest.mark.parametrize("g,h", [(1, 2), (2, 1), (1, 0), (0, 3), (3, 5)])
def test_one_instruction_recurrences_equal_the_literal_tables(g, h):
    go = g + h
    for a, b in _pairs(17 * g + h, 30):
        m, n = len(a), len(b)
        _, (T1, T2, T3) = po.align(a, b, g, h, want_tables=True)
        # row 0 / column 0 of start type -1 (subproblem_alignment.cpp:259-292)
        hg = [(-h - g * j if j else 0) - go for j in range(n + 1)]          # H[i-1][j] - go
        ff = [NEG] * (n + 1)                                                # F[i-1][j]
        for i in range(1, m + 1):
            h_left = -h - g * i                                             # column 0: T3[i][0]
            tf, e = h_left - go, NEG                                        # TF and E of column 0
            diag = hg[0]
            hg[0] = h_left - go
            for j in range(1, n + 1):
                t1g = diag + (1 if a[i - 1] == b[j - 1] else 0)             # T1 - go
                f = max(ff[j] - g, hg[j])                                   # viaddmax(ff, -g, hg)
                tfn = max(f - go, t1g)                                      # viaddmax(F, -go, t1g)
                e = max(e - g, tf)                                          # the only op on the row's chain
                tf = tfn
                hgo = max(e - go, tf)                                       # viaddmax(E, -go, TF)
                diag, hg[j], ff[j] = hg[j], hgo, f
                assert _finite(t1g + go) == _finite(T1[i, j]) and _finite(e) == _finite(T2[i, j]) and _finite(f) == _finite(T3[i, j]), (a, b, i, j)
                assert _finite(hgo + go) == _finite(max(T1[i, j], T2[i, j], T3[i, j]))


@pytest.mark.parametrize("g,h", [(1, 2), (2, 1), (1, 0), (2, 2)])
def test_local_mode_by_zero_borders_equals_the_floored_spec(g, h):
    go = g + h
    for a, b in _pairs(5 + 31 * g + h, 30):
        m, n = len(a), len(b)
        want, (T1, T2, T3) = po.align(a, b, g, h, mode=po.LOCAL, want_tables=True)
        Hs = np.maximum(T1, np.maximum(T2, T3)).astype(np.int64)
        H = np.zeros((m + 1, n + 1), dtype=np.int64)                        # H = 0 on both borders, E = F = -inf there
        E = np.full((m + 1, n + 1), NEG, dtype=np.int64)
        F = np.full((m + 1, n + 1), NEG, dtype=np.int64)
        K1 = np.full((m + 1, n + 1), NEG, dtype=np.int64)
        for i in range(1, m + 1):
            for j in range(1, n + 1):
                K1[i, j] = H[i - 1, j - 1] + (1 if a[i - 1] == b[j - 1] else 0)   # no max(0, .)
                E[i, j] = max(H[i, j - 1] - go, E[i, j - 1] - g)
                F[i, j] = max(H[i - 1, j] - go, F[i - 1, j] - g)
                H[i, j] = max(K1[i, j], E[i, j], F[i, j])
        inner = (slice(1, None), slice(1, None))
        assert np.array_equal(K1[inner], T1[inner])                          # T1 identical everywhere
        assert np.array_equal(H[inner], np.maximum(0, Hs[inner]))            # H = max(0, H_spec)
        pos2, pos3 = T2[inner] > 0, T3[inner] > 0                            # every state a path can visit is positive
        assert np.array_equal(E[inner][pos2], T2[inner][pos2]) and np.array_equal(F[inner][pos3], T3[inner][pos3])
        best = int(K1[inner].max()) if m and n else 0
        assert best == want.score
        if best > 0:
            ii, jj = np.nonzero(K1 == best)
            k = np.lexsort((jj, ii))[0]                                      # smallest i, then smallest j
            assert (int(ii[k]), int(jj[k])) == (want.end_i, want.end_j)


@pytest.mark.parametrize("g,h", [(1, 2), (2, 1), (1, 0), (1, 4)])
def test_forward_plus_reverse_crossing_equals_the_optimum_on_every_row(g, h):
    for a, b in _pairs(3 + 7 * g + h, 30, lo=2):
        m, n = len(a), len(b)
        want, (T1, T2, T3) = po.align(a, b, g, h, want_tables=True)
        _, (R1, R2, R3) = po.align(a[::-1], b[::-1], g, h, want_tables=True)
        Hf = _finite(np.maximum(T1, np.maximum(T2, T3)))
        Hr = _finite(np.maximum(R1, np.maximum(R2, R3)))
        Ff, Fr = _finite(T3), _finite(R3)
        for r in range(1, m):
            js = np.arange(n + 1)
            through_node = Hf[r, js] + Hr[m - r, n - js]                     # the path leaves row r by a fresh step
            inside_gap = Ff[r, js] + Fr[m - r, n - js] + h                   # a vertical gap runs across row r: its open cost was paid twice
            val = np.maximum(through_node, inside_gap)
            assert int(val.max()) == want.score, (a, b, r)
            assert int(val.max()) > NEG // 2
