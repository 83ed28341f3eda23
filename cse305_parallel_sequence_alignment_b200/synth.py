"""Synthetic DNA workloads of the shapes BASELINE.json names: uniform random
ACGT reads and "mutated copies" (per base 5 % substitution to a different base, 1 % insertion of
a geometric(1/2)-length random run, 1 % deletion of a geometric(1/2)-length run, then padded with
random bases / truncated to the target length).  Vectorised numpy, fixed seeds."""
from __future__ import annotations

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)

SEED_C2 = 20250002
SEED_C3 = 20250003
SEED_C4 = 20250004
SEED_C5 = 20250005


def random_codes(rng: np.random.Generator, n_pairs: int, length: int) -> np.ndarray:
    return rng.integers(0, 4, size=(n_pairs, length), dtype=np.uint8)


def mutate_codes(rng: np.random.Generator, a: np.ndarray, out_len: int, sub=0.05, ins=0.01, dele=0.01) -> np.ndarray:
    """Row-wise mutated copies of the 2-bit code matrix a[n, L] -> [n, out_len]."""
    n, L = a.shape
    src = a.copy()
    smask = rng.random((n, L)) < sub
    src[smask] = (src[smask] + rng.integers(1, 4, size=int(smask.sum()), dtype=np.uint8)) & 3
    # deletions: a run of geometric(1/2) length starting at a base with probability `dele`
    keep = np.ones((n, L), dtype=bool)
    dstart = rng.random((n, L)) < dele
    dlen = np.where(dstart, rng.geometric(0.5, size=(n, L)), 0)
    for d in range(0, 12):
        hit = dlen > d
        if not hit.any():
            break
        keep[:, d:] &= ~hit[:, :L - d] if d else ~hit
    # insertions: geometric(1/2) random bases in front of a base with probability `ins`
    ilen = np.where(rng.random((n, L)) < ins, rng.geometric(0.5, size=(n, L)), 0)
    width = ilen + keep
    pos = np.cumsum(width, axis=1) - 1          # output position of base i when it is kept
    out = rng.integers(0, 4, size=(n, out_len), dtype=np.uint8)   # insertions and padding are random
    ok = keep & (pos < out_len)
    rows = np.broadcast_to(np.arange(n)[:, None], (n, L))
    out[rows[ok], pos[ok]] = src[ok]
    return out


def read_pair_batch(n_pairs: int, length: int = 150, seed: int = SEED_C2, letters: bool = True):
    """BASELINE config 2/5 shape: n_pairs pairs of length x length; even pair index: B is a
    mutated copy of A, odd: independent random.  Returns (A, B) as uint8 [n_pairs, length]
    matrices of ASCII letters (or 0..3 codes)."""
    rng = np.random.default_rng(seed)
    a = random_codes(rng, n_pairs, length)
    b = random_codes(rng, n_pairs, length)
    b[0::2] = mutate_codes(rng, a[0::2], length)
    if letters:
        return ACGT[a], ACGT[b]
    return a, b


def mutated_pair(length: int, seed: int, letters: bool = True):
    """One long pair (configs 3/4): B = mutated copy of A, both `length` long."""
    rng = np.random.default_rng(seed)
    a = random_codes(rng, 1, length)
    b = mutate_codes(rng, a, length)
    if letters:
        return ACGT[a[0]], ACGT[b[0]]
    return a[0], b[0]


def fixed_length_layout(n_pairs: int, length: int):
    """offsets/lengths of a [n_pairs, length] matrix viewed as back-to-back sequences."""
    off = np.arange(n_pairs, dtype=np.int64) * length
    ln = np.full(n_pairs, length, dtype=np.int32)
    return off, ln
