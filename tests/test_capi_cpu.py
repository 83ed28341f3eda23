"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol declared in
include/psa.h, and refuses to compute without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

import cse305_parallel_sequence_alignment_b200 as psa
from cse305_parallel_sequence_alignment_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(psa.library_path()):
        psa.build_library()
    return psa.load_library()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "psa.h")).read()
    declared = set(re.findall(r"\b(psa_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"psa_ctx"}
    assert declared, "no prototypes found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/psa.h but not exported by libpsa.so"
    assert declared == set(capi.EXPORTS)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(psa.PsaError) as e:
        psa.Context(0)
    assert e.value.code == -3


def test_pack_and_unpack_roundtrip():
    seqs = [b"ACGT", b"", b"GGA"]
    bases, off, ln = psa.pack_pairs(seqs)
    assert bases.tobytes() == b"ACGTGGA" and list(off) == [0, 4, 4] and list(ln) == [4, 0, 3]
    fwd = bytes([1, 1, 2, 3, 1] * 7)
    words = np.zeros(4, dtype=np.uint32)
    for k, code in enumerate(fwd[::-1]):
        words[k >> 4] |= np.uint32(code << (2 * (k & 15)))
    assert psa.unpack_ops(words, len(fwd)) == fwd


def test_render_rows_matches_print_seq(lib):
    # G2 (main_alignment.cpp:356-362): AGGA vs AGTGC -> AG-GA / AGTGC
    ra, rb = psa.render_rows(b"AGGA", b"AGTGC", bytes([1, 1, 2, 1, 1]), 1, 1)
    assert (ra, rb) == (b"AG-GA", b"AGTGC")


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/psa.h must compile as C99 on its own (no C++, no CUDA, no torch types)."""
    import subprocess
    src = tmp_path / "use_psa.c"
    src.write_text('#include "psa.h"\nint main(void) { psa_bp b; psa_result r; psa_batch_item it; (void)b; (void)r; (void)it; return 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-c",
                    str(src), "-o", str(tmp_path / "use_psa.o")], check=True)


def test_pack_bases_2bit_layout():
    """psa_pack_bases: 16 bases per word, base r in bits 2*(r%16), A=0 C=1 T=2 G=3; counts unrepresentable bytes; the
    vectorised numpy packer used by bench.py produces the same words."""
    seq = b"ACGTTGCAACGTACGTA" + b"GG"
    words, bad = psa.pack_bases(seq)
    assert bad == 0 and len(words) == 2
    code = {65: 0, 67: 1, 84: 2, 71: 3}
    for r, ch in enumerate(seq):
        assert (int(words[r // 16]) >> (2 * (r % 16))) & 3 == code[ch]
    assert psa.pack_bases(b"ACGNacgt")[1] == 5
    mat = np.frombuffer(seq, dtype=np.uint8).reshape(1, -1)
    assert np.array_equal(psa.pack_reads_2bit(mat)[0], words)


def test_compact_ops_offsets_are_the_running_word_counts():
    """PSA_OPS_COMPACT layout: pair k's ceil(aln_len/16) words start at the sum of the earlier pairs' word counts."""
    import numpy as np
    from cse305_parallel_sequence_alignment_b200.capi import PACKED_ITEM_DTYPE
    items = np.zeros(6, dtype=PACKED_ITEM_DTYPE)
    items["aln_len"] = [0, 1, 16, 17, 150, 300]
    off = psa.compact_ops_offsets(items)
    assert off.tolist() == [0, 0, 1, 2, 4, 14, 33]
    assert psa.OPS_COMPACT == 4


def test_pack_reads_batch_equals_pack_bases(lib):
    """psa_pack_reads (8 bases per 64-bit operation, host threads) == psa_pack_bases read by read == the numpy packer:
    every length around the 16-base word edge, strided rows, one thread and many, and the count of bytes that are not
    upper-case ACGT (the reads a caller must route to psa_align_batch instead)."""
    rng = np.random.default_rng(99)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    for L in (1, 7, 8, 15, 16, 17, 31, 32, 33, 150, 151, 512):
        for n in (1, 5, 9000):
            wide = letters[rng.integers(0, 4, size=(n, L + 5))]
            reads = wide[:, :L]                                   # row stride L + 5 bytes
            for threads in (1, 0, 3):
                got, bad = psa.pack_reads(reads, threads)
                assert bad == 0 and got.shape == (n, (L + 15) // 16)
                assert np.array_equal(got, psa.pack_reads_2bit(np.ascontiguousarray(reads)))
            for k in (0, n // 2, n - 1):
                assert np.array_equal(got[k], psa.pack_bases(reads[k].tobytes())[0])
    # unrepresentable bytes: lower case, N, NUL, 0xFF, and the look-alikes that share a 2-bit code with a base
    dirty = letters[rng.integers(0, 4, size=(9000, 150))]
    junk = np.frombuffer(b"acgtN\x00\xff@BEFUVWSD", dtype=np.uint8)
    where = rng.random(dirty.shape) < 0.01
    dirty[where] = junk[rng.integers(0, len(junk), size=int(where.sum()))]
    want_bad = sum(psa.pack_bases(dirty[k].tobytes())[1] for k in range(0, 9000, 50))
    got, bad = psa.pack_reads(dirty, 0)
    assert bad == int(np.count_nonzero(~np.isin(dirty, letters))) > 0
    assert sum(psa.pack_reads(dirty[k:k + 1], 1)[1] for k in range(0, 9000, 50)) == want_bad
    assert np.array_equal(got, psa.pack_reads_2bit(dirty))         # packed as (c >> 1) & 3 all the same
    for byte in range(256):                                        # exhaustively: exactly four byte values are clean
        row = np.full((1, 16), byte, dtype=np.uint8)
        assert psa.pack_reads(row, 1)[1] == (0 if byte in b"ACGT" else 16), byte
