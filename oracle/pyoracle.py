"""ctypes doorway to the CPU checker (oracle/liboracle.so) and, when present, to the compiled
reference (oracle/_ref/libref_align.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NEG = -(2 ** 31) // 2  # ORC_NEG: stands for -infinity

GLOBAL, LOCAL = 0, 1


class OrcResult(C.Structure):
    _fields_ = [("t1", C.c_int32), ("t2", C.c_int32), ("t3", C.c_int32), ("score", C.c_int32),
                ("end_state", C.c_int32), ("end_i", C.c_int64), ("end_j", C.c_int64),
                ("start_i", C.c_int64), ("start_j", C.c_int64), ("aln_len", C.c_int64)]


@dataclass
class Alignment:
    t1: int
    t2: int
    t3: int
    score: int
    end_state: int
    end_i: int
    end_j: int
    start_i: int
    start_j: int
    ops: bytes          # forward order, values 1/2/3
    row_a: bytes
    row_b: bytes


def build(force: bool = False) -> None:
    """Compile liboracle.so (always) and oracle/_ref (only where /root/reference exists)."""
    if force or not os.path.exists(os.path.join(HERE, "liboracle.so")) or \
            os.path.getmtime(os.path.join(HERE, "liboracle.so")) < os.path.getmtime(os.path.join(HERE, "gotoh_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if os.path.isdir("/root/reference/alignment_algorithm"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_lib = None
_ref = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(os.path.join(HERE, "liboracle.so"))
        _lib.orc_align_full.restype = C.c_int
        _lib.orc_align_full.argtypes = [C.c_char_p, C.c_char_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.POINTER(OrcResult), C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]
        _lib.orc_score_linear.restype = C.c_int
        _lib.orc_score_linear.argtypes = [C.c_char_p, C.c_char_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(OrcResult)]
        _lib.orc_score_batch.restype = C.c_int
        _lib.orc_score_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return _lib


def have_ref() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref_align.so"))


def ref() -> C.CDLL:
    global _ref
    if _ref is None:
        _ref = C.CDLL(os.path.join(HERE, "_ref", "libref_align.so"))
        _ref.ref_subproblem.restype = C.c_int64
        _ref.ref_subproblem.argtypes = [C.c_char_p, C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                        C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_void_p]
        _ref.ref_main_alignment_capture.restype = C.c_int64
        _ref.ref_main_alignment_capture.argtypes = [C.c_char_p, C.c_char_p, C.c_int64, C.c_int64, C.c_int64,
                                                    C.c_double, C.c_double, C.c_char_p, C.c_int64]
        _ref.ref_time_batch.restype = C.c_double
        _ref.ref_time_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_int]
    return _ref


def align(a: bytes, b: bytes, g: int = 1, h: int = 2, mode: int = GLOBAL, start_type: int = -1,
          end_type: int = -1, want_tables: bool = False):
    """Full-matrix oracle: fill + traceback + rows.  Returns Alignment (and the 3 tables)."""
    m, n = len(a), len(b)
    res = OrcResult()
    ops = np.zeros(m + n + 1, dtype=np.uint8)
    ra = np.zeros(m + n + 1, dtype=np.uint8)
    rb = np.zeros(m + n + 1, dtype=np.uint8)
    tables = np.zeros(3 * (m + 1) * (n + 1), dtype=np.int32) if want_tables else None
    rc = lib().orc_align_full(a, b, m, n, g, h, start_type, end_type, mode, C.byref(res), ops.ctypes.data,
                              ra.ctypes.data, rb.ctypes.data, tables.ctypes.data if want_tables else None)
    if rc != 0:
        raise RuntimeError(f"orc_align_full rc={rc}")
    L = res.aln_len
    out = Alignment(res.t1, res.t2, res.t3, res.score, res.end_state, res.end_i, res.end_j, res.start_i,
                    res.start_j, ops[:L].tobytes(), ra[:L].tobytes(), rb[:L].tobytes())
    if want_tables:
        return out, tables.reshape(3, m + 1, n + 1)
    return out


def score_linear(a: bytes, b: bytes, g: int = 1, h: int = 2, mode: int = GLOBAL) -> OrcResult:
    res = OrcResult()
    rc = lib().orc_score_linear(a, b, len(a), len(b), g, h, mode, C.byref(res))
    if rc != 0:
        raise RuntimeError(f"orc_score_linear rc={rc}")
    return res


def score_batch(a: np.ndarray, off_a: np.ndarray, len_a: np.ndarray, b: np.ndarray, off_b: np.ndarray,
                len_b: np.ndarray, g: int = 1, h: int = 2, mode: int = GLOBAL):
    """Linear-space scores for a packed batch (uint8 bases back to back)."""
    n = len(len_a)
    out = (OrcResult * n)()
    rc = lib().orc_score_batch(a.ctypes.data, off_a.ctypes.data, len_a.ctypes.data, b.ctypes.data,
                               off_b.ctypes.data, len_b.ctypes.data, n, g, h, mode, C.addressof(out))
    if rc != 0:
        raise RuntimeError(f"orc_score_batch rc={rc}")
    return out


# ---- the compiled reference --------------------------------------------------------------

def ref_subproblem(a: bytes, b: bytes, g: float = 1, h: float = 2, p: int = 1, start_type: int = -1,
                   end_type: int = -1, want_tables: bool = False):
    """Drive the reference's Subproblem class.  Returns (corner(3), end_state, nodes[(i,j,t)])."""
    m, n = len(a), len(b)
    corner = np.zeros(3, dtype=np.int32)
    end_state = C.c_int32(0)
    cap = m + n + 2
    nodes = np.zeros(3 * cap, dtype=np.int64)
    tables = np.zeros(3 * (m + 1) * (n + 1), dtype=np.int32) if want_tables else None
    cnt = ref().ref_subproblem(a, b, m, n, p, start_type, end_type, float(g), float(h), corner.ctypes.data,
                               C.addressof(end_state), nodes.ctypes.data, cap,
                               tables.ctypes.data if want_tables else None)
    if cnt < 0:
        raise RuntimeError("ref_subproblem: node capacity")
    nodes = nodes[:3 * cnt].reshape(cnt, 3)
    if want_tables:
        return corner, end_state.value, nodes, tables.reshape(3, m + 1, n + 1)
    return corner, end_state.value, nodes


def ref_rows(a: bytes, b: bytes, nodes: np.ndarray):
    """print_seq (main_alignment.cpp:32-55) applied to a node list, for comparison."""
    ra = bytes(a[i - 1] if t in (1, 3) else 0x2D for i, j, t in nodes)
    rb = bytes(b[j - 1] if t in (1, 2) else 0x2D for i, j, t in nodes)
    return ra, rb


def ref_main_alignment_stdout(a: bytes, b: bytes, p: int = 32, g: float = 1, h: float = 2) -> bytes:
    cap = 4 * (len(a) + len(b)) + 256
    buf = C.create_string_buffer(cap)
    n = ref().ref_main_alignment_capture(a, b, len(a), len(b), p, float(g), float(h), buf, cap)
    if n < 0 or n > cap:
        raise RuntimeError("ref_main_alignment_capture failed")
    return buf.raw[:n]


def ref_time_batch(a: np.ndarray, off_a, len_a, b: np.ndarray, off_b, len_b, p: int, g: float, h: float,
                   n_threads: int) -> float:
    return ref().ref_time_batch(a.ctypes.data, off_a.ctypes.data, len_a.ctypes.data, b.ctypes.data,
                                off_b.ctypes.data, len_b.ctypes.data, len(len_a), p, float(g), float(h), n_threads)
