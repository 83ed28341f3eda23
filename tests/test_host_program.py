"""The reference-shaped C++ program (host/testing): builds on CPU; on the GPU its stdout and CSV
match the reference program's golden output for BASELINE config 1."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

import cse305_parallel_sequence_alignment_b200 as psa
from tests.helpers import GOLDEN

HOST = os.path.join(os.path.dirname(psa.library_path()), "host")


def test_host_program_builds_and_links():
    psa.build_library()
    exe = os.path.join(HOST, "testing")
    assert os.path.exists(exe)
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libpsa.so" in out and "not found" not in out


def test_host_headers_keep_reference_signatures():
    h = open(os.path.join(HOST, "alignment_algorithm", "main_alignment.h")).read()
    assert "int main_alignment_function(char* A, char* B, size_t m, size_t n, size_t p, double g, double h);" in h
    t = open(os.path.join(HOST, "test_functions", "testing.h")).read()
    for fn in ("test_input_size", "test_input_size_thread", "test_n_cores", "test_n_cores_thread", "test_similarity"):
        assert f"int {fn}(std::vector<std::string>& names, std::vector<std::string>& sequences);" in t
    r = open(os.path.join(HOST, "test_functions", "read_test_data.h")).read()
    assert "int read_and_store_sequences(std::vector<std::string>& names, std::vector<std::string>& sequences, std::string& filename);" in r
    assert "double sequence_similarity(const std::string& sequence1, const std::string& sequence2);" in r


@pytest.mark.gpu
def test_config1_stdout_matches_reference_program(tmp_path):
    shutil.copy(os.path.join(GOLDEN, "dataset_head.fa"), tmp_path / "gene_sequences_test")
    run = subprocess.run([os.path.join(HOST, "testing")], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr
    want = open(os.path.join(GOLDEN, "g1_stdout.txt")).read().split("\n")
    got = run.stdout.split("\n")
    # the main thread's "Joining threads" interleaves freely with the worker's breadcrumbs in the
    # reference too ("stdout interleaving"): compare with that line factored out
    strip = lambda lines: [x for x in lines if x != "Joining threads"]
    assert strip(got) == strip(want)
    assert "Joining threads" in got
    csv = open(tmp_path / "input_size_testing.csv").read().split("\n")
    assert csv[:2] == open(os.path.join(GOLDEN, "g1_csv_head.txt")).read().split("\n")[:2]
    assert csv[2].startswith("0,50,")


def _rescore(row_a: str, row_b: str, g=1, h=2):
    """Score of an alignment as printed (print_seq): +1 per matching column, h + g*k per run of k gap columns."""
    score, prev = 0, 0
    for x, y in zip(row_a, row_b):
        t = 2 if x == "-" else (3 if y == "-" else 1)
        if t == 1:
            score += 1 if x == y else 0
        else:
            score -= g + (0 if t == prev else h)
        prev = t
    return score


def _check_harness_pairs(stdout: str, seqs, expect_pairs: int):
    """Every pair the harness aligned (full-length dataset records, concurrently from hardware_concurrency() host
    threads): the two printed rows spell suffixes of the two records, a truncated head is the border gap
    find_alignment drops (subproblem_alignment.cpp:170), and the rows re-score to the linear-space oracle's global
    score of that pair."""
    from oracle import pyoracle as po
    lines = stdout.split("\n")
    idx = [k for k, x in enumerate(lines) if x == "bp4"]
    assert len(idx) == expect_pairs
    checked = 0
    for k in idx:
        row_a, row_b = lines[k + 1], lines[k + 2]
        assert len(row_a) == len(row_b) and len(row_a) > 0
        sa, sb = row_a.replace("-", ""), row_b.replace("-", "")
        # the harness aligns equal-length prefixes of two records: find them by their tails
        cands = [(x, y) for x in range(len(seqs)) for y in range(len(seqs))
                 if seqs[x][:min(len(seqs[x]), len(seqs[y]))].endswith(sa) and seqs[y][:min(len(seqs[x]), len(seqs[y]))].endswith(sb)]
        assert cands, "printed rows are not suffixes of any pair of records"
        x, y = cands[0]
        L = min(len(seqs[x]), len(seqs[y]))
        a, b = seqs[x][:L].encode(), seqs[y][:L].encode()
        miss_a, miss_b = L - len(sa), L - len(sb)
        assert miss_a == 0 or miss_b == 0            # at most one border gap run is dropped
        lin = po.score_linear(a, b, 1, 2)
        printed = _rescore(row_a, row_b)
        dropped = (2 + miss_a) if miss_a else ((2 + miss_b) if miss_b else 0)
        # the dropped head is a gap run: either a fresh gap (h + g*k) or the extension of the first printed gap run
        first_t = 2 if row_a[0] == "-" else (3 if row_b[0] == "-" else 1)
        extends = (miss_a and first_t == 3) or (miss_b and first_t == 2)
        full = printed - (dropped - (2 if extends else 0))
        assert full == lin.score, (x, y, full, lin.score)
        checked += 1
    return checked


@pytest.mark.gpu
def test_cores_experiment_pairs_match_oracle(tmp_path):
    """f-1 'next' row: test_n_cores_thread (testing.cpp:209-287, disabled in the reference because its tables do not
    fit) through the GPU path -- 16 full-length (13.4 kbp) record pairs aligned from hardware_concurrency() host
    threads at once; every printed alignment re-scores to the oracle's global score."""
    from tests.helpers import dataset
    shutil.copy(os.path.join(GOLDEN, "dataset_head.fa"), tmp_path / "gene_sequences_test")
    env = dict(os.environ, PSA_EXPERIMENT="cores", PSA_TEST_PAIRS="16")
    run = subprocess.run([os.path.join(HOST, "testing")], cwd=tmp_path, capture_output=True, text=True, timeout=900, env=env)
    assert run.returncode == 0, run.stderr[-2000:]
    rows = open(tmp_path / "n_cores_testing.csv").read().strip().split("\n")
    assert rows[1] == "Test number,Number of cores,Execution time" and len(rows) == 2 + 16
    names, seqs = dataset()
    assert _check_harness_pairs(run.stdout, seqs, 16) == 16


@pytest.mark.gpu
def test_similarity_experiment_bounded(tmp_path):
    """f-1 'next' row: test_similarity (testing.cpp:295-369) through the GPU path: CSV schema, and both alignments
    re-score to the oracle."""
    from tests.helpers import dataset
    shutil.copy(os.path.join(GOLDEN, "dataset_head.fa"), tmp_path / "gene_sequences_test")
    env = dict(os.environ, PSA_EXPERIMENT="similarity", PSA_TEST_PAIRS="2")
    run = subprocess.run([os.path.join(HOST, "testing")], cwd=tmp_path, capture_output=True, text=True, timeout=600, env=env)
    assert run.returncode == 0, run.stderr
    rows = open(tmp_path / "similarity_testing.csv").read().strip().split("\n")
    assert rows[1] == "Test number,Similarity,Execution time" and len(rows) == 4
    names, seqs = dataset()
    assert _check_harness_pairs(run.stdout, seqs, 2) == 2


def _nodes_from_ops(ops: bytes, start_i: int, start_j: int):
    """The reference's node list (subproblem_alignment.cpp:151-165: (i,j) for t=1, (0,j) for t=2, (i,0) for t=3)."""
    out, i, j = [], start_i, start_j
    for k, t in enumerate(ops):
        if k > 0:
            if t != 2:
                i += 1
            if t != 3:
                j += 1
        out.append((0 if t == 2 else i, 0 if t == 3 else j, t))
    return out


@pytest.mark.gpu
def test_list_interface_mirrors(tmp_path):
    """compute_alignment / print_seq(align*) / optimal_alignment of host/alignment_algorithm/main_alignment.cpp (the
    mirrors of subproblem_alignment.h:8-13, :33-34 and main_alignment.cpp:32-55, :202-351) driven from C++: node
    lists equal the compiled reference's Subproblem nodes (oracle/_ref) or, without it, the oracle's path; the
    printed rows equal the reference-recorded fixture rows; a 3-piece partition prints the oracle's linked rows."""
    from oracle import pyoracle as po
    from tests.test_gpu_short import _oracle_partition
    exe = os.path.join(HOST, "tests", "align_list_driver")
    assert os.path.exists(exe)
    cases = [c for c in json.load(open(os.path.join(GOLDEN, "random_small.json"))) if c["g"] == int(c["g"])][:60]
    cmds = [f"pair {c['a']} {c['b']} {c['g']} {c['h']}" for c in cases]
    # a 3-piece partition cut along the oracle's own optimal path
    a, b = "GATTACAGATTACAGGATCCATTACA", "GATCACAGGATTAAGGTTCCATACA"
    w = po.align(a.encode(), b.encode(), 1, 2)
    nodes = _nodes_from_ops(w.ops, w.start_i, w.start_j)
    diag = [(i, j, t) for (i, j, t) in nodes if t == 1 and 0 < i < len(a) and 0 < j < len(b)]
    cut = [diag[len(diag) // 3], diag[2 * len(diag) // 3]]
    points = [(0, 0, -1)] + cut + [(len(a), len(b), 1)]
    cmds.append(f"part {a} {b} 1 2 {len(points)} " + " ".join(f"{i} {j} {t}" for i, j, t in points))
    run = subprocess.run([exe], input="\n".join(cmds) + "\n", capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr[-2000:]
    blocks = [blk.strip().split("\n") for blk in run.stdout.split("end\n") if blk.strip()]
    assert len(blocks) == len(cmds)
    for c, blk in zip(cases, blocks[:-1]):
        assert blk[0] == "corner " + " ".join(str(x) for x in c["corner"]), c
        got_nodes = [tuple(int(v) for v in x.split(",")) for x in blk[1].split()[1:]]
        if po.have_ref():
            _, _, ref_nodes = po.ref_subproblem(c["a"].encode(), c["b"].encode(), c["g"], c["h"])
            want_nodes = [tuple(int(v) for v in row) for row in ref_nodes]
        else:
            o = po.align(c["a"].encode(), c["b"].encode(), c["g"], c["h"])
            want_nodes = _nodes_from_ops(o.ops, o.start_i, o.start_j)
        assert got_nodes == want_nodes, c
        assert blk[2] == f"tail ok {len(want_nodes)}"
        rows = blk[3:5] if len(want_nodes) else ["", ""]
        assert [rows[0] if len(rows) > 0 else "", rows[1] if len(rows) > 1 else ""] == [c["row_a"], c["row_b"]], c
    ops, ra, rb, _ = _oracle_partition(a.encode(), b.encode(), points, 1, 2)
    assert blocks[-1][:2] == [ra.decode(), rb.decode()]


@pytest.mark.gpu
def test_reference_own_harness_on_libpsa(tmp_path):
    """Drop-in proof: the reference's OWN main.cpp + testing.cpp + pull_data.cpp (compiled in place
    from /root/reference by oracle/Makefile) linked against libpsa.so through the 20-line binding
    host/dropin/main_alignment_dropin.cpp prints what the reference program prints."""
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "testing_dropin")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/testing_dropin not built (needs /root/reference at build time)")
    shutil.copy(os.path.join(GOLDEN, "dataset_head.fa"), tmp_path / "gene_sequences_test")
    run = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr
    want = open(os.path.join(GOLDEN, "g1_stdout.txt")).read().split("\n")
    strip = lambda lines: [x for x in lines if x != "Joining threads"]
    assert strip(run.stdout.split("\n")) == strip(want)
    assert open(tmp_path / "input_size_testing.csv").read().split("\n")[2].startswith("0,50,")
