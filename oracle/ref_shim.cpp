// oracle/ref_shim.cpp -- extern "C" doorway into the UNMODIFIED reference classes.
//
// TEST INFRASTRUCTURE ONLY (see oracle/gotoh_oracle.c).  This file contains no alignment
// logic of its own: it is compiled by oracle/Makefile together with the reference's own
// alignment_algorithm/subproblem_alignment.cpp and the (2-line-repaired, piped through sed,
// never copied) alignment_algorithm/main_alignment.cpp, straight from /root/reference, into
// oracle/_ref/libref_align.so.  It lets Python drive
//   - class Subproblem  (subproblem_alignment.h:16-97): compute_tables() + find_alignment()
//   - main_alignment_function (main_alignment.h:38) with its stdout captured
// so that the restatement in gotoh_oracle.c can be pinned against the real thing.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <unistd.h>
#include <fcntl.h>

#include "subproblem_alignment.h"   // -I/root/reference/alignment_algorithm

int main_alignment_function(char* A, char* B, size_t m, size_t n, size_t p, double g, double h);

namespace {
const int32_t kNeg = INT32_MIN / 2;
int32_t cell(double v) { return std::isinf(v) ? kNeg : (int32_t)v; }
// the reference wants 1-indexed, unterminated buffers (test_functions/testing.cpp:124-128)
char* one_indexed(const char* s, int64_t len) {
    char* buf = (char*)malloc((size_t)len + 2);
    buf[0] = '#';
    memcpy(buf + 1, s, (size_t)len);
    buf[len + 1] = '#';
    return buf;
}
}  // namespace

extern "C" {

// Runs Subproblem(A,B,m,n,0,0,p,start,end,g,h).compute_tables(); find_alignment().
// corner[3] = T1/T2/T3[m][n]; nodes = (i,j,t) triples of the list alignment_begin..end in
// forward order, capacity node_cap triples; returns the node count or -1 if capacity is short.
// tables (optional) receives T1,T2,T3 as int32, (m+1)*(n+1) each, -inf -> INT32_MIN/2.
int64_t ref_subproblem(const char* a, const char* b, int64_t m, int64_t n, int64_t p, int start_type,
                       int end_type, double g, double h, int32_t* corner, int32_t* end_state,
                       int64_t* nodes, int64_t node_cap, int32_t* tables) {
    char* A = one_indexed(a, m);
    char* B = one_indexed(b, n);
    int64_t count = 0;
    {
        Subproblem sp(A, B, (size_t)m, (size_t)n, 0, 0, (size_t)p, start_type, end_type, g, h);
        sp.compute_tables();
        sp.find_alignment();
        corner[0] = cell(sp.T1[sp.m][sp.n]);
        corner[1] = cell(sp.T2[sp.m][sp.n]);
        corner[2] = cell(sp.T3[sp.m][sp.n]);
        *end_state = sp.alignment_end ? sp.alignment_end->t : 0;
        for (align* q = sp.alignment_begin; q != NULL; q = q->next) {
            if (count < node_cap) {
                nodes[3 * count + 0] = (int64_t)q->i;
                nodes[3 * count + 1] = (int64_t)q->j;
                nodes[3 * count + 2] = (int64_t)q->t;
            }
            ++count;
        }
        if (tables) {
            const size_t cells = (size_t)(sp.m + 1) * (sp.n + 1);
            for (size_t i = 0; i <= sp.m; ++i)
                for (size_t j = 0; j <= sp.n; ++j) {
                    tables[i * (sp.n + 1) + j] = cell(sp.T1[i][j]);
                    tables[cells + i * (sp.n + 1) + j] = cell(sp.T2[i][j]);
                    tables[2 * cells + i * (sp.n + 1) + j] = cell(sp.T3[i][j]);
                }
        }
        for (align* q = sp.alignment_begin; q != NULL;) { align* nx = q->next; free(q); q = nx; }
    }
    free(A); free(B);
    return count <= node_cap ? count : -1;
}

// main_alignment_function with fd 1 redirected into `out` (capacity cap).  Returns the number
// of bytes the reference printed (may exceed cap; only cap bytes are stored), or -1.
int64_t ref_main_alignment_capture(const char* a, const char* b, int64_t m, int64_t n, int64_t p, double g,
                                   double h, char* out, int64_t cap) {
    char* A = one_indexed(a, m);
    char* B = one_indexed(b, n);
    fflush(stdout);
    char path[] = "/tmp/ref_capture_XXXXXX";
    int tmp = mkstemp(path);
    if (tmp < 0) return -1;
    unlink(path);
    int saved = dup(1);
    dup2(tmp, 1);
    main_alignment_function(A, B, (size_t)m, (size_t)n, (size_t)p, g, h);
    fflush(stdout);
    dup2(saved, 1);
    close(saved);
    off_t len = lseek(tmp, 0, SEEK_END);
    lseek(tmp, 0, SEEK_SET);
    int64_t want = len < cap ? len : cap, got = 0;
    while (got < want) {
        ssize_t r = read(tmp, out + got, (size_t)(want - got));
        if (r <= 0) break;
        got += r;
    }
    close(tmp);
    free(A); free(B);
    return (int64_t)len;
}

// main_alignment_function with stdout discarded: the timed CPU-baseline entry.  The caller
// redirects fd 1 to /dev/null once (ref_silence_stdout) so that concurrent threads can call this.
int ref_main_alignment(const char* a, const char* b, int64_t m, int64_t n, int64_t p, double g, double h) {
    char* A = one_indexed(a, m);
    char* B = one_indexed(b, n);
    int rc = main_alignment_function(A, B, (size_t)m, (size_t)n, (size_t)p, g, h);
    free(A); free(B);
    return rc;
}

static int g_saved_stdout = -1;
void ref_silence_stdout(int on) {
    fflush(stdout);
    if (on && g_saved_stdout < 0) {
        g_saved_stdout = dup(1);
        int nul = open("/dev/null", O_WRONLY);
        dup2(nul, 1);
        close(nul);
    } else if (!on && g_saved_stdout >= 0) {
        dup2(g_saved_stdout, 1);
        close(g_saved_stdout);
        g_saved_stdout = -1;
    }
}

// Pair-parallel timing helper (shape of test_n_cores_thread, testing.cpp:269-276): n_threads
// host threads, contiguous chunks of pairs, each pair one main_alignment_function call with
// thread budget p.  Sequences are laid out back to back (offsets + lengths).  Returns seconds.
double ref_time_batch(const char* a, const int64_t* off_a, const int32_t* len_a, const char* b,
                      const int64_t* off_b, const int32_t* len_b, int64_t n_pairs, int64_t p, double g, double h,
                      int n_threads);
}  // extern "C"

#include <chrono>
#include <thread>
#include <vector>

extern "C" double ref_time_batch(const char* a, const int64_t* off_a, const int32_t* len_a, const char* b,
                                 const int64_t* off_b, const int32_t* len_b, int64_t n_pairs, int64_t p, double g,
                                 double h, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    ref_silence_stdout(1);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    const int64_t chunk = (n_pairs + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        const int64_t lo = t * chunk, hi = std::min<int64_t>(n_pairs, lo + chunk);
        if (lo >= hi) break;
        pool.emplace_back([=]() {
            for (int64_t q = lo; q < hi; ++q)
                ref_main_alignment(a + off_a[q], b + off_b[q], len_a[q], len_b[q], p, g, h);
        });
    }
    for (auto& th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    ref_silence_stdout(0);
    return std::chrono::duration<double>(t1 - t0).count();
}
