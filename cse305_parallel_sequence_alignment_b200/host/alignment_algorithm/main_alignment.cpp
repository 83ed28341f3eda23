// Thin C++ boundary over libpsa.so: see main_alignment.h.  Replaces the bodies of
// main_alignment_function / optimal_alignment / OptimalAlignmentMapThread
// (/root/reference/alignment_algorithm/main_alignment.cpp:353-410, :202-351, :11-22) and, through
// the C-ABI, Subproblem::compute_tables / find_alignment.  The reference's processor budgeting
// (omega, ParallelPrefix, assign_processors, main_alignment.cpp:81-200) is CPU-thread bookkeeping
// with no effect on results and has no GPU counterpart.
#include "main_alignment.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>

#include "../../../include/psa.h"

namespace {

// One context per host thread: the harness calls the entry point from hardware_concurrency()
// threads at once (test_functions/testing.cpp:145-152); each gets its own stream and scratch.
struct ThreadCtx {
    psa_ctx* ctx = nullptr;
    int status = PSA_OK;
    ThreadCtx() {
        int dev = 0;
        if (const char* e = std::getenv("PSA_DEVICE")) dev = std::atoi(e);
        status = psa_ctx_create(dev, &ctx);
        if (status != PSA_OK) std::fprintf(stderr, "libpsa: %s\n", psa_last_error(nullptr));
    }
    ~ThreadCtx() { psa_ctx_destroy(ctx); }
};

psa_ctx* thread_ctx(int* status) {
    thread_local ThreadCtx tc;
    *status = tc.status;
    return tc.ctx;
}

bool integral_penalty(double v, int* out) {
    if (!(v >= 0.0) || v > 1e6 || std::floor(v) != v) return false;
    *out = (int)v;
    return true;
}

int run_pair(char* A, char* B, size_t m, size_t n, double g, double h, psa_result* res) {
    int gi = 0, hi = 0;
    if (A == nullptr || B == nullptr || !integral_penalty(g, &gi) || !integral_penalty(h, &hi)) return PSA_ERR_ARG;
    int status = PSA_OK;
    psa_ctx* ctx = thread_ctx(&status);
    if (status != PSA_OK) return status;
    // the reference's buffers are 1-indexed: base 1 lives at A[1]
    status = psa_align_pair(ctx, A + 1, B + 1, m, n, PSA_GLOBAL, gi, hi, PSA_WANT_SCORE | PSA_WANT_TRACEBACK, res);
    if (status != PSA_OK) std::fprintf(stderr, "libpsa: %s\n", psa_last_error(ctx));
    return status;
}

std::mutex g_stdout_lock;

}  // namespace

int main_alignment_function(char* A, char* B, size_t m, size_t n, size_t p, double g, double h) {
    (void)p;
    psa_result res;
    const int status = run_pair(A, B, m, n, g, h, &res);
    if (status != PSA_OK) return status;
    // same bytes as the reference: five breadcrumbs, then print_seq's two rows
    std::string out = "bp1\nbp1.2\nbp2\nbp3\nbp4\n";
    out.append(res.row_a, (size_t)res.aln_len);
    out.push_back('\n');
    out.append(res.row_b, (size_t)res.aln_len);
    out.push_back('\n');
    {
        std::lock_guard<std::mutex> hold(g_stdout_lock);
        std::fwrite(out.data(), 1, out.size(), stdout);
        std::fflush(stdout);
    }
    psa_result_free(&res);
    return 0;
}

int compute_alignment(char* A, char* B, size_t m, size_t n, double g, double h, align** begin, align** end,
                      int corner[3]) {
    psa_result res;
    const int status = run_pair(A, B, m, n, g, h, &res);
    if (status != PSA_OK) return status;
    if (corner) { corner[0] = res.t1; corner[1] = res.t2; corner[2] = res.t3; }
    align* head = nullptr;
    align* tail = nullptr;
    size_t i = (size_t)res.start_i, j = (size_t)res.start_j;
    for (int64_t k = 0; k < res.aln_len; ++k) {
        const int t = res.ops[k];
        if (k > 0) { if (t != 2) ++i; if (t != 3) ++j; }
        align* node = (align*)std::malloc(sizeof(align));
        node->t = t;
        node->i = (t == 2) ? 0 : i;      // coordinate conventions of find_alignment (cpp:151-165)
        node->j = (t == 3) ? 0 : j;
        node->next = nullptr;
        if (tail) tail->next = node; else head = node;
        tail = node;
    }
    if (begin) *begin = head;
    if (end) *end = tail;
    psa_result_free(&res);
    return 0;
}

int optimal_alignment(char* A, char* B, std::vector<align> partial_bp, size_t m, size_t n, size_t p, double g, double h) {
    (void)p;
    int gi = 0, hi = 0;
    if (A == nullptr || B == nullptr || !integral_penalty(g, &gi) || !integral_penalty(h, &hi)) return PSA_ERR_ARG;
    int status = PSA_OK;
    psa_ctx* ctx = thread_ctx(&status);
    if (status != PSA_OK) return status;
    std::vector<psa_bp> bp(partial_bp.size());
    for (size_t k = 0; k < bp.size(); ++k) bp[k] = psa_bp{(int64_t)partial_bp[k].i, (int64_t)partial_bp[k].j, partial_bp[k].t, 0};
    psa_result res;
    status = psa_align_partition(ctx, A + 1, B + 1, m, n, bp.data(), bp.size(), gi, hi, &res);
    if (status != PSA_OK) { std::fprintf(stderr, "libpsa: %s\n", psa_last_error(ctx)); return status; }
    std::string out(res.row_a, (size_t)res.aln_len);
    out.push_back('\n');
    out.append(res.row_b, (size_t)res.aln_len);
    out.push_back('\n');
    {
        std::lock_guard<std::mutex> hold(g_stdout_lock);
        std::fwrite(out.data(), 1, out.size(), stdout);
        std::fflush(stdout);
    }
    psa_result_free(&res);
    return 0;
}

void free_alignment(align* begin) {
    while (begin) { align* nx = begin->next; std::free(begin); begin = nx; }
}

void print_align(align* begin) {
    for (; begin != nullptr; begin = begin->next) std::printf("(%ld, %ld, %d)\n", (long)begin->i, (long)begin->j, begin->t);
}

void print_seq(char* A, char* B, align* begin) {
    std::string top, bottom;
    for (align* q = begin; q != nullptr; q = q->next) {
        top.push_back((q->t == 1 || q->t == 3) ? A[q->i] : '-');
        bottom.push_back((q->t == 1 || q->t == 2) ? B[q->j] : '-');
    }
    std::printf("%s\n%s\n", top.c_str(), bottom.c_str());
}
