// The whole reference-side binding (option A): this single translation unit
// replaces alignment_algorithm/main_alignment.cpp + subproblem_alignment.cpp + sequence_alignment/
// of the reference.  oracle/Makefile links it with the reference's OWN, unmodified main.cpp,
// test_functions/testing.cpp and pull_data.cpp into oracle/_ref/testing_dropin, and
// tests/test_host_program.py checks that program's stdout against the reference program's.
#include <cstddef>
#include <cstdio>

#include "psa.h"

int main_alignment_function(char* A, char* B, size_t m, size_t n, size_t /*p*/, double g, double h) {
    thread_local psa_ctx* ctx = nullptr;          // the harness calls from hardware_concurrency() threads
    if (!ctx && psa_ctx_create(0, &ctx) != PSA_OK) {
        std::fprintf(stderr, "libpsa: %s\n", psa_last_error(nullptr));
        return PSA_ERR_CUDA;
    }
    psa_result r;
    const int rc = psa_align_pair(ctx, A + 1, B + 1, m, n, PSA_GLOBAL, (int)g, (int)h,
                                  PSA_WANT_SCORE | PSA_WANT_TRACEBACK, &r);   // 1-indexed buffers
    if (rc != PSA_OK) {
        std::fprintf(stderr, "libpsa: %s\n", psa_last_error(ctx));
        return rc;
    }
    std::printf("bp1\nbp1.2\nbp2\nbp3\nbp4\n%s\n%s\n", r.row_a, r.row_b);   // main_alignment.cpp:12-21, :32-55
    std::fflush(stdout);
    psa_result_free(&r);
    return 0;
}
