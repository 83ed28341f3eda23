// Host-side mirror of the reference's alignment driver interface
// (/root/reference/alignment_algorithm/main_alignment.h:38 and subproblem_alignment.h:8-13).
// Same entry point, same argument meaning, same stdout; the body is a thin call into the C-ABI
// of libpsa.so (include/psa.h) -- the DP fill and traceback run on the GPU, never on the CPU.
#pragma once
#ifndef PSA_HOST_MAIN_ALIGNMENT_H
#define PSA_HOST_MAIN_ALIGNMENT_H

#include <cstddef>
#include <vector>

// Path node, layout-compatible with the reference's `align` (subproblem_alignment.h:8-13):
// t = 1 diagonal (i,j), t = 2 gap in A (0,j), t = 3 gap in B (i,0).
typedef struct alignment_point {
    size_t i;
    size_t j;
    int t;
    struct alignment_point* next = nullptr;
} align;

// Global affine-gap alignment of A[1..m] with B[1..n] (1-indexed, unterminated buffers exactly as
// test_functions/testing.cpp:124-128 builds them); match +1, mismatch 0, gap of length k costs
// h + g*k.  Prints "bp1\nbp1.2\nbp2\nbp3\nbp4\n" and the two alignment rows
// (main_alignment.cpp:12-21, :32-55).  p (the reference's thread budget) is accepted and ignored:
// results never depended on it.  Returns 0, or a negative psa_status (the reference never fails;
// here a missing GPU or non-integral g/h is an error, not a CPU fallback).
int main_alignment_function(char* A, char* B, size_t m, size_t n, size_t p, double g, double h);

// The same computation returning the path as the reference's linked list instead of printing it
// (what Subproblem::alignment_begin/alignment_end exposed, subproblem_alignment.h:33-34).  Nodes
// are malloc'ed; free with free_alignment().  corner[3] receives T1/T2/T3[m][n].
int compute_alignment(char* A, char* B, size_t m, size_t n, double g, double h, align** begin, align** end,
                      int corner[3]);
// optimal_alignment (main_alignment.cpp:202-351): solves the pieces between consecutive points of
// partial_bp (piece k: start type partial_bp[k].t, end type -partial_bp[k+1].t) -- here as one
// batched GPU launch instead of three thread waves -- links their alignments and prints the two rows
// with print_seq's format.  Every piece is linked (the reference's loop at :343 stops one short).
// p is accepted and ignored.  Returns 0 or a negative psa_status.
int optimal_alignment(char* A, char* B, std::vector<align> partial_bp, size_t m, size_t n, size_t p, double g, double h);
void free_alignment(align* begin);
void print_align(align* begin);
void print_seq(char* A, char* B, align* begin);

#endif
