// Test driver for the host-side mirrors of the reference's list interface
// (/root/reference/alignment_algorithm/subproblem_alignment.h:8-13 `align`, :33-34 alignment_begin/end;
//  main_alignment.cpp:32-55 print_seq, :202-351 optimal_alignment).  Reads commands from stdin:
//     pair <A> <B> <g> <h>                      -> "corner t1 t2 t3", "nodes i,j,t i,j,t ...", then print_seq's two rows
//     part <A> <B> <g> <h> <k> i j t ... (k points) -> optimal_alignment's two rows over that partition
// and writes one block per command to stdout.  tests/test_host_program.py compares the blocks with the
// compiled reference's Subproblem (oracle/_ref) and the oracle.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../alignment_algorithm/main_alignment.h"

static char* one_indexed(const std::string& s) {
    char* p = new char[s.size() + 2];
    p[0] = '-';
    std::memcpy(p + 1, s.data(), s.size());
    p[s.size() + 1] = 0;
    return p;
}

int main() {
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream in(line);
        std::string cmd, a, b;
        double g, h;
        if (!(in >> cmd >> a >> b >> g >> h)) continue;
        char* A = one_indexed(a);
        char* B = one_indexed(b);
        if (cmd == "pair") {
            align *begin = nullptr, *end = nullptr;
            int corner[3] = {0, 0, 0};
            const int rc = compute_alignment(A, B, a.size(), b.size(), g, h, &begin, &end, corner);
            if (rc != 0) { std::printf("error %d\n", rc); std::fflush(stdout); continue; }
            std::printf("corner %d %d %d\nnodes", corner[0], corner[1], corner[2]);
            size_t count = 0;
            align* lastseen = nullptr;
            for (align* q = begin; q != nullptr; q = q->next) { std::printf(" %zu,%zu,%d", q->i, q->j, q->t); lastseen = q; ++count; }
            std::printf("\ntail %s %zu\n", lastseen == end ? "ok" : "BAD", count);
            std::fflush(stdout);
            print_seq(A, B, begin);
            free_alignment(begin);
        } else if (cmd == "part") {
            size_t k = 0;
            in >> k;
            std::vector<align> bp(k);
            for (size_t x = 0; x < k; ++x) { in >> bp[x].i >> bp[x].j >> bp[x].t; bp[x].next = nullptr; }
            const int rc = optimal_alignment(A, B, bp, a.size(), b.size(), 8, g, h);
            if (rc != 0) std::printf("error %d\n", rc);
        }
        std::printf("end\n");
        std::fflush(stdout);
        delete[] A;
        delete[] B;
    }
    return 0;
}
