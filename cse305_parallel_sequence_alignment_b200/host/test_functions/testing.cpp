// Benchmark drivers of the harness, re-written against the same observable behaviour as
// /root/reference/test_functions/testing.cpp: which pairs are drawn (unseeded rand(), so the live
// run always aligns records #2 and #15), how the 1-indexed buffers are built, which stdout lines
// appear, and the CSV schemas.  Every alignment goes through main_alignment_function, i.e. through
// libpsa.so on the GPU; each host thread owns a context/stream, so the threads' pairs overlap on
// the device the way they overlapped on CPU cores in the reference.
#include "testing.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <thread>

#include "../alignment_algorithm/main_alignment.h"
#include "read_test_data.h"

namespace {

using Clock = std::chrono::high_resolution_clock;

// new char[len + 2], bases copied to [1..len], slot 0 and the tail left unset (testing.cpp:124-128)
char* one_indexed_copy(const std::string& s, size_t len) {
    char* buf = new char[len + 2];
    std::memcpy(buf + 1, s.data(), len);
    return buf;
}

size_t host_threads() {
    const size_t t = std::thread::hardware_concurrency();
    return t == 0 ? 1 : t;
}

struct PairDraw {
    int first, second;
};

// two rand() calls per pair, modulo (dataset size - 1) -- testing.cpp:115-116
PairDraw draw_pair(int range) {
    PairDraw d;
    d.first = std::rand() % range;
    d.second = std::rand() % range;
    return d;
}

// Shared skeleton of the three threaded experiments: `pairs` tests split in contiguous chunks
// over the host threads (chunk = ceil(pairs / threads), the last thread takes the remainder).
template <typename Body>
void run_chunked(size_t pairs, Body body) {
    const size_t threads = host_threads();
    const size_t chunk = (pairs + threads - 1) / threads;
    std::vector<std::thread> pool;
    for (size_t t = 0; t < threads; ++t) {
        const size_t lo = t * chunk;
        const size_t hi = (t + 1 == threads) ? pairs : lo + chunk;
        pool.emplace_back(body, t, chunk, lo, hi);
    }
    for (std::thread& th : pool) th.join();
}

}  // namespace

int test_input_size(std::vector<std::string>& names, std::vector<std::string>& sequences) {
    std::cout << "Testing with different input sizes\n\n";
    const int batches = 2;
    const size_t increment = 1000;
    const int range = (int)sequences.size() - 1;
    for (int t = 0; t < batches; ++t) {
        std::cout << "Testing batch " << t << "\n\n";
        PairDraw d = draw_pair(range);
        while (d.second == d.first) d.second = std::rand() % range;
        std::cout << "Testing sequences \n" << names[d.first] << "\n and \n" << names[d.second] << "\n\n";
        for (size_t i = 1; i < (size_t)batches; ++i) {
            d = draw_pair(range);
            const std::string& s1 = sequences[d.first];
            const std::string& s2 = sequences[d.second];
            const size_t len = std::min(i * increment, std::min(s1.size(), s2.size()));
            char* a = one_indexed_copy(s1, len);
            char* b = one_indexed_copy(s2, len);
            // the reference prints strlen() of the unterminated buffers here (undefined); print the length
            std::cout << len << " sequence 1 length" << "\n";
            std::cout << len << " sequence 2 length" << "\n";
            std::cout << "input size min " << len << "\n";
            main_alignment_function(a, b, len, len, 32, 1, 2);
            std::cout << "Got res\n";
            delete[] a;
            delete[] b;
        }
    }
    return 0;
}

int test_input_size_thread(std::vector<std::string>& names, std::vector<std::string>& sequences) {
    (void)names;
    std::cout << "Testing with different input sizes\n\n";
    const size_t test_pairs = 1;
    const size_t input_size = 50;
    const int range = (int)sequences.size() - 1;
    std::vector<double> sizes(test_pairs), seconds(test_pairs);

    std::ofstream csv("input_size_testing.csv");
    csv << "Testing with different input sizes\n";
    csv << "Test number,Input size,Execution time\n";

    std::cout << "Starting threads\n";
    const size_t threads = host_threads();
    const size_t chunk = (test_pairs + threads - 1) / threads;
    std::vector<std::thread> pool;
    for (size_t t = 0; t < threads; ++t) {
        const size_t lo = t * chunk;
        const size_t hi = (t + 1 == threads) ? test_pairs : lo + chunk;
        pool.emplace_back([&, lo, hi]() {
            for (size_t k = lo; k < hi && k < test_pairs; ++k) {
                const PairDraw d = draw_pair(range);
                const std::string& s1 = sequences[d.first];
                const std::string& s2 = sequences[d.second];
                const size_t len = std::min(input_size, std::min(s1.size(), s2.size()));
                char* a = one_indexed_copy(s1, len);
                char* b = one_indexed_copy(s2, len);
                const auto t0 = Clock::now();
                main_alignment_function(a, b, len, len, 32, 1, 2);
                std::cout << "Got res\n";
                sizes[k] = (double)len;
                seconds[k] = std::chrono::duration<double>(Clock::now() - t0).count();
                delete[] a;
                delete[] b;
            }
        });
    }
    std::cout << "Joining threads\n";
    for (std::thread& th : pool) th.join();
    std::cout << "Finished threads\n";
    for (size_t k = 0; k < test_pairs; ++k) csv << k << "," << sizes[k] << "," << seconds[k] << "\n";
    csv.close();
    return 0;
}

int test_n_cores(std::vector<std::string>& names, std::vector<std::string>& sequences) {
    // The reference's version times an empty region (testing.cpp:199-201); kept as the same stub.
    std::cout << "Testing with different number of cores\n\n";
    const int batches = 10, n_tests = 5;
    const int range = (int)sequences.size() - 1;
    for (int t = 0; t < batches; ++t) {
        std::cout << "Testing batch " << t << "\n\n";
        PairDraw d = draw_pair(range);
        while (d.second == d.first) d.second = std::rand() % range;
        std::cout << "Testing sequences \n" << names[d.first] << "\n and \n" << names[d.second] << "\n\n";
        for (int i = 1; i <= n_tests; ++i) {
            const int n_cores = std::rand();
            std::cout << "(" << i << "/" << n_tests << ") " << "Testing with number of cores: " << n_cores << "\n";
            const auto t0 = Clock::now();
            const double s = std::chrono::duration<double>(Clock::now() - t0).count();
            std::cout << "Execution time: " << s << " seconds\n";
        }
    }
    return 0;
}

int test_n_cores_thread(std::vector<std::string>& names, std::vector<std::string>& sequences) {
    (void)names;
    std::cout << "Testing with different input sizes\n\n";
    size_t test_pairs = 2000;
    if (const char* e = std::getenv("PSA_TEST_PAIRS")) test_pairs = (size_t)std::atoll(e);   // bounded runs
    const size_t core_increments = 2;
    const int range = (int)sequences.size() - 1;
    std::vector<double> budget(test_pairs), seconds(test_pairs);

    std::ofstream csv("n_cores_testing.csv");
    csv << "Testing with different number of cores\n";
    csv << "Test number,Number of cores,Execution time\n";

    run_chunked(test_pairs, [&](size_t t, size_t chunk, size_t lo, size_t hi) {
        const size_t p = (((t + 1) * chunk) / core_increments) * core_increments;   // testing.cpp:273
        for (size_t k = lo; k < hi && k < test_pairs; ++k) {
            const PairDraw d = draw_pair(range);
            const std::string& s1 = sequences[d.first];
            const std::string& s2 = sequences[d.second];
            const size_t len = std::min(s1.size(), s2.size());    // full-length records
            char* a = one_indexed_copy(s1, len);
            char* b = one_indexed_copy(s2, len);
            budget[k] = (double)p;
            const auto t0 = Clock::now();
            main_alignment_function(a, b, len, len, p, 1, 2);
            seconds[k] = std::chrono::duration<double>(Clock::now() - t0).count();
            delete[] a;
            delete[] b;
        }
    });
    for (size_t k = 0; k < test_pairs; ++k) csv << k << "," << budget[k] << "," << seconds[k] << "\n";
    csv.close();
    return 0;
}

int test_similarity(std::vector<std::string>& names, std::vector<std::string>& sequences) {
    (void)names;
    std::cout << "Testing with similarity computation\n\n";
    size_t test_pairs = 2000;
    if (const char* e = std::getenv("PSA_TEST_PAIRS")) test_pairs = (size_t)std::atoll(e);
    const int range = (int)sequences.size() - 1;
    std::vector<double> similarity(test_pairs), seconds(test_pairs);

    std::ofstream csv("similarity_testing.csv");
    csv << "Testing with similarity computation\n";
    csv << "Test number,Similarity,Execution time\n";

    run_chunked(test_pairs, [&](size_t, size_t, size_t lo, size_t hi) {
        for (size_t k = lo; k < hi && k < test_pairs; ++k) {
            const PairDraw d = draw_pair(range);
            const std::string& s1 = sequences[d.first];
            const std::string& s2 = sequences[d.second];
            const size_t len = std::min(s1.size(), s2.size());
            char* a = one_indexed_copy(s1, len);
            char* b = one_indexed_copy(s2, len);
            similarity[k] = sequence_similarity(s1, s2);
            const auto t0 = Clock::now();
            main_alignment_function(a, b, len, len, 64, 1, 2);
            seconds[k] = std::chrono::duration<double>(Clock::now() - t0).count();
            delete[] a;
            delete[] b;
        }
    });
    for (size_t k = 0; k < test_pairs; ++k) csv << k << "," << similarity[k] << "," << seconds[k] << "\n";
    csv.close();
    return 0;
}
