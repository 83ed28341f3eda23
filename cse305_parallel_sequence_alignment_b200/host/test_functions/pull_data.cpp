// Host-only data plumbing of the harness (no GPU work): FASTA reader and Hamming-style similarity.
// Observable behaviour (stdout lines, return codes, the similarity value) follows
// /root/reference/test_functions/pull_data.cpp; the implementation is independent.
#include "read_test_data.h"

#include <algorithm>
#include <atomic>
#include <fstream>
#include <iostream>
#include <thread>
#include <unordered_set>

int read_and_store_sequences(std::vector<std::string>& names, std::vector<std::string>& sequences, std::string& filename) {
    std::cout << "Opening data file: " << filename << "\n";
    std::ifstream in(filename);
    if (!in) {
        std::cerr << "Error opening file! Check the file path/name!\n";
        return 1;
    }
    std::cout << "File opened successfully!\n";
    std::cout << "Storing sequences...\n";

    std::string record;
    auto flush_record = [&]() {
        if (!record.empty()) {
            sequences.push_back(record);
            record.clear();
        }
    };
    for (std::string line; std::getline(in, line);) {
        if (!line.empty() && line.front() == '>') {
            flush_record();
            names.push_back(line);
        } else {
            record.append(line);
        }
    }
    flush_record();
    in.close();

    if (names.size() != sequences.size()) {
        std::cout << "Error: mismatch in sequences and names list sizes\n";
        return 1;
    }

    std::cout << "Checking for duplicate sequences...\n";
    std::unordered_set<std::string> seen;
    size_t duplicates = 0;
    for (const std::string& s : sequences) {
        if (!seen.insert(s).second) {
            ++duplicates;
            std::cout << "Duplicate sequence found!\n";
        }
    }
    if (duplicates == 0) std::cout << "No duplicate sequences found.\n";
    else std::cout << "There is at least one duplicate sequence found. Please check your data file.\n";
    std::cout << "Dataset read successfully!\n";
    return 0;
}

double sequence_similarity(const std::string& sequence1, const std::string& sequence2) {
    const size_t shared = std::min(sequence1.size(), sequence2.size());
    const size_t longest = std::max(sequence1.size(), sequence2.size());
    if (longest == 0) return 0.0;
    size_t workers = std::thread::hardware_concurrency();
    if (workers == 0) workers = 1;
    workers = std::min(workers, std::max<size_t>(1, shared / 4096));   // short strings: one worker
    std::atomic<long long> matches(0);
    std::vector<std::thread> pool;
    const size_t span = (shared + workers - 1) / workers;
    for (size_t w = 0; w < workers; ++w) {
        const size_t lo = w * span, hi = std::min(shared, lo + span);
        if (lo >= hi) break;
        pool.emplace_back([&, lo, hi]() {
            long long local = 0;
            for (size_t k = lo; k < hi; ++k) local += (sequence1[k] == sequence2[k]);
            matches.fetch_add(local, std::memory_order_relaxed);
        });
    }
    for (std::thread& t : pool) t.join();
    return static_cast<double>(matches.load()) / static_cast<double>(longest);
}
