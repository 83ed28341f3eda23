// FASTA ingest + similarity: same signatures as /root/reference/test_functions/read_test_data.h.
#pragma once
#ifndef PSA_HOST_READ_TEST_DATA_H
#define PSA_HOST_READ_TEST_DATA_H

#include <string>
#include <vector>

// Reads a multi-record FASTA file: '>' lines go to names, the other lines are concatenated into
// the current record.  Prints the reference's progress lines; returns 0 on success, 1 otherwise
// (test_functions/pull_data.cpp:18-71).
int read_and_store_sequences(std::vector<std::string>& names, std::vector<std::string>& sequences, std::string& filename);

// Fraction of positions i < min(len) with sequence1[i] == sequence2[i], divided by max(len)
// (test_functions/pull_data.cpp:97-127).
double sequence_similarity(const std::string& sequence1, const std::string& sequence2);

#endif
