// Timing harness: same entry points as /root/reference/test_functions/testing.h.
#pragma once
#ifndef PSA_HOST_TESTING_H
#define PSA_HOST_TESTING_H

#include <string>
#include <vector>

// Experiment 1 (input size), sequential and threaded variants (testing.cpp:26-80, :81-166).
int test_input_size(std::vector<std::string>& names, std::vector<std::string>& sequences);
int test_input_size_thread(std::vector<std::string>& names, std::vector<std::string>& sequences);
// Experiment 2 (thread budget p), stub and threaded variants (testing.cpp:174-208, :209-287).
int test_n_cores(std::vector<std::string>& names, std::vector<std::string>& sequences);
int test_n_cores_thread(std::vector<std::string>& names, std::vector<std::string>& sequences);
// Experiment 3 (similarity vs time) (testing.cpp:295-369).
int test_similarity(std::vector<std::string>& names, std::vector<std::string>& sequences);

#endif
