// Program entry, same behaviour as /root/reference/main.cpp:6-21: load ./gene_sequences_test and
// run the input-size experiment (one pair, 50 bp x 50 bp).  The other two experiments stay
// available behind PSA_EXPERIMENT=cores|similarity (they are commented out in the reference
// because its O(mn) tables do not fit; here they run on the GPU).
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "alignment_algorithm/main_alignment.h"
#include "test_functions/read_test_data.h"
#include "test_functions/testing.h"

int main(int argc, char** argv) {
    std::string filename = argc > 1 ? argv[1] : "./gene_sequences_test";
    std::vector<std::string> names, sequences;
    read_and_store_sequences(names, sequences, filename);
    if (sequences.size() < 2) return 1;
    const char* which = std::getenv("PSA_EXPERIMENT");
    if (which && std::strcmp(which, "cores") == 0) return test_n_cores_thread(names, sequences);
    if (which && std::strcmp(which, "similarity") == 0) return test_similarity(names, sequences);
    return test_input_size_thread(names, sequences);
}
