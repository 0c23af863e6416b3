// Exercises include/xgm_enquire.hpp (the Xapian-shaped C++ mirror over the C-ABI) on a small synthetic
// index and prints the MSets; tests/test_gpu_cpp_mirror.py compares the output with the oracle.
//   enquire_mirror <ndocs> <vocab> <op: AND|OR> <first> <maxitems> <check_at_least> <term>...
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "xgm_enquire.hpp"

int main(int argc, char** argv) {
    if (argc < 8) { fprintf(stderr, "usage\n"); return 2; }
    uint32_t ndocs = (uint32_t)atoi(argv[1]), vocab = (uint32_t)atoi(argv[2]);
    std::string op = argv[3];
    uint32_t first = (uint32_t)atoi(argv[4]), maxitems = (uint32_t)atoi(argv[5]), cal = (uint32_t)atoi(argv[6]);
    std::vector<std::string> terms(argv + 7, argv + argc);
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[enquire_mirror] %s: %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    };
    try {
        xgm_index* ix = nullptr;
        xgm::check(xgm_index_build_synthetic(ndocs, vocab, 12345, 1, 0, 0, 0, 4, &ix));
        lap("index build + upload");
        xgm::Database db(ix);
        xgm::Enquire enq(db);
        if (terms.size() == 1) enq.set_query(xgm::Query(terms[0]));
        else enq.set_query(xgm::Query(op == "OR" ? xgm::Query::OP_OR : xgm::Query::OP_AND, terms.begin(), terms.end()));
        xgm::MSet m = enq.get_mset(first, maxitems, cal);
        lap("get_mset (searcher creation + search)");
        printf("Q %u %u %u %u %.17g %.17g %d\n", m.size(), m.get_matches_lower_bound(), m.get_matches_estimated(),
               m.get_matches_upper_bound(), m.get_max_possible(), m.get_max_attained(), (int)m.bounds_are_approximate());
        for (auto it = m.begin(); it != m.end(); ++it) printf("%u %.17g %d\n", *it, it.get_weight(), it.get_percent());
        // a shape the device does not cover must be declined loudly, not guessed
        try {
            xgm::Query bad(xgm::Query::OP_FILTER, xgm::Query(terms[0]), xgm::Query(terms[0]));
            printf("NOT DECLINED\n");
            return 1;
        } catch (const xgm::UnimplementedError&) {
            printf("DECLINED\n");
        }
    } catch (const xgm::Error& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
