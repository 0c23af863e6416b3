// Exercises include/xgm_enquire.hpp (the Xapian-shaped C++ mirror over the C-ABI) on a small synthetic
// index and prints the MSets; tests/test_gpu_cpp_mirror.py compares the output with the oracle.
//   enquire_mirror <ndocs> <vocab> <op: AND|OR> <first> <maxitems> <check_at_least> <term>...
// Environment XGM_MIRROR_FILTER / XGM_MIRROR_NOT / XGM_MIRROR_MAYBE: comma-separated terms wrapped around the
// base as OP_FILTER / OP_AND_NOT / OP_AND_MAYBE (in that order, like the reference's operator nesting).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "xgm_enquire.hpp"

static std::vector<std::string> env_terms(const char* name) {
    std::vector<std::string> out;
    const char* e = getenv(name);
    if (!e) return out;
    std::string cur;
    for (const char* c = e;; ++c) {
        if (*c == ',' || *c == 0) { if (!cur.empty()) out.push_back(cur); cur.clear(); if (*c == 0) break; }
        else cur.push_back(*c);
    }
    return out;
}

static xgm::Query group(xgm::Query::op op, const std::vector<std::string>& ts) {
    return ts.size() == 1 ? xgm::Query(ts[0]) : xgm::Query(op, ts.begin(), ts.end());
}

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
        xgm::Query query = terms.size() == 1 ? xgm::Query(terms[0])
                                             : xgm::Query(op == "OR" ? xgm::Query::OP_OR : xgm::Query::OP_AND, terms.begin(), terms.end());
        const std::vector<std::string> ft = env_terms("XGM_MIRROR_FILTER"), nt = env_terms("XGM_MIRROR_NOT"),
                                       mt = env_terms("XGM_MIRROR_MAYBE");
        if (!ft.empty()) query = xgm::Query(xgm::Query::OP_FILTER, query, group(xgm::Query::OP_AND, ft));
        if (!nt.empty()) query = xgm::Query(xgm::Query::OP_AND_NOT, query, group(xgm::Query::OP_OR, nt));
        if (!mt.empty()) query = xgm::Query(xgm::Query::OP_AND_MAYBE, query, group(xgm::Query::OP_OR, mt));
        enq.set_query(query);
        xgm::MSet m = enq.get_mset(first, maxitems, cal);
        lap("get_mset (searcher creation + search)");
        printf("Q %u %u %u %u %.17g %.17g %d\n", m.size(), m.get_matches_lower_bound(), m.get_matches_estimated(),
               m.get_matches_upper_bound(), m.get_max_possible(), m.get_max_attained(), (int)m.bounds_are_approximate());
        for (auto it = m.begin(); it != m.end(); ++it) printf("%u %.17g %d\n", *it, it.get_weight(), it.get_percent());
        // a shape the device does not cover must be declined loudly, not guessed
        try {
            std::vector<std::string> two{terms[0], terms[0] + "x"};
            xgm::Query bad(xgm::Query::OP_AND_NOT, xgm::Query(xgm::Query::OP_OR, two.begin(), two.end()), xgm::Query(terms[0]));
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
