/* xgm_shim.cc — reference-side binding of libxgm.so, compiled INTO the reference's Xapian library
 * (oracle/build_ref.sh --with-xgm → oracle/_ref/libxapian_ref_xgm.so).  See xgm_shim.h for the seam.
 *
 * What it does for one Matcher::get_mset call:
 *   1. decides whether libxgm covers the request (single local shard, BM25, no decider / spies / collapse /
 *      cut-offs, one of the query shapes of include/xgm.h); anything else is left to the reference matcher;
 *   2. finds — or builds, once per (database uuid, revision) — the HBM index: a glass directory is handed to
 *      xgm_index_open (libxgm parses iamglass + the postlist B-tree itself); any other backend, or a directory
 *      that has moved on to a newer revision, is walked through the public iterators (Database::allterms_begin /
 *      postlist_begin / get_doclength / valuestream_begin);
 *   3. translates the Query tree (api/queryinternal.h) and the collated statistics (weight/weightinternal.h)
 *      into an xgm_query, calls xgm_search, and builds the MSet::Internal (api/msetinternal.h:89-99) from the
 *      result exactly as ProtoMSet::finalise does (matcher/protomset.h:672-682).
 * libxgm.so is dlopen()ed (XGM_LIB, else next to this library): the reference keeps working without it.
 */
#include "config.h"

#include "xgm_shim.h"

#include <dlfcn.h>

#include <atomic>

#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "xapian.h"
#include "xapian/api/msetinternal.h"
#include "xapian/backends/backends.h"
#include "xapian/backends/databaseinternal.h"
#include "xapian/api/queryinternal.h"
#include "xapian/api/result.h"
#include "xapian/common/pack.h"
#include "xapian/common/serialise-double.h"
#include "xapian/weight/weightinternal.h"

#include "../../include/xgm.h"

namespace {

/* ---- libxgm entry points, resolved once ---- */
struct Lib {
    bool ok = false;
    std::string why;
#define XGM_FN(name) decltype(&::name) name = nullptr;
    XGM_FN(xgm_last_error) XGM_FN(xgm_builder_new) XGM_FN(xgm_builder_set_docs) XGM_FN(xgm_builder_add_term)
    XGM_FN(xgm_builder_add_value_slot_serialised) XGM_FN(xgm_builder_set_revision) XGM_FN(xgm_builder_finish)
    XGM_FN(xgm_builder_free) XGM_FN(xgm_index_close) XGM_FN(xgm_searcher_new) XGM_FN(xgm_searcher_free)
    XGM_FN(xgm_search) XGM_FN(xgm_value_key) XGM_FN(xgm_value_key_bytes) XGM_FN(xgm_sort_key_bytes)
    XGM_FN(xgm_abi_version) XGM_FN(xgm_index_open) XGM_FN(xgm_index_info_get)
#undef XGM_FN
};

Lib& lib() {
    static Lib L;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* e = getenv("XGM_SHIM");
        if (e && *e == '0') { L.why = "XGM_SHIM=0"; return; }
        std::string path;
        if (const char* p = getenv("XGM_LIB")) path = p;
        else {
            Dl_info info;
            if (dladdr(reinterpret_cast<void*>(&lib), &info) && info.dli_fname) {
                path = info.dli_fname;
                size_t s = path.rfind('/');
                path = (s == std::string::npos ? std::string(".") : path.substr(0, s)) + "/libxgm.so";
            }
        }
        void* h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!h) { L.why = std::string("dlopen failed: ") + dlerror(); return; }
        bool all = true;
#define XGM_LOAD(name) L.name = reinterpret_cast<decltype(L.name)>(dlsym(h, #name)); all = all && L.name;
        XGM_LOAD(xgm_last_error) XGM_LOAD(xgm_builder_new) XGM_LOAD(xgm_builder_set_docs) XGM_LOAD(xgm_builder_add_term)
        XGM_LOAD(xgm_builder_add_value_slot_serialised) XGM_LOAD(xgm_builder_set_revision) XGM_LOAD(xgm_builder_finish)
        XGM_LOAD(xgm_builder_free) XGM_LOAD(xgm_index_close) XGM_LOAD(xgm_searcher_new) XGM_LOAD(xgm_searcher_free)
        XGM_LOAD(xgm_search) XGM_LOAD(xgm_value_key) XGM_LOAD(xgm_value_key_bytes) XGM_LOAD(xgm_sort_key_bytes)
        XGM_LOAD(xgm_abi_version) XGM_LOAD(xgm_index_open) XGM_LOAD(xgm_index_info_get)
#undef XGM_LOAD
        if (!all) { L.why = "libxgm.so lacks an entry point"; return; }
        if (L.xgm_abi_version() != XGM_ABI_VERSION) { L.why = "libxgm.so ABI version mismatch"; return; }
        L.ok = true;
    });
    return L;
}

/* ---- per-thread record of the last call (tests / tools) ---- */
thread_local int t_served = 0;
thread_local unsigned t_flags = 0;
thread_local std::string t_reason;
std::atomic<unsigned long> g_served{0}, g_declined{0};

bool decline(const char* why) {
    t_served = 0; t_flags = 0; t_reason = why;
    ++g_declined;
    return false;
}

/* ---- index registry: one HBM index per (database uuid, revision) ---- */
struct IndexEntry {
    Xapian::rev revision = 0;
    xgm_index* ix = nullptr;
    std::string failed; /* non-empty: building was tried and failed (do not retry every query) */
};
std::mutex g_reg_mu;
std::map<std::string, std::shared_ptr<IndexEntry>> g_registry;

std::shared_ptr<IndexEntry> build_index(const Xapian::Database& db, Xapian::rev revision) {
    Lib& L = lib();
    auto e = std::make_shared<IndexEntry>();
    e->revision = revision;
    int device = 0;
    if (const char* d = getenv("XGM_DEVICE")) device = atoi(d);
    /* a glass directory is read directly by libxgm (no cursors); only if what is on disk is the very revision
     * this Database object sees — otherwise (a writer committed meanwhile) walk the snapshot we were given */
    {
        std::string path;
        const char* direct = getenv("XGM_SHIM_DIRECT_GLASS");
        if (!(direct && *direct == '0') && db.size() == 1 &&
            db.internal->get_backend_info(&path) == BACKEND_GLASS && !path.empty()) {
            xgm_index* ix = nullptr;
            if (L.xgm_index_open(path.c_str(), device, &ix) == XGM_OK) {
                xgm_index_info info;
                if (L.xgm_index_info_get(ix, &info) == XGM_OK && info.revision == revision) { e->ix = ix; return e; }
                L.xgm_index_close(ix);
            }
        }
    }
    xgm_builder* b = nullptr;
    if (L.xgm_builder_new(&b) != XGM_OK) { e->failed = L.xgm_last_error(); return e; }
    const Xapian::docid last = db.get_lastdocid();
    std::vector<uint32_t> doclen((size_t)last + 1, 0);
    for (auto p = db.postlist_begin(std::string()); p != db.postlist_end(std::string()); ++p) doclen[*p] = db.get_doclength(*p);
    xgm_status st = L.xgm_builder_set_docs(b, db.get_doccount(), last, db.get_total_length(), db.get_doclength_lower_bound(),
                                           db.get_doclength_upper_bound(), doclen.data());
    std::vector<uint32_t> dids, wdfs;
    for (auto t = db.allterms_begin(); st == XGM_OK && t != db.allterms_end(); ++t) {
        const std::string term = *t;
        dids.clear(); wdfs.clear();
        for (auto p = db.postlist_begin(term); p != db.postlist_end(term); ++p) { dids.push_back(*p); wdfs.push_back(p.get_wdf()); }
        uint32_t id;
        st = L.xgm_builder_add_term(b, term.data(), (uint32_t)term.size(), dids.data(), wdfs.data(), (uint32_t)dids.size(),
                                    db.get_collection_freq(term), db.get_wdf_upper_bound(term), &id);
    }
    for (Xapian::valueno slot = 0; st == XGM_OK && slot < 8; ++slot) {
        if (db.get_value_freq(slot) == 0) continue;
        std::vector<uint64_t> off((size_t)last + 2, 0);
        std::string bytes;
        Xapian::docid next = 0;
        for (auto v = db.valuestream_begin(slot); v != db.valuestream_end(slot); ++v) {
            const Xapian::docid d = v.get_docid();
            for (; next <= d; ++next) off[next] = bytes.size();
            bytes += *v;
        }
        for (; next <= last + 1; ++next) off[next] = bytes.size();
        bytes.push_back('\0');
        st = L.xgm_builder_add_value_slot_serialised(b, slot, off.data(), reinterpret_cast<const unsigned char*>(bytes.data()));
    }
    if (st == XGM_OK) st = L.xgm_builder_set_revision(b, revision);
    if (st != XGM_OK) { e->failed = L.xgm_last_error(); L.xgm_builder_free(b); return e; }
    st = L.xgm_builder_finish(b, device, &e->ix); /* consumes the builder */
    if (st != XGM_OK) { e->failed = L.xgm_last_error(); e->ix = nullptr; }
    return e;
}

std::shared_ptr<IndexEntry> index_for(const Xapian::Database& db) {
    const std::string uuid = db.get_uuid();
    const Xapian::rev revision = db.get_revision();
    std::lock_guard<std::mutex> lk(g_reg_mu);
    auto it = g_registry.find(uuid);
    if (it != g_registry.end() && it->second->revision == revision) return it->second;
    /* a new snapshot of the database: the old HBM index stays alive for the searches that still hold it
     * (shared_ptr); it is closed when the last one lets go */
    auto e = build_index(db, revision);
    g_registry[uuid] = e;
    return e;
}

/* ---- per-thread searchers (one Enquire per thread in Xapian, one xgm_searcher per thread here) ---- */
struct SearcherSlot {
    std::shared_ptr<IndexEntry> entry;
    xgm_searcher* s = nullptr;
    uint32_t max_topk = 0;
    ~SearcherSlot() { if (s) lib().xgm_searcher_free(s); }
};
thread_local SearcherSlot t_searcher;

/* ---- query translation ---- */
struct Flat {
    uint32_t op = XGM_OP_AND;
    std::vector<std::string> base, filt, nots, maybe;
    std::vector<uint32_t> wqf, maybe_wqf;
    std::vector<double> factors;
    bool any_factor = false;
    uint32_t filter = XGM_FILTER_NONE, filter_slot = 0;
    std::string range_lo, range_hi;
    bool filter_weighted = false;
    double filter_factor = 1.0;
};

const Xapian::Query::Internal* in(const Xapian::Query& q) { return q.internal.get(); }

/* leaf := LEAF_TERM | OP_SCALE_WEIGHT(leaf) ; the factor is the product of the scales above the term
 * (QueryScaleWeight::postlist, api/queryinternal.cc:1075-1080) */
bool leaf_term(const Xapian::Query& q, std::string& term, uint32_t& wqf, double& factor) {
    if (!in(q)) return false;
    if (q.get_type() == Xapian::Query::OP_SCALE_WEIGHT) {
        std::string s;
        in(q)->serialise(s); /* '\x0d' serialise_double(scale) subquery (queryinternal.cc:2042-2049) */
        const char* p = s.data() + 1;
        const double f = unserialise_double(&p, s.data() + s.size());
        if (!(f >= 0.0)) return false;
        factor *= f;
        return leaf_term(q.get_subquery(0), term, wqf, factor);
    }
    if (q.get_type() != Xapian::Query::LEAF_TERM) return false;
    const auto* t = static_cast<const Xapian::Internal::QueryTerm*>(in(q));
    if (t->get_term().empty()) return false; /* MatchAll */
    term = t->get_term();
    wqf = t->get_wqf();
    return true;
}

bool plain_terms(const Xapian::Query& q, Xapian::Query::op op, std::vector<std::string>& out, std::vector<uint32_t>* wq) {
    std::string term; uint32_t w = 1; double f = 1.0;
    if (leaf_term(q, term, w, f)) {
        if (f != 1.0) return false;
        out.push_back(term);
        if (wq) wq->push_back(w);
        return true;
    }
    if (!in(q) || q.get_type() != op) return false;
    for (size_t i = 0; i < q.get_num_subqueries(); ++i) {
        term.clear(); w = 1; f = 1.0;
        if (!leaf_term(q.get_subquery(i), term, w, f) || f != 1.0) return false;
        out.push_back(term);
        if (wq) wq->push_back(w);
    }
    return !out.empty();
}

/* unserialise_length / StringList element walk of Xapiand (src/length.cc:62-85, src/serialise_list.h:301-356) */
bool sl_length(const char*& p, const char* end, size_t& len) {
    if (p == end) return false;
    len = (unsigned char)*p++;
    if (len == 0xff) {
        len = 0;
        unsigned shift = 0;
        unsigned char ch;
        do {
            if (p == end || shift > 63) return false;
            ch = (unsigned char)*p++;
            len |= (size_t)(ch & 0x7f) << shift;
            shift += 7;
        } while ((ch & 0x80) == 0);
        len += 255;
    }
    return true;
}

/* value-range leaf: stock OP_VALUE_RANGE, or a PostingSource named MultipleValueRange whose serialise() is
 * StringList{serialise_length(slot), start, end} (src/multivalue/range.cc:431-436) */
bool range_leaf(const Xapian::Query& q, Flat& f) {
    if (!in(q)) return false;
    std::string s;
    if (q.get_type() == Xapian::Query::OP_VALUE_RANGE) {
        in(q)->serialise(s); /* 0x20|slot [pack_uint(slot-15)] pack_string(begin) pack_string(end) (queryinternal.cc:1131-1141) */
        const char* p = s.data();
        const char* end = p + s.size();
        unsigned slot = (unsigned char)*p++ & 15u;
        if (slot == 15) { unsigned more; if (!unpack_uint(&p, end, &more)) return false; slot += more; }
        if (!unpack_string(&p, end, f.range_lo) || !unpack_string(&p, end, f.range_hi)) return false;
        f.filter = XGM_FILTER_VALUE_RANGE; f.filter_slot = slot;
        return true;
    }
    if (q.get_type() != Xapian::Query::LEAF_POSTING_SOURCE) return false;
    in(q)->serialise(s); /* 0x0c pack_string(name) pack_string(source->serialise()) (queryinternal.cc:2035-2040) */
    const char* p = s.data() + 1;
    const char* end = s.data() + s.size();
    std::string name, ser;
    if (!unpack_string(&p, end, name) || !unpack_string(&p, end, ser)) return false;
    if (name != "MultipleValueRange" || ser.empty() || ser[0] != '\0') return false;
    const char* e = ser.data() + ser.size();
    const char* c = ser.data() + 1;
    std::string part[3];
    for (int i = 0; i < 3; ++i) {
        size_t len;
        if (!sl_length(c, e, len) || len > (size_t)(e - c)) return false;
        part[i].assign(c, len);
        c += len;
    }
    if (c != e) return false;
    const char* lp = part[0].data();
    size_t slot;
    if (!sl_length(lp, part[0].data() + part[0].size(), slot)) return false;
    f.filter = XGM_FILTER_MULTI_RANGE; f.filter_slot = (uint32_t)slot;
    f.range_lo = part[1]; f.range_hi = part[2];
    return true;
}

/* children of an OP_AND, nested OP_ANDs included: QueryAndLike::postlist_sub_and_like adds them all to ONE
 * AndContext (api/queryinternal.cc:1822-1917), i.e. one MultiAndPostList */
bool and_children(const Xapian::Query& q, Flat& f) {
    for (size_t i = 0; i < q.get_num_subqueries(); ++i) {
        const Xapian::Query sub = q.get_subquery(i);
        std::string term; uint32_t w = 1; double fac = 1.0;
        if (leaf_term(sub, term, w, fac)) {
            f.base.push_back(term); f.wqf.push_back(w); f.factors.push_back(fac); f.any_factor |= fac != 1.0;
            continue;
        }
        if (in(sub) && sub.get_type() == Xapian::Query::OP_AND) {
            if (!and_children(sub, f)) return false;
            continue;
        }
        /* OP_AND(terms..., MultipleValueRange): the source on the weighted side (range.cc:110-125 + query DSL) */
        if (f.filter == XGM_FILTER_NONE && range_leaf(sub, f) && f.filter == XGM_FILTER_MULTI_RANGE) {
            f.filter_weighted = true;
            continue;
        }
        return false;
    }
    return true;
}

/* base := leaf | OP_AND(leaf | OP_AND(...) | range source ...) | OP_OR(leaf...) */
bool base_query(const Xapian::Query& q, Flat& f) {
    std::string term; uint32_t w = 1; double fac = 1.0;
    if (leaf_term(q, term, w, fac)) {
        f.op = XGM_OP_AND;
        f.base.push_back(term); f.wqf.push_back(w); f.factors.push_back(fac); f.any_factor |= fac != 1.0;
        return true;
    }
    if (!in(q)) return false;
    const auto t = q.get_type();
    if (t == Xapian::Query::OP_AND) {
        f.op = XGM_OP_AND;
        return and_children(q, f) && !f.base.empty();
    }
    if (t != Xapian::Query::OP_OR) return false;
    f.op = XGM_OP_OR;
    for (size_t i = 0; i < q.get_num_subqueries(); ++i) {
        term.clear(); w = 1; fac = 1.0;
        if (!leaf_term(q.get_subquery(i), term, w, fac)) return false;
        f.base.push_back(term); f.wqf.push_back(w); f.factors.push_back(fac); f.any_factor |= fac != 1.0;
    }
    return !f.base.empty();
}

bool translate(const Xapian::Query& q, Flat& f) {
    if (!in(q)) return false;
    switch (q.get_type()) {
        case Xapian::Query::OP_AND_MAYBE:
            if (q.get_num_subqueries() != 2) return false;
            return translate(q.get_subquery(0), f) && f.maybe.empty() &&
                   plain_terms(q.get_subquery(1), Xapian::Query::OP_OR, f.maybe, &f.maybe_wqf);
        case Xapian::Query::OP_AND_NOT:
            if (q.get_num_subqueries() != 2) return false;
            return translate(q.get_subquery(0), f) && f.maybe.empty() && f.nots.empty() &&
                   plain_terms(q.get_subquery(1), Xapian::Query::OP_OR, f.nots, nullptr);
        case Xapian::Query::OP_FILTER: {
            if (q.get_num_subqueries() != 2) return false;
            if (!translate(q.get_subquery(0), f) || !f.maybe.empty() || !f.nots.empty()) return false;
            const Xapian::Query r = q.get_subquery(1);
            if (f.filter == XGM_FILTER_NONE && f.filt.empty() && range_leaf(r, f)) return true;
            return f.filt.empty() && plain_terms(r, Xapian::Query::OP_AND, f.filt, nullptr);
        }
        default:
            return base_query(q, f);
    }
}

}  // namespace

extern "C" int xgm_shim_last_served(void) { return t_served; }
extern "C" unsigned xgm_shim_last_flags(void) { return t_flags; }
extern "C" const char* xgm_shim_last_reason(void) { return t_reason.c_str(); }
extern "C" unsigned long xgm_shim_served_count(void) { return g_served.load(); }
extern "C" unsigned long xgm_shim_declined_count(void) { return g_declined.load(); }

bool xgm_shim_try_get_mset(const Xapian::Database& db, const Xapian::Query& query, Xapian::Weight::Internal& stats,
                           const XgmShimArgs& a, Xapian::MSet* out) {
    Lib& L = lib();
    if (!L.ok) return decline(L.why.c_str());
    /* ---- Enquire settings libxgm does not cover → the reference matcher ---- */
    if (a.n_locals != 1 || a.n_remotes != 0 || db.size() != 1) return decline("not a single local shard");
    if (a.mdecider) return decline("MatchDecider");
    if (a.n_matchspies) return decline("MatchSpy");
    if (a.collapse_key != Xapian::BAD_VALUENO) return decline("collapse");
    if (a.percent_threshold != 0 || a.weight_threshold != 0.0) return decline("cut-off");
    if (a.order == Xapian::Enquire::DESCENDING) return decline("descending docid order");
    if (a.time_limit != 0.0) return decline("time limit");
    if (a.wtscheme->name() != "Xapian::BM25Weight") return decline("weighting scheme");
    double k1, k2, k3, b, mnl;
    {
        const std::string ser = a.wtscheme->serialise(); /* k1 k2 k3 b min_normlen (weight/bm25weight.cc:143-152) */
        const char* p = ser.data();
        const char* end = p + ser.size();
        k1 = unserialise_double(&p, end); k2 = unserialise_double(&p, end); k3 = unserialise_double(&p, end);
        b = unserialise_double(&p, end); mnl = unserialise_double(&p, end);
        if (p != end || k2 != 0.0) return decline("BM25 k2 != 0");
    }
    using S = Xapian::Enquire::Internal;
    uint32_t sort_by;
    switch (a.sort_by) {
        case S::REL: sort_by = XGM_SORT_REL; break;
        case S::VAL_REL: sort_by = XGM_SORT_VAL_REL; break;
        case S::VAL: sort_by = XGM_SORT_VAL; break;
        case S::REL_VAL: sort_by = XGM_SORT_REL_VAL; break;
        default: return decline("sort setting");
    }
    /* ---- the sorter: a slot (Enquire::set_sort_by_value*) or Xapiand's Multi_MultiValueKeyMaker with ONE
     * SerialiseKey (serialise(): name + BaseKey::serialise() per slot, keymaker.cc:591-600,33-48) ---- */
    uint32_t sort_slot = a.sort_key, sort_reverse = a.sort_val_reverse, sort_use_max = 0;
    uint64_t sort_missing = 0;
    bool keymaker = false;
    if (sort_by != XGM_SORT_REL && a.sorter) {
        if (a.sorter->name() != "Multi_MultiValueKeyMaker") return decline("KeyMaker");
        const std::string ser = a.sorter->serialise();
        const char* p = ser.data();
        const char* end = p + ser.size();
        size_t len;
        if (!sl_length(p, end, len) || len > (size_t)(end - p) || std::string(p, len) != "SerialiseKey") return decline("sort key type");
        p += len;
        if (!sl_length(p, end, len) || len != (size_t)(end - p)) return decline("more than one sort key");
        size_t slot, rev;
        if (!sl_length(p, end, slot) || !sl_length(p, end, rev) || p != end) return decline("sort key encoding");
        if (a.sort_val_reverse) return decline("reversed key sort");
        keymaker = true;
        sort_slot = (uint32_t)slot; sort_reverse = rev != 0; sort_use_max = rev != 0;
        /* a document without a value sorts as MAX_STR_CMPVALUE "\xff" (forward) or MIN_STR_CMPVALUE (reverse) —
         * std::string("\x00"), which is the EMPTY string (keymaker.h:53-54, keymaker.cc:67-92) */
        if (rev) sort_missing = 0; else L.xgm_value_key("\xff", 1, &sort_missing);
    }
    if (sort_by != XGM_SORT_REL && sort_slot >= 8) return decline("sort slot");
    /* ---- the query ---- */
    Flat f;
    if (!translate(query, f)) return decline("query shape");
    const size_t nall = f.base.size() + f.filt.size() + f.nots.size() + f.maybe.size();
    if (nall > XGM_MAX_TERMS) return decline("too many terms");
    if ((uint64_t)a.first + a.maxitems > XGM_MAX_TOPK) return decline("first + maxitems too large");

    std::shared_ptr<IndexEntry> entry = index_for(db);
    if (!entry->ix) return decline(entry->failed.empty() ? "index unavailable" : entry->failed.c_str());
    SearcherSlot& slot = t_searcher;
    const uint32_t need_topk = std::max<uint32_t>(128u, a.first + a.maxitems);
    if (slot.entry != entry || slot.max_topk < need_topk) {
        if (slot.s) { L.xgm_searcher_free(slot.s); slot.s = nullptr; }
        slot.entry = entry;
        slot.max_topk = need_topk;
        if (L.xgm_searcher_new(entry->ix, 1, need_topk, &slot.s) != XGM_OK) { slot.entry.reset(); return decline(L.xgm_last_error()); }
    }

    std::vector<const char*> terms;
    std::vector<uint32_t> lens, wqf, gtf;
    for (const auto* v : {&f.base, &f.filt, &f.nots, &f.maybe})
        for (const std::string& t : *v) {
            terms.push_back(t.data());
            lens.push_back((uint32_t)t.size());
            Xapian::doccount tf = 0;
            stats.get_stats(t, tf); /* collated over all shards (weightinternal.cc:54-121) */
            gtf.push_back(tf);
        }
    wqf = f.wqf;
    wqf.resize(f.base.size() + f.filt.size() + f.nots.size(), 1);
    wqf.insert(wqf.end(), f.maybe_wqf.begin(), f.maybe_wqf.end());
    xgm_stats xs;
    xs.collection_size = stats.collection_size;
    xs.total_length = stats.total_length;
    xs.termfreq = gtf.data();
    xgm_query xq;
    memset(&xq, 0, sizeof(xq));
    xq.op = f.op; xq.nterms = (uint32_t)f.base.size();
    xq.terms = terms.data(); xq.term_lens = lens.data(); xq.wqf = wqf.data();
    xq.first = a.first; xq.maxitems = a.maxitems; xq.check_at_least = a.check_at_least;
    xq.stats = &xs;
    xq.k1 = k1; xq.k3 = k3; xq.b = b; xq.min_normlen = mnl;
    if (k1 == 0 && k3 == 0 && b == 0 && mnl == 0) return decline("all-zero BM25 parameters"); /* libxgm reads that as 'defaults' */
    xq.nfilter = (uint32_t)f.filt.size(); xq.nnot = (uint32_t)f.nots.size(); xq.nmaybe = (uint32_t)f.maybe.size();
    if (f.any_factor) xq.factors = f.factors.data();
    if (f.filter != XGM_FILTER_NONE) {
        xq.filter = f.filter; xq.filter_slot = f.filter_slot;
        /* the comparison is on 8-byte keys: both bounds must be represented exactly */
        if (!L.xgm_value_key(f.range_lo.data(), f.range_lo.size(), &xq.range_lo) ||
            !L.xgm_value_key(f.range_hi.data(), f.range_hi.size(), &xq.range_hi)) {
            /* a longer upper bound is still exact when only its prefix matters; keep it simple: decline */
            return decline("range bound longer than 8 bytes");
        }
        xq.filter_weighted = f.filter_weighted;
        xq.filter_factor = f.filter_factor;
    }
    xq.sort_by = sort_by; xq.sort_slot = sort_slot; xq.sort_reverse = sort_reverse; xq.sort_use_max = sort_use_max;
    xq.sort_missing_key = sort_missing;
    xq.revision = entry->revision;

    const uint32_t cap = a.maxitems;
    std::vector<uint32_t> docids(cap ? cap : 1);
    std::vector<double> weights(cap ? cap : 1);
    std::vector<uint64_t> keys(cap ? cap : 1);
    xgm_mset_info info;
    const xgm_status st = L.xgm_search(slot.s, &xq, docids.data(), weights.data(), keys.data(), cap, &info);
    if (st == XGM_E_CUDA) throw Xapian::DatabaseError(std::string("xgm: ") + L.xgm_last_error());
    if (st != XGM_OK) return decline(L.xgm_last_error());
    if (info.status == XGM_E_STALE) throw Xapian::DatabaseModifiedError("xgm: HBM index is of another revision");
    if (info.status == XGM_E_CUDA) throw Xapian::DatabaseError("xgm: CUDA failure");
    if (info.status != XGM_OK) return decline("declined by libxgm (XGM_E_UNIMPLEMENTED)");
    if ((info.flags & XGM_MSET_BOUNDS_APPROX) && getenv("XGM_SHIM_EXACT_BOUNDS")) return decline("approximate bounds");

    std::vector<Result> items;
    items.reserve(info.n);
    unsigned char kb[20];
    for (uint32_t i = 0; i < info.n; ++i) {
        items.emplace_back(weights[i], docids[i]);
        if (sort_by != XGM_SORT_REL) {
            /* Result::sort_key: the slot's bytes (matcher.cc:507-517), or the KeyMaker's bytes */
            const size_t n = keymaker ? L.xgm_sort_key_bytes(keys[i], sort_reverse, kb) : L.xgm_value_key_bytes(keys[i], kb);
            items.back().set_sort_key(std::string(reinterpret_cast<const char*>(kb), n));
        }
    }
    *out = Xapian::MSet(new Xapian::MSet::Internal(info.first, info.matches_upper_bound, info.matches_lower_bound,
                                                   info.matches_estimated, info.uncollapsed_upper_bound,
                                                   info.uncollapsed_lower_bound, info.uncollapsed_estimated,
                                                   info.max_possible, info.max_attained, std::move(items),
                                                   info.percent_scale_factor));
    t_served = 1; t_flags = info.flags; t_reason.clear();
    ++g_served;
    return true;
}
