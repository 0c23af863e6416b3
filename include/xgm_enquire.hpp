/* xgm_enquire.hpp — header-only C++ mirror of the slice of the Xapian API that sits on the hot path,
 * over the C-ABI of xgm.h.  Same names, argument meaning and error behaviour as the reference
 * (src/xapian/enquire.h, query.h, mset.h, database.h), so code — and tests — written against
 * Xapian::Database / Query / Enquire / MSet read the same here:
 *
 *     xgm::Database db(index);                       // Xapian::Database over an xgm_index
 *     xgm::Enquire enq(db);
 *     enq.set_query(xgm::Query(xgm::Query::OP_AND, terms.begin(), terms.end()));
 *     xgm::MSet m = enq.get_mset(0, 100);
 *     for (auto it = m.begin(); it != m.end(); ++it) use(*it, it.get_weight());
 *
 * Inside the reference itself none of this is needed — the shim of INTEGRATION.md calls xgm_search from
 * Matcher::get_local_mset and the real Xapian classes stay in place.  Query shapes the kernels do not
 * cover throw xgm::UnimplementedError (Xapian::UnimplementedError in the shim).
 */
#ifndef XGM_ENQUIRE_HPP
#define XGM_ENQUIRE_HPP

#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "xgm.h"

namespace xgm {

typedef uint32_t docid;
typedef uint32_t doccount;
typedef uint32_t termcount;
typedef uint32_t valueno;

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };
struct InvalidArgumentError : Error { using Error::Error; };
struct UnimplementedError : Error { using Error::Error; };
struct DatabaseError : Error { using Error::Error; };
struct DatabaseModifiedError : DatabaseError { using DatabaseError::DatabaseError; };

inline void check(xgm_status st) {
    switch (st) {
        case XGM_OK: return;
        case XGM_E_INVALID: throw InvalidArgumentError(xgm_last_error());
        case XGM_E_UNIMPLEMENTED: throw UnimplementedError(xgm_last_error());
        case XGM_E_STALE: throw DatabaseModifiedError(xgm_last_error());
        default: throw DatabaseError(xgm_last_error());
    }
}

/* Xapian::Database (src/xapian/database.h:80): statistics accessors used by the matcher. */
class Database {
    std::shared_ptr<xgm_index> ix;
  public:
    Database() {}
    /* takes ownership of an index built with xgm_builder_* / xgm_index_load_flat */
    explicit Database(xgm_index* index) : ix(index, xgm_index_close) {}
    xgm_index* handle() const { return ix.get(); }
    doccount get_doccount() const { return info().doccount; }
    docid get_lastdocid() const { return info().lastdocid; }
    uint64_t get_total_length() const { return info().total_length; }
    double get_average_length() const { auto i = info(); return i.doccount ? double(i.total_length) / i.doccount : 0.0; }
    termcount get_doclength_lower_bound() const { return info().doclen_lower_bound; }
    termcount get_doclength_upper_bound() const { return info().doclen_upper_bound; }
    doccount get_termfreq(const std::string& t) const { return stats(t).termfreq; }
    uint64_t get_collection_freq(const std::string& t) const { return stats(t).collfreq; }
    termcount get_wdf_upper_bound(const std::string& t) const { return stats(t).wdf_upper_bound; }
    bool term_exists(const std::string& t) const { return stats(t).termfreq != 0; }
    xgm_index_info info() const { xgm_index_info i; check(xgm_index_info_get(ix.get(), &i)); return i; }
    xgm_term_stats stats(const std::string& t) const {
        xgm_term_stats s;
        check(xgm_term_stats_get(ix.get(), t.data(), (uint32_t)t.size(), &s));
        return s;
    }
};

/* Xapian::Query (src/xapian/query.h): leaves, OP_AND / OP_OR of leaves, OP_FILTER with a value range. */
class Query {
  public:
    /* values of Xapian::Query::op (src/xapian/query.h:76-243) */
    enum op { OP_AND = 0, OP_OR = 1, OP_AND_NOT = 2, OP_AND_MAYBE = 4, OP_FILTER = 5, OP_VALUE_RANGE = 8, LEAF_TERM = 100 };
    op type = LEAF_TERM;
    std::vector<std::string> terms;
    std::vector<termcount> wqf;
    bool has_range = false, multi_range = false;
    valueno range_slot = 0;
    uint64_t range_lo = 0, range_hi = 0;
    /* term groups around an AND (or single-term) base: OP_FILTER(q, boolean terms), OP_AND_NOT(q, terms),
     * OP_AND_MAYBE(q, terms) — nesting order FILTER, then AND_NOT, then AND_MAYBE (xgm_query, include/xgm.h) */
    std::vector<std::string> filter_terms, not_terms, maybe_terms;

    Query() {}
    explicit Query(const std::string& term, termcount wqf_ = 1) : terms{term}, wqf{wqf_} {}
    template <class It> Query(op op_, It begin, It end) : type(op_) {
        if (op_ != OP_AND && op_ != OP_OR) throw UnimplementedError("only OP_AND / OP_OR of terms run on the device");
        for (It i = begin; i != end; ++i) { terms.push_back(*i); wqf.push_back(1); }
    }
    Query(op op_, const Query& a, const Query& b) : type(op_) {
        if (op_ == OP_FILTER && b.has_range) {  /* OP_FILTER(q, value-range) — QueryFilter::postlist, queryinternal.cc:2270-2298 */
            if (!b.terms.empty() || a.has_groups()) throw UnimplementedError("OP_FILTER(q, range) over term groups is not covered");
            *this = a;
            has_range = true; multi_range = b.multi_range; range_slot = b.range_slot; range_lo = b.range_lo; range_hi = b.range_hi;
            return;
        }
        if (op_ == OP_FILTER || op_ == OP_AND_NOT || op_ == OP_AND_MAYBE) {
            /* left: an AND of terms (or one term), possibly already carrying inner groups; right: terms only —
             * an AND of boolean terms for OP_FILTER, an OR of terms for the other two */
            const op rkind = op_ == OP_FILTER ? OP_AND : OP_OR;
            const bool left_ok = (a.type == LEAF_TERM || a.type == OP_AND || a.terms.size() == 1) && !a.has_range && !a.terms.empty();
            const bool right_ok = !b.terms.empty() && !b.has_range && !b.has_groups() && (b.terms.size() == 1 || b.type == rkind);
            const bool order_ok = op_ == OP_FILTER ? (a.not_terms.empty() && a.maybe_terms.empty())
                                  : op_ == OP_AND_NOT ? a.maybe_terms.empty() : true;
            if (!left_ok || !right_ok || !order_ok) throw UnimplementedError("operator nesting not covered by the device matcher");
            *this = a;
            type = a.terms.size() == 1 ? LEAF_TERM : OP_AND;
            std::vector<std::string>& dst = op_ == OP_FILTER ? filter_terms : op_ == OP_AND_NOT ? not_terms : maybe_terms;
            dst.insert(dst.end(), b.terms.begin(), b.terms.end());
            return;
        }
        if (op_ != OP_AND && op_ != OP_OR) throw UnimplementedError("operator not covered by the device matcher");
        if ((a.type != LEAF_TERM && a.type != op_) || (b.type != LEAF_TERM && b.type != op_) || a.has_range || b.has_range ||
            a.has_groups() || b.has_groups())
            throw UnimplementedError("nested operators of different kinds are not covered");
        terms = a.terms; terms.insert(terms.end(), b.terms.begin(), b.terms.end());
        wqf = a.wqf; wqf.insert(wqf.end(), b.wqf.begin(), b.wqf.end());
    }
    /* OP_VALUE_RANGE over decoded numeric keys; multi = Xapiand's MultipleValueRange semantics */
    static Query value_range(valueno slot, uint64_t lo, uint64_t hi, bool multi = false) {
        Query q; q.type = OP_VALUE_RANGE; q.has_range = true; q.multi_range = multi; q.range_slot = slot; q.range_lo = lo; q.range_hi = hi;
        return q;
    }
    bool has_groups() const { return !filter_terms.empty() || !not_terms.empty() || !maybe_terms.empty(); }
    bool empty() const { return terms.empty(); }
};

class MSet;

class MSetIterator {
    const MSet* m = nullptr;
    doccount i = 0;
  public:
    MSetIterator() {}
    MSetIterator(const MSet* m_, doccount i_) : m(m_), i(i_) {}
    docid operator*() const;
    double get_weight() const;
    doccount get_rank() const;
    int get_percent() const;
    uint64_t get_sort_key() const;
    MSetIterator& operator++() { ++i; return *this; }
    bool operator==(const MSetIterator& o) const { return i == o.i; }
    bool operator!=(const MSetIterator& o) const { return i != o.i; }
};

/* Xapian::MSet (src/xapian/mset.h:145-405) over the fields of MSet::Internal. */
class MSet {
    friend class Enquire;
    friend class MSetIterator;
    std::vector<docid> docids;
    std::vector<double> weights;
    std::vector<uint64_t> keys;
    xgm_mset_info info{};
  public:
    doccount size() const { return (doccount)docids.size(); }
    bool empty() const { return docids.empty(); }
    MSetIterator begin() const { return MSetIterator(this, 0); }
    MSetIterator end() const { return MSetIterator(this, size()); }
    MSetIterator operator[](doccount i) const { return MSetIterator(this, i); }
    doccount get_firstitem() const { return info.first; }
    doccount get_matches_lower_bound() const { return info.matches_lower_bound; }
    doccount get_matches_upper_bound() const { return info.matches_upper_bound; }
    /* MSet::get_matches_estimated rounds: api/mset.cc:145-153 + api/roundestimate.h:35-64 */
    doccount get_matches_estimated() const {
        uint32_t m = info.matches_lower_bound, M = info.matches_upper_bound, e = info.matches_estimated;
        uint32_t D = M - m;
        if (D == 0 || e == 0) return e;
        uint32_t r = 1;
        for (int k = (int)std::log10((double)D); k > 0; --k) r *= 10;
        while (r > e) r /= 10;
        uint32_t R = e / r * r;
        if (R < m) R += r;
        else if (R > M) R -= r;
        else if (R < e && r % 2 == 0 && e - R == r / 2) { if (e - m < M - e) R += r; }
        if (R < m || R > M) R = e;
        return R;
    }
    doccount get_uncollapsed_matches_lower_bound() const { return info.uncollapsed_lower_bound; }
    doccount get_uncollapsed_matches_estimated() const { return info.uncollapsed_estimated; }
    doccount get_uncollapsed_matches_upper_bound() const { return info.uncollapsed_upper_bound; }
    double get_max_possible() const { return info.max_possible; }
    double get_max_attained() const { return info.max_attained; }
    /* MSet::Internal::convert_to_percent, api/mset.cc:334-361 */
    int convert_to_percent(double weight) const {
        int percent;
        if (info.percent_scale_factor == 0.0) {
            percent = 100;
        } else if (weight <= 0.0) {
            percent = 0;
        } else {
            percent = int(weight * info.percent_scale_factor + 100.0 * 2.220446049250313e-16);
            if (percent <= 0) percent = 1;
            else if (percent > 100) percent = 100;
        }
        return percent;
    }
    bool bounds_are_approximate() const { return info.flags & XGM_MSET_BOUNDS_APPROX; }
};

inline docid MSetIterator::operator*() const { return m->docids[i]; }
inline double MSetIterator::get_weight() const { return m->weights[i]; }
inline doccount MSetIterator::get_rank() const { return m->info.first + i; }
inline int MSetIterator::get_percent() const { return m->convert_to_percent(m->weights[i]); }
inline uint64_t MSetIterator::get_sort_key() const { return i < m->keys.size() ? m->keys[i] : 0; }

/* Xapian::Enquire (src/xapian/enquire.h, api/enquire.cc:71-298).  One per thread, like the reference. */
class Enquire {
    Database db;
    Query query;
    std::shared_ptr<xgm_searcher> searcher;
    uint32_t searcher_topk = 0;
    uint32_t sort_by = XGM_SORT_REL, sort_slot = 0;
    bool sort_reverse = false, sort_use_max = false;
    bool have_stats = false;
    xgm_stats stats{};
    std::vector<uint32_t> stats_tf;
  public:
    explicit Enquire(const Database& db_) : db(db_) {}
    void set_query(const Query& q) { query = q; }
    const Query& get_query() const { return query; }
    void set_sort_by_relevance() { sort_by = XGM_SORT_REL; }
    void set_sort_by_value(valueno slot, bool reverse) { sort_by = XGM_SORT_VAL; sort_slot = slot; sort_reverse = reverse; }
    void set_sort_by_value_then_relevance(valueno slot, bool reverse) { sort_by = XGM_SORT_VAL_REL; sort_slot = slot; sort_reverse = reverse; }
    void set_sort_by_relevance_then_value(valueno slot, bool reverse) { sort_by = XGM_SORT_REL_VAL; sort_slot = slot; sort_reverse = reverse; }
    /* Xapiand's Multi_MultiValueKeyMaker on one numeric field: ascending uses the smallest value of the
     * slot, `-field` the largest (src/multivalue/keymaker.cc:67-92,704-757) */
    void set_sort_by_key_then_relevance(valueno slot, bool descending) {
        sort_by = XGM_SORT_VAL_REL; sort_slot = slot; sort_reverse = descending; sort_use_max = descending;
    }
    /* collated statistics of all shards (Enquire::set_prepared_mset in Xapiand's two-phase scheme) */
    void set_global_stats(doccount collection_size, uint64_t total_length, const std::vector<doccount>& termfreqs) {
        stats_tf = termfreqs; stats.collection_size = collection_size; stats.total_length = total_length; have_stats = true;
    }
    void clear_global_stats() { have_stats = false; }

    MSet get_mset(doccount first, doccount maxitems, doccount checkatleast = 0) {
        MSet out;
        out.info.first = first;
        if (query.empty()) return out;  /* api/enquire.cc:402-406 */
        const uint32_t docs = db.get_doccount();
        const uint32_t f = first < docs ? first : docs;
        const uint32_t need = f + (maxitems < docs - f ? maxitems : docs - f);
        if (!searcher || need > searcher_topk) {
            xgm_searcher* s = nullptr;
            searcher_topk = need < 16 ? 16 : (need > XGM_MAX_TOPK ? XGM_MAX_TOPK : need);
            check(xgm_searcher_new(db.handle(), 1, searcher_topk, &s));
            searcher.reset(s, xgm_searcher_free);
        }
        std::vector<const char*> tp;
        std::vector<uint32_t> tl;
        std::vector<termcount> wq(query.wqf);
        for (auto& t : query.terms) { tp.push_back(t.data()); tl.push_back((uint32_t)t.size()); }
        for (const std::vector<std::string>* g : {&query.filter_terms, &query.not_terms, &query.maybe_terms})
            for (auto& t : *g) { tp.push_back(t.data()); tl.push_back((uint32_t)t.size()); wq.push_back(1); }
        xgm_query q{};
        q.op = query.type == Query::OP_OR ? XGM_OP_OR : XGM_OP_AND;
        q.nterms = (uint32_t)query.terms.size();
        q.nfilter = (uint32_t)query.filter_terms.size(); q.nnot = (uint32_t)query.not_terms.size();
        q.nmaybe = (uint32_t)query.maybe_terms.size();
        q.terms = tp.data(); q.term_lens = tl.data(); q.wqf = wq.data();
        q.first = first; q.maxitems = maxitems; q.check_at_least = checkatleast;
        if (have_stats) {
            /* one termfreq per term in the order base, filter, not, maybe (the weighted groups use theirs) */
            if (stats_tf.size() < tp.size()) throw InvalidArgumentError("global statistics must cover every query term");
            stats.termfreq = stats_tf.data(); q.stats = &stats;
        }
        if (query.has_range) {
            q.filter = query.multi_range ? XGM_FILTER_MULTI_RANGE : XGM_FILTER_VALUE_RANGE;
            q.filter_slot = query.range_slot; q.range_lo = query.range_lo; q.range_hi = query.range_hi;
        }
        q.sort_by = sort_by; q.sort_slot = sort_slot; q.sort_reverse = sort_reverse; q.sort_use_max = sort_use_max;
        out.docids.resize(searcher_topk); out.weights.resize(searcher_topk); out.keys.resize(searcher_topk);
        check(xgm_search(searcher.get(), &q, out.docids.data(), out.weights.data(), out.keys.data(), searcher_topk, &out.info));
        if (out.info.status != XGM_OK) {
            if (out.info.status == XGM_E_UNIMPLEMENTED) throw UnimplementedError("query shape not covered by the device matcher");
            throw InvalidArgumentError("query rejected by the device matcher");
        }
        out.docids.resize(out.info.n); out.weights.resize(out.info.n); out.keys.resize(out.info.n);
        return out;
    }
};

}  // namespace xgm
#endif
