/* xgm_shim.h — the reference-side binding of libxgm.so: what a Xapiand maintainer adds to the vendored Xapian
 * (src/xapian/matcher/) so that Matcher::get_mset hands the local match to the GPU.
 *
 * Seam: Matcher::get_mset (src/xapian/matcher/matcher.cc:544-651) calls get_local_mset (:346-542) once the
 * collated Weight::Internal statistics are set on every LocalSubMatch (:593-597).  The one-line patch
 * (oracle/build_ref.sh --with-xgm applies it with sed to a generated copy of matcher.cc; INTEGRATION.md shows
 * the diff) turns
 *
 *     local_mset = get_local_mset(first, maxitems, check_at_least, ...);
 * into
 *     if (!XGM_SHIM_TRY_LOCAL_MSET(local_mset)) local_mset = get_local_mset(first, maxitems, check_at_least, ...);
 *
 * xgm_shim_try_get_mset returns false — and the reference's own matcher runs, unchanged — whenever the query,
 * the Enquire settings or the database are outside what libxgm covers (XGM_E_UNIMPLEMENTED), when libxgm.so or
 * a CUDA device is missing, or when XGM_SHIM=0.  It throws Xapian::DatabaseModifiedError on XGM_E_STALE and
 * Xapian::DatabaseError on XGM_E_CUDA, which Xapiand's retry loops already handle
 * (src/database/handler.cc:1292-1316).
 */
#ifndef XGM_SHIM_H
#define XGM_SHIM_H

#include "xapian/enquire.h"
#include "xapian/api/enquireinternal.h"

namespace Xapian { class Database; class Query; class Weight; class MatchDecider; class KeyMaker; class MSet; }

struct XgmShimArgs {
    Xapian::doccount first, maxitems, check_at_least;
    const Xapian::Weight* wtscheme;
    const Xapian::MatchDecider* mdecider;
    const Xapian::KeyMaker* sorter;
    Xapian::valueno collapse_key;
    Xapian::doccount collapse_max;
    int percent_threshold;
    double weight_threshold;
    Xapian::Enquire::docid_order order;
    Xapian::valueno sort_key;
    Xapian::Enquire::Internal::sort_setting sort_by;
    bool sort_val_reverse;
    double time_limit;
    size_t n_matchspies, n_locals, n_remotes;
};

bool xgm_shim_try_get_mset(const Xapian::Database& db, const Xapian::Query& query, Xapian::Weight::Internal& stats,
                           const XgmShimArgs& a, Xapian::MSet* out);

/* Introspection for tests and tools (thread-local: the last xgm_shim_try_get_mset call of this thread). */
extern "C" {
int xgm_shim_last_served(void);          /* 1 = answered by libxgm, 0 = left to the reference matcher */
unsigned xgm_shim_last_flags(void);      /* XGM_MSET_* flags of the served MSet */
const char* xgm_shim_last_reason(void);  /* why it was declined ("" when served) */
unsigned long xgm_shim_served_count(void);
unsigned long xgm_shim_declined_count(void);
}

/* the call-site macro: every argument is a parameter or member of Matcher::get_mset */
#define XGM_SHIM_TRY_LOCAL_MSET(local_mset)                                                                        \
    xgm_shim_try_get_mset(db, query, stats,                                                                         \
                          XgmShimArgs{first, maxitems, check_at_least, &wtscheme, mdecider, sorter, collapse_key,   \
                                      collapse_max, percent_threshold, weight_threshold, order, sort_key, sort_by, \
                                      sort_val_reverse, time_limit, matchspies.size(), locals.size(), remotes.size()}, \
                          &(local_mset))

#endif
