/* oracle/ref_mv_glue.cc — TEST INFRASTRUCTURE, not product code.
 *
 * Thin factory functions around Xapiand's own multivalue classes, compiled INTO
 * oracle/_ref/libxapiand_mv_ref.so next to the reference's src/multivalue/{range,keymaker}.cc
 * (oracle/build_ref.sh).  ref_runner.cc calls these instead of including Xapiand's headers itself: those
 * headers define namespace-scope constants whose constructors live in parts of Xapiand that are not built.
 * Every object is made by the reference's own code paths:
 *   MultipleValueRange           via its constructor, as getNumericQuery does       (src/multivalue/range.cc:121, 343-347)
 *   Multi_MultiValueKeyMaker     via Multi_MultiValueKeyMaker::unserialise          (src/multivalue/keymaker.cc:603-702)
 *   value bytes                  via ::sortable_serialise(long double)              (src/sortable_serialise.cc:41-212,
 *                                = Serialise::integer / positive / floating, src/serialise.h:170-186)
 *   slot bytes                   via StringList::serialise                          (src/serialise_list.h:318-333)
 */
#include <string>
#include <vector>

#include "length.h"
#include "multivalue/keymaker.h"
#include "multivalue/range.h"
#include "serialise_list.h"
#include "sortable_serialise.h"
#include "xapian.h"

/* Link-only stub.  database/data.h (reached through multivalue/keymaker.h → database/utils.h) defines
 * namespace-scope ct_type_t constants — HTTP content types — in every translation unit; their constructor
 * lives in database/data.cc, which needs Xapiand's lz4 / xxhash build configuration and is not compiled
 * here.  Nothing on the matching path reads those constants. */
ct_type_t::ct_type_t(std::string_view) {}

namespace xgmref {

std::string serialise_number(long double v) { return ::sortable_serialise(v); }

std::string serialise_slot(const std::vector<std::string>& sorted_unique_values) {
    return StringList::serialise(sorted_unique_values.begin(), sorted_unique_values.end());
}

Xapian::PostingSource* make_multiple_value_range(unsigned slot, const std::string& start, const std::string& end) {
    /* the constructor getNumericQuery uses (range.cc:121, instantiated for std::string in range.cc:343-347).
     * Not unserialise_with_registry: its `new MultipleValueRange(unserialise_length(*it), *(++it), *(++it))`
     * (range.cc:450) relies on left-to-right argument evaluation, which g++ does not provide. */
    return new MultipleValueRange(slot, std::string(start), std::string(end));
}

/* one SerialiseKey per (slot, reverse) pair, in order — Xapiand's sort on plain field values */
Xapian::KeyMaker* make_key_maker(const std::vector<std::pair<unsigned, bool>>& slots) {
    std::string ser;
    for (const auto& s : slots) {
        ser += serialise_string("SerialiseKey");
        ser += serialise_string(serialise_length(s.first) + serialise_length(s.second ? 1 : 0));
    }
    Multi_MultiValueKeyMaker proto;
    return proto.unserialise(ser, Xapian::Registry());
}

}  // namespace xgmref
