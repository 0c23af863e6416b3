/* xgm_glass.cu — direct reader of a glass database directory (host code).
 *
 * Reads `iamglass` and `postlist.glass` itself — no Xapian library, no cursors — and feeds the index builder:
 * the posting lists, the document-length list, the value streams and the statistics all live in the postlist
 * table.  What is restated, with the reference lines it follows (src/xapian/…):
 *
 *   version file   backends/glass/glass_version.cc:60-71,97-172 (magic, format version, uuid, revision, one
 *                  RootInfo per table :451-466), statistics :193-234
 *   B-tree block   backends/glass/glass_table.h:66-110 (REVISION / LEVEL / DIR_END, directory of D2 offsets from
 *                  byte 11), leaf item I2 K1 key [X2] tag :126-175 with the first / last / compressed flags in the
 *                  top bits of I2 :112-120, branch item = block number, K1 key, X2 :257-300; every multi-byte field
 *                  big-endian (common/wordaccess.h:111-123)
 *   keys           common/pack.h:523-535 (pack_string_preserving_sort), :185-220,232-287
 *                  (pack_uint_preserving_sort), :569-593 (term [+ first docid of the chunk]; "\0\xe0" = document
 *                  lengths), backends/glass/glass_values.h:41-47 ("\0\xd8" slot docid = value chunk),
 *                  glass_values.cc:60-68 ("\0\xd0" slot = value statistics)
 *   posting chunk  backends/glass/glass_postlist.cc:86-148,373-397,677-695: first chunk = termfreq, collfreq,
 *                  first docid - 1; every chunk = is_last, last docid - first docid, wdf, (docid delta - 1, wdf)*
 *   value chunk    backends/glass/glass_values.cc:69-91: value, (docid delta - 1, value)*  with pack_string values
 *   wdf bound      backends/glass/glass_postlist.cc:175-190 + glass_database.cc:822-829
 *
 * The leaf blocks are visited in key order by a depth-first walk from the root (the B-tree has no sibling
 * links).  Pinned on the CPU: tests/test_glass_reader.py compares xgm_glass_export_flat byte for byte with
 * `ref_runner export`, which walks the same database through the reference's public iterators.
 */
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "../../include/xgm.h"

xgm_status xgm_fail(xgm_status st, const char* fmt, ...);                                   /* xgm_host.cu */
xgm_status xgm_builder_force_wdf_ub(xgm_builder* b, uint32_t term_id, uint32_t wdf_ub);    /* xgm_host.cu */

namespace {

struct Reader {
    const unsigned char* p;
    const unsigned char* end;
    bool uint(uint64_t& v) { /* unpack_uint, common/pack.h:327-388: little-endian base-128, high bit = more */
        v = 0;
        unsigned shift = 0;
        while (p != end) {
            const unsigned char c = *p++;
            if (shift < 64) v |= (uint64_t)(c & 0x7f) << shift;
            shift += 7;
            if (!(c & 0x80)) return true;
        }
        return false;
    }
    bool str(std::string& s) { /* unpack_string, pack.h:463-484 */
        uint64_t len;
        if (!uint(len) || len > (uint64_t)(end - p)) return false;
        s.assign(reinterpret_cast<const char*>(p), (size_t)len);
        p += len;
        return true;
    }
    bool boolean(bool& b) { /* unpack_bool: one byte '0' / '1' */
        if (p == end || (*p != '0' && *p != '1')) return false;
        b = *p++ == '1';
        return true;
    }
    bool uint_sort(uint64_t& v) { /* unpack_uint_preserving_sort, pack.h:232-287 */
        if (p == end) return false;
        unsigned char lb = *p++;
        if (lb < 0x80) {
            if (p == end) return false;
            v = ((uint64_t)lb << 8) | *p++;
            return true;
        }
        if (lb == 0xff) return false;
        size_t len = 2;
        for (unsigned char m = 0x40; lb & m; m >>= 1) ++len;
        if ((size_t)(end - p) < len) return false;
        const unsigned mask = 0xffu << (9 - len);
        lb &= (unsigned char)~mask;
        if (len > 8) return false;
        uint64_t r = lb;
        for (size_t i = 0; i < len; ++i) r = (r << 8) | *p++;
        v = r;
        return true;
    }
};

inline uint32_t be2(const unsigned char* b) { return ((uint32_t)b[0] << 8) | b[1]; }
inline uint32_t be4(const unsigned char* b) { return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3]; }

struct GlassStats {
    uint64_t revision = 0, doccount = 0, lastdocid = 0, doclen_lb = 0, doclen_ub = 0, wdf_ub = 0, total_doclen = 0;
    uint64_t root = 0, level = 0, blocksize = 0, num_entries = 0;
    bool root_is_fake = false;
};

const unsigned char GLASS_MAGIC[14] = {0x0f, 0x0d, 'X', 'a', 'p', 'i', 'a', 'n', ' ', 'G', 'l', 'a', 's', 's'};

bool read_version(const std::string& dir, GlassStats& g, std::string& err) {
    FILE* f = fopen((dir + "/iamglass").c_str(), "rb");
    if (!f) { err = "cannot open " + dir + "/iamglass"; return false; }
    unsigned char buf[512];
    const size_t n = fread(buf, 1, sizeof(buf), f);
    fclose(f);
    if (n < 33 || memcmp(buf, GLASS_MAGIC, 14) != 0) { err = "not a glass version file"; return false; }
    const unsigned version = ((unsigned)buf[14] << 8) | buf[15];
    if (version != (((2016u - 2014u) << 9) | (3u << 5) | 14u)) { err = "unsupported glass format version"; return false; }
    Reader r{buf + 16 + 16, buf + n}; /* magic + version, then the 16-byte uuid */
    if (!r.uint(g.revision)) { err = "bad revision"; return false; }
    for (int table = 0; table < 6; ++table) { /* POSTLIST, DOCDATA, TERMLIST, POSITION, SPELLING, SYNONYM */
        uint64_t root, val, entries, bs, cmin;
        std::string fl;
        if (!r.uint(root) || !r.uint(val) || !r.uint(entries) || !r.uint(bs) || !r.uint(cmin) || !r.str(fl)) {
            err = "bad root info";
            return false;
        }
        if (table == 0) {
            g.root = root; g.level = val >> 2; g.root_is_fake = val & 1; g.num_entries = entries; g.blocksize = bs << 11;
        }
    }
    uint64_t oldest, spelling;
    if (r.p == r.end) return true; /* empty database: all statistics zero */
    if (!r.uint(g.doccount) || !r.uint(g.lastdocid) || !r.uint(g.doclen_lb) || !r.uint(g.wdf_ub) || !r.uint(g.doclen_ub) ||
        !r.uint(oldest) || !r.uint(g.total_doclen) || !r.uint(spelling)) {
        err = "bad statistics";
        return false;
    }
    g.lastdocid += g.doccount; /* stored as the difference (glass_version.cc:228-233) */
    g.doclen_ub += g.wdf_ub;
    return true;
}

/* what the walk hands over, in key order */
struct Sink {
    std::function<bool(uint32_t did, uint32_t len)> doclen;
    std::function<bool(const std::string& term, uint64_t termfreq, uint64_t collfreq)> term_begin;
    std::function<bool(uint32_t did, uint32_t wdf)> posting;
    std::function<bool(uint32_t slot, uint32_t did, const std::string& value)> value;
};

struct Walker {
    FILE* f = nullptr;
    GlassStats g;
    Sink* sink = nullptr;
    std::string err;
    /* tag being assembled from its components */
    std::string key, tag;
    bool have = false;
    /* term state across chunks */
    bool in_term = false;

    bool fail(const char* m) { err = m; return false; }

    bool deliver() {
        const unsigned char* k = reinterpret_cast<const unsigned char*>(key.data());
        const size_t kl = key.size();
        Reader t{reinterpret_cast<const unsigned char*>(tag.data()), reinterpret_cast<const unsigned char*>(tag.data()) + tag.size()};
        if (kl >= 2 && k[0] == 0 && k[1] != 0xff) {
            if (k[1] == 0xe0) return chunk(t, kl == 2, k + 2, k + kl, true);
            if (k[1] == 0xd8) {
                Reader kr{k + 2, k + kl};
                uint64_t slot, did;
                if (!kr.uint(slot) || !kr.uint_sort(did)) return fail("bad value chunk key");
                std::string v;
                if (!t.str(v)) return fail("bad value chunk");
                for (;;) {
                    if (!sink->value((uint32_t)slot, (uint32_t)did, v)) return false;
                    if (t.p == t.end) break;
                    uint64_t delta;
                    if (!t.uint(delta) || !t.str(v)) return fail("bad value chunk");
                    did += delta + 1;
                }
                return true;
            }
            return true; /* "\0\xc0" user metadata, "\0\xd0" value statistics: not needed */
        }
        /* a term: unescape pack_string_preserving_sort; what follows the terminator is the chunk's first docid */
        std::string term;
        size_t i = 0;
        bool terminated = false;
        while (i < kl) {
            const unsigned char c = k[i++];
            if (c == 0) {
                if (i < kl && k[i] == 0xff) { ++i; term.push_back('\0'); continue; }
                terminated = true;
                break;
            }
            term.push_back((char)c);
        }
        const bool first = !terminated;
        if (first) cur_term = term;
        return chunk(t, first, k + i, k + kl, false);
    }

    std::string cur_term;

    /* one posting / doclen chunk: `first` has the list header in its tag, the others their first docid in the key */
    bool chunk(Reader& t, bool first, const unsigned char* kp, const unsigned char* kend, bool is_doclen) {
        uint64_t did;
        if (first) {
            uint64_t tf, cf, d0;
            if (!t.uint(tf) || !t.uint(cf) || !t.uint(d0)) return fail("bad first chunk");
            did = d0 + 1;
            if (!is_doclen && !sink->term_begin(cur_term, tf, cf)) return false;
        } else {
            Reader kr{kp, kend};
            if (!kr.uint_sort(did)) return fail("bad chunk key");
        }
        bool is_last;
        uint64_t span, wdf;
        if (!t.boolean(is_last) || !t.uint(span)) return fail("bad chunk header");
        (void)span;
        if (!t.uint(wdf)) return fail("bad chunk");
        for (;;) {
            if (did > 0xfffffffeull || wdf > 0xffffffffull) return fail("docid / wdf beyond 32 bits");
            if (!(is_doclen ? sink->doclen((uint32_t)did, (uint32_t)wdf) : sink->posting((uint32_t)did, (uint32_t)wdf))) return false;
            if (t.p == t.end) break;
            uint64_t delta;
            if (!t.uint(delta) || !t.uint(wdf)) return fail("bad chunk");
            did += delta + 1;
        }
        return true;
    }

    const unsigned char* block_base = nullptr;
    size_t block_size = 0;

    bool leaf_item(const unsigned char* it) {
        const unsigned flags = it[0];
        const int size = (int)(be2(it) & 0x1fffu) + 3;
        const int klen = it[2];
        int cd = 2 + 1 + klen;
        const bool firstc = flags & 0x20, lastc = flags & 0x40;
        if (flags & 0x80) return fail("compressed tag in the postlist table"); /* compress_min is 0 there (glass_version.cc:401-408) */
        if (!firstc) cd += 2;
        if (cd > size) return fail("bad item");
        if ((size_t)(it - block_base) + (size_t)size > block_size) return fail("item beyond its block");
        if (klen == 0) return true; /* the null item that opens the first leaf block (LeafItem_wr::fake_root_item) */
        if (firstc) {
            key.assign(reinterpret_cast<const char*>(it + 3), (size_t)klen);
            tag.clear();
            have = true;
        } else if (!have) {
            return fail("tag component without a first one");
        }
        tag.append(reinterpret_cast<const char*>(it + cd), (size_t)(size - cd));
        if (lastc) {
            have = false;
            return deliver();
        }
        return true;
    }

    bool walk(uint64_t blockno, int level_expected) {
        std::vector<unsigned char> blk(g.blocksize);
        if (fseeko(f, (off_t)(blockno * g.blocksize), SEEK_SET) != 0 || fread(blk.data(), 1, blk.size(), f) != blk.size())
            return fail("cannot read block");
        const int level = blk[4];
        const int dir_end = (int)be2(blk.data() + 9);
        if (level != level_expected || dir_end < 11 || (size_t)dir_end > blk.size()) return fail("bad block header");
        for (int c = 11; c < dir_end; c += 2) {
            const int off = (int)be2(blk.data() + c);
            if ((size_t)off + (level == 0 ? 3 : 7) > blk.size()) return fail("bad directory entry");
            const unsigned char* it = blk.data() + off;
            if (level == 0) {
                block_base = blk.data(); block_size = blk.size();
                if (!leaf_item(it)) return false;
            } else {
                if (!walk(be4(it), level - 1)) return false;
            }
        }
        return true;
    }
};

bool glass_walk(const std::string& dir, GlassStats& g, Sink& sink, std::string& err) {
    if (!read_version(dir, g, err)) return false;
    if (g.root_is_fake || g.num_entries == 0) return true; /* empty postlist table */
    Walker w;
    w.g = g;
    w.sink = &sink;
    w.f = fopen((dir + "/postlist.glass").c_str(), "rb");
    if (!w.f) { err = "cannot open " + dir + "/postlist.glass"; return false; }
    const bool ok = w.walk(g.root, (int)g.level);
    fclose(w.f);
    if (!ok) err = w.err.empty() ? "walk aborted" : w.err;
    return ok;
}

/* everything the builder (or the flat export) needs, gathered in one pass */
struct Loaded {
    GlassStats g;
    std::vector<uint32_t> doclen;
    struct Term { std::string name; uint64_t tf, cf; size_t begin; };
    std::vector<Term> terms;
    std::vector<uint32_t> dids, wdfs;
    struct Val { uint32_t did; std::string v; };
    std::vector<std::vector<Val>> slots;
};

bool load(const std::string& dir, Loaded& L, std::string& err) {
    Sink s;
    bool sized = false;
    s.doclen = [&](uint32_t d, uint32_t l) {
        if (!sized) { L.doclen.assign((size_t)L.g.lastdocid + 1, 0); sized = true; }
        if (d >= L.doclen.size()) { err = "docid beyond lastdocid"; return false; }
        L.doclen[d] = l;
        return true;
    };
    s.term_begin = [&](const std::string& t, uint64_t tf, uint64_t cf) {
        L.terms.push_back(Loaded::Term{t, tf, cf, L.dids.size()});
        return true;
    };
    s.posting = [&](uint32_t d, uint32_t w) { L.dids.push_back(d); L.wdfs.push_back(w); return true; };
    s.value = [&](uint32_t slot, uint32_t d, const std::string& v) {
        if (slot >= 4096) { err = "value slot out of range"; return false; }
        if (L.slots.size() <= slot) L.slots.resize(slot + 1);
        L.slots[slot].push_back(Loaded::Val{d, v});
        return true;
    };
    /* the statistics are needed before the first doclen arrives: read_version fills L.g first inside glass_walk */
    if (!read_version(dir, L.g, err)) return false;
    GlassStats g2;
    if (!glass_walk(dir, g2, s, err)) return false;
    if (!sized) L.doclen.assign((size_t)L.g.lastdocid + 1, 0);
    return true;
}

uint32_t term_wdf_ub(const Loaded& L, size_t t) {
    const Loaded::Term& T = L.terms[t];
    uint64_t ub;
    if (T.cf == 0 || T.tf == 1) ub = T.cf;
    else {
        const uint64_t fw = L.wdfs[T.begin];
        ub = T.cf - fw > fw ? T.cf - fw : fw;
    }
    return (uint32_t)(ub < L.g.wdf_ub ? ub : L.g.wdf_ub);
}

}  // namespace

/* XGMFLAT1, the format of `oracle/ref_runner export` (which walks the same database through the reference's
 * public iterators): the CPU pin of this reader. */
extern "C" xgm_status xgm_glass_export_flat(const char* glass_path, const char* out_path) {
    if (!glass_path || !out_path) return xgm_fail(XGM_E_INVALID, "null argument");
    Loaded L;
    std::string err;
    if (!load(glass_path, L, err)) return xgm_fail(XGM_E_IO, "%s: %s", glass_path, err.c_str());
    FILE* f = fopen(out_path, "wb");
    if (!f) return xgm_fail(XGM_E_IO, "cannot write %s", out_path);
    auto w32 = [&](uint32_t v) { fwrite(&v, 4, 1, f); };
    auto w64 = [&](uint64_t v) { fwrite(&v, 8, 1, f); };
    uint32_t nslots = 0;
    for (size_t s = 0; s < L.slots.size() && s < 8; ++s) nslots += !L.slots[s].empty();
    fwrite("XGMFLAT1", 8, 1, f);
    w32((uint32_t)L.g.doccount); w32((uint32_t)L.g.lastdocid); w64(L.g.total_doclen);
    w32((uint32_t)L.terms.size()); w32(nslots); w32((uint32_t)L.g.doclen_lb); w32((uint32_t)L.g.doclen_ub);
    fwrite(L.doclen.data(), 4, L.doclen.size(), f);
    for (size_t t = 0; t < L.terms.size(); ++t) {
        const Loaded::Term& T = L.terms[t];
        const size_t end = t + 1 < L.terms.size() ? L.terms[t + 1].begin : L.dids.size();
        w32((uint32_t)T.name.size()); fwrite(T.name.data(), 1, T.name.size(), f);
        w32((uint32_t)T.tf); w64(T.cf); w32(term_wdf_ub(L, t)); w32((uint32_t)(end - T.begin));
        fwrite(L.dids.data() + T.begin, 4, end - T.begin, f);
        fwrite(L.wdfs.data() + T.begin, 4, end - T.begin, f);
    }
    for (size_t s = 0; s < L.slots.size() && s < 8; ++s) {
        if (L.slots[s].empty()) continue;
        w32((uint32_t)s); w32((uint32_t)L.slots[s].size());
        for (const auto& v : L.slots[s]) { w32(v.did); w32((uint32_t)v.v.size()); fwrite(v.v.data(), 1, v.v.size(), f); }
    }
    fclose(f);
    return XGM_OK;
}

extern "C" xgm_status xgm_glass_revision(const char* glass_path, uint64_t* revision, uint32_t* doccount, uint32_t* lastdocid) {
    if (!glass_path) return xgm_fail(XGM_E_INVALID, "null argument");
    GlassStats g;
    std::string err;
    if (!read_version(glass_path, g, err)) return xgm_fail(XGM_E_IO, "%s: %s", glass_path, err.c_str());
    if (revision) *revision = g.revision;
    if (doccount) *doccount = (uint32_t)g.doccount;
    if (lastdocid) *lastdocid = (uint32_t)g.lastdocid;
    return XGM_OK;
}

extern "C" xgm_status xgm_index_open(const char* glass_path, int device, xgm_index** out) {
    if (!glass_path || !out) return xgm_fail(XGM_E_INVALID, "null argument");
    Loaded L;
    std::string err;
    if (!load(glass_path, L, err)) return xgm_fail(XGM_E_IO, "%s: %s", glass_path, err.c_str());
    xgm_builder* b = nullptr;
    xgm_status st = xgm_builder_new(&b);
    if (st != XGM_OK) return st;
    st = xgm_builder_set_docs(b, (uint32_t)L.g.doccount, (uint32_t)L.g.lastdocid, L.g.total_doclen, (uint32_t)L.g.doclen_lb,
                              (uint32_t)L.g.doclen_ub, L.doclen.data());
    for (size_t t = 0; st == XGM_OK && t < L.terms.size(); ++t) {
        const Loaded::Term& T = L.terms[t];
        const size_t end = t + 1 < L.terms.size() ? L.terms[t + 1].begin : L.dids.size();
        if (end - T.begin != T.tf) { st = xgm_fail(XGM_E_IO, "term %s: %zu postings, termfreq %llu", T.name.c_str(), end - T.begin, (unsigned long long)T.tf); break; }
        /* the reference's own bound is authoritative even when it is 0: pass it through the "given" path */
        uint32_t id;
        st = xgm_builder_add_term(b, T.name.data(), (uint32_t)T.name.size(), L.dids.data() + T.begin, L.wdfs.data() + T.begin,
                                  (uint32_t)T.tf, T.cf, term_wdf_ub(L, t), &id);
        if (st == XGM_OK && term_wdf_ub(L, t) == 0) st = xgm_builder_force_wdf_ub(b, id, 0);
    }
    for (size_t s = 0; st == XGM_OK && s < L.slots.size() && s < 8; ++s) {
        if (L.slots[s].empty()) continue;
        std::vector<uint64_t> off((size_t)L.g.lastdocid + 2, 0);
        std::string bytes;
        uint32_t next = 0;
        for (const auto& v : L.slots[s]) {
            for (; next <= v.did; ++next) off[next] = bytes.size();
            bytes += v.v;
        }
        for (; next <= L.g.lastdocid + 1; ++next) off[next] = bytes.size();
        bytes.push_back('\0');
        st = xgm_builder_add_value_slot_serialised(b, (uint32_t)s, off.data(), reinterpret_cast<const unsigned char*>(bytes.data()));
    }
    if (st == XGM_OK) st = xgm_builder_set_revision(b, L.g.revision);
    if (st != XGM_OK) { xgm_builder_free(b); return st; }
    return xgm_builder_finish(b, device, out);
}
