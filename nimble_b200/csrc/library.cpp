// Library JSON -> host images of the GPU index.
#include "library.hpp"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <thread>
#include <unordered_map>

#include "json.hpp"

namespace nb200 {

uint64_t hash_kmer(uint64_t x) {       // identical to dev_hash_kmer (kernels.cuh)
    x ^= x >> 29;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    return x;
}

uint64_t revcomp_kmer(uint64_t x, int k) {
    uint64_t y = ~x, r = 0;
    for (int i = 0; i < 32; i++) { r = (r << 2) | (y & 3); y >>= 2; }   // reverse the 2-bit groups
    return r >> (64 - 2 * k);
}

int parse_strand_filter(const char *s) {
    if (!s || !*s) return NB200_UNSTRANDED;
    std::string v(s);
    if (v == "unstranded") return NB200_UNSTRANDED;
    if (v == "fiveprime") return NB200_FIVEPRIME;
    if (v == "threeprime") return NB200_THREEPRIME;
    if (v == "none") return NB200_STRAND_NONE;
    return -1;
}

static int cfg_int(const JValue &c, const char *key, int dflt) {
    const JValue *v = c.get(key);
    if (!v) return dflt;
    if (v->type == JValue::Num) return (int)v->num;
    if (v->type == JValue::Bool) return v->b ? 1 : 0;
    return dflt;
}

static double cfg_num(const JValue &c, const char *key, double dflt) {
    const JValue *v = c.get(key);
    return (v && v->type == JValue::Num) ? v->num : dflt;
}

void parse_library_json(const std::string &path, std::vector<std::string> &names, std::vector<std::string> &seqs,
                        std::vector<std::string> &features, nb200_config &cfg) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open library file " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    std::string text = ss.str();
    JValue root = JParser(text.data(), text.size()).parse();
    if (root.type != JValue::Arr || root.arr.size() < 2 || root.arr[0].type != JValue::Obj ||
        root.arr[1].type != JValue::Obj)
        throw std::runtime_error("library JSON must be [config, data] (nimble/__main__.py:64-65)");
    const JValue &c = root.arr[0], &d = root.arr[1];
    // defaults = nimble/types.py:12-25
    cfg.score_threshold = cfg_int(c, "score_threshold", 20);
    cfg.score_filter = cfg_int(c, "score_filter", 25);
    cfg.score_percent = cfg_num(c, "score_percent", 0.5);
    cfg.num_mismatches = cfg_int(c, "num_mismatches", 0);
    cfg.discard_multiple_matches = cfg_int(c, "discard_multiple_matches", 0);
    cfg.intersect_level = cfg_int(c, "intersect_level", 0);
    cfg.discard_multi_hits = cfg_int(c, "discard_multi_hits", 0);
    cfg.require_valid_pair = cfg_int(c, "require_valid_pair", 0);
    cfg.max_hits_to_report = cfg_int(c, "max_hits_to_report", 10);
    std::string group_on;
    if (const JValue *g = c.get("group_on"); g && g->type == JValue::Str) group_on = g->str;

    const JValue *h = d.get("headers"), *cols = d.get("columns");
    if (!h || !cols || h->type != JValue::Arr || cols->type != JValue::Arr || h->arr.size() != cols->arr.size())
        throw std::runtime_error("library data needs matching 'headers' and 'columns' (nimble/types.py:29-32)");
    int i_name = -1, i_seq = -1, i_group = -1;
    for (size_t i = 0; i < h->arr.size(); i++) {
        const std::string &s = h->arr[i].str;
        if (s == "sequence_name") i_name = (int)i;
        else if (s == "sequence") i_seq = (int)i;
        if (!group_on.empty() && s == group_on) i_group = (int)i;
    }
    if (i_name < 0 || i_seq < 0) throw std::runtime_error("library data lacks sequence_name / sequence columns");
    const JValue &cn = cols->arr[i_name], &cs = cols->arr[i_seq];
    if (cn.type != JValue::Arr || cs.type != JValue::Arr || cn.arr.size() != cs.arr.size())
        throw std::runtime_error("library columns have different lengths");
    const JValue *cg = i_group >= 0 ? &cols->arr[i_group] : nullptr;
    if (cg && (cg->type != JValue::Arr || cg->arr.size() != cn.arr.size()))
        throw std::runtime_error("group_on column has a different length");
    size_t n = cn.arr.size();
    names.resize(n); seqs.resize(n); features.resize(n);
    for (size_t i = 0; i < n; i++) {
        names[i] = cn.arr[i].str;
        seqs[i] = cs.arr[i].str;
        features[i] = cg ? cg->arr[i].str : names[i];
    }
}

static void token_ranks(const std::vector<std::string> &names, std::vector<uint32_t> &te, std::vector<uint32_t> &tc) {
    // Order of the comma-joined feature strings pandas sorts by (nimble/__main__.py:248-251,289)
    // == lexicographic order of token sequences, token = name + ',' (inner) or name + NUL (last).
    size_t F = names.size();
    struct Tok { std::string s; uint32_t id; uint8_t kind; };
    std::vector<Tok> toks;
    toks.reserve(2 * F);
    for (size_t i = 0; i < F; i++) {
        toks.push_back({names[i] + std::string(1, '\0'), (uint32_t)i, 0});
        toks.push_back({names[i] + ",", (uint32_t)i, 1});
    }
    std::sort(toks.begin(), toks.end(), [](const Tok &a, const Tok &b) {
        int c = memcmp(a.s.data(), b.s.data(), std::min(a.s.size(), b.s.size()));
        if (c) return c < 0;
        return a.s.size() < b.s.size();
    });
    te.assign(F, 0); tc.assign(F, 0);
    for (size_t r = 0; r < toks.size(); r++) (toks[r].kind ? tc : te)[toks[r].id] = (uint32_t)r;
}

static bool bytes_less(const std::string &a, const std::string &b) {
    int c = memcmp(a.data(), b.data(), std::min(a.size(), b.size()));
    if (c) return c < 0;
    return a.size() < b.size();
}

void build_feature_dictionary(const std::vector<std::string> &sorted_names, HostLibrary &out) {
    for (size_t i = 0; i < sorted_names.size(); i++) {
        const std::string &s = sorted_names[i];
        if (s.find(',') != std::string::npos || s.find('\t') != std::string::npos || s.find('\n') != std::string::npos ||
            s.find('\0') != std::string::npos)
            throw std::runtime_error("feature name contains ',', TAB, NUL or newline: " + s);
        if (i && !bytes_less(sorted_names[i - 1], s)) throw std::runtime_error("feature names must be sorted and unique");
    }
    out.feature_names = sorted_names;
    out.n_features = (uint32_t)sorted_names.size();
    token_ranks(sorted_names, out.tok_end, out.tok_comma);
}

static inline int base_code(char c) {
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

struct Occ { uint64_t kmer; uint32_t ref, pos; };

void build_library(const std::vector<std::string> &names, const std::vector<std::string> &seqs,
                   const std::vector<std::string> &features, const nb200_config &cfg, int host_threads,
                   HostLibrary &L) {
    (void)host_threads;
    const int k = cfg.k;
    if (k < 4 || k > 32) throw std::runtime_error("k must be in 4..32");
    if (cfg.max_hits_to_report < 1 || cfg.max_hits_to_report > 64)
        throw LimitError("max_hits_to_report must be in 1..64");
    if (cfg.strand_filter < 0 || cfg.strand_filter > 3) throw std::runtime_error("bad strand_filter");
    const size_t R = names.size();
    if (seqs.size() != R || features.size() != R) throw std::runtime_error("names/seqs/features differ in length");
    if (R == 0) throw std::runtime_error("library has no sequences");
    if (R > kMaxRefs) throw LimitError("library has more than 2,097,120 sequences (16-bit class word index)");
    L.cfg = cfg;
    // ---- features: id = rank of the name in byte order ------------------------------------------
    std::vector<std::string> fn(features);
    std::sort(fn.begin(), fn.end(), bytes_less);
    fn.erase(std::unique(fn.begin(), fn.end()), fn.end());
    build_feature_dictionary(fn, L);
    std::vector<uint32_t> fid(R);
    for (size_t i = 0; i < R; i++)
        fid[i] = (uint32_t)(std::lower_bound(fn.begin(), fn.end(), features[i], bytes_less) - fn.begin());
    // ---- internal reference order: (feature id, input order) ------------------------------------
    std::vector<uint32_t> order(R);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return fid[a] < fid[b]; });
    L.n_refs = (uint32_t)R;
    L.ref_names.resize(R); L.ref_feature.resize(R); L.ref_len.resize(R); L.ref_gstart.resize(R);
    L.identity_features = (L.n_features == R);
    uint64_t g = kRefPad;
    for (size_t r = 0; r < R; r++) {
        uint32_t src = order[r];
        L.ref_names[r] = names[src];
        L.ref_feature[r] = fid[src];
        if (fid[src] != r) L.identity_features = false;
        L.ref_len[r] = (uint32_t)seqs[src].size();
        if (g + seqs[src].size() + kRefPad + 128 >= 0xFFFFFFFFull) throw LimitError("library exceeds 4 Gbases");
        L.ref_gstart[r] = (uint32_t)g;
        g += seqs[src].size() + kRefPad;
    }
    L.total_gbases = g + 128;
    const size_t n_words = (size_t)((L.total_gbases + 31) / 32) + 2;
    L.ref2bit.assign(n_words, 0);
    L.refN.assign(n_words, 0xFFFFFFFFu);
    // ---- k-mer occurrences -------------------------------------------------------------------
    std::vector<Occ> occ;
    {
        size_t tot = 0;
        for (size_t r = 0; r < R; r++) if (L.ref_len[r] >= (uint32_t)k) tot += L.ref_len[r] - k + 1;
        occ.reserve(tot);
    }
    const uint64_t kmask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    for (size_t r = 0; r < R; r++) {
        const std::string &s = seqs[order[r]];
        uint64_t x = 0; int valid = 0;
        const uint64_t g0 = L.ref_gstart[r];
        for (size_t i = 0; i < s.size(); i++) {
            int c = base_code(s[i]);
            if (c > 3) { valid = 0; x = 0; continue; }
            uint64_t gp = g0 + i;
            L.ref2bit[gp >> 5] |= (uint64_t)c << (2 * (gp & 31));
            L.refN[gp >> 5] &= ~(1u << (gp & 31));
            x = ((x >> 2) | ((uint64_t)c << (2 * (k - 1)))) & kmask;
            if (++valid >= k) occ.push_back({x, (uint32_t)r, (uint32_t)(i + 1 - k)});
        }
    }
    std::sort(occ.begin(), occ.end(), [](const Occ &a, const Occ &b) {
        if (a.kmer != b.kmer) return a.kmer < b.kmer;
        if (a.ref != b.ref) return a.ref < b.ref;
        return a.pos < b.pos;
    });
    // ---- distinct k-mers -> member lists (ascending ref) + first positions ---------------------
    std::vector<uint64_t> kmers;
    std::vector<uint64_t> mem_off;     // per k-mer, into mem_ref / positions
    std::vector<uint32_t> mem_ref;
    L.positions.clear();
    for (size_t i = 0; i < occ.size(); i++) {
        bool newk = (i == 0 || occ[i].kmer != occ[i - 1].kmer);
        if (newk) { kmers.push_back(occ[i].kmer); mem_off.push_back(mem_ref.size()); }
        if (newk || occ[i].ref != occ[i - 1].ref) { mem_ref.push_back(occ[i].ref); L.positions.push_back(occ[i].pos); }
    }
    mem_off.push_back(mem_ref.size());
    if (mem_ref.size() >= 0xFFFFFFFFull) throw LimitError("more than 4 G k-mer occurrences");
    std::vector<Occ>().swap(occ);
    L.n_kmers = kmers.size();
    // ---- equivalence classes = distinct member lists --------------------------------------------
    std::vector<uint32_t> kclass(kmers.size());
    std::vector<uint64_t> class_rep;   // representative k-mer index per class
    {
        std::unordered_map<uint64_t, std::vector<uint32_t>> by_hash;
        by_hash.reserve(kmers.size() / 4 + 16);
        for (size_t q = 0; q < kmers.size(); q++) {
            uint64_t a = mem_off[q], b = mem_off[q + 1];
            uint64_t h = 0x9E3779B97F4A7C15ull ^ (b - a);
            for (uint64_t j = a; j < b; j++) h = hash_kmer(h ^ mem_ref[j]);
            auto &bucket = by_hash[h];
            uint32_t found = kEmptyClass;
            for (uint32_t c : bucket) {
                uint64_t ra = mem_off[class_rep[c]], rb = mem_off[class_rep[c] + 1];
                if (rb - ra == b - a && memcmp(&mem_ref[ra], &mem_ref[a], (b - a) * sizeof(uint32_t)) == 0) { found = c; break; }
            }
            if (found == kEmptyClass) {
                found = (uint32_t)class_rep.size();
                class_rep.push_back(q);
                bucket.push_back(found);
            }
            kclass[q] = found;
        }
    }
    L.n_classes = class_rep.size();
    L.n_words = (uint32_t)((R + 31) / 32);
    L.class_rec.assign(L.n_classes, ClassRec{});
    L.ov_w.clear(); L.ov_b.clear(); L.ov_pre.clear();
    for (size_t c = 0; c < class_rep.size(); c++) {
        std::vector<std::pair<uint32_t, uint32_t>> pairs;     // members are ascending
        for (uint64_t j = mem_off[class_rep[c]]; j < mem_off[class_rep[c] + 1]; j++) {
            const uint32_t w = mem_ref[j] >> 5, bit = 1u << (mem_ref[j] & 31);
            if (pairs.empty() || pairs.back().first != w) pairs.emplace_back(w, bit);
            else pairs.back().second |= bit;
        }
        ClassRec &rec = L.class_rec[c];
        if (pairs.size() <= 5) {
            rec.n = (uint16_t)pairs.size();
            for (size_t i = 0; i < pairs.size(); i++) { rec.w[i] = (uint16_t)pairs[i].first; rec.b[i] = pairs[i].second; }
        } else {
            if (L.ov_w.size() + pairs.size() >= 0xFFFFFFFFull) throw LimitError("class overflow table exceeds 4 G pairs");
            rec.n = (uint16_t)std::min<size_t>(pairs.size(), 65535);
            rec.b[0] = (uint32_t)L.ov_w.size();
            rec.b[1] = (uint32_t)pairs.size();
            uint32_t pre = 0;
            for (auto &pr : pairs) {
                L.ov_w.push_back(pr.first); L.ov_b.push_back(pr.second); L.ov_pre.push_back(pre);
                pre += (uint32_t)__builtin_popcount(pr.second);
            }
        }
    }
    // ---- canonical open-addressing table: one 32 B slot answers both read orientations ----------
    struct Info { uint32_t cls, off; };
    auto info_of = [&](uint64_t x, Info &out) -> bool {
        auto it = std::lower_bound(kmers.begin(), kmers.end(), x);
        if (it == kmers.end() || *it != x) return false;
        size_t q = (size_t)(it - kmers.begin());
        out = Info{kclass[q], (uint32_t)mem_off[q]};
        return true;
    };
    std::vector<Slot> entries;
    entries.reserve(kmers.size());
    for (size_t q = 0; q < kmers.size(); q++) {
        const uint64_t x = kmers[q], y = revcomp_kmer(x, k);
        Slot e{};
        e.cls_s = e.cls_r = kEmptyClass;
        if (x <= y) {                       // x is canonical (or a palindrome)
            e.key = x; e.cls_s = kclass[q]; e.off_s = (uint32_t)mem_off[q];
            Info o;
            if (y != x && info_of(y, o)) { e.cls_r = o.cls; e.off_r = o.off; }
        } else {
            Info o;
            if (info_of(y, o)) continue;    // the canonical partner is in the index: emitted from there
            e.key = y; e.cls_r = kclass[q]; e.off_r = (uint32_t)mem_off[q];
        }
        entries.push_back(e);
    }
    // load factor 0.2..0.4 while the table stays small (L2-resident), <= 0.6 once it is HBM-sized
    uint64_t slots = 1024;
    while (slots * 2 < 5 * entries.size()) slots <<= 1;            // LF <= 0.4
    if (slots * sizeof(Slot) > (1ull << 30)) { slots = 1024; while (slots * 3 < 5 * entries.size()) slots <<= 1; }
    L.n_slots = slots;
    Slot empty{};
    empty.key = kEmptyKey; empty.cls_s = empty.cls_r = kEmptyClass;
    L.table.assign(slots, empty);
    for (const Slot &e : entries) {
        uint64_t s = hash_kmer(e.key) & (slots - 1);
        while (L.table[s].key != kEmptyKey) s = (s + 1) & (slots - 1);
        L.table[s] = e;
    }
    L.has_index = true;
}

}  // namespace nb200
