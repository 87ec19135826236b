// Host side of `nimble report` (nimble/__main__.py:213-293): per-read TSV -> rows for the device
// UMI stage (agg.cuh) -> counts TSV.  The arithmetic (merge, thresholding, intersection, counting)
// runs on the GPU; this file only parses, interns strings and writes.
#include "ingest.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string_view>
#include <unordered_map>

namespace nb200 {

void slurp_maybe_gz(const std::string &path, std::string &out);   // ingest.cpp

// what pandas.read_csv turns into NaN by default (check_df_from_input, __main__.py:219)
static bool pandas_na(std::string_view v) {
    static const char *na[] = {"", "#N/A", "#N/A N/A", "#NA", "-1.#IND", "-1.#QNAN", "-NaN", "-nan", "1.#IND", "1.#QNAN", "<NA>",
                               "N/A", "NA", "NULL", "NaN", "None", "n/a", "nan", "null"};
    for (const char *s : na) if (v == s) return true;
    return false;
}

// strings -> dense ids in ascending byte order (pandas groupby order == output order of the counts TSV)
static void intern_sorted(const std::vector<std::string_view> &vals, std::vector<uint32_t> &ids, std::vector<std::string> &uniq) {
    std::unordered_map<std::string_view, uint32_t> seen;
    seen.reserve(vals.size() / 4 + 16);
    std::vector<std::string_view> keys;
    for (auto v : vals) if (seen.emplace(v, 0u).second) keys.push_back(v);
    std::sort(keys.begin(), keys.end());
    uniq.clear();
    uniq.reserve(keys.size());
    for (uint32_t i = 0; i < keys.size(); i++) { seen[keys[i]] = i; uniq.emplace_back(keys[i]); }
    ids.resize(vals.size());
    for (size_t i = 0; i < vals.size(); i++) ids[i] = seen[vals[i]];
}

// returns false when there is nothing to report (empty file / header only / no usable row)
bool parse_per_read_tsv(const std::string &path, ReportRows &R) {
    slurp_maybe_gz(path, R.text);
    const std::string &t = R.text;
    if (t.empty()) return false;
    size_t p = 0;
    auto next_line = [&](std::string_view &line) -> bool {
        if (p >= t.size()) return false;
        size_t e = t.find('\n', p);
        if (e == std::string::npos) e = t.size();
        size_t b = e;
        if (b > p && t[b - 1] == '\r') b--;
        line = std::string_view(t.data() + p, b - p);
        p = e + 1;
        return true;
    };
    auto split = [](std::string_view line, std::vector<std::string_view> &out) {
        out.clear();
        size_t a = 0;
        for (;;) {
            const size_t e = line.find('\t', a);
            if (e == std::string_view::npos) { out.push_back(line.substr(a)); break; }
            out.push_back(line.substr(a, e - a));
            a = e + 1;
        }
    };
    std::string_view line;
    std::vector<std::string_view> f;
    if (!next_line(line)) return false;
    split(line, f);
    int fi = -1, ui = -1, ci = -1, si = -1;
    for (int i = 0; i < (int)f.size(); i++) {
        if (f[i] == "nimble_features") fi = i; else if (f[i] == "r1_UB") ui = i;
        else if (f[i] == "r1_CB") ci = i; else if (f[i] == "nimble_score") si = i;
    }
    if (fi < 0 || ui < 0 || ci < 0 || si < 0) throw std::runtime_error("per-read TSV lacks nimble_features / r1_UB / r1_CB / nimble_score");
    const int need = std::max(std::max(fi, ui), std::max(ci, si));
    std::vector<std::string_view> cbs, umis, feats;
    while (next_line(line)) {
        split(line, f);
        if ((int)f.size() <= need) continue;                 // short row: its missing cells are NaN (dropna, :244)
        if (pandas_na(f[fi]) || pandas_na(f[ui]) || pandas_na(f[ci]) || pandas_na(f[si])) continue;
        std::string num(f[si]);
        char *end = nullptr;
        const double s = strtod(num.c_str(), &end);
        if (end == num.c_str() || *end != '\0' || s != s) continue;
        cbs.push_back(f[ci]); umis.push_back(f[ui]); feats.push_back(f[fi]);
        R.score.push_back(s);
    }
    const size_t n = cbs.size();
    if (!n) return false;
    // feature names: every comma-separated token of every row
    std::vector<std::string_view> toks;
    std::vector<uint32_t> row_off(n + 1, 0);
    for (size_t i = 0; i < n; i++) {
        std::string_view v = feats[i];
        size_t a = 0;
        for (;;) {
            const size_t e = v.find(',', a);
            toks.push_back(v.substr(a, e == std::string_view::npos ? std::string_view::npos : e - a));
            if (e == std::string_view::npos) break;
            a = e + 1;
        }
        row_off[i + 1] = (uint32_t)toks.size();
    }
    std::vector<uint32_t> tok_ids;
    intern_sorted(toks, tok_ids, R.feature_names);
    for (size_t i = 0; i < n; i++) std::sort(tok_ids.begin() + row_off[i], tok_ids.begin() + row_off[i + 1]);   // sorted names (:248)
    R.off = std::move(row_off);
    R.ids = std::move(tok_ids);
    std::vector<uint32_t> cid, uid;
    std::vector<std::string> umi_names;
    intern_sorted(cbs, cid, R.cells);
    intern_sorted(umis, uid, umi_names);
    R.key.resize(n);
    for (size_t i = 0; i < n; i++) R.key[i] = ((uint64_t)cid[i] << 32) | uid[i];
    return true;
}

void write_counts_tsv(const std::string &out_path, const nb200_counts *c, const std::vector<std::string> &feature_names,
                      const std::vector<std::string> &cells) {
    const std::string tmp = out_path + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) throw IoError("cannot write " + tmp);
    std::string buf;
    const uint64_t n = c ? c->n_rows : 0;
    for (uint64_t i = 0; i < n; i++) {
        for (uint32_t j = c->feat_off[i]; j < c->feat_off[i + 1]; j++) { if (j > c->feat_off[i]) buf += ','; buf += feature_names[c->feat_ids[j]]; }
        buf += '\t'; buf += std::to_string(c->count[i]); buf += '\t'; buf += cells[c->cell[i]]; buf += '\n';
        if (buf.size() > (1u << 22)) { if (fwrite(buf.data(), 1, buf.size(), f) != buf.size()) { fclose(f); throw IoError("write failed on " + tmp); } buf.clear(); }
    }
    if (!buf.empty() && fwrite(buf.data(), 1, buf.size(), f) != buf.size()) { fclose(f); throw IoError("write failed on " + tmp); }
    if (fclose(f) != 0) throw IoError("close failed on " + tmp);
    if (rename(tmp.c_str(), out_path.c_str()) != 0) throw IoError("cannot rename " + tmp);
}

}  // namespace nb200
