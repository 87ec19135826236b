// Read trimming for `--trim <TARGET_LENGTH>:<STRICTNESS>` (nimble/__main__.py:191-192,400; defaults nimble/types.py:24-25).
// The reference hands the flag to the aligner binary; its two parameters are those of Trimmomatic's MAXINFO step, and the
// published MaxInfo criterion (Bolger et al. 2014) is what runs here — DESIGN.md §2.9, restated in oracle/trim_py.py.
// PARITY UNPINNED for this piece (no vector of the real aligner exists); bit-exact against the restatement.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nimble_b200.h"

namespace nb200 {

struct TrimTable {
    bool on = false;
    int target = 0;
    double strictness = 0.0;
    double ls[NB200_MAX_READ_LEN + 1];   // length term of a prefix of i + 1 bases
    double qs[60];                       // quality term of one base
    void init(int t, double s) {
        on = true; target = t; strictness = s;
        for (int i = 0; i <= NB200_MAX_READ_LEN; i++)
            ls[i] = std::log(1.0 / (1.0 + std::exp((double)(t - i - 1)))) + (1.0 - s) * std::log((double)(i + 1));
        for (int q = 0; q < 60; q++) qs[q] = s * std::log(1.0 - std::pow(0.1, (0.5 + q) / 10.0));
    }
    // bases to keep from the 5' end of the read AS SEQUENCED.  qual: n values; offset is subtracted (33 for FASTQ text);
    // reversed: the array is stored 3' -> 5' (BAM records of reverse-strand alignments)
    uint32_t keep(const uint8_t *qual, uint32_t n, int offset, bool reversed) const {
        double acc = 0.0, best = -INFINITY;
        uint32_t pos = 0;
        if (n > (uint32_t)NB200_MAX_READ_LEN) n = NB200_MAX_READ_LEN;
        for (uint32_t i = 0; i < n; i++) {
            int q = (int)qual[reversed ? n - 1 - i : i] - offset;
            q = q < 0 ? 0 : (q > 59 ? 59 : q);
            acc += qs[q];
            const double score = ls[i] + acc;
            if (score >= best) { best = score; pos = i + 1; }
        }
        return pos;
    }
};

// "L:S[,L:S...]" -> one (target, strictness) per library; empty string = no trimming.  Throws on a malformed value or
// an entry count that is not the library count.
inline std::vector<std::pair<int, double>> parse_trim_arg(const char *arg, size_t n_libs) {
    std::vector<std::pair<int, double>> out;
    if (!arg || !*arg) return out;
    const std::string s(arg);
    size_t a = 0;
    for (;;) {
        const size_t e = s.find(',', a);
        const std::string item = s.substr(a, e == std::string::npos ? std::string::npos : e - a);
        const size_t c = item.find(':');
        char *end1 = nullptr, *end2 = nullptr;
        const long t = c == std::string::npos ? -1 : strtol(item.c_str(), &end1, 10);
        const double st = c == std::string::npos ? -1.0 : strtod(item.c_str() + c + 1, &end2);
        if (c == std::string::npos || c == 0 || end1 != item.c_str() + c || !end2 || *end2 != '\0' || end2 == item.c_str() + c + 1 ||
            t < 0 || t > 100000 || !(st >= 0.0 && st <= 1.0))
            throw std::runtime_error("--trim expects <TARGET_LENGTH>:<STRICTNESS> (strictness 0..1), comma-separated, one entry per library: '" + s + "'");
        out.emplace_back((int)t, st);
        if (e == std::string::npos) break;
        a = e + 1;
    }
    if (out.size() != n_libs) throw std::runtime_error("--trim needs one <TARGET_LENGTH>:<STRICTNESS> entry per library");
    return out;
}

}  // namespace nb200
