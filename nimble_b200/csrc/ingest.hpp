// Native read ingest / TSV output used by nb200_align_files and the `aligner` executable.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nimble_b200.h"

namespace nb200 {

struct IoError : std::runtime_error { using std::runtime_error::runtime_error; };

struct Arena {                 // strings back to back: data + offsets (n + 1)
    std::string data;
    std::vector<int64_t> off{0};
    void add(const char *p, size_t n);
    size_t size() const { return off.size() - 1; }
    const char *ptr(size_t i) const { return data.data() + off[i]; }
    size_t len(size_t i) const { return (size_t)(off[i + 1] - off[i]); }
};

struct ReadSet {
    Arena names, r1, r2, cb, ub, ur, gn;
    std::vector<int64_t> pos1, pos2;
    bool paired = false, has_tags = false;
};

void load_reads(const std::vector<std::string> &inputs, int threads, ReadSet &out);
void write_per_read_tsv(const std::string &out_path, const ReadSet &R, const nb200_read_result *res, const int32_t *feats,
                        int max_hits, const std::vector<std::string> &feature_names);
void write_bulk_tsv(const std::string &out_path, const nb200_counts &c, const std::vector<std::string> &feature_names);

}  // namespace nb200
