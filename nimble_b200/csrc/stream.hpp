// Streaming file pipeline behind nb200_align_files / nb200_align_files_multi (§8e):
//   reader + block-parallel inflate -> record walker (mate pairing) -> parse + 2-bit pack straight into page-locked slabs
//   -> one thread per GPU (two slabs in flight each, slab_api.hpp) -> TSV formatters -> ordered writer.
// Bounded memory: a fixed pool of slabs; every stage hands work on, nothing holds the whole file.
// Reference boundary: what the aligner process does between its argv and its exit code (nimble/__main__.py:177-196).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/nimble_b200.h"

namespace nb200 {

struct FileJob {
    std::vector<std::string> inputs;      // one or two FASTQ(.gz), or one BAM
    std::vector<nb200_ctx *> ctxs;        // one context per GPU; empty = dry run (no device stage, reader statistics only)
    std::vector<int32_t> lib_ids;         // the same ids on every context (libraries loaded in the same order)
    std::vector<std::string> outputs;     // one per library
    int host_threads = 1;
};

struct FileStats {
    uint64_t n_reads = 0, paired = 0, has_tags = 0, bases1 = 0, bases2 = 0, checksum = 0, n_slabs = 0, called = 0;
    double seconds = 0.0;
};

// throws IoError / std::exception
void run_file_pipeline(const FileJob &job, FileStats *stats);

}  // namespace nb200
