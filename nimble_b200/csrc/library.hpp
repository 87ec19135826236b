// Host-side library model: nimble's `[config, data]` JSON -> features -> k-mer index images that
// engine.cu uploads verbatim into HBM.  Reference surface: nimble/types.py:10-32 (schema),
// nimble/__main__.py:64-65 (file layout), nimble/__main__.py:182-189 (`-r LIB.json`).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nimble_b200.h"

namespace nb200 {

constexpr uint32_t kEmptyClass = 0xFFFFFFFFu;
constexpr uint32_t kRefPad = 640;        // invalid bases before/after every reference (>= 500 + band + 64)
constexpr uint32_t kMaxRefs = 65536u * 32u - 32u;  // class words are indexed with 16 bits

struct Slot {            // 32 B = one L2 sector; open addressing, keyed by the CANONICAL k-mer
    uint64_t key;        // min(x, revcomp x); base j at bits [2j, 2j+2); ~0 = empty (never canonical)
    uint32_t cls_s;      // class of references containing `key` itself on their forward strand
    uint32_t off_s;      // index into positions[]: first position in each member of cls_s
    uint32_t cls_r;      // class of references containing revcomp(key) (kEmptyClass = none)
    uint32_t off_r;
    uint32_t pad[2];
};
static_assert(sizeof(Slot) == 32, "slot must be one sector");

// Equivalence class = sparse bitset over references: sorted (word index, 32 member bits) pairs.
// One 32 B record (one sector) holds up to 5 pairs inline; wider classes point into ov_*.
struct ClassRec {
    uint16_t n;          // number of pairs (saturates at 65535 for the overflow form)
    uint16_t w[5];       // inline: word indices (unused = 0 with zero bits)
    uint32_t b[5];       // inline: bits.  overflow (n > 5): b[0] = offset into ov_w/ov_b/ov_pre, b[1] = pairs
};
static_assert(sizeof(ClassRec) == 32, "class record must be one sector");
constexpr uint64_t kEmptyKey = ~0ull;

struct HostLibrary {
    nb200_config cfg{};
    bool has_index = false;
    // features
    std::vector<std::string> feature_names;   // ascending byte order; id = rank
    std::vector<uint32_t> tok_end, tok_comma; // rank of name+NUL / name+',' among all 2F tokens
    // references, INTERNAL order = (feature id, input order): ref -> feature is monotone
    std::vector<std::string> ref_names;
    std::vector<uint32_t> ref_feature;
    std::vector<uint32_t> ref_len, ref_gstart;
    bool identity_features = false;           // feature id == internal ref id
    // index images
    uint32_t n_refs = 0, n_features = 0, n_words = 1;
    uint64_t n_kmers = 0, n_classes = 0, n_slots = 0;
    std::vector<Slot> table;
    std::vector<ClassRec> class_rec;          // n_classes
    std::vector<uint32_t> ov_w, ov_b, ov_pre; // overflow pairs: word index, bits, members before the pair
    std::vector<uint32_t> positions;
    std::vector<uint64_t> ref2bit;            // global coordinate space, 32 bases / word
    std::vector<uint32_t> refN;               // 1 = not ACGT (or padding), 32 bases / word
    uint64_t total_gbases = 0;
};

uint64_t hash_kmer(uint64_t x);
uint64_t revcomp_kmer(uint64_t x, int k);

// throws std::runtime_error (EINVAL-class) / LimitError
struct LimitError : std::runtime_error { using std::runtime_error::runtime_error; };

void parse_library_json(const std::string &path, std::vector<std::string> &names, std::vector<std::string> &seqs,
                        std::vector<std::string> &features, nb200_config &cfg);
void build_library(const std::vector<std::string> &names, const std::vector<std::string> &seqs,
                   const std::vector<std::string> &features, const nb200_config &cfg, int host_threads,
                   HostLibrary &out);
void build_feature_dictionary(const std::vector<std::string> &sorted_names, HostLibrary &out);
int parse_strand_filter(const char *s);

}  // namespace nb200
