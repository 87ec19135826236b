// GPU side of the streaming file pipeline (stream.cpp): a context runs up to two slabs (batches of parsed reads) at
// a time, each on its own "lane" of device buffers: H2D of slab k+1 and D2H of slab k-1 overlap the kernels of slab k.
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../include/nimble_b200.h"

namespace nb200 {

constexpr int kFileLanes = 2;

// The calling thread becomes the one that drives this context's device (cudaSetDevice).
void lane_bind_thread(nb200_ctx *c);
// Enqueue one slab: copy the packed reads to the device, align against every library in lib_ids (the reads go up
// once), copy the per-read records / feature ids back into the caller's PINNED buffers (one pair per library).
// Asynchronous; throws std::exception on error.
// len1 / len2 (may be null): per library, the read lengths that library's kernels use instead of r1->len / r2->len
// (--trim keeps a per-library prefix of every read; the packed records are shared).
void lane_submit(nb200_ctx *c, int lane, const nb200_reads *r1, const nb200_reads *r2, const int32_t *lib_ids, int n_libs,
                 nb200_read_result *const *out_res, int32_t *const *out_feats, uint16_t *const *len1 = nullptr,
                 uint16_t *const *len2 = nullptr);
// --trim setting of a library (nb200_library_set_trim); false = none
bool lane_trim(nb200_ctx *c, int32_t lib_id, int *target, double *strictness);
// Block until the slab of this lane is done and its outputs are in the host buffers (re-runs it with a larger
// Smith-Waterman work list if that overflowed).
void lane_wait(nb200_ctx *c, int lane);
int lane_max_hits(nb200_ctx *c, int32_t lib_id);
const char *lane_feature_name(nb200_ctx *c, int32_t lib_id, uint32_t fid, uint32_t *len);
uint32_t lane_n_features(nb200_ctx *c, int32_t lib_id);
int lane_host_threads(nb200_ctx *c);
// page-lock / release an existing host range for every context of the process (cudaHostRegisterPortable); false = not pinned
bool lane_pin(void *p, size_t bytes);
void lane_unpin(void *p);

}  // namespace nb200
