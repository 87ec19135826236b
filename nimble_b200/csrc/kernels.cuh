// sm_100a kernels of the read-assignment hot path.
// Semantics: DESIGN.md §2 (SPEC); bit-exact against oracle/nimble_oracle.c.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/nimble_b200.h"

namespace nb200 {

constexpr int kBand = 8;                 // SW half band
constexpr int kNB = 2 * kBand + 1;       // cells per row
constexpr int kVW = 64;                  // V = 64*score - edits
constexpr uint32_t kXP = 0xFF7FFF7Fu;    // packed s16x2 (-129,-129) mismatch
constexpr uint32_t kGP = 0xFF3FFF3Fu;    // packed s16x2 (-193,-193) gap
constexpr uint32_t kMatchDelta = 193;    // match - mismatch = 64 + 129
constexpr uint32_t kInvalid = 0xFFFFFFFFu;

struct LibDev {
    const uint4 *table;        // Slot[n_slots], 32 B each (two uint4)
    uint64_t tmask;
    const uint32_t *class_bits;  // (n_classes + 1) rows of wpad words; last row = universe
    const uint32_t *positions;
    const uint64_t *ref2bit;
    const uint32_t *refN;
    const uint32_t *ref_gstart;
    const uint32_t *ref_feature;
    uint32_t wpad, n_refs, n_features, n_classes;
    int32_t k, identity;
};

struct ReadsDev {
    const uint8_t *packed;
    const uint16_t *len;
    uint32_t stride, words;
};

struct __align__(16) RoRec {     // one mate in one orientation, written by the probe kernel for DEFERRED reads
    uint32_t ncand;              // |B| (0 = empty intersection or no hit)
    uint32_t item_off;           // first SW item (kInvalid when none)
    uint16_t n_hits;
    uint16_t len;
    uint16_t seed_i;
    uint8_t full;                // every k-mer position hit -> no SW
    uint8_t pad;
};

struct __align__(16) SwItem {
    uint32_t ro;                 // orientation record index (deferred slot * n_ro + orientation)
    uint32_t ref;                // candidate reference (kInvalid = padding)
    uint32_t gwin;               // global coordinate of band cell (row 0, b 0)
    uint32_t v;                  // out: best V
};

struct CallParams {
    int32_t score_threshold, score_filter, num_mismatches, discard_multiple_matches, intersect_level,
        discard_multi_hits, require_valid_pair, max_hits, strand_filter;
    double score_percent;
};

constexpr int kCtrSpread = 64;   // statistics counters are spread over 64 slots to avoid same-address REDs
struct Counters {
    // alloc = (deferred reads << 40) | SW items : one atomic hands out both cursors
    unsigned long long alloc, overflow, dropped_empty, max_nf, sw_pairs, sw_cells, items_max, deferred_total;
    unsigned long long probes[kCtrSpread], probe_slots[kCtrSpread];
};
constexpr unsigned long long kItemMask = (1ull << 40) - 1;

__device__ __forceinline__ uint64_t dev_hash_kmer(uint64_t x) {   // identical to hash_kmer (library.cpp)
    x ^= x >> 29;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    return x;
}

__device__ __forceinline__ uint64_t dev_revcomp(uint64_t x, int k) {
    uint64_t r = __brevll(~x);
    r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
    return r >> (64 - 2 * k);
}

// one 32 B slot = one L2 sector, fetched with a single 256-bit load that bypasses L1
__device__ __forceinline__ void ldg_slot(const uint4 *p, uint4 &lo, uint4 &hi) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(p));
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
__device__ __forceinline__ uint32_t warp_max(uint32_t v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}
__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, int lane, uint32_t &total) {
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += t;
    }
    total = __shfl_sync(0xFFFFFFFFu, inc, 31);
    return inc - v;
}

enum { ST_NONE = 0, ST_PASS = 1, ST_NO_MATCH = 2, ST_EMPTY = 3, ST_SCORE = 4, ST_PERCENT = 5, ST_MULTI = 6 };
enum { RS_CALLED = 0, RS_NO_PASS = 1, RS_NOT_VALID_PAIR = 2, RS_FORCE_INTERSECT = 3, RS_SCORE_FILTER = 4,
       RS_MULTI_HITS = 5, RS_MAX_HITS = 6 };

// Everything the score/feature stage needs about one read (pair): 4 orientations.
template <int WPL>
struct ReadState {
    uint32_t cls[4][WPL];       // B' per orientation (lane holds words j*32+lane)
    uint32_t nc[4], nh[4];
    int vbest[4];               // -1: no class (no hit / empty)
    int len[4];
    int n_sw;
};

// ---------------------------------------------------------------------------------------------
// X4: score / strand / pair filter + feature calling.  Warp-cooperative.
// ---------------------------------------------------------------------------------------------
template <int WPL>
__device__ __forceinline__ void call_read(const LibDev &lib, const CallParams &cp, bool paired, ReadState<WPL> &S,
                                          uint32_t *sb, int lane, nb200_read_result *res_out, int32_t *fout,
                                          uint16_t *nf_out, Counters *ctr) {
    int st[4], sc[4], ed[4];
#pragma unroll
    for (int o = 0; o < 4; o++) {
        st[o] = ST_NONE; sc[o] = 0; ed[o] = 0;
        if (!paired && o >= 2) continue;
        if (S.nh[o] == 0) { st[o] = ST_NO_MATCH; continue; }
        if (S.vbest[o] < 0) { st[o] = ST_EMPTY; continue; }
        sc[o] = (S.vbest[o] + kVW - 1) / kVW;
        ed[o] = sc[o] * kVW - S.vbest[o];
        if (sc[o] < cp.score_threshold) st[o] = ST_SCORE;
        else if ((double)sc[o] / (double)S.len[o] < cp.score_percent) st[o] = ST_PERCENT;
        else if (cp.discard_multiple_matches && S.nc[o] > 1) st[o] = ST_MULTI;
        else st[o] = ST_PASS;
    }
    int order[4], n_cfg = 0;
    switch (cp.strand_filter) {
    case NB200_FIVEPRIME: order[n_cfg++] = 0; break;
    case NB200_THREEPRIME: order[n_cfg++] = 1; break;
    case NB200_STRAND_NONE: order[n_cfg++] = 0; order[n_cfg++] = 1; if (paired) { order[n_cfg++] = 2; order[n_cfg++] = 3; } break;
    default: order[n_cfg++] = 0; order[n_cfg++] = 1; break;
    }
    uint32_t bestc[WPL];
#pragma unroll
    for (int j = 0; j < WPL; j++) bestc[j] = 0;
    int chosen = -1, chosen_max = 0, first_fail = RS_NO_PASS;
    uint32_t chosen_score = 0;
#pragma unroll
    for (int ci = 0; ci < 4; ci++) {
        if (ci >= n_cfg) continue;
        const int c = order[ci];
        const int ia = (c == 0 || c == 2) ? 0 : 1;            // F,FF: r1 fwd ; R,RR: r1 rc
        const int ib = (c == 0 || c == 3) ? 3 : 2;            // F,RR: r2 rc  ; R,FF: r2 fwd
        const int sa = ia == 0 ? sc[0] : sc[1], sbv = ib == 3 ? sc[3] : sc[2];
        const bool pa = (ia == 0 ? st[0] : st[1]) == ST_PASS, pb = paired && (ib == 3 ? st[3] : st[2]) == ST_PASS;
        int fail = -1, maxmate = 0;
        uint32_t s = 0;
        uint32_t tmp[WPL], A[WPL], B[WPL];
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            tmp[j] = 0;
            A[j] = ia == 0 ? S.cls[0][j] : S.cls[1][j];
            B[j] = ib == 3 ? S.cls[3][j] : S.cls[2][j];
        }
        if (!paired) {
            if (!pa) fail = RS_NO_PASS;
            else {
#pragma unroll
                for (int j = 0; j < WPL; j++) tmp[j] = A[j];
                s = (uint32_t)sa; maxmate = sa;
            }
        } else if (cp.require_valid_pair && !(pa && pb)) fail = RS_NOT_VALID_PAIR;
        else if (!pa && !pb) fail = RS_NO_PASS;
        else if (pa && pb) {
            s = (uint32_t)(sa + sbv); maxmate = max(sa, sbv);
            if (cp.intersect_level == 0) {
#pragma unroll
                for (int j = 0; j < WPL; j++) tmp[j] = A[j] | B[j];
            } else {
                uint32_t any = 0;
#pragma unroll
                for (int j = 0; j < WPL; j++) { tmp[j] = A[j] & B[j]; any |= tmp[j]; }
                if (!__any_sync(0xFFFFFFFFu, any != 0)) {
                    if (cp.intersect_level >= 2) fail = RS_FORCE_INTERSECT;
                    else {
                        const bool useB = sbv > sa;
#pragma unroll
                        for (int j = 0; j < WPL; j++) tmp[j] = useB ? B[j] : A[j];
                    }
                }
            }
        } else {
            if (cp.intersect_level >= 2) fail = RS_FORCE_INTERSECT;
            else {
#pragma unroll
                for (int j = 0; j < WPL; j++) tmp[j] = pa ? A[j] : B[j];
                s = (uint32_t)(pa ? sa : sbv); maxmate = (int)s;
            }
        }
        if (fail >= 0) { if (ci == 0) first_fail = fail; continue; }
        if (chosen < 0 || s > chosen_score) {
            chosen = c; chosen_score = s; chosen_max = maxmate;
#pragma unroll
            for (int j = 0; j < WPL; j++) bestc[j] = tmp[j];
        }
    }
    int reason = first_fail, n_feat = 0;
    const int mh = cp.max_hits;
    if (chosen >= 0) {
        if (chosen_max < cp.score_filter) reason = RS_SCORE_FILTER;
        else {
            uint32_t fw[WPL];
            if (lib.identity) {
#pragma unroll
                for (int j = 0; j < WPL; j++) fw[j] = bestc[j];
            } else {
                for (uint32_t j = lane; j < lib.wpad; j += 32) sb[j] = 0;
                __syncwarp();
#pragma unroll
                for (int j = 0; j < WPL; j++) {
                    uint32_t bits = bestc[j];
                    while (bits) {
                        const int b = __ffs(bits) - 1;
                        bits &= bits - 1;
                        const uint32_t f = __ldg(lib.ref_feature + (uint32_t)((j * 32 + lane) * 32 + b));
                        atomicOr(&sb[f >> 5], 1u << (f & 31));
                    }
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < WPL; j++) fw[j] = sb[j * 32 + lane];
                __syncwarp();
            }
            uint32_t cnt = 0;
#pragma unroll
            for (int j = 0; j < WPL; j++) cnt += __popc(fw[j]);
            const int nf = (int)warp_sum(cnt);
            if (cp.discard_multi_hits > 0 && nf > cp.discard_multi_hits) reason = RS_MULTI_HITS;
            else if (nf > mh) reason = RS_MAX_HITS;
            else {
                reason = RS_CALLED; n_feat = nf;
                uint32_t rowoff = 0;
#pragma unroll
                for (int j = 0; j < WPL; j++) {
                    uint32_t tot;
                    uint32_t ex = warp_excl_scan(__popc(fw[j]), lane, tot);
                    uint32_t bits = fw[j];
                    while (bits) {
                        const int b = __ffs(bits) - 1;
                        bits &= bits - 1;
                        fout[rowoff + ex++] = (int32_t)((j * 32 + lane) * 32 + b);
                    }
                    rowoff += tot;
                }
            }
        }
    }
    for (int t = n_feat + lane; t < mh; t += 32) fout[t] = -1;
    if (lane == 0) {
        nb200_read_result res;
#pragma unroll
        for (int o = 0; o < 4; o++) {
            res.score[o] = (uint16_t)sc[o]; res.n_hits[o] = (uint16_t)S.nh[o];
            res.n_cand[o] = (uint16_t)(S.nc[o] > 65535u ? 65535u : S.nc[o]);
            res.edits[o] = (uint8_t)ed[o]; res.status[o] = (uint8_t)st[o];
        }
        res.reason = (uint8_t)reason; res.config = (uint8_t)(chosen < 0 ? 255 : chosen);
        res.n_feat = (uint8_t)n_feat; res.n_sw = (uint8_t)S.n_sw;
        res.pair_score = chosen < 0 ? 0u : chosen_score;
        *res_out = res;
        *nf_out = (uint16_t)n_feat;
        // the aggregation only needs the batch maximum: skip the atomic once it is already there
        if (n_feat && (unsigned long long)n_feat > *(volatile unsigned long long *)&ctr->max_nf)
            atomicMax(&ctr->max_nf, (unsigned long long)n_feat);
    }
}

// Per-mate probe result kept in registers by the fused kernel.
template <int WPL>
struct MateProbe {
    uint32_t acc[2][WPL];
    uint32_t nh[2], seed_cls[2], seed_off[2];
    int seed_i[2];              // position of the seed k-mer in the ORIENTED read
    int L, P;
};

// X3a + X3b for one mate: every k-mer position, ONE probe of the canonical table serves both
// orientations; equivalence classes are ANDed four rows at a time (independent loads in flight).
template <int WPL>
__device__ __forceinline__ void probe_mate(const LibDev &lib, const ReadsDev &R, uint64_t read, int lane,
                                           MateProbe<WPL> &M, uint32_t &n_probe, uint32_t &slots_read) {
    const uint8_t *rec = R.packed + read * R.stride;
    const uint64_t *seq = reinterpret_cast<const uint64_t *>(rec);
    const uint32_t *nm = reinterpret_cast<const uint32_t *>(rec + (size_t)R.words * 8);
    const int L = R.len[read];
    const int k = lib.k;
    const int P = L - k + 1;
    M.L = L; M.P = P;
    const uint64_t kmask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    const uint64_t kbits = k == 32 ? 0xFFFFFFFFull : ((1ull << k) - 1);
#pragma unroll
    for (int o = 0; o < 2; o++) {
#pragma unroll
        for (int j = 0; j < WPL; j++) M.acc[o][j] = 0xFFFFFFFFu;
        M.nh[o] = 0; M.seed_cls[o] = 0; M.seed_off[o] = 0; M.seed_i[o] = -1;
    }
    uint32_t last[2] = {kInvalid, kInvalid}, anded[2] = {kInvalid, kInvalid};
    bool dead[2] = {false, false};          // intersection already empty: stop fetching rows

    for (int base = 0; base < P; base += 32) {
        const int i = base + lane;
        const int w = base >> 5;
        const uint64_t s0 = seq[w];
        const uint64_t s1 = (w + 1 < (int)R.words) ? seq[w + 1] : 0ull;
        const uint64_t m01 = (uint64_t)nm[w] | ((w + 1 < (int)R.words) ? ((uint64_t)nm[w + 1] << 32) : 0ull);
        const int sh = lane * 2;
        uint64_t x = sh ? ((s0 >> sh) | (s1 << (64 - sh))) : s0;
        x &= kmask;
        const bool valid = (i < P) && (((m01 >> lane) & kbits) == 0);
        const uint64_t y = dev_revcomp(x, k);
        const uint64_t c = x < y ? x : y;
        uint64_t slot = dev_hash_kmer(c) & lib.tmask;
        uint32_t cl[2] = {kInvalid, kInvalid}, of[2] = {0, 0};
        bool done = !valid;
        if (valid) n_probe++;
        while (!done) {
            uint4 lo, hi;
            ldg_slot(lib.table + 2 * slot, lo, hi);
            slots_read++;
            const uint64_t key = ((uint64_t)lo.y << 32) | lo.x;
            if (key == c) {
                // forward read k-mer x: canonical -> "same strand" info, else the revcomp info
                const bool xs = (x == c), ys = (y == c);
                cl[0] = xs ? lo.z : hi.x; of[0] = xs ? lo.w : hi.y;
                cl[1] = ys ? lo.z : hi.x; of[1] = ys ? lo.w : hi.y;
                done = true;
            } else if (key == ~0ull) done = true;
            else slot = (slot + 1) & lib.tmask;
        }
#pragma unroll
        for (int o = 0; o < 2; o++) {
            const bool hit = cl[o] != kInvalid;
            const unsigned hb = __ballot_sync(0xFFFFFFFFu, hit);
            if (hb) {
                M.nh[o] += __popc(hb);
                // orientation 0 reads left to right: its seed is the FIRST hit; the reverse
                // complement visits positions right to left: its first hit is the LAST one here
                if (o == 0) {
                    if (M.seed_i[0] < 0) {
                        const int src = __ffs(hb) - 1;
                        M.seed_i[0] = base + src;
                        M.seed_cls[0] = __shfl_sync(0xFFFFFFFFu, cl[0], src);
                        M.seed_off[0] = __shfl_sync(0xFFFFFFFFu, of[0], src);
                    }
                } else {
                    const int src = 31 - __clz(hb);
                    M.seed_i[1] = P - 1 - (base + src);
                    M.seed_cls[1] = __shfl_sync(0xFFFFFFFFu, cl[1], src);
                    M.seed_off[1] = __shfl_sync(0xFFFFFFFFu, of[1], src);
                }
            }
            uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, cl[o], 1);
            if (lane == 0) prev = last[o];
            last[o] = __shfl_sync(0xFFFFFFFFu, cl[o], 31);
            unsigned nb = __ballot_sync(0xFFFFFFFFu, hit && cl[o] != prev);
            while (nb && !dead[o]) {
                // up to four class rows per trip, all loads issued before the ANDs
                uint32_t cid[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (nb) {
                        const int src = __ffs(nb) - 1;
                        nb &= nb - 1;
                        cid[u] = __shfl_sync(0xFFFFFFFFu, cl[o], src);
                        if (cid[u] == anded[o]) cid[u] = lib.n_classes;      // immediate repeat -> universe row
                        else anded[o] = cid[u];
                    } else cid[u] = lib.n_classes;
                }
                uint32_t rows[4][WPL];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t *row = lib.class_bits + (size_t)cid[u] * lib.wpad + lane;
#pragma unroll
                    for (int j = 0; j < WPL; j++) rows[u][j] = __ldg(row + j * 32);
                }
                uint32_t any = 0;
#pragma unroll
                for (int j = 0; j < WPL; j++) {
                    M.acc[o][j] &= rows[0][j] & rows[1][j] & rows[2][j] & rows[3][j];
                    any |= M.acc[o][j];
                }
                if (!__any_sync(0xFFFFFFFFu, any != 0)) dead[o] = true;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fused X3a/X3b/X4 kernel: one warp per read (pair).  Reads whose orientations all resolve without
// Smith-Waterman are called right here; the others emit SW work items, park their state and are
// finished by call_deferred_kernel after sw_kernel.
// ---------------------------------------------------------------------------------------------
template <int WPL>
__global__ void __launch_bounds__(256)
probe_kernel(LibDev lib, CallParams cp, ReadsDev r1, ReadsDev r2, uint64_t read0, uint64_t n_reads, int n_mates,
             RoRec *__restrict__ ro, uint32_t *__restrict__ roB, uint32_t *__restrict__ deferred,
             SwItem *__restrict__ items, uint32_t items_cap,
             nb200_read_result *__restrict__ results, int32_t *__restrict__ feats, uint16_t *__restrict__ row_nf,
             Counters *__restrict__ ctr) {
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t *sb = smem + (size_t)wib * lib.wpad;
    const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gw >= n_reads) return;
    const uint64_t read = read0 + gw;
    const bool paired = n_mates == 2;
    const int n_ro = n_mates * 2;

    ReadState<WPL> S;
    S.n_sw = 0;
    uint32_t n_probe = 0, slots_read = 0;
    uint32_t seed_cls[4], seed_off[4];
    int seed_i[4];
    bool partial[4];
    uint32_t n_items = 0;
#pragma unroll
    for (int m = 0; m < 2; m++) {
        if (m >= n_mates) {
#pragma unroll
            for (int o = 0; o < 2; o++) {
                const int q = m * 2 + o;
#pragma unroll
                for (int j = 0; j < WPL; j++) S.cls[q][j] = 0;
                S.nc[q] = 0; S.nh[q] = 0; S.vbest[q] = -1; S.len[q] = 0; partial[q] = false;
                seed_cls[q] = 0; seed_off[q] = 0; seed_i[q] = 0;
            }
            continue;
        }
        MateProbe<WPL> M;
        probe_mate<WPL>(lib, m ? r2 : r1, read, lane, M, n_probe, slots_read);
#pragma unroll
        for (int o = 0; o < 2; o++) {
            const int q = m * 2 + o;
            uint32_t cnt = 0;
            if (M.nh[o]) {
#pragma unroll
                for (int j = 0; j < WPL; j++) cnt += __popc(M.acc[o][j]);
                cnt = warp_sum(cnt);
            }
#pragma unroll
            for (int j = 0; j < WPL; j++) S.cls[q][j] = cnt ? M.acc[o][j] : 0u;
            S.nc[q] = cnt; S.nh[q] = M.nh[o]; S.len[q] = M.L;
            const bool full = M.nh[o] && (int)M.nh[o] == M.P;
            S.vbest[q] = cnt ? (full ? M.L * kVW : 0) : -1;
            partial[q] = cnt && !full;
            seed_cls[q] = M.seed_cls[o]; seed_off[q] = M.seed_off[o]; seed_i[q] = M.seed_i[o];
            if (partial[q]) n_items += (cnt + 1) & ~1u;
        }
    }
    n_probe = warp_sum(n_probe);
    slots_read = warp_sum(slots_read);
    if (lane == 0 && n_probe) {
        atomicAdd(&ctr->probes[blockIdx.x & (kCtrSpread - 1)], (unsigned long long)n_probe);
        atomicAdd(&ctr->probe_slots[blockIdx.x & (kCtrSpread - 1)], (unsigned long long)slots_read);
    }
    if (n_items == 0) {   // fast path: nothing to align, call the read now
        call_read<WPL>(lib, cp, paired, S, sb, lane, results + gw, feats + gw * cp.max_hits, row_nf + gw, ctr);
        return;
    }
    // ---- deferred: one atomic hands out the deferred slot and the SW item range -------------------
    unsigned long long a = 0;
    if (lane == 0) a = atomicAdd(&ctr->alloc, (1ull << 40) | (unsigned long long)n_items);
    a = __shfl_sync(0xFFFFFFFFu, a, 0);
    const uint32_t dslot = (uint32_t)(a >> 40);
    uint32_t off = (uint32_t)(a & kItemMask);
    const bool fits = (a & kItemMask) + n_items <= items_cap;
    if (!fits && lane == 0) atomicAdd(&ctr->overflow, 1ull);
    if (lane == 0) deferred[dslot] = (uint32_t)gw;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        if (q >= n_ro) continue;
        const uint64_t ro_idx = (uint64_t)dslot * n_ro + q;
        RoRec rr;
        rr.ncand = S.nc[q]; rr.item_off = kInvalid; rr.n_hits = (uint16_t)S.nh[q]; rr.len = (uint16_t)S.len[q];
        rr.seed_i = (uint16_t)(S.nh[q] ? seed_i[q] : 0); rr.full = (S.nc[q] && !partial[q]) ? 1 : 0; rr.pad = 0;
        if (S.nc[q] && !partial[q]) {         // resolved orientation: park its class for the deferred call
            uint32_t *dst = roB + ro_idx * lib.wpad + lane;
#pragma unroll
            for (int j = 0; j < WPL; j++) dst[j * 32] = S.cls[q][j];
        }
        if (partial[q]) {
            const uint32_t cnt = S.nc[q];
            if (fits) {
                rr.item_off = off;
                uint32_t rowB = 0, rowS = 0;
                const uint32_t *srow = lib.class_bits + (size_t)seed_cls[q] * lib.wpad + lane;
#pragma unroll
                for (int j = 0; j < WPL; j++) {
                    const uint32_t wv = S.cls[q][j];
                    const uint32_t sv = __ldg(srow + j * 32);
                    uint32_t totB, totS;
                    const uint32_t exB = warp_excl_scan(__popc(wv), lane, totB);
                    const uint32_t exS = warp_excl_scan(__popc(sv), lane, totS);
                    uint32_t bits = wv;
                    while (bits) {
                        const int b = __ffs(bits) - 1;
                        bits &= bits - 1;
                        const uint32_t below = (1u << b) - 1;
                        const uint32_t rankB = rowB + exB + __popc(wv & below);
                        const uint32_t rankS = rowS + exS + __popc(sv & below);
                        const uint32_t r = (uint32_t)((j * 32 + lane) * 32 + b);
                        const uint32_t pos = __ldg(lib.positions + seed_off[q] + rankS);
                        SwItem it;
                        it.ro = (uint32_t)ro_idx; it.ref = r;
                        it.gwin = __ldg(lib.ref_gstart + r) + pos - (uint32_t)seed_i[q] - (uint32_t)kBand;
                        it.v = 0;
                        items[off + rankB] = it;
                    }
                    rowB += totB; rowS += totS;
                }
                if (lane == 0 && (cnt & 1)) {
                    SwItem it; it.ro = (uint32_t)ro_idx; it.ref = kInvalid; it.gwin = 0; it.v = 0;
                    items[off + cnt] = it;
                }
            }
            off += (cnt + 1) & ~1u;
        }
        if (lane == 0) ro[ro_idx] = rr;
    }
}

// ---------------------------------------------------------------------------------------------
// X3c: banded Smith-Waterman, linear gap, two candidates of the same oriented read per thread
// in the two s16 halves of DPX s16x2 instructions (VIADDMNMX / VIMNMX3).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t spread_bits32(uint32_t v) {   // bit b -> bit 2b
    uint64_t x = v;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
}

__device__ __forceinline__ void load_ref_window(const LibDev &lib, uint32_t g, uint64_t &bases, uint64_t &nspread) {
    // 32 bases starting at global coordinate g
    const uint32_t w = g >> 5, sh = (g & 31) * 2;
    const uint64_t a = __ldg(lib.ref2bit + w), b = __ldg(lib.ref2bit + w + 1);
    bases = sh ? ((a >> sh) | (b << (64 - sh))) : a;
    const uint64_t n01 = (uint64_t)__ldg(lib.refN + w) | ((uint64_t)__ldg(lib.refN + w + 1) << 32);
    nspread = spread_bits32((uint32_t)(n01 >> (g & 31)));
}

__global__ void __launch_bounds__(128)
sw_kernel(LibDev lib, ReadsDev r1, ReadsDev r2, uint64_t read0, int n_mates, const uint32_t *__restrict__ deferred,
          SwItem *__restrict__ items, uint32_t items_cap, Counters *__restrict__ ctr) {
    unsigned long long total = ctr->alloc & kItemMask;
    if (total > items_cap) total = 0;          // overflowed batch: nothing valid, the host retries
    const uint32_t n_pairs = (uint32_t)(total >> 1);
    const int n_ro = n_mates * 2;
    unsigned long long my_cells = 0, my_pairs = 0;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_pairs; t += gridDim.x * blockDim.x) {
        const SwItem ia = items[2 * t], ib = items[2 * t + 1];
        const bool hasB = ib.ref != kInvalid;
        const uint32_t sub = ia.ro % n_ro;
        const uint64_t read = read0 + deferred[ia.ro / n_ro];
        const int mate = sub >> 1, ori = sub & 1;
        const ReadsDev R = mate ? r2 : r1;
        const uint8_t *rec = R.packed + read * R.stride;
        const uint64_t *seq = reinterpret_cast<const uint64_t *>(rec);
        const uint32_t *nm = reinterpret_cast<const uint32_t *>(rec + (size_t)R.words * 8);
        const int L = R.len[read];
        const uint32_t gA = ia.gwin, gB = hasB ? ib.gwin : ia.gwin;
        uint32_t H[kNB];
#pragma unroll
        for (int b = 0; b < kNB; b++) H[b] = 0;
        uint32_t best = 0;
        uint64_t wa = 0, na = 0, wb = 0, nb = 0;
        for (int i = 0; i < L; i++) {
            if ((i & 15) == 0) {               // one 32-base window serves 16 rows of 17 cells
                load_ref_window(lib, gA + i, wa, na);
                load_ref_window(lib, gB + i, wb, nb);
            }
            const int idx = ori ? (L - 1 - i) : i;
            uint32_t q = (uint32_t)(seq[idx >> 5] >> (2 * (idx & 31))) & 3u;
            if (ori) q = 3u - q;
            const uint32_t qn = (nm[idx >> 5] >> (idx & 31)) & 1u;
            const uint64_t qrep = (0ull - (uint64_t)(q & 1)) & 0x5555555555555555ull |
                                  (0ull - (uint64_t)(q >> 1)) & 0xAAAAAAAAAAAAAAAAull;
            const int rsh = 2 * (i & 15);
            const uint64_t ya = (wa >> rsh) ^ qrep, yb = (wb >> rsh) ^ qrep;
            uint64_t za = ~(ya | (ya >> 1)) & ~(na >> rsh) & 0x5555555555555555ull;
            uint64_t zb = ~(yb | (yb >> 1)) & ~(nb >> rsh) & 0x5555555555555555ull;
            if (qn) { za = 0; zb = 0; }
            const uint64_t zab = za | (zb << 1);   // bit 2b: A matches at band cell b; bit 2b+1: B
            uint32_t left = 0;
#pragma unroll
            for (int b = 0; b < kNB; b++) {
                const uint32_t f2 = (uint32_t)(zab >> (2 * b)) & 3u;
                const uint32_t mf = (f2 * 0x8001u) & 0x00010001u;          // match flag per s16 half
                const uint32_t a = mf * kMatchDelta + H[b];                 // diag + (match - mismatch)
                const uint32_t up = (b + 1 < kNB) ? H[b + 1] : 0u;
                uint32_t h = __viaddmax_s16x2(up, kGP, 0u);                 // max(up + gap, 0)
                h = __viaddmax_s16x2(a, kXP, h);                            // max(diag + s, .)
                h = __viaddmax_s16x2(left, kGP, h);                         // max(left + gap, .)
                H[b] = h;
                left = h;
                best = __vimax3_s16x2(best, h, h);
            }
        }
        items[2 * t].v = best & 0xFFFFu;
        items[2 * t + 1].v = best >> 16;
        my_pairs += hasB ? 2 : 1;
        my_cells += (unsigned long long)(hasB ? 2 : 1) * (unsigned long long)L * kNB;
    }
    // one atomic per warp
    for (int o = 16; o; o >>= 1) {
        my_cells += __shfl_xor_sync(0xFFFFFFFFu, my_cells, o);
        my_pairs += __shfl_xor_sync(0xFFFFFFFFu, my_pairs, o);
    }
    if ((threadIdx.x & 31) == 0 && my_pairs) {
        atomicAdd(&ctr->sw_cells, my_cells);
        atomicAdd(&ctr->sw_pairs, my_pairs);
    }
}

// ---------------------------------------------------------------------------------------------
// Deferred X4: reads that went through Smith-Waterman.  One warp per deferred read.
// ---------------------------------------------------------------------------------------------
template <int WPL>
__global__ void __launch_bounds__(256)
call_deferred_kernel(LibDev lib, CallParams cp, int n_mates, const RoRec *__restrict__ ro,
                     const uint32_t *__restrict__ roB, const uint32_t *__restrict__ deferred,
                     const SwItem *__restrict__ items, uint32_t items_cap,
                     nb200_read_result *__restrict__ results, int32_t *__restrict__ feats,
                     uint16_t *__restrict__ row_nf, Counters *__restrict__ ctr) {
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t *sb = smem + (size_t)wib * lib.wpad;
    const unsigned long long alloc = ctr->alloc;
    if ((alloc & kItemMask) > items_cap) return;         // overflowed batch: the host retries
    const uint32_t n_def = (uint32_t)(alloc >> 40);
    const int n_ro = n_mates * 2;
    const bool paired = n_mates == 2;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t d = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); d < n_def; d += warps) {
        const uint32_t gw = deferred[d];
        ReadState<WPL> S;
        S.n_sw = 0;
#pragma unroll
        for (int o = 0; o < 4; o++) {
#pragma unroll
            for (int j = 0; j < WPL; j++) S.cls[o][j] = 0;
            S.nc[o] = 0; S.nh[o] = 0; S.vbest[o] = -1; S.len[o] = 0;
            if (o >= n_ro) continue;
            const RoRec rr = ro[(uint64_t)d * n_ro + o];
            S.nh[o] = rr.n_hits; S.len[o] = rr.len;
            if (rr.n_hits == 0 || rr.ncand == 0) continue;
            if (rr.full) {
                S.vbest[o] = (int)rr.len * kVW;
                const uint32_t *src = roB + ((uint64_t)d * n_ro + o) * lib.wpad + lane;
#pragma unroll
                for (int j = 0; j < WPL; j++) S.cls[o][j] = src[j * 32];
                S.nc[o] = rr.ncand;
            } else {
                S.n_sw++;
                const SwItem *seg = items + rr.item_off;
                uint32_t vb = 0;
                for (uint32_t t = lane; t < rr.ncand; t += 32) vb = max(vb, seg[t].v);
                const uint32_t vbest = warp_max(vb);
                const uint32_t slack = (uint32_t)cp.num_mismatches * kMatchDelta;
                const uint32_t vmin = vbest > slack ? vbest - slack : 0u;
                for (uint32_t j = lane; j < lib.wpad; j += 32) sb[j] = 0;
                __syncwarp();
                for (uint32_t t = lane; t < rr.ncand; t += 32) {
                    const SwItem it = seg[t];
                    if (it.v >= vmin) atomicOr(&sb[it.ref >> 5], 1u << (it.ref & 31));
                }
                __syncwarp();
                uint32_t cnt = 0;
#pragma unroll
                for (int j = 0; j < WPL; j++) { S.cls[o][j] = sb[j * 32 + lane]; cnt += __popc(S.cls[o][j]); }
                S.nc[o] = warp_sum(cnt);
                S.vbest[o] = (int)vbest;
                __syncwarp();
            }
        }
        call_read<WPL>(lib, cp, paired, S, sb, lane, results + gw, feats + (uint64_t)gw * cp.max_hits, row_nf + gw, ctr);
    }
}

}  // namespace nb200
