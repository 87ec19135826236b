// A6 on device: (cell, umi, feature list, score) rows -> per-cell UMI counts
// (reference nimble/__main__.py:234-293 + nimble/utils.py:119-224; pinned by oracle/ + tests/golden).
// Ordering trick: pandas sorts rows by the comma-joined feature STRING; lexicographic order of
// per-position token ranks (name+',' inner / name+NUL last) reproduces that byte order exactly, so
// an LSD radix sort over token-rank columns replaces every string comparison.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace nb200 {

constexpr int kMaxRounds = 64;   // bound on the reference's `while True` (nimble/utils.py:141)

__global__ void mark_rows_kernel(uint64_t n, const uint64_t *__restrict__ key, const uint16_t *__restrict__ nf,
                                 const double *__restrict__ score, uint8_t *__restrict__ flag) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool ok = nf[i] > 0;
    if (key && key[i] == NB200_NO_BARCODE) ok = false;
    if (score && score[i] != score[i]) ok = false;
    flag[i] = ok ? 1 : 0;
}

// token-rank columns p0 .. p0+cols-1 of every selected row packed into one radix key, the earlier
// position in the more significant bits (0 = list shorter than that position): one stable sort on
// the packed key orders by all of its columns at once
__global__ void gather_tok_kernel(uint32_t m, const uint32_t *__restrict__ perm, const int32_t *__restrict__ feats,
                                  uint32_t stride, const uint16_t *__restrict__ nf, uint32_t p0, uint32_t cols, uint32_t bits,
                                  const uint32_t *__restrict__ tok_end, const uint32_t *__restrict__ tok_comma,
                                  uint32_t *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t row = perm[i];
    const uint32_t n = nf[row];
    uint32_t key = 0;
    for (uint32_t j = 0; j < cols; j++) {
        const uint32_t p = p0 + j;
        uint32_t v = 0;
        if (p < n) {
            const uint32_t f = (uint32_t)feats[(uint64_t)row * stride + p];
            v = (p == n - 1 ? tok_end[f] : tok_comma[f]) + 1;
        }
        key = (key << bits) | v;
    }
    out[i] = key;
}

// CSR rows (off, ids) -> the padded [n, stride] layout the A6 kernels read; checks what the host loop
// used to check (ids in range, ascending inside a row): bit 0 / bit 1 of *err
__global__ void expand_rows_kernel(uint64_t n, const uint32_t *__restrict__ off, const uint32_t *__restrict__ ids, uint32_t stride,
                                   uint32_t n_features, int32_t *__restrict__ feats, uint16_t *__restrict__ nf,
                                   unsigned int *__restrict__ err) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t a = off[i], b = off[i + 1];
    nf[i] = (uint16_t)(b - a);
    int32_t *dst = feats + i * stride;
    uint32_t prev = 0, bad = 0;
    for (uint32_t j = 0; j < stride; j++) {
        int32_t v = -1;
        if (a + j < b) {
            const uint32_t f = ids[a + j];
            if (f >= n_features) bad |= 1u;
            if (j && f < prev) bad |= 2u;
            prev = f;
            v = (int32_t)f;
        }
        dst[j] = v;
    }
    if (bad) atomicOr(err, bad);
}

__global__ void gather_key64_kernel(uint32_t m, const uint32_t *__restrict__ perm, const uint64_t *__restrict__ key,
                                    uint64_t *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    out[i] = key ? key[perm[i]] : 0ull;
}

__global__ void gather_u32_kernel(uint32_t m, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ src,
                                  uint32_t *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    out[i] = src[perm[i]];
}

__global__ void key_heads_kernel(uint32_t m, const uint64_t *__restrict__ sorted_key, uint8_t *__restrict__ head) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    head[i] = (i == 0 || sorted_key[i] != sorted_key[i - 1]) ? 1 : 0;
}

__device__ __forceinline__ bool same_list(const int32_t *a, uint32_t na, const int32_t *b, uint32_t nb) {
    if (na != nb) return false;
    for (uint32_t j = 0; j < na; j++) if (a[j] != b[j]) return false;
    return true;
}

// order of two rows by the byte order of their comma-joined feature strings (same order the LSD token sort
// produces): token ranks position by position, name+NUL for the last name of a row, name+',' otherwise
__device__ __forceinline__ int compare_rows(const int32_t *la, uint32_t na, const int32_t *lb, uint32_t nb,
                                            const uint32_t *__restrict__ tok_end, const uint32_t *__restrict__ tok_comma) {
    const uint32_t mn = na < nb ? na : nb;
    for (uint32_t p = 0; p < mn; p++) {
        const uint32_t ta = (p == na - 1) ? tok_end[la[p]] : tok_comma[la[p]];
        const uint32_t tb = (p == nb - 1) ? tok_end[lb[p]] : tok_comma[lb[p]];
        if (ta != tb) return ta < tb ? -1 : 1;
    }
    return na < nb ? -1 : (na > nb ? 1 : 0);       // only reached for empty rows
}

// largest (cell, umi) group: decides between the per-group sort below and the global token sort
__global__ void max_group_kernel(uint32_t n_groups, const uint32_t *__restrict__ gstart, uint32_t m, unsigned int *__restrict__ out) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t sz = 0;
    if (g < n_groups) sz = ((g + 1 < n_groups) ? gstart[g + 1] : m) - gstart[g];
    sz = __reduce_max_sync(0xFFFFFFFFu, sz);
    if ((threadIdx.x & 31) == 0 && sz > *(volatile unsigned int *)out) atomicMax(out, sz);
}

// rows of one (cell, umi) group -> ascending feature-string order, stable (insertion sort on the permutation;
// groups are a handful of rows, the caller falls back to the global sort when one exceeds kLocalSortMax)
constexpr uint32_t kLocalSortMax = 48;
__global__ void __launch_bounds__(128)
group_sort_kernel(uint32_t n_groups, const uint32_t *__restrict__ gstart, uint32_t m, uint32_t *__restrict__ perm,
                  const int32_t *__restrict__ feats, uint32_t stride, const uint16_t *__restrict__ nf,
                  const uint32_t *__restrict__ tok_end, const uint32_t *__restrict__ tok_comma) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint32_t g0 = gstart[g], g1 = (g + 1 < n_groups) ? gstart[g + 1] : m;
    for (uint32_t i = g0 + 1; i < g1; i++) {
        const uint32_t r = perm[i];
        const int32_t *lr = feats + (uint64_t)r * stride;
        const uint32_t nr = nf[r];
        uint32_t j = i;
        while (j > g0) {
            const uint32_t q = perm[j - 1];
            if (compare_rows(feats + (uint64_t)q * stride, nf[q], lr, nr, tok_end, tok_comma) <= 0) break;   // stable
            perm[j] = q;
            j--;
        }
        perm[j] = r;
    }
}

__device__ inline void heap_sort_u32(uint32_t *a, int n) {
    for (int start = n / 2 - 1; start >= 0; start--) {
        int root = start;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= n) break;
            if (child + 1 < n && a[child] < a[child + 1]) child++;
            if (a[root] >= a[child]) break;
            uint32_t t = a[root]; a[root] = a[child]; a[child] = t;
            root = child;
        }
    }
    for (int end = n - 1; end > 0; end--) {
        uint32_t t = a[0]; a[0] = a[end]; a[end] = t;
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && a[child] < a[child + 1]) child++;
            if (a[root] >= a[child]) break;
            uint32_t u = a[root]; a[root] = a[child]; a[child] = u;
            root = child;
        }
    }
}

__device__ __forceinline__ int find_u32(const uint32_t *U, int nu, uint32_t v) {
    int lo = 0, hi = nu - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (U[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct UmiScratch {          // slices indexed by sorted-row position (x stride for per-feature arrays)
    uint32_t *rep;           // [m]     merged row -> representative sorted position
    double *S;               // [m]     merged score
    uint32_t *U;             // [m*stride] feature universe of the group
    double *fs, *fc;         // [m*stride] Kahan sum / compensation per feature
    uint8_t *flags;          // [m*stride] bit0 present, bit1 drop-now, bit2 keep
};

// nimble/utils.py:125-131 and :158-165 — feature scores from the ORIGINAL merged rows minus the
// names flagged drop-now.  Kahan per feature in order of appearance == pandas group_sum.
__device__ inline bool umi_scores(int m, const uint32_t *rep, const double *S, const uint32_t *perm,
                                  const int32_t *feats, uint32_t stride, const uint16_t *nf,
                                  const uint32_t *U, int nu, double *fs, double *fc, uint8_t *flags,
                                  double &total_out) {
    double total = 0.0;
    bool any = false;
    for (int j = 0; j < nu; j++) { fs[j] = 0.0; fc[j] = 0.0; flags[j] &= ~1; }
    for (int d = 0; d < m; d++) {
        const uint32_t row = perm[rep[d]];
        const int32_t *l = feats + (uint64_t)row * stride;
        const int n = nf[row];
        int len = 0;
        for (int j = 0; j < n; j++) if (!(flags[find_u32(U, nu, (uint32_t)l[j])] & 2)) len++;
        if (!len) continue;
        any = true;
        const double share = S[d] / (double)len;
        total += S[d];
        for (int j = 0; j < n; j++) {
            const int q = find_u32(U, nu, (uint32_t)l[j]);
            if (flags[q] & 2) continue;
            const double y = share - fc[q];
            const double t = fs[q] + y;
            double c = (t - fs[q]) - y;
            if (c != c) c = 0.0;
            fc[q] = c; fs[q] = t;
            flags[q] |= 1;
        }
    }
    total_out = total;
    return any;
}

// One thread per (cell, umi) group.  Rows arrive sorted by (key, feature-string order, input order).
__global__ void __launch_bounds__(128)
umi_kernel(uint32_t n_groups, const uint32_t *__restrict__ gstart, uint32_t m, const uint32_t *__restrict__ perm,
           const uint64_t *__restrict__ sorted_key, const int32_t *__restrict__ feats, uint32_t stride,
           const uint16_t *__restrict__ nf, const double *__restrict__ score, double threshold, int disable,
           UmiScratch sc, uint32_t *__restrict__ out_cell, uint16_t *__restrict__ out_n,
           int32_t *__restrict__ out_list, Counters *__restrict__ ctr) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint32_t g0 = gstart[g], g1 = (g + 1 < n_groups) ? gstart[g + 1] : m;
    uint32_t *rep = sc.rep + g0;
    double *S = sc.S + g0;
    uint32_t *U = sc.U + (uint64_t)g0 * stride;
    double *fs = sc.fs + (uint64_t)g0 * stride, *fc = sc.fc + (uint64_t)g0 * stride;
    uint8_t *flags = sc.flags + (uint64_t)g0 * stride;
    // merged rows: runs of identical lists; Kahan in input order (pandas groupby-sum, __main__.py:251)
    int md = 0;
    for (uint32_t t = g0; t < g1;) {
        const uint32_t ra = perm[t];
        double s = 0.0, c = 0.0;
        uint32_t u = t;
        while (u < g1) {
            const uint32_t rb = perm[u];
            if (!same_list(feats + (uint64_t)ra * stride, nf[ra], feats + (uint64_t)rb * stride, nf[rb])) break;
            const double v = score ? score[rb] : 1.0;
            const double y = v - c, tt = s + y;
            c = (tt - s) - y;
            if (c != c) c = 0.0;
            s = tt;
            u++;
        }
        rep[md] = t; S[md] = s; md++;
        t = u;
    }
    // Fast path (most UMIs): one merged row without duplicate names.  Every feature then has the same
    // ratio (S/len)/S, computed exactly as the general path would (Kahan of one term is the term).
    if (md == 1) {
        const uint32_t row = perm[rep[0]];
        const int n = nf[row];
        const int32_t *l = feats + (uint64_t)row * stride;
        bool dup = false;
        for (int j = 1; j < n; j++) dup |= (l[j] == l[j - 1]);
        if (!dup) {
            bool keep_all = true;
            if (!disable) {
                const double share = S[0] / (double)n;
                const double ratio = share / S[0];
                keep_all = !(ratio < threshold);
            }
            int32_t *dst = out_list + (uint64_t)g * stride;
            if (keep_all) for (int j = 0; j < n; j++) dst[j] = l[j];
            out_cell[g] = (uint32_t)(sorted_key[g0] >> 32);
            out_n[g] = (uint16_t)(keep_all ? n : 0);
            return;
        }
    }
    // feature universe
    int nu = 0;
    for (int d = 0; d < md; d++) {
        const uint32_t row = perm[rep[d]];
        for (int j = 0; j < nf[row]; j++) U[nu++] = (uint32_t)feats[(uint64_t)row * stride + j];
    }
    heap_sort_u32(U, nu);
    {
        int w = 0;
        for (int j = 0; j < nu; j++) if (j == 0 || U[j] != U[j - 1]) U[w++] = U[j];
        nu = w;
    }
    for (int j = 0; j < nu; j++) flags[j] = 0;
    if (disable) {
        for (int j = 0; j < nu; j++) flags[j] = 4;
    } else {
        double total = 0.0;
        bool any = umi_scores(md, rep, S, perm, feats, stride, nf, U, nu, fs, fc, flags, total);
        bool done = false;
        for (int round = 0; round < kMaxRounds && !done; round++) {
            if (!any) { done = true; break; }          // keep = {}
            int nd = 0;
            for (int j = 0; j < nu; j++) {
                flags[j] &= ~2;
                if (!(flags[j] & 1)) continue;
                const double ratio = fs[j] / total;       // IEEE: x/0 = inf, 0/0 = nan -> never < thr
                if (ratio < threshold) { flags[j] |= 2; nd++; }
            }
            if (nd == 0) {
                for (int j = 0; j < nu; j++) if (flags[j] & 1) flags[j] |= 4;
                done = true;
                break;
            }
            any = umi_scores(md, rep, S, perm, feats, stride, nf, U, nu, fs, fc, flags, total);
        }
        if (!done && any) for (int j = 0; j < nu; j++) if (flags[j] & 1) flags[j] |= 4;
    }
    // per merged row: filtered = sorted(set(row) & keep); empty rows vanish; intersect the rest
    int32_t *inter = out_list + (uint64_t)g * stride;
    int ni = -1;
    for (int d = 0; d < md; d++) {
        const uint32_t row = perm[rep[d]];
        const int32_t *l = feats + (uint64_t)row * stride;
        const int n = nf[row];
        if (ni < 0) {
            int w = 0;
            for (int j = 0; j < n; j++) {
                if (!(flags[find_u32(U, nu, (uint32_t)l[j])] & 4)) continue;
                if (w && inter[w - 1] == l[j]) continue;
                inter[w++] = l[j];
            }
            if (w) ni = w;
        } else {
            // does the row keep anything at all?  (rows left empty are dropped, utils.py:205)
            bool nonempty = false;
            for (int j = 0; j < n && !nonempty; j++) nonempty = (flags[find_u32(U, nu, (uint32_t)l[j])] & 4) != 0;
            if (!nonempty) continue;
            int w = 0;
            for (int x = 0; x < ni; x++) {
                bool in = false;
                for (int j = 0; j < n && !in; j++) in = (l[j] == inter[x]);   // kept by construction
                if (in) inter[w++] = inter[x];
            }
            ni = w;
        }
    }
    if (ni == 0) atomicAdd(&ctr->dropped_empty, 1ull);
    out_cell[g] = (uint32_t)(sorted_key[g0] >> 32);
    out_n[g] = (uint16_t)(ni > 0 ? ni : 0);
}

// run heads over UMI rows sorted by (cell, feature-string order)
__global__ void run_heads_kernel(uint32_t m, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ cell,
                                 const int32_t *__restrict__ list, uint32_t stride, const uint16_t *__restrict__ n,
                                 uint8_t *__restrict__ head) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    bool h = true;
    if (i) {
        const uint32_t a = perm[i], b = perm[i - 1];
        h = cell[a] != cell[b] || !same_list(list + (uint64_t)a * stride, n[a], list + (uint64_t)b * stride, n[b]);
    }
    head[i] = h ? 1 : 0;
}

// run i -> (cell, count, n_feat) ; n_feat feeds the exclusive scan that lays out the CSR ids
__global__ void emit_counts_kernel(uint32_t n_out, const uint32_t *__restrict__ starts, uint32_t m,
                                   const uint32_t *__restrict__ perm, const uint32_t *__restrict__ cell,
                                   const uint16_t *__restrict__ n, uint32_t *__restrict__ o_cell,
                                   uint32_t *__restrict__ o_count, uint32_t *__restrict__ o_n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_out) return;
    if (i == n_out) { o_n[i] = 0; return; }          // scan sentinel -> o_off[n_out] = total ids
    const uint32_t s = starts[i], e = (i + 1 < n_out) ? starts[i + 1] : m;
    const uint32_t row = perm[s];
    o_cell[i] = cell[row];
    o_count[i] = e - s;
    o_n[i] = n[row];
}

__global__ void emit_ids_kernel(uint32_t n_out, const uint32_t *__restrict__ starts, const uint32_t *__restrict__ perm,
                                const int32_t *__restrict__ list, uint32_t stride, const uint32_t *__restrict__ o_off,
                                uint32_t *__restrict__ o_ids) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const uint32_t row = perm[starts[i]];
    const uint32_t a = o_off[i], b = o_off[i + 1];
    for (uint32_t j = a; j < b; j++) o_ids[j] = (uint32_t)list[(uint64_t)row * stride + (j - a)];
}

// bulk data (no barcodes): every called read is its own "UMI" row with cell 0
__global__ void bulk_rows_kernel(uint32_t m, const uint32_t *__restrict__ perm, const int32_t *__restrict__ feats,
                                 uint32_t stride, const uint16_t *__restrict__ nf, uint32_t *__restrict__ out_cell,
                                 uint16_t *__restrict__ out_n, int32_t *__restrict__ out_list) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t row = perm[i];
    out_cell[i] = 0;
    out_n[i] = nf[row];
    for (uint32_t j = 0; j < stride; j++) out_list[(uint64_t)i * stride + j] = feats[(uint64_t)row * stride + j];
}

}  // namespace nb200
