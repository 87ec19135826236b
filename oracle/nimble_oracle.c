/* ORACLE — test infrastructure only.  Never linked into, imported by, or executed from the
 * product path (nimble_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file's shared object.
 *
 * What it is: a plain-C CPU restatement of nimble's read-assignment hot path
 * (SURVEY.md §8a rows X1, X3a, X3b, X3c, X4 and A6), following the frozen semantics in
 * DESIGN.md §2 ("SPEC").
 *
 * Parity status:
 *   - A6 (UMI merge / proportional threshold / intersection / count; reference
 *     nimble/__main__.py:234-293, nimble/utils.py:119-224) is PINNED: orc_a6 is checked against
 *     the 24 known-answer tests of /root/reference/test/test.py and against 260 fixtures
 *     produced by running the reference's own pandas report() (tests/golden/).
 *   - X1..X4 (index, k-mer lookup, equivalence-class intersection, scoring, feature calling):
 *     PARITY UNPINNED.  The reference arithmetic lives in the un-vendored third-party binary
 *     BimberLab/nimble-aligner (version unpinned: releases/latest, nimble/__main__.py:127-131),
 *     absent from /root/reference; the reference's tests hold no vector for it.  This file
 *     restates the published pseudoalignment algorithm (k-mer -> equivalence class,
 *     intersection over matched k-mers, score in bp) anchored on the reference's call sites:
 *     config keys nimble/types.py:12-25, argv nimble/__main__.py:177-192, consumed columns
 *     nimble/__main__.py:237-241.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- SPEC constants ------------------------------------------------------ */
#define SW_BAND 8          /* half-width w: band cells per row = 2w+1                       */
#define SW_W 64            /* V = SW_W*score - edits                                        */
#define V_MATCH 64         /* +1 * W                                                         */
#define V_MISMATCH (-129)  /* -(2*W + 1)                                                     */
#define V_GAP (-193)       /* -(3*W + 1), linear gap                                         */
#define MAX_READ_LEN 500
#define MAX_ROUNDS 64

enum { ST_NONE = 0, ST_PASS = 1, ST_NO_MATCH = 2, ST_EMPTY_CLASS = 3, ST_SCORE = 4, ST_PERCENT = 5,
       ST_MULTI = 6 };
enum { RS_CALLED = 0, RS_NO_PASS = 1, RS_NOT_VALID_PAIR = 2, RS_FORCE_INTERSECT = 3,
       RS_SCORE_FILTER = 4, RS_MULTI_HITS = 5, RS_MAX_HITS = 6 };
enum { SF_UNSTRANDED = 0, SF_FIVEPRIME = 1, SF_THREEPRIME = 2, SF_NONE = 3 };

typedef struct {
    int32_t k;
    int32_t score_threshold;
    int32_t score_filter;
    double score_percent;
    int32_t num_mismatches;
    int32_t discard_multiple_matches;
    int32_t intersect_level;
    int32_t discard_multi_hits;
    int32_t require_valid_pair;
    int32_t max_hits_to_report;
    int32_t strand_filter;
    int32_t pad_;
} orc_config;

typedef struct {
    uint16_t score[4];   /* r1 fwd, r1 rc, r2 fwd, r2 rc */
    uint16_t n_hits[4];
    uint16_t n_cand[4];  /* |B'| after SW refinement, saturated at 65535 */
    uint8_t edits[4];
    uint8_t status[4];
    uint8_t reason;
    uint8_t config;
    uint8_t n_feat;
    uint8_t n_sw;        /* orientations that went through Smith-Waterman */
    uint32_t pair_score;
} orc_read_result;

/* ---- index ------------------------------------------------------------------------------- */
typedef struct {
    int32_t k, n_refs, n_features;
    int64_t n_kmers, n_classes;
    /* open addressing: slot -> kmer index or -1 */
    uint64_t tmask;
    int32_t *slot;
    uint64_t *kmer;      /* distinct k-mers                                   */
    int32_t *kclass;     /* class id per distinct k-mer                       */
    int64_t *kpos_off;   /* offset into pos[] (one entry per class member)    */
    int32_t *pos;        /* first position of the k-mer in each member ref    */
    int64_t *class_off;  /* CSR: members of each class, ascending ref id      */
    int32_t *class_ref;
    int32_t *ref_len;
    uint8_t **ref_code;  /* per ref: 0..3, 4 = not ACGT                       */
    int32_t *ref_feature;
} orc_index;

static inline int base_code(char c) {
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 31; x *= 0x7fb5d329728ea185ULL;
    x ^= x >> 27; x *= 0x81dadef4bc2dd44dULL;
    x ^= x >> 33;
    return x;
}

typedef struct { uint64_t kmer; int32_t ref; int32_t pos; } occ_t;

static int occ_cmp(const void *a, const void *b) {
    const occ_t *x = (const occ_t *)a, *y = (const occ_t *)b;
    if (x->kmer != y->kmer) return x->kmer < y->kmer ? -1 : 1;
    if (x->ref != y->ref) return x->ref < y->ref ? -1 : 1;
    return (x->pos > y->pos) - (x->pos < y->pos);
}

/* first-order k-mer encoding: base i of the k-mer sits at bits [2(k-1-i), 2(k-i)) — i.e. the
 * usual big-endian 2-bit number.  Only equality matters to the oracle. */
orc_index *orc_index_build(int32_t n_refs, const char *const *seqs, const int32_t *ref_feature,
                           int32_t n_features, int32_t k) {
    if (k < 4 || k > 32) return NULL;
    orc_index *ix = (orc_index *)calloc(1, sizeof(orc_index));
    ix->k = k; ix->n_refs = n_refs; ix->n_features = n_features;
    ix->ref_len = (int32_t *)calloc(n_refs > 0 ? n_refs : 1, sizeof(int32_t));
    ix->ref_code = (uint8_t **)calloc(n_refs > 0 ? n_refs : 1, sizeof(uint8_t *));
    ix->ref_feature = (int32_t *)calloc(n_refs > 0 ? n_refs : 1, sizeof(int32_t));
    int64_t total = 0;
    for (int32_t r = 0; r < n_refs; r++) {
        int32_t len = (int32_t)strlen(seqs[r]);
        ix->ref_len[r] = len;
        ix->ref_feature[r] = ref_feature[r];
        ix->ref_code[r] = (uint8_t *)malloc(len > 0 ? len : 1);
        for (int32_t i = 0; i < len; i++) ix->ref_code[r][i] = (uint8_t)base_code(seqs[r][i]);
        if (len >= k) total += len - k + 1;
    }
    occ_t *occ = (occ_t *)malloc(sizeof(occ_t) * (size_t)(total > 0 ? total : 1));
    int64_t n_occ = 0;
    const uint64_t kmask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    for (int32_t r = 0; r < n_refs; r++) {
        uint64_t x = 0; int32_t valid = 0;
        for (int32_t i = 0; i < ix->ref_len[r]; i++) {
            int c = ix->ref_code[r][i];
            if (c > 3) { valid = 0; x = 0; continue; }
            x = ((x << 2) | (uint64_t)c) & kmask;
            if (++valid >= k) {
                occ[n_occ].kmer = x; occ[n_occ].ref = r; occ[n_occ].pos = i - k + 1; n_occ++;
            }
        }
    }
    qsort(occ, (size_t)n_occ, sizeof(occ_t), occ_cmp);
    /* distinct k-mers, their member lists (ascending ref) and first positions */
    int64_t n_kmers = 0, n_members = 0;
    for (int64_t i = 0; i < n_occ; i++) {
        if (i == 0 || occ[i].kmer != occ[i - 1].kmer) { n_kmers++; n_members++; }
        else if (occ[i].ref != occ[i - 1].ref) n_members++;
    }
    ix->n_kmers = n_kmers;
    ix->kmer = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n_kmers + 1));
    ix->kclass = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_kmers + 1));
    ix->kpos_off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_kmers + 1));
    ix->pos = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_members + 1));
    int32_t *mem_ref = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_members + 1));
    int64_t *mem_off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_kmers + 2));
    int64_t ki = -1, mi = 0;
    for (int64_t i = 0; i < n_occ; i++) {
        int newk = (i == 0 || occ[i].kmer != occ[i - 1].kmer);
        if (newk) { ki++; ix->kmer[ki] = occ[i].kmer; mem_off[ki] = mi; ix->kpos_off[ki] = mi; }
        if (newk || occ[i].ref != occ[i - 1].ref) { mem_ref[mi] = occ[i].ref; ix->pos[mi] = occ[i].pos; mi++; }
    }
    mem_off[n_kmers] = mi;
    free(occ);
    /* class dedup: hash member lists */
    uint64_t cmask = 16; while (cmask < (uint64_t)(2 * n_kmers + 2)) cmask <<= 1; cmask -= 1;
    int64_t *ctab = (int64_t *)malloc(sizeof(int64_t) * (size_t)(cmask + 1));   /* -> representative kmer */
    for (uint64_t i = 0; i <= cmask; i++) ctab[i] = -1;
    int64_t n_classes = 0;
    int64_t *class_rep = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_kmers + 1));
    for (int64_t q = 0; q < n_kmers; q++) {
        int64_t a = mem_off[q], b = mem_off[q + 1];
        uint64_t h = 0x9e3779b97f4a7c15ULL ^ (uint64_t)(b - a);
        for (int64_t j = a; j < b; j++) h = mix64(h ^ (uint64_t)(uint32_t)mem_ref[j]);
        uint64_t s = h & cmask;
        for (;;) {
            int64_t rep = ctab[s];
            if (rep < 0) { ctab[s] = q; ix->kclass[q] = (int32_t)n_classes; class_rep[n_classes++] = q; break; }
            int64_t ra = mem_off[rep], rb = mem_off[rep + 1];
            if (rb - ra == b - a && memcmp(mem_ref + ra, mem_ref + a, sizeof(int32_t) * (size_t)(b - a)) == 0) {
                ix->kclass[q] = ix->kclass[rep]; break;
            }
            s = (s + 1) & cmask;
        }
    }
    free(ctab);
    ix->n_classes = n_classes;
    ix->class_off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_classes + 1));
    int64_t tot = 0;
    for (int64_t c = 0; c < n_classes; c++) { ix->class_off[c] = tot; tot += mem_off[class_rep[c] + 1] - mem_off[class_rep[c]]; }
    ix->class_off[n_classes] = tot;
    ix->class_ref = (int32_t *)malloc(sizeof(int32_t) * (size_t)(tot + 1));
    for (int64_t c = 0; c < n_classes; c++) {
        int64_t a = mem_off[class_rep[c]], n = mem_off[class_rep[c] + 1] - a;
        memcpy(ix->class_ref + ix->class_off[c], mem_ref + a, sizeof(int32_t) * (size_t)n);
    }
    free(class_rep); free(mem_ref); free(mem_off);
    /* k-mer hash */
    uint64_t tm = 16; while (tm < (uint64_t)(2 * n_kmers + 2)) tm <<= 1;
    ix->tmask = tm - 1;
    ix->slot = (int32_t *)malloc(sizeof(int32_t) * (size_t)tm);
    for (uint64_t i = 0; i < tm; i++) ix->slot[i] = -1;
    for (int64_t q = 0; q < n_kmers; q++) {
        uint64_t s = mix64(ix->kmer[q]) & ix->tmask;
        while (ix->slot[s] >= 0) s = (s + 1) & ix->tmask;
        ix->slot[s] = (int32_t)q;
    }
    return ix;
}

void orc_index_free(orc_index *ix) {
    if (!ix) return;
    for (int32_t r = 0; r < ix->n_refs; r++) free(ix->ref_code[r]);
    free(ix->ref_code); free(ix->ref_len); free(ix->ref_feature); free(ix->slot); free(ix->kmer);
    free(ix->kclass); free(ix->kpos_off); free(ix->pos); free(ix->class_off); free(ix->class_ref);
    free(ix);
}

int64_t orc_index_n_kmers(const orc_index *ix) { return ix->n_kmers; }
int64_t orc_index_n_classes(const orc_index *ix) { return ix->n_classes; }
int64_t orc_index_class_members(const orc_index *ix) { return ix->class_off[ix->n_classes]; }

static inline int64_t lookup(const orc_index *ix, uint64_t x) {
    uint64_t s = mix64(x) & ix->tmask;
    for (;;) {
        int32_t q = ix->slot[s];
        if (q < 0) return -1;
        if (ix->kmer[q] == x) return q;
        s = (s + 1) & ix->tmask;
    }
}

/* ---- Smith-Waterman (SPEC §2.4) --------------------------------------------------------------
 * Banded local alignment of read q[0..L) against ref r; row i covers ref columns
 * j = start + i + b, b in [0, 2w].  Linear gap.  Cells outside the band or outside the
 * reference are unreachable (a base outside the reference never matches).  Returns max V. */
static int32_t sw_banded(const uint8_t *q, int32_t L, const uint8_t *ref, int32_t ref_len, int64_t start) {
    int32_t prev[2 * SW_BAND + 3], cur[2 * SW_BAND + 3];
    const int nb = 2 * SW_BAND + 1;
    for (int b = 0; b < nb + 2; b++) prev[b] = 0;
    int32_t best = 0;
    for (int32_t i = 0; i < L; i++) {
        cur[0] = 0;                      /* cur[b+1] holds band cell b; cur[0] = left of cell 0 */
        for (int b = 0; b < nb; b++) {
            int64_t j = start + i + b;
            int match = 0;
            if (j >= 0 && j < ref_len && q[i] < 4 && ref[j] == q[i]) match = 1;
            int32_t diag = prev[b + 1];              /* (i-1, j-1) is band cell b of row i-1 */
            int32_t up = (b + 1 < nb) ? prev[b + 2] : 0; /* (i-1, j) is band cell b+1          */
            int32_t left = cur[b];
            int32_t h = diag + (match ? V_MATCH : V_MISMATCH);
            if (up + V_GAP > h) h = up + V_GAP;
            if (left + V_GAP > h) h = left + V_GAP;
            if (h < 0) h = 0;
            cur[b + 1] = h;
            if (h > best) best = h;
        }
        memcpy(prev, cur, sizeof(cur));
    }
    return best;
}

/* ---- one mate in one orientation (SPEC §2.2-2.5) ---------------------------------------------- */
typedef struct {
    int status; int score; int edits; int n_hits; int sw;
    int32_t *cls; int32_t n_cls;          /* B' ascending ref ids */
} ori_t;

static void align_orientation(const orc_index *ix, const orc_config *cfg, const uint8_t *q, int32_t L,
                              ori_t *o, int32_t *scratchA, int32_t *scratchB, int32_t *vbuf) {
    const int k = ix->k;
    o->status = ST_NO_MATCH; o->score = 0; o->edits = 0; o->n_hits = 0; o->n_cls = 0; o->sw = 0;
    if (L < k) return;
    const uint64_t kmask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    uint64_t x = 0; int valid = 0;
    int n_hits = 0; int64_t seed_q = -1; int32_t seed_i = -1;
    int32_t last_class = -1;
    int32_t *B = scratchA; int32_t nB = -1;   /* -1 = universe */
    for (int32_t i = 0; i < L; i++) {
        int c = q[i];
        if (c > 3) { valid = 0; x = 0; continue; }
        x = ((x << 2) | (uint64_t)c) & kmask;
        if (++valid < k) continue;
        int64_t e = lookup(ix, x);
        if (e < 0) continue;
        n_hits++;
        if (seed_q < 0) { seed_q = e; seed_i = i - k + 1; }
        int32_t cl = ix->kclass[e];
        if (cl == last_class) continue;
        last_class = cl;
        const int32_t *m = ix->class_ref + ix->class_off[cl];
        int32_t nm = (int32_t)(ix->class_off[cl + 1] - ix->class_off[cl]);
        if (nB < 0) { memcpy(B, m, sizeof(int32_t) * (size_t)nm); nB = nm; }
        else {
            int32_t a = 0, b = 0, n = 0;
            while (a < nB && b < nm) {
                if (B[a] < m[b]) a++; else if (B[a] > m[b]) b++; else { B[n++] = B[a]; a++; b++; }
            }
            nB = n;
        }
    }
    o->n_hits = n_hits;
    if (n_hits == 0) return;
    if (nB <= 0) { o->status = ST_EMPTY_CLASS; return; }
    int32_t vbest = 0;
    if (n_hits == L - k + 1) {
        /* every k-mer of the read is in the index: exact containment is assumed, no SW */
        vbest = L * SW_W;
        for (int32_t t = 0; t < nB; t++) vbuf[t] = vbest;
    } else {
        o->sw = 1;
        /* seed = first hit; its per-member first positions give each candidate's diagonal */
        int32_t scl = ix->kclass[seed_q];
        const int32_t *sm = ix->class_ref + ix->class_off[scl];
        int32_t snm = (int32_t)(ix->class_off[scl + 1] - ix->class_off[scl]);
        const int32_t *spos = ix->pos + ix->kpos_off[seed_q];
        int32_t a = 0;
        for (int32_t t = 0; t < nB; t++) {
            while (a < snm && sm[a] < B[t]) a++;
            /* B is a subset of the seed's class by construction */
            int32_t r = B[t];
            int64_t start = (int64_t)spos[a] - seed_i - SW_BAND;
            vbuf[t] = sw_banded(q, L, ix->ref_code[r], ix->ref_len[r], start);
            if (vbuf[t] > vbest) vbest = vbuf[t];
        }
    }
    int32_t slack = cfg->num_mismatches * (V_MATCH - V_MISMATCH);   /* one more mismatched base costs 64 + 129 */
    int32_t n = 0;
    for (int32_t t = 0; t < nB; t++) if (vbuf[t] >= vbest - slack) scratchB[n++] = B[t];
    o->cls = scratchB; o->n_cls = n;
    o->score = (vbest + SW_W - 1) / SW_W;
    o->edits = o->score * SW_W - vbest;
    if (o->score < cfg->score_threshold) { o->status = ST_SCORE; return; }
    if ((double)o->score / (double)L < cfg->score_percent) { o->status = ST_PERCENT; return; }
    if (cfg->discard_multiple_matches && n > 1) { o->status = ST_MULTI; return; }
    o->status = ST_PASS;
}

static void encode_read(const char *s, int32_t L, uint8_t *fwd, uint8_t *rc) {
    for (int32_t i = 0; i < L; i++) fwd[i] = (uint8_t)base_code(s[i]);
    for (int32_t i = 0; i < L; i++) { uint8_t c = fwd[L - 1 - i]; rc[i] = c > 3 ? 4 : (uint8_t)(3 - c); }
}

static int32_t set_union(const int32_t *a, int32_t na, const int32_t *b, int32_t nb, int32_t *out) {
    int32_t i = 0, j = 0, n = 0;
    while (i < na || j < nb) {
        if (j >= nb || (i < na && a[i] < b[j])) out[n++] = a[i++];
        else if (i >= na || b[j] < a[i]) out[n++] = b[j++];
        else { out[n++] = a[i]; i++; j++; }
    }
    return n;
}
static int32_t set_inter(const int32_t *a, int32_t na, const int32_t *b, int32_t nb, int32_t *out) {
    int32_t i = 0, j = 0, n = 0;
    while (i < na && j < nb) {
        if (a[i] < b[j]) i++; else if (a[i] > b[j]) j++; else { out[n++] = a[i]; i++; j++; }
    }
    return n;
}

static int int_cmp(const void *a, const void *b) { int32_t x = *(const int32_t *)a, y = *(const int32_t *)b; return (x > y) - (x < y); }

/* SPEC §2.6: strand configurations, mate combination, config choice, feature calling */
static void call_read(const orc_index *ix, const orc_config *cfg, ori_t o[4], int paired,
                      int32_t *tmp, int32_t *best_cls, orc_read_result *res, int32_t *feats) {
    static const int cfg_tab[4][2] = { {0, 3}, {1, 2}, {0, 2}, {1, 3} };  /* F, R, FF, RR : (r1 ori idx, r2 ori idx) */
    int order[4], n_cfg = 0;
    switch (cfg->strand_filter) {
    case SF_FIVEPRIME: order[n_cfg++] = 0; break;
    case SF_THREEPRIME: order[n_cfg++] = 1; break;
    case SF_NONE: order[n_cfg++] = 0; order[n_cfg++] = 1; if (paired) { order[n_cfg++] = 2; order[n_cfg++] = 3; } break;
    default: order[n_cfg++] = 0; order[n_cfg++] = 1; break;
    }
    int chosen = -1; uint32_t chosen_score = 0; int32_t n_best = 0; int chosen_maxmate = 0;
    int first_fail = RS_NO_PASS;
    for (int ci = 0; ci < n_cfg; ci++) {
        int c = order[ci];
        ori_t *a = &o[cfg_tab[c][0]];
        ori_t *b = paired ? &o[cfg_tab[c][1]] : NULL;
        int pa = a->status == ST_PASS, pb = b && b->status == ST_PASS;
        int fail = -1; int32_t n = 0; uint32_t sc = 0; int maxmate = 0;
        if (!paired) {
            if (!pa) fail = RS_NO_PASS;
            else { memcpy(tmp, a->cls, sizeof(int32_t) * (size_t)a->n_cls); n = a->n_cls; sc = (uint32_t)a->score; maxmate = a->score; }
        } else if (cfg->require_valid_pair && !(pa && pb)) fail = RS_NOT_VALID_PAIR;
        else if (!pa && !pb) fail = RS_NO_PASS;
        else if (pa && pb) {
            sc = (uint32_t)(a->score + b->score); maxmate = a->score > b->score ? a->score : b->score;
            if (cfg->intersect_level == 0) n = set_union(a->cls, a->n_cls, b->cls, b->n_cls, tmp);
            else {
                n = set_inter(a->cls, a->n_cls, b->cls, b->n_cls, tmp);
                if (n == 0) {
                    if (cfg->intersect_level >= 2) fail = RS_FORCE_INTERSECT;
                    else {
                        ori_t *w = (b->score > a->score) ? b : a;   /* tie -> mate 1 */
                        memcpy(tmp, w->cls, sizeof(int32_t) * (size_t)w->n_cls); n = w->n_cls;
                    }
                }
            }
        } else {
            if (cfg->intersect_level >= 2) fail = RS_FORCE_INTERSECT;
            else {
                ori_t *w = pa ? a : b;
                memcpy(tmp, w->cls, sizeof(int32_t) * (size_t)w->n_cls); n = w->n_cls; sc = (uint32_t)w->score; maxmate = w->score;
            }
        }
        if (fail >= 0) { if (ci == 0) first_fail = fail; continue; }
        if (chosen < 0 || sc > chosen_score) {
            chosen = c; chosen_score = sc; n_best = n; chosen_maxmate = maxmate;
            memcpy(best_cls, tmp, sizeof(int32_t) * (size_t)n);
        }
    }
    res->n_feat = 0; res->pair_score = 0; res->config = 255;
    if (chosen < 0) { res->reason = (uint8_t)first_fail; return; }
    res->config = (uint8_t)chosen; res->pair_score = chosen_score;
    if (chosen_maxmate < cfg->score_filter) { res->reason = RS_SCORE_FILTER; return; }
    /* refs -> features: dedup + ascending feature id (ids are name-rank ordered by the caller) */
    for (int32_t t = 0; t < n_best; t++) tmp[t] = ix->ref_feature[best_cls[t]];
    qsort(tmp, (size_t)n_best, sizeof(int32_t), int_cmp);
    int32_t nf = 0;
    for (int32_t t = 0; t < n_best; t++) if (t == 0 || tmp[t] != tmp[t - 1]) tmp[nf++] = tmp[t];
    if (cfg->discard_multi_hits > 0 && nf > cfg->discard_multi_hits) { res->reason = RS_MULTI_HITS; return; }
    if (nf > cfg->max_hits_to_report) { res->reason = RS_MAX_HITS; return; }
    res->reason = RS_CALLED; res->n_feat = (uint8_t)nf;
    for (int32_t t = 0; t < nf; t++) feats[t] = tmp[t];
}

/* r1/r2: concatenated ASCII bases; off[n+1].  r2 may be NULL (single-end).  feats: n * max_hits_to_report.
 * returns 0, or -1 on a read longer than MAX_READ_LEN / bad config. */
int32_t orc_align(const orc_index *ix, const orc_config *cfg, int64_t n_reads,
                  const char *r1, const int64_t *r1_off, const char *r2, const int64_t *r2_off,
                  int32_t n_threads, orc_read_result *out, int32_t *feats) {
    if (cfg->max_hits_to_report < 1 || cfg->max_hits_to_report > 255) return -1;
    for (int64_t i = 0; i < n_reads; i++) {
        if (r1_off[i + 1] - r1_off[i] > MAX_READ_LEN) return -1;
        if (r2 && r2_off[i + 1] - r2_off[i] > MAX_READ_LEN) return -1;
    }
    const int mh = cfg->max_hits_to_report;
    const int nr = ix->n_refs > 0 ? ix->n_refs : 1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        int32_t *sa = (int32_t *)malloc(sizeof(int32_t) * (size_t)nr * 11);
        int32_t *cls[4] = { sa + nr, sa + 2 * nr, sa + 3 * nr, sa + 4 * nr };
        int32_t *vbuf = sa + 5 * nr, *tmp = sa + 6 * nr, *best = sa + 8 * nr, *work = sa + 9 * nr;
        (void)work;
        uint8_t fwd[MAX_READ_LEN + 1], rc[MAX_READ_LEN + 1];
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < n_reads; i++) {
            ori_t o[4]; memset(o, 0, sizeof(o));
            orc_read_result *res = &out[i]; memset(res, 0, sizeof(*res));
            int32_t L1 = (int32_t)(r1_off[i + 1] - r1_off[i]);
            encode_read(r1 + r1_off[i], L1, fwd, rc);
            align_orientation(ix, cfg, fwd, L1, &o[0], sa, cls[0], vbuf);
            align_orientation(ix, cfg, rc, L1, &o[1], sa, cls[1], vbuf);
            int paired = r2 != NULL;
            if (paired) {
                int32_t L2 = (int32_t)(r2_off[i + 1] - r2_off[i]);
                encode_read(r2 + r2_off[i], L2, fwd, rc);
                align_orientation(ix, cfg, fwd, L2, &o[2], sa, cls[2], vbuf);
                align_orientation(ix, cfg, rc, L2, &o[3], sa, cls[3], vbuf);
            }
            for (int t = 0; t < 4; t++) {
                res->score[t] = (uint16_t)o[t].score; res->edits[t] = (uint8_t)o[t].edits;
                res->status[t] = (uint8_t)o[t].status; res->n_hits[t] = (uint16_t)o[t].n_hits;
                res->n_cand[t] = (uint16_t)(o[t].n_cls > 65535 ? 65535 : o[t].n_cls);
                res->n_sw += (uint8_t)o[t].sw;
            }
            for (int t = 0; t < mh; t++) feats[i * mh + t] = -1;
            call_read(ix, cfg, o, paired, tmp, best, res, feats + i * mh);
        }
        free(sa);
    }
    return 0;
}

/* ---- A6 on ids (pinned; mirrors oracle/a6_py.py) -------------------------------------------
 * Rows: key = (cell << 32 | umi); feature list = ids ascending in NAME order (duplicates
 * allowed); score (NULL => 1.0).  tok_end[id] / tok_comma[id] = rank of "name\0" / "name,"
 * among all 2F such tokens in byte order: lexicographic order of token sequences equals byte
 * order of the comma-joined strings the reference's pandas groupby sorts by
 * (nimble/__main__.py:248-251, 289). */
typedef struct {
    const uint64_t *key; const int32_t *off; const uint32_t *ids; const uint32_t *tok_end, *tok_comma;
} a6_ctx;
static a6_ctx g_a6;   /* qsort has no context argument; orc_a6 is not re-entrant */
#pragma omp threadprivate(g_a6)

static int fset_cmp_tok(const uint32_t *a, int32_t na, const uint32_t *b, int32_t nb,
                        const uint32_t *te, const uint32_t *tc) {
    int32_t n = na < nb ? na : nb;
    for (int32_t i = 0; i < n; i++) {
        uint32_t ta = (i == na - 1) ? te[a[i]] : tc[a[i]];
        uint32_t tb = (i == nb - 1) ? te[b[i]] : tc[b[i]];
        if (ta != tb) return ta < tb ? -1 : 1;
    }
    return (na > nb) - (na < nb);
}

static int a6_row_cmp(const void *pa, const void *pb) {
    int64_t a = *(const int64_t *)pa, b = *(const int64_t *)pb;
    if (g_a6.key[a] != g_a6.key[b]) return g_a6.key[a] < g_a6.key[b] ? -1 : 1;
    int c = fset_cmp_tok(g_a6.ids + g_a6.off[a], g_a6.off[a + 1] - g_a6.off[a],
                         g_a6.ids + g_a6.off[b], g_a6.off[b + 1] - g_a6.off[b], g_a6.tok_end, g_a6.tok_comma);
    if (c) return c;
    return (a > b) - (a < b);   /* stable: original order inside a merged row */
}

typedef struct { uint32_t cell; int32_t n; uint32_t *ids; } umi_out_t;
static const uint32_t *g_te, *g_tc;
#pragma omp threadprivate(g_te, g_tc)
static int umi_out_cmp(const void *pa, const void *pb) {
    const umi_out_t *a = (const umi_out_t *)pa, *b = (const umi_out_t *)pb;
    if (a->cell != b->cell) return a->cell < b->cell ? -1 : 1;
    return fset_cmp_tok(a->ids, a->n, b->ids, b->n, g_te, g_tc);
}

static int u32_cmp(const void *a, const void *b) { uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b; return (x > y) - (x < y); }

/* feature scores from the ORIGINAL merged rows minus the names flagged in drop[]
 * (nimble/utils.py:125-131 first pass, :158-165 re-passes).  Kahan per feature = pandas group_sum. */
static int a6_scores(int64_t m, const int64_t *rep, const double *S, const int32_t *off, const uint32_t *ids,
                     const uint32_t *U, int64_t nu, const uint8_t *drop, double *fs, double *fc,
                     uint8_t *present, double *total_out) {
    double total = 0.0; int any = 0;
    for (int64_t j = 0; j < nu; j++) { fs[j] = 0.0; fc[j] = 0.0; present[j] = 0; }
    for (int64_t d = 0; d < m; d++) {
        int32_t a = off[rep[d]], b = off[rep[d] + 1], len = 0;
        for (int32_t j = a; j < b; j++) {
            const uint32_t *p = (const uint32_t *)bsearch(&ids[j], U, (size_t)nu, sizeof(uint32_t), u32_cmp);
            if (!drop[p - U]) len++;
        }
        if (len == 0) continue;
        any = 1;
        double share = S[d] / (double)len;
        total += S[d];
        for (int32_t j = a; j < b; j++) {
            const uint32_t *p = (const uint32_t *)bsearch(&ids[j], U, (size_t)nu, sizeof(uint32_t), u32_cmp);
            int64_t q = p - U;
            if (drop[q]) continue;
            double y = share - fc[q], tt = fs[q] + y; fc[q] = tt - fs[q] - y; if (fc[q] != fc[q]) fc[q] = 0.0; fs[q] = tt;
            present[q] = 1;
        }
    }
    *total_out = total;
    return any;
}

void orc_free(void *p) { free(p); }

int32_t orc_max_threads(void);
/* the whole A6 stage over the usable rows listed in idx[0..n) (consumed) */
static int64_t a6_run(int64_t *idx, int64_t n, const uint64_t *key, const int32_t *off, const uint32_t *ids,
                      const double *score, const uint32_t *tok_end, const uint32_t *tok_comma,
                      double threshold, int32_t disable_thresholding,
                      uint32_t **out_cell, uint32_t **out_count, int32_t **out_off, uint32_t **out_ids,
                      int64_t *dropped_empty) {
    g_a6.key = key; g_a6.off = off; g_a6.ids = ids; g_a6.tok_end = tok_end; g_a6.tok_comma = tok_comma;
    qsort(idx, (size_t)n, sizeof(int64_t), a6_row_cmp);
    umi_out_t *umis = (umi_out_t *)malloc(sizeof(umi_out_t) * (size_t)(n + 1));
    int64_t n_umi = 0, dropped = 0;
    int64_t g0 = 0;
    while (g0 < n) {
        int64_t g1 = g0;
        while (g1 < n && key[idx[g1]] == key[idx[g0]]) g1++;
        /* merged rows: runs of identical lists */
        int64_t m = 0;
        int64_t *rep = (int64_t *)malloc(sizeof(int64_t) * (size_t)(g1 - g0));
        double *S = (double *)malloc(sizeof(double) * (size_t)(g1 - g0));
        for (int64_t t = g0; t < g1;) {
            int64_t u = t; double s = 0.0, c = 0.0;
            while (u < g1) {
                int64_t a = idx[t], b = idx[u];
                int32_t na = off[a + 1] - off[a], nb = off[b + 1] - off[b];
                if (na != nb || memcmp(ids + off[a], ids + off[b], sizeof(uint32_t) * (size_t)na) != 0) break;
                double v = score ? score[b] : 1.0;
                double y = v - c, tt = s + y; c = tt - s - y; if (c != c) c = 0.0; s = tt;
                u++;
            }
            rep[m] = idx[t]; S[m] = s; m++; t = u;
        }
        /* universe of feature ids in the group */
        int64_t tot = 0;
        for (int64_t d = 0; d < m; d++) tot += off[rep[d] + 1] - off[rep[d]];
        uint32_t *U = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(tot + 1));
        int64_t nu = 0;
        for (int64_t d = 0; d < m; d++) for (int32_t j = off[rep[d]]; j < off[rep[d] + 1]; j++) U[nu++] = ids[j];
        qsort(U, (size_t)nu, sizeof(uint32_t), u32_cmp);
        int64_t w = 0; for (int64_t j = 0; j < nu; j++) if (j == 0 || U[j] != U[j - 1]) U[w++] = U[j];
        nu = w;
        uint8_t *keep = (uint8_t *)malloc((size_t)nu + 1);      /* final kept set            */
        uint8_t *dropnow = (uint8_t *)calloc((size_t)nu + 1, 1); /* this round's to_drop     */
        uint8_t *present = (uint8_t *)malloc((size_t)nu + 1);   /* feature_scores.index      */
        double *fs = (double *)malloc(sizeof(double) * (size_t)(nu + 1));
        double *fc = (double *)malloc(sizeof(double) * (size_t)(nu + 1));
        double total_dummy = 0.0;
        if (disable_thresholding) { for (int64_t j = 0; j < nu; j++) keep[j] = 1; }
        else {
            /* mirrors oracle/a6_py.py: threshold_group (nimble/utils.py:120-171) */
            int any = a6_scores(m, rep, S, off, ids, U, nu, dropnow, fs, fc, present, &total_dummy);
            double total = total_dummy;
            int done = 0;
            for (int round = 0; round < MAX_ROUNDS && !done; round++) {
                if (!any) { for (int64_t j = 0; j < nu; j++) keep[j] = 0; done = 1; break; }
                int nd = 0;
                for (int64_t j = 0; j < nu; j++) {
                    dropnow[j] = 0;
                    if (!present[j]) continue;
                    double ratio = fs[j] / total;     /* IEEE: x/0 = inf, 0/0 = nan -> never < thr */
                    if (ratio < threshold) { dropnow[j] = 1; nd++; }
                }
                if (nd == 0) { for (int64_t j = 0; j < nu; j++) keep[j] = present[j]; done = 1; break; }
                any = a6_scores(m, rep, S, off, ids, U, nu, dropnow, fs, fc, present, &total_dummy);
                total = total_dummy;
            }
            if (!done) { for (int64_t j = 0; j < nu; j++) keep[j] = any ? present[j] : 0; }
        }
        /* per merged row: filtered = sorted(set(row) & keep); rows left empty vanish; intersect */
        uint32_t *inter = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(nu + 1));
        int64_t ni = -1;
        for (int64_t d = 0; d < m; d++) {
            int32_t a = off[rep[d]], b = off[rep[d] + 1];
            uint32_t *f = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(b - a + 1)); int32_t nf = 0;
            for (int32_t j = a; j < b; j++) {
                uint32_t *p = (uint32_t *)bsearch(&ids[j], U, (size_t)nu, sizeof(uint32_t), u32_cmp);
                if (!keep[p - U]) continue;
                if (nf && f[nf - 1] == ids[j]) continue;
                f[nf++] = ids[j];
            }
            if (nf == 0) { free(f); continue; }
            if (ni < 0) { memcpy(inter, f, sizeof(uint32_t) * (size_t)nf); ni = nf; }
            else {
                int64_t x = 0, y = 0, z = 0;
                while (x < ni && y < nf) { if (inter[x] < f[y]) x++; else if (inter[x] > f[y]) y++; else { inter[z++] = inter[x]; x++; y++; } }
                ni = z;
            }
            free(f);
        }
        if (ni == 0) dropped++;
        if (ni > 0) {
            umis[n_umi].cell = (uint32_t)(key[idx[g0]] >> 32); umis[n_umi].n = (int32_t)ni;
            umis[n_umi].ids = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)ni);
            memcpy(umis[n_umi].ids, inter, sizeof(uint32_t) * (size_t)ni); n_umi++;
        }
        free(inter); free(keep); free(dropnow); free(present); free(fs); free(fc); free(U); free(rep); free(S);
        g0 = g1;
    }
    free(idx);
    g_te = tok_end; g_tc = tok_comma;
    qsort(umis, (size_t)n_umi, sizeof(umi_out_t), umi_out_cmp);
    int64_t n_out = 0, n_ids = 0;
    for (int64_t i = 0; i < n_umi; i++) if (i == 0 || umi_out_cmp(&umis[i], &umis[i - 1]) != 0) { n_out++; n_ids += umis[i].n; }
    *out_cell = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n_out + 1));
    *out_count = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n_out + 1));
    *out_off = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_out + 2));
    *out_ids = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n_ids + 1));
    int64_t o = -1, w = 0;
    for (int64_t i = 0; i < n_umi; i++) {
        if (i == 0 || umi_out_cmp(&umis[i], &umis[i - 1]) != 0) {
            o++; (*out_cell)[o] = umis[i].cell; (*out_count)[o] = 0; (*out_off)[o] = (int32_t)w;
            memcpy(*out_ids + w, umis[i].ids, sizeof(uint32_t) * (size_t)umis[i].n); w += umis[i].n;
        }
        (*out_count)[o]++;
    }
    (*out_off)[n_out] = (int32_t)w;
    for (int64_t i = 0; i < n_umi; i++) free(umis[i].ids);
    free(umis);
    if (dropped_empty) *dropped_empty = dropped;
    return n_out;
}

static int a6_usable(int64_t i, const uint64_t *key, const int32_t *off, const double *score) {
    if (key[i] == ~0ULL || off[i + 1] == off[i]) return 0;
    if (score && score[i] != score[i]) return 0;
    return 1;
}

int64_t orc_a6(int64_t n_rows, const uint64_t *key, const int32_t *off, const uint32_t *ids,
               const double *score, const uint32_t *tok_end, const uint32_t *tok_comma,
               double threshold, int32_t disable_thresholding,
               uint32_t **out_cell, uint32_t **out_count, int32_t **out_off, uint32_t **out_ids,
               int64_t *dropped_empty) {
    int64_t *idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_rows + 1));
    int64_t n = 0;
    for (int64_t i = 0; i < n_rows; i++) if (a6_usable(i, key, off, score)) idx[n++] = i;
    return a6_run(idx, n, key, off, ids, score, tok_end, tok_comma, threshold, disable_thresholding,
                  out_cell, out_count, out_off, out_ids, dropped_empty);
}

/* Same result on n_threads host threads (benchmark CPU arm).  Cells are independent in A6 (every group key
 * and every output row carries its cell), so rows are dealt to shards by a hash of the cell, each shard runs
 * the serial stage, and the shard tables - disjoint in cells, each already in output order - are merged by
 * cell. */
typedef struct { uint32_t cell; int32_t shard; int64_t row; } a6_merge_t;
static int a6_merge_cmp(const void *pa, const void *pb) {
    const a6_merge_t *a = (const a6_merge_t *)pa, *b = (const a6_merge_t *)pb;
    if (a->cell != b->cell) return a->cell < b->cell ? -1 : 1;
    return (a->row > b->row) - (a->row < b->row);
}
int64_t orc_a6_mt(int32_t n_threads, int64_t n_rows, const uint64_t *key, const int32_t *off, const uint32_t *ids,
                  const double *score, const uint32_t *tok_end, const uint32_t *tok_comma,
                  double threshold, int32_t disable_thresholding,
                  uint32_t **out_cell, uint32_t **out_count, int32_t **out_off, uint32_t **out_ids,
                  int64_t *dropped_empty) {
    int T = n_threads > 0 ? n_threads : orc_max_threads();
    if (T > 256) T = 256;
    if (T <= 1 || n_rows < 4096)
        return orc_a6(n_rows, key, off, ids, score, tok_end, tok_comma, threshold, disable_thresholding,
                      out_cell, out_count, out_off, out_ids, dropped_empty);
    int64_t *cnt = (int64_t *)calloc((size_t)T + 1, sizeof(int64_t));
    uint8_t *sh = (uint8_t *)malloc((size_t)n_rows + 1);
    for (int64_t i = 0; i < n_rows; i++) {
        if (!a6_usable(i, key, off, score)) { sh[i] = 255; continue; }
        uint32_t h = (uint32_t)(key[i] >> 32) * 0x9E3779B1u;
        sh[i] = (uint8_t)(((uint64_t)h * (uint64_t)T) >> 32);
        cnt[sh[i]]++;
    }
    int64_t **lists = (int64_t **)malloc(sizeof(int64_t *) * (size_t)T);
    int64_t *fill = (int64_t *)calloc((size_t)T, sizeof(int64_t));
    for (int t = 0; t < T; t++) lists[t] = (int64_t *)malloc(sizeof(int64_t) * (size_t)(cnt[t] + 1));
    for (int64_t i = 0; i < n_rows; i++) if (sh[i] != 255) lists[sh[i]][fill[sh[i]]++] = i;
    free(sh);
    uint32_t **pc = (uint32_t **)calloc((size_t)T, sizeof(void *)), **pn = (uint32_t **)calloc((size_t)T, sizeof(void *));
    uint32_t **pi = (uint32_t **)calloc((size_t)T, sizeof(void *));
    int32_t **po = (int32_t **)calloc((size_t)T, sizeof(void *));
    int64_t *rows = (int64_t *)calloc((size_t)T, sizeof(int64_t)), *drop = (int64_t *)calloc((size_t)T, sizeof(int64_t));
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(T)
#endif
    for (int t = 0; t < T; t++)
        rows[t] = a6_run(lists[t], cnt[t], key, off, ids, score, tok_end, tok_comma, threshold, disable_thresholding,
                         &pc[t], &pn[t], &po[t], &pi[t], &drop[t]);
    int64_t n_out = 0, n_ids = 0, dropped = 0;
    for (int t = 0; t < T; t++) { n_out += rows[t]; n_ids += po[t][rows[t]]; dropped += drop[t]; }
    a6_merge_t *mg = (a6_merge_t *)malloc(sizeof(a6_merge_t) * (size_t)(n_out + 1));
    int64_t w = 0;
    for (int t = 0; t < T; t++) for (int64_t r = 0; r < rows[t]; r++) { mg[w].cell = pc[t][r]; mg[w].shard = t; mg[w].row = r; w++; }
    qsort(mg, (size_t)n_out, sizeof(a6_merge_t), a6_merge_cmp);
    *out_cell = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n_out + 1));
    *out_count = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n_out + 1));
    *out_off = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_out + 2));
    *out_ids = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n_ids + 1));
    int64_t at = 0;
    for (int64_t i = 0; i < n_out; i++) {
        const int t = mg[i].shard; const int64_t r = mg[i].row;
        (*out_cell)[i] = pc[t][r]; (*out_count)[i] = pn[t][r]; (*out_off)[i] = (int32_t)at;
        const int32_t a = po[t][r], b = po[t][r + 1];
        memcpy(*out_ids + at, pi[t] + a, sizeof(uint32_t) * (size_t)(b - a));
        at += b - a;
    }
    (*out_off)[n_out] = (int32_t)at;
    for (int t = 0; t < T; t++) { free(pc[t]); free(pn[t]); free(po[t]); free(pi[t]); }
    free(pc); free(pn); free(po); free(pi); free(rows); free(drop); free(mg); free(lists); free(fill); free(cnt);
    if (dropped_empty) *dropped_empty = dropped;
    return n_out;
}

int32_t orc_sizeof_result(void) { return (int32_t)sizeof(orc_read_result); }
int32_t orc_sizeof_config(void) { return (int32_t)sizeof(orc_config); }
int32_t orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
