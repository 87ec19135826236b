/* ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * C restatement of nimble's whitelist cell-barcode correction ("A5" in SURVEY.md §8a),
 * nimble/fastq_barcode_processor.py:17-36 (build_hamming_index), :73-128 (correct_cell_barcode).
 * Sequential on purpose: the reference's correction_cache makes the FIRST read carrying a raw
 * barcode decide its correction, and this file keeps that loop as it is.
 * Checked against oracle/a5_py.py, which is pinned on tests/golden/a5_cases.json (outputs of the
 * reference's own functions).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CB_SKIPPED 0
#define CB_PERFECT 1
#define CB_CORRECTED 2
#define CB_NONE 3

static const char ALPHA[5] = {'A', 'C', 'G', 'T', 'N'};

static uint64_t fnv(const char *s, int n) {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < n; i++) { h ^= (unsigned char)s[i]; h *= 1099511628211ull; }
    return h ^ (h >> 29);
}

/* string -> int32 map, open addressing; keys are cb_len bytes stored in `keys` */
typedef struct {
    int cb_len;
    uint64_t cap, n;       /* cap = power of two */
    char *keys;
    int32_t *val;
    uint8_t *aux, *used;
} smap;

static void smap_init(smap *m, int cb_len, uint64_t want) {
    uint64_t cap = 16;
    while (cap < want * 2) cap <<= 1;
    m->cb_len = cb_len; m->cap = cap; m->n = 0;
    m->keys = (char *)malloc(cap * (size_t)cb_len);
    m->val = (int32_t *)malloc(cap * sizeof(int32_t));
    m->aux = (uint8_t *)malloc(cap);
    m->used = (uint8_t *)calloc(cap, 1);
}
static void smap_free(smap *m) { free(m->keys); free(m->val); free(m->aux); free(m->used); }

static int64_t smap_find(const smap *m, const char *k) {
    uint64_t s = fnv(k, m->cb_len) & (m->cap - 1);
    while (m->used[s]) {
        if (memcmp(m->keys + s * (size_t)m->cb_len, k, (size_t)m->cb_len) == 0) return (int64_t)s;
        s = (s + 1) & (m->cap - 1);
    }
    return -1;
}
static void smap_put_raw(smap *m, const char *k, int32_t v, uint8_t aux) {
    uint64_t s = fnv(k, m->cb_len) & (m->cap - 1);
    while (m->used[s]) s = (s + 1) & (m->cap - 1);
    m->used[s] = 1;
    memcpy(m->keys + s * (size_t)m->cb_len, k, (size_t)m->cb_len);
    m->val[s] = v; m->aux[s] = aux;
    m->n++;
}
static void smap_put(smap *m, const char *k, int32_t v, uint8_t aux) {
    if ((m->n + 1) * 2 > m->cap) {
        smap old = *m;
        smap_init(m, old.cb_len, old.cap);
        for (uint64_t s = 0; s < old.cap; s++)
            if (old.used[s]) smap_put_raw(m, old.keys + s * (size_t)old.cb_len, old.val[s], old.aux[s]);
        smap_free(&old);
    }
    smap_put_raw(m, k, v, aux);
}

/* wl: n_wl entries of cb_len bytes (same-length whitelist lines), wl_idx[i] = index reported for entry i
 * (first occurrence wins for duplicate lines).  cb / qual: n x cb_len.  eligible may be NULL. */
int32_t orc_cb_correct(const char *wl, const int32_t *wl_idx, int64_t n_wl, int32_t cb_len, const char *cb,
                       const uint8_t *qual, const uint8_t *eligible, int64_t n, int32_t *out_idx, uint8_t *out_status,
                       int64_t *stats4 /* perfect, corrected, none, cache size */) {
    if (cb_len < 1 || cb_len > 64) return -1;
    smap W, C;
    smap_init(&W, cb_len, (uint64_t)n_wl + 1);
    for (int64_t i = 0; i < n_wl; i++)
        if (smap_find(&W, wl + i * (size_t)cb_len) < 0) smap_put(&W, wl + i * (size_t)cb_len, wl_idx[i], 0);
    smap_init(&C, cb_len, 1024);
    int64_t st[4] = {0, 0, 0, 0};
    char var[64], best[64];
    for (int64_t r = 0; r < n; r++) {
        if (eligible && !eligible[r]) { out_idx[r] = -1; out_status[r] = CB_SKIPPED; continue; }
        const char *raw = cb + r * (size_t)cb_len;
        const uint8_t *q = qual + r * (size_t)cb_len;
        int32_t res; uint8_t status;
        int64_t s = smap_find(&C, raw);                               /* :90-91 cache */
        if (s >= 0) { res = C.val[s]; status = C.aux[s]; }
        else {
            int64_t w = smap_find(&W, raw);                           /* :94-96 perfect match */
            if (w >= 0) { res = W.val[w]; status = CB_PERFECT; }
            else {
                /* :99 candidates = whitelist entries one substitution away (raw's base at that position in ACGTN) */
                int n_cand = 0, best_q = 1 << 30;
                int32_t best_idx = -1;
                for (int i = 0; i < cb_len; i++) {
                    const char ch = raw[i];
                    if (!(ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T' || ch == 'N')) continue;
                    for (int a = 0; a < 5; a++) {
                        if (ALPHA[a] == ch) continue;
                        memcpy(var, raw, (size_t)cb_len);
                        var[i] = ALPHA[a];
                        const int64_t v = smap_find(&W, var);
                        if (v < 0) continue;
                        n_cand++;
                        /* :113-125 strictly lowest quality at the differing position; candidates in ascending
                         * string order (SPEC) == on equal quality the smaller string wins */
                        if ((int)q[i] < best_q || ((int)q[i] == best_q && memcmp(var, best, (size_t)cb_len) < 0)) {
                            best_q = q[i]; best_idx = W.val[v]; memcpy(best, var, (size_t)cb_len);
                        }
                    }
                }
                if (n_cand == 0) { res = -1; status = CB_NONE; }      /* :101-103 */
                else { res = best_idx; status = CB_CORRECTED; }       /* :105-109 single, :111-128 several */
            }
            smap_put(&C, raw, res, status);
        }
        out_idx[r] = res; out_status[r] = status;
        st[status == CB_PERFECT ? 0 : (status == CB_CORRECTED ? 1 : 2)]++;
    }
    st[3] = (int64_t)C.n;
    if (stats4) memcpy(stats4, st, sizeof st);
    smap_free(&W); smap_free(&C);
    return 0;
}
