// Raw DEFLATE decoder for BGZF blocks (RFC 1951), written for the streaming reader's inflate tasks: they are three
// quarters of the host work per read, and a BGZF block is small (<= 64 KB), whole in memory on both sides and of known
// inflated size — so the decoder keeps a 64-bit bit buffer refilled by unaligned 8-byte loads, decodes with one table
// lookup per symbol (11-bit primary table for literals/lengths, 8-bit for distances, subtables for longer codes) and
// copies matches 8 bytes at a time, with a careful byte-wise loop only near the ends of the buffers.
//
// Contract: fast_inflate() returns true iff the stream decoded without anomaly into EXACTLY out_len bytes.  On false
// the caller falls back to zlib (which also produces the error message for really corrupt input); the decoder never
// reads outside [in, in + in_len) nor writes outside [out, out + out_len).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace nb200 {
namespace fi {

constexpr int kLitBits = 11, kDistBits = 8, kPreBits = 7;
constexpr uint32_t kLit = 1u << 10, kEob = 1u << 11, kSub = 1u << 12;      // entry flags; entry 0 = invalid code
// entry: bits 0-4 total code length, bits 5-9 extra bits (or subtable bits for kSub), bits 10-12 flags, bits 16-31 payload
// (literal byte / base length / base distance / subtable start)
static inline int e_len(uint32_t e) { return (int)(e & 0x1F); }
static inline int e_extra(uint32_t e) { return (int)((e >> 5) & 0x1F); }
static inline uint32_t e_payload(uint32_t e) { return e >> 16; }

struct Tables {
    uint32_t lit[(1 << kLitBits) + 288 * 16];
    uint32_t dist[(1 << kDistBits) + 32 * 128];
};

static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

static inline uint32_t bitrev(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; i++) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}

// canonical Huffman code lengths -> decode table.  kind: 0 literal/length, 1 distance, 2 precode (payload = symbol).
// Returns false for an over-subscribed code or a symbol outside the alphabet; incomplete codes leave entries 0.
static inline bool build_table(const uint8_t *lens, int n_sym, int kind, int P, uint32_t *table, size_t cap) {
    uint16_t count[16] = {0}, next[16];
    for (int s = 0; s < n_sym; s++) count[lens[s]]++;
    count[0] = 0;
    uint32_t kraft = 0;
    for (int l = 1; l <= 15; l++) kraft += (uint32_t)count[l] << (15 - l);
    if (kraft > (1u << 15)) return false;
    if (kraft < (1u << 15)) {                        // incomplete code: zlib accepts it only as "no code at all" for distances or one
        int max_len = 0;                             // single 1-bit code (literal/length, distance); follow it, so that damaged
        for (int l = 1; l <= 15; l++) if (count[l]) max_len = l;      // streams are declined alike
        if (!(max_len == 0 && kind == 1) && !(max_len == 1 && kind != 2)) return false;
    }
    uint32_t code = 0;
    for (int l = 1; l <= 15; l++) { code = (code + count[l - 1]) << 1; next[l] = (uint16_t)code; }
    memset(table, 0, sizeof(uint32_t) << P);
    // subtable sizes: longest code below every primary prefix
    uint8_t sub_bits[1 << kLitBits];
    bool any_long = false;
    for (int l = P + 1; l <= 15; l++) any_long |= count[l] != 0;
    if (any_long) memset(sub_bits, 0, (size_t)1 << P);
    uint16_t codes[288];
    for (int s = 0; s < n_sym; s++) {
        const int l = lens[s];
        if (!l) continue;
        codes[s] = (uint16_t)bitrev(next[l]++, l);
        if (l > P) { uint8_t &b = sub_bits[codes[s] & ((1u << P) - 1)]; if (l - P > b) b = (uint8_t)(l - P); }
    }
    size_t used = (size_t)1 << P;
    if (any_long)
        for (uint32_t p = 0; p < (1u << P); p++)
            if (sub_bits[p]) {
                if (used + ((size_t)1 << sub_bits[p]) > cap) return false;
                table[p] = kSub | ((uint32_t)used << 16) | ((uint32_t)sub_bits[p] << 5) | (uint32_t)P;
                memset(table + used, 0, sizeof(uint32_t) << sub_bits[p]);
                used += (size_t)1 << sub_bits[p];
            }
    for (int s = 0; s < n_sym; s++) {
        const int l = lens[s];
        if (!l) continue;
        uint32_t e;
        if (kind == 0) {
            if (s < 256) e = kLit | ((uint32_t)s << 16);
            else if (s == 256) e = kEob;
            else if (s <= 285) e = ((uint32_t)kLenBase[s - 257] << 16) | ((uint32_t)kLenExtra[s - 257] << 5);
            else continue;                           // 286, 287: never valid in data; left invalid
        } else if (kind == 1) {
            if (s >= 30) continue;
            e = ((uint32_t)kDistBase[s] << 16) | ((uint32_t)kDistExtra[s] << 5);
        } else e = (uint32_t)s << 16;
        e |= (uint32_t)l;
        const uint32_t c = codes[s];
        if (l <= P) {
            for (uint32_t i = c; i < (1u << P); i += 1u << l) table[i] = e;
        } else {
            const uint32_t pe = table[c & ((1u << P) - 1)];
            uint32_t *sub = table + e_payload(pe);
            const int sb = e_extra(pe);
            for (uint32_t i = c >> P; i < (1u << sb); i += 1u << (l - P)) sub[i] = e;
        }
    }
    return true;
}

struct Bits {
    const uint8_t *p, *end;
    uint64_t buf = 0;
    int cnt = 0;
    inline void refill() {
        if (end - p >= 8) {
            uint64_t v;
            memcpy(&v, p, 8);
            buf |= v << cnt;
            p += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56 && p < end) { buf |= (uint64_t)*p++ << cnt; cnt += 8; }
        }
    }
    inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
    inline void drop(int n) { buf >>= n; cnt -= n; }
};

}  // namespace fi

static inline bool fast_inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) {
    using namespace fi;
    static thread_local Tables T;
    static thread_local Tables Tfixed;
    static thread_local bool fixed_ready = false;
    Bits b;
    b.p = in; b.end = in + in_len;
    uint8_t *o = out, *const o_end = out + out_len;
    for (;;) {
        b.refill();
        if (b.cnt < 3) return false;
        const uint32_t bfinal = b.peek(1);
        b.drop(1);
        const uint32_t btype = b.peek(2);
        b.drop(2);
        const Tables *tab = &T;
        if (btype == 0) {                            // stored: back to the byte boundary, LEN / NLEN, bytes
            b.drop(b.cnt & 7);
            b.p -= b.cnt >> 3;                       // whole bytes still in the buffer go back
            b.buf = 0; b.cnt = 0;
            if (b.end - b.p < 4) return false;
            const uint32_t len = b.p[0] | (b.p[1] << 8), nlen = b.p[2] | (b.p[3] << 8);
            b.p += 4;
            if ((len ^ 0xFFFFu) != nlen || (size_t)(b.end - b.p) < len || (size_t)(o_end - o) < len) return false;
            memcpy(o, b.p, len);
            o += len; b.p += len;
            if (bfinal) break;
            continue;
        } else if (btype == 1) {
            if (!fixed_ready) {
                uint8_t l[288], d[32];
                for (int i = 0; i < 144; i++) l[i] = 8;
                for (int i = 144; i < 256; i++) l[i] = 9;
                for (int i = 256; i < 280; i++) l[i] = 7;
                for (int i = 280; i < 288; i++) l[i] = 8;
                for (int i = 0; i < 32; i++) d[i] = 5;
                if (!build_table(l, 288, 0, kLitBits, Tfixed.lit, sizeof Tfixed.lit / 4) || !build_table(d, 32, 1, kDistBits, Tfixed.dist, sizeof Tfixed.dist / 4)) return false;
                fixed_ready = true;
            }
            tab = &Tfixed;
        } else if (btype == 2) {
            b.refill();
            if (b.cnt < 14) return false;
            const int hlit = (int)b.peek(5) + 257; b.drop(5);
            const int hdist = (int)b.peek(5) + 1; b.drop(5);
            const int hclen = (int)b.peek(4) + 4; b.drop(4);
            if (hlit > 286 || hdist > 30) return false;
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t pl[19] = {0};
            for (int i = 0; i < hclen; i++) {
                if (b.cnt < 3) { b.refill(); if (b.cnt < 3) return false; }
                pl[order[i]] = (uint8_t)b.peek(3); b.drop(3);
            }
            uint32_t pre[1 << kPreBits];
            if (!build_table(pl, 19, 2, kPreBits, pre, 1 << kPreBits)) return false;
            uint8_t lens[288 + 32];
            int n = 0;
            const int total = hlit + hdist;
            while (n < total) {
                b.refill();
                const uint32_t e = pre[b.peek(kPreBits)];
                const int l = e_len(e);
                if (!l || b.cnt < l + 7) return false;
                b.drop(l);
                const int sym = (int)e_payload(e);
                if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
                int rep; uint8_t v = 0;
                if (sym == 16) { if (!n) return false; v = lens[n - 1]; rep = 3 + (int)b.peek(2); b.drop(2); }
                else if (sym == 17) { rep = 3 + (int)b.peek(3); b.drop(3); }
                else { rep = 11 + (int)b.peek(7); b.drop(7); }
                if (n + rep > total) return false;
                memset(lens + n, v, (size_t)rep);
                n += rep;
            }
            if (lens[256] == 0) return false;        // no end-of-block code
            uint8_t ll[288] = {0}, dl[32] = {0};
            memcpy(ll, lens, (size_t)hlit);
            memcpy(dl, lens + hlit, (size_t)hdist);
            if (!build_table(ll, 288, 0, kLitBits, T.lit, sizeof T.lit / 4) || !build_table(dl, 32, 1, kDistBits, T.dist, sizeof T.dist / 4)) return false;
        } else return false;

        // ---- symbols of the block ----
        const uint32_t *lt = tab->lit, *dt = tab->dist;
        bool block_done = false;
        // Fast loop: while at least 16 input bytes and 320 output bytes remain, no per-symbol bounds checks are needed —
        // one refill (56+ valid bits) covers the longest symbol sequence handled per iteration (literal 15 + literal 15, or
        // length 15 + 5, distance 15 + 13 = 48 bits), a match writes at most 258 + 15 bytes, and the next table entry is
        // looked up before the current symbol is finished.
        if ((size_t)(b.end - b.p) >= 16 && (size_t)(o_end - o) >= 320) {
            const uint8_t *const in_fast = b.end - 16;
            uint8_t *const out_fast = o_end - 320;
            b.refill();
            uint32_t e = lt[b.peek(kLitBits)];
            do {
                if (e & kSub) e = lt[e_payload(e) + ((b.buf >> kLitBits) & ((1u << e_extra(e)) - 1))];
                if (!e_len(e)) return false;
                b.drop(e_len(e));
                if (e & kLit) {
                    *o++ = (uint8_t)(e >> 16);
                    e = lt[b.peek(kLitBits)];
                    if (e & kLit) {                                        // a second literal out of the same refill (41+ bits left)
                        b.drop(e_len(e));
                        *o++ = (uint8_t)(e >> 16);
                        e = lt[b.peek(kLitBits)];
                        if (e & kLit) {                                    // and a third (26+ bits left, a direct literal code has <= 11)
                            b.drop(e_len(e));
                            *o++ = (uint8_t)(e >> 16);
                            b.refill();
                            e = lt[b.peek(kLitBits)];
                            continue;
                        }
                    }
                    b.refill();
                    e = lt[b.peek(kLitBits)];
                    continue;
                }
                if (e & kEob) { block_done = true; break; }
                uint32_t len = e_payload(e) + b.peek(e_extra(e));
                b.drop(e_extra(e));
                uint32_t d = dt[b.peek(kDistBits)];
                if (d & kSub) d = dt[e_payload(d) + ((b.buf >> kDistBits) & ((1u << e_extra(d)) - 1))];
                if (!e_len(d)) return false;
                b.drop(e_len(d));
                const uint32_t dist = e_payload(d) + b.peek(e_extra(d));
                b.drop(e_extra(d));
                b.refill();
                e = lt[b.peek(kLitBits)];                                  // next symbol's entry, before the copy
                if (dist > (size_t)(o - out)) return false;
                const uint8_t *src = o - dist;
                uint8_t *dst = o;
                o += len;
                if (dist >= 8) {
                    uint64_t w0, w1;
                    memcpy(&w0, src, 8);
                    if (dist >= 16) {
                        memcpy(&w1, src + 8, 8);
                        memcpy(dst, &w0, 8); memcpy(dst + 8, &w1, 8);
                        dst += 16; src += 16;
                        while (dst < o) { memcpy(&w0, src, 8); memcpy(&w1, src + 8, 8); memcpy(dst, &w0, 8); memcpy(dst + 8, &w1, 8); dst += 16; src += 16; }
                    } else {
                        do { memcpy(&w0, src, 8); memcpy(dst, &w0, 8); dst += 8; src += 8; } while (dst < o);
                    }
                } else if (dist == 1) {
                    memset(dst, *src, len);
                } else {
                    do { *dst++ = *src++; } while (dst < o);
                }
            } while (b.p <= in_fast && o <= out_fast);
            // the entry looked up ahead is dropped: the careful loop below starts from the bit buffer again
        }
        while (!block_done) {
            b.refill();
            uint32_t e = lt[b.peek(kLitBits)];
            if (e & kSub) e = lt[e_payload(e) + ((b.buf >> kLitBits) & ((1u << e_extra(e)) - 1))];
            int l = e_len(e);
            if (!l || b.cnt < l) return false;
            b.drop(l);
            if (e & kLit) {
                if (o >= o_end) return false;
                *o++ = (uint8_t)(e >> 16);
                continue;
            }
            if (e & kEob) break;
            const int le = e_extra(e);
            uint32_t len = e_payload(e);
            if (b.cnt < le + 15 + 13) { b.refill(); }
            if (b.cnt < le) return false;
            len += b.peek(le); b.drop(le);
            uint32_t d = dt[b.peek(kDistBits)];
            if (d & kSub) d = dt[e_payload(d) + ((b.buf >> kDistBits) & ((1u << e_extra(d)) - 1))];
            l = e_len(d);
            if (!l || b.cnt < l) return false;
            b.drop(l);
            const int de = e_extra(d);
            if (b.cnt < de) { b.refill(); if (b.cnt < de) return false; }
            const uint32_t dist = e_payload(d) + b.peek(de);
            b.drop(de);
            if (dist > (size_t)(o - out) || len > (size_t)(o_end - o)) return false;
            const uint8_t *src = o - dist;
            for (uint32_t i = 0; i < len; i++) o[i] = src[i];
            o += len;
        }
        if (bfinal) break;
    }
    return o == o_end;
}

}  // namespace nb200
