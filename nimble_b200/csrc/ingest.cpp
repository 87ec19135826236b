// Native read ingest and per-read TSV output: FASTQ(.gz) and
// BAM (BGZF, inflated block-parallel, no htslib) -> names / sequences / CB,UB,UR,GN tags ->
// nb200_pack_reads -> nb200_align -> the TSV nimble's `report` consumes
// (nimble/__main__.py:219,237-241).  Used by nb200_align_files and the `aligner` executable.
#include "ingest.hpp"

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace nb200 {

static bool ends_with(const std::string &s, const char *suf) {
    const size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

// whole (optionally gzip-compressed) file -> memory; gzread handles plain files transparently
static void slurp_gz(const std::string &path, std::string &out) {
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) throw IoError("cannot open " + path);
    gzbuffer(f, 1 << 20);
    out.clear();
    std::vector<char> buf(1 << 22);
    for (;;) {
        int n = gzread(f, buf.data(), (unsigned)buf.size());
        if (n < 0) { gzclose(f); throw IoError("read error in " + path); }
        if (n == 0) break;
        out.append(buf.data(), (size_t)n);
    }
    gzclose(f);
}

void slurp_maybe_gz(const std::string &path, std::string &out) { slurp_gz(path, out); }

void Arena::add(const char *p, size_t n) {
    data.append(p, n);
    off.push_back((int64_t)data.size());
}

static void read_fastq(const std::string &path, Arena &names, Arena &seqs) {
    std::string text;
    slurp_gz(path, text);
    const char *p = text.data(), *end = p + text.size();
    while (p < end) {
        const char *l1 = (const char *)memchr(p, '\n', end - p);
        if (!l1) break;
        const char *l2 = (const char *)memchr(l1 + 1, '\n', end - (l1 + 1));
        if (!l2) l2 = end;
        if (*p != '@') throw std::runtime_error("malformed FASTQ record in " + path);
        const char *ne = p + 1;
        while (ne < l1 && *ne != ' ' && *ne != '\t' && *ne != '\r') ne++;
        names.add(p + 1, ne - (p + 1));
        const char *se = l2;
        while (se > l1 + 1 && (se[-1] == '\r' || se[-1] == '\n')) se--;
        seqs.add(l1 + 1, se - (l1 + 1));
        // skip '+' line and quality line
        const char *q = l2 < end ? l2 + 1 : end;
        for (int k = 0; k < 2 && q < end; k++) {
            const char *nl = (const char *)memchr(q, '\n', end - q);
            q = nl ? nl + 1 : end;
        }
        p = q;
    }
}

// ---- BAM -----------------------------------------------------------------------------------------
static void inflate_bgzf(const std::string &path, int threads, std::string &out) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) throw IoError("cannot open " + path);
    std::string raw;
    {
        std::vector<char> buf(1 << 22);
        size_t n;
        while ((n = fread(buf.data(), 1, buf.size(), f)) > 0) raw.append(buf.data(), n);
        fclose(f);
    }
    struct Blk { size_t off, clen, uoff, ulen; };
    std::vector<Blk> blks;
    size_t p = 0, utotal = 0;
    bool bgzf = true;
    while (p + 18 <= raw.size()) {
        const unsigned char *h = (const unsigned char *)raw.data() + p;
        if (h[0] != 31 || h[1] != 139 || !(h[3] & 4)) { bgzf = false; break; }
        const size_t xlen = h[10] | (h[11] << 8);
        size_t bsize = 0, q = 12;
        while (q + 4 <= 12 + xlen) {
            const size_t slen = h[q + 2] | (h[q + 3] << 8);
            if (h[q] == 'B' && h[q + 1] == 'C' && slen == 2) bsize = (size_t)(h[q + 4] | (h[q + 5] << 8)) + 1;
            q += 4 + slen;
        }
        if (!bsize || p + bsize > raw.size()) { bgzf = false; break; }
        const unsigned char *t = h + bsize - 4;
        const size_t isize = t[0] | (t[1] << 8) | (t[2] << 16) | ((size_t)t[3] << 24);
        blks.push_back({p + 12 + xlen, bsize - 12 - xlen - 8, utotal, isize});
        utotal += isize;
        p += bsize;
    }
    if (!bgzf || blks.empty()) {       // plain gzip (e.g. written by python's gzip module): one stream
        slurp_gz(path, out);
        return;
    }
    out.assign(utotal, '\0');
    std::atomic<size_t> next{0};
    std::atomic<int> bad{0};
    auto work = [&] {
        z_stream zs;
        for (size_t i; (i = next.fetch_add(1)) < blks.size();) {
            if (!blks[i].ulen) continue;
            memset(&zs, 0, sizeof(zs));
            if (inflateInit2(&zs, -15) != Z_OK) { bad = 1; return; }
            zs.next_in = (Bytef *)raw.data() + blks[i].off; zs.avail_in = (uInt)blks[i].clen;
            zs.next_out = (Bytef *)&out[blks[i].uoff]; zs.avail_out = (uInt)blks[i].ulen;
            if (inflate(&zs, Z_FINISH) != Z_STREAM_END) bad = 1;
            inflateEnd(&zs);
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < std::max(1, threads); t++) th.emplace_back(work);
    for (auto &x : th) x.join();
    if (bad) throw IoError("corrupt BGZF block in " + path);
}

static inline uint32_t le32(const unsigned char *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

static void read_bam(const std::string &path, int threads, ReadSet &R) {
    std::string buf;
    inflate_bgzf(path, threads, buf);
    const unsigned char *b = (const unsigned char *)buf.data();
    const size_t n = buf.size();
    if (n < 12 || memcmp(b, "BAM\1", 4) != 0) throw std::runtime_error(path + " is not a BAM file");
    size_t p = 4;
    p += 4 + le32(b + p);
    const uint32_t n_ref = le32(b + p); p += 4;
    for (uint32_t i = 0; i < n_ref && p + 4 <= n; i++) p += 4 + le32(b + p) + 4;
    static const char code[] = "=ACMGRSVTWYHKDBN";
    struct Mate { std::string seq; std::string tag[4]; int64_t pos = -1; bool have = false; };
    // records of a pair are adjacent in unaligned / name-sorted BAMs; keep a small pending map keyed by name
    bool any_paired = false;
    struct Out { std::string name; Mate m[2]; };
    std::vector<Out> outs;
    auto find_pending = [&](const std::string &name) -> int {
        for (int i = (int)outs.size() - 1, k = 0; i >= 0 && k < 8; i--, k++) if (outs[i].name == name) return i;
        return -1;
    };
    while (p + 4 <= n) {
        const uint32_t bs = le32(b + p); p += 4;
        if (p + bs > n || bs < 32) throw std::runtime_error("truncated BAM record in " + path);
        const unsigned char *r = b + p, *end = r + bs;
        const int64_t pos = (int32_t)le32(r + 4);
        const uint32_t l_name = r[8];
        const uint32_t n_cigar = r[12] | (r[13] << 8);
        const uint32_t flag = r[14] | (r[15] << 8);
        const uint32_t l_seq = le32(r + 16);
        const unsigned char *q = r + 32;
        std::string name((const char *)q, l_name ? l_name - 1 : 0); q += l_name;
        q += 4 * (size_t)n_cigar;
        Mate m;
        m.seq.resize(l_seq);
        for (uint32_t i = 0; i < l_seq; i++) m.seq[i] = code[(q[i >> 1] >> ((i & 1) ? 0 : 4)) & 15];
        q += (l_seq + 1) / 2 + l_seq;
        while (q + 3 <= end) {
            const char t0 = (char)q[0], t1 = (char)q[1], ty = (char)q[2]; q += 3;
            size_t adv = 0;
            const char *zs = nullptr;
            switch (ty) {
            case 'Z': case 'H': { zs = (const char *)q; adv = strnlen(zs, end - q) + 1; break; }
            case 'A': case 'c': case 'C': adv = 1; break;
            case 's': case 'S': adv = 2; break;
            case 'i': case 'I': case 'f': adv = 4; break;
            case 'B': { const char sub = (char)q[0]; const uint32_t cnt = le32(q + 1);
                        const int es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4; adv = 5 + (size_t)cnt * es; break; }
            default: throw std::runtime_error("unknown BAM tag type in " + path);
            }
            if (zs) {
                int slot = -1;
                if (t0 == 'C' && t1 == 'B') slot = 0; else if (t0 == 'U' && t1 == 'B') slot = 1;
                else if (t0 == 'U' && t1 == 'R') slot = 2; else if (t0 == 'G' && t1 == 'N') slot = 3;
                if (slot >= 0) m.tag[slot].assign(zs, adv - 1);
            }
            q += adv;
        }
        p += bs;
        if (flag & 0x900) continue;                         // secondary / supplementary
        if (flag & 0x10) {                                  // stored reverse-complemented: restore the read as sequenced
            std::reverse(m.seq.begin(), m.seq.end());
            for (char &c : m.seq) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
        }
        m.pos = pos + 1; m.have = true;
        const int which = (flag & 0x80) ? 1 : 0;
        if (which) any_paired = true;
        int i = find_pending(name);
        if (i < 0 || outs[i].m[which].have) { outs.push_back(Out{name, {}}); i = (int)outs.size() - 1; }
        outs[i].m[which] = std::move(m);
    }
    R.paired = any_paired;
    R.has_tags = true;
    for (auto &o : outs) {
        Mate &a = (o.m[0].have || any_paired) ? o.m[0] : o.m[1];
        R.names.add(o.name.data(), o.name.size());
        R.r1.add(a.seq.data(), a.seq.size());
        if (any_paired) R.r2.add(o.m[1].seq.data(), o.m[1].seq.size());
        R.cb.add(a.tag[0].data(), a.tag[0].size());
        const std::string &ub = a.tag[1].empty() ? a.tag[2] : a.tag[1];
        R.ub.add(ub.data(), ub.size());
        R.ur.add(a.tag[2].data(), a.tag[2].size());
        R.gn.add(a.tag[3].data(), a.tag[3].size());
        R.pos1.push_back(a.pos);
        R.pos2.push_back(any_paired && o.m[1].have ? o.m[1].pos : -1);
    }
}

void load_reads(const std::vector<std::string> &inputs, int threads, ReadSet &R) {
    if (inputs.empty() || inputs.size() > 2) throw std::runtime_error("expected one or two --input files");
    std::string lower = inputs[0];
    std::transform(lower.begin(), lower.end(), lower.begin(), ::tolower);
    if (ends_with(lower, ".bam")) { read_bam(inputs[0], threads, R); return; }
    read_fastq(inputs[0], R.names, R.r1);
    if (inputs.size() == 2) {
        Arena n2;
        read_fastq(inputs[1], n2, R.r2);
        if (R.r2.size() != R.r1.size()) throw std::runtime_error("R1 and R2 FASTQ files hold different numbers of reads");
        R.paired = true;
    }
}

// ---- TSV output ------------------------------------------------------------------------------------
struct Sink {
    FILE *f = nullptr;
    gzFile g = nullptr;
    std::string buf;
    void open(const std::string &path, bool gz) {
        if (gz) { g = gzopen(path.c_str(), "wb"); if (!g) throw IoError("cannot write " + path); gzbuffer(g, 1 << 20); }
        else { f = fopen(path.c_str(), "wb"); if (!f) throw IoError("cannot write " + path); }
    }
    void flush() {
        if (buf.empty()) return;
        if (g) { if (gzwrite(g, buf.data(), (unsigned)buf.size()) <= 0) throw IoError("write failed"); }
        else if (fwrite(buf.data(), 1, buf.size(), f) != buf.size()) throw IoError("write failed");
        buf.clear();
    }
    void put(const std::string &s) { buf += s; if (buf.size() > (1 << 22)) flush(); }
    void close() { flush(); if (g) gzclose(g); if (f) fclose(f); g = nullptr; f = nullptr; }
    ~Sink() { if (g) gzclose(g); if (f) fclose(f); }
};

void write_per_read_tsv(const std::string &out_path, const ReadSet &R, const nb200_read_result *res, const int32_t *feats,
                        int max_hits, const std::vector<std::string> &feature_names) {
    const std::string tmp = out_path + ".tmp";
    Sink s;
    s.open(tmp, ends_with(out_path, ".gz"));
    s.put("nimble_features\tnimble_score\tr1_forward_score\tr1_reverse_score\tr2_forward_score\tr2_reverse_score\t"
          "r1_QNAME\tr1_CB\tr1_UB\tr1_UR\tr1_GN\tr1_POS\tr2_POS\n");
    std::string line;
    const size_t n = R.r1.size();
    for (size_t i = 0; i < n; i++) {
        if (!res[i].n_feat) continue;
        line.clear();
        for (int j = 0; j < res[i].n_feat; j++) { if (j) line += ','; line += feature_names[feats[i * (size_t)max_hits + j]]; }
        line += "\t1";
        for (int o = 0; o < 4; o++) { line += '\t'; line += std::to_string(res[i].score[o]); }
        line += '\t'; line.append(R.names.ptr(i), R.names.len(i));
        line += '\t'; line.append(R.cb.ptr(i), R.cb.len(i));
        line += '\t'; line.append(R.ub.ptr(i), R.ub.len(i));
        line += '\t'; line.append(R.ur.ptr(i), R.ur.len(i));
        line += '\t'; line.append(R.gn.ptr(i), R.gn.len(i));
        line += '\t'; if (R.pos1[i] > 0) line += std::to_string(R.pos1[i]);
        line += '\t'; if (R.pos2[i] > 0) line += std::to_string(R.pos2[i]);
        line += '\n';
        s.put(line);
    }
    s.close();
    if (rename(tmp.c_str(), out_path.c_str()) != 0) throw IoError("cannot rename " + tmp);
}

void write_bulk_tsv(const std::string &out_path, const nb200_counts &c, const std::vector<std::string> &feature_names) {
    const std::string tmp = out_path + ".tmp";
    Sink s;
    s.open(tmp, ends_with(out_path, ".gz"));
    s.put("nimble_features\tnimble_score\n");
    std::string line;
    for (uint64_t i = 0; i < c.n_rows; i++) {
        line.clear();
        for (uint32_t j = c.feat_off[i]; j < c.feat_off[i + 1]; j++) { if (j > c.feat_off[i]) line += ','; line += feature_names[c.feat_ids[j]]; }
        line += '\t'; line += std::to_string(c.count[i]); line += '\n';
        s.put(line);
    }
    s.close();
    if (rename(tmp.c_str(), out_path.c_str()) != 0) throw IoError("cannot rename " + tmp);
}

}  // namespace nb200
