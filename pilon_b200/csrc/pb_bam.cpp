// pb_bam.cpp -- native BAM ingest (and a BAM / BAI / FASTA writer for synthetic inputs), host side of libpilonb200.so.
//
// The reference reads its BAMs through htsjdk 2.23.0 (un-vendored): BamFile.process asks the reader for the records that
// overlap the region +-10 kb, drops those that fail validateRead and hands the rest to PileUpRegion.addRead one by one
// (BamFile.scala:101-148).  pb_bam_query_pack is that loop for the engine: BGZF inflate, BAI linear-index seek, BAM
// record decode, validateRead, and straight into a pb_packer -- a BAM record (l_seq, n_cigar_op, uint32 cigar, 4-bit seq,
// qual) is already most of a pb_batch row.  Formats restated from the SAM/BAM specification (SAMv1,
// sections 4.1 BGZF, 4.2 BAM, 5.2 BAI); htsjdk semantics restated where the reference depends on them:
//   * SAMRecord.getAlignmentStart = pos + 1; getAlignmentEnd = start + (reference length of the CIGAR) - 1, 0 if unmapped;
//   * queryOverlapping(name, start, end): records whose [alignmentStart, alignmentEnd] meets [start, end] (1-based,
//     inclusive; a record without an alignment end is treated as one base long), in file order;
//   * bases decode through "=ACMGRSVTWYHKDBN"; a quality array of 0xFF bytes means "no qualities".
// The writer exists so that synthetic inputs can be given to the real Pilon JVM wherever one is available
// (tools/run_real_pilon.sh) and so that the reader can be round-trip tested here.
#include <chrono>
#include <future>
#include <thread>
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/pilon_b200.h"

namespace {

thread_local std::string g_err_bam;
int fail_bam(int code, const std::string& msg) { g_err_bam = msg; return code; }

inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }
inline void wr32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 0; i < 4; i++) v.push_back((uint8_t)(x >> (8 * i))); }
inline void wr64(std::vector<uint8_t>& v, uint64_t x) { for (int i = 0; i < 8; i++) v.push_back((uint8_t)(x >> (8 * i))); }

// ---- BGZF reader: random access by virtual offset (compressed block start << 16 | offset in the inflated block) ----
struct Bgzf {
    FILE* f = nullptr;
    std::vector<uint8_t> block;        // inflated current block
    uint64_t block_addr = 0;           // file offset of the current block
    uint32_t block_csize = 0;          // its compressed size
    size_t at = 0;                     // read position inside `block`
    bool eof = false;
    // Read-ahead window: the next blocks' compressed bytes are read in one sequential sweep and inflated by a few threads
    // (BGZF blocks are independent deflate streams); load() then serves consecutive blocks out of the window.  Inflate was
    // 60 % of a query's time on one thread.
    struct Ahead { uint64_t addr = 0; uint32_t csize = 0; size_t data_len = 0; uint32_t isize = 0; bool ok = false, end = false, bad = false, live = false;
                   std::vector<uint8_t> cdata, out; };
    std::vector<Ahead> win; size_t win_n = 0;
    std::vector<Ahead> nxt; size_t nxt_n = 0;          // the window after `win`, being read and inflated while `win` is consumed
    std::future<bool> pending; bool has_pending = false; uint64_t nxt_addr = 0;
    int threads = 4;
    static constexpr size_t WINDOW = 48;
    double t_load = 0;                 // PB_BAM_TRACE: seconds inside load()

    // header + compressed payload of the block at `addr` into a; false = corrupt / short file.  a.end = clean end of file.
    bool fetch(uint64_t addr, Ahead& a) {
        a.addr = addr; a.ok = false; a.end = false; a.csize = 0; a.isize = 0; a.data_len = 0;
        if (fseeko(f, (off_t)addr, SEEK_SET) != 0) return false;
        uint8_t hdr[18];
        const size_t got = fread(hdr, 1, 18, f);
        if (got == 0) { a.end = true; return true; }
        if (got != 18 || hdr[0] != 31 || hdr[1] != 139 || hdr[2] != 8 || !(hdr[3] & 4)) return false;
        const unsigned xlen = rd16(hdr + 10);
        // the BC subfield is the first (and in practice only) extra subfield: SI1 = 66, SI2 = 67, SLEN = 2, BSIZE
        std::vector<uint8_t> extra(xlen);
        memcpy(extra.data(), hdr + 12, std::min<size_t>(6, xlen));
        if (xlen > 6 && fread(extra.data() + 6, 1, xlen - 6, f) != xlen - 6) return false;
        int bsize = -1;
        for (size_t i = 0; i + 4 <= extra.size();) {
            const unsigned slen = rd16(extra.data() + i + 2);
            if (extra[i] == 66 && extra[i + 1] == 67 && slen == 2) bsize = rd16(extra.data() + i + 4);
            i += 4 + slen;
        }
        if (bsize < 0) return false;
        const size_t csize = (size_t)bsize + 1;
        if (csize < 12 + (size_t)xlen + 8) return false;                 // corrupt block header
        a.data_len = csize - 12 - xlen - 8;
        a.cdata.resize(a.data_len + 8);
        if (xlen <= 6) {                                     // part of the payload may already sit in hdr (never: xlen == 6 exactly)
            if (fseeko(f, (off_t)(addr + 12 + xlen), SEEK_SET) != 0) return false;
        }
        if (fread(a.cdata.data(), 1, a.data_len + 8, f) != a.data_len + 8) return false;
        a.isize = rd32(a.cdata.data() + a.data_len + 4);
        if (a.isize > 65536u) return false;                              // a BGZF block inflates to at most 64 KiB
        a.csize = (uint32_t)csize;
        return true;
    }
    static void inflate_one(Ahead& a) {
        a.out.resize(a.isize);
        a.ok = true;
        if (!a.isize) return;
        z_stream zs; memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -15) != Z_OK) { a.ok = false; return; }
        zs.next_in = a.cdata.data(); zs.avail_in = (uInt)a.data_len; zs.next_out = a.out.data(); zs.avail_out = a.isize;
        const int rc = inflate(&zs, Z_FINISH);
        inflateEnd(&zs);
        if (rc != Z_STREAM_END || crc32(crc32(0L, Z_NULL, 0), a.out.data(), a.isize) != rd32(a.cdata.data() + a.data_len)) a.ok = false;
    }
    // window = the blocks from `addr` on; the first one must be readable, later failures surface when they are reached
    bool fill(std::vector<Ahead>& W, size_t& n, uint64_t addr) {
        if (W.size() < WINDOW) W.resize(WINDOW);
        n = 0;
        uint64_t a = addr;
        while (n < WINDOW) {
            Ahead& w = W[n];
            const bool got = fetch(a, w);
            if (!got) { if (n == 0) return false; w.addr = a; w.ok = false; w.end = false; w.csize = 0; w.bad = true; n++; break; }
            w.bad = false;
            n++;
            if (w.end) break;
            a += w.csize;
        }
        size_t n_inf = 0;
        for (size_t k = 0; k < n; k++) if (!W[k].end && W[k].csize) n_inf++;
        const int nt = (int)std::min<size_t>((size_t)std::max(1, threads), n_inf);
        if (nt <= 1) { for (size_t k = 0; k < n; k++) if (!W[k].end && W[k].csize) inflate_one(W[k]); }
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; t++) th.emplace_back([&W, n, nt, t]() { for (size_t k = (size_t)t; k < n; k += (size_t)nt) if (!W[k].end && W[k].csize) inflate_one(W[k]); });
            for (auto& x : th) x.join();
        }
        return true;
    }
    void drain() { if (has_pending) { pending.get(); has_pending = false; } }       // nobody else may touch `f` while a prefetch runs
    void prefetch_after_win() {
        if (win_n == 0) return;
        const Ahead& last = win[win_n - 1];
        if (last.end || last.bad || !last.csize) return;
        nxt_addr = last.addr + last.csize;
        pending = std::async(std::launch::async, [this]() { return fill(nxt, nxt_n, nxt_addr); });
        has_pending = true;
    }
    ~Bgzf() { drain(); }
    bool load(uint64_t addr) {
        const auto t0 = std::chrono::steady_clock::now();
        const bool ok = load_(addr);
        t_load += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return ok;
    }
    bool load_(uint64_t addr) {
        size_t k = 0;
        for (; k < win_n; k++) if (win[k].addr == addr && (win[k].end || win[k].bad || win[k].live)) break;
        if (k == win_n) {
            bool have = false;
            if (has_pending) {                               // the sequential case: the next window is (being) prepared
                const bool okp = pending.get(); has_pending = false;
                if (okp && nxt_n && nxt_addr == addr) { win.swap(nxt); win_n = nxt_n; nxt_n = 0; have = true; }
            }
            if (!have && !fill(win, win_n, addr)) { win_n = 0; return false; }
            for (size_t i = 0; i < win_n; i++) win[i].live = !win[i].end && !win[i].bad;
            prefetch_after_win();
            k = 0;
        }
        Ahead& a = win[k];
        if (a.end) { eof = true; block.clear(); at = 0; block_addr = addr; block_csize = 0; return true; }
        if (a.bad || !a.ok) return false;
        block.swap(a.out);
        block_addr = addr; block_csize = a.csize; at = 0; eof = false;
        a.live = false;                                      // served: its bytes now live in `block`
        return true;
    }
    bool seek(uint64_t voff) {
        if (!load(voff >> 16)) return false;
        at = (size_t)(voff & 0xFFFF);
        return at <= block.size();
    }
    // reads exactly n bytes; false at a clean end of file (n == 0 bytes read) or on error (*err set)
    bool read(uint8_t* dst, size_t n, bool* err) {
        size_t done = 0;
        while (done < n) {
            if (at == block.size()) {
                if (eof) { if (done) *err = true; return false; }
                if (!load(block_addr + block_csize)) { *err = true; return false; }
                if (eof) { if (done) *err = true; return false; }
                continue;
            }
            const size_t k = std::min(n - done, block.size() - at);
            memcpy(dst + done, block.data() + at, k);
            done += k; at += k;
        }
        return true;
    }
};

// ---- BGZF writer -----------------------------------------------------------------------------------
struct BgzfWriter {
    FILE* f = nullptr;
    std::vector<uint8_t> buf;          // pending uncompressed bytes of the current block
    uint64_t written = 0;              // compressed bytes written = file offset of the current block
    static constexpr size_t BLOCK = 0xFF00;

    uint64_t tell() const { return (written << 16) | (uint64_t)buf.size(); }
    bool flush_block() {
        std::vector<uint8_t> out(buf.size() + buf.size() / 100 + 64 + 26);
        z_stream zs; memset(&zs, 0, sizeof zs);
        if (deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
        zs.next_in = buf.data(); zs.avail_in = (uInt)buf.size(); zs.next_out = out.data() + 18; zs.avail_out = (uInt)(out.size() - 26);
        const int rc = deflate(&zs, Z_FINISH);
        const size_t clen = zs.total_out;
        deflateEnd(&zs);
        if (rc != Z_STREAM_END) return false;
        const size_t total = 18 + clen + 8;
        const uint8_t head[18] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, (uint8_t)((total - 1) & 0xFF), (uint8_t)((total - 1) >> 8)};
        memcpy(out.data(), head, 18);
        const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), buf.data(), (uInt)buf.size());
        for (int i = 0; i < 4; i++) { out[18 + clen + i] = (uint8_t)(crc >> (8 * i)); out[18 + clen + 4 + i] = (uint8_t)((uint32_t)buf.size() >> (8 * i)); }
        if (fwrite(out.data(), 1, total, f) != total) return false;
        written += total; buf.clear();
        return true;
    }
    bool write(const uint8_t* p, size_t n) {
        while (n) {
            const size_t k = std::min(n, BLOCK - buf.size());
            buf.insert(buf.end(), p, p + k); p += k; n -= k;
            if (buf.size() == BLOCK && !flush_block()) return false;
        }
        return true;
    }
    bool finish() {
        if (!buf.empty() && !flush_block()) return false;
        return flush_block();              // the empty block that marks the end of a BGZF file
    }
};

inline int reg2bin(int64_t beg, int64_t end) {       // SAMv1 5.3 (0-based, end exclusive)
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

}  // namespace

// =================================================================================================
// reader
// =================================================================================================
struct pb_bam {
    Bgzf z;
    std::vector<std::string> ref_names; std::vector<int64_t> ref_lens;
    std::vector<std::vector<uint64_t>> linear;      // BAI linear index per reference (16 kb windows)
    uint64_t first_record = 0;                       // virtual offset of the first alignment record
    std::vector<uint8_t> rec, seq, qual;
    std::vector<uint32_t> cig;
};

extern "C" const char* pb_bam_last_error(void) { return g_err_bam.c_str(); }

extern "C" int pb_bam_open(const char* bam_path, const char* bai_path, pb_bam** out) {
    if (!bam_path || !out) return fail_bam(PB_ERR_INVALID, "null argument");
    pb_bam* b = new pb_bam();
    b->z.f = fopen(bam_path, "rb");
    if (!b->z.f) { delete b; return fail_bam(PB_ERR_INVALID, std::string("cannot open ") + bam_path); }
    bool err = false;
    uint8_t w[8];
    auto bad = [&](const char* m) { b->z.drain(); fclose(b->z.f); delete b; return fail_bam(PB_ERR_INVALID, std::string(bam_path) + ": " + m); };
    if (!b->z.load(0) || !b->z.read(w, 8, &err) || memcmp(w, "BAM\1", 4) != 0) return bad("not a BAM file");
    const uint32_t l_text = rd32(w + 4);
    if (l_text > (1u << 30)) return bad("implausible header length");
    std::vector<uint8_t> text(l_text);
    if (l_text && !b->z.read(text.data(), l_text, &err)) return bad("truncated header");
    if (!b->z.read(w, 4, &err)) return bad("truncated header");
    const uint32_t n_ref = rd32(w);
    for (uint32_t i = 0; i < n_ref; i++) {
        if (!b->z.read(w, 4, &err)) return bad("truncated reference list");
        const uint32_t l_name = rd32(w);
        if (l_name > (1u << 16)) return bad("implausible reference name length");
        std::vector<uint8_t> nm(l_name + 4);
        if (!b->z.read(nm.data(), l_name + 4, &err)) return bad("truncated reference list");
        b->ref_names.emplace_back(reinterpret_cast<const char*>(nm.data()), l_name ? l_name - 1 : 0);
        b->ref_lens.push_back((int64_t)rd32(nm.data() + l_name));
    }
    // a block boundary reached exactly at the end of the header: the first record starts the next block
    if (b->z.at == b->z.block.size() && !b->z.eof) { if (!b->z.load(b->z.block_addr + b->z.block_csize)) return bad("corrupt BGZF block"); }
    b->first_record = (b->z.block_addr << 16) | (uint64_t)b->z.at;
    // ---- BAI: only the linear index is used (a region query scans forward from the window's first record) ----
    const std::string bai = bai_path ? std::string(bai_path) : std::string(bam_path) + ".bai";
    if (FILE* fi = fopen(bai.c_str(), "rb")) {
        std::vector<uint8_t> d;
        uint8_t tmp[1 << 16]; size_t k;
        while ((k = fread(tmp, 1, sizeof tmp, fi)) > 0) d.insert(d.end(), tmp, tmp + k);
        fclose(fi);
        size_t p = 0;
        auto need = [&](size_t n) { return p + n <= d.size(); };
        if (need(8) && memcmp(d.data(), "BAI\1", 4) == 0) {
            const uint32_t nr = rd32(d.data() + 4); p = 8;
            b->linear.resize(nr);
            for (uint32_t r = 0; r < nr && need(4); r++) {
                const uint32_t n_bin = rd32(d.data() + p); p += 4;
                for (uint32_t j = 0; j < n_bin && need(8); j++) { const uint32_t n_chunk = rd32(d.data() + p + 4); p += 8 + (size_t)n_chunk * 16; }
                if (!need(4)) break;
                const uint32_t n_intv = rd32(d.data() + p); p += 4;
                for (uint32_t j = 0; j < n_intv && need(8); j++, p += 8) b->linear[r].push_back(rd64(d.data() + p));
            }
        }
    }
    *out = b;
    return PB_OK;
}

extern "C" int pb_bam_close(pb_bam* b) { if (b) { b->z.drain(); if (b->z.f) fclose(b->z.f); delete b; } return PB_OK; }
extern "C" int pb_bam_n_refs(const pb_bam* b, int32_t* n) { if (!b || !n) return fail_bam(PB_ERR_INVALID, "null argument"); *n = (int32_t)b->ref_names.size(); return PB_OK; }
extern "C" int pb_bam_ref(const pb_bam* b, int32_t i, const char** name, int64_t* len) {
    if (!b || i < 0 || i >= (int32_t)b->ref_names.size()) return fail_bam(PB_ERR_INVALID, "bad reference index");
    if (name) *name = b->ref_names[(size_t)i].c_str();
    if (len) *len = b->ref_lens[(size_t)i];
    return PB_OK;
}

// BamFile.process's reader loop (BamFile.scala:117-139) for the records of reference `ref_id` that overlap
// [start, stop] (1-based, inclusive; the caller applies the +-10 kb of BamFile.scala:118-119): validateRead
// (BamFile.scala:101-105) and pb_packer_add.  n_records = records packed, n_rejected = records validateRead dropped.
extern "C" int pb_bam_query_pack(pb_bam* b, int32_t ref_id, int32_t start, int32_t stop, int non_pf, int duplicates,
                                 pb_packer* packer, int64_t* n_records, int64_t* n_rejected) {
    if (!b || !packer) return fail_bam(PB_ERR_INVALID, "null argument");
    if (ref_id < 0 || ref_id >= (int32_t)b->ref_names.size()) return fail_bam(PB_ERR_INVALID, "bad reference index");
    if (start < 1) start = 1;                                  // htsjdk: start 0 = from the beginning of the contig
    uint64_t from = b->first_record;
    if ((size_t)ref_id < b->linear.size() && !b->linear[(size_t)ref_id].empty()) {
        const auto& lin = b->linear[(size_t)ref_id];
        size_t w = (size_t)((start - 1) >> 14);
        if (w >= lin.size()) w = lin.size() - 1;
        while (w > 0 && lin[w] == 0) w--;                      // an empty window: fall back to an earlier one
        if (lin[w]) from = lin[w];
    }
    if (!b->z.seek(from)) return fail_bam(PB_ERR_INVALID, "corrupt BGZF block");
    int64_t n_ok = 0, n_rej = 0;
    bool err = false;
    uint8_t w4[4];
    while (b->z.read(w4, 4, &err)) {
        const uint32_t bs = rd32(w4);
        if (bs < 32 || bs > (1u << 28)) return fail_bam(PB_ERR_INVALID, "corrupt BAM record");
        b->rec.resize(bs);
        if (!b->z.read(b->rec.data(), bs, &err)) return fail_bam(PB_ERR_INVALID, "truncated BAM record");
        const uint8_t* r = b->rec.data();
        const int32_t refID = (int32_t)rd32(r), pos0 = (int32_t)rd32(r + 4);
        const uint32_t l_read_name = r[8], mapq = r[9], n_cigar = rd16(r + 12), flag = rd16(r + 14), l_seq = rd32(r + 16);
        const int32_t next_ref = (int32_t)rd32(r + 20), tlen = (int32_t)rd32(r + 28);
        if (refID < ref_id && refID >= 0) continue;            // (only when no index told us where the reference starts)
        if (refID != ref_id) break;                            // past the reference (or into the unplaced reads)
        const int32_t aStart = pos0 + 1;
        if (aStart > stop) break;                              // coordinate-sorted: nothing further can overlap
        if (32 + (size_t)l_read_name + 4 * (size_t)n_cigar + ((size_t)l_seq + 1) / 2 + (size_t)l_seq > bs) return fail_bam(PB_ERR_INVALID, "corrupt BAM record");
        const uint8_t* cig = r + 32 + l_read_name;
        int64_t reflen = 0;
        for (uint32_t k = 0; k < n_cigar; k++) { const uint32_t e = rd32(cig + 4 * k); const uint32_t op = e & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) reflen += e >> 4; }
        const bool unmapped = flag & 0x4;
        const int64_t aEnd = (unmapped || reflen == 0) ? aStart : (int64_t)aStart + reflen - 1;
        if (aEnd < start) continue;
        if ((!non_pf && (flag & 0x200)) || (!duplicates && (flag & 0x400)) || (flag & 0x100)) { n_rej++; continue; }   // validateRead
        const uint8_t* sq = cig + 4 * (size_t)n_cigar;
        const uint8_t* ql = sq + ((size_t)l_seq + 1) / 2;
        b->cig.resize(n_cigar);
        if (n_cigar) memcpy(b->cig.data(), cig, 4 * (size_t)n_cigar);       // (BAM is little endian, and so are we; the record is unaligned)
        const uint32_t f = ((flag & 0x1) ? PB_F_PAIRED : 0) | ((flag & 0x2) ? PB_F_PROPER : 0) | ((refID == next_ref) ? PB_F_MATE_SAME_REF : 0) |
                           (unmapped ? PB_F_UNMAPPED : 0) | ((flag & 0x10) ? PB_F_REVERSE : 0);
        const int rc = pb_packer_add_bam(packer, aStart, tlen, (int32_t)mapq, f, b->cig.data(), (int32_t)n_cigar, sq, l_seq ? ql : nullptr, (int32_t)l_seq);
        if (rc != PB_OK) return fail_bam(rc, pb_last_error());
        n_ok++;
    }
    if (err) return fail_bam(PB_ERR_INVALID, "corrupt or truncated BGZF stream");
    if (n_records) *n_records = n_ok;
    if (n_rejected) *n_rejected = n_rej;
    if (getenv("PB_BAM_TRACE")) { fprintf(stderr, "pb_bam_query_pack: %lld records, %.3f s in BGZF block loads so far\n", (long long)n_ok, b->z.t_load); }
    return PB_OK;
}

// =================================================================================================
// writer (synthetic inputs for the reference JVM and for round-trip tests)
// =================================================================================================
struct pb_bam_writer {
    BgzfWriter z;
    std::vector<std::string> ref_names; std::vector<int64_t> ref_lens;
    struct RefIndex { std::map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins; std::vector<uint64_t> linear; };
    std::vector<RefIndex> index;
    int32_t last_ref = -1; int32_t last_pos = 0; uint64_t n_records = 0;
};

extern "C" int pb_bam_writer_open(const char* path, const char* const* ref_names, const int64_t* ref_lens, int32_t n_refs,
                                  const char* program_line, pb_bam_writer** out) {
    if (!path || !out || (n_refs && (!ref_names || !ref_lens))) return fail_bam(PB_ERR_INVALID, "null argument");
    pb_bam_writer* w = new pb_bam_writer();
    w->z.f = fopen(path, "wb");
    if (!w->z.f) { delete w; return fail_bam(PB_ERR_INVALID, std::string("cannot create ") + path); }
    std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
    for (int32_t i = 0; i < n_refs; i++) {
        w->ref_names.emplace_back(ref_names[i]); w->ref_lens.push_back(ref_lens[i]);
        text += std::string("@SQ\tSN:") + ref_names[i] + "\tLN:" + std::to_string(ref_lens[i]) + "\n";
    }
    if (program_line) text += std::string(program_line) + "\n";
    std::vector<uint8_t> h;
    h.insert(h.end(), {'B', 'A', 'M', 1});
    wr32(h, (uint32_t)text.size()); h.insert(h.end(), text.begin(), text.end());
    wr32(h, (uint32_t)n_refs);
    for (int32_t i = 0; i < n_refs; i++) {
        wr32(h, (uint32_t)w->ref_names[(size_t)i].size() + 1);
        h.insert(h.end(), w->ref_names[(size_t)i].begin(), w->ref_names[(size_t)i].end()); h.push_back(0);
        wr32(h, (uint32_t)ref_lens[i]);
    }
    w->index.resize((size_t)n_refs);
    if (!w->z.write(h.data(), h.size()) || !w->z.flush_block()) { fclose(w->z.f); delete w; return fail_bam(PB_ERR_INVALID, "write failed"); }
    *out = w;
    return PB_OK;
}

// Appends the reads of a host batch (coordinate order within and across calls) as records of reference `ref_id`.
// Fields a pb_batch does not carry are synthesised: QNAME "r<ordinal>", RNEXT = RNAME (or unset when the mate flag says
// another reference), PNEXT = POS + TLEN for a positive TLEN, no auxiliary tags.  extra_flags[r] (may be NULL) is OR-ed
// into the SAM flag word (e.g. 0x100 secondary, 0x200 QC fail, 0x400 duplicate, to exercise validateRead).
extern "C" int pb_bam_writer_add_batch(pb_bam_writer* w, int32_t ref_id, const pb_batch* b, const uint16_t* extra_flags) {
    if (!w || !b) return fail_bam(PB_ERR_INVALID, "null argument");
    if (b->mem != PB_MEM_HOST || !b->quals || !b->bases2) return fail_bam(PB_ERR_INVALID, "pb_bam_writer_add_batch needs a host batch with quals and bases2");
    if (ref_id < 0 || ref_id >= (int32_t)w->ref_names.size() || ref_id < w->last_ref) return fail_bam(PB_ERR_INVALID, "bad reference index / order");
    static const uint8_t ENC[4] = {1, 2, 4, 8};                                    // A C G T in the 4-bit alphabet
    if (ref_id != w->last_ref) { w->last_ref = ref_id; w->last_pos = 0; }
    pb_bam_writer::RefIndex& ix = w->index[(size_t)ref_id];
    std::vector<uint8_t> rec;
    int64_t ei = 0;
    for (int64_t r = 0; r < b->n_reads; r++) {
        const int32_t pos0 = b->pos[r] - 1, L = b->read_len[r];
        if (b->pos[r] < w->last_pos) return fail_bam(PB_ERR_UNSORTED, "records must be added in coordinate order");
        w->last_pos = b->pos[r];
        const uint32_t c0 = b->cigar_off[r], nc = b->cigar_off[r + 1] - c0;
        const uint8_t fl = b->flags[r];
        int64_t reflen = 0;
        for (uint32_t k = 0; k < nc; k++) { const uint32_t e = b->cigar[c0 + k]; const uint32_t op = e & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) reflen += e >> 4; }
        const int64_t end0 = pos0 + ((fl & PB_F_UNMAPPED) || reflen == 0 ? 1 : reflen);    // exclusive
        uint32_t flag = ((fl & PB_F_PAIRED) ? 0x1 : 0) | ((fl & PB_F_PROPER) ? 0x2 : 0) | ((fl & PB_F_UNMAPPED) ? 0x4 : 0) | ((fl & PB_F_REVERSE) ? 0x10 : 0);
        if (fl & PB_F_PAIRED) flag |= (b->tlen[r] > 0 || (b->tlen[r] == 0 && !(fl & PB_F_REVERSE))) ? 0x40 : 0x80;
        if (extra_flags) flag |= extra_flags[r];
        const std::string qname = "r" + std::to_string(w->n_records);
        rec.clear();
        wr32(rec, 0);                                                               // block_size, patched below
        wr32(rec, (uint32_t)ref_id); wr32(rec, (uint32_t)pos0);
        rec.push_back((uint8_t)(qname.size() + 1)); rec.push_back(b->mapq[r]);
        const int bin = reg2bin(pos0, end0);
        rec.push_back((uint8_t)(bin & 0xFF)); rec.push_back((uint8_t)(bin >> 8));
        rec.push_back((uint8_t)(nc & 0xFF)); rec.push_back((uint8_t)(nc >> 8));
        rec.push_back((uint8_t)(flag & 0xFF)); rec.push_back((uint8_t)(flag >> 8));
        wr32(rec, (uint32_t)L);
        const bool paired = fl & PB_F_PAIRED;
        const int32_t n_refs = (int32_t)w->ref_names.size();
        const int32_t other_ref = n_refs >= 2 ? (ref_id + 1) % n_refs : -1;          // "mate on another reference"
        wr32(rec, (uint32_t)((fl & PB_F_MATE_SAME_REF) ? ref_id : other_ref));
        wr32(rec, paired ? (uint32_t)std::max<int64_t>(0, (int64_t)pos0 + (b->tlen[r] > 0 ? b->tlen[r] - L : b->tlen[r] < 0 ? b->tlen[r] + L : 0)) : (uint32_t)-1);
        wr32(rec, (uint32_t)b->tlen[r]);
        rec.insert(rec.end(), qname.begin(), qname.end()); rec.push_back(0);
        for (uint32_t k = 0; k < nc; k++) wr32(rec, b->cigar[c0 + k]);
        // bases and qualities: 2-bit codes + the exception table give back the original letters and bytes
        const uint32_t s0 = b->seq_off[r];
        const size_t seq_at = rec.size();
        rec.resize(seq_at + ((size_t)L + 1) / 2 + (size_t)L, 0);
        const bool hasq = fl & PB_F_HAS_QUALS;
        for (int32_t j = 0; j < L; j++) {
            const uint32_t i = s0 + (uint32_t)j;
            uint8_t q = b->quals[i], code4;
            if (q & 0x80) {
                while (ei < b->n_exc && b->exc_idx[ei] < i) ei++;
                if (ei >= b->n_exc || b->exc_idx[ei] != i) return fail_bam(PB_ERR_INVALID, "exception table does not cover a marked base");
                const char* p = strchr("=ACMGRSVTWYHKDBN", (int)b->exc_base[ei]);
                code4 = p ? (uint8_t)(p - "=ACMGRSVTWYHKDBN") : 15;
                q = b->exc_qual[ei];
            } else code4 = ENC[(b->bases2[i >> 2] >> (2 * (i & 3))) & 3];
            rec[seq_at + (size_t)(j >> 1)] |= (uint8_t)(code4 << ((~j & 1) << 2));
            rec[seq_at + ((size_t)L + 1) / 2 + (size_t)j] = hasq ? q : 0xFF;
        }
        const uint32_t bs = (uint32_t)rec.size() - 4;
        for (int i = 0; i < 4; i++) rec[(size_t)i] = (uint8_t)(bs >> (8 * i));
        const uint64_t v0 = w->z.tell();
        if (!w->z.write(rec.data(), rec.size())) return fail_bam(PB_ERR_INVALID, "write failed");
        const uint64_t v1 = w->z.tell();
        // index: the record's bin gets the chunk [v0, v1); every 16 kb window it overlaps remembers the smallest v0
        auto& chunks = ix.bins[(uint32_t)bin];
        if (!chunks.empty() && chunks.back().second == v0) chunks.back().second = v1; else chunks.emplace_back(v0, v1);
        const size_t w0 = (size_t)(pos0 >> 14), w1 = (size_t)((end0 - 1) >> 14);
        if (ix.linear.size() <= w1) ix.linear.resize(w1 + 1, 0);
        for (size_t k = w0; k <= w1; k++) if (ix.linear[k] == 0 || v0 < ix.linear[k]) ix.linear[k] = v0;
        w->n_records++;
    }
    return PB_OK;
}

extern "C" int pb_bam_writer_close(pb_bam_writer* w, const char* bai_path) {
    if (!w) return PB_OK;
    const bool ok = w->z.finish();
    fclose(w->z.f);
    int rc = ok ? PB_OK : fail_bam(PB_ERR_INVALID, "write failed");
    if (ok && bai_path) {
        std::vector<uint8_t> d = {'B', 'A', 'I', 1};
        wr32(d, (uint32_t)w->index.size());
        for (auto& ix : w->index) {
            // windows no record starts or overlaps inherit the following offset convention of samtools: keep the previous
            uint64_t prev = 0;
            for (auto& v : ix.linear) { if (v == 0) v = prev; else prev = v; }
            wr32(d, (uint32_t)ix.bins.size());
            for (auto& kv : ix.bins) {
                wr32(d, kv.first); wr32(d, (uint32_t)kv.second.size());
                for (auto& c : kv.second) { wr64(d, c.first); wr64(d, c.second); }
            }
            wr32(d, (uint32_t)ix.linear.size());
            for (uint64_t v : ix.linear) wr64(d, v);
        }
        FILE* fi = fopen(bai_path, "wb");
        if (!fi || fwrite(d.data(), 1, d.size(), fi) != d.size()) rc = fail_bam(PB_ERR_INVALID, std::string("cannot write ") + bai_path);
        if (fi) fclose(fi);
    }
    delete w;
    return rc;
}

// FASTA + .fai for the assembly (what GenomeFile loads, GenomeFile.scala:31-42): 60 bases to the line
extern "C" int pb_fasta_write(const char* path, const char* const* names, const uint8_t* const* seqs, const int64_t* lens, int32_t n) {
    if (!path || (n && (!names || !seqs || !lens))) return fail_bam(PB_ERR_INVALID, "null argument");
    FILE* f = fopen(path, "wb");
    FILE* fi = fopen((std::string(path) + ".fai").c_str(), "wb");
    if (!f || !fi) { if (f) fclose(f); if (fi) fclose(fi); return fail_bam(PB_ERR_INVALID, std::string("cannot create ") + path); }
    int64_t off = 0;
    for (int32_t i = 0; i < n; i++) {
        off += fprintf(f, ">%s\n", names[i]);
        fprintf(fi, "%s\t%lld\t%lld\t60\t61\n", names[i], (long long)lens[i], (long long)off);
        for (int64_t p = 0; p < lens[i]; p += 60) {
            const size_t k = (size_t)std::min<int64_t>(60, lens[i] - p);
            fwrite(seqs[i] + p, 1, k, f); fputc('\n', f);
            off += (int64_t)k + 1;
        }
    }
    fclose(f); fclose(fi);
    return PB_OK;
}
