// pb_synth.cpp -- deterministic synthetic genome + read generator for the BASELINE.json configs
// (SURVEY.md 8d).  Host-only bench/test tooling: writes packed pb_batch arrays directly.
//
// Everything is a pure function of (seed, coordinates): the assembly base at a locus, the planted
// variants, and the content of fragment k.  A region's reads can therefore be generated on their own
// (streaming, any genome size) and a read that falls in the halo of two adjacent chunks is identical
// in both.
//
// Model: the ASSEMBLY is what Pilon is given; reads are sampled from a TRUTH that differs from it by
// planted SNPs (1 / snp_block), indels (1 / indel_block, 1..10 bp, 30 % inside homopolymers where half
// of the reads report the indel right-aligned so the engine's left shift has work to do), with
// sequencing errors, soft clips, improper pairs, low MAPQ and N bases at the configured rates.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <omp.h>

#include "../../include/pilon_b200.h"

namespace {

inline uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
inline uint64_t h2(uint64_t a, uint64_t b) { return mix(a ^ mix(b)); }
inline uint64_t h3(uint64_t a, uint64_t b, uint64_t c) { return mix(a ^ mix(b ^ mix(c))); }

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() { return mix(s++); }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
};

}  // namespace

extern "C" {

typedef struct ps_params {
    uint64_t seed;
    int64_t contig_len;
    int64_t chunk_size;       // Pilon chunking of this contig (indel sites keep away from chunk edges)
    int32_t read_len;         // 150
    int32_t snp_block;        // 1000
    int32_t indel_block;      // 5000
    int32_t n_period;         // 20000: a 20-base N run every n_period loci (0.1 %)
    double depth;             // x coverage of this library
    double ins_mean, ins_sd;  // fragment length distribution
    double het_frac;          // fraction of planted sites that are heterozygous (per-fragment coin)
    double sub_err, indel_err, clip_frac, improper_frac, lowmq_frac, n_frac;
    uint64_t lib_seed;        // distinguishes libraries (frags / jumps) over the same genome
} ps_params;

}  // extern "C"

namespace {

constexpr int MAXK = 10;

struct IndelSite { int64_t s; int k; bool ins, homop, het, ok; uint8_t X; int run; uint8_t bytes[MAXK]; };

inline bool near_chunk_edge(const ps_params& P, int64_t x) {
    if (P.chunk_size <= 0) return false;
    const int64_t m = (x - 1) % P.chunk_size;
    return m < 256 || m > P.chunk_size - 256 || x < 300 || x > P.contig_len - 300;
}

inline IndelSite indel_site(const ps_params& P, int64_t j) {
    IndelSite st;
    const uint64_t h = h3(P.seed, 0x1D, (uint64_t)j);
    const int B = P.indel_block;
    st.s = j * B + B / 5 + (int64_t)(h % (uint64_t)(B * 3 / 5));
    st.k = 1 + (int)((h >> 20) % MAXK);
    st.ins = (h >> 30) & 1;
    st.homop = ((h >> 32) % 100) < 30;
    st.het = ((h >> 40) % 10000) < (uint64_t)(P.het_frac * 10000);
    st.X = (uint8_t)"ACGT"[(h >> 48) & 3];
    st.run = st.k + 2 + (int)((h >> 52) % 5);
    for (int t = 0; t < MAXK; t++) st.bytes[t] = st.homop ? st.X : (uint8_t)"ACGT"[(h3(P.seed, 0x1E, (uint64_t)j * 16 + t)) & 3];
    st.ok = st.s > 300 && st.s + 64 < P.contig_len && !near_chunk_edge(P, st.s) && !near_chunk_edge(P, st.s + 40);
    return st;
}

// assembly base at 1-based locus x
inline uint8_t asm_base(const ps_params& P, int64_t x) {
    if (P.n_period > 0) { const int64_t m = x % P.n_period; if (m >= P.n_period / 2 && m < P.n_period / 2 + 20) return 'N'; }
    const IndelSite st = indel_site(P, (x - 1) / P.indel_block);
    if (st.ok && st.homop && x >= st.s - 1 && x <= st.s + st.run) {
        if (x == st.s - 1 || x == st.s + st.run) return st.X == 'A' ? 'C' : 'A';     // fence the run
        return st.X;
    }
    return (uint8_t)"ACGT"[h2(P.seed, (uint64_t)x) & 3];
}

inline bool snp_at(const ps_params& P, int64_t x, uint8_t* alt, bool* het) {
    const int64_t j = (x - 1) / P.snp_block;
    const uint64_t h = h3(P.seed, 0x51, (uint64_t)j);
    const int64_t t = j * P.snp_block + P.snp_block / 10 + (int64_t)(h % (uint64_t)(P.snp_block * 8 / 10));
    if (t != x) return false;
    const IndelSite st = indel_site(P, (x - 1) / P.indel_block);
    if (st.ok && x > st.s - 50 && x < st.s + 50) return false;
    const uint8_t a = asm_base(P, x);
    if (a == 'N') return false;
    const int code = a == 'A' ? 0 : a == 'C' ? 1 : a == 'G' ? 2 : 3;
    *alt = (uint8_t)"ACGT"[(code + 1 + (int)((h >> 24) % 3)) & 3];
    *het = ((h >> 40) % 10000) < (uint64_t)(P.het_frac * 10000);
    return true;
}

inline int64_t snp_locus_of_block(const ps_params& P, int64_t j) {
    const uint64_t h = h3(P.seed, 0x51, (uint64_t)j);
    return j * P.snp_block + P.snp_block / 10 + (int64_t)(h % (uint64_t)(P.snp_block * 8 / 10));
}

// Window cache: assembly bytes and SNP markers (0 none; else alt letter | 0x80 if heterozygous)
struct Window { int64_t w0, w1; const uint8_t* A; const uint8_t* snp; };

// truth base for an aligned (M) position of fragment `frag`
inline uint8_t truth_base(const ps_params& P, const Window& W, int64_t x, uint64_t frag) {
    uint8_t a = W.A[x - W.w0];
    const uint8_t m = W.snp[x - W.w0];
    if (m) { if (!(m & 0x80) || (h3(P.seed, frag, (uint64_t)x) & 1)) return (uint8_t)(m & 0x7F); }
    if (a == 'N') a = (uint8_t)"ACGT"[h2(P.seed ^ 0x77, (uint64_t)x) & 3];
    return a;
}

struct Frag { int64_t start; int32_t ins; };

inline Frag fragment(const ps_params& P, int64_t k, double delta) {
    const uint64_t h = h3(P.seed ^ P.lib_seed, 0xF2, (uint64_t)k);
    const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
    Frag f;
    f.start = (int64_t)std::floor(((double)k + u) * delta) + 1;
    const uint64_t g = mix(h);
    const double u1 = ((double)(g >> 11) + 1.0) * (1.0 / 9007199254740993.0);
    const double u2 = (double)(mix(g) >> 11) * (1.0 / 9007199254740992.0);
    const double z = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    int64_t ins = (int64_t)std::llround(P.ins_mean + P.ins_sd * z);
    if (ins < P.read_len) ins = P.read_len;
    f.ins = (int32_t)ins;
    return f;
}

struct ReadKey { int64_t pos; int64_t k; int32_t mate; int32_t ins; };

constexpr int CIG_STRIDE = 24;

struct Gen {
    std::vector<int32_t> pos, tlen, read_len;
    std::vector<uint8_t> mapq, flags, quals, bases2, exc_base, exc_qual;
    std::vector<uint32_t> cigar_off, cigar, seq_off, exc_idx;
    std::vector<uint8_t> qual_codes;        // packed transport of quals (pb_batch.qual_codes) when <= 16 distinct bytes occur
    uint8_t lut[16] = {0}; bool q4_ok = false; int code_bits = 0;
    int64_t aligned = 0;
};

// generate one read into fixed-stride slots; returns number of cigar ops
int gen_read(const ps_params& P, const Window& W, const ReadKey& rk, uint8_t* q_out, uint8_t* b2_out, uint32_t* cig_out,
             int32_t* rlen_out, uint8_t* mapq_out, uint8_t* flags_out, int32_t* tlen_out,
             std::vector<uint32_t>& exc_idx, std::vector<uint8_t>& exc_base, std::vector<uint8_t>& exc_qual,
             uint32_t seq_base, int64_t* aligned) {
    const int L = P.read_len;
    Rng rng(h3(P.seed ^ P.lib_seed, 0xA0 + (uint64_t)rk.mate, (uint64_t)rk.k));
    const uint64_t frag = (uint64_t)rk.k ^ (P.lib_seed << 20);
    uint8_t bases[512];
    int nb = 0, ncig = 0;
    uint32_t ops[CIG_STRIDE];
    auto push = [&](int op, int len) {
        if (len <= 0) return;
        if (ncig && (int)(ops[ncig - 1] & 15) == op) ops[ncig - 1] += (uint32_t)len << 4;
        else if (ncig < CIG_STRIDE) ops[ncig++] = ((uint32_t)len << 4) | (uint32_t)op;
    };
    // soft clip decision
    int clipL = 0, clipR = 0;
    if (rng.uni() < P.clip_frac) { const int c = 5 + (int)rng.below(26); if (rng.next() & 1) clipL = c; else clipR = c; }
    const int La = L - clipL - clipR;
    for (int i = 0; i < clipL; i++) bases[nb++] = (uint8_t)"ACGT"[rng.below(4)];
    push(4, clipL);
    // private sequencing indel
    int64_t err_locus = -1; bool err_ins = false;
    if (rng.uni() < 1.0 - std::pow(1.0 - P.indel_err, (double)La)) { err_locus = rk.pos + 20 + rng.below((uint32_t)std::max(1, La - 40)); err_ins = rng.next() & 1; }
    const bool right_aligned = rng.next() & 1;     // aligner placement inside homopolymers
    int64_t cur = rk.pos; int na = 0;
    while (na < La && cur <= P.contig_len) {
        // next indel event at locus e > cur
        int64_t e = INT64_MAX; int ek = 0; bool eins = false; uint8_t ebytes[MAXK];
        for (int64_t j = (cur - 1) / P.indel_block; j <= (cur + La) / P.indel_block; j++) {
            const IndelSite st = indel_site(P, j);
            if (!st.ok) continue;
            if (st.het && !(h3(P.seed, frag, (uint64_t)st.s ^ 0xABCD) & 1)) continue;
            int64_t loc = st.s;
            if (st.homop && right_aligned) loc = st.ins ? st.s + st.run : st.s + st.run - st.k;
            if (loc > cur && loc < e) { e = loc; ek = st.k; eins = st.ins; memcpy(ebytes, st.bytes, MAXK); }
        }
        if (err_locus > cur && err_locus < e && (e == INT64_MAX || std::llabs(e - err_locus) > 12)) {
            e = err_locus; ek = 1; eins = err_ins; ebytes[0] = (uint8_t)"ACGT"[rng.below(4)];
        }
        int64_t mlen = std::min<int64_t>(La - na, P.contig_len - cur + 1);
        if (e != INT64_MAX) mlen = std::min<int64_t>(mlen, e - cur);
        for (int64_t i = 0; i < mlen; i++) bases[nb++] = truth_base(P, W, cur + i, frag);
        push(0, (int)mlen); cur += mlen; na += (int)mlen; *aligned += mlen;
        if (na >= La || cur > P.contig_len || e == INT64_MAX || cur != e) continue;
        if (eins) {
            if (na + ek < La) { for (int t = 0; t < ek; t++) bases[nb++] = ebytes[t]; push(1, ek); na += ek; }
            else { err_locus = -1; if (e == cur) { /* not applied: consume one aligned base to make progress */
                    bases[nb++] = truth_base(P, W, cur, frag); push(0, 1); cur++; na++; (*aligned)++; } }
        } else {
            if (cur + ek <= P.contig_len && na < La) { push(2, ek); cur += ek; }
            else { bases[nb++] = truth_base(P, W, cur, frag); push(0, 1); cur++; na++; (*aligned)++; }
        }
        if (e == err_locus) err_locus = -1;
    }
    // a read may not end with I or D
    while (ncig && ((ops[ncig - 1] & 15) == 2)) ncig--;
    if (ncig && (ops[ncig - 1] & 15) == 1) { ops[ncig - 1] = (ops[ncig - 1] & ~15u) | 4u; }
    for (int i = 0; i < clipR; i++) bases[nb++] = (uint8_t)"ACGT"[rng.below(4)];
    push(4, clipR);
    // sequencing substitutions, N calls, qualities
    memset(q_out, 0, (size_t)((L + 3) & ~3));
    memset(b2_out, 0, (size_t)(((L + 3) & ~3) / 4));
    const uint32_t sub_thr = (uint32_t)(P.sub_err * 1048576.0), n_thr = (uint32_t)(P.n_frac * 1048576.0);
    for (int i = 0; i < nb; i++) {
        uint8_t b = bases[i];
        const uint64_t r = rng.next();                        // one draw per base: [0,20) sub, [20,40) N, [40,56) qual
        if ((uint32_t)(r & 0xFFFFF) < sub_thr) b = (uint8_t)"ACGT"[((b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : 3) + 1 + (int)((r >> 60) % 3)) & 3];
        const uint32_t qu = (uint32_t)((r >> 40) & 0xFFFF);
        uint8_t q = qu < 49152 ? 37 : qu < 58982 ? 32 : qu < 63570 ? 25 : qu < 64880 ? 12 : 2;   // mean ~35, tail to 2
        const bool isN = (uint32_t)((r >> 20) & 0xFFFFF) < n_thr;
        if (isN) { b = 'N'; q = 2; }
        const int code = b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : b == 'T' ? 3 : -1;
        if (code < 0) {
            q_out[i] = 0x80;
            exc_idx.push_back(seq_base + (uint32_t)i); exc_base.push_back(b); exc_qual.push_back(q);
        } else { q_out[i] = q; b2_out[i >> 2] |= (uint8_t)(code << (2 * (i & 3))); }
    }
    *rlen_out = nb;
    *mapq_out = rng.uni() < P.lowmq_frac ? (uint8_t)rng.below(31) : 60;
    const bool improper = (h3(P.seed ^ P.lib_seed, 0x1A, (uint64_t)rk.k) % 10000) < (uint64_t)(P.improper_frac * 10000);
    *flags_out = (uint8_t)(PB_F_PAIRED | (improper ? 0 : PB_F_PROPER) | PB_F_MATE_SAME_REF | PB_F_HAS_QUALS | (rk.mate ? PB_F_REVERSE : 0));
    *tlen_out = rk.mate ? -rk.ins : rk.ins;
    memcpy(cig_out, ops, sizeof(uint32_t) * (size_t)ncig);
    return ncig;
}

}  // namespace

extern "C" {

void ps_contig(const ps_params* P, int64_t lo, int64_t hi, uint8_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t x = lo; x <= hi; x++) out[x - lo] = asm_base(*P, x);
}

// reads of this library whose pos lies in [lo, hi], coordinate-sorted, packed
void* ps_generate(const ps_params* Pp, int64_t lo, int64_t hi) {
    const ps_params P = *Pp;
    Gen* g = new Gen();
    const int L = P.read_len;
    const double delta = 2.0 * L / P.depth;
    if (lo < 1) lo = 1;
    if (hi > P.contig_len) hi = P.contig_len;
    const int64_t span = (int64_t)(P.ins_mean + 8 * P.ins_sd) + 8;
    const int64_t k0 = std::max<int64_t>(0, (int64_t)std::floor((double)(lo - span - 1) / delta) - 2);
    const int64_t k1 = (int64_t)std::ceil((double)hi / delta) + 2;
    std::vector<ReadKey> keys;
    keys.reserve((size_t)((k1 - k0 + 1) * 2));
    for (int64_t k = k0; k <= k1; k++) {
        const Frag f = fragment(P, k, delta);
        if (f.start + f.ins - 1 > P.contig_len) continue;
        const int64_t p2 = f.start + f.ins - L;
        if (f.start >= lo && f.start <= hi) keys.push_back(ReadKey{f.start, k, 0, f.ins});
        if (p2 >= lo && p2 <= hi) keys.push_back(ReadKey{p2, k, 1, f.ins});
    }
    std::sort(keys.begin(), keys.end(), [](const ReadKey& a, const ReadKey& b) {
        return a.pos != b.pos ? a.pos < b.pos : a.k != b.k ? a.k < b.k : a.mate < b.mate; });
    // window cache of the assembly and the SNP markers
    const int64_t w0 = lo, w1 = std::min<int64_t>(P.contig_len, hi + L + 2 * MAXK + 64);
    std::vector<uint8_t> WA((size_t)(w1 - w0 + 1)), WS((size_t)(w1 - w0 + 1), 0);
#pragma omp parallel for schedule(static)
    for (int64_t x = w0; x <= w1; x++) {
        WA[x - w0] = asm_base(P, x);
        if (snp_locus_of_block(P, (x - 1) / P.snp_block) == x) {
            uint8_t alt; bool het;
            if (snp_at(P, x, &alt, &het)) WS[x - w0] = (uint8_t)(alt | (het ? 0x80 : 0));
        }
    }
    const Window W{w0, w1, WA.data(), WS.data()};
    const size_t n = keys.size();
    const size_t stride = (size_t)((L + 3) & ~3);
    g->pos.resize(n); g->tlen.resize(n); g->read_len.resize(n); g->mapq.resize(n); g->flags.resize(n);
    g->seq_off.resize(n); g->cigar_off.assign(n + 1, 0);
    g->quals.resize(n * stride); g->bases2.resize(n * stride / 4);
    std::vector<uint32_t> cig_tmp(n * CIG_STRIDE);
    std::vector<uint32_t> ncig(n);
    const int T = omp_get_max_threads();
    std::vector<std::vector<uint32_t>> ti(T); std::vector<std::vector<uint8_t>> tb(T), tq(T);
    std::vector<int64_t> al(T, 0);
#pragma omp parallel
    {
        const int t = omp_get_thread_num();
#pragma omp for schedule(static)
        for (int64_t r = 0; r < (int64_t)n; r++) {
            g->pos[r] = (int32_t)keys[r].pos; g->seq_off[r] = (uint32_t)((size_t)r * stride);
            ncig[r] = (uint32_t)gen_read(P, W, keys[r], &g->quals[(size_t)r * stride], &g->bases2[(size_t)r * stride / 4],
                                        &cig_tmp[(size_t)r * CIG_STRIDE], &g->read_len[r], &g->mapq[r], &g->flags[r],
                                        &g->tlen[r], ti[t], tb[t], tq[t], (uint32_t)((size_t)r * stride), &al[t]);
        }
    }
    for (size_t r = 0; r < n; r++) g->cigar_off[r + 1] = g->cigar_off[r] + ncig[r];
    g->cigar.resize(g->cigar_off[n]);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)n; r++)
        memcpy(&g->cigar[g->cigar_off[r]], &cig_tmp[(size_t)r * CIG_STRIDE], sizeof(uint32_t) * ncig[r]);
    for (int t = 0; t < T; t++) {   // static schedule: thread t owns an ascending read range
        g->exc_idx.insert(g->exc_idx.end(), ti[t].begin(), ti[t].end());
        g->exc_base.insert(g->exc_base.end(), tb[t].begin(), tb[t].end());
        g->exc_qual.insert(g->exc_qual.end(), tq[t].begin(), tq[t].end());
        g->aligned += al[t];
    }
    {   // the same alphabet rule as pb_packer_view: at most 16 distinct stored quality bytes -> 4-bit codes
        bool seen[256] = {false};
        const int64_t ns = (int64_t)g->quals.size();
#pragma omp parallel
        {
            bool mine[256] = {false};
#pragma omp for schedule(static)
            for (int64_t i = 0; i < ns; i++) mine[g->quals[i]] = true;
#pragma omp critical
            for (int v = 0; v < 256; v++) seen[v] = seen[v] || mine[v];
        }
        uint8_t code_of[256]; int n_codes = 0; seen[0] = true;
        for (int v = 0; v < 256; v++) if (seen[v]) { if (n_codes < 16) { code_of[v] = (uint8_t)n_codes; g->lut[n_codes] = (uint8_t)v; } n_codes++; }
        g->q4_ok = n_codes <= 16;
        if (g->q4_ok) {
            const int bits = n_codes <= 8 ? 3 : 4;
            g->code_bits = bits;
            g->qual_codes.assign(((size_t)ns * bits + 7) / 8 + 32, 0);
            if (bits == 4) {
#pragma omp parallel for schedule(static)
                for (int64_t j = 0; j < ns / 2; j++)
                    g->qual_codes[j] = (uint8_t)(code_of[g->quals[2 * j]] | (code_of[g->quals[2 * j + 1]] << 4));
            } else {                                             // 8 codes = 24 bits = 3 bytes: groups never share a byte
#pragma omp parallel for schedule(static)
                for (int64_t j = 0; j < (ns + 7) / 8; j++) {
                    uint32_t v = 0;
                    for (int t = 0; t < 8; t++) if (8 * j + t < ns) v |= (uint32_t)code_of[g->quals[8 * j + t]] << (3 * t);
                    g->qual_codes[3 * j] = (uint8_t)v; g->qual_codes[3 * j + 1] = (uint8_t)(v >> 8); g->qual_codes[3 * j + 2] = (uint8_t)(v >> 16);
                }
            }
        }
    }
    return g;
}

void ps_view(void* h, pb_batch* b, int64_t* aligned) {
    Gen* g = (Gen*)h;
    memset(b, 0, sizeof(*b));
    b->n_reads = (int64_t)g->pos.size(); b->n_cigar = (int64_t)g->cigar.size();
    b->n_seq = (int64_t)g->quals.size(); b->n_exc = (int64_t)g->exc_idx.size();
    b->pos = g->pos.data(); b->tlen = g->tlen.data(); b->read_len = g->read_len.data();
    b->mapq = g->mapq.data(); b->flags = g->flags.data(); b->cigar_off = g->cigar_off.data();
    b->cigar = g->cigar.data(); b->seq_off = g->seq_off.data(); b->quals = g->quals.data();
    b->bases2 = g->bases2.data(); b->exc_idx = g->exc_idx.data(); b->exc_base = g->exc_base.data();
    b->exc_qual = g->exc_qual.data(); b->mem = PB_MEM_HOST;
    if (g->q4_ok) { b->qual_codes = g->qual_codes.data(); b->qual_code_bits = g->code_bits; memcpy(b->qual_lut, g->lut, 16); }
    if (aligned) *aligned = g->aligned;
}

void ps_free(void* h) { delete (Gen*)h; }

}  // extern "C"
