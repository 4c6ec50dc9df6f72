// pb_device.cuh -- device-side data layout and arithmetic helpers of the B200 pileup engine.
//
// Reference semantics restated here (citations relative to
// /root/reference/src/main/scala/org/broadinstitute/pilon/):
//   Utils.scala:22-27 (roundDiv / pct), PileUp.scala:132-247 (BaseCall),
//   GenomeRegion.scala:255-271 (pass-1 classification).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/pilon_b200.h"

namespace pb {

// ---------------------------------------------------------------------------------------------
// HBM layout
// ---------------------------------------------------------------------------------------------
// One CIGAR op of one read becomes at most one "segment": a run of consecutive loci that receive
// consecutive read bases through PileUpRegion.add (PileUpRegion.scala:38-48).  The trusted-flank
// test (:118), the region test (:40) and the read validity (:107) are folded in by k_prep, so the
// hot kernel only tests  (unsigned)(locus - loc0) < len.
//   x = loc0   first locus index (0-based in region)
//   y = len    number of loci
//   z = src    batch base index of the read base that lands on loc0
//   w = mq1 (bits 0..15: adjMq + 1) | SEG_VALID | SEG_HASQ
struct __align__(16) Seg { int32_t loc0; int32_t len; uint32_t src; uint32_t w; };
static constexpr uint32_t SEG_VALID = 1u << 16;   // valid read: PileUp.add + region baseCount; else badPair++
static constexpr uint32_t SEG_HASQ  = 1u << 17;   // read has base qualities (else Pilon.defaultQual)
static constexpr uint32_t SEG_READD = 1u << 18;   // bases a deletion's left shift re-adds (PileUpRegion.scala:170-178): they keep their
                                                  // own quality even in a nanopore batch (:172 reads quals(rloc), not the :190 rule)

// One insertion / deletion observation (PileUp.addInsertion / addDeletion, PileUp.scala:98-114).
// `lk`  = locus index << 1 | (kind - 1); `h` identifies the string (exact 2-bit image for
// insertions of <= 28 plain bases, the length for deletions, a 63-bit hash otherwise).
struct __align__(16) EventKey { uint64_t lk; uint64_t h; };
struct __align__(16) Event { uint32_t src; uint32_t len; uint32_t rot; uint32_t batch; };

struct __align__(8) Group {          // device image of pb_indel (+ the winning event)
    int32_t loc, kind, list_len, win_count, win_len, win_has_n;
    uint32_t win_ev; uint32_t pad;
    int64_t str_off;
};

struct __align__(32) Rare { int32_t ins, insq, del, delq, q, mq, clips, delfrag; };

// Per-locus contributions of long-read batches (--nanopore / --pacbio; PileUpRegion.scala:120-134,160,180-181,190): their
// bases are added by k_long with global atomics into this plane, which the pileup epilogue merges and re-zeroes.  Only
// allocated for regions that get such a batch: the short-read hot path never sees it.
struct __align__(16) Extra { unsigned long long qs[4]; uint32_t cnt[4]; uint32_t mq, q, bp, frag; };

// What the pileup kernel needs to know about a batch, passed BY VALUE in the kernel parameters
// (constant bank): no dependent global load stands between a warp and its first descriptor.
struct PileBatch {
    const Seg* seg; const uint8_t* quals; const uint8_t* bases2; const uint32_t* win_first;
    const int32_t* reach;      // device: [0] max forward reach, [1] max backward reach of the batch's segments (k_fold)
    uint32_t n_cigar; uint32_t flags;      // flags: 1 = counts toward fragCoverage, 2 = has reads
};
static constexpr int PB_MAXB = 20;
// `ext` != nullptr (more than PB_MAXB batches): the table lives in device memory instead
// `spread` (scatter kernels): a grab takes every ng-th descriptor of the position-sorted list instead of 16 neighbours (deep pile-ups)
struct PileBatches { int32_t n; int32_t spread; const PileBatch* ext; PileBatch b[PB_MAXB]; };
__device__ __forceinline__ const PileBatch& pile_batch(const PileBatches& PB, int i) { return PB.ext ? PB.ext[i] : PB.b[i]; }

struct DevBatch {
    int64_t n_reads, n_cigar, n_seq, n_exc;
    const int32_t *pos, *tlen, *read_len;
    const uint8_t *mapq, *flags;
    const uint32_t *cigar_off, *cigar, *seq_off;
    const uint8_t *quals, *bases2;
    const uint32_t* exc_idx;
    const uint8_t *exc_base, *exc_qual;
    Seg* seg;               // [n_cigar]
    uint32_t* win_first;    // [n_win + 1]  cigar_off[#reads with (pos - start) < 32*k]: first candidate segment slot
    int32_t* insert_out;    // [n_reads]    addRead return values (PileUpRegion.scala:219)
    int32_t* reach;         // [2] device scalars: max forward reach, max backward reach (loci)
    int32_t frag;           // counts toward fragCoverage (GenomeRegion.scala:291,296)
    int32_t fwd, back;      // host copy of reach[0..1], filled in after k_prep (saves a dependent load per tile)
    int32_t long_read;      // BamFile.longReadType: 0, 1 = nanopore, 2 = pacbio (BamFile.scala:43-47)
};

// k_prep spreads its per-warp partial sums over SC_SLOTS slots (same-address L2 atomics serialise);
// k_scalars folds them into the Scalars block.
static constexpr int SC_SLOTS = 64;
static constexpr int BC_SPREAD = 64;     // partial sums per batch of the region baseCount (RegionDev.batch_bc): same-address atomics serialise (8 slots cost k_prep 2x)
struct ScalarSlot { unsigned long long aligned_bases; int read_count, unknown_ops, dropped_oob; unsigned n_work; int fwd[8], back[8]; };

struct Scalars {
    unsigned long long base_count;      // PileUpRegion.baseCount
    unsigned long long aligned_bases;
    unsigned long long str_bytes;       // bytes handed out by k_indel_strings
    long long coverage;                 // roundDiv(baseCount, size)
    int read_count;
    int phys_cov_start, insert_size_start;   // PileUpRegion.scala:59-60
    int unknown_ops, dropped_oob;
    unsigned n_events, n_groups, n_cand, n_work;
    int n_calls;                        // loci with a changing / ambiguous call (pb_region_result.calls), written by the select
    int error;
    int min_depth;
};

struct Cfg {
    int min_qual, min_mq, flank, default_qual, min_min_depth, old_indel, fix_amb;
    double min_depth;
};

struct RegionDev {
    int32_t start, stop;
    int64_t size;
    int32_t n_win;              // ceil(size / 32)
    int32_t ref_locus0;         // locus of ref[0]  (= max(start - PB_REF_HALO, 1))
    int32_t ref_end;            // last locus held   (= min(stop + PB_REF_HALO, contig length))
    const uint8_t* ref;         // raw contig bytes for loci [ref_locus0, ref_end]
    Cfg cfg;
    int32_t read_count, min_depth;   // host copies of the region scalars, valid for kernels launched after k_scalars' read-back
    Scalars* sc;
    // sparse ("rare") per-locus contributions, zero between regions; written by k_prep with atomics,
    // consumed and re-zeroed by the pileup epilogue wherever rare_bits says so.  One 32-byte struct per
    // locus = one DRAM sector (eight separate planes cost eight sector round trips per touched locus).
    Rare* rare;
    uint32_t *r_gins, *r_gdel;  // group index + 1 of the locus' insertion / deletion evidence
    uint32_t* rare_bits;        // [n_win] bit l set = locus 32*w + l has any rare contribution
    Extra* extra;               // [size] or nullptr: what long-read batches contributed (k_long)
    const uint8_t* head;        // long-read mode: raw contig bytes [0, head_len) -- the reference indexes refBases with REGION
    int64_t head_len;           //   indices in homoRun / nanoporeExclude (PileUpRegion.scala:120-134,180-181,190)
    int64_t contig_len;
    int2* pc_diff;              // physCov / insertSize difference array (PileUpRegion.scala:62-88)
    // events
    EventKey* ev_key; Event* ev; uint32_t ev_cap;
    Group* groups; uint32_t groups_cap;
    uint8_t* str_pool; uint64_t str_cap;
    int4* work;  uint32_t work_cap;     // trusted in-region I / D ops queued by k_prep for k_indel: (read, op slot, readOffset | batch << 24, locus)
    uint32_t work_slots, work_sub;      // the queue is work_slots sub-queues of work_sub entries, one counter each (ScalarSlot.n_work):
                                        // a single counter is a hot address every warp with an indel would wait on
    int4* cand;  uint32_t cand_cap;     // unordered pass-1 DEL candidates: (locus index, deletions, length, -)
    ScalarSlot* slots;                  // [SC_SLOTS] k_prep partial sums
    unsigned long long* batch_bc;       // [batches * BC_SPREAD] PileUpRegion.baseCount contributed by each batch (BamFile.scala:120,146)
    int32_t exp_flags;                  // PB_EXP timing experiments (results invalid): 1 skip epilogue, 2 skip compute, 4 skip staging copies
    // outputs (final state)
    int32_t* o_cnt;   // [size*4]
    int64_t* o_qs;    // [size*4]
    int32_t *o_mq, *o_q, *o_pc, *o_is, *o_bp, *o_del, *o_delq, *o_ins, *o_insq, *o_clips;
    int32_t *o_cov, *o_frag;
    int8_t *o_wq, *o_wmq;
    uint8_t* o_flags;
    uint64_t* o_call;
};

// ---------------------------------------------------------------------------------------------
// Reference-predicted 2-bit codes of one read (pb_batch.base_delta_idx): emit(code) is called once per stored base
// slot of the read, padding included, in order.  Shared by the host encoder and the device rebuild kernel.
// ---------------------------------------------------------------------------------------------
template <class Emit>
__host__ __device__ __forceinline__ void predicted_codes(int32_t pos, const uint32_t* cigar, uint32_t n_ops, int32_t read_len,
                                                         const uint8_t* ref, int64_t ref_lo, int64_t ref_hi, Emit&& emit) {
    int64_t locus = pos; int32_t done = 0;
    for (uint32_t k = 0; k < n_ops && done < read_len; k++) {
        const uint32_t e = cigar[k]; const int op = (int)(e & 15); const int64_t len = (int64_t)(e >> 4);
        if (op == 0 || op == 7 || op == 8) {                         // M = X: predicted from the reference
            for (int64_t j = 0; j < len && done < read_len; j++, done++) {
                const int64_t l = locus + j;
                uint32_t code = 0;
                if (l >= ref_lo && l <= ref_hi) { const uint8_t b = ref[l - ref_lo]; code = b == 'C' ? 1u : b == 'G' ? 2u : b == 'T' ? 3u : 0u; }
                emit(code);
            }
            locus += len;
        } else if (op == 1 || op == 4) {                             // I S: no reference counterpart
            for (int64_t j = 0; j < len && done < read_len; j++, done++) emit(0u);
        } else if (op == 2 || op == 3) locus += len;                 // D N
    }
    for (const int32_t padded = (read_len + 3) & ~3; done < padded; done++) emit(0u);
}

// ---------------------------------------------------------------------------------------------
// JVM arithmetic (Utils.scala:22-27)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int32_t wrap32(int64_t x) { return (int32_t)(uint32_t)(uint64_t)x; }
__host__ __device__ __forceinline__ int64_t roundDivL(int64_t n, int64_t d) {
    return d > 0 ? (int64_t)((uint64_t)n + (uint64_t)(d / 2)) / d : 0;
}
__host__ __device__ __forceinline__ int32_t roundDivI(int32_t n, int32_t d) {
    return d > 0 ? wrap32((int64_t)n + d / 2) / d : 0;
}
__host__ __device__ __forceinline__ int32_t pctI(int32_t n, int32_t d) { return roundDivI(wrap32(100LL * n), d); }
__host__ __device__ __forceinline__ int64_t abs64(int64_t x) { return x < 0 ? -x : x; }

// ---------------------------------------------------------------------------------------------
// PileUp.BaseCall (PileUp.scala:132-247) on flat counters
// ---------------------------------------------------------------------------------------------
struct CallIn {
    int64_t c[4], q[4];
    int32_t mqSum, qSum, ins, del, insQual, delQual;
    const Group* gins;   // evidence groups of this locus or nullptr
    const Group* gdel;
};

// hetIndelCall (PileUp.scala:209-247): 0 none, 1 homozygous, 2 heterozygous
__host__ __device__ __forceinline__ int het_indel_call(const Cfg& cfg, int64_t depth, const Group* g, int32_t pct) {
    if (depth < cfg.min_min_depth || pct < 5 || g == nullptr || g->list_len == 0) return 0;   // :213
    if (g->win_count < 2 || g->win_count <= g->list_len / 2) return 0;                         // :220
    if (g->win_has_n) return 0;                                                                // :222
    const int32_t wl = g->win_len;
    if (cfg.old_indel) return (pct >= 33 && pct >= 50 - wl) ? 1 : 0;                           // :223-228
    const int32_t middle = 45 - wl > 10 ? 45 - wl : 10;                                        // :232
    const int32_t low = middle / 2, high = middle + middle - low;                              // :234-236
    if (pct > high) return 1;
    if (pct >= low) return 2;
    return 0;
}

// returns the packed call record (include/pilon_b200.h); *indel_len = length of the called indel
__host__ __device__ __forceinline__ uint64_t compute_call(const Cfg& cfg, const CallIn& in, int32_t* indel_len) {
    const int64_t n = in.c[0] + in.c[1] + in.c[2] + in.c[3];                                   // :133
    const int64_t* s = in.qSum > 0 ? in.q : in.c;                                              // :135
    // BaseSum.order (BaseSum.scala:57-60) is a stable descending sort: ties keep A<C<G<T
    int o0 = 0;
#pragma unroll
    for (int i = 1; i < 4; i++) if (s[i] > s[o0]) o0 = i;
    int o1 = o0 == 0 ? 1 : 0;
#pragma unroll
    for (int i = 0; i < 4; i++) if (i != o0 && i != o1 && s[i] > s[o1]) o1 = i;
    const int base = n > 0 ? o0 : 4;                                                           // :138
    const int64_t baseSum = in.q[o0], altSum = in.q[o1];                                       // :139,141
    const int64_t total = in.q[0] + in.q[1] + in.q[2] + in.q[3];                               // :143
    const int64_t homoScore = baseSum - (total - baseSum);                                     // :144
    const int64_t half = total / 2;                                                            // :145
    const int64_t heteroScore = total - abs64(half - baseSum) - abs64(half - altSum);          // :146
    const int homo = homoScore >= heteroScore;                                                 // :147
    const int64_t score = in.mqSum > 0
        ? (int64_t)((uint64_t)abs64(homoScore - heteroScore) * (uint64_t)n) / in.mqSum : 0;    // :148
    const int64_t depth = n + in.del;                                                          // :44
    int indel = 0, homoIndel = 1, ilen = 0, res = 0;
    if (in.ins > 2 && in.ins > in.del) {                                                       // :183-186
        const int32_t p1 = pctI(in.insQual, in.mqSum), p2 = pctI(in.ins, wrap32(n));           // :122
        res = het_indel_call(cfg, depth, in.gins, p1 > p2 ? p1 : p2);
        if (res) { indel = 1; homoIndel = res == 1; ilen = in.gins->win_len; }
    }
    if (!res && in.del > 2 && in.del > in.ins) {                                               // :188-191
        const int32_t p1 = pctI(in.delQual, in.mqSum);
        const int32_t p2 = pctI(in.del, wrap32((int64_t)wrap32(n) + in.del));                  // :123
        res = het_indel_call(cfg, depth, in.gdel, p1 > p2 ? p1 : p2);
        if (res) { indel = 2; homoIndel = res == 1; ilen = in.gdel->win_len; }
    }
    const int called = (base != 4) || indel;                                                   // :165
    const int64_t q = n > 0 ? score / n : 0;                                                   // :166
    const int hi = q >= 10;                                                                    // :167
    if (indel_len) *indel_len = ilen;
    return (uint64_t)base | ((uint64_t)o1 << 3) | ((uint64_t)homo << 5) | ((uint64_t)indel << 6) |
           ((uint64_t)homoIndel << 8) | ((uint64_t)called << 9) | ((uint64_t)hi << 10) | ((uint64_t)score << 16);
}

// reference base class: 0..3 = ACGT, 4 = 'N', 5 = anything else; upper-cased as GenomeRegion.refBase does (:783-787)
__host__ __device__ __forceinline__ int ref_class(uint8_t b) {
    if (b >= 'a' && b <= 'z') b -= 32;
    return b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : b == 'T' ? 3 : b == 'N' ? 4 : 5;
}

// GenomeRegion.postProcess pass 1, one locus, ignoring `deleted` (GenomeRegion.scala:255-271)
__host__ __device__ __forceinline__ uint32_t classify(uint64_t call, int64_t depth, int32_t min_depth, int rbi, int fixamb) {
    if (!(depth >= min_depth && rbi != 4 && PB_CALL_CALLED(call))) return 0;
    const int base = PB_CALL_BASE(call), alt = PB_CALL_ALT(call), homo = PB_CALL_HOMO(call);
    const int indel = PB_CALL_INDEL(call), homoIndel = PB_CALL_HOMOINDEL(call);
    const int b_eq_r = base == rbi;
    if (homo && b_eq_r && PB_CALL_HICONF(call) && !indel) return PB_FL_CONFIRMED;
    if (indel == 1 && homoIndel) return PB_FL_CHANGED | (PB_KIND_INS << PB_FL_KIND_SHIFT);
    if (indel == 2 && homoIndel) return PB_FL_CHANGED | (PB_KIND_DEL << PB_FL_KIND_SHIFT);
    if (!b_eq_r && PB_CALL_SCORE(call) > 0) {
        if (homo) return PB_FL_CHANGED | (PB_KIND_SNP << PB_FL_KIND_SHIFT);
        if (fixamb || alt != rbi) return PB_FL_AMBIGUOUS | (PB_KIND_AMB << PB_FL_KIND_SHIFT);
    }
    return 0;
}

}  // namespace pb
