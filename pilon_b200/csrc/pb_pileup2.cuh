// pb_pileup2.cuh -- the hot kernel, second generation.
//
// Same contract as k_pileup (pb_kernels.cuh): one warp owns a window of 32 loci, gathers every
// overlapping segment, then runs BaseCall + pass-1 classification and writes each per-locus output
// once.  What changed is how the per-base work (PileUp.add, PileUp.scala:75-84) is done:
//
//  * staging: the 32 candidate descriptors of a chunk are examined one per lane; each lane that
//    holds an overlapping "fast" segment copies the few 16-byte blocks of quality bytes (and the
//    12 bytes of 2-bit codes) its window needs into shared memory with cp.async, double-buffered
//    against the compute of the previous chunk -- no register staging, ~100 copies in flight per warp.
//  * compute: lane = (row group g = lane>>3, column quad k = lane&7).  Each group walks its own
//    rows (segments); a lane handles 4 consecutive loci of one row per step with byte-SIMD
//    arithmetic: the 4 quality bytes, the 4 codes, a 4-bit in-range mask.  Bases that equal the
//    lane's primary letter (the reference base of the locus) are accumulated in packed registers
//    (4 x 8-bit counts, 4 x 16-bit quality sums) for the current mapping quality of the group.
//  * anything else is exact but slower: a base that differs from the primary letter updates the
//    per-window table in shared memory directly; a change of (adjMq + 1) or 255 rows flushes the
//    packed registers into that table; segments of invalid reads, reads without qualities and
//    soft clips take the scalar lane-per-locus path of the first-generation kernel.
//
// All arithmetic is integer and order-independent, so the result is bit-identical to the scalar walk.
#pragma once
#include "pb_kernels.cuh"

namespace pb {

static constexpr int P2_WARPS = 8;
static constexpr int P2_ROWS = 32;
static constexpr int P2_ROWB = 64;       // 48 quality bytes + 12 code bytes (+4 pad) per staged row

struct __align__(16) RowHdr { uint32_t qoff_mq; uint32_t colmask; uint32_t clo, chi; };

struct __align__(16) WarpSmem {
    uint8_t pad[32];                              // loads of columns left of a row's first byte land here
    uint8_t rows[2][P2_ROWS][P2_ROWB];
    uint8_t tail[16];
    RowHdr hdr[2][P2_ROWS];
    unsigned long long tqs[32][4];                // qualSum per locus, letter
    uint32_t tcnt[32][4];                         // baseCount
    uint32_t tmq[32], tq[32], tbp[32];            // mqSum, qSum, badPair contributions
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------------
// per-locus epilogue shared by both kernel generations: merge the sparse contributions, BaseCall,
// pass-1 classification, write every output plane of locus `loc` once.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t udiv_fast(uint64_t n, uint64_t d) {      // d > 0
    if (((n | d) >> 32) == 0) return (int64_t)((uint32_t)n / (uint32_t)d);
    return (int64_t)(n / d);
}

__device__ __forceinline__ void finish_locus(const RegionDev& R, int64_t w, int lane, int32_t loc,
                                             const uint32_t c[4], const uint64_t q[4],
                                             uint32_t mqS, uint32_t qS, uint32_t bp, uint32_t fragN,
                                             uint32_t rb, uint8_t refb, int2 rc_md = make_int2(-1, 0)) {
    // rc_md = the region's {read count, minDepth} (k_fold's device scalars) when the caller already holds them
    const bool inr = loc < R.size;
    int32_t r_ins = 0, r_insq = 0, r_del = 0, r_delq = 0, r_q = 0, r_mq = 0, r_clips = 0, r_delfrag = 0;
    uint32_t gi = 0, gd = 0;
    if (inr && ((rb >> lane) & 1) && !(R.exp_flags & 32)) {
        int4* rp = reinterpret_cast<int4*>(&R.rare[loc]);
        const int4 ra = rp[0], rb2 = rp[1];
        r_ins = ra.x; r_insq = ra.y; r_del = ra.z; r_delq = ra.w; r_q = rb2.x; r_mq = rb2.y; r_clips = rb2.z; r_delfrag = rb2.w;
        if (r_ins > 2 || r_del > 2) { gi = R.r_gins[loc]; gd = R.r_gdel[loc]; }     // only an indel call needs the evidence groups
        rp[0] = make_int4(0, 0, 0, 0); rp[1] = make_int4(0, 0, 0, 0);
    }
    if (rb && lane == 0) R.rare_bits[w] = 0;
    if (!inr) return;
    const int32_t mqSum = (int32_t)(mqS + (uint32_t)r_mq), qSum = (int32_t)(qS + (uint32_t)r_q);
    const int64_t n = (int64_t)c[0] + c[1] + c[2] + c[3];
    const int64_t depth = n + r_del;
    const int64_t qtot = (int64_t)(q[0] + q[1] + q[2] + q[3]);
    uint64_t call;
    int32_t ilen = 0;
    if (R.exp_flags & 16) { call = (uint64_t)(c[0] + mqSum); }
    else if (r_ins <= 2 && r_del <= 2) {
        // no indel can be called (PileUp.scala:183-191): the plain-base BaseCall with select-based ordering
        const bool useq = qSum > 0;                                                         // :135
        const int64_t s0 = useq ? (int64_t)q[0] : c[0], s1 = useq ? (int64_t)q[1] : c[1];
        const int64_t s2 = useq ? (int64_t)q[2] : c[2], s3 = useq ? (int64_t)q[3] : c[3];
        int o0 = 0; int64_t m0 = s0;
        if (s1 > m0) { m0 = s1; o0 = 1; }
        if (s2 > m0) { m0 = s2; o0 = 2; }
        if (s3 > m0) { m0 = s3; o0 = 3; }
        int o1 = o0 == 0 ? 1 : 0; int64_t m1 = o0 == 0 ? s1 : s0;
        if (o0 != 1 && o1 != 1 && s1 > m1) { m1 = s1; o1 = 1; }
        if (o0 != 2 && s2 > m1) { m1 = s2; o1 = 2; }
        if (o0 != 3 && s3 > m1) { m1 = s3; o1 = 3; }
        const int64_t baseSum = o0 == 0 ? q[0] : o0 == 1 ? q[1] : o0 == 2 ? q[2] : q[3];     // :139
        const int64_t altSum = o1 == 0 ? q[0] : o1 == 1 ? q[1] : o1 == 2 ? q[2] : q[3];      // :141
        const int64_t homoScore = baseSum - (qtot - baseSum);                               // :144
        const int64_t half = qtot / 2;
        const int64_t hetero = qtot - abs64(half - baseSum) - abs64(half - altSum);         // :146
        const int homo = homoScore >= hetero;
        const int64_t score = mqSum > 0 ? udiv_fast((uint64_t)abs64(homoScore - hetero) * (uint64_t)n, (uint64_t)mqSum) : 0;   // :148
        const int base = n > 0 ? o0 : 4;
        const int64_t qq = n > 0 ? udiv_fast((uint64_t)score, (uint64_t)n) : 0;             // :166
        call = (uint64_t)base | ((uint64_t)o1 << 3) | ((uint64_t)homo << 5) | (1ull << 8) |
               ((uint64_t)(base != 4) << 9) | ((uint64_t)(qq >= 10) << 10) | ((uint64_t)score << 16);
    } else {
        CallIn in;
        in.c[0] = c[0]; in.c[1] = c[1]; in.c[2] = c[2]; in.c[3] = c[3];
        in.q[0] = (int64_t)q[0]; in.q[1] = (int64_t)q[1]; in.q[2] = (int64_t)q[2]; in.q[3] = (int64_t)q[3];
        in.mqSum = mqSum; in.qSum = qSum; in.ins = r_ins; in.del = r_del; in.insQual = r_insq; in.delQual = r_delq;
        in.gins = gi ? &R.groups[gi - 1] : nullptr; in.gdel = gd ? &R.groups[gd - 1] : nullptr;
        call = compute_call(R.cfg, in, &ilen);
    }
    uint32_t fl = 0;
    if (rc_md.x < 0) rc_md = make_int2(R.sc->read_count, R.sc->min_depth);
    if (rc_md.x != 0)                                                                       // GenomeRegion.scala:229-231
        fl = classify(call, depth, rc_md.y, ref_class(refb), R.cfg.fix_amb);
    if ((R.exp_flags & 8) && call != 0x1234567812345678ull) return;
    reinterpret_cast<int4*>(R.o_cnt)[loc] = make_int4((int)c[0], (int)c[1], (int)c[2], (int)c[3]);
    reinterpret_cast<longlong2*>(R.o_qs)[2 * (int64_t)loc] = make_longlong2((long long)q[0], (long long)q[1]);
    reinterpret_cast<longlong2*>(R.o_qs)[2 * (int64_t)loc + 1] = make_longlong2((long long)q[2], (long long)q[3]);
    R.o_mq[loc] = mqSum; R.o_q[loc] = qSum; R.o_bp[loc] = (int32_t)bp;
    R.o_del[loc] = r_del; R.o_delq[loc] = r_delq; R.o_ins[loc] = r_ins; R.o_insq[loc] = r_insq;
    R.o_clips[loc] = r_clips;
    R.o_cov[loc] = wrap32(depth);                                                           // GenomeRegion.scala:247
    R.o_frag[loc] = (int32_t)(fragN + (uint32_t)r_delfrag);                                 // GenomeRegion.scala:296-298
    R.o_wq[loc] = (int8_t)(uint8_t)(mqSum > 0 ? udiv_fast((uint64_t)qtot + (uint64_t)(mqSum / 2), (uint64_t)mqSum) : 0);   // PileUp.scala:60-62
    R.o_wmq[loc] = (int8_t)(uint8_t)(qSum > 0 ? udiv_fast((uint64_t)qtot + (uint64_t)(qSum / 2), (uint64_t)qSum) : 0);     // PileUp.scala:56-58
    R.o_flags[loc] = (uint8_t)fl;
    R.o_call[loc] = call;
    if ((fl & PB_FL_CHANGED) && ((fl >> PB_FL_KIND_SHIFT) & 3) == PB_KIND_DEL) {
        const uint32_t ci = atomicAdd(&R.sc->n_cand, 1u);
        if (ci < R.cand_cap) R.cand[ci] = make_int4(loc, r_del, ilen, 0); else atomicOr(&R.sc->error, 2);
    }
}

// ---------------------------------------------------------------------------------------------
// k_pileup2
// ---------------------------------------------------------------------------------------------
template <bool MINQ>
__global__ void __launch_bounds__(P2_WARPS * 32) k_pileup2(RegionDev R, const DevBatch* __restrict__ batches, int n_batches) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * P2_WARPS + (threadIdx.x >> 5);
    if (w >= R.n_win) return;                       // warps are independent: no block-level barrier below
    WarpSmem& S = reinterpret_cast<WarpSmem*>(smem_raw)[threadIdx.x >> 5];
    const int32_t w0 = (int32_t)(w << 5);
    const int g = lane >> 3, k = lane & 7, kk = k << 2;
    const int min_qual = R.cfg.min_qual;
    const uint32_t defq = (uint32_t)R.cfg.default_qual;
    const uint32_t minq_add = (uint32_t)(0x80 - (min_qual > 128 ? 128 : min_qual)) * 0x01010101u;

    // zero the per-window table (lane <-> locus)
#pragma unroll
    for (int b = 0; b < 4; b++) { S.tqs[lane][b] = 0; S.tcnt[lane][b] = 0; }
    S.tmq[lane] = 0; S.tq[lane] = 0; S.tbp[lane] = 0;
    // primary letters of my four loci = reference bases (anything not ACGT -> A; correctness never depends on it)
    uint32_t P8 = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int64_t l = (int64_t)w0 + kk + j;
        const int rc = l < R.size ? ref_class(ref_at(R, (int64_t)R.start + l)) : 0;
        P8 |= (uint32_t)(rc & 3 & (rc < 4 ? 3 : 0)) << (2 * j);
    }
    __syncwarp();

    // packed accumulators of the group's current mapping quality
    uint32_t cnt4 = 0, QLo = 0, QHi = 0, cur_mq = 0, nrows = 0;
    uint32_t fragN = 0, nprev = 0;

    auto flush = [&]() {
        if (cnt4) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t cj = (cnt4 >> (8 * j)) & 0xFF;
                if (cj) {
                    const uint32_t Qsel = (j == 0) ? (QLo & 0xFFFF) : (j == 1) ? (QHi & 0xFFFF) : (j == 2) ? (QLo >> 16) : (QHi >> 16);
                    const int l = kk + j; const uint32_t letter = (P8 >> (2 * j)) & 3;
                    atomicAdd(&S.tcnt[l][letter], cj);
                    atomicAdd(&S.tqs[l][letter], (unsigned long long)Qsel * cur_mq);
                    atomicAdd(&S.tmq[l], cj * cur_mq);
                    atomicAdd(&S.tq[l], Qsel);
                }
            }
        }
        cnt4 = 0; QLo = 0; QHi = 0; nrows = 0;
    };

    for (int b = 0; b < n_batches; b++) {
        const DevBatch& B = batches[b];
        if (B.n_reads == 0) continue;
        const int32_t fwd = B.reach[0], back = B.reach[1];
        const int64_t x = (int64_t)w0 - fwd + 1;
        int64_t khi = (((int64_t)w0 + 32 + back) + 31) >> 5; if (khi > R.n_win) khi = R.n_win;
        const uint32_t slo = x <= 0 ? 0u : B.win_first[x >> 5];
        const uint32_t shi = (((int64_t)w0 + 32 + back) > ((int64_t)R.n_win << 5)) ? (uint32_t)B.n_cigar : B.win_first[khi];
        const uint8_t* __restrict__ quals = B.quals;
        const uint8_t* __restrict__ bases2 = B.bases2;
        const int nchunks = (int)((shi - slo + 31) >> 5);

        // ---- stage one chunk of 32 candidate descriptors into buffer `buf`; returns #fast rows ----
        auto stage = [&](int c, int buf) -> int {
            const uint32_t s = slo + ((uint32_t)c << 5) + lane;
            Seg mine; mine.loc0 = 0; mine.len = 0; mine.src = 0; mine.w = 0;
            if (s < shi) mine = B.seg[s];
            const bool ov = mine.len > 0 && mine.loc0 < w0 + 32 && mine.loc0 + mine.len > w0;
            const bool fast = ov && (mine.w & SEG_VALID) && (mine.w & SEG_HASQ);
            // scalar path (lane <-> locus) for the odd ones: invalid reads / soft clips / no qualities
            unsigned m = __ballot_sync(FULL, ov && !fast);
            while (m) {
                const int j = __ffs(m) - 1; m &= m - 1;
                const int32_t sl0 = __shfl_sync(FULL, mine.loc0, j);
                const int32_t sln = __shfl_sync(FULL, mine.len, j);
                const uint32_t ssrc = __shfl_sync(FULL, mine.src, j);
                const uint32_t sw = __shfl_sync(FULL, mine.w, j);
                const uint32_t off = (uint32_t)(w0 + lane - sl0);
                if (off < (uint32_t)sln) {
                    if (sw & SEG_VALID) {
                        const uint32_t idx = ssrc + off;
                        const uint32_t qb = quals[idx];
                        const uint32_t code = (bases2[idx >> 2] >> ((idx & 3) << 1)) & 3;
                        if (!(qb & 0x80)) {
                            const uint32_t q = (sw & SEG_HASQ) ? qb : defq;
                            if (!MINQ || (int)q >= min_qual) {
                                const uint32_t mq1 = sw & 0xFFFF;
                                S.tcnt[lane][code] += 1; S.tqs[lane][code] += (unsigned long long)(q * mq1);
                                S.tmq[lane] += mq1; S.tq[lane] += q;
                            }
                        }
                    } else S.tbp[lane] += 1;                       // PileUpRegion.scala:45
                }
            }
            const unsigned fm = __ballot_sync(FULL, fast);
            if (fast) {
                const int row = __popc(fm & ((1u << lane) - 1));
                const int cA = mine.loc0 > w0 ? mine.loc0 - w0 : 0;
                const int cB = mine.loc0 + mine.len - w0 < 32 ? mine.loc0 + mine.len - w0 : 32;
                const uint32_t i0 = mine.src + (uint32_t)(w0 + cA - mine.loc0);      // base index of column cA
                uint8_t* rq = S.rows[buf][row];
                const uint32_t ga = i0 & ~15u;
                const int nblk = (int)(((i0 + (uint32_t)(cB - cA) - 1) >> 4) - (i0 >> 4)) + 1;
                cp_async16(rq, quals + ga);
                if (nblk > 1) cp_async16(rq + 16, quals + ga + 16);
                if (nblk > 2) cp_async16(rq + 32, quals + ga + 32);
                const uint32_t b0 = i0 >> 2, ba = b0 & ~3u;
                cp_async4(rq + 48, bases2 + ba); cp_async4(rq + 52, bases2 + ba + 4); cp_async4(rq + 56, bases2 + ba + 8);
                RowHdr h;
                h.qoff_mq = (uint32_t)((int)(i0 - ga) - cA + 32) | ((mine.w & 0xFFFF) << 16);   // +32 bias keeps it unsigned
                h.colmask = (cB == 32 ? 0xFFFFFFFFu : ((1u << cB) - 1)) & ~((1u << cA) - 1);
                h.clo = (8 * (b0 & 3) + 2 * (i0 & 3)) | ((uint32_t)cA << 8);           // code realignment recipe, applied after landing
                h.chi = 0;
                S.hdr[buf][row] = h;
            }
            return __popc(fm);
        };

        // ---- compute one staged chunk ----
        auto compute = [&](int buf, int n) {
            if (lane < n) {                                   // realign the 2-bit codes of row `lane` to column 0
                RowHdr& h = S.hdr[buf][lane];
                const uint32_t sft = h.clo & 0xFF, cA = (h.clo >> 8) & 0xFF;
                const uint32_t* wc = reinterpret_cast<const uint32_t*>(S.rows[buf][lane] + 48);
                const uint32_t W0 = wc[0], W1 = wc[1], W2 = wc[2];
                const uint64_t raw = ((uint64_t)__funnelshift_r(W1, W2, sft) << 32) | __funnelshift_r(W0, W1, sft);
                const uint64_t al = raw << (2 * cA);
                h.clo = (uint32_t)al; h.chi = (uint32_t)(al >> 32);
            }
            __syncwarp();
            for (int r = g; r < n; r += 4) {
                const uint4 h = *reinterpret_cast<const uint4*>(&S.hdr[buf][r]);
                const uint32_t mq1 = h.x >> 16;
                if (mq1 != cur_mq || nrows == 255) { flush(); cur_mq = mq1; }
                nrows++;
                const uint8_t* rq = S.rows[buf][r];
                const int qo = (int)(h.x & 0xFFFF) - 32 + kk;
                const uint32_t* pw = reinterpret_cast<const uint32_t*>(rq + (qo & ~3));   // floor to a word, also for qo < 0
                const uint32_t Q4 = __funnelshift_r(pw[0], pw[1], (qo & 3) << 3);
                const uint32_t C8 = ((k < 4 ? h.z : h.w) >> ((k & 3) << 3)) & 0xFF;
                const uint32_t in4 = (((h.y >> kk) & 15u) * 0x00204081u) & 0x01010101u;
                const uint32_t X = C8 ^ P8;
                const uint32_t mis4 = (((X | (X >> 1)) & 0x55u) * 0x00041041u) & 0x01010101u;
                uint32_t val4 = (~Q4 >> 7) & 0x01010101u;
                if (MINQ) val4 &= (((Q4 & 0x7F7F7F7Fu) + minq_add) >> 7);
                const uint32_t act4 = val4 & in4;
                const uint32_t mat4 = act4 & ~mis4;
                uint32_t mm4 = act4 & mis4;
                if (mm4) {                                       // bases that differ from the primary letter: exact, direct
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if ((mm4 >> (8 * j)) & 1) {
                            const uint32_t q = (Q4 >> (8 * j)) & 0x7F, letter = (C8 >> (2 * j)) & 3;
                            const int l = kk + j;
                            atomicAdd(&S.tcnt[l][letter], 1u);
                            atomicAdd(&S.tqs[l][letter], (unsigned long long)(q * mq1));
                            atomicAdd(&S.tmq[l], mq1);
                            atomicAdd(&S.tq[l], q);
                        }
                    }
                }
                const uint32_t Qm = Q4 & (mat4 * 0xFFu);
                cnt4 += mat4;
                QLo += Qm & 0x00FF00FFu;
                QHi += (Qm >> 8) & 0x00FF00FFu;
            }
        };

        int n_cur = 0;
        if (nchunks > 0) n_cur = stage(0, 0);
        cp_async_commit();
        for (int c = 0; c < nchunks; c++) {
            int n_next = 0;
            if (c + 1 < nchunks) n_next = stage(c + 1, (c + 1) & 1);
            cp_async_commit();
            cp_async_wait<1>();
            __syncwarp();
            compute(c & 1, n_cur);
            __syncwarp();
            n_cur = n_next;
        }
        flush(); cur_mq = 0;
        __syncwarp();
        const uint32_t nnow = S.tcnt[lane][0] + S.tcnt[lane][1] + S.tcnt[lane][2] + S.tcnt[lane][3];
        if (B.frag) fragN += nnow - nprev;
        nprev = nnow;
        __syncwarp();
    }
    cp_async_wait<0>();
    __syncwarp();
    uint32_t c[4]; uint64_t q[4];
#pragma unroll
    for (int b = 0; b < 4; b++) { c[b] = S.tcnt[lane][b]; q[b] = S.tqs[lane][b]; }
    finish_locus(R, w, lane, w0 + lane, c, q, S.tmq[lane], S.tq[lane], S.tbp[lane], fragN, R.rare_bits[w],
                 (int64_t)w0 + lane < R.size ? ref_at(R, (int64_t)R.start + w0 + lane) : (uint8_t)'N');
}

}  // namespace pb
