// pb_epilogue.cuh -- what every pileup kernel formulation shares: small PTX helpers and the per-locus epilogue
// (sparse merge + PileUp.BaseCall + pass-1 classification + the single write of every output plane).
//
// Reference citations are relative to /root/reference/src/main/scala/org/broadinstitute/pilon/.
#pragma once
#include "pb_kernels.cuh"

namespace pb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
                                             const uint32_t c_in[4], const uint64_t q_in[4],
                                             uint32_t mqS, uint32_t qS, uint32_t bp, uint32_t fragN,
                                             uint32_t rb, uint8_t refb, int2 rc_md = make_int2(-1, 0)) {
    // rc_md = the region's {read count, minDepth} (k_fold's device scalars) when the caller already holds them
    const bool inr = loc < R.size;
    uint32_t c[4] = {c_in[0], c_in[1], c_in[2], c_in[3]}; uint64_t q[4] = {q_in[0], q_in[1], q_in[2], q_in[3]};
    if (R.extra != nullptr && inr) {             // (kernel-uniform) the region has long-read batches: add what k_long counted
        Extra* xp = &R.extra[loc];
        const Extra x = *xp;
#pragma unroll
        for (int b = 0; b < 4; b++) { c[b] += x.cnt[b]; q[b] += x.qs[b]; }
        mqS += x.mq; qS += x.q; bp += x.bp; fragN += x.frag;
        if (x.cnt[0] | x.cnt[1] | x.cnt[2] | x.cnt[3] | x.bp) *xp = Extra{};               // self-cleaning, like the rare plane
    }
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
    // every quality sum below 2^30 (any ordinary depth: 2^30 / (127 * 256) = 33 k bases): the whole BaseCall in 32-bit arithmetic
    const bool small = ((uint64_t)qtot >> 30) == 0;
    if (R.exp_flags & 16) { call = (uint64_t)(c[0] + mqSum); }
    else if (r_ins <= 2 && r_del <= 2 && small) {
        // no indel can be called (PileUp.scala:183-191) and nothing needs 64 bits
        const bool useq = qSum > 0;                                                         // :135
        const uint32_t q0 = (uint32_t)q[0], q1 = (uint32_t)q[1], q2 = (uint32_t)q[2], q3 = (uint32_t)q[3];
        const uint32_t s0 = useq ? q0 : c[0], s1 = useq ? q1 : c[1], s2 = useq ? q2 : c[2], s3 = useq ? q3 : c[3];
        // BaseSum.order (BaseSum.scala:57-60): stable descending, ties keep A < C < G < T
        int o0 = 0; uint32_t m0 = s0;
        if (s1 > m0) { m0 = s1; o0 = 1; }
        if (s2 > m0) { m0 = s2; o0 = 2; }
        if (s3 > m0) { m0 = s3; o0 = 3; }
        int o1 = o0 == 0 ? 1 : 0; uint32_t m1 = o0 == 0 ? s1 : s0;
        if (o0 != 1 && o1 != 1 && s1 > m1) { m1 = s1; o1 = 1; }
        if (o0 != 2 && s2 > m1) { m1 = s2; o1 = 2; }
        if (o0 != 3 && s3 > m1) { m1 = s3; o1 = 3; }
        const int32_t baseSum = (int32_t)(o0 == 0 ? q0 : o0 == 1 ? q1 : o0 == 2 ? q2 : q3);  // :139
        const int32_t altSum = (int32_t)(o1 == 0 ? q0 : o1 == 1 ? q1 : o1 == 2 ? q2 : q3);   // :141
        const int32_t total = (int32_t)qtot;
        const int32_t homoScore = baseSum - (total - baseSum);                              // :144
        const int32_t half = total >> 1;                                                    // :145 (total >= 0)
        const int32_t hetero = total - abs(half - baseSum) - abs(half - altSum);            // :146
        const int homo = homoScore >= hetero;                                               // :147
        const uint32_t diff = (uint32_t)abs(homoScore - hetero);                            // |.| <= 2^31
        uint64_t score = 0;
        if (mqSum > 0) {                                                                    // :148
            const uint64_t prod = (uint64_t)diff * (uint64_t)(uint32_t)n;
            score = (prod >> 32) == 0 ? (uint64_t)((uint32_t)prod / (uint32_t)mqSum) : prod / (uint64_t)mqSum;
        }
        const int base = n > 0 ? o0 : 4;
        const int hi = n > 0 && score >= 10ull * (uint64_t)n;                               // :166-167: score / n >= 10
        call = (uint64_t)base | ((uint64_t)o1 << 3) | ((uint64_t)homo << 5) | (1ull << 8) |
               ((uint64_t)(base != 4) << 9) | ((uint64_t)hi << 10) | (score << 16);
    } else if (r_ins <= 2 && r_del <= 2) {
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
    if (small) {                                                                            // numerators below 2^31
        R.o_wq[loc] = (int8_t)(uint8_t)(mqSum > 0 ? ((uint32_t)qtot + (uint32_t)(mqSum / 2)) / (uint32_t)mqSum : 0u);      // PileUp.scala:60-62
        R.o_wmq[loc] = (int8_t)(uint8_t)(qSum > 0 ? ((uint32_t)qtot + (uint32_t)(qSum / 2)) / (uint32_t)qSum : 0u);        // PileUp.scala:56-58
    } else {
        R.o_wq[loc] = (int8_t)(uint8_t)(mqSum > 0 ? udiv_fast((uint64_t)qtot + (uint64_t)(mqSum / 2), (uint64_t)mqSum) : 0);
        R.o_wmq[loc] = (int8_t)(uint8_t)(qSum > 0 ? udiv_fast((uint64_t)qtot + (uint64_t)(qSum / 2), (uint64_t)qSum) : 0);
    }
    R.o_flags[loc] = (uint8_t)fl;
    R.o_call[loc] = call;
    if ((fl & PB_FL_CHANGED) && ((fl >> PB_FL_KIND_SHIFT) & 3) == PB_KIND_DEL) {
        const uint32_t ci = atomicAdd(&R.sc->n_cand, 1u);
        if (ci < R.cand_cap) R.cand[ci] = make_int4(loc, r_del, ilen, 0); else atomicOr(&R.sc->error, 2);
    }
}

}  // namespace pb
