// pb_kernels.cuh -- sm_100a kernels of the pileup + BaseCall engine.
//
// Pipeline per region:
//   k_prep      thread per read: CIGAR walk -> segments, sparse updates, physCov diffs, window -> first segment table
//   k_indel     thread per trusted I / D op: left shift, sparse updates, event records
//   k_fold      region scalars: coverage / minDepth, reach of the segments
//   k_scan1/2/3 physCov prefix sums (PileUpRegion.computePhysCov)
//   (radix sort of the indel event keys) -> k_groups -> k_indel_strings
//   the pileup kernel (pb_pileup7.cuh scatter / pb_pileup5.cuh gather) + per-locus epilogue (pb_epilogue.cuh)
//   k_spill     sequential deletion-spill resolution (GenomeRegion.scala:259-264) + fix-up
//
// Reference citations are relative to /root/reference/src/main/scala/org/broadinstitute/pilon/.
#pragma once
#include "pb_device.cuh"

namespace pb {

static constexpr unsigned FULL = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t ref_at(const RegionDev& R, int64_t locus) { return R.ref[locus - R.ref_locus0]; }

__device__ __forceinline__ int64_t exc_find(const DevBatch& B, uint32_t idx) {
    int64_t lo = 0, hi = B.n_exc;
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (B.exc_idx[m] < idx) lo = m + 1; else hi = m; }
    return lo;   // contract: exc_idx[lo] == idx
}

// ASCII read byte + raw quality byte of batch base `idx` (what htsjdk's getReadBases / getBaseQualities hold)
__device__ __forceinline__ void read_base(const DevBatch& B, uint32_t idx, uint8_t* base, uint8_t* qraw) {
    const uint8_t q = B.quals[idx];
    if (q & 0x80) {
        const int64_t p = exc_find(B, idx);
        *base = B.exc_base[p]; *qraw = B.exc_qual[p];
    } else {
        const uint32_t code = (B.bases2[idx >> 2] >> (2 * (idx & 3))) & 3;
        *base = (uint8_t)("ACGT"[code]); *qraw = q;
    }
}

__device__ __forceinline__ void mark_rare(const RegionDev& R, int64_t i) {
    atomicOr(&R.rare_bits[i >> 5], 1u << (i & 31));
}

// physCov difference array (int2 {count, insert size} per locus) updated with ONE 64-bit add: the count (|sum| < 2^31,
// there are fewer reads than that) sits in the low word sign-extended, so a negative running count borrows exactly 1
// from the high word and pc_unpack gives it back; the high word wraps mod 2^32 like the reference's Int.
__device__ __forceinline__ unsigned long long pc_pack(int32_t dcount, int32_t dins) {
    return ((unsigned long long)(uint32_t)dins << 32) + (unsigned long long)(long long)dcount;
}
__device__ __forceinline__ int2 pc_unpack(int2 d) { d.y += d.x < 0 ? 1 : 0; return d; }

__device__ __forceinline__ uint64_t fnv64(uint64_t h, uint8_t b) { return (h ^ b) * 1099511628211ull; }

// ---- long-read helpers (PileUpRegion.scala:120-134).  Both index `refBases` -- the whole contig, 0-based -- with whatever
// the caller passes: a 1-based locus for insertions (:160), a REGION index for deletions and aligned bases (:180-181,190).
// They are restated with the same argument, so a region that does not start at locus 1 sees the same (odd) bases. ----
__device__ __forceinline__ uint8_t contig_at(const RegionDev& R, int64_t i0) {          // refBases(i0); 0 outside the contig
    if (i0 < 0 || i0 >= R.contig_len) return 0;
    if (i0 < R.head_len) return R.head[i0];
    const int64_t locus = i0 + 1;
    return (locus >= R.ref_locus0 && locus <= R.ref_end) ? R.ref[locus - R.ref_locus0] : 0;
}
__device__ __forceinline__ bool homo_run_ge4(const RegionDev& R, int64_t i0) {          // homoRun(i0) >= 4  (:120-126)
    if (i0 < 0 || i0 + 3 >= R.contig_len) return false;       // fewer than four bases left: the run cannot reach 4 (i0 outside: JVM AIOOBE)
    const uint8_t b = contig_at(R, i0);
    return contig_at(R, i0 + 1) == b && contig_at(R, i0 + 2) == b && contig_at(R, i0 + 3) == b;
}
__device__ __forceinline__ bool nanopore_exclude(const RegionDev& R, int64_t i0) {      // :128-134
    return i0 - 2 >= 0 && i0 + 2 < R.size &&                    // inRegion(locus(i0 - 2)) && inRegion(locus(i0 + 2))
           contig_at(R, i0 - 2) == 'C' && contig_at(R, i0 - 1) == 'C' && contig_at(R, i0 + 1) == 'G' && contig_at(R, i0 + 2) == 'G';
}

// number of 32-locus windows that start at or before `pos` (index into win_first)
__device__ __forceinline__ int64_t kmin_of(const RegionDev& R, int32_t pos) {
    const int64_t d = (int64_t)pos - R.start;
    if (d < 0) return 0;
    const int64_t k = (d >> 5) + 1;
    return k > (int64_t)R.n_win + 1 ? (int64_t)R.n_win + 1 : k;
}

// ---------------------------------------------------------------------------------------------
// k_prep: one thread per read.  PileUpRegion.addRead (PileUpRegion.scala:102-220) minus the
// per-base adds, which become segments for k_pileup.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_prep(RegionDev R, DevBatch B, uint32_t batch_id) {
    // CIGAR operator sets as bit masks over the BAM op code (MIDNSHP=X = 0..8)
    constexpr uint32_t OPS_ALN = 0x181u /* M = X */, OPS_REF = 0x18Du /* M D N = X */, OPS_READ = 0x193u /* M I S = X */;
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bc = 0, aligned = 0;
    int rc = 0, unk = 0, fwd = 0, back = 0;
    // win_first[k] = segment slot of the first read with (pos - start) >= 32 k: thread r owns the windows between its
    // predecessor's start and its own (usually none or one; long runs are filled by the whole warp further down)
    int64_t wf0 = 0, wf1 = 0; uint32_t wfv = 0;
    if (r < B.n_reads) {
        const Cfg& cfg = R.cfg;
        const int32_t length = B.read_len[r];
        const int mq = B.mapq[r];
        const uint8_t fl = B.flags[r];
        const bool paired = fl & PB_F_PAIRED;
        const bool valid = (mq >= cfg.min_mq) && (!paired || ((fl & PB_F_PROPER) && (fl & PB_F_MATE_SAME_REF)));  // :107
        const bool hasq = fl & PB_F_HAS_QUALS;
        const int32_t aStart = B.pos[r];
        const int32_t prevStart = r > 0 ? B.pos[r - 1] : aStart;
        const uint32_t c0 = B.cigar_off[r], c1 = B.cigar_off[r + 1];
        const uint32_t seq0 = B.seq_off[r];
        const int32_t tlen = B.tlen[r];                                    // every per-read load is issued before the first use
        if (prevStart > aStart) atomicOr(&R.sc->error, 1);                 // batch not sorted by pos
        wf0 = r == 0 ? 0 : kmin_of(R, prevStart); wf1 = kmin_of(R, aStart); wfv = c0;
        if (wf1 - wf0 <= 4) { for (int64_t k = wf0; k < wf1; k++) B.win_first[k] = wfv; wf1 = wf0; }
        const int32_t flank = cfg.flank;
        int64_t clipped = 0, reflen = 0;
        for (uint32_t k = c0; k < c1; k++) {
            const uint32_t e = B.cigar[k]; const int op = e & 15; const int64_t len = e >> 4;
            if (op == 4) clipped += len;                                                        // :139
            if ((OPS_REF >> op) & 1) reflen += len;
            if ((OPS_ALN >> op) & 1) aligned += len;
        }
        const int32_t aEnd = (fl & PB_F_UNMAPPED) ? 0 : wrap32((int64_t)aStart + reflen - 1);    // getAlignmentEnd
        // :141; without soft clips roundDiv(mq * length, length) == mq whenever the product cannot wrap
        const int32_t adjMq = (clipped == 0 && length > 0 && length < (1 << 22) && mq >= 0 && mq < 256)
                                  ? mq : roundDivI(wrap32((int64_t)mq * (length - clipped)), length);
        const uint32_t segw = (uint32_t)((adjMq + 1) & 0xFFFF) | (hasq ? SEG_HASQ : 0u);
        const int64_t tlo = flank, thi = (int64_t)length - flank;      // trusted read offsets [tlo, thi)  :118
        int64_t readOffset = 0, refOffset = 0;
        for (uint32_t k = c0; k < c1; k++) {
            const uint32_t e = B.cigar[k]; const int op = e & 15; const int64_t len = e >> 4;
            const int64_t locus = (int64_t)aStart + refOffset;                                   // :148
            Seg sg; sg.loc0 = 0; sg.len = 0; sg.src = 0; sg.w = 0;
            if ((OPS_ALN >> op) & 1) {                                                           // M = X  :184-193
                int64_t o0 = readOffset > tlo ? readOffset : tlo;
                int64_t o1 = readOffset + len < thi ? readOffset + len : thi;
                if (o1 > o0) {
                    int64_t l0 = locus + (o0 - readOffset), l1 = l0 + (o1 - o0) - 1;
                    if (l0 < R.start) { o0 += R.start - l0; l0 = R.start; }
                    if (l1 > R.stop) l1 = R.stop;
                    if (l1 >= l0) {
                        sg.loc0 = (int32_t)(l0 - R.start); sg.len = (int32_t)(l1 - l0 + 1);
                        sg.src = seq0 + (uint32_t)o0; sg.w = segw | (valid ? SEG_VALID : 0u);
                        if (valid) bc += (unsigned long long)sg.len;                             // :43
                    }
                }
            } else if (op == 4) {                                                                // S  :194-206
                const int64_t clipStart = readOffset == 0 ? locus - len : locus;
                const int64_t clipEnd = clipStart + len - 1;
                if (clipStart >= R.start && clipStart <= R.stop) { atomicAdd(&R.rare[clipStart - R.start].clips, 1); mark_rare(R, clipStart - R.start); }
                if (clipEnd >= R.start && clipEnd <= R.stop) { atomicAdd(&R.rare[clipEnd - R.start].clips, 1); mark_rare(R, clipEnd - R.start); }
                const int64_t l0 = clipStart > R.start ? clipStart : R.start;
                const int64_t l1 = clipEnd < R.stop ? clipEnd : R.stop;
                if (l1 >= l0) { sg.loc0 = (int32_t)(l0 - R.start); sg.len = (int32_t)(l1 - l0 + 1); sg.w = 0; }   // badPair++ each
            } else if (op == 1 || op == 2) {                                                     // I, D  :150-183
                // the rare, divergent part (left shift, exception look-ups, event records) runs in k_indel, one
                // thread per op, so that it does not hold 31 idle lanes hostage here
                const bool inr = op == 1 ? (locus >= R.start && locus <= R.stop && len > 0)
                                         : (locus >= R.start && locus <= R.stop && locus + len - 1 >= R.start && locus + len - 1 <= R.stop);
                if (valid && readOffset >= tlo && readOffset < thi && inr) {
                    if (readOffset >= (1 << 24) || batch_id >= 256) atomicOr(&R.sc->error, 16);
                    const uint32_t sq = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & (R.work_slots - 1);
                    const uint32_t wi = atomicAdd(&R.slots[sq].n_work, 1u);
                    if (wi < R.work_sub) R.work[(size_t)sq * R.work_sub + wi] = make_int4((int)r, (int)k, (int)(readOffset & 0xFFFFFF) | (int)(batch_id << 24), (int)locus);
                    else atomicOr(&R.sc->error, 2);
                }
            } else if (op == 5 || op == 3) {                                                     // H, N  :207-210
            } else unk++;                                                                        // :211-212
            if (sg.len > 0) {
                const int64_t f = (int64_t)R.start + sg.loc0 + sg.len - aStart;    // exclusive end relative to pos
                const int64_t bk = (int64_t)aStart - (R.start + sg.loc0);
                if (f > fwd) fwd = (int)(f > 0x3fffffff ? 0x3fffffff : f);
                if (bk > back) back = (int)(bk > 0x3fffffff ? 0x3fffffff : bk);
            }
            if (!(R.exp_flags & 128)) B.seg[k] = sg;
            if ((OPS_READ >> op) & 1) readOffset += len;                                         // :214
            if ((OPS_REF >> op) & 1) refOffset += len;                                           // :215
        }
        rc = 1;                                                                                  // :218
        // physCovIncr, PileUpRegion.scala:62-88
        int32_t ins = 0;
        if (valid && !(paired && tlen <= 0) && !(R.exp_flags & 64)) {
            int64_t s, e;
            if (!paired) { s = aStart < aEnd ? aStart : aEnd; e = aStart > aEnd ? aStart : aEnd; }
            else { s = aStart; e = (int64_t)aStart + tlen; }
            ins = wrap32(e - s); s = wrap32(s); e = wrap32(e);
            // one 64-bit RED per end instead of two 32-bit ones: count diff in the low word, insert-size diff in the high
            // word; pc_unpack() removes the borrow a negative count leaves in the high word, so both sums stay exact
            // (mod 2^32, as the JVM's Int fields are)
            if (s >= R.start && s <= R.stop) atomicAdd(reinterpret_cast<unsigned long long*>(&R.pc_diff[s - R.start]), pc_pack(1, ins));
            else if (s < R.start && !(e < R.start)) { atomicAdd(&R.sc->phys_cov_start, 1); atomicAdd(&R.sc->insert_size_start, ins); }
            if (e >= R.start && e <= R.stop) atomicAdd(reinterpret_cast<unsigned long long*>(&R.pc_diff[e - R.start]), pc_pack(-1, wrap32(-(int64_t)ins)));
        }
        B.insert_out[r] = ins;
    }
    {   // long window runs (gaps in the coverage, the stretch before the first read) and the tail after the last read
        unsigned longm = __ballot_sync(FULL, wf1 > wf0);
        while (longm) {
            const int j = __ffs(longm) - 1; longm &= longm - 1;
            const long long k0 = __shfl_sync(FULL, (long long)wf0, j), k1 = __shfl_sync(FULL, (long long)wf1, j);
            const uint32_t v = __shfl_sync(FULL, wfv, j);
            for (long long k = k0 + (threadIdx.x & 31); k < k1; k += 32) B.win_first[k] = v;
        }
        const unsigned lastm = __ballot_sync(FULL, r == B.n_reads - 1);
        if (lastm) {
            const long long k1 = __shfl_sync(FULL, (long long)kmin_of(R, r < B.n_reads ? B.pos[r] : 0), __ffs(lastm) - 1);
            for (long long k = k1 + (threadIdx.x & 31); k <= (long long)R.n_win; k += 32) B.win_first[k] = (uint32_t)B.n_cigar;
        }
    }
    // warp-aggregate the region scalars, then one atomic per quantity per warp into one of SC_SLOTS slots
    // (same-address L2 atomics serialise; a block-level reduction would make every warp wait for the slowest)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bc += __shfl_xor_sync(FULL, bc, o); aligned += __shfl_xor_sync(FULL, aligned, o);
        rc += __shfl_xor_sync(FULL, rc, o); unk += __shfl_xor_sync(FULL, unk, o);
        fwd = max(fwd, __shfl_xor_sync(FULL, fwd, o)); back = max(back, __shfl_xor_sync(FULL, back, o));
    }
    if ((threadIdx.x & 31) == 0) {
        ScalarSlot* sl = &R.slots[(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & (SC_SLOTS - 1)];
        if (bc) atomicAdd(&R.batch_bc[batch_id * BC_SPREAD + ((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & (BC_SPREAD - 1))], bc);
        if (aligned) atomicAdd(&sl->aligned_bases, aligned);
        if (rc) atomicAdd(&sl->read_count, rc);
        if (unk) atomicAdd(&sl->unknown_ops, unk);
        if (fwd) atomicMax(&sl->fwd[batch_id & 7], fwd);          // fire-and-forget: reading the slot back would wait on a hot line
        if (back) atomicMax(&sl->back[batch_id & 7], back);
    }
}

// ---------------------------------------------------------------------------------------------
// k_indel: one thread per trusted, in-region insertion / deletion op recorded by k_prep
// (PileUpRegion.scala:150-183): left shift, PileUp.addInsertion / addDeletion, the deletion's re-added
// bases (as the segment of the D op's own slot) and the event record for the majority vote.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_indel(RegionDev R, const DevBatch* __restrict__ batches) {
    __shared__ uint32_t pre[SC_SLOTS + 1];                    // flat item index -> (sub-queue, entry)
    static_assert(SC_SLOTS == 64, "two warps load the sub-queue counts");
    if (threadIdx.x < 64) {                                   // exclusive scan of the 64 sub-queue lengths: one load latency, not 64
        const uint32_t q = threadIdx.x, ln = threadIdx.x & 31;
        const uint32_t cnt = q < R.work_slots ? min(R.slots[q].n_work, R.work_sub) : 0u;
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, inc, o); if (ln >= (uint32_t)o) inc += v; }
        if (q == 31) pre[SC_SLOTS] = inc;                     // first half's total, parked in the last slot for a moment
        pre[q] = inc - cnt;
    }
    __syncthreads();
    const uint32_t half = pre[SC_SLOTS];
    __syncthreads();
    if (threadIdx.x >= 32 && threadIdx.x < 64) {
        pre[threadIdx.x] += half;
        if (threadIdx.x == 63) pre[SC_SLOTS] = pre[63] + (63u < R.work_slots ? min(R.slots[63].n_work, R.work_sub) : 0u);
    }
    __syncthreads();
    const uint32_t n_work = pre[SC_SLOTS];
    if (blockIdx.x == 0 && threadIdx.x == 0) R.sc->n_events = n_work;     // event slots in use (one per work item, some stay empty)
    int drop = 0;
    for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < n_work; wi += gridDim.x * blockDim.x) {
        if (wi < R.ev_cap) R.ev_key[wi].lk = ~0ull;                      // "no event" until this item records one
        uint32_t q = 0;
#pragma unroll
        for (uint32_t stp = SC_SLOTS / 2; stp > 0; stp >>= 1) if (pre[q + stp] <= wi) q += stp;
        const int4 it = R.work[(size_t)q * R.work_sub + (wi - pre[q])];
        const int64_t r = it.x; const uint32_t k = (uint32_t)it.y;
        const uint32_t batch_id = (uint32_t)it.z >> 24;
        const int64_t readOffset = it.z & 0xFFFFFF, locus = it.w;
        const DevBatch& B = batches[batch_id];
        const Cfg& cfg = R.cfg;
        const int32_t length = B.read_len[r];
        const int mq = B.mapq[r];
        const bool hasq = B.flags[r] & PB_F_HAS_QUALS;
        const int32_t aStart = B.pos[r];
        const uint32_t c0 = B.cigar_off[r], c1 = B.cigar_off[r + 1];
        const uint32_t seq0 = B.seq_off[r];
        int64_t clipped = 0;
        for (uint32_t kk = c0; kk < c1; kk++) { const uint32_t e = B.cigar[kk]; if ((e & 15) == 4) clipped += e >> 4; }
        const int32_t adjMq = roundDivI(wrap32((int64_t)mq * (length - clipped)), length);       // :141
        const int longRead = B.long_read;
        const int32_t indelMq = longRead > 0 ? (adjMq < 8 ? adjMq : 8) : adjMq;                  // :142
        const uint32_t segw = (uint32_t)((adjMq + 1) & 0xFFFF) | (hasq ? SEG_HASQ : 0u);
        const int64_t tlo = cfg.flank, thi = (int64_t)length - cfg.flank;
        const uint32_t e = B.cigar[k]; const int op = e & 15; const int64_t len = e >> 4;
        Seg sg; sg.loc0 = 0; sg.len = 0; sg.src = 0; sg.w = 0;
        if (op == 1) {                                                                // I  :150-162
                int64_t iloc = locus;
                {
                    const uint32_t src = seq0 + (uint32_t)readOffset;
                    int64_t j = len - 1; uint32_t rot = 0; bool dropped = false;
                    while (iloc > 1) {
                        uint8_t b, q; read_base(B, src + (uint32_t)j, &b, &q);
                        if (ref_at(R, iloc - 1) != b) break;                  // refBases(iloc - 2) == insertion(len - 1)
                        iloc -= 1; rot += 1; j = j == 0 ? len - 1 : j - 1;
                        if (iloc < R.start) { dropped = true; break; }        // JVM: AIOOBE at pileups(index(iloc))
                    }
                    if (dropped) drop++;
                    else if (longRead > 0 && homo_run_ge4(R, iloc)) {}         // :160 (homoRun is handed the LOCUS as an index)
                    else {
                        const int64_t i = iloc - R.start;
                        uint8_t b0, q0; read_base(B, src, &b0, &q0);
                        const int qual = hasq ? (int)(int8_t)q0 : (int)(int8_t)cfg.default_qual;
                        atomicAdd(&R.rare[i].insq, indelMq + 1);                 // PileUp.addInsertion, PileUp.scala:98-105
                        atomicAdd(&R.rare[i].q, qual);
                        atomicAdd(&R.rare[i].ins, 1);
                        mark_rare(R, i);
                        // identity of the rotated string: final[t] = orig[(t - rot) mod len]
                        const uint32_t rm = (uint32_t)(rot % (uint32_t)len);
                        uint64_t h = 0; bool exact = len <= 28;
                        uint64_t hh = 1469598103934665603ull;
                        for (int s2 = 0; s2 < 4; s2++) hh = fnv64(hh, (uint8_t)((uint64_t)len >> (8 * s2)));
                        for (int64_t t = 0; t < len; t++) {
                            int64_t sidx = t - rm; if (sidx < 0) sidx += len;
                            uint8_t b, q; read_base(B, src + (uint32_t)sidx, &b, &q);
                            const int code = b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : b == 'T' ? 3 : -1;
                            if (code < 0) exact = false;
                            if (exact) h |= (uint64_t)code << (2 * t);
                            hh = fnv64(hh, b);
                        }
                        h = exact ? (h | ((uint64_t)len << 56)) : ((hh >> 1) | (1ull << 63));
                        const uint32_t ei = wi;                              // event slot = work item: no shared counter
                        if (ei < R.ev_cap) {
                            R.ev_key[ei].lk = ((uint64_t)i << 1); R.ev_key[ei].h = h;
                            Event ev; ev.src = src; ev.len = (uint32_t)len; ev.rot = rm; ev.batch = batch_id;
                            R.ev[ei] = ev;
                        } else atomicOr(&R.sc->error, 2);
                    }
                }
            } else if (op == 2) {                                                                // D  :163-183
                int64_t dloc = locus, rloc = readOffset;
                {
                    bool dropped = false;
                    while (dloc > 1 && rloc > 0 && ref_at(R, dloc - 1) == ref_at(R, dloc + len - 1)) {   // refBases(dloc-2) == refBases(dloc+len-2)
                        dloc -= 1; rloc -= 1;
                        if (dloc < R.start) { dropped = true; break; }        // JVM: AIOOBE at :182
                    }
                    if (dropped) drop++;
                    else {
                        // the shift re-adds read bases [rloc, readOffset) at loci rloc' + (locus + len - readOffset)
                        // when trusted (:174-178); remove() is a no-op for valid reads (:50-58)
                        int64_t a = rloc > tlo ? rloc : tlo, b = readOffset < thi ? readOffset : thi;
                        if (b > a) {
                            sg.loc0 = (int32_t)(a + (locus + len - readOffset) - R.start); sg.len = (int32_t)(b - a);
                            sg.src = seq0 + (uint32_t)a; sg.w = segw | SEG_VALID | SEG_READD;
                            atomicAdd(&R.batch_bc[batch_id * BC_SPREAD + (wi & (BC_SPREAD - 1))], (unsigned long long)sg.len);   // :43 (rare)
                        }
                        const int64_t i = dloc - R.start;
                        // :180-181: long reads drop deletions in homopolymers (and, nanopore, at CC.GG motifs)
                        const bool lr_skip = longRead > 0 && (homo_run_ge4(R, i) || (longRead == 1 && nanopore_exclude(R, i)));
                        uint8_t b0, q0; read_base(B, seq0 + (uint32_t)readOffset, &b0, &q0);
                        const int qual = hasq ? (int)(int8_t)q0 : (int)(int8_t)cfg.default_qual;
                        if (!lr_skip) {
                        atomicAdd(&R.rare[i].mq, indelMq + 1);                   // PileUp.addDeletion, PileUp.scala:107-114
                        atomicAdd(&R.rare[i].delq, indelMq + 1);
                        atomicAdd(&R.rare[i].q, qual);
                        atomicAdd(&R.rare[i].del, 1);
                        if (B.frag) atomicAdd(&R.rare[i].delfrag, 1);
                        mark_rare(R, i);
                        const uint32_t ei = wi;                              // event slot = work item: no shared counter
                        if (ei < R.ev_cap) {
                            R.ev_key[ei].lk = ((uint64_t)i << 1) | 1; R.ev_key[ei].h = (uint64_t)len;
                            Event ev; ev.src = 0; ev.len = (uint32_t)len; ev.rot = 0; ev.batch = batch_id;
                            R.ev[ei] = ev;
                        } else atomicOr(&R.sc->error, 2);
                        }
                    }
                }
            }
        if (sg.len > 0) {
            if (!(R.exp_flags & 128)) B.seg[k] = sg;
            const int64_t f = (int64_t)R.start + sg.loc0 + sg.len - aStart;
            const int fi = (int)(f > 0x3fffffff ? 0x3fffffff : f);
            if (fi > B.reach[0]) atomicMax(&B.reach[0], fi);     // (never beyond the read's own aligned span in practice)
            // the shift walks read offsets backwards across earlier CIGAR elements (:167): after a long insertion or a
            // leading soft clip inside a repeat the re-added bases can start LEFT of the read's pos
            const int64_t bk = (int64_t)aStart - ((int64_t)R.start + sg.loc0);
            const int bi = (int)(bk > 0x3fffffff ? 0x3fffffff : bk);
            if (bi > B.reach[1]) atomicMax(&B.reach[1], bi);
        }
    }
    // few threads: plain atomics into the slots that the last k_fold folds
    if (drop) atomicAdd(&R.slots[threadIdx.x & (SC_SLOTS - 1)].dropped_oob, drop);
}

// ---------------------------------------------------------------------------------------------
// k_long: the aligned bases of a LONG-READ batch (PileUpRegion.scala:184-193 with longRead > 0), warp per segment, lane per
// base, global atomics into the Extra plane.  Not a hot path: no BASELINE config has long reads; it exists so that
// --nanopore / --pacbio inputs are served instead of refused.  Nanopore: the quality of a base that lands on a CC.GG motif
// (nanoporeExclude, region-index quirk included) counts as 0 (:190).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_long(RegionDev R, DevBatch B) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int min_qual = R.cfg.min_qual;
    for (int64_t s = warp0; s < B.n_cigar; s += nwarps) {
        const Seg sg = B.seg[s];
        if (sg.len <= 0) continue;
        const bool valid = sg.w & SEG_VALID, hasq = sg.w & SEG_HASQ;
        const uint32_t mq1 = sg.w & 0xFFFFu;
        for (int32_t j = lane; j < sg.len; j += 32) {
            const int64_t i = (int64_t)sg.loc0 + j;
            Extra& X = R.extra[i];
            if (!valid) { atomicAdd(&X.bp, 1u); continue; }                                   // PileUpRegion.scala:45
            const uint32_t idx = sg.src + (uint32_t)j;
            const uint8_t qb = B.quals[idx];
            const bool zeroed = B.long_read == 1 && !(sg.w & SEG_READD) && nanopore_exclude(R, i);    // :190: the quality counts as 0
            uint32_t code = (B.bases2[idx >> 2] >> (2 * (idx & 3))) & 3u;
            int q = hasq ? (int)qb : R.cfg.default_qual;
            if (qb & 0x80) {
                // marked uncountable: not A C G T (PileUp.scala:46-52), or a quality byte the JVM reads as negative (:77).  The
                // second kind counts after all when the motif rule replaces its quality by 0
                if (!zeroed) continue;
                uint8_t letter, qraw; read_base(B, idx, &letter, &qraw);
                if (letter == 'A') code = 0; else if (letter == 'C') code = 1; else if (letter == 'G') code = 2; else if (letter == 'T') code = 3; else continue;
            }
            if (zeroed) q = 0;
            if (q < min_qual) continue;                                                       // PileUp.scala:77
            atomicAdd(&X.cnt[code], 1u);
            atomicAdd(&X.qs[code], (unsigned long long)((uint32_t)q * mq1));
            atomicAdd(&X.mq, mq1); atomicAdd(&X.q, (uint32_t)q);
            if (B.frag) atomicAdd(&X.frag, 1u);
        }
    }
}

// region coverage and minDepth (PileUpRegion.scala:36; GenomeRegion.scala:221-224)
// launched once per group of <= 8 batches with 32 threads: folds the slots, then (last group only)
// coverage and minDepth
__global__ void k_fold(RegionDev R, int32_t* reach0, int nb, int last, int nb_total) {
    const int lane = threadIdx.x;
    unsigned long long bc = 0, al = 0; int rc = 0, unk = 0, drop = 0; int fw[8], bk[8];
    if (last) for (int i = lane; i < nb_total * BC_SPREAD; i += 32) bc += R.batch_bc[i];      // every batch's k_prep and k_indel have run
#pragma unroll
    for (int j = 0; j < 8; j++) { fw[j] = 0; bk[j] = 0; }
    for (int i = lane; i < SC_SLOTS; i += 32) {
        ScalarSlot& sl = R.slots[i];
        al += sl.aligned_bases; rc += sl.read_count; unk += sl.unknown_ops; drop += sl.dropped_oob;
#pragma unroll
        for (int j = 0; j < 8; j++) { fw[j] = max(fw[j], sl.fwd[j]); bk[j] = max(bk[j], sl.back[j]); }
        ScalarSlot z = {}; z.n_work = sl.n_work; sl = z;     // the I/D sub-queue counters live until k_indel has run
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bc += __shfl_xor_sync(FULL, bc, o); al += __shfl_xor_sync(FULL, al, o);
        rc += __shfl_xor_sync(FULL, rc, o); unk += __shfl_xor_sync(FULL, unk, o); drop += __shfl_xor_sync(FULL, drop, o);
#pragma unroll
        for (int j = 0; j < 8; j++) { fw[j] = max(fw[j], __shfl_xor_sync(FULL, fw[j], o)); bk[j] = max(bk[j], __shfl_xor_sync(FULL, bk[j], o)); }
    }
    if (lane == 0) {
        Scalars* sc = R.sc;
        sc->base_count += bc; sc->aligned_bases += al; sc->read_count += rc; sc->unknown_ops += unk; sc->dropped_oob += drop;
#pragma unroll
        for (int j = 0; j < 8; j++) if (j < nb) { reach0[2 * j] = max(reach0[2 * j], fw[j]); reach0[2 * j + 1] = max(reach0[2 * j + 1], bk[j]); }
        if (last) {
            // region coverage and minDepth (PileUpRegion.scala:36; GenomeRegion.scala:221-224)
            const long long cov = roundDivL((long long)sc->base_count, R.size);
            sc->coverage = cov;
            int md;
            if (R.cfg.min_depth >= 1) md = (int)R.cfg.min_depth;
            else {
                const double v = floor(__dadd_rn(__dmul_rn(R.cfg.min_depth, (double)cov), 0.5));   // Double.round, no FMA contraction
                md = (int)v > R.cfg.min_min_depth ? (int)v : R.cfg.min_min_depth;
            }
            sc->min_depth = md;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// indel evidence: events -> one Group per (locus, kind).  The events are radix-sorted by (locus, kind) only;
// inside a group the strict-majority string (PileUp.scala:219-220, the only one hetIndelCall can accept) is
// found with a Boyer-Moore vote over the 64-bit string identities -- no ordering of the strings is needed.
// ---------------------------------------------------------------------------------------------
// `cap` slots are sorted whatever the number of events turns out to be (it is only known on the device): the
// unused ones get the largest key and sort to the end, so no host round trip sizes the sort
__global__ void __launch_bounds__(256) k_event_keys(RegionDev R, uint32_t* keys, uint32_t* idx, uint32_t cap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    const uint32_t n = min(R.sc->n_events, R.ev_cap);
    const uint64_t lk = i < n ? R.ev_key[i].lk : ~0ull;              // ~0: a work item that recorded no event
    keys[i] = lk == ~0ull ? 0xFFFFFFFFu : (uint32_t)lk;               // lk = locus index << 1 | kind  < 2^32 - 1
    idx[i] = i;
}

// byte t of the string of event e (insertion: rotated read bases; deletion: raw reference bytes)
__device__ __forceinline__ uint8_t event_byte(const RegionDev& R, const DevBatch* batches, uint64_t lk, const Event& e, uint32_t t) {
    if (lk & 1) return ref_at(R, (int64_t)R.start + (int64_t)(lk >> 1) + t);     // refBases.slice(dloc-1, dloc+len-1)
    int64_t s = (int64_t)t - e.rot; if (s < 0) s += e.len;
    uint8_t b, q; read_base(batches[e.batch], e.src + (uint32_t)s, &b, &q);
    return b;
}

// one thread: record the group [g0, ge) of key lk whose most frequent tested string identity is h (cnt events, lowest
// event index rep)
__device__ __forceinline__ void emit_group(const RegionDev& R, const DevBatch* batches, const uint32_t* perm, uint64_t lk,
                                           uint32_t g0, uint32_t ge, uint64_t h, uint32_t cnt, uint32_t rep) {
    const uint32_t len = ge - g0;
    const Event wev = R.ev[rep];
    // hashed identities (long / non-ACGT insertions): make sure the candidate's events really are one string
    if ((h >> 63) && !(lk & 1)) {
        uint32_t same = 0;
        for (uint32_t j = g0; j < ge; j++) {
            if (R.ev_key[perm[j]].h != h) continue;
            const Event e2 = R.ev[perm[j]];
            bool eq = e2.len == wev.len;
            for (uint32_t t = 0; eq && t < wev.len; t++) eq = event_byte(R, batches, lk, e2, t) == event_byte(R, batches, lk, wev, t);
            same += eq;
        }
        if (same != cnt) atomicOr(&R.sc->error, 4);                  // 63-bit hash collision (never observed)
    }
    Group g;
    g.loc = (int32_t)(lk >> 1); g.kind = (int32_t)(lk & 1) + 1; g.list_len = (int32_t)len;
    const bool majority = cnt >= 2 && cnt > len / 2;
    g.win_count = majority ? (int32_t)cnt : 0;                       // only a strict majority is ever consumed
    g.win_len = majority ? (int32_t)wev.len : 0;
    g.win_ev = rep; g.pad = 0; g.str_off = 0;
    int has_n = 0;
    if (majority) for (uint32_t t = 0; t < wev.len; t++) has_n |= event_byte(R, batches, lk, wev, t) == 'N';   // PileUp.scala:222
    g.win_has_n = has_n;
    const uint32_t gi = atomicAdd(&R.sc->n_groups, 1u);
    if (gi < R.groups_cap) {
        R.groups[gi] = g;
        ((lk & 1) ? R.r_gdel : R.r_gins)[g.loc] = gi + 1;            // locus already has its rare bit (k_prep)
    } else atomicOr(&R.sc->error, 2);
}

// `keys` = (locus << 1 | kind) sorted, `perm[i]` = original event index of sorted position i.
// Thread per event finds the group starts.  A group of at most 64 events (any ordinary depth) is handled by its own
// thread: Boyer-Moore vote, recount, record.  Longer groups (a 5000x pile-up has thousands of events per indel) are
// worked on by the whole warp: each lane votes over its strided share -- a strict majority of the group is a strict
// majority of at least one share -- and the distinct surviving candidates are counted exactly.
// Evidence lists up to this long are voted by the thread that finds their start (two dependent loads per entry, one after
// the other); longer ones -- a real indel at depth 70 has 30-70 entries -- by the whole warp, a few load latencies in all.
static constexpr uint32_t GROUP_SERIAL_MAX = 8;
__global__ void __launch_bounds__(256) k_groups(RegionDev R, const DevBatch* batches, const uint32_t* keys,
                                                const uint32_t* perm, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    n = min(n, min(R.sc->n_events, R.ev_cap));                           // `n` slots were sorted, the events come first
    const bool is_start = i < n && keys[i] != 0xFFFFFFFFu && (i == 0 || keys[i - 1] != keys[i]);   // padding keys sort last
    bool big = false;
    if (is_start) {
        const uint32_t key = keys[i];
        uint64_t h = 0; uint32_t votes = 0, ge = i;
        for (; ge < n && ge < i + GROUP_SERIAL_MAX + 1 && keys[ge] == key; ge++) {
            const uint64_t hj = R.ev_key[perm[ge]].h;
            if (votes == 0) { h = hj; votes = 1; } else if (hj == h) votes++; else votes--;
        }
        big = ge == i + GROUP_SERIAL_MAX + 1;                            // longer: the warp takes it below
        if (!big) {
            uint32_t cnt = 0, rep = 0xFFFFFFFFu;
            for (uint32_t j = i; j < ge; j++) {
                const uint32_t ej = perm[j];
                if (R.ev_key[ej].h == h) { cnt++; rep = min(rep, ej); }
            }
            emit_group(R, batches, perm, (uint64_t)key, i, ge, h, cnt, rep);
        }
    }
    unsigned starts = __ballot_sync(FULL, big);
    while (starts) {
        const int js = __ffs(starts) - 1; starts &= starts - 1;
        const uint32_t g0 = __shfl_sync(FULL, i, js);
        const uint32_t key = keys[g0];
        // group end = first index > g0 whose key differs (keys are sorted, equal keys contiguous): the 32 lanes probe
        // g0 + 2^lane at once, then split the bracket 32 ways per step -- three or four load latencies for any group size
        uint32_t lo, hi;
        {
            const uint64_t p = (uint64_t)g0 + (1ull << lane);
            const bool mism = p >= n || keys[p] != key;
            const int kq = __ffs(__ballot_sync(FULL, mism)) - 1;         // lane 31 always reports a mismatch (n < 2^31)
            const uint64_t h0 = (uint64_t)g0 + (1ull << kq);
            hi = h0 < (uint64_t)n ? (uint32_t)h0 : n;
            lo = kq == 0 ? g0 + 1 : g0 + (1u << (kq - 1)) + 1;
        }
        while (lo < hi) {                                                // every index < lo matches, index hi does not (or is n)
            const uint32_t step = (hi - lo + 31) / 32;
            const uint64_t p = (uint64_t)lo + (uint64_t)lane * step;
            const bool mism = p >= hi || keys[p] != key;
            const int j = __ffs(__ballot_sync(FULL, mism)) - 1;          // >= 0: the last lanes reach hi
            const uint64_t h1 = (uint64_t)lo + (uint64_t)j * step;
            const uint32_t nhi = h1 < (uint64_t)hi ? (uint32_t)h1 : hi;
            lo = j == 0 ? lo : lo + (uint32_t)(j - 1) * step + 1;
            hi = j == 0 ? lo : nhi;
        }
        const uint32_t ge = lo, len = ge - g0;
        uint64_t cand = 0; uint32_t votes = 0;
        for (uint32_t t = g0 + lane; t < ge; t += 32) {
            const uint64_t hj = R.ev_key[perm[t]].h;
            if (votes == 0) { cand = hj; votes = 1; } else if (hj == cand) votes++; else votes--;
        }
        uint64_t h = 0; uint32_t cnt = 0, rep = perm[g0];
        unsigned todo = __ballot_sync(FULL, votes > 0);
        while (todo) {
            const int kc = __ffs(todo) - 1;
            const uint64_t c = ((uint64_t)__shfl_sync(FULL, (uint32_t)(cand >> 32), kc) << 32) | __shfl_sync(FULL, (uint32_t)cand, kc);
            todo &= ~__ballot_sync(FULL, votes > 0 && cand == c);        // every lane that voted for c is settled
            uint32_t my = 0, myrep = 0xFFFFFFFFu;
            for (uint32_t t = g0 + lane; t < ge; t += 32) {
                const uint32_t ej = perm[t];
                if (R.ev_key[ej].h == c) { my++; myrep = min(myrep, ej); }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { my += __shfl_xor_sync(FULL, my, o); myrep = min(myrep, __shfl_xor_sync(FULL, myrep, o)); }
            if (my > cnt) { h = c; cnt = my; rep = myrep; }
            if (my > len / 2) break;                                     // the strict majority, unique
        }
        if (lane == 0) emit_group(R, batches, perm, (uint64_t)key, g0, ge, h, cnt, rep);
    }
}

__global__ void __launch_bounds__(128) k_indel_strings(RegionDev R, const DevBatch* batches, uint32_t n_groups) {
    const uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= n_groups) return;
    Group g = R.groups[gi];
    const uint64_t off = atomicAdd(&R.sc->str_bytes, (unsigned long long)g.win_len);
    R.groups[gi].str_off = (int64_t)off;
    if (off + g.win_len > R.str_cap) { atomicOr(&R.sc->error, 2); return; }
    const uint64_t lk = ((uint64_t)g.loc << 1) | (uint64_t)(g.kind - 1);
    const Event e = R.ev[g.win_ev];
    for (int32_t t = 0; t < g.win_len; t++) R.str_pool[off + t] = event_byte(R, batches, lk, e, (uint32_t)t);
}

// sparse cleanup of the group-index planes once every consumer is done
__global__ void __launch_bounds__(128) k_groups_clear(RegionDev R, uint32_t n_groups) {
    const uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= n_groups) return;
    const Group g = R.groups[gi];
    (g.kind == 2 ? R.r_gdel : R.r_gins)[g.loc] = 0;
}

__global__ void __launch_bounds__(128) k_groups_clear_dev(RegionDev R) {      // same, count read on the device
    const uint32_t n = min(R.sc->n_groups, R.groups_cap);
    for (uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x; gi < n; gi += gridDim.x * blockDim.x) {
        const Group g = R.groups[gi];
        (g.kind == 2 ? R.r_gdel : R.r_gins)[g.loc] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// physCov: PileUpRegion.computePhysCov (PileUpRegion.scala:90-100) as a 3-kernel prefix sum over
// int2 (physCov, insertSize) with 32-bit wrap-around; consumes and re-zeroes the diff plane.
// ---------------------------------------------------------------------------------------------
static constexpr int SCAN_THREADS = 256;
static constexpr int SCAN_ITEMS = 8;
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint2 warp_incl_scan(uint2 v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned x = __shfl_up_sync(FULL, v.x, o), y = __shfl_up_sync(FULL, v.y, o);
        if (lane >= o) { v.x += x; v.y += y; }
    }
    return v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan1(RegionDev R, uint2* block_sums) {
    __shared__ uint2 wsum[SCAN_THREADS / 32];
    // thread t owns items [t*ITEMS, t*ITEMS + ITEMS) of the tile: four 16-byte loads, whole sectors used
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint2 acc = make_uint2(0, 0);
    if (base + SCAN_ITEMS <= R.size) {
        const int4* p = reinterpret_cast<const int4*>(R.pc_diff + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 2; k++) {
            const int4 v = p[k];
            const int2 a = pc_unpack(make_int2(v.x, v.y)), b = pc_unpack(make_int2(v.z, v.w));
            acc.x += (unsigned)a.x + (unsigned)b.x; acc.y += (unsigned)a.y + (unsigned)b.y;
        }
    } else {
        for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < R.size) { const int2 d = pc_unpack(R.pc_diff[base + k]); acc.x += (unsigned)d.x; acc.y += (unsigned)d.y; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { acc.x += __shfl_xor_sync(FULL, acc.x, o); acc.y += __shfl_xor_sync(FULL, acc.y, o); }
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint2 t = make_uint2(0, 0);
        for (int w = 0; w < SCAN_THREADS / 32; w++) { t.x += wsum[w].x; t.y += wsum[w].y; }
        block_sums[blockIdx.x] = t;
    }
}

// single block: exclusive scan of the block sums, seeded with the carry-in (:91-92)
__global__ void __launch_bounds__(1024) k_scan2(RegionDev R, uint2* block_sums, int nblocks) {
    __shared__ uint2 wtot[32];
    __shared__ uint2 carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = make_uint2((unsigned)R.sc->phys_cov_start, (unsigned)R.sc->insert_size_start);
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + threadIdx.x;
        uint2 v = i < nblocks ? block_sums[i] : make_uint2(0, 0);
        const uint2 inc = warp_incl_scan(v, lane);
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        if (warp == 0) { uint2 t = wtot[lane]; t = warp_incl_scan(t, lane); wtot[lane] = t; }
        __syncthreads();
        uint2 ex = make_uint2(inc.x - v.x, inc.y - v.y);
        if (warp > 0) { ex.x += wtot[warp - 1].x; ex.y += wtot[warp - 1].y; }
        const uint2 c = carry_s;
        if (i < nblocks) block_sums[i] = make_uint2(ex.x + c.x, ex.y + c.y);
        __syncthreads();
        if (threadIdx.x == 0) { carry_s.x += wtot[31].x; carry_s.y += wtot[31].y; }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan3(RegionDev R, const uint2* block_sums) {
    __shared__ uint2 wtot[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // blocked arrangement: thread t owns items [t*ITEMS, t*ITEMS + ITEMS) of the tile; 16-byte loads / stores
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    const bool full = base + SCAN_ITEMS <= R.size;
    uint2 v[SCAN_ITEMS];
    uint2 run = make_uint2(0, 0);
    if (full) {
        int4* p = reinterpret_cast<int4*>(R.pc_diff + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 2; k++) {
            const int4 d = p[k];
            p[k] = make_int4(0, 0, 0, 0);
            const int2 a = pc_unpack(make_int2(d.x, d.y)), b = pc_unpack(make_int2(d.z, d.w));
            run.x += (unsigned)a.x; run.y += (unsigned)a.y; v[2 * k] = run;
            run.x += (unsigned)b.x; run.y += (unsigned)b.y; v[2 * k + 1] = run;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            const int64_t i = base + k;
            int2 d = make_int2(0, 0);
            if (i < R.size) { d = pc_unpack(R.pc_diff[i]); R.pc_diff[i] = make_int2(0, 0); }
            run.x += (unsigned)d.x; run.y += (unsigned)d.y;
            v[k] = run;
        }
    }
    const uint2 inc = warp_incl_scan(run, lane);
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    uint2 off = block_sums[blockIdx.x];
    off.x += inc.x - run.x; off.y += inc.y - run.y;
    for (int w = 0; w < warp; w++) { off.x += wtot[w].x; off.y += wtot[w].y; }
    int32_t pc[SCAN_ITEMS], is[SCAN_ITEMS];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        pc[k] = (int32_t)(v[k].x + off.x);
        is[k] = (int32_t)(v[k].y + off.y);
        if (pc[k] > 0) is[k] /= pc[k];                                                          // :97-99
    }
    if (full) {
        int4* op = reinterpret_cast<int4*>(R.o_pc + base); int4* oi = reinterpret_cast<int4*>(R.o_is + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; k++) {
            op[k] = make_int4(pc[4 * k], pc[4 * k + 1], pc[4 * k + 2], pc[4 * k + 3]);
            oi[k] = make_int4(is[4 * k], is[4 * k + 1], is[4 * k + 2], is[4 * k + 3]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < R.size) { R.o_pc[base + k] = pc[k]; R.o_is[base + k] = is[k]; }
    }
}

// ---------------------------------------------------------------------------------------------
// k_spill: GenomeRegion.scala:259-264.  Pass 1 walks loci in ascending order; a homozygous
// deletion called at locus i marks i+1 .. i+len-1 deleted and adds its `deletions` to theirs;
// deleted loci make no calls.  Candidates (computed in parallel, ignoring `deleted`) are sorted
// by locus and accepted/rejected by the reference's watermark rule, then applied in parallel.
// The watermark is sequential only inside a CHAIN of overlapping candidates: a candidate that
// starts right of the end of every earlier candidate (a prefix maximum) is accepted whatever
// happened before it, and the rule restarts there -- so one thread per chain head walks its
// chain (chains are one or two entries long).  The list is small (one entry per homozygous
// deletion call): every CTA sorts its own copy in shared memory (up to SPILL_SMEM_CAP entries;
// beyond that one CTA works in a global scratch) and applies its share of the accepted ones.
// ---------------------------------------------------------------------------------------------
static constexpr uint32_t SPILL_SMEM_CAP = 8192;      // 128 KB of int4
static constexpr uint32_t SPILL_CTAS = 37;            // every CTA sorts its own copy of the list, then applies its share
__device__ __forceinline__ void bitonic_sort_by_x(int4* a, uint32_t npow2) {
    for (uint32_t k = 2; k <= npow2; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < npow2; t += blockDim.x) {
                const uint32_t ixj = t ^ j;
                if (ixj > t) {
                    const int4 A = a[t], Bv = a[ixj];
                    const bool up = (t & k) == 0;
                    if ((A.x > Bv.x) == up) { a[t] = Bv; a[ixj] = A; }
                }
            }
            __syncthreads();
        }
}

__global__ void __launch_bounds__(1024) k_spill(RegionDev R, int4* scratch, uint32_t scratch_cap) {
    extern __shared__ int4 sm4[];
    const uint32_t n = min(R.sc->n_cand, R.cand_cap);
    if (n == 0) return;
    uint32_t p2 = 1; while (p2 < n) p2 <<= 1;
    int4* a = (p2 <= SPILL_SMEM_CAP) ? sm4 : scratch;     // else the global scratch, and one CTA does it all
    if (a == scratch && blockIdx.x) return;
    if (p2 > scratch_cap && a == scratch) { if (threadIdx.x == 0) atomicOr(&R.sc->error, 2); return; }
    for (uint32_t t = threadIdx.x; t < p2; t += blockDim.x) a[t] = t < n ? R.cand[t] : make_int4(0x7fffffff, 0, 0, 0);
    __syncthreads();
    bitonic_sort_by_x(a, p2);
    // chain heads: x > max over ALL earlier candidates of their last deleted locus (thread t owns entries [t*per, t*per+per))
    __shared__ long long s_max[1024];
    const uint32_t per = (n + blockDim.x - 1) / blockDim.x, b0 = threadIdx.x * per, b1 = min(b0 + per, n);
    auto end_of = [&](const int4& c) { return (long long)c.x + c.z - 1; };
    long long m = -1;
    for (uint32_t t = b0; t < b1; t++) m = max(m, end_of(a[t]));
    s_max[threadIdx.x] = m;
    __syncthreads();
    for (uint32_t o = 1; o < blockDim.x; o <<= 1) {           // inclusive max-scan of the per-thread maxima
        const long long v = threadIdx.x >= o ? s_max[threadIdx.x - o] : -1;
        __syncthreads();
        s_max[threadIdx.x] = max(s_max[threadIdx.x], v);
        __syncthreads();
    }
    m = threadIdx.x ? s_max[threadIdx.x - 1] : -1;            // everything left of my entries
    for (uint32_t t = b0; t < b1; t++) {
        const int4 c = a[t];
        a[t].w = (long long)c.x > m ? 2 : 0;                  // w: 2 = chain head, 1 = rejected
        m = max(m, end_of(c));
    }
    __syncthreads();
    // the reference's watermark over (locus, length), restarted at every chain head
    for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) {
        if (a[t].w != 2) continue;
        long long deleted_until = end_of(a[t]);
        for (uint32_t u = t + 1; u < n; u++) {
            const int4 c = a[u];
            if (c.w == 2) break;                              // the next chain: somebody else's
            if ((long long)c.x <= deleted_until) { a[u].w = 1; continue; }      // it is deleted: makes no call
            deleted_until = max(deleted_until, end_of(c));
        }
    }
    __syncthreads();
    // apply: every deleted locus belongs to exactly one accepted deletion.  Warp per candidate, lane per deleted locus --
    // the per-locus work is a chain of dependent loads and a BaseCall (~2 us): a thread per candidate made the kernel as
    // slow as the longest deletion -- and the candidates are dealt to all CTAs of the grid (each has sorted its own copy).
    const uint32_t lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (uint32_t t = blockIdx.x * wpb + (threadIdx.x >> 5); t < n; t += gridDim.x * wpb) {
        const int4 cd = a[t];
        if (cd.w == 1) continue;
        const int32_t loc = cd.x, d = cd.y, L = cd.z;
        for (int32_t j = 1 + (int32_t)lane; j < L; j += 32) {
            const int64_t i = (int64_t)loc + j;
            const int32_t nd = wrap32((int64_t)R.o_del[i] + d);                                  // :263
            R.o_del[i] = nd;
            const int4 c = reinterpret_cast<const int4*>(R.o_cnt)[i];
            const longlong2 qa = reinterpret_cast<const longlong2*>(R.o_qs)[2 * i], qb = reinterpret_cast<const longlong2*>(R.o_qs)[2 * i + 1];
            CallIn in;
            in.c[0] = c.x; in.c[1] = c.y; in.c[2] = c.z; in.c[3] = c.w;
            in.q[0] = qa.x; in.q[1] = qa.y; in.q[2] = qb.x; in.q[3] = qb.y;
            in.mqSum = R.o_mq[i]; in.qSum = R.o_q[i]; in.ins = R.o_ins[i]; in.del = nd;
            in.insQual = R.o_insq[i]; in.delQual = R.o_delq[i];
            const uint32_t gi = R.r_gins[i], gd = R.r_gdel[i];
            in.gins = gi ? &R.groups[gi - 1] : nullptr; in.gdel = gd ? &R.groups[gd - 1] : nullptr;
            R.o_call[i] = compute_call(R.cfg, in, nullptr);          // what Vcf.writeRecord recomputes (Vcf.scala:78-79)
            R.o_cov[i] = wrap32((int64_t)c.x + c.y + c.z + c.w + nd);                            // :247 sees the spilled count
            R.o_flags[i] = PB_FL_DELETED;                                                        // :262, and :255 makes no call
        }
    }
}

}  // namespace pb
