// pb_pileup7c.cuh -- the scatter kernel for DEEP, NARROW regions (an amplicon at 5000x: BASELINE config 5): a thread-block
// CLUSTER shares one tile.
//
// k_pileup7 gives a 2048-locus tile to one CTA; a 200 kb region has 98 of them -- two thirds of a B200's 148 SMs x 2 CTA slots
// stay empty -- and at 5000x a tile sees 80 000 descriptors, i.e. twenty folds of its 12-bit counters through global memory.
// Here a tile is 512 loci and belongs to a cluster of up to 8 CTAs.  Every CTA of the cluster keeps its OWN copy of the
// tile's counters in its shared memory and draws grabs of 16 descriptors from ONE cursor for the whole cluster (a word of
// rank 0's shared memory, remote atomics) -- the reductions commute, so any split is exact.  The 12-bit / 20-bit packed
// words are folded into wide counters (32-bit counts, 64-bit quality sums) that also live in shared memory, so depth never
// touches global memory.  A grab takes every ng-th descriptor of the position-sorted list (PileBatches.spread): at 5000x
// sixteen NEIGHBOURS start within half a locus of each other and every lane of a reduction would hit the same address.
// After a cluster barrier the CTAs split the tile's windows among themselves and each sums its loci's wide counters over
// all copies through DISTRIBUTED SHARED MEMORY (cluster.map_shared_rank -> ld.shared::cluster) before the usual per-locus
// epilogue.  Same arithmetic as k_pileup7, same results bit for bit.
#pragma once
#include <cooperative_groups.h>

#include "pb_pileup7.cuh"

namespace pb {

namespace cg = cooperative_groups;

static constexpr int P7C_TILE = 512;                 // default tile (PB_CTILE=1024 for experiments)
static constexpr int P7C_MAX_CLUSTER = 8;            // portable cluster size limit

struct __align__(16) Wide7c { unsigned long long q[4]; uint32_t c[4]; uint32_t mq, qs, bp, nf; };    // 64 B per locus

template <int T>
struct __align__(16) Tile7c {
    uint32_t A[4][T];     // count << 20 | sum of quals, per letter          } same layout as Tile7: scatter_chunk and
    int32_t Bq[4][T];     // sum of qual * (mq1 - dom)                        } scatter_chunk_dmq address it the same way
    int32_t C[T];         // sum of (mq1 - dom)
    uint32_t X[T];        // badPair << 16 | counted bases outside fragCoverage
    Wide7c W[T];          // what the passes so far have folded
    uint32_t grab0[PB_MAXB + 1];
    uint32_t slo[PB_MAXB], nseg[PB_MAXB];
    uint32_t dom, next;          // next: grabs this CTA has taken in the current pass
    uint32_t cnext;              // rank 0's copy: the CLUSTER's cursor over the tile's flat grab list (remote atomics from the other ranks)
    uint32_t exhausted;          // somebody saw the cluster's cursor run past the end
    int32_t read_count, min_depth;
    uint2 slow[P7_WARPS][P7_SLOW_CAP];
};

template <bool MINQ, int T>
__global__ void __launch_bounds__(P7_WARPS * 32, 2) k_pileup7c(const RegionDev R, const PileBatches PB) {
    extern __shared__ __align__(16) uint8_t smem_raw7c[];
    Tile7c<T>& S = *reinterpret_cast<Tile7c<T>*>(smem_raw7c);
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t CL = cluster.num_blocks(), rank = cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t t0 = (int32_t)(blockIdx.x / CL) * T;
    const int n_batches = PB.n;
    const int min_qual = R.cfg.min_qual;
    const uint32_t defq = (uint32_t)R.cfg.default_qual;
    const uint32_t minq_add = (uint32_t)(0x80 - (min_qual > 128 ? 128 : min_qual)) * 0x01010101u;
    constexpr uint32_t OFF_X = 36u * T;

    // ---- tile set-up (every CTA of the cluster computes the same candidate ranges) ----
    {
        uint4* z = reinterpret_cast<uint4*>(&S.A[0][0]);
        for (int i = tid; i < (int)((10 * T * 4 + sizeof(Wide7c) * T) / 16); i += P7_WARPS * 32) z[i] = make_uint4(0, 0, 0, 0);
        if (tid == 0) { S.dom = 0; S.next = 0; S.cnext = 0; S.exhausted = 0; S.read_count = R.sc->read_count; S.min_depth = R.sc->min_depth; }
        if (warp == 0) {
            uint32_t my_slo = 0, my_nseg = 0;
            if (lane < n_batches) {
                const PileBatch& Bl = PB.b[lane];
                if (Bl.flags & 2) {
                    const int64_t x = (int64_t)t0 - Bl.reach[0] + 1;
                    const int64_t y = (int64_t)t0 + T + Bl.reach[1];
                    int64_t khi = (y + 31) >> 5; if (khi > R.n_win) khi = R.n_win;
                    my_slo = x <= 0 ? 0u : Bl.win_first[x >> 5];
                    const uint32_t shi = (y > ((int64_t)R.n_win << 5)) ? Bl.n_cigar : Bl.win_first[khi];
                    my_nseg = shi > my_slo ? shi - my_slo : 0u;
                }
            }
            uint32_t ng = (my_nseg + P7_GRAB - 1) / P7_GRAB, pre = ng;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, pre, o); if (lane >= o) pre += v; }
            if (lane < PB_MAXB) { S.slo[lane] = my_slo; S.nseg[lane] = my_nseg; S.grab0[lane] = pre - ng; }
            if (lane == PB_MAXB - 1) S.grab0[PB_MAXB] = pre;
        }
    }
    cluster.sync();                 // the set-up is visible, and rank 0's cursor is zero before anybody draws from it
    const uint32_t total_grabs = S.grab0[PB_MAXB];
    // The tile's grabs are handed out by ONE cursor for the whole cluster (in rank 0's shared memory, drawn from with
    // remote atomics): the CTAs of a cluster share their SMs with CTAs of other clusters and run at different speeds, and a
    // static split left a third of all warp time waiting at the cluster barrier (profiles/README.md, r2q).
    uint32_t* const cursor = cluster.map_shared_rank(&S.cnext, 0);
    // fold the 12/20-bit tile into the wide counters and clear it
    auto fold = [&]() {
        const uint32_t dom = S.dom;
        for (int l = tid; l < T; l += P7_WARPS * 32) {
            uint32_t c[4], sq[4]; Wide7c w = S.W[l];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint32_t a = S.A[b][l]; c[b] = a >> 20; sq[b] = a & 0xFFFFFu;
                w.c[b] += c[b];
                w.q[b] += (unsigned long long)((long long)((uint64_t)dom * sq[b]) + (long long)S.Bq[b][l]);
                S.A[b][l] = 0; S.Bq[b][l] = 0;
            }
            const uint32_t n = c[0] + c[1] + c[2] + c[3], x = S.X[l];
            w.mq += dom * n + (uint32_t)S.C[l]; w.qs += sq[0] + sq[1] + sq[2] + sq[3];
            w.bp += x >> 16; w.nf += x & 0xFFFFu;
            S.C[l] = 0; S.X[l] = 0;
            S.W[l] = w;
        }
    };

    const uint32_t sA = smem_u32(&S.A[0][0]);
    uint32_t dom_r = 0;
    constexpr uint32_t PASS_GRABS = P7_PASS_DESC / P7_GRAB;
    uint32_t slow_n = 0;
    auto drain_slow = [&]() {
        __syncwarp();
        const uint32_t dom = dom_r;
        for (uint32_t e = (uint32_t)(lane >> 4); e < slow_n; e += 2u) {
            const uint2 en = S.slow[warp][e];
            const PileBatch& Bb = PB.b[en.x];
            const Seg sg = Bb.seg[en.y];
            const int32_t cA = sg.loc0 > t0 ? sg.loc0 : t0;
            const int32_t cBx = sg.loc0 + sg.len < t0 + T ? sg.loc0 + sg.len : t0 + T;
            const int32_t n = cBx - cA;
            const uint32_t src = sg.src + (uint32_t)(cA - sg.loc0), last = src + (uint32_t)n - 1u;
            const int32_t col = cA - t0;
            const int32_t dmq = (int32_t)(sg.w & 0xFFFFu) - (int32_t)dom;
            const bool hasq = sg.w & SEG_HASQ;
            const uint32_t qand = hasq ? 0x7Fu : 0u, qor = hasq ? 0u : defq;
            const uint32_t nohq_pass = (!hasq && (int)defq >= min_qual) ? 0x01010101u : 0u;
            const uint4* qp = reinterpret_cast<const uint4*>(Bb.quals);
            const uint32_t* cp = reinterpret_cast<const uint32_t*>(Bb.bases2);
            for (uint32_t k = (src >> 4) + (uint32_t)(lane & 15); k <= (last >> 4); k += 16) {
                const uint4 Q = qp[k]; const uint32_t cw = cp[k];
                const uint32_t okm = chunk_mask<MINQ>(Q, k, src, last, minq_add, nohq_pass);
                scatter_chunk_dmq<T>(Q, cw, okm, sA + 4u * (uint32_t)(col + (int32_t)(16u * k - src)), qand, qor, dmq);
            }
        }
        slow_n = 0;
        __syncwarp();
    };

    // ---- scatter: passes of <= 4064 descriptors per CTA (12-bit counts), then a fold into the wide counters ----
    for (bool first_pass = true;; first_pass = false) {
        if (!first_pass) {
            fold();
            if (tid == 0) S.next = 0;
            __syncthreads();
        }
        constexpr uint32_t p1 = PASS_GRABS;
        uint32_t g_nx = p1, sidx_nx = 0; int b_nx = 0; Seg seg_nx = {0, 0, 0, 0};
        auto fetch = [&]() {
            uint32_t g = 0xFFFFFFFFu;
            if (lane == 0) {
                if (atomicAdd(&S.next, 1u) < p1) {                // a local ticket of this pass, then the cluster's cursor
                    g = atomicAdd(cursor, 1u);
                    if (g >= total_grabs) { g = 0xFFFFFFFFu; S.exhausted = 1u; }
                }
            }
            g = __shfl_sync(FULL, g, 0);
            g_nx = g == 0xFFFFFFFFu ? p1 : 0u; seg_nx = Seg{0, 0, 0, 0};
            if (g == 0xFFFFFFFFu) return;
            int b = 0;
            while (g >= S.grab0[b + 1]) b++;
            b_nx = b;
            // PB.spread: the 16 descriptors of a grab are not neighbours of the position-sorted list (at 5000x those start
            // within half a locus of each other: every lane of a reduction would hit the same one or two addresses, and
            // same-address shared-memory reductions serialise -- 6.7 wavefronts per instruction measured) but ng apart,
            // i.e. spread evenly over the tile
            const uint32_t gl = g - S.grab0[b], ngb = S.grab0[b + 1] - S.grab0[b];
            const uint32_t di = PB.spread ? (uint32_t)(lane >> 1) * ngb + gl : gl * P7_GRAB + (uint32_t)(lane >> 1);
            sidx_nx = S.slo[b] + di;
            if (di < S.nseg[b]) seg_nx = PB.b[b].seg[sidx_nx];
        };
        fetch();
        while (g_nx < p1) {
            const Seg mine = seg_nx;
            const int b_cur = b_nx; const uint32_t sidx = sidx_nx;
            const PileBatch& Bb = PB.b[b_cur];
            fetch();
            const int h = lane & 1;
            const uint8_t* __restrict__ gquals = Bb.quals;
            const uint8_t* __restrict__ gbases = Bb.bases2;
            const bool nf = !(Bb.flags & 1);
            const int32_t cA = mine.loc0 > t0 ? mine.loc0 : t0;
            const int32_t cBx = mine.loc0 + mine.len < t0 + T ? mine.loc0 + mine.len : t0 + T;
            const int32_t n = mine.len > 0 ? (cBx > cA ? cBx - cA : 0) : 0;
            const uint32_t src = mine.src + (uint32_t)(cA - mine.loc0);
            const int32_t col = cA - t0;
            const bool valid = mine.w & SEG_VALID;
            const bool live = n > 0 && valid;
            const uint32_t last = src + (uint32_t)n - 1u;
            const uint32_t c0 = src >> 4, c1 = last >> 4, mid = c0 + ((c1 - c0 + 2u) >> 1);
            uint32_t k = h ? mid : c0;
            const uint32_t k1 = h ? c1 : mid - 1u;
            const bool work = live && k <= k1;
            const uint4* qp = reinterpret_cast<const uint4*>(gquals) + k;
            const uint32_t* cp = reinterpret_cast<const uint32_t*>(gbases) + k;
            uint4 Q = make_uint4(0, 0, 0, 0); uint32_t cw = 0;
            if (work) { Q = *qp; cw = *cp; }
            unsigned badm = __ballot_sync(FULL, n > 0 && !valid && h == 0);
            while (badm) {                                        // PileUpRegion.scala:45
                const int j = __ffs(badm) - 1; badm &= badm - 1;
                const int32_t bn = __shfl_sync(FULL, n, j), bcol = __shfl_sync(FULL, col, j);
                for (int i = lane; i < bn; i += 32) red_shared_add(sA + OFF_X + 4u * (uint32_t)(bcol + i), 0x10000u);
            }
            const unsigned livem = __ballot_sync(FULL, live);
            if (livem == 0) continue;
            const uint32_t mq1 = mine.w & 0xFFFFu;
            if (dom_r == 0) {
                const uint32_t first = __shfl_sync(FULL, mq1, __ffs(livem) - 1);
                uint32_t old = 0;
                if (lane == 0) old = atomicCAS(&S.dom, 0u, first);
                old = __shfl_sync(FULL, old, 0);
                dom_r = old ? old : first;
            }
            const int32_t dmq = (int32_t)mq1 - (int32_t)dom_r;
            int inl = 0;
            {
                const unsigned qm = __ballot_sync(FULL, live && dmq != 0 && h == 0);
                if (qm) {
                    const uint32_t e = slow_n + (uint32_t)__popc(qm & ((1u << lane) - 1u));
                    if ((qm >> lane) & 1u) { if (e < P7_SLOW_CAP) S.slow[warp][e] = make_uint2((uint32_t)b_cur, sidx); else inl = 1; }
                    slow_n = min(slow_n + (uint32_t)__popc(qm), (uint32_t)P7_SLOW_CAP);
                }
            }
            inl = __shfl_sync(FULL, inl, lane & ~1);
            const bool hasq = mine.w & SEG_HASQ;
            const bool allhq = __all_sync(FULL, hasq || !live);
            const uint32_t qand = hasq ? 0x7Fu : 0u, qor = (1u << 20) | (hasq ? 0u : defq);
            const uint32_t nohq_pass = (!hasq && (int)defq >= min_qual) ? 0x01010101u : 0u;
            if (work) {
                uint32_t sa = sA + 4u * (uint32_t)(col + (int32_t)(16u * k - src));
                for (;;) {
                    uint4 Qn = make_uint4(0, 0, 0, 0); uint32_t cwn = 0;
                    const bool more = k < k1;
                    if (more) { Qn = qp[1]; cwn = cp[1]; }
                    const uint32_t okm = chunk_mask<MINQ>(Q, k, src, last, minq_add, nohq_pass);
                    if (allhq) { if (nf) scatter_chunk<true, true, T>(Q, cw, okm, sa, qand, qor); else scatter_chunk<false, true, T>(Q, cw, okm, sa, qand, qor); }
                    else { if (nf) scatter_chunk<true, false, T>(Q, cw, okm, sa, qand, qor); else scatter_chunk<false, false, T>(Q, cw, okm, sa, qand, qor); }
                    if (inl) scatter_chunk_dmq<T>(Q, cw, okm, sa, qand, qor, dmq);
                    if (!more) break;
                    Q = Qn; cw = cwn; k++; qp++; cp++; sa += 64;
                }
            }
        }
        drain_slow();
        __syncthreads();
        if (S.exhausted) break;     // written before the barrier by whoever drew past the end
    }
    fold();                         // everything of mine is in the wide counters now
    cluster.sync();                 // ... and everybody else's in theirs (barrier.cluster: also orders the shared-memory writes)

    // ---- epilogue: the tile's 16 windows are dealt round-robin to the CTAs of the cluster; a warp per window sums its loci's
    // wide counters over all CL copies through distributed shared memory ----
    const int2 rc_md = make_int2(S.read_count, S.min_depth);
    for (int wl = (int)rank + (int)CL * warp; wl < T / 32; wl += (int)CL * P7_WARPS) {
        const int64_t w = ((int64_t)t0 >> 5) + wl;
        if (w >= R.n_win) break;
        const int l = wl * 32 + lane;
        const int64_t loc = (int64_t)t0 + l;
        const bool inr = loc < R.size;
        const uint32_t pre_rb = R.rare_bits[w];
        const uint8_t pre_ref = inr ? ref_at(R, (int64_t)R.start + loc) : (uint8_t)'N';
        uint32_t c[4] = {0, 0, 0, 0}; uint64_t q[4] = {0, 0, 0, 0};
        uint32_t mqS = 0, qS = 0, bp = 0, nfc = 0;
        for (uint32_t rr = 0; rr < CL; rr++) {
            const Wide7c* wp = cluster.map_shared_rank(&S.W[l], rr);             // the same locus in CTA rr's shared memory
            const uint4 a0 = reinterpret_cast<const uint4*>(wp)[0], a1 = reinterpret_cast<const uint4*>(wp)[1];
            const uint4 a2 = reinterpret_cast<const uint4*>(wp)[2], a3 = reinterpret_cast<const uint4*>(wp)[3];
            q[0] += (uint64_t)a0.x | ((uint64_t)a0.y << 32); q[1] += (uint64_t)a0.z | ((uint64_t)a0.w << 32);
            q[2] += (uint64_t)a1.x | ((uint64_t)a1.y << 32); q[3] += (uint64_t)a1.z | ((uint64_t)a1.w << 32);
            c[0] += a2.x; c[1] += a2.y; c[2] += a2.z; c[3] += a2.w;
            mqS += a3.x; qS += a3.y; bp += a3.z; nfc += a3.w;
        }
        const uint32_t n = c[0] + c[1] + c[2] + c[3];
        finish_locus(R, w, lane, (int32_t)loc, c, q, mqS, qS, bp, n - nfc, pre_rb, pre_ref, rc_md);
    }
    cluster.sync();                 // nobody leaves while a neighbour may still read its counters
}

}  // namespace pb
