// pb_pileup7.cuh -- the hot kernel, seventh generation: scatter into a shared-memory tile.
//
// The gather kernels (k_pileup .. k_pileup5) give a 32-locus window to a warp and bring every
// overlapping read segment to it; a 130-base segment meets five windows, and each of those meetings
// costs descriptor handling, staging geometry and realignment -- two thirds of k_pileup5's issue slots
// (profiles/README.md, r1p).  This kernel turns the loop around:
//
//   * a CTA owns a tile of T loci and keeps PileUp's hot counters for it in shared memory;
//   * its warps take the segments that overlap the tile, 32 descriptors per grab from a shared cursor,
//     LANE <-> HALF A SEGMENT: every per-segment computation (clipping, addresses, flags) is done once per
//     lane for 16 segments at a time, and a lane then walks its half in aligned 16-base chunks --
//     one 16-byte load of quality bytes, one 4-byte load of 2-bit codes -- and issues for every counted
//     base ONE native 32-bit shared-memory reduction,
//         A[letter][locus] += 1 << 20 | qual        (12-bit count | 20-bit quality sum).
//     The 32 lanes of a reduction hit 32 unrelated loci (~3 wavefronts per instruction instead of 1);
//     that is the price for paying ~8 thread instructions per base and nothing per (segment, window);
//   * PileUp.add's other three sums (PileUp.scala:75-84) follow from that pair for every read whose
//     (adjMq + 1) equals the tile's reference value `dom` (the first one the CTA meets):
//         qualSum[b] = dom * sum_q[b] + Bq[b],   mqSum = dom * count + C,   qSum = sum_b sum_q[b]
//     where Bq / C collect `qual * (mq1 - dom)` and `mq1 - dom` of the other reads (two more reductions
//     per base for those only).  badPair and the bases outside fragCoverage share a fourth word;
//   * the loads of a lane's next chunk are issued before the reductions of the current one;
//   * every segment is met once per tile (T = 2048: 6 % halo instead of 5x), quality bytes go from
//     L2 straight into registers, nothing is staged or flushed;
//   * 12-bit counts: after 4064 descriptors the tile is folded into the 32/64-bit output planes
//     (deep pile-ups only), then the epilogue (finish_locus) runs per locus as before.
//
// Integer atomics commute, so the result is bit-identical to the sequential walk of the reference.
#pragma once
#include "pb_pileup5.cuh"

namespace pb {

static constexpr int P7_WARPS = 16;
static constexpr int P7_TILE = 2048;                // loci per CTA
static constexpr int P7_GRAB = 16;                  // descriptors per grab from the tile's cursor (two lanes per descriptor)
static constexpr int P7_PASS_DESC = 4064;           // descriptors per pass <= 4095 (12-bit count)
static constexpr int P7_SLOW_CAP = 32;              // per warp and pass: queued segments with another mapping quality

__device__ __forceinline__ void red_shared_add(uint32_t saddr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ void red_shared_add_if(uint32_t nz, uint32_t saddr, uint32_t v) {   // predicated, no branch
    asm volatile("{\n .reg .pred pp;\n setp.ne.u32 pp, %0, 0;\n @pp red.shared.add.u32 [%1], %2;\n}" ::"r"(nz), "r"(saddr), "r"(v) : "memory");
}

template <int T>
struct __align__(16) Tile7 {
    uint32_t A[4][T];            // count << 20 | sum of quals, per letter
    int32_t Bq[4][T];            // sum of qual * (mq1 - dom)
    int32_t C[T];                // sum of (mq1 - dom)
    uint32_t X[T];               // badPair << 16 | counted bases outside fragCoverage
    uint32_t grab0[PB_MAXB + 1]; // first flat grab index of every batch
    uint32_t slo[PB_MAXB], nseg[PB_MAXB];
    uint32_t dom;                // 0 = not chosen yet ((adjMq + 1) >= 1 always)
    uint32_t next;               // next flat grab index to hand out
    int32_t read_count, min_depth;   // the region's scalars (k_fold), fetched during set-up for the epilogue
    uint2 slow[P7_WARPS][P7_SLOW_CAP];   // per warp: (batch, descriptor index) of segments whose (adjMq + 1) != dom
};

// if (bits & mask) { A word += v;  (NF) X word += 1 }   -- one predicate for both reductions
template <bool NF>
__device__ __forceinline__ void red_counted(uint32_t bits, uint32_t mask, uint32_t saddr, uint32_t v, uint32_t xaddr) {
    if (NF)
        asm volatile("{\n .reg .pred pp;\n .reg .b32 tt;\n and.b32 tt, %0, %1;\n setp.ne.u32 pp, tt, 0;\n @pp red.shared.add.u32 [%2], %3;\n @pp red.shared.add.u32 [%4], 1;\n}"
                     ::"r"(bits), "r"(mask), "r"(saddr), "r"(v), "r"(xaddr) : "memory");
    else
        asm volatile("{\n .reg .pred pp;\n .reg .b32 tt;\n and.b32 tt, %0, %1;\n setp.ne.u32 pp, tt, 0;\n @pp red.shared.add.u32 [%2], %3;\n}"
                     ::"r"(bits), "r"(mask), "r"(saddr), "r"(v) : "memory");
}

// 16 bases of one lane's segment: Q = their quality bytes, cw = their 2-bit codes, okm = which of them are counted,
// sa = shared byte address of A[0][locus of base 0 of the chunk].  Per base: PRMT (value), LOP + IMAD (address of
// the letter's word), LOP3 -> predicate, RED.  HQ: every lane's read has qualities (the byte goes straight into the
// packed value 1 << 20 | q); otherwise qand / qor substitute default_qual per lane.
template <bool NF, bool HQ, int T>
__device__ __forceinline__ void scatter_chunk(const uint4 Q, uint32_t cw, uint32_t okm, uint32_t sa, uint32_t qand, uint32_t qor) {
    constexpr uint32_t OFF_X = 36u * T;                                        // byte offset of X[.] from A[0][.]
    constexpr int LOG = T == 512 ? 11 : T == 1024 ? 12 : 13;                   // log2 of the letter stride in bytes
    const uint32_t cwm = cw >> 14, cwh = cw >> 28;                             // codes of bases 7.. and 14.. at bit 0
#pragma unroll
    for (int b = 0; b < 16; b++) {
        const uint32_t Qw = b < 4 ? Q.x : b < 8 ? Q.y : b < 12 ? Q.z : Q.w;
        const uint32_t val = HQ ? __byte_perm(Qw, 0x00100000u, 0x7650 | (b & 3))
                                : ((__byte_perm(Qw, 0, 0x4440 | (b & 3)) & qand) | qor);   // 1 << 20 | q
        const int pos = b < 7 ? 2 * b : b < 14 ? 2 * (b - 7) : 2 * (b - 14);   // bit position of the code in its copy
        const uint32_t code = (b < 7 ? cw : b < 14 ? cwm : cwh) & (3u << pos);
        const uint32_t a = (LOG >= pos ? code * (1u << (LOG >= pos ? LOG - pos : 0)) : code >> (pos > LOG ? pos - LOG : 0)) + sa;   // + letter * 4 T
        red_counted<NF>(okm, 1u << b, a + 4u * b, val, sa + 4u * b + OFF_X);
    }
}

// Which of the 16 bases of aligned chunk k are counted: inside [src, last], quality byte without the 0x80 mark, >= minQual.
template <bool MINQ>
__device__ __forceinline__ uint32_t chunk_mask(const uint4 Q, uint32_t k, uint32_t src, uint32_t last, uint32_t minq_add, uint32_t nohq_pass) {
    const uint32_t lo = k == (src >> 4) ? (src & 15u) : 0u, hi = k == (last >> 4) ? (last & 15u) + 1u : 16u;
    // bit 7 of every byte = "not counted": the 0x80 mark, or (MINQ) a quality below minQual unless default_qual decides
    uint32_t u0 = Q.x, u1 = Q.y, u2 = Q.z, u3 = Q.w;
    if (MINQ) {
        u0 |= ~((((Q.x & 0x7F7F7F7Fu) + minq_add) | (nohq_pass << 7))); u1 |= ~((((Q.y & 0x7F7F7F7Fu) + minq_add) | (nohq_pass << 7)));
        u2 |= ~((((Q.z & 0x7F7F7F7Fu) + minq_add) | (nohq_pass << 7))); u3 |= ~((((Q.w & 0x7F7F7F7Fu) + minq_add) | (nohq_pass << 7)));
    }
    // the four bit-7s of a word gathered into its top nibble by one multiplication (partial products never collide:
    // bit 7 -> 28, bit 15 -> 29, bit 23 -> 30, bit 31 -> 31)
    const uint32_t n0 = ((u0 & 0x80808080u) * 0x00204081u) >> 28, n1 = ((u1 & 0x80808080u) * 0x00204081u) >> 28;
    const uint32_t n2 = ((u2 & 0x80808080u) * 0x00204081u) >> 28, n3 = ((u3 & 0x80808080u) * 0x00204081u) >> 28;
    const uint32_t bad = n0 | (n1 << 4) | (n2 << 8) | (n3 << 12);
    return ((1u << hi) - 1u) & ~((1u << lo) - 1u) & ~bad;
}

// The same 16 bases for a segment whose (adjMq + 1) differs from the tile's reference value by dmq: Bq / C terms.
template <int T>
__device__ __forceinline__ void scatter_chunk_dmq(const uint4 Q, uint32_t cw, uint32_t okm, uint32_t sa, uint32_t qand, uint32_t qor, int32_t dmq) {
    constexpr uint32_t OFF_BQ = 16u * T, OFF_C = 32u * T;
    constexpr int LOG = T == 512 ? 11 : T == 1024 ? 12 : 13;
#pragma unroll
    for (int b = 0; b < 16; b++) {
        const uint32_t Qw = b < 4 ? Q.x : b < 8 ? Q.y : b < 12 ? Q.z : Q.w;
        const uint32_t q = ((__byte_perm(Qw, 0, 0x4440 | (b & 3)) & qand) | qor) & 0xFFu;
        const uint32_t letter = (2 * b <= LOG ? (cw << (LOG - 2 * b)) : (cw >> (2 * b - LOG))) & (3u << LOG);
        asm volatile("{\n .reg .pred pp;\n .reg .b32 tt;\n and.b32 tt, %0, %1;\n setp.ne.u32 pp, tt, 0;\n @pp red.shared.add.u32 [%2], %3;\n @pp red.shared.add.u32 [%4], %5;\n}"
                     ::"r"(okm), "r"(1u << b), "r"(sa + letter + 4u * b + OFF_BQ), "r"((uint32_t)((int32_t)q * dmq)),
                       "r"(sa + 4u * b + OFF_C), "r"((uint32_t)dmq) : "memory");
    }
}

template <bool MINQ, int T>
__global__ void __launch_bounds__(P7_WARPS * 32, 2) k_pileup7(const RegionDev R, const PileBatches PB) {
    extern __shared__ __align__(16) uint8_t smem_raw7[];
    Tile7<T>& S = *reinterpret_cast<Tile7<T>*>(smem_raw7);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t t0 = (int32_t)blockIdx.x * T;
    const int n_batches = PB.n;
    const int min_qual = R.cfg.min_qual;
    const uint32_t defq = (uint32_t)R.cfg.default_qual;
    const uint32_t minq_add = (uint32_t)(0x80 - (min_qual > 128 ? 128 : min_qual)) * 0x01010101u;
    constexpr uint32_t OFF_X = 36u * T;                         // byte offset of X[.] from A[0][.]

    // ---- tile set-up: zero the counters, candidate descriptor range of every batch ----
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(&S.A[0][0]);
        for (int i = tid; i < 10 * T; i += P7_WARPS * 32) z[i] = 0;
        if (tid == 0) { S.dom = 0; S.next = 0; S.read_count = R.sc->read_count; S.min_depth = R.sc->min_depth; }
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
            uint32_t ng = (my_nseg + P7_GRAB - 1) / P7_GRAB, pre = ng;      // inclusive scan of the grab counts
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, pre, o); if (lane >= o) pre += v; }
            if (lane < PB_MAXB) { S.slo[lane] = my_slo; S.nseg[lane] = my_nseg; S.grab0[lane] = pre - ng; }
            if (lane == PB_MAXB - 1) S.grab0[PB_MAXB] = pre;
        }
    }
    __syncthreads();
    const uint32_t total_grabs = S.grab0[PB_MAXB];
    bool folded = false;

    // fold the 12/20-bit tile into the output planes (used as 32/64-bit accumulators) and clear it
    auto fold = [&]() {
        const uint32_t dom = S.dom;
        for (int l = tid; l < T; l += P7_WARPS * 32) {
            const int64_t loc = (int64_t)t0 + l;
            const int x_ = l;
            if (loc < R.size) {
                uint32_t c[4], sq[4]; long long q[4];
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const uint32_t a = S.A[b][x_]; c[b] = a >> 20; sq[b] = a & 0xFFFFFu;
                    q[b] = (long long)((uint64_t)dom * sq[b]) + (long long)S.Bq[b][x_];
                }
                const uint32_t n = c[0] + c[1] + c[2] + c[3];
                const uint32_t mq = dom * n + (uint32_t)S.C[x_], qs = sq[0] + sq[1] + sq[2] + sq[3];
                const uint32_t x = S.X[x_];
                int4* oc = reinterpret_cast<int4*>(R.o_cnt) + loc;
                long long* oq = reinterpret_cast<long long*>(R.o_qs) + 4 * loc;
                if (folded) {
                    const int4 p = *oc;
                    *oc = make_int4(p.x + (int)c[0], p.y + (int)c[1], p.z + (int)c[2], p.w + (int)c[3]);
#pragma unroll
                    for (int b = 0; b < 4; b++) oq[b] += q[b];
                    R.o_mq[loc] += (int32_t)mq; R.o_q[loc] += (int32_t)qs;
                    R.o_bp[loc] += (int32_t)(x >> 16); R.o_frag[loc] += (int32_t)(x & 0xFFFFu);
                } else {
                    *oc = make_int4((int)c[0], (int)c[1], (int)c[2], (int)c[3]);
#pragma unroll
                    for (int b = 0; b < 4; b++) oq[b] = q[b];
                    R.o_mq[loc] = (int32_t)mq; R.o_q[loc] = (int32_t)qs;
                    R.o_bp[loc] = (int32_t)(x >> 16); R.o_frag[loc] = (int32_t)(x & 0xFFFFu);
                }
            }
#pragma unroll
            for (int b = 0; b < 4; b++) { S.A[b][x_] = 0; S.Bq[b][x_] = 0; }
            S.C[x_] = 0; S.X[x_] = 0;
        }
        folded = true;
    };

    // ---- scatter: passes of <= 4064 descriptors; a warp takes the next 16 descriptors when it is free ----
    const uint32_t sA = smem_u32(&S.A[0][0]);
    uint32_t dom_r = 0;                                          // register copy of S.dom once it is known
    static_assert(T == 1024 || T == 2048, "tile size");
    constexpr uint32_t PASS_GRABS = P7_PASS_DESC / P7_GRAB;
    // A warp drains its own queue right after its last grab of a pass (no barrier needed: the reductions commute):
    // two queued segments per iteration, lane <-> chunk (16 lanes per segment)
    uint32_t slow_n = 0;                                         // warp-uniform: entries in S.slow[warp]
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

    for (uint32_t p0 = 0; p0 < total_grabs; p0 += PASS_GRABS) {
        if (p0) {
            fold();
            if (tid == 0) S.next = p0;
            __syncthreads();
        }
        const uint32_t p1 = min(p0 + PASS_GRABS, total_grabs);
        // the next grab (cursor value, batch, my descriptor) is fetched before the current one is processed
        uint32_t g_nx = p1, sidx_nx = 0; int b_nx = 0; Seg seg_nx = {0, 0, 0, 0};
        auto fetch = [&]() {
            uint32_t g = 0;
            if (lane == 0) g = atomicAdd(&S.next, 1u);
            g = __shfl_sync(FULL, g, 0);
            g_nx = g; seg_nx = Seg{0, 0, 0, 0};
            if (g >= p1) return;
            int b = 0;
            while (g >= S.grab0[b + 1]) b++;                      // batch of this grab (grab0 is non-decreasing)
            b_nx = b;
            const uint32_t gl = g - S.grab0[b], ngb = S.grab0[b + 1] - S.grab0[b];
            const uint32_t di = PB.spread ? (uint32_t)(lane >> 1) * ngb + gl : gl * P7_GRAB + (uint32_t)(lane >> 1);   // two lanes per descriptor
            sidx_nx = S.slo[b] + di;
            if (di < S.nseg[b]) seg_nx = PB.b[b].seg[sidx_nx];
        };
        if (!(R.exp_flags & 2)) fetch();
        while (g_nx < p1) {
            const Seg mine = seg_nx;
            const int b_cur = b_nx; const uint32_t sidx = sidx_nx;
            const PileBatch& Bb = PB.b[b_cur];
            fetch();
            // lane (d, h) walks half h of the chunks of descriptor d
            const int h = lane & 1;
            const uint8_t* __restrict__ gquals = Bb.quals;
            const uint8_t* __restrict__ gbases = Bb.bases2;
            const bool nf = !(Bb.flags & 1);                      // warp-uniform: this batch is outside fragCoverage
            // my segment clipped to the tile: n bases from base index src, first locus = tile column col
            const int32_t cA = mine.loc0 > t0 ? mine.loc0 : t0;
            const int32_t cBx = mine.loc0 + mine.len < t0 + T ? mine.loc0 + mine.len : t0 + T;
            const int32_t n = mine.len > 0 ? (cBx > cA ? cBx - cA : 0) : 0;
            const uint32_t src = mine.src + (uint32_t)(cA - mine.loc0);
            const int32_t col = cA - t0;
            const bool valid = mine.w & SEG_VALID;
            const bool live = n > 0 && valid;
            // aligned 16-base chunks c0..c1 of the batch's base stream (chunk k = bases [16 k, 16 k + 16)); my half of them.
            // The first chunk's loads are issued now, the rest of the per-grab set-up runs under their latency.
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
            while (badm) {                                        // PileUpRegion.scala:45: badPair++ on every locus, lane <-> locus
                const int j = __ffs(badm) - 1; badm &= badm - 1;
                const int32_t bn = __shfl_sync(FULL, n, j), bcol = __shfl_sync(FULL, col, j);
                for (int i = lane; i < bn; i += 32) red_shared_add(sA + OFF_X + 4u * (uint32_t)(bcol + i), 0x10000u);
            }
            const unsigned livem = __ballot_sync(FULL, live);
            if (livem == 0) continue;
            const uint32_t mq1 = mine.w & 0xFFFFu;
            if (dom_r == 0) {                                     // the tile's reference (adjMq + 1): first one met
                const uint32_t first = __shfl_sync(FULL, mq1, __ffs(livem) - 1);
                uint32_t old = 0;
                if (lane == 0) old = atomicCAS(&S.dom, 0u, first);
                old = __shfl_sync(FULL, old, 0);
                dom_r = old ? old : first;
            }
            const int32_t dmq = (int32_t)mq1 - (int32_t)dom_r;
            // a segment with another mapping quality (~5 % of the reads) is queued: its Bq / C terms are added after the
            // scatter, lane <-> chunk, instead of dragging the whole warp through a second reduction block here
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
                {
                    uint32_t sa = sA + 4u * (uint32_t)(col + (int32_t)(16u * k - src));      // A[0][locus of base 16 k] (virtual before col)
                    // two register stages, used alternately: the loads of the next chunk are in flight while the reductions of
                    // this one issue, and nothing is copied between the stages
                    auto chunk = [&](const uint4& Qc, const uint32_t cwc) {
                        const uint32_t okm = chunk_mask<MINQ>(Qc, k, src, last, minq_add, nohq_pass);
                        if (allhq) { if (nf) scatter_chunk<true, true, T>(Qc, cwc, okm, sa, qand, qor); else scatter_chunk<false, true, T>(Qc, cwc, okm, sa, qand, qor); }
                        else { if (nf) scatter_chunk<true, false, T>(Qc, cwc, okm, sa, qand, qor); else scatter_chunk<false, false, T>(Qc, cwc, okm, sa, qand, qor); }
                        if (inl) scatter_chunk_dmq<T>(Qc, cwc, okm, sa, qand, qor, dmq);         // slow list full (deep pile-ups)
                    };
                    uint4 Q2 = make_uint4(0, 0, 0, 0); uint32_t cw2 = 0;
                    for (;;) {
                        bool more = k < k1;
                        if (more) { Q2 = qp[1]; cw2 = cp[1]; }
                        chunk(Q, cw);
                        if (!more) break;
                        k++; sa += 64;
                        more = k < k1;
                        if (more) { Q = qp[2]; cw = cp[2]; }
                        chunk(Q2, cw2);
                        if (!more) break;
                        k++; sa += 64; qp += 2; cp += 2;
                    }
                }
            }
        }
        drain_slow();
        __syncthreads();
    }

    // ---- epilogue: warp per 32-locus window of the tile ----
    const uint32_t dom = S.dom;
    const int2 rc_md = make_int2(S.read_count, S.min_depth);
    for (int wl = warp; wl < T / 32; wl += P7_WARPS) {
        const int64_t w = ((int64_t)t0 >> 5) + wl;
        if (w >= R.n_win) break;
        const int l = wl * 32 + lane, x_ = l;
        const int64_t loc = (int64_t)t0 + l;
        const bool inr = loc < R.size;
        const uint32_t pre_rb = R.rare_bits[w];
        const uint8_t pre_ref = inr ? ref_at(R, (int64_t)R.start + loc) : (uint8_t)'N';
        uint32_t c[4], sq[4]; uint64_t q[4];
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const uint32_t a = S.A[b][x_]; c[b] = a >> 20; sq[b] = a & 0xFFFFFu;
            q[b] = (uint64_t)((long long)((uint64_t)dom * sq[b]) + (long long)S.Bq[b][x_]);
        }
        uint32_t n = c[0] + c[1] + c[2] + c[3];
        uint32_t mqS = dom * n + (uint32_t)S.C[x_], qS = sq[0] + sq[1] + sq[2] + sq[3];
        const uint32_t x = S.X[x_];
        uint32_t bp = x >> 16, nfc = x & 0xFFFFu;
        if (folded && inr) {
            const int4 p = reinterpret_cast<const int4*>(R.o_cnt)[loc];
            c[0] += (uint32_t)p.x; c[1] += (uint32_t)p.y; c[2] += (uint32_t)p.z; c[3] += (uint32_t)p.w;
#pragma unroll
            for (int b = 0; b < 4; b++) q[b] += (uint64_t)R.o_qs[4 * loc + b];
            mqS += (uint32_t)R.o_mq[loc]; qS += (uint32_t)R.o_q[loc];
            bp += (uint32_t)R.o_bp[loc]; nfc += (uint32_t)R.o_frag[loc];
            n = c[0] + c[1] + c[2] + c[3];
        }
        if (R.exp_flags & 1) { if (c[0] == 0xdeadbeef) R.o_mq[loc] = (int32_t)q[0]; continue; }
        finish_locus(R, w, lane, (int32_t)loc, c, q, mqS, qS, bp, n - nfc, pre_rb, pre_ref, rc_md);
    }
}

}  // namespace pb
