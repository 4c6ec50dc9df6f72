// pb_pileup9.cuh -- the hot kernel: scatter into a shared-memory tile, every lane busy.
//
// k_pileup7 (pb_pileup7.cuh) met every read segment once per CTA tile and let the shared-memory reduction units
// accumulate, but its lanes were mapped to half segments: a lane whose segment is short, clipped by the tile edge or
// not an aligned-base segment at all idles while its neighbours walk their chunks (21 of 32 lanes active on average,
// profiles/pileup7_r1s_raw.csv).  This kernel keeps the tile, the packed counters and the per-base reduction and
// changes the unit of work from the segment to the CHUNK (16 consecutive bases of a batch's base stream, aligned:
// one 16-byte load of quality bytes + one 4-byte load of 2-bit codes):
//
//   grab      a warp takes 16 (8) consecutive descriptors of a batch from the tile's cursor, LANE <-> HALF A DESCRIPTOR:
//             clipping to the tile, addresses, flags -- once per segment, 16 segments per instruction;
//   expand    a warp-wide prefix sum of the chunk counts turns them into a dense list of self-contained 8-byte work
//             items (chunk index, tile address, first / last counted position) in the warp's corner of shared memory;
//   scatter   LANE <-> ITEM: 32 chunks per trip whatever the shape of the segments they came from; per base one PRMT
//             (value), LOP + IMAD (address of the letter's word), one predicate, one shared-memory reduction
//                 A[letter][locus] += 1 << 20 | qual            (12-bit count | 20-bit quality sum);
//             the loads of the next trip are in flight while the reductions of this one issue;
//   minority  PileUp.add's other sums follow from that pair for every read whose (adjMq + 1) equals the tile's
//             reference value `dom` (pb_pileup7.cuh has the algebra).  Segments with another value are queued per
//             warp and, 32 at a time, go through the same expand / scatter machinery once more for their Bq / C
//             terms only -- dense as well, so a BAM whose mapping qualities are all over the place degrades
//             gracefully instead of falling off a divergent branch;
//   fold      12-bit counts: after 4064 descriptors the tile is folded into the output planes (deep pile-ups only);
//   epilogue  finish_locus() per locus: sparse merge, BaseCall, pass-1 classification, one write per plane.
//
// Integer reductions commute, so the result is bit-identical to the sequential walk of the reference
// (PileUpRegion.scala:184-193 -> PileUp.scala:75-84).
#pragma once
#include <type_traits>
#include "pb_pileup7.cuh"

namespace pb {

static constexpr int P9_WARPS = 16;
static constexpr int P9_TILE = 2048;                // loci per CTA
static constexpr int P9_PASS_DESC = 4064;           // descriptors per pass <= 4095 (12-bit count)
static constexpr int P9_RING = 96;                  // work items per ring: < 32 pending + 32 lanes x <= 2 new ones

template <int T>
struct __align__(16) Tile9 {
    uint32_t pad[16];            // items of a segment's first chunk address up to 15 words before A[0][0]
    uint32_t A[4][T];            // count << 20 | sum of quals, per letter
    int32_t Bq[4][T];            // sum of qual * (mq1 - dom)
    int32_t C[T];                // sum of (mq1 - dom)
    uint32_t X[T];               // badPair << 16 | counted bases outside fragCoverage
    uint2 ringF[P9_WARPS][P9_RING];       // per warp: chunk work items that need no mask
    uint2 ringE[P9_WARPS][P9_RING];       // per warp: chunk work items for the masked path
    uint2 slow[P9_WARPS][32];             // per warp: descriptor indices of queued segments (16 minority + 16 general)
    const Seg* seg[PB_MAXB]; const uint8_t* quals[PB_MAXB]; const uint8_t* bases2[PB_MAXB];
    uint32_t grab0[PB_MAXB + 1]; // first flat grab index of every batch
    uint32_t slo[PB_MAXB], nseg[PB_MAXB], nf[PB_MAXB];
    uint32_t dom;                // 0 = not chosen yet ((adjMq + 1) >= 1 always)
    uint32_t next;               // next flat grab index to hand out
    uint32_t glog;               // log2 of the descriptors per grab (3..4)
    int32_t read_count, min_depth;   // the region's scalars (k_fold), fetched during set-up for the epilogue
};

// 16 bases of one chunk, every one of them counted and carrying a quality: PRMT (1 << 20 | q), LOP + IMAD (address of the
// letter's word), RED -- no predicate.  NF: the batch is outside fragCoverage (second reduction into X).
template <bool NF, int T>
__device__ __forceinline__ void scatter16(const uint4 Q, uint32_t cw, uint32_t sa) {
    constexpr uint32_t OFF_X = 36u * T;
    constexpr int LOG = T == 512 ? 11 : T == 1024 ? 12 : 13;
    const uint32_t cwm = cw >> 14, cwh = cw >> 28;
#pragma unroll
    for (int b = 0; b < 16; b++) {
        const uint32_t Qw = b < 4 ? Q.x : b < 8 ? Q.y : b < 12 ? Q.z : Q.w;
        const uint32_t val = __byte_perm(Qw, 0x00100000u, 0x7650 | (b & 3));
        const int pos = b < 7 ? 2 * b : b < 14 ? 2 * (b - 7) : 2 * (b - 14);
        const uint32_t code = (b < 7 ? cw : b < 14 ? cwm : cwh) & (3u << pos);
        red_shared_add(code * (1u << (LOG - pos)) + sa + 4u * b, val);
        if (NF) red_shared_add(sa + 4u * b + OFF_X, 1u);
    }
}

// Which of the 16 bases of a chunk are counted: positions [lo, hi), quality byte without the 0x80 mark, >= minQual.
template <bool MINQ>
__device__ __forceinline__ uint32_t chunk_mask9(const uint4 Q, uint32_t lo, uint32_t hi, uint32_t minq_add, uint32_t nohq_pass) {
    uint32_t v0 = ~Q.x >> 7, v1 = ~Q.y >> 7, v2 = ~Q.z >> 7, v3 = ~Q.w >> 7;
    if (MINQ) {                                                  // reads without qualities: default_qual decides
        v0 &= (((Q.x & 0x7F7F7F7Fu) + minq_add) >> 7) | nohq_pass; v1 &= (((Q.y & 0x7F7F7F7Fu) + minq_add) >> 7) | nohq_pass;
        v2 &= (((Q.z & 0x7F7F7F7Fu) + minq_add) >> 7) | nohq_pass; v3 &= (((Q.w & 0x7F7F7F7Fu) + minq_add) >> 7) | nohq_pass;
    }
    const uint32_t m0 = ((v0 & 0x01010101u) * 0x01020408u) >> 24, m1 = ((v1 & 0x01010101u) * 0x01020408u) >> 24;
    const uint32_t m2 = ((v2 & 0x01010101u) * 0x01020408u) >> 24, m3 = ((v3 & 0x01010101u) * 0x01020408u) >> 24;
    return ((1u << hi) - 1u) & ~((1u << lo) - 1u) & ((m0 & 15u) | ((m1 & 15u) << 4) | ((m2 & 15u) << 8) | ((m3 & 15u) << 12));
}

template <bool MINQ, int T>
__global__ void __launch_bounds__(P9_WARPS * 32, 2) k_pileup9(const RegionDev R, const PileBatches PB) {
    extern __shared__ __align__(16) uint8_t smem_raw9[];
    Tile9<T>& S = *reinterpret_cast<Tile9<T>*>(smem_raw9);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t t0 = (int32_t)blockIdx.x * T;
    const int n_batches = PB.n;
    const int min_qual = R.cfg.min_qual;
    const uint32_t defq = (uint32_t)R.cfg.default_qual;
    const uint32_t minq_add = (uint32_t)(0x80 - (min_qual > 128 ? 128 : min_qual)) * 0x01010101u;
    constexpr uint32_t OFF_X = 36u * T;                         // byte offset of X[.] from A[0][.]
    static_assert(T == 1024 || T == 2048, "tile size");

    // ---- tile set-up: zero the counters, candidate descriptor range of every batch, grab size ----
    {
        uint4* z = reinterpret_cast<uint4*>(&S.A[0][0]);
        for (int i = tid; i < 10 * T / 4; i += P9_WARPS * 32) z[i] = make_uint4(0, 0, 0, 0);
        if (tid == 0) { S.dom = 0; S.next = 0; S.read_count = R.sc->read_count; S.min_depth = R.sc->min_depth; }
        if (warp == 0) {
            uint32_t my_slo = 0, my_nseg = 0;
            if (lane < n_batches) {
                const PileBatch& Bl = PB.b[lane];
                S.seg[lane] = Bl.seg; S.quals[lane] = Bl.quals; S.bases2[lane] = Bl.bases2; S.nf[lane] = (Bl.flags & 1) ? 0u : 1u;
                if (Bl.flags & 2) {
                    const int64_t x = (int64_t)t0 - Bl.reach[0] + 1;
                    const int64_t y = (int64_t)t0 + T + Bl.reach[1];
                    int64_t khi = (y + 31) >> 5; if (khi > R.n_win) khi = R.n_win;
                    my_slo = x <= 0 ? 0u : Bl.win_first[x >> 5];
                    const uint32_t shi = (y > ((int64_t)R.n_win << 5)) ? Bl.n_cigar : Bl.win_first[khi];
                    my_nseg = shi > my_slo ? shi - my_slo : 0u;
                }
            }
            // descriptors per grab (two lanes each): 16, or 8 on a shallow tile so that every warp still gets a few grabs
            // and the warps reach the end of the scatter together
            uint32_t tot = my_nseg;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
            const uint32_t glog = tot >= 4u * 16u * P9_WARPS ? 4u : 3u;
            uint32_t ng = (my_nseg + (1u << glog) - 1u) >> glog, pre = ng;      // inclusive scan of the grab counts
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, pre, o); if (lane >= o) pre += v; }
            if (lane < PB_MAXB) { S.slo[lane] = my_slo; S.nseg[lane] = my_nseg; S.grab0[lane] = pre - ng; }
            if (lane == PB_MAXB - 1) S.grab0[PB_MAXB] = pre;
            if (lane == 0) S.glog = glog;
        }
    }
    __syncthreads();
    const uint32_t total_grabs = S.grab0[PB_MAXB];
    const uint32_t glog = S.glog, G = 1u << glog;
    bool folded = false;

    // fold the 12/20-bit tile into the output planes (used as 32/64-bit accumulators) and clear it
    auto fold = [&]() {
        const uint32_t dom = S.dom;
        for (int l = tid; l < T; l += P9_WARPS * 32) {
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

    const uint32_t sA = smem_u32(&S.A[0][0]);
    uint2* const ringF = S.ringF[warp];                          // {k, sa}: interior chunks, every base in range
    uint2* const ringE = S.ringE[warp];                          // {k | lo << 28, sa | hi << 18}: edge chunks, chunks with uncountable bases
    uint32_t* const queue = reinterpret_cast<uint32_t*>(S.slow[warp]);     // [0, 32) minority segments, [32, 64) general segments
    uint32_t f_head = 0, f_n = 0, e_head = 0, e_n = 0;           // warp-uniform ring cursors
    uint32_t g_n = 0; bool g_dmq = false;                        // transient list of the general path (ring F's storage, ring F empty)
    uint32_t dom_r = 0;                                          // register copy of S.dom once it is known
    uint32_t slow_n = 0, gen_n = 0;                              // warp-uniform: queued minority / general segments
    const int h = lane & 1;                                      // LANE <-> HALF of a descriptor's chunks
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint4* qp = nullptr; const uint32_t* cp = nullptr;     // base streams of the batch the rings hold chunks of
    bool nf = false; int ring_batch = -1;
    auto wrap = [](uint32_t x) { return x >= (uint32_t)P9_RING ? x - (uint32_t)P9_RING : x; };
    // the half segment this lane is handing out chunks of: d_c chunks from chunk d_k0 on, tile address d_sa0 of chunk 0,
    // first counted position d_lo of chunk 0, end position d_hi in chunk d_c - 1; chunks [cur, cur + rem) still to hand out
    uint32_t d_c = 0, d_k0 = 0, d_sa0 = 0, d_lo = 0, d_hi = 16, cur = 0, rem = 0; bool d_hasq = true; int32_t d_dmq = 0;

    // ---- scatter: passes of <= 4064 descriptors.  One loop; every trip of it first PRODUCES work items (a grab's
    // edge chunks, a round of interior chunks, a round of a queued segment's chunks) and then PUMPS the rings: each
    // reduction body exists once in the code (inlined copies at every call site had made the kernel outgrow the
    // instruction cache: 21 k SASS instructions, "no instruction" the top stall reason) ----
    const uint32_t PASS_GRABS = (uint32_t)P9_PASS_DESC >> glog;
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
            const uint32_t di = ((g - S.grab0[b]) << glog) + (uint32_t)(lane >> 1);     // two lanes per descriptor
            sidx_nx = S.slo[b] + di;
            if ((uint32_t)(lane >> 1) < G && di < S.nseg[b]) seg_nx = PB.b[b].seg[sidx_nx];
        };
        if (!(R.exp_flags & 2)) fetch();
        int state = 0;                                           // 0: decide, 1: a round of ring F items, 2: a round of general-path items
        bool need_flush = false, pass_end = false;
        for (;;) {
            bool flush = false;
            if (state == 1) {
                // every lane hands out at most two interior chunks, so that ring F never holds more than 31 + 64 items
                const uint32_t take = min(rem, 2u);
                uint32_t pre = take;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, pre, o); if (lane >= o) pre += v; }
                const uint32_t at = wrap(f_head + f_n) + pre - take;
                if (take) ringF[wrap(at)] = make_uint2(d_k0 + cur, d_sa0 + 64u * cur);
                if (take > 1u) ringF[wrap(at + 1u)] = make_uint2(d_k0 + cur + 1u, d_sa0 + 64u * cur + 64u);
                f_n += __shfl_sync(FULL, pre, 31); cur += take; rem -= take;
                if (!__any_sync(FULL, rem != 0u)) state = 0;
                __syncwarp();
            } else if (state == 2) {
                // general path: at most three chunks per lane into the transient list {k | lo << 28, sa | hi << 18 | lane << 24}
                const uint32_t take = min(rem, (uint32_t)(P9_RING / 32));
                uint32_t pre = take;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, pre, o); if (lane >= o) pre += v; }
                uint2* at = ringF + (pre - take);
                for (uint32_t j = 0; j < take; j++) {
                    const uint32_t ord = cur + j;
                    *at++ = make_uint2((d_k0 + ord) | (ord == 0u ? d_lo << 28 : 0u),
                                       (d_sa0 + 64u * ord) | ((ord + 1u == d_c ? d_hi : 16u) << 18) | ((uint32_t)lane << 24));
                }
                g_n = __shfl_sync(FULL, pre, 31); cur += take; rem -= take;
                if (!__any_sync(FULL, rem != 0u)) state = 0;
                __syncwarp();
            } else {
                const bool drain_slow = slow_n >= 16u || (need_flush && slow_n);
                const bool drain_gen = !drain_slow && (gen_n >= 16u || (need_flush && gen_n));
                if (need_flush && (f_n | e_n)) flush = true;                       // batch switch / end of the pass: empty the rings
                else if ((drain_slow || drain_gen) && f_n) flush = true;           // the general path borrows ring F's storage
                else if (!drain_slow && !drain_gen && need_flush) {
                    if (pass_end) break;
                    ring_batch = -1; need_flush = false;                           // (the branch below installs the next batch)
                } else if (!drain_slow && !drain_gen && g_nx >= p1) {
                    if (ring_batch < 0) break;
                    need_flush = true; pass_end = true;
                } else if (!drain_slow && !drain_gen && b_nx != ring_batch && ring_batch >= 0) need_flush = true;
                else {
                    // a segment half per lane: from the queues (lane pair <-> queued segment) or from the next grab
                    const bool drain = drain_slow || drain_gen;
                    Seg mine = {0, 0, 0, 0}; uint32_t sidx = 0; int b_cur = ring_batch;
                    if (drain) {
                        uint32_t& qn = drain_slow ? slow_n : gen_n;
                        const uint32_t cnt = min(qn, 16u);
                        if ((uint32_t)(lane >> 1) < cnt) mine = S.seg[ring_batch][queue[(drain_slow ? 0u : 32u) + qn - cnt + (uint32_t)(lane >> 1)]];
                        qn -= cnt;
                    } else {
                        mine = seg_nx; sidx = sidx_nx; b_cur = b_nx;
                        fetch();
                        if (ring_batch < 0) {                    // (warp-uniform) the rings hold chunks of one batch at a time
                            ring_batch = b_cur;
                            qp = reinterpret_cast<const uint4*>(S.quals[b_cur]); cp = reinterpret_cast<const uint32_t*>(S.bases2[b_cur]);
                            nf = S.nf[b_cur] != 0u;              // this batch is outside fragCoverage
                        }
                    }
                    // my half of the descriptor clipped to the tile
                    const int32_t cA = mine.loc0 > t0 ? mine.loc0 : t0;
                    const int32_t cBx = mine.loc0 + mine.len < t0 + T ? mine.loc0 + mine.len : t0 + T;
                    const int32_t n = mine.len > 0 ? (cBx > cA ? cBx - cA : 0) : 0;
                    const int32_t col = cA - t0;
                    const bool valid = mine.w & SEG_VALID;
                    const bool live = n > 0 && valid;
                    {
                        const uint32_t src = mine.src + (uint32_t)(cA - mine.loc0);
                        const uint32_t first = src & 15u, lastrel = first + (uint32_t)n - 1u;
                        const uint32_t c = live ? (lastrel >> 4) + 1u : 0u;        // chunks of the whole segment
                        const uint32_t c0 = (c + 1u) >> 1, c1 = c - c0;            // first half / second half
                        const uint32_t off = h ? c0 : 0u;
                        d_c = h ? c1 : c0;
                        d_k0 = (src >> 4) + off;
                        d_sa0 = sA + 4u * (uint32_t)(col - (int32_t)first) + 64u * off;
                        d_lo = h ? 0u : first;
                        d_hi = (h ? c1 > 0u : c1 == 0u) ? (lastrel & 15u) + 1u : 16u;      // the half that holds the segment's last chunk
                        d_hasq = mine.w & SEG_HASQ;
                        d_dmq = (int32_t)(mine.w & 0xFFFFu) - (int32_t)dom_r;
                    }
                    cur = 0; rem = 0;
                    if (drain) {
                        g_dmq = drain_slow;
                        rem = d_c;
                        if (__any_sync(FULL, rem != 0u)) state = 2;
                    } else {
                        unsigned badm = __ballot_sync(FULL, n > 0 && !valid && h == 0);
                        while (badm) {                            // PileUpRegion.scala:45: badPair++ on every locus, lane <-> locus
                            const int j = __ffs(badm) - 1; badm &= badm - 1;
                            const int32_t bn = __shfl_sync(FULL, n, j), bcol = __shfl_sync(FULL, col, j);
                            for (int i = lane; i < bn; i += 32) red_shared_add(sA + OFF_X + 4u * (uint32_t)(bcol + i), 0x10000u);
                        }
                        const unsigned livem = __ballot_sync(FULL, live);
                        if (livem) {
                            const uint32_t mq1 = mine.w & 0xFFFFu;
                            if (dom_r == 0) {                     // the tile's reference (adjMq + 1): first one met
                                const uint32_t first = __shfl_sync(FULL, mq1, __ffs(livem) - 1);
                                uint32_t old = 0;
                                if (lane == 0) old = atomicCAS(&S.dom, 0u, first);
                                old = __shfl_sync(FULL, old, 0);
                                dom_r = old ? old : first;
                            }
                            // a segment with another mapping quality (~5 % of the reads) is queued for its Bq / C terms, a
                            // read without qualities for the general path (fewer than 16 are queued at this point)
                            const unsigned qm = __ballot_sync(FULL, live && mq1 != dom_r && h == 0);
                            const unsigned gm = __ballot_sync(FULL, live && !d_hasq && h == 0);
                            if ((qm >> lane) & 1u) queue[slow_n + (uint32_t)__popc(qm & lt_mask)] = sidx;
                            if ((gm >> lane) & 1u) queue[32u + gen_n + (uint32_t)__popc(gm & lt_mask)] = sidx;
                            slow_n += (uint32_t)__popc(qm); gen_n += (uint32_t)__popc(gm);
                            // chunks of reads with qualities: my half's first / last chunk goes to ring E when it is partial
                            // (never both unless d_c == 1), the rest to ring F
                            const bool ring = live && d_hasq && d_c > 0u;
                            const bool e_first = ring && d_lo != 0u, e_last = ring && d_hi != 16u;
                            const unsigned em = __ballot_sync(FULL, e_first || e_last);
                            if (e_first || e_last) {
                                const uint32_t ord = e_first ? 0u : d_c - 1u;
                                ringE[wrap(wrap(e_head + e_n) + (uint32_t)__popc(em & lt_mask))] =
                                    make_uint2((d_k0 + ord) | ((e_first ? d_lo : 0u) << 28), (d_sa0 + 64u * ord) | ((e_last ? d_hi : 16u) << 18));
                            }
                            e_n += (uint32_t)__popc(em);
                            if (ring) { cur = e_first ? 1u : 0u; rem = d_c - cur - ((e_last && !e_first) ? 1u : 0u); }
                            if (__any_sync(FULL, rem != 0u)) state = 1;
                            __syncwarp();
                        }
                    }
                }
            }

            // ---- pump ------------------------------------------------------------------------------------
            if (g_n) {          // general path (ring F is empty, its storage holds the transient list): all of it, masked
                for (uint32_t base = 0; base < g_n; base += 32u) {
                    const uint32_t i = base + (uint32_t)lane;
                    uint2 it = make_uint2(0u, 0u); uint4 Q = make_uint4(0, 0, 0, 0); uint32_t cw = 0;
                    if (i < g_n) { it = ringF[i]; const uint32_t k = it.x & 0x0FFFFFFFu; Q = qp[k]; cw = cp[k]; }
                    const int src_lane = (int)((it.y >> 24) & 31u);
                    const bool hq = __shfl_sync(FULL, (int)d_hasq, src_lane) != 0;
                    const int32_t dq = __shfl_sync(FULL, d_dmq, src_lane);
                    const uint32_t nohq_pass = (!hq && (int)defq >= min_qual) ? 0x01010101u : 0u;
                    const uint32_t okm = chunk_mask9<MINQ>(Q, it.x >> 28, (it.y >> 18) & 31u, minq_add, nohq_pass);   // hi == 0 for idle lanes
                    const uint32_t sa = it.y & 0x3FFFFu;
                    if (g_dmq) scatter_chunk_dmq<T>(Q, cw, okm, sa, hq ? 0x7Fu : 0u, hq ? 0u : defq, dq);
                    else if (nf) scatter_chunk<true, false, T>(Q, cw, okm, sa, hq ? 0x7Fu : 0u, (1u << 20) | (hq ? 0u : defq));
                    else scatter_chunk<false, false, T>(Q, cw, okm, sa, hq ? 0x7Fu : 0u, (1u << 20) | (hq ? 0u : defq));
                }
                g_n = 0;
                __syncwarp();
            }
            for (;;) {
                // ring E first (it is below 32 entries afterwards, and two trips of ring F add at most 64): the masked path
                while (e_n >= 32u || (flush && e_n)) {
                    const uint32_t cnt = min(e_n, 32u);
                    uint2 it = make_uint2(0u, 0u); uint4 Q = make_uint4(0, 0, 0, 0); uint32_t cw = 0;
                    if ((uint32_t)lane < cnt) { it = ringE[wrap(e_head + (uint32_t)lane)]; const uint32_t k = it.x & 0x0FFFFFFFu; Q = qp[k]; cw = cp[k]; }
                    const uint32_t okm = chunk_mask9<MINQ>(Q, it.x >> 28, (it.y >> 18) & 31u, minq_add, 0u);          // hi == 0 for idle lanes
                    if (nf) scatter_chunk<true, true, T>(Q, cw, okm, it.y & 0x3FFFFu, 0x7Fu, 1u << 20);
                    else scatter_chunk<false, true, T>(Q, cw, okm, it.y & 0x3FFFFu, 0x7Fu, 1u << 20);
                    e_head = wrap(e_head + cnt); e_n -= cnt;
                }
                // ring F: up to two trips, both loads issued before the first trip's reductions; 16 unconditional
                // reductions per lane; a chunk with a marked byte moves to ring E
                const uint32_t todo = min(flush ? f_n : (f_n & ~31u), 64u);
                if (todo == 0u) break;
                const uint32_t cntA = min(todo, 32u), cntB = todo - cntA;
                uint2 itA = make_uint2(0u, 0u), itB = itA; uint4 QA = make_uint4(0, 0, 0, 0), QB = QA; uint32_t cwA = 0, cwB = 0;
                if ((uint32_t)lane < cntA) { itA = ringF[wrap(f_head + (uint32_t)lane)]; QA = qp[itA.x]; cwA = cp[itA.x]; }
                if ((uint32_t)lane < cntB) { itB = ringF[wrap(f_head + 32u + (uint32_t)lane)]; QB = qp[itB.x]; cwB = cp[itB.x]; }
#pragma unroll 1
                for (int t = 0; t < 2; t++) {
                    const uint32_t cnt = t ? cntB : cntA;
                    if (cnt == 0u) break;
                    const uint2 it = t ? itB : itA; const uint4 Q = t ? QB : QA; const uint32_t cw = t ? cwB : cwA;
                    const bool act = (uint32_t)lane < cnt;
                    uint32_t marked = (Q.x | Q.y | Q.z | Q.w) & 0x80808080u;                // uncountable bases (PileUp.scala:46-52)
                    if (MINQ) marked |= ~(((Q.x & 0x7F7F7F7Fu) + minq_add) & ((Q.y & 0x7F7F7F7Fu) + minq_add) &
                                          ((Q.z & 0x7F7F7F7Fu) + minq_add) & ((Q.w & 0x7F7F7F7Fu) + minq_add)) & 0x80808080u;   // below minQual (:77)
                    const bool dirty = act && marked != 0u;
                    if (act && !dirty) { if (nf) scatter16<true, T>(Q, cw, it.y); else scatter16<false, T>(Q, cw, it.y); }
                    const unsigned dm = __ballot_sync(FULL, dirty);
                    if (dm) {
                        if (dirty) ringE[wrap(wrap(e_head + e_n) + (uint32_t)__popc(dm & lt_mask))] = make_uint2(it.x, it.y | (16u << 18));
                        e_n += (uint32_t)__popc(dm);
                    }
                }
                f_head = wrap(f_head + todo); f_n -= todo;
                __syncwarp();
            }
        }
        __syncthreads();
    }

    // ---- epilogue: warp per 32-locus window of the tile ----
    const uint32_t dom = S.dom;
    const int2 rc_md = make_int2(S.read_count, S.min_depth);
    for (int wl = warp; wl < T / 32; wl += P9_WARPS) {
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
