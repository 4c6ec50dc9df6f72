// pb_pileup4.cuh -- the hot kernel, fourth generation: TMA-fed warp-specialised CTA tile.
//
// A CTA owns a tile of 7 windows x 32 loci.  Reads are stored back to back in a batch, so the bytes
// that 32 consecutive segment descriptors need form ONE contiguous range of the quality array (and of
// the 2-bit code array).  Warp 0, the PRODUCER, therefore does almost nothing per chunk: each lane
// clips its (prefetched) descriptor to the tile, four warp reductions give the byte range and the
// column span, and one elected thread issues two TMA bulk copies (cp.async.bulk, SASS UBLKCP) that
// land the range in a 4-deep shared-memory ring and complete the chunk's full-mbarrier by byte
// count.  A chunk whose range would not fit (long reads) is shortened to 16, 8, ... 1 descriptors.
//
// Warps 1..7 are CONSUMERS, one per window.  Per chunk a consumer skips it outright if its column
// span misses the window; otherwise lane j derives the geometry of row j (in-window column mask,
// shared-memory address of window column 0, window-aligned 2-bit codes), the warp votes the
// dominant (adjMq + 1) of its overlapping rows, and the rows are accumulated with the byte-SIMD
// scheme described in pb_pileup2.cuh: 4 rows per step (lane group <-> row), 4 loci per lane, packed
// 8-bit counts / 16-bit quality sums for bases that equal the locus' primary letter.  Exact slow
// paths cover the rest (other letters, other mapping qualities, reads without qualities, invalid
// reads and soft clips); packed registers are flushed by a shuffle reduction, never by atomics.
// The epilogue is finish_locus(): sparse merge + BaseCall + pass-1 classification + one write per plane.
#pragma once
#include "pb_pileup3.cuh"

namespace pb {

static constexpr int P4_CW = 7;
static constexpr int P4_TILE = P4_CW * 32;
static constexpr int P4_NS = 4;
static constexpr int P4_QCAP = 6144;             // quality bytes per chunk (32 x 152-byte reads = 4864 + alignment)
static constexpr int P4_CCAP = P4_QCAP / 4 + 64;
static constexpr int P4_PAD = 256;               // masked loads may fall this far outside a buffer

struct __align__(16) Chunk4 { uint32_t type, frag; int32_t lo, hi; uint32_t qbase, cbase, pad0, pad1; };

struct __align__(128) Slot4 {
    uint8_t front[P4_PAD];
    uint8_t qbuf[P4_QCAP];
    uint8_t mid[P4_PAD];
    uint8_t cbuf[P4_CCAP];
    uint8_t back[P4_PAD - 64];
    Seg seg[32];                                 // clipped to the tile; loc0 = tile column; len = 0: nothing
    Chunk4 ch;
};

// Per-window accumulation table.  The running sums are 32-bit so that every update is a native,
// fire-and-forget shared-memory atomic (64-bit shared atomics are CAS loops); a quality sum grows by
// at most 127 * 256 per row, so it is folded into the 64-bit table every SPILL_ROWS rows.
struct __align__(16) Warp4 {
    unsigned long long tqs64[32][4];
    uint32_t tqs[32][4];
    uint32_t tcnt[32][4];
    uint32_t tmq[32], tq[32], tbp[32];
    unsigned long long codes[32];
};
static constexpr uint32_t P4_SPILL_ROWS = 120000;   // 120000 * 32512 < 2^32

struct __align__(128) Smem4 {
    Slot4 slot[P4_NS];
    unsigned long long full[P4_NS], empty[P4_NS];
    Warp4 warp[P4_CW];
};

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <bool MINQ>
__global__ void __launch_bounds__((P4_CW + 1) * 32) k_pileup4(RegionDev R, const DevBatch* __restrict__ batches, int n_batches) {
    extern __shared__ __align__(128) uint8_t smem_raw4[];
    Smem4& S = *reinterpret_cast<Smem4*>(smem_raw4);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t t0 = (int32_t)blockIdx.x * P4_TILE;
    int* err = &R.sc->error;
    // debug timeline (off unless PB_DEBUG_TILE selects this tile)
    const bool dbg_on = R.dbg != nullptr && (int32_t)blockIdx.x == R.dbg_tile && lane == 0;
    int dbg_n = 0;
    auto mark = [&](int tag, uint32_t n) {
        if (dbg_on && dbg_n < 255) { R.dbg[(warp * 256 + dbg_n) * 2] = ((long long)tag << 32) | n; R.dbg[(warp * 256 + dbg_n) * 2 + 1] = clock64(); dbg_n++; }
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < P4_NS; s++) { mbar_init(&S.full[s], 32); mbar_init(&S.empty[s], P4_CW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        // ================================ PRODUCER ================================
        uint32_t n = 0;
        auto acquire = [&](uint32_t slot) -> bool {
            if (n >= P4_NS) return mbar_wait(&S.empty[slot], ((n / P4_NS) + 1) & 1, err);
            return true;
        };
        auto marker = [&](uint32_t type, uint32_t frag) -> bool {
            const uint32_t slot = n % P4_NS;
            if (!acquire(slot)) return false;
            if (lane == 0) { Chunk4 ch = {}; ch.type = type; ch.frag = frag; S.slot[slot].ch = ch; }
            mbar_arrive(&S.full[slot]);
            n++;
            return true;
        };
        bool alive = true;
        for (int bb = 0; bb < n_batches && alive; bb += 32) {
            uint32_t my_slo = 0, my_shi = 0;          // candidate segment range per batch, lane <-> batch
            if (bb + lane < n_batches) {
                const DevBatch& Bl = batches[bb + lane];
                if (Bl.n_reads) {
                    const int64_t x = (int64_t)t0 - Bl.reach[0] + 1;
                    const int64_t y = (int64_t)t0 + P4_TILE + Bl.reach[1];
                    int64_t khi = (y + 31) >> 5; if (khi > R.n_win) khi = R.n_win;
                    my_slo = x <= 0 ? 0u : Bl.win_first[x >> 5];
                    my_shi = (y > ((int64_t)R.n_win << 5)) ? (uint32_t)Bl.n_cigar : Bl.win_first[khi];
                }
            }
            const int nbb = n_batches - bb < 32 ? n_batches - bb : 32;
            for (int bi = 0; bi < nbb && alive; bi++) {
                const Seg* __restrict__ segs = batches[bb + bi].seg;
                const uint8_t* __restrict__ gquals = batches[bb + bi].quals;
                const uint8_t* __restrict__ gbases = batches[bb + bi].bases2;
                const uint32_t bfrag = (uint32_t)batches[bb + bi].frag;
                const bool bempty = batches[bb + bi].n_reads == 0;
                const uint32_t slo = __shfl_sync(FULL, my_slo, bi), shi = __shfl_sync(FULL, my_shi, bi);
                if (bempty) continue;
                const Seg none = {0, 0, 0, 0};
                Seg next = none;
                if (slo + lane < shi) next = segs[slo + lane];
                uint32_t sb = slo;
                while (sb < shi && alive) {
                    mark(0, n);
                    Seg mine = next;
                    next = none;
                    if (sb + 32 + lane < shi) next = segs[sb + 32 + lane];       // prefetch the next chunk's descriptors
                    const uint32_t slot = n % P4_NS;
                    if (mine.len == 0x7fffffff) mark(99, n);                    // (forces the descriptor load to complete before mark 1)
                    mark(1, n);
                    alive = acquire(slot);
                    if (!alive) break;
                    mark(2, n);
                    Slot4& SL = S.slot[slot];
                    // clip to the tile
                    const int a = mine.loc0 > t0 ? mine.loc0 : t0;
                    const int e = mine.loc0 + mine.len < t0 + P4_TILE ? mine.loc0 + mine.len : t0 + P4_TILE;
                    const bool ovt = mine.len > 0 && e > a;
                    const bool needs = ovt && (mine.w & SEG_VALID);           // invalid rows carry no bytes
                    const uint32_t csrc = mine.src + (uint32_t)(a - mine.loc0);
                    uint32_t take = 32;
                    uint32_t qlo, qhi;
                    for (;;) {
                        const bool in = (uint32_t)lane < take;
                        qlo = __reduce_min_sync(FULL, (needs && in) ? csrc : 0xFFFFFFFFu);
                        qhi = __reduce_max_sync(FULL, (needs && in) ? csrc + (uint32_t)(e - a) : 0u);
                        if (qhi == 0 || ((qhi + 15) & ~15u) - (qlo & ~15u) <= (uint32_t)P4_QCAP || take == 1) break;
                        take >>= 1;                                              // long reads: shorten the chunk
                    }
                    const bool in = (uint32_t)lane < take;
                    const int32_t lo = __reduce_min_sync(FULL, (ovt && in) ? a - t0 : 0x7fffffff);
                    const int32_t hi = __reduce_max_sync(FULL, (ovt && in) ? e - t0 : 0);
                    Seg out = none;
                    if (ovt && in) { out.loc0 = a - t0; out.len = e - a; out.src = csrc; out.w = mine.w; }
                    SL.seg[lane] = out;
                    if (lane == 0) {
                        Chunk4 ch = {}; ch.type = CH_DATA; ch.lo = lo; ch.hi = hi;
                        if (qhi) {
                            const uint32_t qbase = qlo & ~15u, qbytes = ((qhi + 15) & ~15u) - qbase;
                            const uint32_t cbase = (qbase >> 2) & ~15u, cbytes = ((((qhi + 3) >> 2) + 15) & ~15u) - cbase;
                            ch.qbase = qbase; ch.cbase = cbase;
                            SL.ch = ch;                                            // before the (releasing) arrive
                            asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}"
                                         ::"r"(smem_u32(&S.full[slot])), "r"(qbytes + cbytes) : "memory");
                            bulk_g2s(SL.qbuf, gquals + qbase, qbytes, &S.full[slot]);
                            bulk_g2s(SL.cbuf, gbases + cbase, cbytes, &S.full[slot]);
                        } else {
                            SL.ch = ch;
                            mbar_arrive(&S.full[slot]);
                        }
                    } else mbar_arrive(&S.full[slot]);
                    mark(3, n);
                    n++;
                    sb += take;
                    if (take != 32) { next = none; if (sb + lane < shi) next = segs[sb + lane]; }
                }
                if (!alive) break;
                alive = marker(CH_EOB, bfrag);       // consumers flush + fragCoverage snapshot (GenomeRegion.scala:290-298)
            }
        }
        if (alive) marker(CH_EOT, 0);
        return;
    }

    // ================================== CONSUMERS ==================================
    const int cw = warp - 1;
    const int32_t wc = cw * 32;
    const int64_t w = (int64_t)blockIdx.x * P4_CW + cw;
    const bool active = w < R.n_win;
    const int32_t w0 = t0 + wc;
    Warp4& W = S.warp[cw];
    const int g = lane >> 3, k = lane & 7, kk = k << 2;
    const int min_qual = R.cfg.min_qual;
    const uint32_t defq = (uint32_t)R.cfg.default_qual;
    const uint32_t minq_add = (uint32_t)(0x80 - (min_qual > 128 ? 128 : min_qual)) * 0x01010101u;

#pragma unroll
    for (int b = 0; b < 4; b++) { W.tqs64[lane][b] = 0; W.tqs[lane][b] = 0; W.tcnt[lane][b] = 0; }
    W.tmq[lane] = 0; W.tq[lane] = 0; W.tbp[lane] = 0;
    const uint32_t pre_rb = active ? R.rare_bits[w] : 0u;
    const uint8_t pre_ref = (active && (int64_t)w0 + lane < R.size) ? ref_at(R, (int64_t)R.start + w0 + lane) : (uint8_t)'N';
    // primary letters of my 4 loci = reference bases (lane <-> locus byte fetched above, regrouped by shuffle)
    uint32_t P8 = 0;
    {
        const int rc = ref_class(pre_ref);
        const uint32_t code = (uint32_t)(rc < 4 ? rc : 0);
#pragma unroll
        for (int j = 0; j < 4; j++) P8 |= __shfl_sync(FULL, code, kk + j) << (2 * j);
    }
    __syncwarp();

    uint32_t cnt4 = 0, QLo = 0, QHi = 0, cur_mq = 0, nrows = 0, fragN = 0, nprev = 0, rows_total = 0;

    auto spill = [&]() {       // fold the 32-bit quality sums into the 64-bit table (lane <-> locus)
        __syncwarp();
#pragma unroll
        for (int b = 0; b < 4; b++) { W.tqs64[lane][b] += W.tqs[lane][b]; W.tqs[lane][b] = 0; }
        __syncwarp();
        rows_total = 0;
    };
    auto flush = [&]() {       // warp-uniform: reduce the 4 row groups with shuffles, then lane (g,k) owns locus 4k+g
        if (__any_sync(FULL, cnt4 != 0)) {
            uint32_t c02 = cnt4 & 0x00FF00FFu, c13 = (cnt4 >> 8) & 0x00FF00FFu;
            uint32_t q0 = QLo & 0xFFFF, q2 = QLo >> 16, q1 = QHi & 0xFFFF, q3 = QHi >> 16;
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
                c02 += __shfl_xor_sync(FULL, c02, o); c13 += __shfl_xor_sync(FULL, c13, o);
                q0 += __shfl_xor_sync(FULL, q0, o); q1 += __shfl_xor_sync(FULL, q1, o);
                q2 += __shfl_xor_sync(FULL, q2, o); q3 += __shfl_xor_sync(FULL, q3, o);
            }
            const uint32_t cj = g == 0 ? (c02 & 0xFFFF) : g == 1 ? (c13 & 0xFFFF) : g == 2 ? (c02 >> 16) : (c13 >> 16);
            const uint32_t Qj = g == 0 ? q0 : g == 1 ? q1 : g == 2 ? q2 : q3;
            __syncwarp();
            if (cj) {
                const int l = kk + g; const uint32_t letter = (P8 >> (2 * g)) & 3;
                W.tcnt[l][letter] += cj;
                W.tqs[l][letter] += Qj * cur_mq;            // <= 1020 rows * 127 * 256 per flush
                W.tmq[l] += cj * cur_mq;
                W.tq[l] += Qj;
            }
            __syncwarp();
        }
        cnt4 = 0; QLo = 0; QHi = 0; nrows = 0;
    };

    for (uint32_t n = 0;; n++) {
        const uint32_t slot = n % P4_NS;
        mark(4, n);
        if (!mbar_wait(&S.full[slot], (n / P4_NS) & 1, err)) return;
        mark(5, n);
        Slot4& SL = S.slot[slot];
        const Chunk4 ch = SL.ch;
        if (ch.type == CH_EOT) break;
        if (ch.type == CH_EOB) {
            flush();
            const uint32_t nnow = W.tcnt[lane][0] + W.tcnt[lane][1] + W.tcnt[lane][2] + W.tcnt[lane][3];
            if (ch.frag) fragN += nnow - nprev;
            nprev = nnow;
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.empty[slot]);
            continue;
        }
        if (active && ch.lo < wc + 32 && ch.hi > wc) {
            // ---- per-row geometry, lane <-> row ----
            const Seg sg = SL.seg[lane];
            const int x0 = sg.loc0 - wc, x1 = x0 + sg.len;                   // window-relative columns
            const bool ov = sg.len > 0 && x0 < 32 && x1 > 0;
            const bool valid = sg.w & SEG_VALID, hasq = sg.w & SEG_HASQ;
            const uint32_t mq1 = sg.w & 0xFFFF;
            const uint32_t lo = x0 > 0 ? (uint32_t)x0 : 0u, hi = x1 < 32 ? (uint32_t)x1 : 32u;
            const uint32_t colmask = ov ? ((hi == 32 ? 0xFFFFFFFFu : ((1u << hi) - 1)) & ~((1u << lo) - 1)) : 0u;
            const uint32_t iw = sg.src - (uint32_t)x0;                          // base index of window column 0
            // rows that carry no bytes get a harmless address: the fast loop loads (masked) from every row in its range
            const int32_t qaddr = (int32_t)smem_u32(SL.qbuf) + ((ov && valid) ? (int32_t)(iw - ch.qbase) : 0);
            if (ov && valid) {
                const int32_t cbw = 2 * (int32_t)(iw - 4u * ch.cbase);          // bit offset of column 0's code in cbuf
                const uint32_t ca = smem_u32(SL.cbuf) + (uint32_t)((cbw >> 5) * 4);
                uint32_t W0, W1, W2;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(W0) : "r"(ca));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(W1) : "r"(ca + 4));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(W2) : "r"(ca + 8));
                const uint32_t sft = (uint32_t)cbw & 31;
                W.codes[lane] = ((unsigned long long)__funnelshift_r(W1, W2, sft) << 32) | __funnelshift_r(W0, W1, sft);
            }
            // dominant (adjMq + 1) among this window's eligible rows; voted only when some row disagrees
            const bool elig = ov && valid && hasq;
            if (__any_sync(FULL, elig && mq1 != cur_mq)) {
                const unsigned peers = __match_any_sync(FULL, elig ? mq1 : (0x10000u + lane));
                const uint32_t votes = elig ? (((uint32_t)__popc(peers) << 17) | ((mq1 == cur_mq) ? 0x10000u : 0u) | mq1) : 0u;
                const uint32_t best = __reduce_max_sync(FULL, votes);
                if (best && (best & 0xFFFF) != cur_mq) { flush(); cur_mq = best & 0xFFFF; }
            }
            const bool fast = elig && mq1 == cur_mq;
            const unsigned fastm = __ballot_sync(FULL, fast);
            unsigned scalm = __ballot_sync(FULL, ov && !fast);
            rows_total += 32;
            if (rows_total > P4_SPILL_ROWS) { flush(); spill(); }
            const uint32_t cm_fast = fast ? colmask : 0u;
            __syncwarp();
            mark(10, (uint32_t)__popc(scalm));
            // ---- odd rows: lane <-> locus ----
            while (scalm) {
                const int j = __ffs(scalm) - 1; scalm &= scalm - 1;
                const uint32_t cmj = __shfl_sync(FULL, colmask, j);
                const uint32_t swj = __shfl_sync(FULL, sg.w, j);
                const int32_t qb = __shfl_sync(FULL, qaddr, j);
                if ((cmj >> lane) & 1) {
                    if (!(swj & SEG_VALID)) atomicAdd(&W.tbp[lane], 1u);           // PileUpRegion.scala:45
                    else {
                        uint32_t qv;
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(qv) : "r"((uint32_t)(qb + lane)));
                        if (!(qv & 0x80)) {
                            const uint32_t code = (uint32_t)(W.codes[j] >> (2 * lane)) & 3;
                            const uint32_t q = (swj & SEG_HASQ) ? qv : defq;
                            if (!MINQ || (int)q >= min_qual) {
                                const uint32_t m1 = swj & 0xFFFF;
                                atomicAdd(&W.tcnt[lane][code], 1u); atomicAdd(&W.tqs[lane][code], q * m1);
                                atomicAdd(&W.tmq[lane], m1); atomicAdd(&W.tq[lane], q);
                            }
                        }
                    }
                }
            }
            // ---- fast rows: 4 rows per step (one per lane group), 4 loci per lane ----
            mark(11, (uint32_t)__popc(fastm));
            if (fastm) {
                const int r_lo = __ffs(fastm) - 1, r_hi = 32 - __clz(fastm);
                const int iters = (r_hi - r_lo + 3) >> 2;
                mark(12, (uint32_t)iters);
                if (nrows + (uint32_t)iters > 255) flush();
                nrows += (uint32_t)iters;
                // 4 independent rows per lane per trip: all shuffles, then all loads, then the arithmetic,
                // so that the latencies of the four chains overlap (the single chain was ~400 cycles long)
                for (int it0 = 0; it0 < iters; it0 += 4) {
                    uint32_t cm[4], wlo[4], whi[4], sh[4], C8[4];
                    int32_t qb[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int r = r_lo + g + 4 * (it0 + u);
                        cm[u] = __shfl_sync(FULL, cm_fast, r & 31);
                        qb[u] = __shfl_sync(FULL, qaddr, r & 31);
                        if (r >= r_hi) cm[u] = 0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int r = r_lo + g + 4 * (it0 + u);
                        const uint32_t a = (uint32_t)(qb[u] + kk);
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wlo[u]) : "r"(a & ~3u));
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(whi[u]) : "r"((a & ~3u) + 4));
                        sh[u] = (a & 3) << 3;
                        C8[u] = reinterpret_cast<const uint8_t*>(&W.codes[r & 31])[k];
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint32_t in4 = (((cm[u] >> kk) & 15u) * 0x00204081u) & 0x01010101u;
                        const uint32_t Q4 = __funnelshift_r(wlo[u], whi[u], sh[u]);
                        const uint32_t X = C8[u] ^ P8;
                        const uint32_t mis4 = (((X | (X >> 1)) & 0x55u) * 0x00041041u) & 0x01010101u;
                        uint32_t val4 = (~Q4 >> 7) & 0x01010101u;
                        if (MINQ) val4 &= (((Q4 & 0x7F7F7F7Fu) + minq_add) >> 7);
                        const uint32_t act4 = val4 & in4;
                        const uint32_t mat4 = act4 & ~mis4;
                        uint32_t mm4 = act4 & mis4;
                        while (mm4) {                            // bases that differ from the primary letter: exact, direct
                            const int j = (__ffs(mm4) - 1) >> 3; mm4 &= mm4 - 1;
                            const uint32_t q = (Q4 >> (8 * j)) & 0x7F, letter = (C8[u] >> (2 * j)) & 3;
                            const int l = kk + j;
                            atomicAdd(&W.tcnt[l][letter], 1u);
                            atomicAdd(&W.tqs[l][letter], q * cur_mq);
                            atomicAdd(&W.tmq[l], cur_mq);
                            atomicAdd(&W.tq[l], q);
                        }
                        const uint32_t Qm = Q4 & (mat4 * 0xFFu);
                        cnt4 += mat4;
                        QLo += Qm & 0x00FF00FFu;
                        QHi += (Qm >> 8) & 0x00FF00FFu;
                    }
                }
            }
        }
        __syncwarp();
        mark(6, n);
        if (lane == 0) mbar_arrive(&S.empty[slot]);
    }
    if (!active) return;
    mark(7, 0);
    flush();
    uint32_t c[4]; uint64_t q[4];
    __syncwarp();
#pragma unroll
    for (int b = 0; b < 4; b++) { c[b] = W.tcnt[lane][b]; q[b] = W.tqs64[lane][b] + W.tqs[lane][b]; }
    finish_locus(R, w, lane, w0 + lane, c, q, W.tmq[lane], W.tq[lane], W.tbp[lane], fragN, pre_rb, pre_ref);
    mark(8, 0);
}

}  // namespace pb
