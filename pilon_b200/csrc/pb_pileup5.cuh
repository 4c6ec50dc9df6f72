// pb_pileup5.cuh -- the hot kernel, fifth generation: independent warps, everything learned so far.
//
// Measurements that shaped it (profiles/README.md): the per-base work is a long dependency chain, so a
// warp runs at ~0.12 IPC and the SM only fills when EVERY resident warp has work all the time.  The
// tile kernels (k_pileup3/4) serialise start-up, batch switches and the epilogue behind mbarriers and
// keep at most 3 of 7 consumer warps busy; this kernel goes back to k_pileup2's structure -- one warp
// owns one 32-locus window and never waits for another warp -- and keeps the later improvements:
//
//   staging    lane <-> candidate descriptor (prefetched one chunk ahead); overlapping rows are
//              compacted; each lane copies the <= 3 16-byte blocks of quality bytes and 12 bytes of
//              2-bit codes its row contributes to this window with cp.async into a double buffer.
//   compute    lane <-> row geometry (column mask, shared address of column 0, flags); rows whose
//              (adjMq + 1) equals the chunk's dominant value are accumulated 4 rows x 4 loci per
//              step with byte-SIMD arithmetic, four independent chains in flight per lane;
//              bases that differ from the locus' primary letter, rows with another mapping quality,
//              reads without qualities, invalid reads and soft clips take exact slower paths that
//              update the per-window table with native 32-bit shared-memory atomics.
//   flush      warp-uniform shuffle reduction of the packed registers into the table.
//   epilogue   finish_locus(): sparse merge + BaseCall + pass-1 classification, one write per plane.
#pragma once
#include "pb_pileup4.cuh"

namespace pb {

static constexpr int P5_WARPS = 8;

struct __align__(8) Geo5 { uint32_t colmask; uint16_t qoff; uint16_t flags; };   // flags: mq1 | valid<<9 | hasq<<10 | fast<<11

struct __align__(16) Warp5 {
    uint8_t pad[32];
    uint8_t rows[2][32][64];                     // 48 quality bytes + 12 code bytes (+4) per staged row
    uint8_t tail[16];
    Geo5 geo[2][32];
    uint32_t tqs[32][4];                         // 32-bit running tables (see Warp4): native shared atomics
    uint32_t tcnt[32][4];
    uint32_t tmq[32], tq[32], tbp[32];
    unsigned long long codes[32];
};

template <bool MINQ>
__global__ void __launch_bounds__(P5_WARPS * 32, 4) k_pileup5(const RegionDev R, const PileBatches PB) {
    const int n_batches = PB.n;
    extern __shared__ __align__(16) uint8_t smem_raw5[];
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * P5_WARPS + (threadIdx.x >> 5);
    if (w >= R.n_win) return;                        // warps are independent: no block-level barrier below
    Warp5& W = reinterpret_cast<Warp5*>(smem_raw5)[threadIdx.x >> 5];
    const int32_t w0 = (int32_t)(w << 5);
    const int g = lane >> 3, k = lane & 7, kk = k << 2;
    const int min_qual = R.cfg.min_qual;
    const uint32_t defq = (uint32_t)R.cfg.default_qual;
    const uint32_t minq_add = (uint32_t)(0x80 - (min_qual > 128 ? 128 : min_qual)) * 0x01010101u;
    const uint32_t rows_base = smem_u32(&W.rows[0][0][0]);

#pragma unroll
    for (int b = 0; b < 4; b++) { W.tqs[lane][b] = 0; W.tcnt[lane][b] = 0; }
    W.tmq[lane] = 0; W.tq[lane] = 0; W.tbp[lane] = 0;
    // epilogue inputs, fetched now so that their latency hides behind the accumulation
    const uint32_t pre_rb = R.rare_bits[w];
    const uint8_t pre_ref = ((int64_t)w0 + lane < R.size) ? ref_at(R, (int64_t)R.start + w0 + lane) : (uint8_t)'N';
    uint32_t P8 = 0;                                  // primary letters of my 4 loci = reference bases
    {
        const int rc = ref_class(pre_ref);
        const uint32_t code = (uint32_t)(rc < 4 ? rc : 0);
#pragma unroll
        for (int j = 0; j < 4; j++) P8 |= __shfl_sync(FULL, code, kk + j) << (2 * j);
    }
    // candidate segment range of every batch (lane <-> batch): one load latency for all of them
    uint32_t my_slo = 0, my_shi = 0;
    if (lane < n_batches) {
        const PileBatch& Bl = PB.b[lane];
        if (Bl.flags & 2) {
            const int64_t x = (int64_t)w0 - Bl.fwd + 1;
            const int64_t y = (int64_t)w0 + 32 + Bl.back;
            int64_t khi = (y + 31) >> 5; if (khi > R.n_win) khi = R.n_win;
            my_slo = x <= 0 ? 0u : Bl.win_first[x >> 5];
            my_shi = (y > ((int64_t)R.n_win << 5)) ? Bl.n_cigar : Bl.win_first[khi];
        }
    }
    __syncwarp();

    uint32_t cnt4 = 0, QLo = 0, QHi = 0, cur_mq = 0, nrows = 0, fragN = 0, nprev = 0, rows_total = 0;
    bool spilled = false;

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
                W.tqs[l][letter] += Qj * cur_mq;             // <= 1020 rows * 127 * 256 per flush
                W.tmq[l] += cj * cur_mq;
                W.tq[l] += Qj;
            }
            __syncwarp();
        }
        cnt4 = 0; QLo = 0; QHi = 0; nrows = 0;
    };
    // 32-bit quality sums are folded into the 64-bit output plane long before they could overflow
    auto spill = [&]() {
        __syncwarp();
        if ((int64_t)w0 + lane < R.size) {
            long long* o = reinterpret_cast<long long*>(R.o_qs) + 4 * ((int64_t)w0 + lane);
#pragma unroll
            for (int b = 0; b < 4; b++) { o[b] = (spilled ? o[b] : 0) + (long long)W.tqs[lane][b]; W.tqs[lane][b] = 0; }
        }
        spilled = true; rows_total = 0;
        __syncwarp();
    };

    for (int bb = 0; bb < n_batches; bb++) {
        const uint32_t slo = __shfl_sync(FULL, my_slo, bb), shi = __shfl_sync(FULL, my_shi, bb);
        const Seg* __restrict__ segs = PB.b[bb].seg;
        const uint8_t* __restrict__ gquals = PB.b[bb].quals;
        const uint8_t* __restrict__ gbases = PB.b[bb].bases2;
        const bool bfrag = PB.b[bb].flags & 1;
        if (!(PB.b[bb].flags & 2)) continue;
        const int nchunks = (int)((shi - slo + 31) >> 5);
        const Seg none = {0, 0, 0, 0};
        Seg next = none;
        if (slo + lane < shi) next = segs[slo + lane];
        uint32_t dom_stage = cur_mq;        // dominant (adjMq + 1) of the chunk being staged

        // ---- stage chunk c into buffer `buf`; returns (#rows, dominant mq of the chunk) ----
        auto stage = [&](int c, int buf, uint32_t& dom_out) -> int {
            const Seg mine = next;
            next = none;
            const uint32_t nx = slo + ((uint32_t)(c + 1) << 5) + lane;
            if (nx < shi) next = segs[nx];                                  // prefetch the next chunk's descriptors
            const bool ov = mine.len > 0 && mine.loc0 < w0 + 32 && mine.loc0 + mine.len > w0;
            const bool valid = mine.w & SEG_VALID, hasq = mine.w & SEG_HASQ;
            const uint32_t mq1 = mine.w & 0xFFFF;
            const bool elig = ov && valid && hasq;
            if (__any_sync(FULL, elig && mq1 != dom_stage)) {               // vote only when some row disagrees
                const unsigned peers = __match_any_sync(FULL, elig ? mq1 : (0x10000u + lane));
                const uint32_t votes = elig ? (((uint32_t)__popc(peers) << 17) | ((mq1 == dom_stage) ? 0x10000u : 0u) | mq1) : 0u;
                const uint32_t best = __reduce_max_sync(FULL, votes);
                if (best) dom_stage = best & 0xFFFF;
            }
            dom_out = dom_stage;
            const unsigned ovm = __ballot_sync(FULL, ov);
            if (ov) {
                const int row = __popc(ovm & ((1u << lane) - 1));
                const int cA = mine.loc0 > w0 ? mine.loc0 - w0 : 0;
                const int cB = mine.loc0 + mine.len - w0 < 32 ? mine.loc0 + mine.len - w0 : 32;
                Geo5 ge;
                ge.colmask = (cB == 32 ? 0xFFFFFFFFu : ((1u << cB) - 1)) & ~((1u << cA) - 1);
                const bool fast = elig && mq1 == dom_stage;
                ge.flags = (uint16_t)((mq1 & 0x1FF) | (valid ? 0x200u : 0u) | (hasq ? 0x400u : 0u) | (fast ? 0x800u : 0u));
                ge.qoff = 0;
                if (valid && !(R.exp_flags & 4)) {
                    const uint32_t i0 = mine.src + (uint32_t)(w0 + cA - mine.loc0);      // base index of column cA
                    uint8_t* rq = W.rows[buf][row];
                    const uint32_t ga = i0 & ~15u;
                    const int nblk = (int)(((i0 + (uint32_t)(cB - cA) - 1) >> 4) - (i0 >> 4)) + 1;
                    cp_async16(rq, gquals + ga);
                    if (nblk > 1) cp_async16(rq + 16, gquals + ga + 16);
                    if (nblk > 2) cp_async16(rq + 32, gquals + ga + 32);
                    const uint32_t b0 = i0 >> 2, ba = b0 & ~3u;
                    cp_async4(rq + 48, gbases + ba); cp_async4(rq + 52, gbases + ba + 4); cp_async4(rq + 56, gbases + ba + 8);
                    // shared offset (from rows_base, biased by 32) of window column 0's quality byte
                    ge.qoff = (uint16_t)((buf * 32 + row) * 64 + (int)(i0 - ga) - cA + 32);
                    // code realignment recipe (applied after landing): bit shift | cA << 8, parked in the row's spare bytes
                    *reinterpret_cast<uint32_t*>(rq + 60) = (8 * (b0 & 3) + 2 * (i0 & 3)) | ((uint32_t)cA << 8);
                }
                W.geo[buf][row] = ge;
            }
            return __popc(ovm);
        };

        // ---- compute one staged chunk of n rows (lane <-> row for the geometry) ----
        auto compute = [&](int buf, int n, uint32_t dom) {
            if (dom != cur_mq) { flush(); cur_mq = dom; }
            rows_total += 32;
            if (rows_total > P4_SPILL_ROWS) { flush(); spill(); }
            uint32_t gm = 0, gf = 0; int32_t gq = (int32_t)rows_base;
            if (lane < n) {
                const Geo5 ge = W.geo[buf][lane];
                gm = ge.colmask; gf = ge.flags; gq = (int32_t)rows_base + (int32_t)ge.qoff - 32;
                if (gf & 0x200u) {                                   // realign the row's 2-bit codes to window column 0
                    const uint32_t* wc = reinterpret_cast<const uint32_t*>(W.rows[buf][lane] + 48);
                    const uint32_t W0 = wc[0], W1 = wc[1], W2 = wc[2], rec = wc[3];
                    const uint32_t sft = rec & 0xFF, cA = (rec >> 8) & 0xFF;
                    const uint64_t raw = ((uint64_t)__funnelshift_r(W1, W2, sft) << 32) | __funnelshift_r(W0, W1, sft);
                    W.codes[lane] = raw << (2 * cA);
                }
            }
            const unsigned fastm = __ballot_sync(FULL, (gf & 0x800u) != 0);
            unsigned scalm = __ballot_sync(FULL, lane < n && !(gf & 0x800u));
            const uint32_t cm_fast = (gf & 0x800u) ? gm : 0u;
            __syncwarp();
            // ---- odd rows: lane <-> locus ----
            while (scalm) {
                const int j = __ffs(scalm) - 1; scalm &= scalm - 1;
                const uint32_t cmj = __shfl_sync(FULL, gm, j);
                const uint32_t fj = __shfl_sync(FULL, gf, j);
                const int32_t qb = __shfl_sync(FULL, gq, j);
                if ((cmj >> lane) & 1) {
                    if (!(fj & 0x200u)) atomicAdd(&W.tbp[lane], 1u);               // PileUpRegion.scala:45
                    else {
                        uint32_t qv;
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(qv) : "r"((uint32_t)(qb + lane)));
                        if (!(qv & 0x80)) {
                            const uint32_t code = (uint32_t)(W.codes[j] >> (2 * lane)) & 3;
                            const uint32_t q = (fj & 0x400u) ? qv : defq;
                            if (!MINQ || (int)q >= min_qual) {
                                const uint32_t m1 = fj & 0x1FF;
                                atomicAdd(&W.tcnt[lane][code], 1u); atomicAdd(&W.tqs[lane][code], q * m1);
                                atomicAdd(&W.tmq[lane], m1); atomicAdd(&W.tq[lane], q);
                            }
                        }
                    }
                }
            }
            // ---- fast rows: 4 rows per step (lane group <-> row), 4 loci per lane, 4 steps in flight ----
            if (fastm) {
                const int r_lo = __ffs(fastm) - 1, r_hi = 32 - __clz(fastm);
                const int iters = (r_hi - r_lo + 3) >> 2;
                if (nrows + (uint32_t)iters > 255) flush();
                nrows += (uint32_t)iters;
                for (int it0 = 0; it0 < iters; it0 += 4) {
                    uint32_t cm[4], wlo[4], whi[4], sh[4], C8[4];
                    int32_t qb[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int r = r_lo + g + 4 * (it0 + u);
                        cm[u] = __shfl_sync(FULL, cm_fast, r & 31);
                        qb[u] = __shfl_sync(FULL, gq, r & 31);
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
        };

        uint32_t dom_cur = cur_mq, dom_next = cur_mq;
        int n_cur = 0;
        if (nchunks > 0) n_cur = stage(0, 0, dom_cur);
        cp_async_commit();
        for (int c = 0; c < nchunks; c++) {
            int n_next = 0;
            if (c + 1 < nchunks) n_next = stage(c + 1, (c + 1) & 1, dom_next);
            cp_async_commit();
            cp_async_wait<1>();
            __syncwarp();
            if (!(R.exp_flags & 2)) compute(c & 1, n_cur, dom_cur);
            __syncwarp();
            n_cur = n_next; dom_cur = dom_next;
        }
        flush();
        __syncwarp();
        const uint32_t nnow = W.tcnt[lane][0] + W.tcnt[lane][1] + W.tcnt[lane][2] + W.tcnt[lane][3];
        if (bfrag) fragN += nnow - nprev;
        nprev = nnow;
        __syncwarp();
    }
    cp_async_wait<0>();
    __syncwarp();
    uint32_t c[4]; uint64_t q[4];
    const bool inr = (int64_t)w0 + lane < R.size;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        c[b] = W.tcnt[lane][b];
        q[b] = (uint64_t)W.tqs[lane][b] + ((spilled && inr) ? (uint64_t)R.o_qs[4 * ((int64_t)w0 + lane) + b] : 0ull);
    }
    if (R.exp_flags & 1) { if (c[0] == 0xdeadbeef) R.o_mq[w0 + lane] = (int32_t)q[0]; return; }
    finish_locus(R, w, lane, w0 + lane, c, q, W.tmq[lane], W.tq[lane], W.tbp[lane], fragN, pre_rb, pre_ref);
}

}  // namespace pb
