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
#include "pb_epilogue.cuh"

namespace pb {

static constexpr int P5_WARPS = 8;
static constexpr uint32_t P5_SPILL_ROWS = 120000;   // rows between folds of the 32-bit quality sums: 120000 * 32512 < 2^32

struct __align__(8) Geo5 { uint32_t colmask; uint16_t qoff; uint16_t flags; };   // flags: mq1 | valid<<9 | hasq<<10 | fast<<11

struct __align__(16) Warp5 {
    uint8_t pad[32];
    uint8_t rows[2][32][64];                     // 48 quality bytes + 12 code bytes (+4) per staged row
    uint8_t tail[16];
    Geo5 geo[2][32];
    uint32_t tqs[32][4];                         // 32-bit running tables : native shared atomics
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
    // candidate segment range of 32 batches at a time (lane <-> batch): one load latency for all of them
    uint32_t my_slo = 0, my_nseg = 0;
    auto load_ranges = [&](int b0) {
        my_slo = 0; my_nseg = 0;
        if (b0 + lane < n_batches) {
            const PileBatch& Bl = pile_batch(PB, b0 + lane);
            if (Bl.flags & 2) {
                const int64_t x = (int64_t)w0 - Bl.reach[0] + 1;
                const int64_t y = (int64_t)w0 + 32 + Bl.reach[1];
                int64_t khi = (y + 31) >> 5; if (khi > R.n_win) khi = R.n_win;
                my_slo = x <= 0 ? 0u : Bl.win_first[x >> 5];
                const uint32_t shi = (y > ((int64_t)R.n_win << 5)) ? Bl.n_cigar : Bl.win_first[khi];
                my_nseg = shi > my_slo ? shi - my_slo : 0u;
            }
        }
    };
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
    auto batch_end = [&](bool frag) {                   // fragCoverage snapshot (GenomeRegion.scala:290-298)
        flush();
        const uint32_t nnow = W.tcnt[lane][0] + W.tcnt[lane][1] + W.tcnt[lane][2] + W.tcnt[lane][3];
        if (frag) fragN += nnow - nprev;
        nprev = nnow;
        __syncwarp();
    };

    // ---- one flat, software-pipelined sequence of (batch, chunk) pairs: stage(i+1) | wait | compute(i) ----
    // staging cursor
    int s_b = -1; uint32_t s_c = 0, s_nch = 0, s_slo = 0, s_nseg = 0;
    auto advance = [&]() -> bool {                      // move the cursor to the next non-empty chunk; false at the end
        s_c++;
        while (s_b < 0 || s_c >= s_nch) {
            s_b++; s_c = 0;
            if (s_b >= n_batches) return false;
            if ((s_b & 31) == 0) load_ranges(s_b);      // (warp-uniform) the next 32 batches' ranges
            s_nseg = __shfl_sync(FULL, my_nseg, s_b & 31);
            s_nch = (s_nseg + 31) >> 5;
            s_slo = __shfl_sync(FULL, my_slo, s_b & 31);
        }
        return true;
    };
    const Seg none = {0, 0, 0, 0};
    Seg next = none;
    bool have_next = advance();
    if (have_next && (s_c << 5) + lane < s_nseg) next = pile_batch(PB, s_b).seg[s_slo + (s_c << 5) + lane];
    uint32_t dom_stage = 0;

    // per staged buffer: number of rows, dominant mq, batch index (-1: nothing staged)
    int n_b0 = 0, n_b1 = 0; uint32_t dom_b0 = 0, dom_b1 = 0; int b_b0 = -1, b_b1 = -1;     // (scalars: no local-memory arrays)

    auto stage = [&](int buf) {                         // stages the chunk under the cursor, then advances it
        const Seg mine = next;
        const int my_b = s_b;
        const uint8_t* __restrict__ gquals = pile_batch(PB, my_b).quals;
        const uint8_t* __restrict__ gbases = pile_batch(PB, my_b).bases2;
        have_next = advance();
        next = none;
        if (have_next && (s_c << 5) + lane < s_nseg) next = pile_batch(PB, s_b).seg[s_slo + (s_c << 5) + lane];   // prefetch
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
        if (buf) { n_b1 = __popc(ovm); dom_b1 = dom_stage; b_b1 = my_b; } else { n_b0 = __popc(ovm); dom_b0 = dom_stage; b_b0 = my_b; }
    };

    // one SIMD step for row r of the buffer (lane group g), 4 loci per lane
    auto simd_step = [&](uint32_t cm, uint32_t wlo, uint32_t whi, uint32_t sh, uint32_t C8) {
        const uint32_t in4 = (((cm >> kk) & 15u) * 0x00204081u) & 0x01010101u;
        const uint32_t Q4 = __funnelshift_r(wlo, whi, sh);
        const uint32_t X = C8 ^ P8;
        const uint32_t mis4 = (((X | (X >> 1)) & 0x55u) * 0x00041041u) & 0x01010101u;
        uint32_t val4 = (~Q4 >> 7) & 0x01010101u;
        if (MINQ) val4 &= (((Q4 & 0x7F7F7F7Fu) + minq_add) >> 7);
        const uint32_t act4 = val4 & in4;
        const uint32_t mat4 = act4 & ~mis4;
        uint32_t mm4 = act4 & mis4;
        while (mm4) {                            // bases that differ from the primary letter: exact, direct
            const int j = (__ffs(mm4) - 1) >> 3; mm4 &= mm4 - 1;
            const uint32_t q = (Q4 >> (8 * j)) & 0x7F, letter = (C8 >> (2 * j)) & 3;
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
    };

    int last_b = -1;
    auto compute = [&](int buf) {
        const int n = buf ? n_b1 : n_b0;
        const uint32_t dom = buf ? dom_b1 : dom_b0;
        const int bb = buf ? b_b1 : b_b0;
        if (bb != last_b) {                             // first chunk of another batch: close the previous one
            if (last_b >= 0) batch_end(pile_batch(PB, last_b).flags & 1);
            last_b = bb;
        }
        if (n == 0) return;
        if (dom != cur_mq) { flush(); cur_mq = dom; }
        rows_total += 32;
        if (rows_total > P5_SPILL_ROWS) { flush(); spill(); }
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
        // ---- fast rows: 4 rows per step (lane group <-> row), 4 loci per lane ----
        if (fastm) {
            const int r_lo = __ffs(fastm) - 1, r_hi = 32 - __clz(fastm);
            const int iters = (r_hi - r_lo + 3) >> 2;
            if (nrows + (uint32_t)iters > 255) flush();
            nrows += (uint32_t)iters;
            int it0 = 0;
            // full trips: 4 independent steps in flight (all shuffles, then all loads, then the arithmetic)
            for (; it0 + 4 <= iters; it0 += 4) {
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
                for (int u = 0; u < 4; u++) simd_step(cm[u], wlo[u], whi[u], sh[u], C8[u]);
            }
            // remainder: one step per trip
            for (; it0 < iters; it0++) {
                const int r = r_lo + g + 4 * it0;
                uint32_t cm = __shfl_sync(FULL, cm_fast, r & 31);
                const int32_t qb = __shfl_sync(FULL, gq, r & 31);
                if (r >= r_hi) cm = 0;
                const uint32_t a = (uint32_t)(qb + kk);
                uint32_t wlo, whi;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wlo) : "r"(a & ~3u));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(whi) : "r"((a & ~3u) + 4));
                simd_step(cm, wlo, whi, (a & 3) << 3, reinterpret_cast<const uint8_t*>(&W.codes[r & 31])[k]);
            }
        }
    };

    int cur = 0;
    bool have_cur = false;
    if (have_next) { stage(0); have_cur = true; }
    cp_async_commit();
    while (have_cur) {
        bool staged = false;
        if (have_next) { stage(cur ^ 1); staged = true; }
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        if (!(R.exp_flags & 2)) compute(cur);
        __syncwarp();
        cur ^= 1; have_cur = staged;
    }
    if (last_b >= 0) batch_end(pile_batch(PB, last_b).flags & 1);
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
