// pb_pileup3.cuh -- the hot kernel, third generation: warp-specialised CTA tile.
//
// A CTA owns a tile of 7 windows x 32 loci.  Warp 0 is the PRODUCER: it walks the tile's candidate
// segment descriptors once, 32 per chunk (lane <-> descriptor), and for every segment that overlaps
// the tile copies the 16-byte-aligned run of quality bytes and of 2-bit codes the tile needs with
// per-lane cp.async (SASS LDGSTS; cp.async.bulk/UBLKCP takes warp-uniform operands, so 32 different
// rows per chunk would serialise -- measured, see profiles/) into a 4-deep ring of shared-memory
// chunks guarded by full/empty mbarriers; copy completion is signalled to the full barrier with
// cp.async.mbarrier.arrive.  Warps 1..7 are CONSUMERS, one per window: for every chunk they pick the
// rows that overlap their window and accumulate them with the byte-SIMD scheme of k_pileup2
// (lane = (row group, column quad); packed 8-bit counts / 16-bit quality sums for bases that equal
// the locus' primary letter and carry the chunk's dominant mapping quality).  Everything else stays
// exact through slower paths: other letters update the per-window table directly, rows with another
// mapping quality / without qualities / of invalid reads go through a lane-per-locus loop, and the
// packed registers are flushed with a shuffle reduction (no atomics) when the dominant quality
// changes, before they could overflow, and at the end of each batch (fragCoverage snapshot).
// The epilogue (BaseCall + pass-1 classification + single write of every output) is finish_locus().
//
// Each read byte is fetched from HBM/L2 once per tile; descriptor handling is paid once per tile
// instead of once per window; global-load latency is hidden by the ring instead of by occupancy.
#pragma once
#include "pb_pileup2.cuh"

namespace pb {

static constexpr int P3_CW = 7;                  // consumer warps = windows per tile
static constexpr int P3_TILE = P3_CW * 32;       // loci per tile
static constexpr int P3_NS = 4;                  // ring depth (chunks)
static constexpr int P3_QB = 176;                // quality bytes per staged row (<= 150 + 15, 16-byte blocks)
static constexpr int P3_CB = 48;                 // code bytes per staged row
static constexpr int P3_ROWB = P3_QB + P3_CB;    // 224
static constexpr int P3_FRONT = 256;             // slack for (masked) loads left of a row

enum : uint32_t { ROW_NONE = 0, ROW_FAST = 1, ROW_SCALAR = 2, ROW_INVALID = 3 };
enum : uint32_t { CH_DATA = 0, CH_EOB = 1, CH_EOT = 2 };

// c01: c0 | c1 << 16 (tile columns, c0 < c1);  q: qoff | mq1 << 16;  k: kind | hasq << 8 | cbit << 16
struct __align__(16) Row3 { uint32_t c01, q, k, pad; };
struct __align__(16) Chunk3 { uint32_t type, dom, frag, pad; };

struct __align__(16) Warp3 {                     // per consumer warp
    unsigned long long tqs[32][4];
    uint32_t tcnt[32][4];
    uint32_t tmq[32], tq[32], tbp[32];
    unsigned long long codes[32];                // window-aligned 2-bit codes of the chunk's rows
};

struct __align__(128) Smem3 {
    uint8_t front[P3_FRONT];
    uint8_t rows[P3_NS][32][P3_ROWB];
    uint8_t tail[128];
    Row3 hdr[P3_NS][32];
    Chunk3 chunk[P3_NS];
    unsigned long long full[P3_NS], empty[P3_NS];
    Warp3 warp[P3_CW];
};

// ---- mbarrier / bulk-copy primitives (PTX ISA, sm_90+) ----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(unsigned long long* b, uint32_t bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// bounded spin: a protocol bug must surface as an error flag, never as a hung GPU
__device__ __forceinline__ bool mbar_wait(unsigned long long* b, uint32_t parity, int* err) {
    const uint32_t a = smem_u32(b);
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 26); spin++) {
        uint32_t ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return true;
    }
    atomicOr(err, 8);
    return false;
}
// every cp.async issued so far by this thread arrives on `b` when it lands; the barrier's expected
// count already includes this arrival (.noinc)
__device__ __forceinline__ void cp_async_arrive_noinc(unsigned long long* b) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}

// ---------------------------------------------------------------------------------------------
template <bool MINQ>
__global__ void __launch_bounds__((P3_CW + 1) * 32) k_pileup3(RegionDev R, const DevBatch* __restrict__ batches, int n_batches) {
    extern __shared__ __align__(128) uint8_t smem_raw3[];
    Smem3& S = *reinterpret_cast<Smem3*>(smem_raw3);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t t0 = (int32_t)blockIdx.x * P3_TILE;               // first locus index of the tile
    int* err = &R.sc->error;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P3_NS; s++) { mbar_init(&S.full[s], 64); mbar_init(&S.empty[s], P3_CW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        // ================================ PRODUCER ================================
        uint32_t n = 0;                                              // chunk counter
        uint32_t dom = 0;
        auto acquire = [&](uint32_t slot) -> bool {
            if (n >= P3_NS) return mbar_wait(&S.empty[slot], ((n / P3_NS) + 1) & 1, err);
            return true;
        };
        bool alive = true;
        for (int bb = 0; bb < n_batches && alive; bb += 32) {
            // candidate segment range of up to 32 batches at once (lane <-> batch): one load latency for all
            uint32_t my_slo = 0, my_shi = 0;
            if (bb + lane < n_batches) {
                const DevBatch& Bl = batches[bb + lane];
                if (Bl.n_reads) {
                    const int64_t x = (int64_t)t0 - Bl.reach[0] + 1;
                    const int64_t y = (int64_t)t0 + P3_TILE + Bl.reach[1];
                    int64_t khi = (y + 31) >> 5; if (khi > R.n_win) khi = R.n_win;
                    my_slo = x <= 0 ? 0u : Bl.win_first[x >> 5];
                    my_shi = (y > ((int64_t)R.n_win << 5)) ? (uint32_t)Bl.n_cigar : Bl.win_first[khi];
                }
            }
            const int nbb = n_batches - bb < 32 ? n_batches - bb : 32;
            for (int bi = 0; bi < nbb && alive; bi++) {
                // registers, not the struct in global memory: the asm memory clobbers below would reload every field per chunk
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
                for (uint32_t sb = slo; sb < shi && alive; sb += 32, n++) {
                    const Seg mine = next;
                    next = none;
                    if (sb + 32 + lane < shi) next = segs[sb + 32 + lane];      // prefetch the next chunk's descriptors
                    const uint32_t slot = n % P3_NS;
                    alive = acquire(slot);
                    if (!alive) break;
                    const bool ov = mine.len > 0 && mine.loc0 < t0 + P3_TILE && mine.loc0 + mine.len > t0;
                    const bool valid = mine.w & SEG_VALID, hasq = mine.w & SEG_HASQ;
                    const uint32_t mq1 = mine.w & 0xFFFF;
                    const bool elig = ov && valid && hasq;
                    // dominant (adjMq + 1) of the chunk: the value most eligible rows carry; keep the previous one on ties
                    const unsigned peers = __match_any_sync(FULL, elig ? mq1 : (0x10000u + lane));
                    const uint32_t votes = elig ? (((uint32_t)__popc(peers) << 17) | ((mq1 == dom) ? 0x10000u : 0u) | mq1) : 0u;
                    const uint32_t best = __reduce_max_sync(FULL, votes);
                    if (best) dom = best & 0xFFFF;
                    Row3 h; h.c01 = 0; h.q = 0; h.k = ROW_NONE; h.pad = 0;
                    if (ov) {
                        const int a = mine.loc0 > t0 ? mine.loc0 : t0;
                        const int e = mine.loc0 + mine.len < t0 + P3_TILE ? mine.loc0 + mine.len : t0 + P3_TILE;
                        const uint32_t c0 = (uint32_t)(a - t0), c1 = (uint32_t)(e - t0);
                        h.c01 = c0 | (c1 << 16);
                        if (!valid) h.k = ROW_INVALID;
                        else {
                            const uint32_t i0 = mine.src + (uint32_t)(a - mine.loc0), nb = c1 - c0;
                            const uint32_t ga = i0 & ~15u, qblk = (((i0 + nb - 1) >> 4) - (i0 >> 4)) + 1;
                            const uint32_t b0 = i0 >> 2, gb = b0 & ~15u, cblk = ((((i0 + nb - 1) >> 2) >> 4) - (b0 >> 4)) + 1;
                            uint8_t* row = S.rows[slot][lane];
                            const uint8_t* qs = gquals + ga;
                            for (uint32_t t = 0; t < qblk; t++) cp_async16(row + 16 * t, qs + 16 * t);
                            const uint8_t* cs = gbases + gb;
                            for (uint32_t t = 0; t < cblk; t++) cp_async16(row + P3_QB + 16 * t, cs + 16 * t);
                            h.q = (i0 - ga) | (mq1 << 16);
                            const uint32_t cbit = 8 * (b0 - gb) + 2 * (i0 & 3);
                            h.k = ((hasq && mq1 == dom) ? ROW_FAST : ROW_SCALAR) | ((hasq ? 1u : 0u) << 8) | (cbit << 16);
                        }
                    }
                    cp_async_arrive_noinc(&S.full[slot]);
                    S.hdr[slot][lane] = h;
                    if (lane == 0) { Chunk3 ch; ch.type = CH_DATA; ch.dom = dom; ch.frag = 0; ch.pad = 0; S.chunk[slot] = ch; }
                    mbar_arrive(&S.full[slot]);
                }
                if (!alive) break;
                {   // end of batch: consumers flush and take the fragCoverage snapshot (GenomeRegion.scala:290-298)
                    const uint32_t slot = n % P3_NS;
                    alive = acquire(slot);
                    if (!alive) break;
                    if (lane == 0) { Chunk3 ch; ch.type = CH_EOB; ch.dom = dom; ch.frag = bfrag; ch.pad = 0; S.chunk[slot] = ch; }
                    cp_async_arrive_noinc(&S.full[slot]);
                    mbar_arrive(&S.full[slot]);
                    n++;
                }
            }
        }
        if (alive) {
            const uint32_t slot = n % P3_NS;
            if (acquire(slot)) {
                if (lane == 0) { Chunk3 ch; ch.type = CH_EOT; ch.dom = 0; ch.frag = 0; ch.pad = 0; S.chunk[slot] = ch; }
                cp_async_arrive_noinc(&S.full[slot]);
                mbar_arrive(&S.full[slot]);
            }
        }
        return;
    }

    // ================================== CONSUMERS ==================================
    const int cw = warp - 1;
    const int32_t wc = cw * 32;                                     // tile column of my window
    const int64_t w = (int64_t)blockIdx.x * P3_CW + cw;             // global window index
    const bool active = w < R.n_win;
    const int32_t w0 = t0 + wc;
    Warp3& W = S.warp[cw];
    const int g = lane >> 3, k = lane & 7, kk = k << 2;
    const int min_qual = R.cfg.min_qual;
    const uint32_t defq = (uint32_t)R.cfg.default_qual;
    const uint32_t minq_add = (uint32_t)(0x80 - (min_qual > 128 ? 128 : min_qual)) * 0x01010101u;

#pragma unroll
    for (int b = 0; b < 4; b++) { W.tqs[lane][b] = 0; W.tcnt[lane][b] = 0; }
    W.tmq[lane] = 0; W.tq[lane] = 0; W.tbp[lane] = 0;
    // epilogue inputs, fetched now so that their latency hides behind the accumulation
    const uint32_t pre_rb = active ? R.rare_bits[w] : 0u;
    const uint8_t pre_ref = (active && (int64_t)w0 + lane < R.size) ? ref_at(R, (int64_t)R.start + w0 + lane) : (uint8_t)'N';
    uint32_t P8 = 0;                                                // primary letters of my 4 loci = reference bases
    if (active) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t l = (int64_t)w0 + kk + j;
            const int rc = l < R.size ? ref_class(ref_at(R, (int64_t)R.start + l)) : 0;
            P8 |= (uint32_t)(rc < 4 ? rc : 0) << (2 * j);
        }
    }
    __syncwarp();

    uint32_t cnt4 = 0, QLo = 0, QHi = 0, cur_mq = 0, nrows = 0, fragN = 0, nprev = 0;

    // warp-uniform flush: reduce the 4 row groups with shuffles, then lane (g,k) owns locus 4k+g
    auto flush = [&]() {
        if (__any_sync(FULL, cnt4 != 0)) {
            uint32_t c02 = cnt4 & 0x00FF00FFu, c13 = (cnt4 >> 8) & 0x00FF00FFu;     // counts of loci (0,2) and (1,3)
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
                W.tqs[l][letter] += (unsigned long long)Qj * cur_mq;
                W.tmq[l] += cj * cur_mq;
                W.tq[l] += Qj;
            }
            __syncwarp();
        }
        cnt4 = 0; QLo = 0; QHi = 0; nrows = 0;
    };

    for (uint32_t n = 0;; n++) {
        const uint32_t slot = n % P3_NS;
        if (!mbar_wait(&S.full[slot], (n / P3_NS) & 1, err)) return;
        const Chunk3 ch = S.chunk[slot];
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
        if (active) {
            if (ch.dom != cur_mq) { flush(); cur_mq = ch.dom; }
            // ---- per-row setup, lane <-> row ----
            const Row3 h = S.hdr[slot][lane];
            const uint32_t kind = h.k & 0xFF;
            const int x0 = (int)(h.c01 & 0xFFFF) - wc, x1 = (int)(h.c01 >> 16) - wc;     // window-relative columns
            const bool ov = kind != ROW_NONE && x0 < 32 && x1 > 0;
            const uint32_t lo = x0 > 0 ? (uint32_t)x0 : 0u, hi = x1 < 32 ? (uint32_t)x1 : 32u;
            const uint32_t colmask = ov ? ((hi == 32 ? 0xFFFFFFFFu : ((1u << hi) - 1)) & ~((1u << lo) - 1)) : 0u;
            const uint32_t rowaddr = smem_u32(S.rows[slot][lane]);
            const int32_t qbase = (int32_t)rowaddr + (int32_t)(h.q & 0xFFFF) - x0;       // smem address of window column 0
            if (ov && kind != ROW_INVALID) {
                const int32_t cbw = (int32_t)(h.k >> 16) - 2 * x0;                      // bit offset of column 0's code
                const int32_t wi = cbw >> 5;                                              // floor: may be negative (masked columns)
                const uint32_t ca = rowaddr + P3_QB + (uint32_t)(wi * 4);
                uint32_t W0, W1, W2;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(W0) : "r"(ca));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(W1) : "r"(ca + 4));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(W2) : "r"(ca + 8));
                const uint32_t sft = (uint32_t)cbw & 31;
                W.codes[lane] = ((unsigned long long)__funnelshift_r(W1, W2, sft) << 32) | __funnelshift_r(W0, W1, sft);
            }
            const unsigned fastm = __ballot_sync(FULL, ov && kind == ROW_FAST);
            unsigned scalm = __ballot_sync(FULL, ov && kind != ROW_FAST);
            const uint32_t cm_fast = (ov && kind == ROW_FAST) ? colmask : 0u;
            __syncwarp();
            // ---- odd rows: lane <-> locus (other mapping quality, no qualities, invalid reads / soft clips) ----
            while (scalm) {
                const int j = __ffs(scalm) - 1; scalm &= scalm - 1;
                const uint32_t cmj = __shfl_sync(FULL, colmask, j);
                const uint32_t hk = __shfl_sync(FULL, h.k, j);
                const uint32_t hq = __shfl_sync(FULL, h.q, j);
                const int32_t qb = __shfl_sync(FULL, qbase, j);
                if ((cmj >> lane) & 1) {
                    if ((hk & 0xFF) == ROW_INVALID) W.tbp[lane] += 1;                     // PileUpRegion.scala:45
                    else {
                        uint32_t qv;
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(qv) : "r"((uint32_t)(qb + lane)));
                        if (!(qv & 0x80)) {
                            const uint32_t code = (uint32_t)(W.codes[j] >> (2 * lane)) & 3;
                            const uint32_t q = ((hk >> 8) & 1) ? qv : defq;
                            if (!MINQ || (int)q >= min_qual) {
                                const uint32_t mq1 = hq >> 16;
                                W.tcnt[lane][code] += 1; W.tqs[lane][code] += (unsigned long long)(q * mq1);
                                W.tmq[lane] += mq1; W.tq[lane] += q;
                            }
                        }
                    }
                }
            }
            // ---- fast rows: 4 rows per step (one per lane group), 4 loci per lane ----
            if (fastm) {
                const int r_lo = __ffs(fastm) - 1, r_hi = 32 - __clz(fastm);
                const int iters = (r_hi - r_lo + 3) >> 2;
                if (nrows + (uint32_t)iters > 255) flush();
                nrows += (uint32_t)iters;
                for (int it = 0, r = r_lo + g; it < iters; it++, r += 4) {
                    const uint32_t cm = __shfl_sync(FULL, cm_fast, r & 31);
                    const int32_t qb = __shfl_sync(FULL, qbase, r & 31);
                    const uint32_t in4 = r < r_hi ? ((((cm >> kk) & 15u) * 0x00204081u) & 0x01010101u) : 0u;
                    const uint32_t a = (uint32_t)(qb + kk);
                    uint32_t wlo, whi;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wlo) : "r"(a & ~3u));
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(whi) : "r"((a & ~3u) + 4));
                    const uint32_t Q4 = __funnelshift_r(wlo, whi, (a & 3) << 3);
                    const uint32_t C8 = reinterpret_cast<const uint8_t*>(&W.codes[r & 31])[k];
                    const uint32_t X = C8 ^ P8;
                    const uint32_t mis4 = (((X | (X >> 1)) & 0x55u) * 0x00041041u) & 0x01010101u;
                    uint32_t val4 = (~Q4 >> 7) & 0x01010101u;
                    if (MINQ) val4 &= (((Q4 & 0x7F7F7F7Fu) + minq_add) >> 7);
                    const uint32_t act4 = val4 & in4;
                    const uint32_t mat4 = act4 & ~mis4;
                    const uint32_t mm4 = act4 & mis4;
                    if (mm4) {                                   // bases that differ from the primary letter: exact, direct
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            if ((mm4 >> (8 * j)) & 1) {
                                const uint32_t q = (Q4 >> (8 * j)) & 0x7F, letter = (C8 >> (2 * j)) & 3;
                                const int l = kk + j;
                                atomicAdd(&W.tcnt[l][letter], 1u);
                                atomicAdd(&W.tqs[l][letter], (unsigned long long)(q * cur_mq));
                                atomicAdd(&W.tmq[l], cur_mq);
                                atomicAdd(&W.tq[l], q);
                            }
                        }
                    }
                    const uint32_t Qm = Q4 & (mat4 * 0xFFu);
                    cnt4 += mat4;
                    QLo += Qm & 0x00FF00FFu;
                    QHi += (Qm >> 8) & 0x00FF00FFu;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.empty[slot]);
    }
    if (!active) return;
    flush();
    uint32_t c[4]; uint64_t q[4];
#pragma unroll
    for (int b = 0; b < 4; b++) { c[b] = W.tcnt[lane][b]; q[b] = W.tqs[lane][b]; }
    finish_locus(R, w, lane, w0 + lane, c, q, W.tmq[lane], W.tq[lane], W.tbp[lane], fragN, pre_rb, pre_ref);
}

}  // namespace pb
