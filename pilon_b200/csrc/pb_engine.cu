// pb_engine.cu -- host orchestration + C ABI of libpilonb200.so (see include/pilon_b200.h).
//
// One engine = one CUDA stream on one GPU.  Region lifetime mirrors
// GenomeRegion.initializePileUps / finalizePileUps (GenomeRegion.scala:149-155): begin, add the
// region's read batches (BamFile.process, BamFile.scala:108-148), finish (PileUpRegion.postProcess
// + GenomeRegion.postProcess pass 1, GenomeRegion.scala:214-272) and read everything back.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <chrono>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/transform_iterator.h>
#include <thrust/iterator/counting_iterator.h>

#include "pb_pileup7.cuh"
#include "pb_pileup7c.cuh"

using namespace pb;


static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(PB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));          \
    } while (0)

namespace {

// device/pinned scalar block: Scalars followed by the per-batch reach pairs (k_prep's atomicMax targets)
// then the per-batch base counts (BC_SPREAD partial sums per batch: same-address L2 atomics serialise)
constexpr size_t MAX_BATCHES = 256;
constexpr size_t SC_REACH_OFF = (sizeof(Scalars) + sizeof(ScalarSlot) * SC_SLOTS + 15) & ~(size_t)15;
constexpr size_t SC_BC_OFF = SC_REACH_OFF + MAX_BATCHES * 8;
constexpr size_t SC_BYTES = SC_BC_OFF + MAX_BATCHES * BC_SPREAD * 8;

// grow-only device buffer, zero-filled on growth when asked (the "rare" planes rely on it)
struct DBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes, bool zero, cudaStream_t s) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaError_t e = cudaFree(p); if (e != cudaSuccess) return e; p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) return e;
        cap = want;
        if (zero) return cudaMemsetAsync(p, 0, cap, s);
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Device memory of one region's batches (uploaded arrays, segment tables, ...): bump allocation out of blocks the engine
// keeps.  The stream-ordered allocator it replaces cost ~30 calls per region and, with several regions in flight, grew its
// pool in the middle of timed steps (an implicit device synchronisation: add_batch calls of 70-560 ms, profiles/README.md
// r2q).  Everything that used the arena is stream-ordered before the next region's uploads, so reset() is just a rewind;
// when a region needed more than one block the blocks are merged into one (cudaFree / cudaMalloc: warm-up only).
struct Arena {
    struct Block { void* p; size_t cap; };
    std::vector<Block> blocks;
    size_t cur = 0, off = 0, used = 0, high = 0;
    cudaError_t alloc(size_t bytes, void** out) {
        bytes = (bytes + 255) & ~(size_t)255;
        while (cur < blocks.size() && off + bytes > blocks[cur].cap) { cur++; off = 0; }
        if (cur == blocks.size()) {
            const size_t cap = std::max(bytes, (size_t)64 << 20);
            void* p = nullptr;
            cudaError_t e = cudaMalloc(&p, cap);
            if (e != cudaSuccess) return e;
            blocks.push_back({p, cap}); off = 0;
        }
        *out = static_cast<uint8_t*>(blocks[cur].p) + off;
        off += bytes; used += bytes;
        return cudaSuccess;
    }
    cudaError_t reset() {
        high = std::max(high, used);
        if (blocks.size() > 1) {                     // (cudaFree waits for the device: nothing is using the blocks afterwards)
            for (auto& b : blocks) { cudaError_t e = cudaFree(b.p); if (e != cudaSuccess) return e; }
            blocks.clear();
            const size_t cap = high + high / 8 + ((size_t)1 << 20);
            void* p = nullptr;
            cudaError_t e = cudaMalloc(&p, cap);
            if (e != cudaSuccess) return e;
            blocks.push_back({p, cap});
        }
        cur = off = used = 0;
        return cudaSuccess;
    }
    void release() { for (auto& b : blocks) cudaFree(b.p); blocks.clear(); cur = off = used = 0; }
};

struct HostBatch {
    DevBatch d;                      // device view
};

}  // namespace

struct pb_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // side stream: the physCov scans run beside the indel / sort / group chain (fork-join inside a pass)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evp0 = nullptr, evp1 = nullptr, ev_sc = nullptr, ev_fork = nullptr, ev_join = nullptr;
    Cfg cfg{};
    bool in_region = false;
    RegionDev R{};
    std::vector<HostBatch> batches;
    DBuf meta_flag; bool meta_used = false;     // what k_meta_expand found wrong with a batch's compact metadata (checked by the pass)
    Arena arena;                     // device memory of the current region's batches
    DBuf d_batches, d_pile;          // DevBatch[] image; PileBatch[] image (only for > PB_MAXB batches)
    // per-locus buffers
    DBuf ref, rare, gplane[2], rare_bits, pc_diff, block_sums, scalars, extra, head;
    const uint8_t* contig_host = nullptr;   // the caller's contig (valid until pb_region_finish): head upload of long-read regions
    DBuf o_cnt, o_qs, o_i32[12], o_wq, o_wmq, o_flags, o_call;
    // event buffers
    DBuf ev_key, ev, perm, sort_buf, groups, cand, work, spill_scratch, str_pool, cub_tmp, call_idx, call_entries;
    Scalars* h_sc = nullptr;         // pinned
    int64_t launches = 0;
    float last_pileup_ms = 0.f;
    bool dirty = false;              // rare planes may be non-zero after a failed run
    bool unverified = false;         // pb_region_compute passes whose device error flag nobody has read yet
    // pb_region_compute replays a captured CUDA graph of the pass while the region's batches stay the same
    int graph_state = 0;             // 0: next pass runs plainly (and grows every buffer), 1: next pass is captured, 2: replay, 3: never
    cudaGraphExec_t graph_exec = nullptr;
    int64_t graph_launches = 0;      // kernels inside the captured pass
    std::vector<DevBatch> img_host;  // image of the batch table; a captured H2D copy reads it at every replay
    // PB_PHASE_TIMING=1 (diagnostics): device time of upload / pass / download per region, printed by pb_destroy
    bool phase_timing = false; cudaEvent_t ph[4] = {}; double ph_ms[3] = {0, 0, 0}; int64_t ph_regions = 0;
    uint8_t* ref_stage = nullptr; size_t ref_stage_cap = 0;     // pinned staging for the reference window of a pageable contig
    cudaEvent_t ev_stage = nullptr;
    bool host_trace = false;         // PB_HOST_TRACE=1: host microseconds per section of pb_region_begin, printed by pb_destroy
    double ht[6] = {0, 0, 0, 0, 0, 0}; long ht_n = 0;
    double hf[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long hf_n = 0, ht_calls = 0, hf_calls = 0;
    int ctile = P7C_TILE;            // tile of the cluster scatter kernel (PB_CTILE: A/B runs)
    int spread_order = -1;           // -1 = by depth
    int64_t deep_depth = 1000;       // mean depth above which a region goes to the cluster scatter kernel ...
    int64_t spread_depth = 1000;     // ... and above which a grab spreads its descriptors over the tile (measured: profiles/README.md, r2q)
    int pileup_version = 0;          // 0 = choose per region (k_pileup7 scatter / k_pileup5 gather); PB_PILEUP=5 or 7 forces one (A/B runs)
    std::vector<PileBatch> pile_host; // image of the pileup kernels' batch table (device-side copy when there are more than PB_MAXB batches)
};

static int clean_sparse_planes(pb_engine* e);
static void drop_graph(pb_engine* e) {
    if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }
    if (e->graph_state != 3) e->graph_state = 0;
}
static int free_batches(pb_engine* e) {
    drop_graph(e);
    e->batches.clear();
    CK(e->arena.reset());
    return PB_OK;
}

extern "C" int pb_abi_version(void) { return PB_ABI_VERSION; }
extern "C" const char* pb_last_error(void) { return g_err.c_str(); }
extern "C" int pb_device_count(int* n_out) { CK(cudaGetDeviceCount(n_out)); return PB_OK; }

extern "C" int pb_create(int device, const pb_config* c, pb_engine** out) {
    if (!c || !out) return fail(PB_ERR_INVALID, "null argument");
    // the packed format folds "negative quality byte" into the uncountable flag, which is only
    // equivalent to PileUp.scala:77 when minQual >= 0; defaultQual is a phred value
    if (c->min_qual < 0) return fail(PB_ERR_UNSUPPORTED, "min_qual < 0 is not supported");
    if (c->default_qual < 0 || c->default_qual > 127) return fail(PB_ERR_UNSUPPORTED, "default_qual must be in 0..127");
    if (c->flank < 0) return fail(PB_ERR_INVALID, "flank < 0");
    CK(cudaSetDevice(device));
    pb_engine* e = new pb_engine();
    e->device = device;
    e->cfg.min_qual = c->min_qual; e->cfg.min_mq = c->min_mq; e->cfg.flank = c->flank;
    e->cfg.default_qual = c->default_qual; e->cfg.min_min_depth = c->min_min_depth;
    e->cfg.old_indel = c->old_indel; e->cfg.fix_amb = c->fix_amb; e->cfg.min_depth = c->min_depth;
    if (const char* v = getenv("PB_HOST_TRACE")) e->host_trace = atoi(v) != 0;
    if (const char* v = getenv("PB_SPREAD")) e->spread_order = atoi(v);           // A/B runs: 0 / 1 forces the descriptor order of a grab
    if (const char* v = getenv("PB_DEEP")) e->deep_depth = atoi(v);
    if (const char* v = getenv("PB_PILEUP")) { const int pv = atoi(v); if (pv == 5 || pv == 7 || pv == 8) e->pileup_version = pv; }
    CK(cudaFuncSetAttribute(k_pileup7<false, P7_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tile7<P7_TILE>)));
    CK(cudaFuncSetAttribute(k_pileup7<true, P7_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tile7<P7_TILE>)));

    CK(cudaFuncSetAttribute(k_pileup7c<false, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tile7c<512>)));
    CK(cudaFuncSetAttribute(k_pileup7c<true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tile7c<512>)));
    CK(cudaFuncSetAttribute(k_pileup7c<false, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tile7c<1024>)));
    CK(cudaFuncSetAttribute(k_pileup7c<true, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tile7c<1024>)));
    if (const char* v = getenv("PB_CTILE")) e->ctile = atoi(v) == 1024 ? 1024 : 512;
    CK(cudaFuncSetAttribute(k_spill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SPILL_SMEM_CAP * sizeof(int4))));
    CK(cudaFuncSetAttribute(k_pileup5<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_pileup5<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->ev_stage, cudaEventDisableTiming));
    if (getenv("PB_PHASE_TIMING")) { e->phase_timing = true; for (auto& ev : e->ph) CK(cudaEventCreate(&ev)); }
    CK(cudaEventCreate(&e->ev0)); CK(cudaEventCreate(&e->ev1));
    CK(cudaEventCreate(&e->evp0)); CK(cudaEventCreate(&e->evp1));
    CK(cudaEventCreateWithFlags(&e->ev_sc, cudaEventDisableTiming));
    CK(cudaMallocHost(&e->h_sc, SC_BYTES + 16));          // + the compact-metadata flag word
    CK(e->meta_flag.ensure(16, true, e->stream));
    *out = e;
    return PB_OK;
}

extern "C" int pb_destroy(pb_engine* e) {
    if (e && e->phase_timing && e->ph_regions)
        fprintf(stderr, "pilon_b200 phases over %lld regions: upload %.2f ms, pass %.2f ms, download %.2f ms per region\n",
                (long long)e->ph_regions, e->ph_ms[0] / e->ph_regions, e->ph_ms[1] / e->ph_regions, e->ph_ms[2] / e->ph_regions);
    if (!e) return PB_OK;
    if (e->host_trace && e->ht_n)
        fprintf(stderr, "pilon_b200 pb_region_begin over %ld calls, host us per call: free_batches %.0f, reference window copy %.0f, flag read-back %.0f, buffers %.0f\n",
                e->ht_n, e->ht[0] / e->ht_n, e->ht[1] / e->ht_n, e->ht[2] / e->ht_n, e->ht[3] / e->ht_n);
    if (e->host_trace && e->hf_n)
        fprintf(stderr, "pilon_b200 pb_region_finish over %ld calls, host us per call: enqueue pass %.0f, wait pass %.0f, indel groups (copy, strings) %.0f, "
                "enqueue download %.0f, wait download %.0f, scalars + string pool %.0f, sort + fill indels %.0f, free batches %.0f\n",
                e->hf_n, e->hf[0] / e->hf_n, e->hf[1] / e->hf_n, e->hf[2] / e->hf_n, e->hf[3] / e->hf_n, e->hf[4] / e->hf_n, e->hf[5] / e->hf_n,
                e->hf[6] / e->hf_n, e->hf[7] / e->hf_n);
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    free_batches(e);
    cudaStreamSynchronize(e->stream);
    DBuf* all[] = {&e->d_batches, &e->d_pile, &e->extra, &e->head, &e->ref, &e->rare_bits, &e->pc_diff, &e->block_sums, &e->scalars,
                   &e->o_cnt, &e->o_qs, &e->o_wq, &e->o_wmq, &e->o_flags, &e->o_call, &e->ev_key, &e->ev, &e->perm,
                   &e->groups, &e->cand, &e->work, &e->spill_scratch, &e->str_pool, &e->cub_tmp, &e->call_idx, &e->call_entries, &e->meta_flag};
    for (DBuf* b : all) b->release();
    e->arena.release();
    e->rare.release(); for (auto& b : e->gplane) b.release();
    for (auto& b : e->o_i32) b.release();
    if (e->h_sc) cudaFreeHost(e->h_sc);
    cudaEventDestroy(e->ev0); cudaEventDestroy(e->ev1); cudaEventDestroy(e->evp0); cudaEventDestroy(e->evp1);
    cudaEventDestroy(e->ev_sc); cudaEventDestroy(e->ev_fork); cudaEventDestroy(e->ev_join); cudaEventDestroy(e->ev_stage);
    if (e->ref_stage) cudaFreeHost(e->ref_stage);
    cudaStreamDestroy(e->stream2);
    cudaStreamDestroy(e->stream);
    delete e;
    return PB_OK;
}

extern "C" int pb_stream(pb_engine* e, void** s) { if (!e || !s) return fail(PB_ERR_INVALID, "null"); *s = e->stream; return PB_OK; }

extern "C" int pb_region_begin(pb_engine* e, const uint8_t* contig, int64_t contig_len, int32_t start, int32_t stop) {
    if (!e || !contig) return fail(PB_ERR_INVALID, "null argument");
    if (start < 1 || stop < start || stop > contig_len) return fail(PB_ERR_INVALID, "region must satisfy 1 <= start <= stop <= contig_len");
    const int64_t S = (int64_t)stop + 1 - start;
    if (S >= (1ll << 30)) return fail(PB_ERR_INVALID, "region too large (size must be < 2^30)");
    CK(cudaSetDevice(e->device));
    cudaStream_t s = e->stream;
    auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_prev = e->host_trace ? now() : 0.0;
    auto lap = [&](int k) { if (e->host_trace) { const double t = now(); if (e->ht_calls >= 8) { e->ht[k] += t - t_prev; if (t - t_prev > 30e3) fprintf(stderr, "pilon_b200 slow: pb_region_begin section %d took %.1f ms (call %ld, size %lld)\n", k, (t - t_prev) / 1e3, e->ht_calls, (long long)S); } t_prev = t; } };   // (the first calls grow buffers)
    if (free_batches(e) != PB_OK) return PB_ERR_CUDA;
    lap(0);
    if (e->phase_timing) CK(cudaEventRecord(e->ph[0], s));
    RegionDev& R = e->R;
    R = RegionDev{};
    R.start = start; R.stop = stop; R.size = S; R.n_win = (int32_t)((S + 31) >> 5);
    R.ref_locus0 = start - PB_REF_HALO > 1 ? start - PB_REF_HALO : 1;
    R.ref_end = (int32_t)std::min<int64_t>(contig_len, (int64_t)stop + PB_REF_HALO);
    R.cfg = e->cfg;
    const size_t ref_bytes = (size_t)(R.ref_end - R.ref_locus0 + 1);
    CK(e->ref.ensure(ref_bytes + 8, false, s));                                     // k_rebuild_bases reads whole words
    {
        // A cudaMemcpyAsync from PAGEABLE memory is synchronous and waits its turn on the copy engine -- behind the other host
        // threads' batch uploads (measured: 1.5-3.8 ms per call, 20 calls per C2 step).  Unless the caller's contig is pinned,
        // the window goes through a pinned staging buffer of the engine's own: a host memcpy, then a truly asynchronous copy.
        const uint8_t* src = contig + (R.ref_locus0 - 1);
        cudaPointerAttributes at;
        const bool pinned = cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost;
        if (!pinned) {
            cudaGetLastError();
            if (ref_bytes > e->ref_stage_cap) {
                if (e->ref_stage) { CK(cudaEventSynchronize(e->ev_stage)); CK(cudaFreeHost(e->ref_stage)); e->ref_stage = nullptr; e->ref_stage_cap = 0; }
                const size_t want = ref_bytes + ref_bytes / 8 + 4096;
                CK(cudaHostAlloc((void**)&e->ref_stage, want, cudaHostAllocDefault));
                e->ref_stage_cap = want;
            } else {
                CK(cudaEventSynchronize(e->ev_stage));        // the previous region's copy out of the staging buffer is done
            }
            memcpy(e->ref_stage, src, ref_bytes);
            src = e->ref_stage;
        }
        CK(cudaMemcpyAsync(e->ref.p, src, ref_bytes, cudaMemcpyHostToDevice, s));
        if (!pinned) CK(cudaEventRecord(e->ev_stage, s));
    }
    lap(1);
    R.ref = e->ref.as<uint8_t>();
    R.contig_len = contig_len; R.extra = nullptr; R.head = nullptr; R.head_len = 0;
    e->contig_host = contig;
    if (e->meta_used) { CK(cudaMemsetAsync(e->meta_flag.p, 0, 4, s)); e->meta_used = false; }
    if (e->unverified && e->scalars.p) {              // asynchronous passes since the last read-back: did one of them raise a flag?
        CK(cudaMemcpyAsync(e->h_sc, e->scalars.p, sizeof(Scalars), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (e->h_sc->error) e->dirty = true;
    }
    e->unverified = false;
    lap(2);
    if (e->dirty) { int rcc = clean_sparse_planes(e); if (rcc != PB_OK) return rcc; e->dirty = false; }
    const size_t n4 = (size_t)S * 4;
    CK(e->rare.ensure((size_t)S * sizeof(Rare), true, s));
    for (auto& b : e->gplane) CK(b.ensure(n4, true, s));
    CK(e->rare_bits.ensure((size_t)R.n_win * 4, true, s));
    CK(e->pc_diff.ensure((size_t)S * 8, true, s));
    CK(e->scalars.ensure(SC_BYTES, false, s));
    CK(cudaMemsetAsync(e->scalars.p, 0, SC_BYTES, s));
    CK(e->o_cnt.ensure((size_t)S * 16, false, s)); CK(e->o_qs.ensure((size_t)S * 32, false, s));
    for (auto& b : e->o_i32) CK(b.ensure(n4, false, s));
    CK(e->o_wq.ensure((size_t)S, false, s)); CK(e->o_wmq.ensure((size_t)S, false, s));
    CK(e->o_flags.ensure((size_t)S, false, s)); CK(e->o_call.ensure((size_t)S * 8, false, s));
    const int nblocks = (int)((S + SCAN_TILE - 1) / SCAN_TILE);
    CK(e->block_sums.ensure((size_t)nblocks * 8, false, s));
    lap(3); if (e->ht_calls++ >= 8) e->ht_n++;
    R.sc = e->scalars.as<Scalars>();
    R.slots = reinterpret_cast<ScalarSlot*>(static_cast<uint8_t*>(e->scalars.p) + sizeof(Scalars));
    R.batch_bc = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(e->scalars.p) + SC_BC_OFF);
    R.rare = e->rare.as<Rare>();
    R.r_gins = e->gplane[0].as<uint32_t>(); R.r_gdel = e->gplane[1].as<uint32_t>();
    R.rare_bits = e->rare_bits.as<uint32_t>();
    R.pc_diff = e->pc_diff.as<int2>();
    R.o_cnt = e->o_cnt.as<int32_t>(); R.o_qs = e->o_qs.as<int64_t>();
    int32_t** o32[] = {&R.o_mq, &R.o_q, &R.o_pc, &R.o_is, &R.o_bp, &R.o_del, &R.o_delq, &R.o_ins, &R.o_insq,
                       &R.o_clips, &R.o_cov, &R.o_frag};
    for (int i = 0; i < 12; i++) *o32[i] = e->o_i32[i].as<int32_t>();
    R.o_wq = e->o_wq.as<int8_t>(); R.o_wmq = e->o_wmq.as<int8_t>();
    R.o_flags = e->o_flags.as<uint8_t>(); R.o_call = e->o_call.as<uint64_t>();
    e->in_region = true;
    return PB_OK;
}

// Make one batch array visible to the device: device pointers are used in place, host arrays are
// copied asynchronously into the engine's batch arena.
template <class T>
static int stage(pb_engine* e, HostBatch& hb, const T* src, size_t n, int mem, const T** dst) {
    if (mem == PB_MEM_DEVICE) { *dst = src; return PB_OK; }
    void* p = nullptr;
    CK(e->arena.alloc(n * sizeof(T) + 64, &p));               // kernels fetch aligned 16-byte blocks
    if (n) CK(cudaMemcpyAsync(p, src, n * sizeof(T), cudaMemcpyHostToDevice, e->stream));
    *dst = (const T*)p;
    return PB_OK;
}

// bases2 from the reference prediction + the batch's base deltas (pb_batch.base_delta_idx): thread per read; a read owns
// whole bytes of bases2 (its first base index is a multiple of 4), so no two threads touch the same byte.  Same rule as
// predicted_codes() (pb_device.cuh, which the host encoder uses), with a fast path that turns four reference bytes into
// one output byte at a time: x = (b >> 1) & 3, code = x ^ (x >> 1) maps A C G T to 0 1 2 3; any other byte gives 0.
__device__ __forceinline__ uint32_t ref_code4(uint32_t w) {          // four reference bytes -> four codes, one per byte
    const uint32_t x = (w >> 1) & 0x03030303u;
    uint32_t c = x ^ ((x >> 1) & 0x01010101u);
    const uint32_t t = w ^ 0x41414141u;                             // A C G T -> 0x00 0x02 0x06 0x15
    uint32_t keep = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t ti = (t >> (8 * i)) & 0xFFu;
        const bool ok = ti < 32u && ((0x00200045u >> ti) & 1u);
        keep |= ok ? (0xFFu << (8 * i)) : 0u;
    }
    return c & keep;
}

// The 128 reads of a block own one contiguous run of output bytes: it is assembled in shared memory (byte stores are
// cheap there) and written out with coalesced 16-byte stores; a run longer than the staging buffer (long reads) is
// written directly.
static constexpr int RB_THREADS = 128;
static constexpr uint32_t RB_STAGE = 16384;         // bytes of staging per block (128 reads of up to 512 bases)

__global__ void __launch_bounds__(RB_THREADS) k_rebuild_bases(RegionDev R, DevBatch B, const uint32_t* __restrict__ didx,
                                                              const uint8_t* __restrict__ dcode, int64_t nd, uint8_t* out) {
    __shared__ __align__(16) uint8_t stage[RB_STAGE + 16];
    const int64_t r0 = (int64_t)blockIdx.x * RB_THREADS;
    const int64_t r = r0 + threadIdx.x;
    const int64_t rl = min(r0 + RB_THREADS, B.n_reads) - 1;             // last read of the block
    const uint32_t byte0 = B.seq_off[r0] >> 2;                          // the block's run of output bytes: [byte0, byte1)
    const uint32_t byte1 = (B.seq_off[rl] >> 2) + (uint32_t)((B.read_len[rl] + 3) >> 2);
    const bool staged = byte1 - byte0 <= RB_STAGE;                      // block-uniform
    if (r < B.n_reads) {
        const uint32_t soff = B.seq_off[r];
        const int32_t L = B.read_len[r];
        const uint32_t c0 = B.cigar_off[r], c1 = B.cigar_off[r + 1];
        uint8_t* o = staged ? stage + ((soff >> 2) - byte0) : out + (soff >> 2);
        const int64_t lo = R.ref_locus0, hi = R.ref_end;
        uint32_t acc = 0; int nacc = 0, bytei = 0; int32_t done = 0;
        auto emit = [&](uint32_t code) { acc |= code << (2 * nacc); if (++nacc == 4) { o[bytei++] = (uint8_t)acc; acc = 0; nacc = 0; } };
        auto one = [&](int64_t l) -> uint32_t {
            if (l < lo || l > hi) return 0u;
            const uint8_t b = R.ref[l - lo];
            return b == 'C' ? 1u : b == 'G' ? 2u : b == 'T' ? 3u : 0u;
        };
        int64_t locus = B.pos[r];
        for (uint32_t k = c0; k < c1 && done < L; k++) {
            const uint32_t e = B.cigar[k]; const int op = (int)(e & 15); const int64_t len = (int64_t)(e >> 4);
            if (op == 0 || op == 7 || op == 8) {
                int64_t j = 0;
                for (; j < len && done < L && nacc != 0; j++, done++) emit(one(locus + j));      // up to the next output byte
                // one output byte per trip from a sliding window over aligned reference words
                int64_t l = locus + j;
                if (j + 4 <= len && done + 4 <= L && l >= lo && l + 3 <= hi) {
                    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(R.ref);         // R.ref is 256-byte aligned (cudaMalloc)
                    uint64_t wi = (uint64_t)(l - lo) >> 2;
                    const uint32_t sh = (uint32_t)((l - lo) & 3) << 3;
                    const uint64_t wlast = (uint64_t)(hi - lo) >> 2;
                    uint32_t cur = w32[wi];
                    for (; j + 4 <= len && done + 4 <= L && l + 3 <= hi; j += 4, done += 4, l += 4) {
                        const uint32_t nxt = wi + 1 <= wlast ? w32[wi + 1] : 0u;
                        const uint32_t c = ref_code4(__funnelshift_r(cur, nxt, sh));
                        o[bytei++] = (uint8_t)((c * 0x01041040u) >> 24);                      // c0 | c1 << 2 | c2 << 4 | c3 << 6
                        cur = nxt; wi++;
                    }
                }
                for (; j < len && done < L; j++, done++) emit(one(locus + j));
                locus += len;
            } else if (op == 1 || op == 4) {
                for (int64_t j = 0; j < len && done < L; j++, done++) emit(0u);
            } else if (op == 2 || op == 3) locus += len;
        }
        for (const int32_t padded = (L + 3) & ~3; done < padded; done++) emit(0u);
        const uint32_t end = soff + (uint32_t)((L + 3) & ~3);
        int64_t dl = 0, dh = nd;
        while (dl < dh) { const int64_t m = (dl + dh) >> 1; if (didx[m] < soff) dl = m + 1; else dh = m; }
        for (int64_t i = dl; i < nd && didx[i] < end; i++) {
            const uint32_t bi = didx[i] - soff, sh = 2 * (bi & 3);
            o[bi >> 2] = (uint8_t)((o[bi >> 2] & ~(3u << sh)) | ((uint32_t)dcode[i] << sh));
        }
    }
    if (!staged) return;
    __syncthreads();
    // coalesced write-out of [byte0, byte1): bytes up to the first 16-byte boundary, 16-byte stores, the rest
    const uint32_t nbytes = byte1 - byte0;
    const uint32_t head = min(nbytes, (16u - (byte0 & 15u)) & 15u);
    for (uint32_t i = threadIdx.x; i < head; i += RB_THREADS) out[byte0 + i] = stage[i];
    const uint32_t nvec = (nbytes - head) >> 4;
    for (uint32_t v = threadIdx.x; v < nvec; v += RB_THREADS) {
        const uint8_t* sp = stage + head + 16u * v;                     // (not aligned in shared memory: assembled from byte loads)
        uint4 x;
        x.x = (uint32_t)sp[0] | ((uint32_t)sp[1] << 8) | ((uint32_t)sp[2] << 16) | ((uint32_t)sp[3] << 24);
        x.y = (uint32_t)sp[4] | ((uint32_t)sp[5] << 8) | ((uint32_t)sp[6] << 16) | ((uint32_t)sp[7] << 24);
        x.z = (uint32_t)sp[8] | ((uint32_t)sp[9] << 8) | ((uint32_t)sp[10] << 16) | ((uint32_t)sp[11] << 24);
        x.w = (uint32_t)sp[12] | ((uint32_t)sp[13] << 8) | ((uint32_t)sp[14] << 16) | ((uint32_t)sp[15] << 24);
        *reinterpret_cast<uint4*>(out + byte0 + head + 16u * v) = x;
    }
    for (uint32_t i = head + 16u * nvec + threadIdx.x; i < nbytes; i += RB_THREADS) out[byte0 + i] = stage[i];
}

extern "C" void pb_free(void* p) { free(p); }

extern "C" int pb_base_delta_encode(const pb_batch* b, const uint8_t* contig, int64_t contig_len, int32_t start, int32_t stop,
                                    uint32_t** idx_out, uint8_t** code_out, int64_t* n_out) {
    if (!b || !contig || !idx_out || !code_out || !n_out) return fail(PB_ERR_INVALID, "null argument");
    if (b->mem != PB_MEM_HOST || !b->bases2) return fail(PB_ERR_INVALID, "pb_base_delta_encode needs a host batch with bases2");
    if (start < 1 || stop < start || stop > contig_len) return fail(PB_ERR_INVALID, "region must satisfy 1 <= start <= stop <= contig_len");
    const int64_t ref_lo = start - PB_REF_HALO > 1 ? start - PB_REF_HALO : 1;
    const int64_t ref_hi = std::min<int64_t>(contig_len, (int64_t)stop + PB_REF_HALO);
    const uint8_t* ref = contig + (ref_lo - 1);
    const int64_t n = b->n_reads;
    const unsigned nt = (unsigned)std::max<int64_t>(1, std::min<int64_t>(16, n / 4096));
    std::vector<std::vector<uint32_t>> vi(nt); std::vector<std::vector<uint8_t>> vc(nt);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) th.emplace_back([&, t]() {
        const int64_t r0 = n * t / nt, r1 = n * (t + 1) / nt;
        for (int64_t r = r0; r < r1; r++) {
            uint32_t bi = b->seq_off[r];
            predicted_codes(b->pos[r], b->cigar + b->cigar_off[r], b->cigar_off[r + 1] - b->cigar_off[r], b->read_len[r], ref, ref_lo, ref_hi,
                            [&](uint32_t code) {
                                const uint32_t have = (b->bases2[bi >> 2] >> (2 * (bi & 3))) & 3u;
                                if (have != code) { vi[t].push_back(bi); vc[t].push_back((uint8_t)have); }
                                bi++;
                            });
        }
    });
    for (auto& x : th) x.join();
    size_t total = 0;
    for (auto& v : vi) total += v.size();
    uint32_t* oi = (uint32_t*)malloc(total * 4 + 64); uint8_t* oc = (uint8_t*)malloc(total + 64);
    if (!oi || !oc) { free(oi); free(oc); return fail(PB_ERR_OOM, "out of host memory"); }
    size_t at = 0;
    for (unsigned t = 0; t < nt; t++) {
        if (!vi[t].empty()) { memcpy(oi + at, vi[t].data(), vi[t].size() * 4); memcpy(oc + at, vc[t].data(), vc[t].size()); }
        at += vi[t].size();
    }
    *idx_out = oi; *code_out = oc; *n_out = (int64_t)total;
    return PB_OK;
}

// Compact per-read metadata (pb_batch.meta_codes): the plain arrays are 22 B per read + 4 B per CIGAR op, 29 % of what a
// short-read batch uploads once qualities and bases travel packed; a position-sorted batch of reads with mostly "150M"
// CIGARs needs 8 B per read.
extern "C" int pb_meta_encode(const pb_batch* b, uint64_t** codes_out, uint32_t** cigar_out, int64_t* n_cigar_out,
                              int32_t** esc_out, int64_t* n_esc_out, int32_t* pos0_out, int32_t* seq_stride_out) {
    if (!b || !codes_out || !cigar_out || !n_cigar_out || !esc_out || !n_esc_out || !pos0_out || !seq_stride_out)
        return fail(PB_ERR_INVALID, "null argument");
    if (b->mem != PB_MEM_HOST) return fail(PB_ERR_INVALID, "pb_meta_encode needs a host batch");
    const int64_t n = b->n_reads;
    if (n && (!b->pos || !b->tlen || !b->read_len || !b->mapq || !b->flags || !b->cigar_off || !b->seq_off || (b->n_cigar && !b->cigar)))
        return fail(PB_ERR_INVALID, "null per-read array in pb_batch");
    // seq_off layout: cumulative padded lengths, or a constant stride
    int32_t stride = 0;
    {
        bool canonical = true, strided = n > 1;
        uint64_t acc = 0;
        const int64_t st = n > 1 ? (int64_t)b->seq_off[1] - (int64_t)b->seq_off[0] : 0;
        for (int64_t r = 0; r < n; r++) {
            if ((uint64_t)b->seq_off[r] != acc) canonical = false;
            if (strided && (int64_t)b->seq_off[r] != r * st) strided = false;
            if (b->read_len[r] < 0 || b->read_len[r] > 255) return fail(PB_ERR_UNSUPPORTED, "a read is longer than 255 bases");
            acc += (uint64_t)((b->read_len[r] + 3) & ~3);
        }
        if (!canonical) {
            if (!strided || st <= 0 || st > 0x7FFFFFFF) return fail(PB_ERR_UNSUPPORTED, "seq_off is neither cumulative nor a constant stride");
            stride = (int32_t)st;
        }
    }
    std::vector<uint32_t> cig; std::vector<int32_t> esc;
    uint64_t* codes = (uint64_t*)malloc((size_t)std::max<int64_t>(n, 1) * 8 + 64);
    if (!codes) return fail(PB_ERR_OOM, "out of host memory");
    int32_t prev = n ? b->pos[0] : 0;
    for (int64_t r = 0; r < n; r++) {
        const int64_t delta = (int64_t)b->pos[r] - prev;
        const uint32_t c0 = b->cigar_off[r], c1 = b->cigar_off[r + 1];
        if (delta < 0 || c1 < c0 || c1 - c0 > 255 || r > 0x7FFFFFFF) { free(codes); return fail(PB_ERR_UNSUPPORTED, delta < 0 ? "the batch is not sorted by pos" : "a read has more than 255 CIGAR ops"); }
        prev = b->pos[r];
        uint64_t code = 0;
        if (delta >= 0xFFFF) { code |= 0xFFFFull; esc.push_back((int32_t)r); esc.push_back(0); esc.push_back((int32_t)delta); }
        else code |= (uint64_t)delta;
        const int32_t tl = b->tlen[r];
        if (tl <= -32768 || tl > 32767) { code |= 0x8000ull << 16; esc.push_back((int32_t)r); esc.push_back(1); esc.push_back(tl); }
        else code |= (uint64_t)(uint16_t)(int16_t)tl << 16;
        code |= (uint64_t)(uint8_t)b->read_len[r] << 32 | (uint64_t)b->mapq[r] << 40 | (uint64_t)b->flags[r] << 48;
        const bool simple = c1 - c0 == 1 && b->cigar[c0] == ((uint32_t)b->read_len[r] << 4);
        if (!simple) {
            if (c1 == c0) { free(codes); return fail(PB_ERR_UNSUPPORTED, "a read has no CIGAR op"); }
            code |= (uint64_t)(c1 - c0) << 56;
            cig.insert(cig.end(), b->cigar + c0, b->cigar + c1);
        }
        codes[r] = code;
    }
    uint32_t* oc = (uint32_t*)malloc(cig.size() * 4 + 64); int32_t* oe = (int32_t*)malloc(esc.size() * 4 + 64);
    if (!oc || !oe) { free(codes); free(oc); free(oe); return fail(PB_ERR_OOM, "out of host memory"); }
    if (!cig.empty()) memcpy(oc, cig.data(), cig.size() * 4);
    if (!esc.empty()) memcpy(oe, esc.data(), esc.size() * 4);
    *codes_out = codes; *cigar_out = oc; *n_cigar_out = (int64_t)cig.size(); *esc_out = oe; *n_esc_out = (int64_t)(esc.size() / 3);
    *pos0_out = n ? b->pos[0] : 0; *seq_stride_out = stride;
    return PB_OK;
}

// ---- device side of the compact metadata: one scan over (pos delta, padded length, ops, listed ops), then a thread per read ----
namespace {
struct MetaSums { uint32_t pos, seq, ops, listed; };
struct MetaAdd { __host__ __device__ MetaSums operator()(const MetaSums& a, const MetaSums& b) const { return MetaSums{a.pos + b.pos, a.seq + b.seq, a.ops + b.ops, a.listed + b.listed}; } };
__device__ __forceinline__ bool meta_escape(const int32_t* __restrict__ esc, int64_t n_esc, int32_t r, int32_t field, int32_t* v) {
    int64_t lo = 0, hi = n_esc;
    while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (esc[3 * m] < r) lo = m + 1; else hi = m; }
    for (; lo < n_esc && esc[3 * lo] == r; lo++) if (esc[3 * lo + 1] == field) { *v = esc[3 * lo + 2]; return true; }
    return false;
}
struct MetaTerm {
    const uint64_t* codes; const int32_t* esc; int64_t n_esc;
    __device__ MetaSums operator()(int32_t r) const {
        const uint64_t c = codes[r];
        int32_t d = (int32_t)(c & 0xFFFFu);
        if (d == 0xFFFF && !meta_escape(esc, n_esc, r, 0, &d)) d = 0;
        const uint32_t len = (uint32_t)(c >> 32) & 0xFFu, nl = (uint32_t)(c >> 56);
        return MetaSums{(uint32_t)d, (len + 3u) & ~3u, nl ? nl : 1u, nl};
    }
};
// flag bits: 1 = the records describe more CIGAR ops / base slots than n_cigar / n_seq say, 2 = an escape is missing
__global__ void __launch_bounds__(256) k_meta_expand(const uint64_t* __restrict__ codes, const MetaSums* __restrict__ ex, const uint32_t* __restrict__ listed,
                                                     const int32_t* __restrict__ esc, int64_t n_esc, int64_t n, int64_t n_cigar, int64_t n_seq, int64_t n_listed,
                                                     int32_t pos0, int32_t stride, int32_t* pos, int32_t* tlen, int32_t* read_len, uint8_t* mapq, uint8_t* flags,
                                                     uint32_t* cigar_off, uint32_t* cigar, uint32_t* seq_off, int* flag) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint64_t c = codes[r];
    const MetaSums e = ex[r];
    int32_t d = (int32_t)(c & 0xFFFFu);
    if (d == 0xFFFF && !meta_escape(esc, n_esc, (int32_t)r, 0, &d)) { d = 0; atomicOr(flag, 2); }
    int32_t tl = (int32_t)(int16_t)(uint16_t)(c >> 16);
    if (tl == -32768 && !meta_escape(esc, n_esc, (int32_t)r, 1, &tl)) { tl = 0; atomicOr(flag, 2); }
    const uint32_t len = (uint32_t)(c >> 32) & 0xFFu, nl = (uint32_t)(c >> 56), nops = nl ? nl : 1u;
    pos[r] = pos0 + (int32_t)(e.pos + (uint32_t)d); tlen[r] = tl; read_len[r] = (int32_t)len;
    mapq[r] = (uint8_t)(c >> 40); flags[r] = (uint8_t)(c >> 48);
    const uint64_t so = stride > 0 ? (uint64_t)r * (uint64_t)stride : (uint64_t)e.seq;
    seq_off[r] = (uint32_t)so;
    cigar_off[r] = e.ops;
    if (r == n - 1) cigar_off[n] = e.ops + nops;
    if ((uint64_t)e.ops + nops > (uint64_t)n_cigar || so + ((len + 3u) & ~3u) > (uint64_t)n_seq || (uint64_t)e.listed + nl > (uint64_t)n_listed) {
        atomicOr(flag, 1);                                   // the pass will be refused; until then this read is empty and harmless
        read_len[r] = 0; seq_off[r] = 0;
        for (uint32_t k = 0; k < nops; k++) if ((uint64_t)e.ops + k < (uint64_t)n_cigar) cigar[e.ops + k] = 0;
        if ((uint64_t)e.ops > (uint64_t)n_cigar) cigar_off[r] = (uint32_t)n_cigar;
        if (r == n - 1) cigar_off[n] = (uint32_t)min((uint64_t)e.ops + nops, (uint64_t)n_cigar);
        return;
    }
    if (nl == 0) cigar[e.ops] = len << 4;
    else for (uint32_t k = 0; k < nl; k++) cigar[e.ops + k] = listed[e.listed + k];
}
}  // namespace

static int stage_meta(pb_engine* e, const pb_batch* b, DevBatch& d) {
    cudaStream_t s = e->stream;
    const int64_t n = b->n_reads;
    void *pc = nullptr, *pl = nullptr, *pe = nullptr, *px = nullptr, *p = nullptr;
    CK(e->arena.alloc((size_t)n * 8 + 64, &pc));
    CK(e->arena.alloc((size_t)b->n_meta_cigar * 4 + 64, &pl));
    CK(e->arena.alloc((size_t)b->n_meta_esc * 12 + 64, &pe));
    CK(e->arena.alloc((size_t)n * sizeof(MetaSums) + 64, &px));
    if (n) CK(cudaMemcpyAsync(pc, b->meta_codes, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    if (b->n_meta_cigar) CK(cudaMemcpyAsync(pl, b->meta_cigar, (size_t)b->n_meta_cigar * 4, cudaMemcpyHostToDevice, s));
    if (b->n_meta_esc) CK(cudaMemcpyAsync(pe, b->meta_esc, (size_t)b->n_meta_esc * 12, cudaMemcpyHostToDevice, s));
    CK(e->arena.alloc((size_t)n * 4 + 64, &p)); d.pos = (const int32_t*)p;
    CK(e->arena.alloc((size_t)n * 4 + 64, &p)); d.tlen = (const int32_t*)p;
    CK(e->arena.alloc((size_t)n * 4 + 64, &p)); d.read_len = (const int32_t*)p;
    CK(e->arena.alloc((size_t)n + 64, &p)); d.mapq = (const uint8_t*)p;
    CK(e->arena.alloc((size_t)n + 64, &p)); d.flags = (const uint8_t*)p;
    CK(e->arena.alloc((size_t)(n + 1) * 4 + 64, &p)); d.cigar_off = (const uint32_t*)p;
    CK(e->arena.alloc((size_t)b->n_cigar * 4 + 64, &p)); d.cigar = (const uint32_t*)p;
    CK(e->arena.alloc((size_t)n * 4 + 64, &p)); d.seq_off = (const uint32_t*)p;
    if (n == 0) { CK(cudaMemsetAsync((void*)d.cigar_off, 0, 4, s)); return PB_OK; }
    MetaTerm term{(const uint64_t*)pc, (const int32_t*)pe, b->n_meta_esc};
    auto in = thrust::make_transform_iterator(thrust::counting_iterator<int32_t>(0), term);
    size_t tmp = 0;
    CK(cub::DeviceScan::ExclusiveScan(nullptr, tmp, in, (MetaSums*)px, MetaAdd(), MetaSums{0, 0, 0, 0}, (int)n, s));
    void* pt = nullptr;
    CK(e->arena.alloc(tmp + 64, &pt));
    CK(cub::DeviceScan::ExclusiveScan(pt, tmp, in, (MetaSums*)px, MetaAdd(), MetaSums{0, 0, 0, 0}, (int)n, s));
    k_meta_expand<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const uint64_t*)pc, (const MetaSums*)px, (const uint32_t*)pl, (const int32_t*)pe, b->n_meta_esc, n,
                                                              b->n_cigar, b->n_seq, b->n_meta_cigar, b->meta_pos0, b->meta_seq_stride,
                                                              (int32_t*)d.pos, (int32_t*)d.tlen, (int32_t*)d.read_len, (uint8_t*)d.mapq, (uint8_t*)d.flags,
                                                              (uint32_t*)d.cigar_off, (uint32_t*)d.cigar, (uint32_t*)d.seq_off, e->meta_flag.as<int>());
    e->launches += 3;
    e->meta_used = true;
    return PB_OK;
}

// 4-bit quality codes -> quality bytes (pb_batch.qual_codes, qual_code_bits == 4).  A thread expands 16 bases: the low / high half of an input
// word is directly a PRMT selector (one code per nibble); codes 0..7 and 8..15 come from two 8-byte pools, bit 3 picks.
__global__ void __launch_bounds__(256) k_unpack_quals4(const uint2* __restrict__ in, uint4* __restrict__ out, size_t groups,
                                                       uint32_t l0, uint32_t l1, uint32_t l2, uint32_t l3) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= groups) return;
    const uint2 w = in[i];
    auto four = [&](uint32_t sel16) {
        const uint32_t lo = __byte_perm(l0, l1, sel16 & 0x7777u), hi = __byte_perm(l2, l3, sel16 & 0x7777u);
        const uint32_t m = __byte_perm(0x0000FF00u, 0u, (sel16 >> 3) & 0x1111u);          // 0xFF where the code is >= 8
        return (hi & m) | (lo & ~m);
    };
    out[i] = make_uint4(four(w.x & 0xFFFFu), four(w.x >> 16), four(w.y & 0xFFFFu), four(w.y >> 16));
}

// 3-bit quality codes: a thread expands 32 bases = 96 bits = three input words into two 16-byte outputs.  Four codes
// (12 bits) are spread into the four nibbles of a PRMT selector over the 8-entry table.
__global__ void __launch_bounds__(256) k_unpack_quals3(const uint32_t* __restrict__ in, uint4* __restrict__ out, size_t groups,
                                                       uint32_t l0, uint32_t l1) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= groups) return;
    const uint32_t w0 = in[3 * i], w1 = in[3 * i + 1], w2 = in[3 * i + 2];
    auto four = [&](uint32_t x) {                                    // x: 12 bits = 4 codes
        const uint32_t sel = (x & 0x7u) | ((x & 0x38u) << 1) | ((x & 0x1C0u) << 2) | ((x & 0xE00u) << 3);
        return __byte_perm(l0, l1, sel);
    };
    out[2 * i] = make_uint4(four(w0 & 0xFFFu), four((w0 >> 12) & 0xFFFu), four(__funnelshift_r(w0, w1, 24) & 0xFFFu), four((w1 >> 4) & 0xFFFu));
    out[2 * i + 1] = make_uint4(four((w1 >> 16) & 0xFFFu), four(__funnelshift_r(w1, w2, 28) & 0xFFFu), four((w2 >> 8) & 0xFFFu), four((w2 >> 20) & 0xFFFu));
}

static int stage_batch(pb_engine* e, const pb_batch* b, int frag, int long_read_type);

extern "C" int pb_region_add_batch(pb_engine* e, const pb_batch* b, int frag, int long_read_type) {
    if (!e || !b) return fail(PB_ERR_INVALID, "null argument");
    if (!e->in_region) return fail(PB_ERR_INVALID, "pb_region_begin has not been called");
    if (long_read_type < 0 || long_read_type > 2) return fail(PB_ERR_INVALID, "long_read_type must be 0, 1 (nanopore) or 2 (pacbio)");
    if (b->n_seq >= (1ll << 32) || (b->n_seq & 3)) return fail(PB_ERR_INVALID, "n_seq must be a multiple of 4 and < 2^32");
    if (b->n_cigar >= (1ll << 32) || b->n_reads >= (1ll << 31)) return fail(PB_ERR_INVALID, "batch too large");
    if (b->n_reads < 0 || b->n_cigar < 0 || b->n_seq < 0 || b->n_exc < 0) return fail(PB_ERR_INVALID, "negative count in pb_batch");
    if (e->batches.size() >= MAX_BATCHES) return fail(PB_ERR_INVALID, "too many batches in one region");
    if (b->mem != PB_MEM_HOST && b->mem != PB_MEM_DEVICE) return fail(PB_ERR_INVALID, "pb_batch.mem must be PB_MEM_HOST or PB_MEM_DEVICE");
    const bool deltas = b->mem == PB_MEM_HOST && b->base_delta_idx, qcodes = b->mem == PB_MEM_HOST && b->qual_codes;
    if (deltas && (b->n_base_delta < 0 || (b->n_base_delta && !b->base_delta_code))) return fail(PB_ERR_INVALID, "bad base delta arrays");
    if (!deltas && !b->bases2 && b->n_seq) return fail(PB_ERR_INVALID, "batch has neither bases2 nor base deltas");
    if (qcodes && b->qual_code_bits != 3 && b->qual_code_bits != 4) return fail(PB_ERR_INVALID, "qual_code_bits must be 3 or 4 when qual_codes is given");
    if (!qcodes && !b->quals && b->n_seq) return fail(PB_ERR_INVALID, "batch has neither quals nor qual_codes");
    const bool meta = b->mem == PB_MEM_HOST && b->meta_codes;
    if (meta && (b->n_meta_cigar < 0 || b->n_meta_esc < 0 || (b->n_meta_cigar && !b->meta_cigar) || (b->n_meta_esc && !b->meta_esc) || b->meta_seq_stride < 0))
        return fail(PB_ERR_INVALID, "bad compact metadata arrays");
    if (!meta && b->n_reads && (!b->pos || !b->tlen || !b->read_len || !b->mapq || !b->flags || !b->cigar_off || !b->seq_off))
        return fail(PB_ERR_INVALID, "null per-read array in pb_batch");
    if ((!meta && b->n_cigar && !b->cigar) || (b->n_exc && (!b->exc_idx || !b->exc_base || !b->exc_qual))) return fail(PB_ERR_INVALID, "null array in pb_batch");
    CK(cudaSetDevice(e->device));
    const double t_ab = e->host_trace ? std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count() : 0.0;
    drop_graph(e);                                   // the captured pass belongs to the previous batch set
    // every argument has been validated: from here on only CUDA calls can fail, and then the half-staged batch is withdrawn
    e->batches.emplace_back();
    const int rc_stage = stage_batch(e, b, frag, long_read_type);
    if (rc_stage != PB_OK) e->batches.pop_back();      // (its arena space comes back with the region's)
    if (e->host_trace) {
        const double dt = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count() - t_ab;
        if (dt > 30e3 && e->ht_calls >= 8) fprintf(stderr, "pilon_b200 slow: pb_region_add_batch took %.1f ms (%lld reads)\n", dt / 1e3, (long long)b->n_reads);
    }
    return rc_stage;
}

static int stage_meta(pb_engine* e, const pb_batch* b, DevBatch& d);
static int stage_batch(pb_engine* e, const pb_batch* b, int frag, int long_read_type) {
    HostBatch& hb = e->batches.back();
    if (long_read_type != 0) {
        // long-read batches are counted by k_long into the Extra plane (zero between regions, like the rare planes); the
        // reference's homoRun / nanoporeExclude index the contig with REGION indices, so a region that does not start at
        // locus 1 also needs the head of the contig on the device
        RegionDev& R = e->R;
        CK(e->extra.ensure((size_t)R.size * sizeof(Extra), true, e->stream));
        R.extra = e->extra.as<Extra>();
        if (R.start > 1 && !R.head) {
            const size_t hl = (size_t)std::min<int64_t>(R.contig_len, R.size + 8);
            CK(e->head.ensure(hl + 8, false, e->stream));
            CK(cudaMemcpyAsync(e->head.p, e->contig_host, hl, cudaMemcpyHostToDevice, e->stream));
            R.head = e->head.as<uint8_t>(); R.head_len = (int64_t)hl;
        }
    }
    DevBatch& d = hb.d;
    memset(&d, 0, sizeof(d));
    d.n_reads = b->n_reads; d.n_cigar = b->n_cigar; d.n_seq = b->n_seq; d.n_exc = b->n_exc; d.frag = frag ? 1 : 0;
    d.long_read = long_read_type;
    const size_t n = (size_t)b->n_reads;
    int rc;
#define ST(field, count) if ((rc = stage(e, hb, b->field, (size_t)(count), b->mem, &d.field)) != PB_OK) return rc
    if (b->mem == PB_MEM_HOST && b->meta_codes) {       // compact transport of the eight per-read arrays, rebuilt on the device
        if ((rc = stage_meta(e, b, d)) != PB_OK) return rc;
    } else {
        ST(pos, n); ST(tlen, n); ST(read_len, n); ST(mapq, n); ST(flags, n); ST(cigar_off, n + 1);
        ST(cigar, b->n_cigar); ST(seq_off, n);
    }
    if (b->mem == PB_MEM_HOST && b->base_delta_idx) {   // compact transport: upload the deltas, rebuild bases2 on the device
        void *pi = nullptr, *pc = nullptr, *pout = nullptr;
        const size_t nd = (size_t)b->n_base_delta;
        CK(e->arena.alloc(nd * 4 + 64, &pi));
        CK(e->arena.alloc(nd + 64, &pc));
        CK(e->arena.alloc((size_t)b->n_seq / 4 + 64, &pout));
        if (nd) {
            CK(cudaMemcpyAsync(pi, b->base_delta_idx, nd * 4, cudaMemcpyHostToDevice, e->stream));
            CK(cudaMemcpyAsync(pc, b->base_delta_code, nd, cudaMemcpyHostToDevice, e->stream));
        }
        d.bases2 = (const uint8_t*)pout;
        if (n) {
            k_rebuild_bases<<<(unsigned)((n + RB_THREADS - 1) / RB_THREADS), RB_THREADS, 0, e->stream>>>(e->R, d, (const uint32_t*)pi, (const uint8_t*)pc, (int64_t)nd, (uint8_t*)pout);
            e->launches++;
        }
    } else {
        ST(bases2, b->n_seq / 4);
    }
    if (b->mem == PB_MEM_HOST && b->qual_codes) {      // compact transport: upload 3- / 4-bit codes, expand on the device
        const int bits = b->qual_code_bits;
        const size_t per = bits == 4 ? 16 : 32;                       // bases expanded by one thread
        const size_t groups = ((size_t)b->n_seq + per - 1) / per;
        const size_t in_bytes = ((size_t)b->n_seq * bits + 7) / 8;
        void *pin = nullptr, *pout = nullptr;
        CK(e->arena.alloc(groups * (bits == 4 ? 8 : 12) + 64, &pin));
        CK(e->arena.alloc(groups * per + 64, &pout));
        if (b->n_seq) {
            CK(cudaMemcpyAsync(pin, b->qual_codes, in_bytes, cudaMemcpyHostToDevice, e->stream));
            uint32_t l[4]; memcpy(l, b->qual_lut, 16);
            if (bits == 4) k_unpack_quals4<<<(unsigned)((groups + 255) / 256), 256, 0, e->stream>>>((const uint2*)pin, (uint4*)pout, groups, l[0], l[1], l[2], l[3]);
            else k_unpack_quals3<<<(unsigned)((groups + 255) / 256), 256, 0, e->stream>>>((const uint32_t*)pin, (uint4*)pout, groups, l[0], l[1]);
            e->launches++;
        }
        d.quals = (const uint8_t*)pout;
    } else {
        ST(quals, b->n_seq);
    }
    ST(exc_idx, b->n_exc); ST(exc_base, b->n_exc); ST(exc_qual, b->n_exc);
#undef ST
    void* p = nullptr;
    CK(e->arena.alloc(((size_t)b->n_cigar + 1) * sizeof(Seg), &p)); d.seg = (Seg*)p;
    CK(e->arena.alloc(((size_t)e->R.n_win + 2) * 4, &p)); d.win_first = (uint32_t*)p;
    CK(e->arena.alloc((n + 1) * 4, &p)); d.insert_out = (int32_t*)p;
    d.reach = reinterpret_cast<int32_t*>(static_cast<uint8_t*>(e->scalars.p) + SC_REACH_OFF) + 2 * (e->batches.size() - 1);
    return PB_OK;
}

// ---------------------------------------------------------------------------------------------
// the compute pipeline (no host<->device data copies except one 80-byte scalar read-back)
// ---------------------------------------------------------------------------------------------

static int clean_sparse_planes(pb_engine* e) {   // a failed / abandoned pass may have left the sparse planes populated
    cudaStream_t s = e->stream;
    if (e->rare.p) CK(cudaMemsetAsync(e->rare.p, 0, e->rare.cap, s));
    for (auto& b : e->gplane) if (b.p) CK(cudaMemsetAsync(b.p, 0, b.cap, s));
    if (e->rare_bits.p) CK(cudaMemsetAsync(e->rare_bits.p, 0, e->rare_bits.cap, s));
    if (e->pc_diff.p) CK(cudaMemsetAsync(e->pc_diff.p, 0, e->pc_diff.cap, s));
    if (e->extra.p) CK(cudaMemsetAsync(e->extra.p, 0, e->extra.cap, s));
    return PB_OK;
}

// error flags raised by the kernels of a pass (read back by the caller once the pass has been synchronised)
static int pass_error(uint32_t err) {
    if (err & 1) return fail(PB_ERR_UNSORTED, "a batch is not sorted by pos");
    if (err & 8) return fail(PB_ERR_CUDA, "pipeline barrier timed out (internal protocol error)");
    if (err & 16) return fail(PB_ERR_UNSUPPORTED, "read longer than 2^24 bases or more than 256 batches in a region");
    if (err & 4) return fail(PB_ERR_HASH, "two different long insertions share a 63-bit hash");
    if (err) return fail(PB_ERR_CUDA, "internal capacity error flag set by a kernel");
    return PB_OK;
}

// No host round trip inside: everything the later kernels need from the earlier ones (event count, reach of the
// segments, read count, minDepth) stays on the device, so the whole pass can be enqueued at once (and captured in a
// CUDA graph, pb_region_compute).  The indel event list is sized on the host by an upper bound -- CIGAR operations
// minus reads: every read has at least one operation that is not I or D -- and a read set that breaks that bound
// (I/D-only CIGARs) raises the capacity flag, which makes the callers repeat the pass with `full_cap`.
static int compute(pb_engine* e, bool time_pileup, bool full_cap = false) {
    cudaStream_t s = e->stream;
    RegionDev& R = e->R;
    const int nb = (int)e->batches.size();
    size_t total_cigar = 0, total_reads = 0, total_seq = 0;
    for (auto& hb : e->batches) { total_cigar += (size_t)hb.d.n_cigar; total_reads += (size_t)hb.d.n_reads; total_seq += (size_t)hb.d.n_seq; }
    const uint32_t cap = (uint32_t)std::min<size_t>(total_cigar, 0xFFFFFFF0u);
    const uint32_t evcap = full_cap ? cap : (uint32_t)std::min<size_t>((size_t)cap, total_cigar - std::min(total_cigar, total_reads) + 64);
    CK(e->ev_key.ensure((size_t)cap * sizeof(EventKey) + 16, false, s));
    CK(e->ev.ensure((size_t)cap * sizeof(Event) + 16, false, s));
    CK(e->perm.ensure((size_t)cap * 4 + 16, false, s));
    CK(e->sort_buf.ensure((size_t)cap * 12 + 64, false, s));       // radix sort: keys in, keys out, indices out
    CK(e->groups.ensure((size_t)cap * sizeof(Group) + 16, false, s));
    CK(e->cand.ensure((size_t)cap * sizeof(int4) + 16, false, s));
    CK(e->work.ensure((size_t)cap * sizeof(int4) + 16, false, s));
    R.ev_key = e->ev_key.as<EventKey>(); R.ev = e->ev.as<Event>(); R.ev_cap = evcap;
    R.groups = e->groups.as<Group>(); R.groups_cap = cap;
    R.cand = e->cand.as<int4>(); R.cand_cap = cap;
    R.work = e->work.as<int4>(); R.work_cap = evcap;
    R.work_slots = full_cap ? 1u : (uint32_t)SC_SLOTS;           // the retry pass must not fail on an unlucky spread
    R.work_sub = std::max<uint32_t>(1u, cap / R.work_slots);
    R.str_pool = e->str_pool.as<uint8_t>(); R.str_cap = e->str_pool.cap;
    std::vector<DevBatch>& img = e->img_host;
    img.resize(nb);
    for (int i = 0; i < nb; i++) img[i] = e->batches[i].d;
    CK(e->d_batches.ensure(sizeof(DevBatch) * (size_t)std::max(nb, 1), false, s));
    if (nb) CK(cudaMemcpyAsync(e->d_batches.p, img.data(), sizeof(DevBatch) * nb, cudaMemcpyHostToDevice, s));
    const DevBatch* dB = e->d_batches.as<DevBatch>();
    e->dirty = true;
    CK(cudaMemsetAsync(e->scalars.p, 0, SC_BYTES, s));

    if (const char* xf = getenv("PB_EXP")) R.exp_flags = atoi(xf);     // knock-out experiments (tools/knockout_timing.py)
    int32_t* reach_base = reinterpret_cast<int32_t*>(static_cast<uint8_t*>(e->scalars.p) + SC_REACH_OFF);
    if (nb == 0) { k_fold<<<1, 32, 0, s>>>(R, reach_base, 0, 1, 0); e->launches++; }
    for (int i0 = 0; i0 < nb; i0 += 8) {     // k_prep folds its block partials into slots that carry 8 batches' reach
        const int i1 = std::min(nb, i0 + 8);
        for (int i = i0; i < i1; i++) {
            const DevBatch& d = e->batches[i].d;
            CK(cudaMemsetAsync(d.win_first, 0, ((size_t)R.n_win + 2) * 4, s));
            if (d.n_reads == 0) continue;
            k_prep<<<(unsigned)((d.n_reads + 127) / 128), 128, 0, s>>>(R, d, (uint32_t)i);      // also builds win_first
            e->launches += 1;
        }
        if (i1 == nb) {
            // every k_prep has run: the physCov scans only need its difference array, so they go to the side stream and
            // run beside k_indel -> k_fold -> sort -> k_groups -> pileup (joined before the deletion spill)
            CK(cudaEventRecord(e->ev_fork, s));
            CK(cudaStreamWaitEvent(e->stream2, e->ev_fork, 0));
            const int nblocks = (int)((R.size + SCAN_TILE - 1) / SCAN_TILE);
            k_scan1<<<nblocks, SCAN_THREADS, 0, e->stream2>>>(R, e->block_sums.as<uint2>());
            k_scan2<<<1, 1024, 0, e->stream2>>>(R, e->block_sums.as<uint2>(), nblocks);
            k_scan3<<<nblocks, SCAN_THREADS, 0, e->stream2>>>(R, e->block_sums.as<uint2>());
            e->launches += 3;
            CK(cudaEventRecord(e->ev_join, e->stream2));
            k_indel<<<148 * 12, 128, 0, s>>>(R, dB); e->launches++;    // queued I / D ops of every batch; a chain of dependent loads per op: as many threads in flight as 40 registers allow
        }
        k_fold<<<1, 32, 0, s>>>(R, reach_base + 2 * i0, i1 - i0, i1 == nb, nb); e->launches++;
    }
    const bool forked = nb > 0;
    if (!forked) {              // no batch at all: the scans still have to produce (zero) planes
        const int nblocks = (int)((R.size + SCAN_TILE - 1) / SCAN_TILE);
        k_scan1<<<nblocks, SCAN_THREADS, 0, s>>>(R, e->block_sums.as<uint2>());
        k_scan2<<<1, 1024, 0, s>>>(R, e->block_sums.as<uint2>(), nblocks);
        k_scan3<<<nblocks, SCAN_THREADS, 0, s>>>(R, e->block_sums.as<uint2>());
        e->launches += 3;
    }
    if (evcap) {
        // events -> (locus, kind) groups: radix sort of `evcap` 32-bit keys (the unused slots carry the largest key), k_groups
        uint32_t* keys_in = e->sort_buf.as<uint32_t>();
        uint32_t* keys_out = keys_in + cap + 4;
        uint32_t* idx_out = keys_out + cap + 4;
        uint32_t* idx_in = e->perm.as<uint32_t>();
        k_event_keys<<<(evcap + 255) / 256, 256, 0, s>>>(R, keys_in, idx_in, evcap); e->launches++;
        int end_bit = 1;
        while (end_bit < 32 && (((uint64_t)R.size << 1) >> end_bit)) end_bit++;
        if (end_bit < 32) end_bit++;                                     // the padding key's extra bit
        size_t tmp = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, keys_in, keys_out, idx_in, idx_out, (int64_t)evcap, 0, end_bit, s));
        CK(e->cub_tmp.ensure(tmp + 16, false, s));
        CK(cub::DeviceRadixSort::SortPairs(e->cub_tmp.p, tmp, keys_in, keys_out, idx_in, idx_out, (int64_t)evcap, 0, end_bit, s));
        k_groups<<<(evcap + 255) / 256, 256, 0, s>>>(R, dB, keys_out, idx_out, evcap); e->launches++;
    }
    if (const char* xf = getenv("PB_EXP")) R.exp_flags = atoi(xf);
    if (time_pileup) CK(cudaEventRecord(e->evp0, s));
    // Three formulations of the same arithmetic: pile-ups deeper than ~1000x go to the cluster scatter kernel
    // (wide counters in shared memory, a cluster per 512-locus tile); shallower regions with enough 2048-locus tiles to
    // fill the GPU to the scatter kernel; narrow shallow ones to the gather kernel.
    int pv = e->pileup_version;
    const int64_t depth = R.size > 0 ? (int64_t)(total_seq / (size_t)R.size) : 0;          // stored bases per locus: >= depth
    if (pv == 0) {
        const int64_t tiles = (R.size + P7_TILE - 1) / P7_TILE;
        pv = depth > e->deep_depth ? 8 : tiles >= 256 ? 7 : 5;
    }
    if ((pv == 7 || pv == 8) && nb > PB_MAXB) pv = 5;   // the scatter kernels keep their per-batch cursors in shared memory: <= PB_MAXB batches
    PileBatches PBt; memset(&PBt, 0, sizeof(PBt)); PBt.n = nb;
    PBt.spread = e->spread_order >= 0 ? e->spread_order : (depth > e->spread_depth ? 1 : 0);
    {
        std::vector<PileBatch>& pile = e->pile_host;
        pile.resize((size_t)std::max(nb, 1));
        for (int i = 0; i < nb; i++) {
            const DevBatch& d = img[i];
            PileBatch& pb_ = pile[i];
            pb_.seg = d.seg; pb_.quals = d.quals; pb_.bases2 = d.bases2; pb_.win_first = d.win_first;
            pb_.n_cigar = (uint32_t)d.n_cigar; pb_.reach = d.reach;
            pb_.flags = (d.frag ? 1u : 0u) | ((d.n_reads && !d.long_read) ? 2u : 0u);     // long-read batches: k_long, not the tile kernels
            if (i < PB_MAXB) PBt.b[i] = pb_;
        }
        if (nb > PB_MAXB) {                // more BAMs than the by-value table holds: the kernel reads a device-side table
            CK(e->d_pile.ensure(sizeof(PileBatch) * (size_t)nb, false, s));
            CK(cudaMemcpyAsync(e->d_pile.p, pile.data(), sizeof(PileBatch) * (size_t)nb, cudaMemcpyHostToDevice, s));
            PBt.ext = e->d_pile.as<PileBatch>();
        }
    }
    for (int i = 0; i < nb; i++) {         // long-read batches: plain global-atomic accumulation, merged by the epilogue
        const DevBatch& d = img[i];
        if (d.long_read && d.n_cigar) { k_long<<<148 * 4, 256, 0, s>>>(R, d); e->launches++; }
    }
    if (R.exp_flags & 256) {               // knock-out: everything but the pileup kernel (what the rest of the pass costs)
    } else if (pv == 7) {
        const unsigned grid = (unsigned)((R.n_win * 32 + P7_TILE - 1) / P7_TILE);
        if (e->cfg.min_qual > 0) k_pileup7<true, P7_TILE><<<grid, P7_WARPS * 32, sizeof(Tile7<P7_TILE>), s>>>(R, PBt);
        else k_pileup7<false, P7_TILE><<<grid, P7_WARPS * 32, sizeof(Tile7<P7_TILE>), s>>>(R, PBt);
    } else if (pv == 8) {
        // a cluster per 512-locus tile; the cluster grows until the grid fills the GPU's 148 x 2 CTA slots a few times over
        const unsigned T = (unsigned)e->ctile;
        const unsigned tiles = (unsigned)((R.n_win * 32 + T - 1) / T);
        unsigned cl = 1;
        if (const char* cs = getenv("PB_CLUSTER")) cl = (unsigned)atoi(cs);
        else while (cl < (unsigned)P7C_MAX_CLUSTER && (uint64_t)tiles * cl < 148u * 2u * 4u) cl <<= 1;
        if (cl < 1 || cl > (unsigned)P7C_MAX_CLUSTER || (cl & (cl - 1))) cl = P7C_MAX_CLUSTER;
        cudaLaunchConfig_t lc; memset(&lc, 0, sizeof(lc));
        lc.gridDim = dim3(tiles * cl); lc.blockDim = dim3(P7_WARPS * 32); lc.stream = s;
        lc.dynamicSmemBytes = T == 1024 ? sizeof(Tile7c<1024>) : sizeof(Tile7c<512>);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        if (tiles) {
            const bool mq = e->cfg.min_qual > 0;
            if (T == 1024) { if (mq) CK(cudaLaunchKernelEx(&lc, k_pileup7c<true, 1024>, R, PBt)); else CK(cudaLaunchKernelEx(&lc, k_pileup7c<false, 1024>, R, PBt)); }
            else { if (mq) CK(cudaLaunchKernelEx(&lc, k_pileup7c<true, 512>, R, PBt)); else CK(cudaLaunchKernelEx(&lc, k_pileup7c<false, 512>, R, PBt)); }
        }
    } else {
        const unsigned grid = (unsigned)((R.n_win + P5_WARPS - 1) / P5_WARPS);
        size_t smem = sizeof(Warp5) * P5_WARPS;
        if (const char* sp = getenv("PB_SMEM_PAD")) smem += (size_t)atoi(sp);    // occupancy experiments
        if (e->cfg.min_qual > 0) k_pileup5<true><<<grid, P5_WARPS * 32, smem, s>>>(R, PBt);
        else k_pileup5<false><<<grid, P5_WARPS * 32, smem, s>>>(R, PBt);
    }
    e->launches++;
    if (time_pileup) CK(cudaEventRecord(e->evp1, s));
    if (forked) CK(cudaStreamWaitEvent(s, e->ev_join, 0));          // join: the scans are part of the pass
    // deletion spill: candidates are bounded by the number of deletion groups
    uint32_t p2 = 1; while (p2 < evcap) p2 <<= 1;
    CK(e->spill_scratch.ensure((size_t)p2 * sizeof(int4) + 16, false, s));
    if (evcap) { k_spill<<<SPILL_CTAS, 1024, (size_t)std::min(p2, SPILL_SMEM_CAP) * sizeof(int4), s>>>(R, e->spill_scratch.as<int4>(), p2); e->launches++; }
    CK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_region_compute(pb_engine* e) {
    if (!e || !e->in_region) return fail(PB_ERR_INVALID, "no region");
    CK(cudaSetDevice(e->device));
    // plain launches under Nsight Compute (it sets NV_COMPUTE_PROFILER_PERFWORKS_DIR): capturing on several host threads
    // while the profiler serialises kernels crashed the tool, and a launch list wants to see every kernel anyway
    static const bool use_graph = getenv("PB_NOGRAPH") == nullptr && getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") == nullptr;
    if (use_graph && e->graph_state == 2) {                  // the pass has no host round trip: one call replays it
        CK(cudaGraphLaunch(e->graph_exec, e->stream));
        e->launches += e->graph_launches;
        e->dirty = false; e->unverified = true;              // a pass without error flags leaves the sparse planes zero
        return PB_OK;
    }
    const bool capture = use_graph && e->graph_state == 1;
    const int64_t l0 = e->launches;
    if (capture) CK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    int rc = compute(e, false);
    // the group count is only known on the device here: clear through the whole capacity-bounded list
    if (rc == PB_OK) { k_groups_clear_dev<<<64, 128, 0, e->stream>>>(e->R); e->launches++; }
    if (capture) {
        cudaGraph_t g = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
        if (rc != PB_OK || ce != cudaSuccess || !g) {        // something in the pass cannot be captured: stay with plain launches
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            e->graph_state = 3;
            e->launches = l0;
            return pb_region_compute(e);
        }
        const cudaError_t ci = cudaGraphInstantiate(&e->graph_exec, g, 0);
        cudaGraphDestroy(g);
        if (ci != cudaSuccess) { cudaGetLastError(); e->graph_exec = nullptr; e->graph_state = 3; e->launches = l0; return pb_region_compute(e); }
        e->graph_launches = e->launches - l0;
        e->graph_state = 2;
        CK(cudaGraphLaunch(e->graph_exec, e->stream));
        e->dirty = false; e->unverified = true;
        return PB_OK;
    }
    if (rc != PB_OK) return rc;
    e->dirty = false; e->unverified = true;
    if (use_graph && e->graph_state == 0) e->graph_state = 1;      // buffers have their final size now
    return PB_OK;
}

extern "C" int pb_region_compute_timed(pb_engine* e, int iters, float* total_ms, float* pileup_ms, int64_t* launches) {
    if (!e || !e->in_region) return fail(PB_ERR_INVALID, "no region");
    CK(cudaSetDevice(e->device));
    const int64_t l0 = e->launches;
    float tot = 0.f, pil = 0.f;
    static const bool use_graph = getenv("PB_NOGRAPH") == nullptr && getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") == nullptr;
    auto check_pass = [&]() -> int {
        CK(cudaMemcpyAsync(e->h_sc, e->scalars.p, sizeof(Scalars), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        e->unverified = false;
        if (e->h_sc->error) { e->dirty = true; return pass_error((uint32_t)e->h_sc->error); }
        return PB_OK;
    };
    for (int it = 0; it < iters; it++) {
        // (1) plain launches with events around the pileup kernel (events recorded inside a captured graph cannot be
        // timed): the kernel's own duration, and the pass's when no graph is in use
        CK(cudaEventRecord(e->ev0, e->stream));
        int rc = compute(e, true);
        if (rc != PB_OK) return rc;
        k_groups_clear_dev<<<64, 128, 0, e->stream>>>(e->R); e->launches++;
        CK(cudaEventRecord(e->ev1, e->stream));
        if ((rc = check_pass()) != PB_OK) return rc;
        e->dirty = false;
        float a = 0, b = 0;
        CK(cudaEventElapsedTime(&a, e->ev0, e->ev1));
        CK(cudaEventElapsedTime(&b, e->evp0, e->evp1));
        pil += b;
        // (2) the pass the way pb_region_compute runs it -- replayed as one CUDA graph once it has been captured: its duration
        if (use_graph) {
            while (e->graph_state < 2) { if ((rc = pb_region_compute(e)) != PB_OK) return rc; if ((rc = check_pass()) != PB_OK) return rc; }
            if (e->graph_state == 2) {
                CK(cudaEventRecord(e->ev0, e->stream));
                if ((rc = pb_region_compute(e)) != PB_OK) return rc;
                CK(cudaEventRecord(e->ev1, e->stream));
                if ((rc = check_pass()) != PB_OK) return rc;
                CK(cudaEventElapsedTime(&a, e->ev0, e->ev1));
            }
        }
        tot += a;
    }
    if (total_ms) *total_ms = tot;
    if (pileup_ms) *pileup_ms = pil;
    if (launches) *launches = e->launches - l0;
    return PB_OK;
}

// pb_region_result.calls: the loci whose call changes or questions the reference, in locus order -- an ordered stream
// compaction over the flags plane (1 byte per locus), then a gather of (locus, flags, call).  The count goes into the
// region scalars and comes back with them: no extra synchronisation.
namespace {
struct ChangedOrAmbiguous {
    const uint8_t* fl;
    __device__ bool operator()(int32_t i) const { return (fl[i] & (PB_FL_CHANGED | PB_FL_AMBIGUOUS)) != 0; }
};
__global__ void __launch_bounds__(256) k_call_entries(RegionDev R, const int32_t* __restrict__ idx, pb_call_entry* __restrict__ out, int64_t cap) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = R.sc->n_calls < cap ? R.sc->n_calls : cap;
    if (t >= n) return;
    const int32_t i = idx[t];
    pb_call_entry en; en.locus_index = i; en.flags = R.o_flags[i]; en.call = R.o_call[i];
    out[t] = en;
}
}  // namespace

static int select_calls(pb_engine* e, int64_t cap) {
    cudaStream_t s = e->stream;
    RegionDev& R = e->R;
    const int S = (int)R.size;
    CK(e->call_idx.ensure((size_t)S * 4 + 16, false, s));
    CK(e->call_entries.ensure((size_t)cap * sizeof(pb_call_entry) + 16, false, s));
    thrust::counting_iterator<int32_t> it(0);
    ChangedOrAmbiguous pred{R.o_flags};
    size_t tmp = 0;
    CK(cub::DeviceSelect::If(nullptr, tmp, it, e->call_idx.as<int32_t>(), &R.sc->n_calls, S, pred, s));
    CK(e->cub_tmp.ensure(tmp + 16, false, s));
    CK(cub::DeviceSelect::If(e->cub_tmp.p, tmp, it, e->call_idx.as<int32_t>(), &R.sc->n_calls, S, pred, s));
    if (cap) k_call_entries<<<(unsigned)((cap + 255) / 256), 256, 0, s>>>(R, e->call_idx.as<int32_t>(), e->call_entries.as<pb_call_entry>(), cap);
    e->launches += 3;
    return PB_OK;
}

extern "C" int pb_region_finish(pb_engine* e, pb_region_result* res, int32_t* const* insert_sizes_out) {
    if (!e || !res) return fail(PB_ERR_INVALID, "null argument");
    if (!e->in_region) return fail(PB_ERR_INVALID, "pb_region_begin has not been called");
    CK(cudaSetDevice(e->device));
    cudaStream_t s = e->stream;
    RegionDev& R = e->R;
    if (res->calls_cap < 0) return fail(PB_ERR_INVALID, "calls_cap < 0");
    auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_prev = e->host_trace ? now() : 0.0;
    auto lap = [&](int k) { if (e->host_trace) { const double t = now(); if (e->hf_calls >= 8) { e->hf[k] += t - t_prev; if (t - t_prev > 60e3) fprintf(stderr, "pilon_b200 slow: pb_region_finish section %d took %.1f ms (call %ld, size %lld)\n", k, (t - t_prev) / 1e3, e->hf_calls, (long long)R.size); } t_prev = t; } };
    const int64_t calls_cap = res->calls ? std::min<int64_t>(res->calls_cap, R.size) : 0;
    for (int attempt = 0;; attempt++) {
        if (e->phase_timing && attempt == 0) CK(cudaEventRecord(e->ph[1], s));
        int rc = compute(e, false, attempt > 0);
        if (rc == PB_OK && calls_cap) rc = select_calls(e, calls_cap);
        if (rc != PB_OK) { cudaStreamSynchronize(s); return rc; }
        lap(0);
        CK(cudaMemcpyAsync(e->h_sc, e->scalars.p, sizeof(Scalars), cudaMemcpyDeviceToHost, s));
        int* h_meta = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(e->h_sc) + SC_BYTES);
        *h_meta = 0;
        if (e->meta_used) CK(cudaMemcpyAsync(h_meta, e->meta_flag.p, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        lap(1);
        if (*h_meta) {
            e->dirty = true;
            return fail(PB_ERR_INVALID, (*h_meta & 1) ? "compact metadata (pb_batch.meta_codes) describes more CIGAR ops or bases than n_cigar / n_seq / n_meta_cigar"
                                                      : "compact metadata (pb_batch.meta_codes) refers to an escape that meta_esc does not hold");
        }
        if (e->h_sc->error == 2 && attempt == 0) {       // the I/D bound did not hold (I/D-only CIGARs): once more, full capacity
            if ((rc = clean_sparse_planes(e)) != PB_OK) return rc;
            continue;
        }
        if (e->h_sc->error) return pass_error((uint32_t)e->h_sc->error);
        break;
    }
    if (e->phase_timing) CK(cudaEventRecord(e->ph[2], s));
    const uint32_t ng = e->h_sc->n_groups;
    // winning strings: sized exactly, then gathered
    std::vector<Group> groups(ng);
    if (ng) {
        CK(cudaMemcpyAsync(groups.data(), R.groups, sizeof(Group) * ng, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        size_t need = 0;
        for (auto& g : groups) need += (size_t)g.win_len;
        CK(e->str_pool.ensure(need + 16, false, s));
        R.str_pool = e->str_pool.as<uint8_t>(); R.str_cap = e->str_pool.cap;
        k_indel_strings<<<(ng + 127) / 128, 128, 0, s>>>(R, e->d_batches.as<DevBatch>(), ng); e->launches++;
        k_groups_clear<<<(ng + 127) / 128, 128, 0, s>>>(R, ng); e->launches++;
        CK(cudaMemcpyAsync(groups.data(), R.groups, sizeof(Group) * ng, cudaMemcpyDeviceToHost, s));
    }
    lap(2);
    const size_t S = (size_t)R.size;
#define D2H(dst, src, bytes) if (res->dst) CK(cudaMemcpyAsync(res->dst, src, (bytes), cudaMemcpyDeviceToHost, s))
    D2H(base_count4, R.o_cnt, S * 16); D2H(qual_sum4, R.o_qs, S * 32);
    D2H(mq_sum, R.o_mq, S * 4); D2H(q_sum, R.o_q, S * 4); D2H(phys_cov, R.o_pc, S * 4); D2H(insert_size, R.o_is, S * 4);
    D2H(bad_pair, R.o_bp, S * 4); D2H(deletions, R.o_del, S * 4); D2H(del_qual, R.o_delq, S * 4);
    D2H(insertions, R.o_ins, S * 4); D2H(ins_qual, R.o_insq, S * 4); D2H(clips, R.o_clips, S * 4);
    D2H(coverage_arr, R.o_cov, S * 4); D2H(frag_coverage, R.o_frag, S * 4);
    D2H(weighted_qual, R.o_wq, S); D2H(weighted_mq, R.o_wmq, S); D2H(flags, R.o_flags, S); D2H(call, R.o_call, S * 8);
#undef D2H
    res->n_calls = calls_cap ? (int64_t)e->h_sc->n_calls : 0;
    if (calls_cap && res->n_calls)
        CK(cudaMemcpyAsync(res->calls, e->call_entries.p, (size_t)std::min(res->n_calls, calls_cap) * sizeof(pb_call_entry), cudaMemcpyDeviceToHost, s));
    if (insert_sizes_out)
        for (size_t i = 0; i < e->batches.size(); i++)
            if (insert_sizes_out[i] && e->batches[i].d.n_reads)
                CK(cudaMemcpyAsync(insert_sizes_out[i], e->batches[i].d.insert_out, (size_t)e->batches[i].d.n_reads * 4, cudaMemcpyDeviceToHost, s));
    std::vector<uint8_t> pool;
    if (e->phase_timing) CK(cudaEventRecord(e->ph[3], s));
    lap(3);
    CK(cudaStreamSynchronize(s));
    lap(4);
    if (e->phase_timing) {
        for (int i = 0; i < 3; i++) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, e->ph[i], e->ph[i + 1])); e->ph_ms[i] += ms; }
        e->ph_regions++;
    }
    CK(cudaMemcpyAsync(e->h_sc, e->scalars.p, sizeof(Scalars), cudaMemcpyDeviceToHost, s));
    const size_t nbat = e->batches.size();
    unsigned long long* h_bc = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(e->h_sc) + SC_BC_OFF);
    if (nbat) CK(cudaMemcpyAsync(h_bc, R.batch_bc, nbat * BC_SPREAD * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (e->h_sc->error) return fail(PB_ERR_CUDA, "internal error flag set by a kernel");
    const size_t nbytes = (size_t)e->h_sc->str_bytes;
    if (nbytes) { pool.resize(nbytes); CK(cudaMemcpy(pool.data(), R.str_pool, nbytes, cudaMemcpyDeviceToHost)); }
    e->dirty = false;
    lap(5);
    // scalars
    const Scalars& sc = *e->h_sc;
    res->size = R.size; res->base_count = (int64_t)sc.base_count; res->coverage = sc.coverage;
    res->aligned_bases = (int64_t)sc.aligned_bases; res->read_count = sc.read_count; res->min_depth = sc.min_depth;
    res->unknown_ops = sc.unknown_ops; res->dropped_oob = sc.dropped_oob;
    // per-BAM deltas around each batch, as BamFile.process takes them (BamFile.scala:120-122,142-146)
    res->n_batches = (int64_t)nbat;
    {
        int64_t cum = 0;
        for (size_t b = 0; b < nbat && (int64_t)b < res->batch_cap; b++) {
            int64_t bc = 0;
            for (int j = 0; j < BC_SPREAD; j++) bc += (int64_t)h_bc[b * BC_SPREAD + j];
            if (res->batch_read_count) res->batch_read_count[b] = (int32_t)e->batches[b].d.n_reads;     // readCount += 1 per addRead (:218)
            if (res->batch_base_count) res->batch_base_count[b] = bc;
            if (res->batch_coverage) res->batch_coverage[b] = roundDivL(cum + bc, R.size) - roundDivL(cum, R.size);   // PileUpRegion.scala:36
            cum += bc;
        }
    }
    // indel evidence, ordered by (locus, kind)
    std::sort(groups.begin(), groups.end(), [](const Group& a, const Group& b) { return a.loc < b.loc || (a.loc == b.loc && a.kind < b.kind); });
    int64_t ni = 0, nbo = 0;
    for (auto& g : groups) {
        if (res->indels && ni < res->indels_cap) {
            pb_indel& o = res->indels[ni];
            o.locus_index = g.loc; o.kind = g.kind; o.list_len = g.list_len; o.win_count = g.win_count;
            o.win_len = g.win_len; o.win_has_n = g.win_has_n; o.str_off = nbo;
            if (res->indel_bytes && nbo + g.win_len <= res->indel_bytes_cap && g.win_len)
                memcpy(res->indel_bytes + nbo, pool.data() + g.str_off, (size_t)g.win_len);
        }
        ni++; nbo += g.win_len;
    }
    res->n_indels = ni; res->n_indel_bytes = nbo;
    lap(6);
    if (free_batches(e) != PB_OK) return PB_ERR_CUDA;
    lap(7); if (e->hf_calls++ >= 8) e->hf_n++;
    e->in_region = false;
    return PB_OK;
}

// ---------------------------------------------------------------------------------------------
// host packer: what a re-plumbed BamFile.process (BamFile.scala:126-139) calls per record
// ---------------------------------------------------------------------------------------------
struct pb_packer {
    std::vector<int32_t> pos, tlen, read_len;
    std::vector<uint8_t> mapq, flags, quals, bases2, exc_base, exc_qual, qual_codes;
    std::vector<uint32_t> cigar_off{0}, cigar, seq_off, exc_idx;
    // alphabet of the stored quality bytes: at most 16 distinct values -> the 4-bit transport is possible
    uint8_t code_of[256]; uint8_t lut[16]; int n_codes = 0; bool q4_ok = true;
    pb_packer() { memset(code_of, 0xFF, sizeof(code_of)); memset(lut, 0, sizeof(lut)); note(0); }
    void note(uint8_t v) {
        if (code_of[v] != 0xFF || !q4_ok) return;
        if (n_codes == 16) { q4_ok = false; return; }
        code_of[v] = (uint8_t)n_codes; lut[n_codes++] = v;
    }
};

extern "C" int pb_packer_create(pb_packer** out) { if (!out) return fail(PB_ERR_INVALID, "null"); *out = new pb_packer(); return PB_OK; }
extern "C" int pb_packer_destroy(pb_packer* p) { delete p; return PB_OK; }
extern "C" int pb_packer_reset(pb_packer* p) {
    if (!p) return fail(PB_ERR_INVALID, "null");
    p->pos.clear(); p->tlen.clear(); p->read_len.clear(); p->mapq.clear(); p->flags.clear(); p->quals.clear();
    p->bases2.clear(); p->exc_base.clear(); p->exc_qual.clear(); p->cigar_off.assign(1, 0); p->cigar.clear();
    p->seq_off.clear(); p->exc_idx.clear(); p->qual_codes.clear();
    memset(p->code_of, 0xFF, sizeof(p->code_of)); memset(p->lut, 0, sizeof(p->lut)); p->n_codes = 0; p->q4_ok = true; p->note(0);
    return PB_OK;
}

// One record into the packer.  BASES gives, for base j, its 2-bit code (or 0x80 for anything that is not exactly A/C/G/T)
// and, for the exception table, its ASCII letter.
namespace {
struct AsciiBases {
    const uint8_t* seq;
    static const uint8_t* lut() { static const struct L { uint8_t v[256]; L() { memset(v, 0x80, 256); v['A'] = 0; v['C'] = 1; v['G'] = 2; v['T'] = 3; } } T; return T.v; }
    uint8_t code(int32_t j) const { return lut()[seq[j]]; }
    uint8_t ascii(int32_t j) const { return seq[j]; }
};
struct BamBases {                       // BAM's own encoding: two bases per byte, high nibble first, =ACMGRSVTWYHKDBN
    const uint8_t* seq4;
    uint8_t nib(int32_t j) const { return (uint8_t)((seq4[j >> 1] >> ((~j & 1) << 2)) & 15); }
    uint8_t code(int32_t j) const { static const uint8_t C4[16] = {0x80, 0, 1, 0x80, 2, 0x80, 0x80, 0x80, 3, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80}; return C4[nib(j)]; }
    uint8_t ascii(int32_t j) const { return (uint8_t)"=ACMGRSVTWYHKDBN"[nib(j)]; }
};
template <class BASES>
int packer_add(pb_packer* p, int32_t pos, int32_t tlen, int32_t mapq, uint32_t flags, const uint32_t* cigar, int32_t n_cigar,
               const BASES& B, const uint8_t* qual, int32_t read_len) {
    if (!p || read_len < 0 || n_cigar < 0) return fail(PB_ERR_INVALID, "bad argument");
    const size_t off = p->quals.size();
    const size_t padded = ((size_t)read_len + 3) & ~(size_t)3;
    if (off + padded >= (1ull << 32)) return fail(PB_ERR_INVALID, "batch full (n_seq must stay < 2^32): start a new batch");
    if (!p->pos.empty() && pos < p->pos.back()) return fail(PB_ERR_UNSORTED, "records must be added in coordinate order");
    const bool hasq = qual != nullptr && read_len > 0 && qual[0] != 0xFF;   // htsjdk: 0xFF.. == no qualities
    p->pos.push_back(pos); p->tlen.push_back(tlen); p->read_len.push_back(read_len);
    p->mapq.push_back((uint8_t)mapq);
    p->flags.push_back((uint8_t)((flags & ~(uint32_t)PB_F_HAS_QUALS) | (hasq ? PB_F_HAS_QUALS : 0)));
    p->cigar.insert(p->cigar.end(), cigar, cigar + n_cigar);
    p->cigar_off.push_back((uint32_t)p->cigar.size());
    p->seq_off.push_back((uint32_t)off);
    p->quals.resize(off + padded, 0);
    p->bases2.resize((off + padded) / 4, 0);
    // four bases per step: one bases2 byte, four quality bytes; anything irregular (a letter that is not exactly A/C/G/T, a
    // quality byte >= 128) takes the base-by-base path for that group of four
    uint8_t* const qo = p->quals.data() + off;
    uint8_t* const bo = p->bases2.data() + off / 4;               // `off` is a multiple of 4
    auto slow = [&](int32_t j) {
        const uint8_t code = B.code(j), q = hasq ? qual[j] : 0;
        if ((code | q) & 0x80) {
            qo[j] = 0x80;
            p->exc_idx.push_back((uint32_t)(off + (size_t)j)); p->exc_base.push_back(B.ascii(j)); p->exc_qual.push_back(q);
        } else qo[j] = q;
        if (!(code & 0x80)) bo[j >> 2] |= (uint8_t)(code << (2 * (j & 3)));      // (a plain letter keeps its code even when its quality is an exception)
        if (p->code_of[qo[j]] == 0xFF) p->note(qo[j]);
    };
    int32_t j = 0;
    for (; j + 4 <= read_len; j += 4) {
        const uint8_t c0 = B.code(j), c1 = B.code(j + 1), c2 = B.code(j + 2), c3 = B.code(j + 3);
        uint8_t q0 = 0, q1 = 0, q2 = 0, q3 = 0;
        if (hasq) { q0 = qual[j]; q1 = qual[j + 1]; q2 = qual[j + 2]; q3 = qual[j + 3]; }
        if ((c0 | c1 | c2 | c3 | q0 | q1 | q2 | q3) & 0x80) { slow(j); slow(j + 1); slow(j + 2); slow(j + 3); continue; }
        bo[j >> 2] = (uint8_t)(c0 | (c1 << 2) | (c2 << 4) | (c3 << 6));
        qo[j] = q0; qo[j + 1] = q1; qo[j + 2] = q2; qo[j + 3] = q3;
        if ((p->code_of[q0] | p->code_of[q1] | p->code_of[q2] | p->code_of[q3]) == 0xFF) { p->note(q0); p->note(q1); p->note(q2); p->note(q3); }
    }
    for (; j < read_len; j++) slow(j);
    return PB_OK;
}
}  // namespace

extern "C" int pb_packer_add(pb_packer* p, int32_t pos, int32_t tlen, int32_t mapq, uint32_t flags,
                             const uint32_t* cigar, int32_t n_cigar, const uint8_t* seq, const uint8_t* qual, int32_t read_len) {
    if (read_len > 0 && !seq) return fail(PB_ERR_INVALID, "bad argument");
    return packer_add(p, pos, tlen, mapq, flags, cigar, n_cigar, AsciiBases{seq}, qual, read_len);
}

extern "C" int pb_packer_add_bam(pb_packer* p, int32_t pos, int32_t tlen, int32_t mapq, uint32_t flags,
                                 const uint32_t* cigar, int32_t n_cigar, const uint8_t* seq4, const uint8_t* qual, int32_t read_len) {
    if (read_len > 0 && !seq4) return fail(PB_ERR_INVALID, "bad argument");
    return packer_add(p, pos, tlen, mapq, flags, cigar, n_cigar, BamBases{seq4}, qual, read_len);
}

extern "C" int pb_packer_add_many(pb_packer* p, int64_t n, const int32_t* pos, const int32_t* tlen, const uint8_t* mapq,
                                  const uint8_t* flags, const int32_t* read_len, const uint32_t* cigar_off,
                                  const uint32_t* cigar, const uint64_t* ascii_off, const uint8_t* seq, const uint8_t* qual) {
    for (int64_t r = 0; r < n; r++) {
        const bool hasq = (flags[r] & PB_F_HAS_QUALS) && qual;
        int rc = pb_packer_add(p, pos[r], tlen[r], mapq[r], flags[r], cigar + cigar_off[r],
                               (int32_t)(cigar_off[r + 1] - cigar_off[r]), seq + ascii_off[r],
                               hasq ? qual + ascii_off[r] : nullptr, read_len[r]);
        if (rc != PB_OK) return rc;
    }
    return PB_OK;
}

extern "C" int pb_packer_view(pb_packer* p, pb_batch* b) {
    if (!p || !b) return fail(PB_ERR_INVALID, "null");
    memset(b, 0, sizeof(*b));
    b->n_reads = (int64_t)p->pos.size(); b->n_cigar = (int64_t)p->cigar.size();
    b->n_seq = (int64_t)p->quals.size(); b->n_exc = (int64_t)p->exc_idx.size();
    b->pos = p->pos.data(); b->tlen = p->tlen.data(); b->read_len = p->read_len.data();
    b->mapq = p->mapq.data(); b->flags = p->flags.data(); b->cigar_off = p->cigar_off.data();
    b->cigar = p->cigar.data(); b->seq_off = p->seq_off.data(); b->quals = p->quals.data();
    b->bases2 = p->bases2.data(); b->exc_idx = p->exc_idx.data(); b->exc_base = p->exc_base.data();
    b->exc_qual = p->exc_qual.data(); b->mem = PB_MEM_HOST;
    if (p->q4_ok) {                                     // binned qualities: offer the packed transport as well
        const size_t ns = p->quals.size();
        const int bits = p->n_codes <= 8 ? 3 : 4;
        p->qual_codes.assign((ns * bits + 7) / 8 + 32, 0);
        const uint8_t* q = p->quals.data();
        uint8_t* o = p->qual_codes.data();
        size_t i = 0;
        if (bits == 3) {
            for (; i + 8 <= ns; i += 8) {                // eight codes -> three bytes
                uint32_t v = 0;
                for (int t = 0; t < 8; t++) v |= (uint32_t)p->code_of[q[i + t]] << (3 * t);
                uint8_t* d = o + 3 * (i >> 3);
                d[0] = (uint8_t)v; d[1] = (uint8_t)(v >> 8); d[2] = (uint8_t)(v >> 16);
            }
        } else {
            for (; i + 2 <= ns; i += 2) o[i >> 1] = (uint8_t)(p->code_of[q[i]] | (p->code_of[q[i + 1]] << 4));
        }
        for (; i < ns; i++) {
            const uint32_t c = p->code_of[p->quals[i]];
            const size_t bit = i * bits;
            o[bit >> 3] |= (uint8_t)(c << (bit & 7));
            if ((bit & 7) + bits > 8) o[(bit >> 3) + 1] |= (uint8_t)(c >> (8 - (bit & 7)));
        }
        b->qual_codes = p->qual_codes.data(); b->qual_code_bits = bits; memcpy(b->qual_lut, p->lut, 16);
    }
    return PB_OK;
}
