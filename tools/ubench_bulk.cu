// Microbenchmark: latency and per-SM throughput of cp.async.bulk (UBLKCP) and per-lane cp.async
// (LDGSTS) for ~5 KB contiguous ranges, as used by the pileup kernels' staging.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const uint8_t* src, size_t span, int bytes, int depth, int iters, int mode, long long* out) {
    extern __shared__ __align__(128) uint8_t sm[];
    unsigned long long* bar = (unsigned long long*)sm;           // [8]
    uint8_t* buf = sm + 128;
    const int lane = threadIdx.x;
    if (lane == 0) { for (int i = 0; i < depth; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar[i])), "r"(mode == 0 ? 1 : 32)); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    uint64_t rng = blockIdx.x * 7919ull + 13;
    long long tsum = 0, tmax = 0; long long t0 = clock64();
    long long issue[8];
    auto issue_one = [&](int slot) {
        rng = rng * 6364136223846793005ull + 1442695040888963407ull;
        size_t off = ((rng >> 20) % (span - bytes - 64)) & ~(size_t)127;
        issue[slot] = clock64();
        if (mode == 0) {
            if (lane == 0) {
                asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(s32(&bar[slot])), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf + slot * 8192)), "l"(src + off), "r"(bytes), "r"(s32(&bar[slot])) : "memory");
            }
        } else {
            for (int b = lane * 16; b < bytes; b += 512)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(buf + slot * 8192 + b)), "l"(src + off + b) : "memory");
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s32(&bar[slot])) : "memory");
        }
    };
    for (int i = 0; i < depth; i++) issue_one(i);
    for (int it = 0; it < iters; it++) {
        int slot = it % depth; uint32_t par = (it / depth) & 1, ok = 0;
        while (!ok) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(s32(&bar[slot])), "r"(par) : "memory");
        long long t = clock64() - issue[slot]; tsum += t; if (t > tmax) tmax = t;
        __syncwarp();
        if (it + depth < iters) issue_one(slot);
    }
    long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x * 3] = tsum / iters; out[blockIdx.x * 3 + 1] = tmax; out[blockIdx.x * 3 + 2] = t1 - t0; }
}
int main() {
    size_t span = 4ull << 30; uint8_t* src; cudaMalloc(&src, span); cudaMemset(src, 1, span);
    long long* out; cudaMallocManaged(&out, 148 * 8 * 3 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + 8 * 8192);
    const int iters = 200;
    for (int mode = 0; mode < 2; mode++)
        for (int bytes : {1024, 5120})
            for (int depth : {1, 4, 8})
                for (int cps : {1, 4}) {
                    int grid = 148 * cps;
                    size_t smem = 128 + depth * 8192; if (cps == 4 && smem > 50000) continue;
                    k<<<grid, 32, smem>>>(src, span, bytes, depth, iters, mode, out);
                    cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
                    double lat = 0, mx = 0, tot = 0; for (int b = 0; b < grid; b++) { lat += out[b * 3]; mx = mx > out[b * 3 + 1] ? mx : out[b * 3 + 1]; tot += out[b * 3 + 2]; }
                    lat /= grid; tot /= grid;
                    printf("%s bytes=%5d depth=%d ctas/sm=%d: avg latency %7.0f cyc, max %7.0f, per-SM %6.2f B/cyc, chip ~%5.2f TB/s @1.9GHz\n", mode ? "LDGSTS" : "UBLKCP", bytes, depth, cps, lat, mx,
                           (double)bytes * iters * cps / tot, (double)bytes * iters * cps / tot * 148 * 1.9e9 / 1e12);
                }
    return 0;
}
