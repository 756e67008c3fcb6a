// Scratch experiment 2 (not product code): 1-read + 1-write copy with the access pattern of the fused
// encode+loss kernel -- every warp streams its own 12 KB map in chunks (ring of TMA slots) -- against
// patterns where the warps of a CTA share one contiguous window. Optional artificial per-chunk delay.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scratch/stream_bench2 scratch/stream_bench2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// PATTERN 0: warp-per-map; CTA owns a contiguous range of maps, warp w takes maps w, w+nwarps, ... of it
// PATTERN 1: chunks of the CTA's contiguous range dealt round-robin to warps (window = nwarps*ring chunks)
// OUT 0: st.global per lane; OUT 1: in-place smem then bulk store (ring slot reused after wait_group.read)
template <int PATTERN, int OUT>
__global__ void __launch_bounds__(1024, 1) copy_kernel(const float* __restrict__ a, float* __restrict__ c, long long nmaps, int map_quads,
                                                       int chunk_quads, int ring, int nwarps, int delay) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunk_bytes = chunk_quads * 16u;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + warp * ring;
    unsigned char* slots = smem + 2048 + (size_t)warp * ring * chunk_bytes;
    if (lane == 0) { for (int r = 0; r < ring; ++r) mbar_init(bars + r, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const int cpm = map_quads / chunk_quads;
    const long long lo = (long long)blockIdx.x * nmaps / gridDim.x, hi = (long long)(blockIdx.x + 1) * nmaps / gridDim.x;
    // chunk sequence of this warp, as global chunk ids
    long long nmine;
    if (PATTERN == 0) { const long long mymaps = (hi - lo - warp + nwarps - 1) / nwarps; nmine = (hi - lo > warp) ? mymaps * cpm : 0; }
    else { const long long tot = (hi - lo) * cpm; nmine = (tot > warp) ? (tot - warp + nwarps - 1) / nwarps : 0; }
    auto chunk_id = [&](long long k) -> long long {
        if (PATTERN == 0) return (lo + warp + (k / cpm) * nwarps) * cpm + (k % cpm);
        return lo * cpm + warp + k * nwarps;
    };
    long long pk = 0; int ps = 0;
    auto issue = [&]() {
        if (pk >= nmine) return;
        mbar_expect_tx(bars + ps, chunk_bytes);
        bulk_g2s(slots + (size_t)ps * chunk_bytes, a + chunk_id(pk) * chunk_quads * 4, chunk_bytes, bars + ps);
        ++pk; if (++ps == ring) ps = 0;
    };
    if (lane == 0) for (int r = 0; r < ring; ++r) issue();
    int cs = 0; uint32_t parity = 0;
    for (long long k = 0; k < nmine; ++k) {
        mbar_wait(bars + cs, parity);
        float4* x = reinterpret_cast<float4*>(slots + (size_t)cs * chunk_bytes);
        float4* out = reinterpret_cast<float4*>(c + chunk_id(k) * chunk_quads * 4);
        for (int q = lane; q < chunk_quads; q += 32) {
            float4 v = x[q];
            for (int d = 0; d < delay; ++d) { v.x = fmaf(v.x, 1.0000001f, 1e-9f); v.y = fmaf(v.y, 1.0000001f, 1e-9f); v.z = fmaf(v.z, 1.0000001f, 1e-9f); v.w = fmaf(v.w, 1.0000001f, 1e-9f); }
            if (OUT == 0) out[q] = v; else x[q] = v;
        }
        if (OUT == 1) fence_async();
        __syncwarp();
        if (lane == 0) {
            if (OUT == 1) { bulk_s2g(out, x, chunk_bytes); bulk_commit(); bulk_wait_read0(); }
            else fence_async();
            issue();
        }
        if (++cs == ring) { cs = 0; parity ^= 1u; }
    }
}

template <typename F> float time_it(F fn, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    fn(); fn(); CK(cudaDeviceSynchronize());
    std::vector<float> t;
    for (int r = 0; r < reps; ++r) { CK(cudaEventRecord(e0)); fn(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); t.push_back(ms); }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

int main(int argc, char** argv) {
    const int map_quads = 768;                                   // 64x48 floats
    const long long nmaps = 17408LL * 8;                         // 8 x (1024 persons x 17 joints): 1.7 GB per array
    const long long n = nmaps * map_quads * 4;
    float *a, *c;
    CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&c, n * 4));
    CK(cudaMemset(a, 1, n * 4)); CK(cudaMemset(c, 0, n * 4));
    printf("cudaMemcpy d2d: %.1f GB/s\n", 2 * n * 4.0 / time_it([&] { CK(cudaMemcpyAsync(c, a, n * 4, cudaMemcpyDeviceToDevice)); }) / 1e6);
    for (int delay : {0, 8, 24})
        for (int pattern = 0; pattern <= 1; ++pattern)
            for (int out = 0; out <= 1; ++out)
                for (int cq : {96, 192, 384, 768})
                    for (int ring : {1, 2, 3, 4})
                        for (int nwarps : {4, 8, 12, 16, 24, 32}) {
                            if (delay && (nwarps < 8 || cq > 384 || ring > 3)) continue;
                            const size_t smem = 2048 + (size_t)nwarps * ring * cq * 16;
                            if (smem > 227 * 1024 || nwarps * ring > 256) continue;
                            float ms;
#define RUN(P, O) { CK(cudaFuncSetAttribute(copy_kernel<P, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                    ms = time_it([&] { copy_kernel<P, O><<<148, nwarps * 32, smem>>>(a, c, nmaps, map_quads, cq, ring, nwarps, delay); }, 3); }
                            if (pattern == 0 && out == 0) RUN(0, 0) else if (pattern == 0) RUN(0, 1) else if (out == 0) RUN(1, 0) else RUN(1, 1)
                            CK(cudaGetLastError());
                            printf("delay=%-2d pattern=%d out=%d chunk=%5dB ring=%d warps=%-2d inflight=%3dKB  %8.3f ms %7.1f GB/s\n", delay, pattern, out, cq * 16, ring, nwarps,
                                   (int)(nwarps * ring * cq * 16 / 1024), ms, 2 * n * 4.0 / ms / 1e6);
                            fflush(stdout);
                        }
    return 0;
}
