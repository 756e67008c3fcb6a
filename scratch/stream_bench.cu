// Scratch experiment (not product code): which access style gets the most out of HBM for a
// "2 reads + 1 write" streaming kernel (c = a - b) and for a pure copy on this B200?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scratch/stream_bench scratch/stream_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <string.h>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float4 ldg_nc(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_cs(float4* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- style 1: register streaming, U independent quads per thread per tile --------------------------
template <int U, int NREAD, bool CS>
__global__ void __launch_bounds__(256) reg_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ c, long long nq) {
    const long long tile = 256LL * U;
    for (long long base = (long long)blockIdx.x * tile; base < nq; base += (long long)gridDim.x * tile) {
        float4 x[U], y[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long q = base + u * 256 + threadIdx.x;
            if (q < nq) { x[u] = ldg_nc(a + q); if (NREAD == 2) y[u] = ldg_nc(b + q); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long q = base + u * 256 + threadIdx.x;
            if (q < nq) {
                float4 v = x[u];
                if (NREAD == 2) { v.x -= y[u].x; v.y -= y[u].y; v.z -= y[u].z; v.w -= y[u].w; }
                if (CS) stg_cs(c + q, v); else c[q] = v;
            }
        }
    }
}

// ---- style 2: per-warp TMA ring in, register compute, direct st.global out -----------------------------
// ---- style 3: per-warp TMA ring in, compute into smem, TMA bulk store out ------------------------------
template <int NREAD, int OUTMODE>
__global__ void __launch_bounds__(1024, 1) tma_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ c,
                                                      long long nchunks, int chunk_quads, int ring, int nwarps) {
    constexpr bool BULK_OUT = (OUTMODE == 1);
    extern __shared__ __align__(128) unsigned char smem[];
    float sink = 0.f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunk_bytes = chunk_quads * 16u;
    const uint32_t slot_bytes = chunk_bytes * (NREAD + (BULK_OUT ? 1 : 0));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + warp * ring;   // <= 256 barriers
    unsigned char* slots = smem + 2048 + (size_t)warp * ring * slot_bytes;
    if (lane == 0) { for (int r = 0; r < ring; ++r) mbar_init(bars + r, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const long long gw = (long long)blockIdx.x * nwarps + warp, total = (long long)gridDim.x * nwarps;
    long long pi = gw; int ps = 0;
    auto issue = [&]() {
        if (pi >= nchunks) return;
        unsigned char* dst = slots + (size_t)ps * slot_bytes;
        mbar_expect_tx(bars + ps, chunk_bytes * NREAD);
        bulk_g2s(dst, a + pi * chunk_quads * 4, chunk_bytes, bars + ps);
        if (NREAD == 2) bulk_g2s(dst + chunk_bytes, b + pi * chunk_quads * 4, chunk_bytes, bars + ps);
        pi += total; if (++ps == ring) ps = 0;
    };
    if (lane == 0) for (int r = 0; r < ring; ++r) issue();
    int cs = 0; uint32_t parity = 0;
    for (long long i = gw; i < nchunks; i += total) {
        mbar_wait(bars + cs, parity);
        unsigned char* slot = slots + (size_t)cs * slot_bytes;
        const float4* xa = reinterpret_cast<const float4*>(slot);
        const float4* xb = reinterpret_cast<const float4*>(slot + chunk_bytes);
        float4* xo = reinterpret_cast<float4*>(slot + chunk_bytes * NREAD);
        float4* out = reinterpret_cast<float4*>(c + i * chunk_quads * 4);
        if (BULK_OUT) {      // the store issued from this slot `ring` chunks ago must have read its smem
            if (lane == 0) { switch (ring) { case 2: bulk_wait_read<1>(); break; case 3: bulk_wait_read<2>(); break; case 4: bulk_wait_read<3>(); break;
                                           case 6: bulk_wait_read<5>(); break; case 8: bulk_wait_read<7>(); break; default: bulk_wait_read<0>(); } }
            __syncwarp();
        }
#pragma unroll 2
        for (int q = lane; q < chunk_quads; q += 32) {
            float4 v = xa[q];
            if (NREAD == 2) { const float4 w = xb[q]; v.x -= w.x; v.y -= w.y; v.z -= w.z; v.w -= w.w; }
            if (OUTMODE == 2) sink += v.x + v.y + v.z + v.w; else if (BULK_OUT) xo[q] = v; else out[q] = v;
        }
        if (BULK_OUT) fence_async();
        __syncwarp();
        if (lane == 0) {
            if (BULK_OUT) { bulk_s2g(out, xo, chunk_bytes); bulk_commit(); }
            else fence_async();
            issue();
        }
        if (++cs == ring) { cs = 0; parity ^= 1u; }
    }
    if (BULK_OUT && lane == 0) bulk_wait_read<0>();
    if (OUTMODE == 2 && sink == 123.456f) c[0] = sink;
}

template <typename F> float time_it(F fn, int reps = 7) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    fn(); fn(); CK(cudaDeviceSynchronize());
    std::vector<float> t;
    for (int r = 0; r < reps; ++r) { CK(cudaEventRecord(e0)); fn(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); t.push_back(ms); }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

int main(int argc, char** argv) {
    const long long n = (argc > 1 ? atoll(argv[1]) : 512LL) << 20;   // floats per array (default 2 GiB each)
    float *a, *b, *c;
    CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&c, n * 4));
    CK(cudaMemset(a, 1, n * 4)); CK(cudaMemset(b, 2, n * 4)); CK(cudaMemset(c, 0, n * 4));
    const long long nq = n / 4;
    auto report = [&](const char* name, float ms, int streams) { printf("%-66s %8.3f ms  %7.1f GB/s\n", name, ms, streams * n * 4.0 / ms / 1e6); fflush(stdout); };
    report("cudaMemcpyAsync d2d (1r+1w)", time_it([&] { CK(cudaMemcpyAsync(c, a, n * 4, cudaMemcpyDeviceToDevice)); }), 2);
#define REG(U, NR, CS, G) report("reg U=" #U " nread=" #NR " cs=" #CS " grid=148*" #G, time_it([&] { reg_kernel<U, NR, CS><<<148 * G, 256>>>((const float4*)a, (const float4*)b, (float4*)c, nq); }), NR + 1)
    REG(1, 1, false, 8); REG(4, 1, false, 8); REG(4, 1, true, 8); REG(8, 1, false, 4);
    REG(1, 2, false, 8); REG(2, 2, false, 8); REG(4, 2, false, 8); REG(4, 2, true, 8); REG(4, 2, false, 4); REG(8, 2, false, 4); REG(3, 2, false, 8); REG(4, 2, false, 16);
    struct Row { char name[128]; float ms; int streams; };
    std::vector<Row> rows;
    for (int nread = 1; nread <= 2; ++nread)
        for (int bulk = 0; bulk <= 2; ++bulk)
            for (int cq : {96, 192, 384, 768, 1536})
                for (int ring : {2, 3, 4, 6, 8})
                    for (int nwarps : {1, 2, 3, 4, 6, 8, 12, 16, 24, 32}) {
                        if (bulk == 2 && nread == 2) continue;
                        const size_t slot = (size_t)cq * 16 * (nread + (bulk == 1));
                        const size_t smem = 2048 + (size_t)nwarps * ring * slot;
                        if (smem > 227 * 1024 || nwarps * ring > 256) continue;
                        if ((size_t)nwarps * ring * cq * 16 * nread < 12 * 1024) continue;   // too little in flight to matter
                        const long long nchunks = nq / cq;
                        Row r;
                        snprintf(r.name, sizeof r.name, "tma nread=%d bulk_out=%d chunk=%dB ring=%d warps=%d inflight=%dKB", nread, bulk, cq * 16, ring, nwarps, (int)(nwarps * ring * cq * 16 * nread / 1024));
                        float ms;
#define TMA(NR, BO) { CK(cudaFuncSetAttribute(tma_kernel<NR, BO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                      ms = time_it([&] { tma_kernel<NR, BO><<<148, nwarps * 32, smem>>>(a, b, c, nchunks, cq, ring, nwarps); }, 3); }
                        if (nread == 1 && bulk == 0) TMA(1, 0) else if (nread == 1 && bulk == 1) TMA(1, 1) else if (nread == 1) TMA(1, 2) else if (!bulk) TMA(2, 0) else TMA(2, 1)
                        CK(cudaGetLastError());
                        r.ms = ms; r.streams = nread + (bulk == 2 ? 0 : 1);
                        rows.push_back(r);
                    }
    for (int nread = 1; nread <= 2; ++nread)
        for (int bulk = 0; bulk <= 2; ++bulk) {
            std::vector<Row> sel;
            char key[64]; snprintf(key, sizeof key, "tma nread=%d bulk_out=%d", nread, bulk);
            for (auto& r : rows) if (!strncmp(r.name, key, strlen(key))) sel.push_back(r);
            std::sort(sel.begin(), sel.end(), [](const Row& x, const Row& y) { return x.ms < y.ms; });
            for (size_t i = 0; i < sel.size() && i < 10; ++i) report(sel[i].name, sel[i].ms, sel[i].streams);
        }
    CK(cudaDeviceSynchronize());
    return 0;
}
