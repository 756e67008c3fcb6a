// SURVEY section 8f ranks 1+2: the training-step half of the path as ONE pass.
//
//   targets, mask = get_heat_map(joints)                      commons/transforms.py:167-191
//   loss = 0.5 * MSELoss(pred * mask, target * mask); backward   processors/dp_pose_hrnet_solver.py:106-107
//   acc  = HeatMapAcc()(pred * mask, target * mask)              metrics/pose_metrics.py:212-245, solver :123-124
//
// The targets are never materialised: each warp owns one (person, joint) map, evaluates the
// separable float64 Gaussian factors exactly like the stand-alone encoder, and while the predicted
// map streams by (one 16-byte load per lane per step) forms target, masked difference, squared
// error, gradient and -- for HeatMapAcc -- the running argmax of both masked maps. HBM traffic per
// person drops from 209 168 (encode) + 626 756 (loss) + 2 x 208 896 (two argmax passes) bytes to
// read pred + write grad = 417 860 bytes. Per-element arithmetic is the same as in the separate
// kernels, so grad/targets/weights are bit-identical to sp_encode_f32 + sp_mse_fwd_bwd_f32 and the
// loss differs only by the float64 summation order.
#include "sp_common.cuh"
#include "sp_gauss.cuh"
#include "sp_reduce.cuh"
#include "sp_lowp.cuh"
#include "sp_train_dev.cuh"
#include <math_constants.h>
#include <stdlib.h>

#ifdef SP_TRAIN_TRACE
// scratch instrumentation (never compiled into the product library): per-warp timestamps of the
// period-tiled kernel, 8 x int64 per warp: [globaltimer at entry, after griddepcontrol.wait, first
// chunk landed, last map done, after the loss reduction, maps processed, clock64 span of the map loop, 0]
__device__ long long* g_trace_ptr = nullptr;
extern "C" int sp_debug_set_trace(void* p) {
    return (int)cudaMemcpyToSymbol(g_trace_ptr, &p, sizeof(p));
}
__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

namespace {

using namespace sp_gauss;
using namespace sp_reduce;
using namespace sp_trn;

// Variant A: predicted maps read straight from global memory (any map size).
// PT = storage type of pred and grad (float32, or float16 / bfloat16 under torch.cuda.amp: the arithmetic stays
// float32 as autocast runs it, the gradient is cast to PT once, after the upstream scale has been applied).
template <bool WRITE_GRAD, bool WRITE_TARGETS, bool ACC, typename PT = float>
__global__ void __launch_bounds__(kThreads)
encode_mse_kernel(const MapIo io, float* __restrict__ loss, MseWorkspace* __restrict__ ws, double inv_count) {
    extern __shared__ __align__(16) double factors[];   // per warp: ex[Wpad] then ey[H]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpad = (io.W + 1) & ~1;
    double* ex = factors + (size_t)warp * (wpad + ((io.H + 1) & ~1));   // even stride: 16-byte aligned ex for every warp (odd H)
    double* ey = ex + wpad;
    const int hw = io.H * io.W;
    const int total_warps = gridDim.x * kWarps;
    double sum_sq = 0.0;
    sp::grid_dep_wait();
    const float half_scale = io.scale_dev ? __fmul_rn(io.half_scale, __ldg(io.scale_dev)) : io.half_scale;
    Joint3 jn = load_joint(io, blockIdx.x * kWarps + warp);
    for (int m = blockIdx.x * kWarps + warp; m < io.nmaps; m += total_warps) {
        const Joint3 jc = jn;
        jn = load_joint(io, m + total_warps);               // next map's joint: latency hidden behind this map
        const JointVerdict jv = prepare_map(io, m, jc, ex, ey, lane);
        MapState st;
        begin_map(io, jv, st, lane, ACC);
        run_quads<WRITE_GRAD, WRITE_TARGETS, ACC, false, PT>(
            io, m, reinterpret_cast<const float4*>(reinterpret_cast<const PT*>(io.pred) + (size_t)m * hw), 0, hw >> 2, jv, ex, ey,
            st, lane, half_scale);
        if (ACC) end_map_acc<PT>(io, m, jv, ex, ey, st, lane);
        sum_sq += (double)st.acc;
    }
    if (loss != nullptr) finish_loss<512>(sum_sq, ws, loss, inv_count);     // every lane carries the quads it visited
}

// Variant B (default): persistent, one CTA per SM, up to 32 warps. Every warp streams its predicted
// maps through a private ring of small shared-memory slots (1-D TMA bulk copies of `chunk_quads`
// quads each, mbarrier complete_tx). The map is consumed strictly in order, so a slot is re-armed
// with the chunk `ring` positions ahead -- possibly of the warp's NEXT map -- the moment it has been
// read: the copy engine runs a full ring ahead of the arithmetic, nothing stalls on individual
// global loads, and a 6 KB ring per warp leaves room for 32 resident warps to hide the float64
// and shared-memory latencies.
// Work distribution: CTA c owns the contiguous map range [c*nmaps/grid, (c+1)*nmaps/grid); its warps
// claim maps from a shared-memory counter (lane 0, one map ahead of the copies it is issuing) and
// pass the claimed indices to the consuming lanes through a small per-warp FIFO.
// dynamic smem: [mbarriers 1024 B][claim FIFOs + work counter 2048 B][factors per warp][ring slots per warp]
constexpr int kFifo = 16;          // entries per warp; the producer is never more than ring+2 <= 10 maps ahead
constexpr int kRingHeader = 1024 + 2048 + 64;

template <bool WRITE_GRAD, bool WRITE_TARGETS, bool ACC>
__global__ void __launch_bounds__(512, 1)
encode_mse_ring_kernel(const MapIo io, float* __restrict__ loss, MseWorkspace* __restrict__ ws, double inv_count,
                       int nwarps, int ring, int chunk_quads) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int hw = io.H * io.W;
    const int nq = hw >> 2;
    const int chunks_per_map = nq / chunk_quads;                 // host guarantees divisibility
    const uint32_t chunk_bytes = (uint32_t)chunk_quads * 16u;
    const int wpad = (io.W + 1) & ~1;
    const size_t fac_bytes = (size_t)(wpad + ((io.H + 1) & ~1)) * sizeof(double);   // multiple of 16 (odd H too)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw) + warp * ring;
    volatile int* fifo = reinterpret_cast<int*>(smem_raw + 1024) + warp * kFifo;
    int& next_map = *reinterpret_cast<int*>(smem_raw + 1024 + 2048);
    double* ex = reinterpret_cast<double*>(smem_raw + kRingHeader + warp * fac_bytes);
    double* ey = ex + wpad;
    unsigned char* slots = smem_raw + kRingHeader + (size_t)nwarps * fac_bytes + (size_t)warp * ring * chunk_bytes;
    const int range_lo = (int)((long long)blockIdx.x * io.nmaps / gridDim.x);
    const int range_hi = (int)((long long)(blockIdx.x + 1) * io.nmaps / gridDim.x);
    if (threadIdx.x == 0) next_map = range_lo;
    if (lane == 0) {
        for (int r = 0; r < ring; ++r) sp::mbar_init(bars + r, 1);
        sp::mbar_fence_init();
    }
    __syncthreads();
    sp::grid_dep_wait();            // the prologue above overlapped the previous kernel's tail

    // producer (lane 0 only): copies chunk pc of map pm into slot ps; pnext = the map claimed after pm
    int pm = -1, pnext = -1, pc = 0, ps = 0, tail = 0;
    auto claim = [&]() {
        const int m = atomicAdd(&next_map, 1);
        const int got = (m < range_hi) ? m : -1;
        fifo[tail & (kFifo - 1)] = got;
        ++tail;
        return got;
    };
    auto issue_next = [&]() {
        if (pm < 0) return;
        sp::mbar_expect_tx(bars + ps, chunk_bytes);
        sp::bulk_g2s(slots + (size_t)ps * chunk_bytes, io.pred + (size_t)pm * hw + (size_t)pc * chunk_quads * 4,
                     chunk_bytes, bars + ps);
        if (++pc == chunks_per_map) {
            pc = 0;
            pm = pnext;
            pnext = (pm >= 0) ? claim() : -1;
        }
        if (++ps == ring) ps = 0;
    };
    if (lane == 0) {
        pm = claim();
        pnext = (pm >= 0) ? claim() : -1;
        for (int r = 0; r < ring; ++r) issue_next();
    }
    __syncwarp();

    double sum_sq = 0.0;
    int cs = 0;                 // consumer slot
    uint32_t parity = 0;
    int head = 0;
    int m = fifo[0];
    Joint3 jn = load_joint(io, m >= 0 ? m : io.nmaps);
    while (m >= 0) {
        const Joint3 jc = jn;
        // the producer claimed this map's successor before it issued this map's first chunk
        const int m_next = fifo[(head + 1) & (kFifo - 1)];
        jn = load_joint(io, m_next >= 0 ? m_next : io.nmaps);           // next map's joint, a whole map ahead
        const JointVerdict jv = prepare_map(io, m, jc, ex, ey, lane);   // overlaps the copies in flight
        MapState st;
        begin_map(io, jv, st, lane, ACC);
        for (int c = 0; c < chunks_per_map; ++c) {
            sp::mbar_wait(bars + cs, parity);
            const float4* chunk = reinterpret_cast<const float4*>(slots + (size_t)cs * chunk_bytes);
            run_quads<WRITE_GRAD, WRITE_TARGETS, ACC, true>(io, m, chunk, c * chunk_quads, (c + 1) * chunk_quads,
                                                            jv, ex, ey, st, lane, io.half_scale);
            __syncwarp();
            if (lane == 0) {
                sp::fence_proxy_async_smem();
                issue_next();                                           // refills the slot just drained
            }
            if (++cs == ring) { cs = 0; parity ^= 1u; }
        }
        if (ACC) end_map_acc(io, m, jv, ex, ey, st, lane);
        sum_sq += (double)st.acc;
        __syncwarp();
        ++head;
        m = m_next;
    }
    finish_loss<512>(sum_sq, ws, loss, inv_count);
}

// Variant C (default for W = 48 and W = 72 when grad is wanted and targets are not): the ring of
// variant B with the per-quad and per-chunk bookkeeping compiled away. With QPR quads per row, the
// (row, quad-in-row) pattern a lane sees repeats every PERIOD = QPR / gcd(32, QPR) warp steps and
// advances by ROWS = 32 / gcd(32, QPR) rows, so the chunk body is fully unrolled over PPC periods
// with per-lane constants: no index arithmetic, for PERIOD <= 3 the lane's x factors live in
// registers (one shared-memory load per quad left: the row factor), ring depth and chunk size are
// template parameters, shared memory is addressed with 32-bit offsets, the draw / unit-mask /
// tracking decisions are made once per map (three specialisations of the map body), and the running
// argmax keeps only (value, quad) -- the position inside the winning quad is recovered at the end
// of the NEXT map from a 16-byte re-read that has had a whole map to complete.
// Why it matters: with 16 warps per SM each warp issues about one instruction every 8 cycles, so
// the kernel's time is (instructions per quad) x latency until the copy engine becomes the limit.
constexpr int cgcd(int a, int b) { return b == 0 ? a : cgcd(b, a % b); }

template <int QPR>
struct Tile {
    static constexpr int G = cgcd(32, QPR);
    static constexpr int PERIOD = QPR / G;
    static constexpr int ROWS = 32 / G;
    static constexpr bool EX_IN_REGS = PERIOD <= 3;
};

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ double lds64f(uint32_t addr) {
    double r;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ double2 lds128d(uint32_t addr) {
    double2 r;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(addr));
    return r;
}
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// Per-warp ring of RING slots of CHUNK_BYTES, all shared-memory addresses 32-bit.
template <int RING, int CHUNK_BYTES>
struct TileRing {
    uint32_t bar0, slot0;        // shared addresses of this warp's first barrier / slot
    uint32_t cs, parity;         // consumer slot and phase
    // producer side (lane 0): next chunk to request
    const char* src;             // global address of the next chunk
    int left;                    // chunks of the current map still to request (0: no map)
    int pnext;                   // map claimed after the current one (-1: none)
    uint32_t ps;
    int tail, nmaps;
    int inflight, depth;         // copies in flight / copies kept in flight while there are unclaimed maps
    int static_next, static_left, static_stride, dyn_base;
    volatile int* fifo;
    unsigned int* next_work;     // grid-wide counter in the caller's workspace
    const char* pred;
    size_t map_bytes;
    int chunks_per_map;

    // Maps are dealt GRID-WIDE: the first two of every warp are fixed (no atomic storm at launch), the
    // rest come from one counter in the caller's workspace, one map ahead of the copies being issued.
    // A per-CTA range would make the slowest SM the critical path, and the SMs are far from equal once
    // the memory system queues: at 96x72 half the TPCs finished equal ranges in ~58 us and the other
    // half in ~86 us (per-warp timestamps, profiles/r1f_fused_timeline.md).
    __device__ __forceinline__ int claim() {
        int m;
        if (static_left > 0) {
            m = static_next;
            static_next += static_stride;
            --static_left;
        } else {
            m = dyn_base + (int)atomicAdd(next_work, 1u);
        }
        const int got = (m >= 0 && m < nmaps) ? m : -1;
        fifo[tail & (kFifo - 1)] = got;
        ++tail;
        return got;
    }
    __device__ __forceinline__ void start(int m) {           // lane 0
        left = 0;
        pnext = -1;
        if (m >= 0) {
            src = pred + (size_t)m * map_bytes;
            left = chunks_per_map;
            pnext = claim();
        }
    }
    __device__ __forceinline__ void issue_next() {           // lane 0
        if (left == 0) return;
        const uint32_t bar = bar0 + ps * 8u, dst = slot0 + ps * (uint32_t)CHUNK_BYTES;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "n"(CHUNK_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "n"(CHUNK_BYTES), "r"(bar) : "memory");
        src += CHUNK_BYTES;
        ++inflight;
        if (--left == 0 && pnext >= 0) {
            src = pred + (size_t)pnext * map_bytes;
            left = chunks_per_map;
            pnext = claim();
        }
        if (RING > 1) ps = (ps + 1 == RING) ? 0u : ps + 1;
    }
    // Keep `depth` copies in flight while the CTA has maps left to hand out; once this warp is on its
    // last map (nothing claimed behind it) use every slot: the SM's other warps are running dry, a
    // lone warp at depth 1 pulls ~1.3 KB/us of the SM's 22 KB/us, and the deeper ring that is slower
    // under full load (HBM read/write turnarounds) is what shortens the launch tail.
    __device__ __forceinline__ void top_up() {               // lane 0
        const int want = (pnext < 0) ? RING : depth;
        while (inflight < want && left > 0) issue_next();
    }
    __device__ __forceinline__ uint32_t wait() {             // returns the shared address of the chunk
        const uint32_t bar = bar0 + cs * 8u;
        while (!mbar_try_wait_a(bar, parity)) {
        }
        return slot0 + cs * (uint32_t)CHUNK_BYTES;
    }
    // BULK variant of release: the lanes have overwritten the chunk in place with the gradient; lane 0 hands the
    // slot to the copy engine (cp.async.bulk shared -> global) and, one release later, re-arms it -- after
    // cp.async.bulk.wait_group.read 1 has confirmed that every store but the newest has left shared memory.
    // Of RING slots one is being consumed, one is draining, RING - 2 are being filled.
    int stores;
    __device__ __forceinline__ void release_store(int lane, void* gdst) {
        __syncwarp();
        if (lane == 0) {
            sp::fence_proxy_async_smem();
            const uint32_t src_slot = slot0 + cs * (uint32_t)CHUNK_BYTES;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(src_slot), "n"(CHUNK_BYTES) : "memory");
            sp::bulk_commit();
            --inflight;
            if (stores++ > 0) {
                sp::bulk_wait_read<1>();
                issue_next();
            }
        }
        if (RING > 1) {
            if (++cs == RING) { cs = 0; parity ^= 1u; }
        } else {
            parity ^= 1u;
        }
    }
    __device__ __forceinline__ void release(int lane) {
        __syncwarp();
        if (lane == 0) {
            sp::fence_proxy_async_smem();
            --inflight;
            top_up();                                        // refills the slot just drained (and more in the tail)
        }
        if (RING > 1) {
            if (++cs == RING) { cs = 0; parity ^= 1u; }
        } else {
            parity ^= 1u;
        }
    }
};

// The predicted-map argmax of one map whose winning quad is still on its way back from L2.
struct PendingAxis {
    int m;              // < 0: nothing pending
    int quad;
    float gmax, mk;
    float4 v;
};

__device__ __forceinline__ void flush_pending(const MapIo& io, PendingAxis& pd, int lane) {
    if (pd.m < 0) return;
    const float a = __fmul_rn(pd.mk, pd.v.x), b = __fmul_rn(pd.mk, pd.v.y), c = __fmul_rn(pd.mk, pd.v.z);
    const int sub = (a == pd.gmax) ? 0 : (b == pd.gmax) ? 1 : (c == pd.gmax) ? 2 : 3;
    if (lane == 0) io.pred_xy[pd.m] = axis_of(pd.gmax, 4 * pd.quad + sub, io.W);
    pd.m = -1;
}

// MODE 0: Gaussian drawn, mask == 1, argmax tracked iff ACC (the common case);
// MODE 1: nothing drawn and mask == 0 (invisible / culled joint): target 0, no tracking;
// MODE 2: anything else (odd mask values, sigma outside the analytic range): run-time flags.
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int QPR, int PPC, int RING, bool ACC, int MODE, bool BULK = false>
__device__ __forceinline__ float tile_map(const MapIo& io, int m, const JointVerdict& jv, const double* ex, const double* ey,
                                          TileRing<RING, PPC * 32 * Tile<QPR>::PERIOD * 16>& rg, PendingAxis& pd, int lane) {
    using T = Tile<QPR>;
    constexpr int PERIOD = T::PERIOD, ROWS = T::ROWS;
    constexpr int CHUNK_QUADS = PPC * 32 * PERIOD;
    const int hw = io.H * io.W;
    const bool draw = (MODE == 0) ? true : (MODE == 1) ? false : jv.draw;
    const bool unit = (MODE == 0);
    const bool track = ACC && ((MODE == 0) ? true : (MODE == 1) ? false : (jv.weight != 0.f));
    const bool analytic_t = track && draw && io.analytic_ok && jv.weight >= 0.5f && jv.weight <= 4.f;
    const bool track_t = (MODE == 2) && track && draw && !analytic_t;
    const float mk = unit ? 1.0f : jv.weight, norm = io.norm, half_scale = io.half_scale;

    // per-lane constants of one period: shared addresses of the row factor and of the x factors
    const uint32_t ex_a = sp::smem_u32(ex), ey_a = sp::smem_u32(ey);
    uint32_t eya[PERIOD], exa[PERIOD];
#pragma unroll
    for (int j = 0; j < PERIOD; ++j) {
        const int q = lane + 32 * j;
        const int y = q / QPR;
        eya[j] = ey_a + 8u * (uint32_t)y;
        exa[j] = ex_a + 16u * (uint32_t)(q - y * QPR);           // (e0, e1) plane; (e2, e3) is 16*QPR bytes further
    }
    double exr[T::EX_IN_REGS ? PERIOD : 1][4];
    if (T::EX_IN_REGS && draw) {
#pragma unroll
        for (int j = 0; j < PERIOD; ++j) {
            const double2 a = lds128d(exa[j]), b = lds128d(exa[j] + 16u * QPR);
            exr[j][0] = a.x; exr[j][1] = a.y; exr[j][2] = b.x; exr[j][3] = b.y;
        }
    }

    float acc = 0.f;
    float best = -CUDART_INF_F, tbest = -CUDART_INF_F;
    int bq = lane, tbq = lane;
    int qlane = lane;                                   // this lane's quad at step 0 of the current chunk
    float4* g4 = reinterpret_cast<float4*>(io.grad + (size_t)m * hw) + lane;
    const int chunks = (hw >> 2) / CHUNK_QUADS;

#pragma unroll 1
    for (int c = 0; c < chunks; ++c) {
        const uint32_t chunk = rg.wait() + 16u * (uint32_t)lane;
#pragma unroll
        for (int it = 0; it < PPC; ++it) {
#pragma unroll
            for (int j = 0; j < PERIOD; ++j) {
                constexpr int kDummy = 0;
                (void)kDummy;
                const int step = it * PERIOD + j;
                const float4 p = lds128(chunk + 512u * step);
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (draw) {
                    const double fy = lds64f(eya[j] + 8u * (uint32_t)(it * ROWS));
                    double e0, e1, e2, e3;
                    if (T::EX_IN_REGS) {
                        e0 = exr[j][0]; e1 = exr[j][1]; e2 = exr[j][2]; e3 = exr[j][3];
                    } else {
                        const double2 a = lds128d(exa[j]), b = lds128d(exa[j] + 16u * QPR);
                        e0 = a.x; e1 = a.y; e2 = b.x; e3 = b.y;
                    }
                    t.x = __double2float_rn(__dmul_rn(e0, fy));
                    t.y = __double2float_rn(__dmul_rn(e1, fy));
                    t.z = __double2float_rn(__dmul_rn(e2, fy));
                    t.w = __double2float_rn(__dmul_rn(e3, fy));
                }
                const float px = unit ? p.x : __fmul_rn(mk, p.x), py = unit ? p.y : __fmul_rn(mk, p.y);
                const float pz = unit ? p.z : __fmul_rn(mk, p.z), pw = unit ? p.w : __fmul_rn(mk, p.w);
                const float tx = unit ? t.x : __fmul_rn(mk, t.x), ty = unit ? t.y : __fmul_rn(mk, t.y);
                const float tz = unit ? t.z : __fmul_rn(mk, t.z), tw = unit ? t.w : __fmul_rn(mk, t.w);
                const float dx = __fsub_rn(px, tx), dy = __fsub_rn(py, ty), dz = __fsub_rn(pz, tz), dw = __fsub_rn(pw, tw);
                acc = fmaf(dx, dx, acc);
                acc = fmaf(dy, dy, acc);
                acc = fmaf(dz, dz, acc);
                acc = fmaf(dw, dw, acc);
                float4 g;
                g.x = __fmul_rn(__fmul_rn(norm, dx), half_scale);
                g.y = __fmul_rn(__fmul_rn(norm, dy), half_scale);
                g.z = __fmul_rn(__fmul_rn(norm, dz), half_scale);
                g.w = __fmul_rn(__fmul_rn(norm, dw), half_scale);
                if (!unit) {
                    g.x = __fmul_rn(g.x, mk); g.y = __fmul_rn(g.y, mk); g.z = __fmul_rn(g.z, mk); g.w = __fmul_rn(g.w, mk);
                }
                if (BULK) sts128(chunk + 512u * step, g);      // in place: this lane has just read these 16 bytes
                else      g4[32 * step] = g;
                if (track) {
                    const float m4 = sp::fmax_nan(sp::fmax_nan(px, py), sp::fmax_nan(pz, pw));
                    if (m4 > best) bq = qlane + 32 * step;
                    best = sp::fmax_nan(best, m4);          // NaN sticks: resolved by the exact scan below
                }
                if (track_t) {
                    const float m4 = fmaxf(fmaxf(tx, ty), fmaxf(tz, tw));
                    if (m4 > tbest) { tbest = m4; tbq = qlane + 32 * step; }
                }
            }
        }
        if (BULK) rg.release_store(lane, g4 - lane);
        else      rg.release(lane);
        g4 += CHUNK_QUADS;
        qlane += CHUNK_QUADS;
#pragma unroll
        for (int j = 0; j < PERIOD; ++j) eya[j] += 8u * (uint32_t)(PPC * ROWS);
    }

    if (ACC) {
        flush_pending(io, pd, lane);                 // the previous map's quad arrived long ago
        float2 lxy = make_float2(0.f, 0.f);
        if (track) {
            const float* src = io.pred + (size_t)m * hw;
            if (__any_sync(SP_FULL, best != best)) {
                float pv;
                int pi;
                MaskedPredView<float> view{src, mk};
                argmax_exact_scan(view, hw, lane, pv, pi);
                if (lane == 0) io.pred_xy[m] = axis_of(pv, pi, io.W);
            } else {
                // every lane kept the first quad holding its own maximum: the smallest quad among
                // the lanes that hold the warp-wide maximum contains torch.max's answer
                const float gmax = warp_max_f32(best);
                const int gq = (int)__reduce_min_sync(SP_FULL, (best == gmax) ? (unsigned)bq : 0x7fffffffu);
                pd.m = m;
                pd.quad = gq;
                pd.gmax = gmax;
                pd.mk = mk;
                pd.v = __ldg(reinterpret_cast<const float4*>(src) + gq);
            }
            if (analytic_t) {
                const int W = io.W;
                const int xn = min(max(__float2int_rn(jv.mx), 0), W - 1), yn = min(max(__float2int_rn(jv.my), 0), io.H - 1);
                const int yy = yn - 1 + lane / 3, xx = xn - 1 + lane % 3;
                const bool in = lane < 9 && yy >= 0 && yy < io.H && xx >= 0 && xx < W;
                float v = -CUDART_INF_F;
                if (in) v = __fmul_rn(mk, __double2float_rn(__dmul_rn(ex[ex_slot<QPR>(xx)], ey[yy])));
                const float gmax = warp_max_f32(v);
                const unsigned gi = __reduce_min_sync(SP_FULL, (in && v == gmax) ? (unsigned)(yy * W + xx) : 0x7fffffffu);
                lxy = axis_of(gmax, (int)gi, W);
            } else if (track_t) {
                const float gmax = warp_max_f32(tbest);
                const int gq = (int)__reduce_min_sync(SP_FULL, (tbest == gmax) ? (unsigned)tbq : 0x7fffffffu);
                const int y = gq / QPR, x4 = 4 * (gq - y * QPR);
                int sub = 3;
#pragma unroll
                for (int e = 2; e >= 0; --e)
                    if (__fmul_rn(mk, __double2float_rn(__dmul_rn(ex[ex_slot<QPR>(x4 + e)], ey[y]))) == gmax) sub = e;
                lxy = axis_of(gmax, 4 * gq + sub, io.W);
            }
        } else if (lane == 0) {
            io.pred_xy[m] = make_float2(0.f, 0.f);
        }
        if (lane == 0) io.label_xy[m] = lxy;
    }
    return acc;
}

// dynamic smem: same layout as variant B
template <int QPR, int PPC, int RING, bool ACC, bool BULK = false>
__global__ void __launch_bounds__(512, 1)
encode_mse_tile_kernel(const MapIo io, float* __restrict__ loss, MseWorkspace* __restrict__ ws, double inv_count, int nwarps,
                       int depth, int static_maps) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int CHUNK_BYTES = PPC * 32 * Tile<QPR>::PERIOD * 16;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
#ifdef SP_TRAIN_TRACE
    long long tr[8] = {gtime(), 0, 0, 0, 0, 0, 0, 0};
#endif
    const int hw = io.H * io.W;
    const int wpad = (io.W + 1) & ~1;
    const size_t fac_bytes = (size_t)(wpad + ((io.H + 1) & ~1)) * sizeof(double);   // multiple of 16 (odd H too)
    volatile int* fifo = reinterpret_cast<int*>(smem_raw + 1024) + warp * kFifo;
    double* ex = reinterpret_cast<double*>(smem_raw + kRingHeader + warp * fac_bytes);
    double* ey = ex + wpad;
    TileRing<RING, CHUNK_BYTES> rg;
    const uint32_t base = sp::smem_u32(smem_raw);
    rg.bar0 = base + (uint32_t)(warp * RING * 8);
    rg.slot0 = base + (uint32_t)(kRingHeader + (size_t)nwarps * fac_bytes + (size_t)warp * RING * CHUNK_BYTES);
    rg.cs = 0; rg.parity = 0; rg.ps = 0; rg.tail = 0; rg.left = 0; rg.pnext = -1; rg.src = nullptr;
    rg.inflight = 0; rg.depth = depth; rg.stores = 0;
    rg.fifo = fifo; rg.next_work = &ws->next_work;
    rg.nmaps = io.nmaps;
    rg.static_stride = (int)gridDim.x * nwarps;
    rg.static_next = (int)blockIdx.x * nwarps + warp;
    rg.static_left = static_maps;
    rg.dyn_base = static_maps * rg.static_stride;
    rg.pred = reinterpret_cast<const char*>(io.pred);
    rg.map_bytes = (size_t)hw * 4;
    rg.chunks_per_map = hw * 4 / CHUNK_BYTES;
    if (lane == 0) {
        for (int r = 0; r < RING; ++r) sp::mbar_init(reinterpret_cast<uint64_t*>(smem_raw) + warp * RING + r, 1);
        sp::mbar_fence_init();
    }
    __syncthreads();
    sp::grid_dep_wait();            // the prologue above overlapped the previous kernel's tail
#ifdef SP_TRAIN_TRACE
    tr[1] = gtime();
#endif

    if (lane == 0) {
        rg.start(rg.claim());
        rg.top_up();
    }
    __syncwarp();
#ifdef SP_TRAIN_TRACE
    if (fifo[0] >= 0) { while (!mbar_try_wait_a(rg.bar0, 0)) { } }
    tr[2] = gtime();
    const long long c0 = clock64();
#endif

    double sum_sq = 0.0;
    PendingAxis pd;
    pd.m = -1; pd.quad = 0; pd.gmax = 0.f; pd.mk = 0.f; pd.v = make_float4(0.f, 0.f, 0.f, 0.f);
    int head = 0;
    int m = fifo[0];
    Joint3 jn = load_joint(io, m >= 0 ? m : io.nmaps);
    while (m >= 0) {
        const Joint3 jc = jn;
        const int m_next = fifo[(head + 1) & (kFifo - 1)];   // claimed before this map's first chunk was issued
        jn = load_joint(io, m_next >= 0 ? m_next : io.nmaps);
        const JointVerdict jv = prepare_map<QPR>(io, m, jc, ex, ey, lane);
        float acc;
        if (jv.draw && jv.weight == 1.0f && (!ACC || io.analytic_ok))
            acc = tile_map<QPR, PPC, RING, ACC, 0, BULK>(io, m, jv, ex, ey, rg, pd, lane);
        else if (!jv.draw && jv.weight == 0.0f)
            acc = tile_map<QPR, PPC, RING, ACC, 1, BULK>(io, m, jv, ex, ey, rg, pd, lane);
        else
            acc = tile_map<QPR, PPC, RING, ACC, 2, BULK>(io, m, jv, ex, ey, rg, pd, lane);
        sum_sq += (double)acc;
        __syncwarp();
        ++head;
        m = m_next;
    }
    if (ACC) flush_pending(io, pd, lane);
    if (BULK && lane == 0) sp::bulk_wait_all<0>();           // every gradient byte of this warp is in global memory
#ifdef SP_TRAIN_TRACE
    tr[3] = gtime(); tr[5] = head; tr[6] = clock64() - c0;
    { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); tr[7] = smid; }
#endif
    finish_loss<512>(sum_sq, ws, loss, inv_count);
#ifdef SP_TRAIN_TRACE
    tr[4] = gtime();
    if (lane == 0 && g_trace_ptr) {
        long long* t = g_trace_ptr + ((size_t)blockIdx.x * 16 + warp) * 8;
        for (int i = 0; i < 8; ++i) t[i] = tr[i];
    }
#endif
}

// HeatMapAcc epilogue (metrics/pose_metrics.py:227-245) on the [B,K] argmax coordinates.
__global__ void __launch_bounds__(256)
heatmap_acc_kernel(const float2* __restrict__ pred_xy, const float2* __restrict__ label_xy, float* __restrict__ acc,
                   int B, int K, float norm_x, float norm_y, float thresh) {
    extern __shared__ int counters[];        // hit[K], valid[K]
    sp::grid_dep_wait();
    int* hit = counters;
    int* valid = counters + K;
    for (int k = threadIdx.x; k < 2 * K; k += blockDim.x) counters[k] = 0;
    __syncthreads();
    const int n = B * K;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float2 l = label_xy[i];
        if (l.x > 1.f && l.y > 1.f) {
            const float2 p = pred_xy[i];
            const float dx = __fsub_rn(__fdiv_rn(p.x, norm_x), __fdiv_rn(l.x, norm_x));
            const float dy = __fsub_rn(__fdiv_rn(p.y, norm_y), __fdiv_rn(l.y, norm_y));
            // torch.norm accumulates float inputs in double on CPU (the pinned oracle); no (dx, dy)
            // pair of the supported map sizes lands on the threshold, so float vs double is moot
            const float dist = (float)sqrt((double)dx * (double)dx + (double)dy * (double)dy);
            const int k = i % K;
            atomicAdd(&valid[k], 1);
            if (dist < thresh) atomicAdd(&hit[k], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float sum = 0.f;
        int used = 0;
        for (int k = 0; k < K; ++k) {
            if (valid[k] < 1) continue;
            sum = __fadd_rn(sum, __fdiv_rn((float)hit[k], (float)valid[k]));
            ++used;
        }
        acc[0] = used > 0 ? __fdiv_rn(sum, (float)used) : 0.f;
    }
}

}  // namespace

static int encode_mse_launch(const float* joints, const void* pred_raw, int pred_dtype, void* grad_raw, float* targets,
                             float* weights, float* loss, float* pred_xy, float* label_xy,
                             void* workspace, size_t workspace_bytes,
                             int B, int K, int H, int W, double sigma, float grad_scale, const float* scale_dev, void* stream) {
    const float* pred = static_cast<const float*>(pred_raw);      // reinterpreted per dtype inside the generic kernel
    float* grad = static_cast<float*>(grad_raw);
    const bool plain_f32 = (pred_dtype == SP_DTYPE_F32) && !scale_dev && loss;
    SP_RETURN_IF(pred_dtype != SP_DTYPE_F32 && pred_dtype != SP_DTYPE_F16 && pred_dtype != SP_DTYPE_BF16, SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF(!joints || !pred || (!loss && !grad) || !workspace, SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF(B <= 0 || K <= 0 || H <= 0 || W <= 0 || !(sigma > 0.0), SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF((pred_xy == nullptr) != (label_xy == nullptr), SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF((long long)B * K > 0x7fffffffLL || (long long)H * W > (1 << 24), SP_ERR_UNSUPPORTED);
    SP_RETURN_IF(W % 4 != 0, SP_ERR_UNSUPPORTED);          // callers compose sp_encode_f32 + sp_mse_fwd_bwd_f32 instead
    SP_RETURN_IF(workspace_bytes < sizeof(MseWorkspace), SP_ERR_WORKSPACE);
    SP_RETURN_IF(!sp_aligned16(workspace) || !sp_aligned16(pred) || (grad && !sp_aligned16(grad)) ||
                 (targets && !sp_aligned16(targets)) || (pred_xy && (!sp_aligned16(pred_xy) || !sp_aligned16(label_xy))),
                 SP_ERR_BAD_ALIGNMENT);
    const int nmaps = B * K;
    const float reach = (float)(sigma * 3.0);
    const double denom = 2.0 * (sigma * sigma);
    const double count = (double)B * (double)K * (double)H * (double)W;
    const float norm = (float)(2.0 / count);
    const float half_scale = 0.5f * grad_scale;
    const int wpad = (W + 1) & ~1;
    const size_t fac_bytes = (size_t)(wpad + ((H + 1) & ~1)) * sizeof(double);          // multiple of 16 (odd H too)
    MapIo io;
    io.joints = joints; io.pred = pred; io.grad = grad; io.targets = targets; io.weights = weights;
    io.pred_xy = reinterpret_cast<float2*>(pred_xy); io.label_xy = reinterpret_cast<float2*>(label_xy);
    io.nmaps = nmaps; io.H = H; io.W = W; io.reach = reach; io.denom = denom; io.norm = norm; io.half_scale = half_scale;
    io.scale_dev = scale_dev;
    io.analytic_ok = (sigma >= 0.25 && sigma <= 64.0) ? 1 : 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MseWorkspace* ws = static_cast<MseWorkspace*>(workspace);
    const int sel = (grad ? 4 : 0) | (targets ? 2 : 0) | (pred_xy ? 1 : 0);

    // Ring variant: needs the map to split into equal chunks of a multiple of 32 quads (<= 4 KB)
    const int nq = (H * W) >> 2;
    const SpTuning& tune = sp_tuning();
    const bool force_ldg = sp_knob(tune.train_force_ldg, 0) == 1;
    int chunk_quads = 0;
    if ((H * W) % 4 == 0 && nq % 32 == 0) {
        int want = sp_knob(tune.train_chunk_quads, 192);  // 3 KB
        if (want < 32) want = 32;
        for (int d = want - want % 32; d >= 32; d -= 32)
            if (nq % d == 0) { chunk_quads = d; break; }
    }
    // Variant C: W = 48 / 72 (12 / 18 quads per row), whole periods per map, grad wanted, targets not
    const int qpr = W >> 2;
    if (plain_f32 && (qpr == 12 || qpr == 18) && grad && !targets && !force_ldg && sp_knob(tune.train_no_tile, 0) == 0) {
        const int period = (qpr == 12) ? Tile<12>::PERIOD : Tile<18>::PERIOD;
        const int rows = (qpr == 12) ? Tile<12>::ROWS : Tile<18>::ROWS;
        // tuned layouts (PPC periods per chunk, ring depth); SP_TRAIN_TILE_CFG picks another compiled one
        const int cfg = sp_knob(tune.train_tile_cfg, 0);
        int ppc = 1, ring = 2;
        if (qpr == 12) {
            if (cfg == 1) { ppc = 1; ring = 2; } else if (cfg == 2) { ppc = 2; ring = 2; } else if (cfg == 3) { ppc = 4; ring = 1; }
            else if (cfg == 4) { ppc = 2; ring = 3; } else if (cfg == 5) { ppc = 2; ring = 4; } else { ppc = 2; ring = 1; }
        } else {
            if (cfg == 1) { ppc = 1; ring = 1; } else if (cfg == 2) { ppc = 1; ring = 3; } else if (cfg == 3) { ppc = 1; ring = 4; } else { ppc = 1; ring = 2; }
        }
        // copies kept in flight per warp in steady state (the whole ring is used once a warp is on its last map)
        int depth = sp_knob(tune.train_depth, ring);
        if (depth < 1) depth = 1;
        if (depth > ring) depth = ring;
        const int periods = nq / (32 * period);
        if (H % rows == 0 && periods % ppc == 0) {
            const size_t chunk_bytes = (size_t)ppc * 32 * period * 16;
            const size_t budget = 226 * 1024 - kRingHeader;
            int nwarps = (int)(budget / (fac_bytes + ring * chunk_bytes));
            if (nwarps > 16) nwarps = 16;
            // 96x72 maps are 27 KB: with 16 warps a map takes ~20 us and the launch ends with a long,
            // thin tail; 8 warps (74 KB in flight per SM) halve it (512 persons: 88 -> 83 us) once there
            // are enough maps to keep the warps fed. Small launches keep 16 warps (more maps in flight).
            int want_warps = (qpr == 18 && nmaps >= 2 * 16 * sp_sm_count()) ? 8 : 16;
            want_warps = sp_knob(tune.train_warps, want_warps);
            if (want_warps < nwarps) nwarps = want_warps;
            if (nwarps >= 1) {
                const size_t smem = kRingHeader + (size_t)nwarps * (fac_bytes + ring * chunk_bytes);
                int grid = sp_sm_count();
                const int need = (nmaps + nwarps - 1) / nwarps;
                if (grid > need) grid = need;
                // maps per warp that are assigned up front (no atomic); the rest are dealt from the grid-wide
                // counter. All atomics hit one address (~3.5 ns each on B200), so the dynamic share is kept to
                // what the spread of SM speeds needs: half of a warp's expected maps, at least 2 fixed.
                int static_maps = (int)((long long)nmaps * sp_knob(tune.train_static_pct, 50) / 100 / ((long long)grid * nwarps));
                if (static_maps < 2) static_maps = 2;
#define SP_LAUNCH_TILE(Q, P, R)                                                                                              \
    do {                                                                                                                     \
        if (pred_xy) {                                                                                                       \
            SP_CUDA(sp_launch_smem(encode_mse_tile_kernel<Q, P, R, true>, dim3(grid), dim3(nwarps * 32), smem, st, io, loss, ws, 1.0 / count, nwarps, depth, static_maps)); \
        } else {                                                                                                             \
            SP_CUDA(sp_launch_smem(encode_mse_tile_kernel<Q, P, R, false>, dim3(grid), dim3(nwarps * 32), smem, st, io, loss, ws, 1.0 / count, nwarps, depth, static_maps)); \
        }                                                                                                                    \
    } while (0)
                // SP_TRAIN_BULK_STORE=1: the gradient leaves through the TMA too (ring >= 2; the start-up depth is the whole ring)
                const bool bulk = sp_knob(tune.train_bulk_store, 0) == 1 && ring >= 2 && ((qpr == 12 && ppc == 2 && ring <= 4) || (qpr == 18 && ring <= 3));
                if (bulk) {
#define SP_LAUNCH_TILE_BULK(Q, P, R)                                                                                         \
    do {                                                                                                                     \
        if (pred_xy) SP_CUDA(sp_launch_smem(encode_mse_tile_kernel<Q, P, R, true, true>, dim3(grid), dim3(nwarps * 32), smem, st, io, loss, ws, 1.0 / count, nwarps, R, static_maps)); \
        else         SP_CUDA(sp_launch_smem(encode_mse_tile_kernel<Q, P, R, false, true>, dim3(grid), dim3(nwarps * 32), smem, st, io, loss, ws, 1.0 / count, nwarps, R, static_maps)); \
    } while (0)
                    if (qpr == 12) { if (ring == 2) SP_LAUNCH_TILE_BULK(12, 2, 2); else if (ring == 3) SP_LAUNCH_TILE_BULK(12, 2, 3); else SP_LAUNCH_TILE_BULK(12, 2, 4); }
                    else           { if (ring == 2) SP_LAUNCH_TILE_BULK(18, 1, 2); else SP_LAUNCH_TILE_BULK(18, 1, 3); }
#undef SP_LAUNCH_TILE_BULK
                    return 0;
                }
                if (qpr == 12) {
                    if (ppc == 1) SP_LAUNCH_TILE(12, 1, 2); else if (ppc == 2 && ring == 2) SP_LAUNCH_TILE(12, 2, 2);
                    else if (ppc == 2 && ring == 3) SP_LAUNCH_TILE(12, 2, 3); else if (ppc == 2 && ring == 4) SP_LAUNCH_TILE(12, 2, 4);
                    else if (ppc == 4) SP_LAUNCH_TILE(12, 4, 1); else SP_LAUNCH_TILE(12, 2, 1);
                } else {
                    if (ring == 1) SP_LAUNCH_TILE(18, 1, 1); else if (ring == 3) SP_LAUNCH_TILE(18, 1, 3);
                    else if (ring == 4) SP_LAUNCH_TILE(18, 1, 4); else SP_LAUNCH_TILE(18, 1, 2);
                }
#undef SP_LAUNCH_TILE
                return 0;
            }
        }
    }
    if (plain_f32 && chunk_quads > 0 && !force_ldg) {
        const size_t chunk_bytes = (size_t)chunk_quads * 16;
        const size_t budget = 226 * 1024 - kRingHeader;        // 1 KB spare for static shared memory
        int ring = sp_knob(tune.train_ring, 2);
        if (ring < 1) ring = 1;
        if (ring > 8) ring = 8;
        int nwarps = (int)(budget / (fac_bytes + ring * chunk_bytes));
        if (nwarps > 16) nwarps = 16;      // 16 x 2 x 3 KB = 96 KB in flight per SM; more warps = more HBM streams = slower
        { const int w = sp_knob(tune.train_warps, 16); if (w < nwarps) nwarps = w; }
        if (nwarps < 1) nwarps = 1;
        SP_RETURN_IF((size_t)nwarps * ring * 8 > 1024, SP_ERR_UNSUPPORTED);
        const size_t smem = kRingHeader + (size_t)nwarps * (fac_bytes + ring * chunk_bytes);
        int grid = sp_sm_count();
        const int need = (nmaps + nwarps - 1) / nwarps;
        if (grid > need) grid = need;
#define SP_LAUNCH_RING(G, T, A)                                                                                              \
    do {                                                                                                                     \
        SP_CUDA(sp_launch_smem(encode_mse_ring_kernel<G, T, A>, dim3(grid), dim3(nwarps * 32), smem, st, io, loss, ws, 1.0 / count, nwarps, ring, chunk_quads)); \
    } while (0)
        switch (sel) {
            case 0: SP_LAUNCH_RING(false, false, false); break;
            case 1: SP_LAUNCH_RING(false, false, true); break;
            case 2: SP_LAUNCH_RING(false, true, false); break;
            case 3: SP_LAUNCH_RING(false, true, true); break;
            case 4: SP_LAUNCH_RING(true, false, false); break;
            case 5: SP_LAUNCH_RING(true, false, true); break;
            case 6: SP_LAUNCH_RING(true, true, false); break;
            default: SP_LAUNCH_RING(true, true, true); break;
        }
#undef SP_LAUNCH_RING
        return 0;
    }

    const size_t smem = (size_t)kWarps * fac_bytes;
    SP_RETURN_IF(smem > 200 * 1024, SP_ERR_UNSUPPORTED);
    int grid = (nmaps + kWarps - 1) / kWarps;
    if (grid > kMaxPartials) grid = kMaxPartials;
#define SP_LAUNCH_TRAIN(G, T, A)                                                                                         \
    do {                                                                                                                 \
        if (pred_dtype == SP_DTYPE_F16)                                                                                  \
            SP_CUDA(sp_launch_smem(encode_mse_kernel<G, T, A, __half>, dim3(grid), dim3(kThreads), smem, st, io, loss, ws, 1.0 / count)); \
        else if (pred_dtype == SP_DTYPE_BF16)                                                                            \
            SP_CUDA(sp_launch_smem(encode_mse_kernel<G, T, A, __nv_bfloat16>, dim3(grid), dim3(kThreads), smem, st, io, loss, ws, 1.0 / count)); \
        else                                                                                                             \
            SP_CUDA(sp_launch_smem(encode_mse_kernel<G, T, A, float>, dim3(grid), dim3(kThreads), smem, st, io, loss, ws, 1.0 / count)); \
    } while (0)
    switch (sel) {
        case 0: SP_LAUNCH_TRAIN(false, false, false); break;
        case 1: SP_LAUNCH_TRAIN(false, false, true); break;
        case 2: SP_LAUNCH_TRAIN(false, true, false); break;
        case 3: SP_LAUNCH_TRAIN(false, true, true); break;
        case 4: SP_LAUNCH_TRAIN(true, false, false); break;
        case 5: SP_LAUNCH_TRAIN(true, false, true); break;
        case 6: SP_LAUNCH_TRAIN(true, true, false); break;
        default: SP_LAUNCH_TRAIN(true, true, true); break;
    }
#undef SP_LAUNCH_TRAIN
    return 0;
}

extern "C" int sp_encode_mse_fwd_bwd_f32(const float* joints, const float* pred, float* grad, float* targets,
                                         float* weights, float* loss, float* pred_xy, float* label_xy,
                                         void* workspace, size_t workspace_bytes,
                                         int B, int K, int H, int W, double sigma, float grad_scale, void* stream) {
    SP_RETURN_IF(!loss, SP_ERR_BAD_ARGUMENT);
    return encode_mse_launch(joints, pred, SP_DTYPE_F32, grad, targets, weights, loss, pred_xy, label_xy, workspace, workspace_bytes,
                             B, K, H, W, sigma, grad_scale, nullptr, stream);
}

extern "C" int sp_encode_mse_fwd_bwd(const float* joints, const void* pred, int pred_dtype, void* grad, float* targets,
                                     float* weights, float* loss, float* pred_xy, float* label_xy,
                                     void* workspace, size_t workspace_bytes,
                                     int B, int K, int H, int W, double sigma, float grad_scale, const float* grad_scale_dev,
                                     void* stream) {
    return encode_mse_launch(joints, pred, pred_dtype, grad, targets, weights, loss, pred_xy, label_xy, workspace, workspace_bytes,
                             B, K, H, W, sigma, grad_scale, grad_scale_dev, stream);
}

extern "C" int sp_heatmap_acc_f32(const float* pred_xy, const float* label_xy, float* acc,
                                  int B, int K, int H, int W, float distance_thresh, float norm_frac, void* stream) {
    SP_RETURN_IF(!pred_xy || !label_xy || !acc, SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF(B <= 0 || K <= 0 || H <= 0 || W <= 0 || K > 4096, SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF((long long)B * K > 0x7fffffffLL, SP_ERR_UNSUPPORTED);
    SP_RETURN_IF(!sp_aligned16(pred_xy) || !sp_aligned16(label_xy), SP_ERR_BAD_ALIGNMENT);
    // norm = tensor([W, H], float32) / norm_frac
    const float nx = (float)W / norm_frac, ny = (float)H / norm_frac;
    SP_CUDA(sp_launch(heatmap_acc_kernel, dim3(1), dim3(256), (size_t)2 * K * sizeof(int), static_cast<cudaStream_t>(stream),
                      reinterpret_cast<const float2*>(pred_xy), reinterpret_cast<const float2*>(label_xy), acc, B, K, nx, ny,
                      distance_thresh));
    return 0;
}
