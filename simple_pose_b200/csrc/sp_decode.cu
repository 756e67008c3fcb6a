// A3/A4/A5/A6: heatmap decoders of the reference's metrics/pose_metrics.py, one fused kernel.
//
//   GAUSS_TAYLOR  GaussTaylorKeyPointDecoder.__call__ (:62-107): argmax on the raw map, 11x11
//                 Gaussian blur, rescale by ori_max/blur_max, clamp 1e-10, log, second-order
//                 Taylor step at the peak, clamp(min=0), affine back-projection.
//   ARGMAX        BasicKeyPointDecoder.heat_map_to_axis (:11-24).
//   BASIC         BasicKeyPointDecoder.__call__ (:26-52): quarter-pixel shift toward the larger
//                 neighbour.
//   flip-test     optional second input: the decoded map is 0.5*(hm + mirror/swap(hm_flip)).
//
// HBM-bound: K*H*W*4 bytes read per person (twice that with the flip input), 12 bytes written
// per joint. Design (persistent, one CTA per SM):
//   * every warp owns a private ring of shared-memory stages and streams whole (person, joint)
//     maps into it with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx); lane 0
//     re-arms the stage for the warp's next map as soon as the warp is done with it, so the
//     copy of map i+1 overlaps the arithmetic on map i and a full SM keeps ~100-200 KB in flight;
//   * argmax: one 16-byte shared load per lane per step, per-quad NaN-propagating max, a
//     (value, quad) warp butterfly with torch.max tie rules (first index wins), and an exact
//     scalar rescan only if a NaN/Inf was seen;
//   * the blur is NOT applied to the whole map. The Taylor step reads the log-blurred map at 13
//     stencil points only, so the 11x11 window is evaluated at those 13 points from a
//     zero-padded 15x15 patch (1573 FMAs per joint instead of 371 712). The reference's
//     ori_max/blur_max factor multiplies every stencil value by the same constant, which
//     cancels in every finite difference of the logs, *provided the 1e-10 clamp does not fire*.
//     Since blur_max <= ori_max * sum(w) up to float32 rounding the factor is >= 1 - 1e-5, so when all 13 blurred
//     values are >= 2e-10 the clamp provably cannot fire and the factor is dropped; if they are
//     all <= 0 every stencil value is clamped to the same constant and the Hessian test
//     (det != 0) rejects the joint, as in the reference; the remaining (mixed) case takes an
//     exact slow path that blurs the whole map of that joint to obtain blur_max;
//   * no tensor cores: nothing here is a dense contraction.
#include "sp_common.cuh"
#include "sp_decode_dev.cuh"
#include <math_constants.h>

#ifdef SP_TRAIN_TRACE
// scratch instrumentation (never compiled into the product library): per-CTA [release, finish, smid, maps]
__device__ long long* g_decode_trace_ptr = nullptr;
extern "C" int sp_debug_set_decode_trace(void* p) {
    return (int)cudaMemcpyToSymbol(g_decode_trace_ptr, &p, sizeof(p));
}
__device__ __forceinline__ long long decode_gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

namespace {

using namespace sp_dec;

// ---------------------------------------------------------------------------------------------
// fast path: persistent CTAs, per-warp TMA rings
// ---------------------------------------------------------------------------------------------
// dynamic shared memory layout:
//   [0, 512)                       mbarriers (nwarps * stages * 8 B)
//   [512, 1016)                    map index held by each (warp, stage); [1016] CTA work counter
//   [1024, 1024 + 1024)            blur weights (<= 225 floats)
//   then per warp                  patch (kPatchFloats floats, padded to 1536 B)
//   then per warp, per stage       map a (hw floats) [+ map b if FLIP]
constexpr int kBarBytes = 1024;
constexpr int kWtsBytes = 1024;
constexpr int kPatchBytes = 1536;

// KS = compile-time blur size (11 = the reference's), 0 = runtime A.ksize
//
// Work distribution: CTA c owns the contiguous map range [c*nmaps/grid, (c+1)*nmaps/grid) (sizes
// differ by at most one map across SMs) and its warps pull maps from a shared-memory counter, so
// the tail of a launch is one map per warp instead of a whole stride of the grid.
template <bool FLIP, int KS>
__global__ void __launch_bounds__(512, 1)
decode_tma_kernel(const DecodeArgs A, int nwarps, int stages, int static_maps) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int hw = A.H * A.W;
    const uint32_t map_bytes = (uint32_t)hw * 4u;
    const uint32_t stage_bytes = FLIP ? 2u * map_bytes : map_bytes;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + warp * stages;
    int* claimed = reinterpret_cast<int*>(smem + kBarBytes / 2) + warp * stages;    // map held by each stage
    int& next_map = *reinterpret_cast<int*>(smem + kBarBytes - 8);                  // CTA work counter
    float* wts = reinterpret_cast<float*>(smem + kBarBytes);
    float* patch = reinterpret_cast<float*>(smem + kBarBytes + kWtsBytes + (size_t)warp * kPatchBytes);
    unsigned char* ring = smem + kBarBytes + kWtsBytes + (size_t)nwarps * kPatchBytes +
                          (size_t)warp * stages * stage_bytes;

    const int range_lo = (int)((long long)blockIdx.x * A.nmaps / gridDim.x);
    const int range_hi = (int)((long long)(blockIdx.x + 1) * A.nmaps / gridDim.x);
    if (threadIdx.x == 0) next_map = range_lo;
    if (lane == 0) {
        for (int s = 0; s < stages; ++s) sp::mbar_init(bars + s, 1);
        sp::mbar_fence_init();
    }
    __syncthreads();                // work counter and barriers are visible to every warp
    sp::grid_dep_wait();            // everything above overlapped the previous kernel's tail
#ifdef SP_TRAIN_TRACE
    const long long trace_t0 = decode_gtime();
    int trace_maps = 0;
#endif

    // Work distribution. Without a workspace: the CTA's own range, claimed from shared memory. With one
    // (sp_decode_ws_f32): maps are dealt GRID-WIDE -- warp w of CTA c owns maps k*grid*nwarps + c*nwarps + w for
    // k < static_maps (>= stages; interleaved across the whole grid, no atomics), the rest come from the counter in
    // the workspace -- because the SMs do not drain HBM at equal rates once the memory system queues
    // (profiles/r1f_fused_timeline.md) and equal ranges make the slowest SM the critical path. The global atomic
    // is issued one map ahead (`ahead` holds its raw result while the current map is processed), so its round trip
    // is never waited for.
    const bool grid_wide = (A.work != nullptr);
    const int round = (int)gridDim.x * nwarps;
    unsigned int ahead = 0;
    long long static_next = (long long)stages * round + (long long)blockIdx.x * nwarps + warp;
    int static_left = static_maps - stages;
    auto claim_grid = [&]() {                       // lane 0, grid-wide mode: this warp's next map or -1
        long long m;
        if (static_left > 0) {
            m = static_next;
            static_next += round;
            if (--static_left == 0) ahead = atomicAdd(A.work, 1u);          // the first dynamic claim, one map ahead
        } else {
            m = (long long)static_maps * round + (long long)ahead;
            if (m < A.nmaps) ahead = atomicAdd(A.work, 1u);
        }
        return (m < A.nmaps) ? (int)m : -1;
    };
    auto issue = [&](int s, int m) {                // lane 0: start the copy of map m into stage s (or park -1)
        if (m < 0) {
            claimed[s] = -1;
            return;
        }
        claimed[s] = m;
        float* dst = reinterpret_cast<float*>(ring + (size_t)s * stage_bytes);
        sp::mbar_expect_tx(bars + s, stage_bytes);
        sp::bulk_g2s(dst, A.hm + (size_t)m * hw, map_bytes, bars + s);
        if (FLIP) {
            const int b = m / A.K, k = m - b * A.K;
            const int src = b * A.K + __ldg(A.perm + k);
            sp::bulk_g2s(dst + hw, A.hm_flip + (size_t)src * hw, map_bytes, bars + s);
        }
    };
    auto claim_local = [&]() {
        const int m = atomicAdd(&next_map, 1);
        return (m < range_hi) ? m : -1;
    };

    // first copies go out before anything else touches global memory: the blur weights (a dependent
    // global load + block barrier, ~0.7 us) are fetched while the first maps are in flight
    if (lane == 0) {
        for (int s = 0; s < stages; ++s) {
            if (grid_wide) {
                const long long m = (long long)s * round + (long long)blockIdx.x * nwarps + warp;
                issue(s, m < A.nmaps ? (int)m : -1);
            } else {
                issue(s, claim_local());
            }
        }
        if (grid_wide && static_left <= 0) ahead = atomicAdd(A.work, 1u);
    }
    if (A.mode == SP_DECODE_GAUSS_TAYLOR || A.mode == SP_DECODE_DARK_ORIGINAL)
        for (int t = threadIdx.x; t < A.ksize * A.ksize; t += blockDim.x) wts[t] = __ldg(A.blur_w + t);
    __syncthreads();
    LaneTaps<(KS > 0 ? KS : 3)> taps;
    if (KS > 0 && (A.mode == SP_DECODE_GAUSS_TAYLOR || A.mode == SP_DECODE_DARK_ORIGINAL)) taps.load(wts, lane);
    int s = 0;
    uint32_t parity = 0;
    for (;;) {
        __syncwarp();
        const int m = *reinterpret_cast<volatile int*>(claimed + s);
        if (m < 0) break;                                  // claims are handed out in order: nothing follows
        const Affine T = load_affine(A, m);
        sp::mbar_wait(bars + s, parity);
        float* a = reinterpret_cast<float*>(ring + (size_t)s * stage_bytes);
        const Peak pk = argmax_smem<FLIP>(a, a + hw, hw, A.W, lane);
        DirectView view{a};
        finish_map(A, view, m, pk, lane, T, [&](int px, int py, float ori_max, float& ox, float& oy) {
            if (KS > 0) return taylor_refine_smem<(KS > 0 ? KS : 3)>(a, wts, patch, taps, A.H, A.W, px, py, ori_max, lane, ox, oy);
            return taylor_refine_generic(view, wts, patch, A.H, A.W, A.ksize, px, py, ori_max, lane, ox, oy);
        });
        __syncwarp();
        if (lane == 0) {
            sp::fence_proxy_async_smem();
            issue(s, grid_wide ? claim_grid() : claim_local());
        }
        if (++s == stages) { s = 0; parity ^= 1u; }
#ifdef SP_TRAIN_TRACE
        ++trace_maps;
#endif
    }
#ifdef SP_TRAIN_TRACE
    if (lane == 0 && g_decode_trace_ptr) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        long long* t = g_decode_trace_ptr + ((size_t)blockIdx.x * 16 + warp) * 4;
        t[0] = trace_t0; t[1] = decode_gtime(); t[2] = smid; t[3] = trace_maps;
    }
#endif
    if (grid_wide) {                 // the last CTA to finish restores the workspace's zero state
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(A.work + 1, 1u) == gridDim.x - 1) {
                A.work[0] = 0u;
                A.work[1] = 0u;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// generic path: any H, W; reads straight from global memory (L1/L2 serve the re-reads)
// ---------------------------------------------------------------------------------------------
template <bool FLIP>
__global__ void __launch_bounds__(256)
decode_generic_kernel(const DecodeArgs A) {
    __shared__ float wts[kMaxKsize * kMaxKsize];
    __shared__ float patches[8][kPatchFloats];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int hw = A.H * A.W;
    sp::grid_dep_wait();
    if (A.mode == SP_DECODE_GAUSS_TAYLOR || A.mode == SP_DECODE_DARK_ORIGINAL)
        for (int t = threadIdx.x; t < A.ksize * A.ksize; t += blockDim.x) wts[t] = __ldg(A.blur_w + t);
    __syncthreads();
    const int total = gridDim.x * 8;
    for (int m = blockIdx.x * 8 + warp; m < A.nmaps; m += total) {
        const Affine T = load_affine(A, m);
        if (FLIP) {
            const int b = m / A.K, k = m - b * A.K;
            FlipAvgView view{A.hm + (size_t)m * hw, A.hm_flip + (size_t)(b * A.K + __ldg(A.perm + k)) * hw, A.W};
            const Peak pk = argmax_exact(view, hw, lane);
            finish_map(A, view, m, pk, lane, T, [&](int px, int py, float ori_max, float& ox, float& oy) {
                return taylor_refine_generic(view, wts, patches[warp], A.H, A.W, A.ksize, px, py, ori_max, lane, ox, oy);
            });
        } else {
            DirectView view{A.hm + (size_t)m * hw};
            const Peak pk = argmax_exact(view, hw, lane);
            finish_map(A, view, m, pk, lane, T, [&](int px, int py, float ori_max, float& ox, float& oy) {
                return taylor_refine_generic(view, wts, patches[warp], A.H, A.W, A.ksize, px, py, ori_max, lane, ox, oy);
            });
        }
    }
}

}  // namespace

extern "C" size_t sp_decode_workspace_bytes(void) { return 16; }

static int decode_launch(const float* hm, const float* hm_flip, const int* perm,
                         const float* trans_inv, const float* blur_w,
                         float* coords, float* maxval, int* argmax, float* rows, int row_stride,
                         int B, int K, int H, int W, int ksize, int mode, unsigned int* work, void* stream) {
    SP_RETURN_IF(B < 0 || K <= 0 || H <= 0 || W <= 0, SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF(B > 0 && !hm, SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF(B > 0 && !rows && (!coords || !maxval), SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF(rows && row_stride < 3 * K, SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF(mode < SP_DECODE_GAUSS_TAYLOR || mode > SP_DECODE_DARK_ORIGINAL, SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF(hm_flip && !perm, SP_ERR_BAD_ARGUMENT);
    if (mode == SP_DECODE_GAUSS_TAYLOR || mode == SP_DECODE_DARK_ORIGINAL) {
        SP_RETURN_IF(!blur_w, SP_ERR_BAD_ARGUMENT);
        SP_RETURN_IF(ksize < 3 || ksize > kMaxKsize || (ksize & 1) == 0, SP_ERR_UNSUPPORTED);
    } else {
        ksize = 0;
    }
    SP_RETURN_IF((long long)B * K > 0x7fffffffLL || (long long)H * W > (1 << 24), SP_ERR_UNSUPPORTED);
    if (B == 0) return 0;
    SP_RETURN_IF(coords && !sp_aligned16(coords), SP_ERR_BAD_ALIGNMENT);

    DecodeArgs A;
    A.hm = hm; A.hm_flip = hm_flip; A.perm = perm; A.trans_inv = trans_inv; A.blur_w = blur_w;
    A.coords = coords; A.maxval = maxval; A.argmax = argmax; A.rows = rows; A.row_stride = row_stride;
    A.nmaps = B * K; A.K = K; A.H = H; A.W = W; A.ksize = ksize; A.mode = mode;
    A.work = work;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool flip = hm_flip != nullptr;

    // fast path needs 16-byte bulk copies and row-aligned quads
    const SpTuning& tune = sp_tuning();
    const size_t map_bytes = (size_t)H * W * 4;
    const size_t stage_bytes = flip ? 2 * map_bytes : map_bytes;
    const size_t budget = 227 * 1024 - kBarBytes - kWtsBytes;
    bool fast = (W % 4 == 0) && sp_aligned16(hm) && (!flip || sp_aligned16(hm_flip)) &&
                (stage_bytes + kPatchBytes <= budget) && sp_knob(tune.decode_force_generic, 0) == 0;
    if (fast) {
        // as many warps as fit (<= 16), then as many stages per warp as still fit (<= 4)
        int nwarps = (int)(budget / (stage_bytes + kPatchBytes));
        if (nwarps > 16) nwarps = 16;
        // large launches of small maps: 14 warps (172 KB in flight per SM) measured 2 % faster than 16
        // (196 KB); small launches keep 16 so that every map is in flight at once
        if (nwarps > 14 && (long long)A.nmaps >= 4LL * 16 * sp_sm_count()) nwarps = 14;
        nwarps = sp_knob(tune.decode_warps, nwarps);
        if (nwarps < 1) nwarps = 1;
        if (nwarps > 16) nwarps = 16;
        int stages = (int)((budget / nwarps - kPatchBytes) / stage_bytes);
        if (stages > 4) stages = 4;
        stages = sp_knob(tune.decode_stages, stages);
        if (stages < 1) stages = 1;
        while (nwarps * stages > 64 && stages > 1) --stages;       // mbarrier / claim tables hold 64 entries
        while ((size_t)nwarps * (stages * stage_bytes + kPatchBytes) > budget && stages > 1) --stages;
        while ((size_t)nwarps * (stages * stage_bytes + kPatchBytes) > budget && nwarps > 1) --nwarps;
        const size_t smem = kBarBytes + kWtsBytes + (size_t)nwarps * (kPatchBytes + stages * stage_bytes);
        int grid = sp_sm_count();
        const int need = (A.nmaps + nwarps - 1) / nwarps;
        if (grid > need) grid = need;
        // Work distribution (needs the caller's workspace, sp_decode_ws_f32 / sp_decode_rows_f32):
        //  * launches of 4 ... 256 maps per warp slot: maps dealt GRID-WIDE, 85 % of a warp's share interleaved statically
        //    (no atomic), the tail claimed from the workspace counter one map ahead. Against equal per-CTA ranges:
        //    1024 x 64x48 36.3 -> 36.0 us, flip 67.8 -> 66.4 us (0.986 of the HBM peak); 4096 x 64x48 130.7 -> 127.2 us;
        //    512 x 96x72 40.1 -> 39.5 us, flip 83.3 -> 80.8 us. Without the static share every map costs a global atomic
        //    on one address (~3 ns each), which loses for 12 KB items (64x48: 36 -> 39 us; 4096 persons: 131 -> 203 us);
        //  * smaller launches: equal per-CTA ranges claimed from shared memory (256 x 64x48: 12.4 us vs 16.1 us
        //    grid-wide), except items >= 40 KB (flip decode of 96x72 maps), which keep the all-dynamic dealing.
        // SP_DECODE_GRID_WIDE=1/0 forces either, SP_DECODE_STATIC_PCT sets the static share.
        int static_pct = 0;
        {
            const int force = sp_knob(tune.decode_grid_wide, -1);
            // ... and <= 256 maps per warp slot: the 104 k-person decode of cfg 5 (852 maps per slot) measured 3.09 ms
            // dealt grid-wide against 3.01 ms with contiguous per-CTA ranges, an eighth of it 0.377 against 0.387 ms
            const long long slots = (long long)nwarps * sp_sm_count();
            const bool large = (long long)A.nmaps >= 4LL * slots && (long long)A.nmaps <= 256LL * slots;
            if (force == 0 || (force < 0 && !large && stage_bytes < 40 * 1024)) A.work = nullptr;
            if (large) static_pct = 85;
            static_pct = sp_knob(tune.decode_static_pct, static_pct);
        }
        // ... but never more than 6 dynamically claimed maps per warp: what the dynamic tail evens out is a few maps per
        // warp whatever the launch size, and every claim is an atomic on one address (the 104 k-person decode of cfg 5
        // slowed from 3.01 to 3.3 ms with 15 % = 128 claims per warp)
        int static_maps = stages;
        if (A.work != nullptr && static_pct > 0) {
            const long long share = (long long)A.nmaps / ((long long)grid * nwarps);
            long long dynamic = share * (100 - (static_pct > 100 ? 100 : static_pct)) / 100;
            if (dynamic > 6) dynamic = 6;
            if (share - dynamic > static_maps) static_maps = (int)(share - dynamic);
        }
#define SP_LAUNCH_DECODE(F, KS)                                                                                     \
    do {                                                                                                            \
        SP_CUDA(sp_launch_smem(decode_tma_kernel<F, KS>, dim3(grid), dim3(nwarps * 32), smem, st, A, nwarps, stages, static_maps)); \
    } while (0)
        const bool ks11 = (ksize == 11) && sp_knob(tune.decode_runtime_ksize, 0) == 0;
        if (flip) { if (ks11) SP_LAUNCH_DECODE(true, 11); else SP_LAUNCH_DECODE(true, 0); }
        else      { if (ks11) SP_LAUNCH_DECODE(false, 11); else SP_LAUNCH_DECODE(false, 0); }
#undef SP_LAUNCH_DECODE
        return 0;
    }
    int grid = sp_sm_count() * 8;
    const int need = (A.nmaps + 7) / 8;
    if (grid > need) grid = need;
    if (flip) SP_CUDA(sp_launch(decode_generic_kernel<true>, dim3(grid), dim3(256), 0, st, A));
    else      SP_CUDA(sp_launch(decode_generic_kernel<false>, dim3(grid), dim3(256), 0, st, A));
    return 0;
}

extern "C" int sp_decode_f32(const float* hm, const float* hm_flip, const int* perm,
                             const float* trans_inv, const float* blur_w,
                             float* coords, float* maxval, int* argmax,
                             int B, int K, int H, int W, int ksize, int mode, void* stream) {
    return decode_launch(hm, hm_flip, perm, trans_inv, blur_w, coords, maxval, argmax, nullptr, 0, B, K, H, W, ksize, mode, nullptr, stream);
}

extern "C" int sp_decode_ws_f32(const float* hm, const float* hm_flip, const int* perm,
                                const float* trans_inv, const float* blur_w,
                                float* coords, float* maxval, int* argmax,
                                int B, int K, int H, int W, int ksize, int mode,
                                void* workspace, size_t workspace_bytes, void* stream) {
    SP_RETURN_IF(!workspace, SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF(workspace_bytes < sp_decode_workspace_bytes(), SP_ERR_WORKSPACE);
    SP_RETURN_IF(!sp_aligned16(workspace), SP_ERR_BAD_ALIGNMENT);
    return decode_launch(hm, hm_flip, perm, trans_inv, blur_w, coords, maxval, argmax, nullptr, 0, B, K, H, W, ksize, mode,
                         static_cast<unsigned int*>(workspace), stream);
}

extern "C" int sp_decode_rows_f32(const float* hm, const float* hm_flip, const int* perm,
                                  const float* trans_inv, const float* blur_w,
                                  float* rows, int row_stride, float* coords, float* maxval,
                                  int B, int K, int H, int W, int ksize, int mode,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    SP_RETURN_IF(!workspace || (B > 0 && !rows), SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF(workspace_bytes < sp_decode_workspace_bytes(), SP_ERR_WORKSPACE);
    SP_RETURN_IF(!sp_aligned16(workspace), SP_ERR_BAD_ALIGNMENT);
    return decode_launch(hm, hm_flip, perm, trans_inv, blur_w, coords, maxval, nullptr, rows, row_stride, B, K, H, W, ksize,
                         mode, static_cast<unsigned int*>(workspace), stream);
}
