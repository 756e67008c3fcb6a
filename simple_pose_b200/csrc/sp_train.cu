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
#include <math_constants.h>

namespace {

using namespace sp_gauss;
using namespace sp_reduce;

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * SP_WARP;

__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ float warp_max_f32(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}

// running (value, quad) argmax over the quads a lane visits in increasing order
struct QuadBest {
    float best;
    int bq;
    float poison;
    __device__ __forceinline__ void init() { best = -CUDART_INF_F; bq = 0x3fffffff; poison = 0.f; }
    __device__ __forceinline__ void push(float a, float b, float c, float d, int q) {
        const float m4 = sp::fmax_nan(sp::fmax_nan(a, b), sp::fmax_nan(c, d));
        poison = fmaf(m4, 0.f, poison);
        if (m4 > best) { best = m4; bq = q; }
    }
};

struct MaskedPredView {          // m * pred[i], straight from global memory (exact fallback only)
    const float* p;
    float m;
    __device__ __forceinline__ float at(int i) const { return __fmul_rn(m, p[i]); }
};

template <typename View>
__device__ __noinline__ void argmax_exact_scan(const View map, int hw, int lane, float& val, int& idx) {
    float bv = -CUDART_INF_F;
    int bi = 0x7fffffff;
    for (int i = lane; i < hw; i += 32) {
        const float v = map.at(i);
        if (sp::better(v, i, bv, bi)) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(SP_FULL, bv, o);
        const int oi = __shfl_xor_sync(SP_FULL, bi, o);
        if (sp::better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    val = bv;
    idx = bi;
}

// heat_map_to_axis on an index/value pair: (x, y) as floats, zeroed when the max is not > 0
__device__ __forceinline__ float2 axis_of(float val, int idx, int W) {
    if (!(val > 0.f)) return make_float2(0.f, 0.f);
    const int y = idx / W;
    return make_float2((float)(idx - y * W), (float)y);
}

template <bool WRITE_GRAD, bool WRITE_TARGETS, bool ACC>
__global__ void __launch_bounds__(kThreads)
encode_mse_kernel(const float* __restrict__ joints, const float* __restrict__ pred, float* __restrict__ grad,
                  float* __restrict__ targets, float* __restrict__ weights, float* __restrict__ loss,
                  MseWorkspace* __restrict__ ws, float2* __restrict__ pred_xy, float2* __restrict__ label_xy,
                  int nmaps, int H, int W, float reach, double denom, float norm, float half_scale, double inv_count) {
    extern __shared__ __align__(16) double factors[];   // per warp: ex[Wpad] then ey[H]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpad = (W + 1) & ~1;
    double* ex = factors + (size_t)warp * (wpad + H);
    double* ey = ex + wpad;
    const int hw = H * W;
    const int nq = hw >> 2;
    const int qpr = W >> 2;
    const int step_y = 32 / qpr, step_x = 32 - step_y * qpr;
    const int total_warps = gridDim.x * kWarps;
    double warp_sum_sq = 0.0;

    for (int m = blockIdx.x * kWarps + warp; m < nmaps; m += total_warps) {
        const float mx = __ldg(joints + 3 * (size_t)m + 0);
        const float my = __ldg(joints + 3 * (size_t)m + 1);
        const float vis = __ldg(joints + 3 * (size_t)m + 2);
        const JointVerdict jv = judge_joint(mx, my, vis, reach, H, W);
        const float mk = jv.weight;
        if (lane == 0 && weights) weights[m] = mk;
        __syncwarp();
        if (jv.draw) {
            for (int i = lane; i < W + H; i += 32) {
                if (i < W) ex[i] = gauss_factor(i, mx, denom);
                else       ey[i - W] = gauss_factor(i - W, my, denom);
            }
        }
        __syncwarp();

        const float4* p4 = reinterpret_cast<const float4*>(pred + (size_t)m * hw);
        float4* g4 = reinterpret_cast<float4*>(grad + (size_t)m * hw);
        float4* t4 = reinterpret_cast<float4*>(targets + (size_t)m * hw);
        int y = lane / qpr;
        int xq = lane - y * qpr;
        float acc = 0.f;
        QuadBest bp, bt;
        bp.init();
        bt.init();
        const bool track = ACC && (mk != 0.f);
#pragma unroll 4
        for (int q = lane; q < nq; q += 32) {
            const float4 p = ldg_stream4(p4 + q);
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (jv.draw) {
                const double2 a = *reinterpret_cast<const double2*>(ex + 4 * xq);
                const double2 b = *reinterpret_cast<const double2*>(ex + 4 * xq + 2);
                const double fy = ey[y];
                t.x = __double2float_rn(__dmul_rn(a.x, fy));
                t.y = __double2float_rn(__dmul_rn(a.y, fy));
                t.z = __double2float_rn(__dmul_rn(b.x, fy));
                t.w = __double2float_rn(__dmul_rn(b.y, fy));
            }
            const float px = __fmul_rn(mk, p.x), py = __fmul_rn(mk, p.y), pz = __fmul_rn(mk, p.z), pw = __fmul_rn(mk, p.w);
            const float tx = __fmul_rn(mk, t.x), ty = __fmul_rn(mk, t.y), tz = __fmul_rn(mk, t.z), tw = __fmul_rn(mk, t.w);
            const float dx = __fsub_rn(px, tx), dy = __fsub_rn(py, ty), dz = __fsub_rn(pz, tz), dw = __fsub_rn(pw, tw);
            acc = fmaf(dx, dx, acc);
            acc = fmaf(dy, dy, acc);
            acc = fmaf(dz, dz, acc);
            acc = fmaf(dw, dw, acc);
            if (WRITE_GRAD) {
                float4 g;
                g.x = __fmul_rn(__fmul_rn(__fmul_rn(norm, dx), half_scale), mk);
                g.y = __fmul_rn(__fmul_rn(__fmul_rn(norm, dy), half_scale), mk);
                g.z = __fmul_rn(__fmul_rn(__fmul_rn(norm, dz), half_scale), mk);
                g.w = __fmul_rn(__fmul_rn(__fmul_rn(norm, dw), half_scale), mk);
                g4[q] = g;
            }
            if (WRITE_TARGETS) t4[q] = t;
            if (track) {
                bp.push(px, py, pz, pw, q);
                bt.push(tx, ty, tz, tw, q);
            }
            xq += step_x;
            y += step_y;
            if (xq >= qpr) { xq -= qpr; ++y; }
        }
        warp_sum_sq += (double)acc;

        if (ACC) {
            float2 pxy = make_float2(0.f, 0.f), lxy = make_float2(0.f, 0.f);
            if (track) {
                // predicted map: fl(m * p)
                float pv;
                int pi;
                if (__any_sync(SP_FULL, bp.poison != bp.poison)) {
                    MaskedPredView view{pred + (size_t)m * hw, mk};
                    argmax_exact_scan(view, hw, lane, pv, pi);
                } else {
                    const float gmax = warp_max_f32(bp.best);
                    const unsigned gq = __reduce_min_sync(SP_FULL, (bp.best == gmax) ? (unsigned)bp.bq : 0x7fffffffu);
                    const float4 w = __ldg(p4 + gq);
                    const float a = __fmul_rn(mk, w.x), b = __fmul_rn(mk, w.y), c = __fmul_rn(mk, w.z);
                    const int sub = (a == gmax) ? 0 : (b == gmax) ? 1 : (c == gmax) ? 2 : 3;
                    pv = gmax;
                    pi = 4 * (int)gq + sub;
                }
                pxy = axis_of(pv, pi, W);
                // target map: fl(m * t); always finite
                if (jv.draw) {
                    const float gmax = warp_max_f32(bt.best);
                    const unsigned gq = __reduce_min_sync(SP_FULL, (bt.best == gmax) ? (unsigned)bt.bq : 0x7fffffffu);
                    const int gy = (int)gq / qpr, gx = 4 * ((int)gq - gy * qpr);
                    const double fy = ey[gy];
                    const float a = __fmul_rn(mk, __double2float_rn(__dmul_rn(ex[gx + 0], fy)));
                    const float b = __fmul_rn(mk, __double2float_rn(__dmul_rn(ex[gx + 1], fy)));
                    const float c = __fmul_rn(mk, __double2float_rn(__dmul_rn(ex[gx + 2], fy)));
                    const int sub = (a == gmax) ? 0 : (b == gmax) ? 1 : (c == gmax) ? 2 : 3;
                    lxy = axis_of(gmax, 4 * (int)gq + sub, W);
                }
            }
            if (lane == 0) {
                pred_xy[m] = pxy;
                label_xy[m] = lxy;
            }
        }
    }
    // every lane carries the squared error of the quads it visited
    finish_loss<kThreads>(warp_sum_sq, ws, loss, inv_count);
}

// HeatMapAcc epilogue (metrics/pose_metrics.py:227-245) on the [B,K] argmax coordinates.
__global__ void __launch_bounds__(256)
heatmap_acc_kernel(const float2* __restrict__ pred_xy, const float2* __restrict__ label_xy, float* __restrict__ acc,
                   int B, int K, float norm_x, float norm_y, float thresh) {
    extern __shared__ int counters[];        // hit[K], valid[K]
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

extern "C" int sp_encode_mse_fwd_bwd_f32(const float* joints, const float* pred, float* grad, float* targets,
                                         float* weights, float* loss, float* pred_xy, float* label_xy,
                                         void* workspace, size_t workspace_bytes,
                                         int B, int K, int H, int W, double sigma, float grad_scale, void* stream) {
    SP_RETURN_IF(!joints || !pred || !loss || !workspace, SP_ERR_BAD_ARGUMENT);
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
    const size_t smem = (size_t)kWarps * (wpad + H) * sizeof(double);
    SP_RETURN_IF(smem > 200 * 1024, SP_ERR_UNSUPPORTED);
    int grid = (nmaps + kWarps - 1) / kWarps;
    if (grid > kMaxPartials) grid = kMaxPartials;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MseWorkspace* ws = static_cast<MseWorkspace*>(workspace);
#define SP_LAUNCH_TRAIN(G, T, A)                                                                                         \
    do {                                                                                                                 \
        if (smem > 48 * 1024)                                                                                            \
            SP_CUDA(cudaFuncSetAttribute(encode_mse_kernel<G, T, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        encode_mse_kernel<G, T, A><<<grid, kThreads, smem, st>>>(joints, pred, grad, targets, weights, loss, ws,         \
            reinterpret_cast<float2*>(pred_xy), reinterpret_cast<float2*>(label_xy), nmaps, H, W, reach, denom, norm,    \
            half_scale, 1.0 / count);                                                                                    \
    } while (0)
    const int sel = (grad ? 4 : 0) | (targets ? 2 : 0) | (pred_xy ? 1 : 0);
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
    return sp_launch_status();
}

extern "C" int sp_heatmap_acc_f32(const float* pred_xy, const float* label_xy, float* acc,
                                  int B, int K, int H, int W, float distance_thresh, float norm_frac, void* stream) {
    SP_RETURN_IF(!pred_xy || !label_xy || !acc, SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF(B <= 0 || K <= 0 || H <= 0 || W <= 0 || K > 4096, SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF((long long)B * K > 0x7fffffffLL, SP_ERR_UNSUPPORTED);
    SP_RETURN_IF(!sp_aligned16(pred_xy) || !sp_aligned16(label_xy), SP_ERR_BAD_ALIGNMENT);
    // norm = tensor([W, H], float32) / norm_frac
    const float nx = (float)W / norm_frac, ny = (float)H / norm_frac;
    heatmap_acc_kernel<<<1, 256, (size_t)2 * K * sizeof(int), static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2*>(pred_xy), reinterpret_cast<const float2*>(label_xy), acc, B, K, nx, ny,
        distance_thresh);
    return sp_launch_status();
}
