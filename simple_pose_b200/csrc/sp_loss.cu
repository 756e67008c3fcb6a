// A2: masked joints-MSE, forward and backward in one pass
// (reference processors/dp_pose_hrnet_solver.py:86,106-107:
//  loss = 0.5 * MSELoss(pred * mask[..., None, None], target * mask[..., None, None]); loss.backward()).
//
// HBM-bound: reads pred and target once, writes grad once (3 streams; the ATen sequence makes
// ~6 passes and two temporaries). Each CTA walks whole (person, joint) maps so the mask is a
// per-map scalar; 16-byte loads, several in flight per thread. The sum of squares is
// accumulated per thread in float32 over a few elements, then in float64 across the block;
// block partials go to the caller's workspace and the last block to finish adds them in a
// fixed order (deterministic, no floating-point atomics) and writes the scalar loss.
//
// Arithmetic follows ATen so that grad is bit-identical for finite inputs:
//   d = fl(m*p) - fl(m*t); grad = ((fl(2/N) * d) * 0.5) * m   (mse_loss_backward, then mul backward)
#include "sp_common.cuh"
#include "sp_reduce.cuh"

namespace {

using namespace sp_reduce;
constexpr int kThreads = 256;

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ float sq_err_and_grad(float p, float t, float m, float norm, float half_scale, float& g) {
    const float d = __fsub_rn(__fmul_rn(m, p), __fmul_rn(m, t));
    g = __fmul_rn(__fmul_rn(__fmul_rn(norm, d), half_scale), m);
    return d;
}

template <bool VEC4, bool WRITE_GRAD>
__global__ void __launch_bounds__(kThreads)
mse_fwd_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                   const float* __restrict__ mask, float* __restrict__ grad, float* __restrict__ loss,
                   MseWorkspace* __restrict__ ws, int nmaps, int hw, float norm, float half_scale,
                   double inv_count, int skip_masked) {
    double block_sum = 0.0;
    sp::grid_dep_wait();
    sp::grid_dep_launch();

    for (int m = blockIdx.x; m < nmaps; m += gridDim.x) {
        const float mk = __ldg(mask + m);
        const size_t base = (size_t)m * hw;
        float acc = 0.f;
        if (skip_masked && mk == 0.f) {
            if (WRITE_GRAD) {
                if (VEC4) {
                    float4* g4 = reinterpret_cast<float4*>(grad + base);
                    for (int q = threadIdx.x; q < (hw >> 2); q += kThreads) g4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
                    for (int i = threadIdx.x; i < hw; i += kThreads) grad[base + i] = 0.f;
                }
            }
            continue;
        }
        if (VEC4) {
            const float4* p4 = reinterpret_cast<const float4*>(pred + base);
            const float4* t4 = reinterpret_cast<const float4*>(target + base);
            float4* g4 = reinterpret_cast<float4*>(grad + base);
            const int nq = hw >> 2;
#pragma unroll 4
            for (int q = threadIdx.x; q < nq; q += kThreads) {
                const float4 p = ldg_stream(p4 + q);
                const float4 t = ldg_stream(t4 + q);
                float4 g;
                float d;
                d = sq_err_and_grad(p.x, t.x, mk, norm, half_scale, g.x); acc = fmaf(d, d, acc);
                d = sq_err_and_grad(p.y, t.y, mk, norm, half_scale, g.y); acc = fmaf(d, d, acc);
                d = sq_err_and_grad(p.z, t.z, mk, norm, half_scale, g.z); acc = fmaf(d, d, acc);
                d = sq_err_and_grad(p.w, t.w, mk, norm, half_scale, g.w); acc = fmaf(d, d, acc);
                if (WRITE_GRAD) g4[q] = g;
            }
        } else {
            for (int i = threadIdx.x; i < hw; i += kThreads) {
                float g;
                const float d = sq_err_and_grad(pred[base + i], target[base + i], mk, norm, half_scale, g);
                acc = fmaf(d, d, acc);
                if (WRITE_GRAD) grad[base + i] = g;
            }
        }
        block_sum += (double)acc;
    }

    finish_loss<kThreads>(block_sum, ws, loss, inv_count);
}

__global__ void __launch_bounds__(256)
scale_inplace_kernel(float* __restrict__ data, long long n, const float* __restrict__ scale_dev) {
    sp::grid_dep_wait();
    sp::grid_dep_launch();
    const float s = __ldg(scale_dev);
    if (s == 1.0f) return;
    const long long n4 = n >> 2;
    float4* d4 = reinterpret_cast<float4*>(data);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
        float4 v = d4[q];
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        d4[q] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) data[(n4 << 2) + threadIdx.x] *= s;
}

}  // namespace

extern "C" size_t sp_mse_workspace_bytes(void) { return sizeof(MseWorkspace); }

extern "C" int sp_mse_fwd_bwd_f32(const float* pred, const float* target, const float* mask,
                                  float* grad, float* loss, void* workspace, size_t workspace_bytes,
                                  int B, int K, int HW, float grad_scale, int flags, void* stream) {
    SP_RETURN_IF(!pred || !target || !mask || !loss || !workspace, SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF(B <= 0 || K <= 0 || HW <= 0, SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF((long long)B * K > 0x7fffffffLL, SP_ERR_UNSUPPORTED);
    SP_RETURN_IF(workspace_bytes < sizeof(MseWorkspace), SP_ERR_WORKSPACE);
    SP_RETURN_IF(!sp_aligned16(workspace), SP_ERR_BAD_ALIGNMENT);
    const int nmaps = B * K;
    const double count = (double)B * (double)K * (double)HW;
    const float norm = (float)(2.0 / count);       // ATen: norm = 2 / numel, applied in float32
    const float half_scale = 0.5f * grad_scale;    // the 0.5 of the loss expression times upstream grad
    const bool vec4 = (HW % 4 == 0) && sp_aligned16(pred) && sp_aligned16(target) && (!grad || sp_aligned16(grad));
    int grid = sp_sm_count() * 8;
    if (grid > nmaps) grid = nmaps;
    if (grid > kMaxPartials) grid = kMaxPartials;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MseWorkspace* ws = static_cast<MseWorkspace*>(workspace);
    const int skip = (flags & SP_MSE_SKIP_MASKED) ? 1 : 0;
#define SP_LAUNCH_MSE(V, G) \
    SP_CUDA(sp_launch(mse_fwd_bwd_kernel<V, G>, dim3(grid), dim3(kThreads), 0, st, pred, target, mask, grad, loss, ws, nmaps, HW, norm, half_scale, 1.0 / count, skip))
    if (vec4) { if (grad) SP_LAUNCH_MSE(true, true); else SP_LAUNCH_MSE(true, false); }
    else      { if (grad) SP_LAUNCH_MSE(false, true); else SP_LAUNCH_MSE(false, false); }
#undef SP_LAUNCH_MSE
    return sp_launch_status();
}

extern "C" int sp_scale_inplace_f32(float* data, long long n, const float* scale_dev, void* stream) {
    SP_RETURN_IF(!data || !scale_dev || n < 0, SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF(!sp_aligned16(data), SP_ERR_BAD_ALIGNMENT);
    if (n == 0) return 0;
    long long blocks = ((n >> 2) + 255) / 256;
    const long long cap = (long long)sp_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    SP_CUDA(sp_launch(scale_inplace_kernel, dim3((int)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), data, n, scale_dev));
    return sp_launch_status();
}
