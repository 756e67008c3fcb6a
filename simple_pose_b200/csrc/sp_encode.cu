// A1: DarkPose-style unbiased Gaussian target encoder (reference commons/transforms.py:167-191).
//
// Store-bound: K*H*W*4 bytes written per person, 12 bytes read per joint. One warp owns one
// (person, joint) map: the lanes evaluate the W + H separable factors exp(-dx^2/(2 s^2)) and
// exp(-dy^2/(2 s^2)) in float64 (the reference computes in float64 and rounds once to float32,
// so the product is formed in float64 too and converted with a single cvt.rn.f32.f64), park
// them in a private slice of shared memory, then stream the map out with coalesced 16-byte
// stores (32 lanes x float4 = 512 contiguous bytes per instruction). No block-level barrier.
// Denormal results are kept (no -ftz / fast-math), as in the reference.
#include "sp_common.cuh"
#include "sp_gauss.cuh"

namespace {

using namespace sp_gauss;
constexpr int kWarpsPerCta = 8;

template <bool VEC4, int WARPS>
__global__ void __launch_bounds__(WARPS* SP_WARP)
encode_refine_kernel(const float* __restrict__ joints, float* __restrict__ targets, float* __restrict__ weights,
                     int nmaps, int H, int W, float reach, double denom, int parts) {
    extern __shared__ __align__(16) double factors[];   // per warp: ex[Wpad] then ey[H]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpad = (W + 1) & ~1;                        // keeps ey 16-byte aligned
    double* ex = factors + (size_t)warp * (wpad + ((H + 1) & ~1));   // even stride: every warp's ex stays 16-byte aligned (odd H)
    double* ey = ex + wpad;
    const int hw = H * W;
    const long long total_warps = (long long)gridDim.x * WARPS;
    const long long units = (long long)nmaps * parts;     // a unit = rows [r0, r1) of one map, one warp
    const int rows_per_part = (H + parts - 1) / parts;
    sp::grid_dep_wait();

    for (long long u = (long long)blockIdx.x * WARPS + warp; u < units; u += total_warps) {
        const int m = (int)(u / parts);
        const int part = (int)(u - (long long)m * parts);
        const int r0 = part * rows_per_part, r1 = min(H, r0 + rows_per_part);
        const float mx = __ldg(joints + 3 * (size_t)m + 0);
        const float my = __ldg(joints + 3 * (size_t)m + 1);
        const float vis = __ldg(joints + 3 * (size_t)m + 2);
        const JointVerdict jv = judge_joint(mx, my, vis, reach, H, W);
        if (lane == 0 && part == 0) weights[m] = jv.weight;
        float* out = targets + (size_t)m * hw;

        if (!jv.draw) {
            if (VEC4) {
                float4* o4 = reinterpret_cast<float4*>(out);
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int q = (r0 * W >> 2) + lane; q < (r1 * W >> 2); q += 32) o4[q] = z;
            } else {
                for (int i = r0 * W + lane; i < r1 * W; i += 32) out[i] = 0.f;
            }
            continue;
        }

        __syncwarp();   // previous unit's readers are done with ex/ey
        for (int i = lane; i < W + (r1 - r0); i += 32) {
            if (i < W) ex[i] = gauss_factor(i, mx, denom);
            else       ey[r0 + i - W] = gauss_factor(r0 + i - W, my, denom);
        }
        __syncwarp();

        if (VEC4) {
            // W % 4 == 0: a quad never straddles rows. Track (row, quad-in-row) incrementally.
            const int qpr = W >> 2;
            const int q0 = r0 * qpr + lane, nq = r1 * qpr;
            int y = q0 / qpr;
            int xq = q0 - y * qpr;
            const int step_y = 32 / qpr, step_x = 32 - step_y * qpr;
            float4* o4 = reinterpret_cast<float4*>(out);
            for (int q = q0; q < nq; q += 32) {
                const double2 a = *reinterpret_cast<const double2*>(ex + 4 * xq);
                const double2 b = *reinterpret_cast<const double2*>(ex + 4 * xq + 2);
                const double fy = ey[y];
                float4 v;
                v.x = __double2float_rn(__dmul_rn(a.x, fy));
                v.y = __double2float_rn(__dmul_rn(a.y, fy));
                v.z = __double2float_rn(__dmul_rn(b.x, fy));
                v.w = __double2float_rn(__dmul_rn(b.y, fy));
                o4[q] = v;
                xq += step_x;
                y += step_y;
                if (xq >= qpr) { xq -= qpr; ++y; }
            }
        } else {
            for (int i = r0 * W + lane; i < r1 * W; i += 32) {
                const int y = i / W, x = i - y * W;
                out[i] = __double2float_rn(__dmul_rn(ex[x], ey[y]));
            }
        }
    }
}


// A1': BasicSimpleTransform.get_heat_map (commons/transforms.py:80-116): the centre is quantised
// to int(x/stride + 0.5) and only a (6 sigma + 1)^2 patch of a fixed Gaussian table is pasted.
// The table is computed on the host exactly as the reference does (NumPy float32) and handed in,
// so the pasted values are bit-identical by construction. One warp per map, float4 stores.
template <bool VEC4>
__global__ void __launch_bounds__(kWarpsPerCta* SP_WARP)
encode_basic_kernel(const float* __restrict__ joints, const float* __restrict__ table, float* __restrict__ targets,
                    float* __restrict__ weights, int nmaps, int H, int W, double reach, float stride, int side) {
    extern __shared__ float tab[];
    sp::grid_dep_wait();
    for (int i = threadIdx.x; i < side * side; i += blockDim.x) tab[i] = __ldg(table + i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int m = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (m >= nmaps) return;
    const float jx = __ldg(joints + 3 * (size_t)m + 0);
    const float jy = __ldg(joints + 3 * (size_t)m + 1);
    const float vis = __ldg(joints + 3 * (size_t)m + 2);
    const int mu_x = (int)__fadd_rn(__fdiv_rn(jx, stride), 0.5f);      // int(x / stride + 0.5), float32
    const int mu_y = (int)__fadd_rn(__fdiv_rn(jy, stride), 0.5f);
    const int lo_x = (int)((double)mu_x - reach), lo_y = (int)((double)mu_y - reach);     // Python float maths
    const int hi_x = (int)((double)mu_x + reach + 1.0), hi_y = (int)((double)mu_y + reach + 1.0);
    const bool culled = lo_x >= W || lo_y >= H || hi_x < 0 || hi_y < 0;
    if (lane == 0) weights[m] = culled ? 0.f : vis;
    const bool draw = !culled && vis > 0.5f;
    // pasted window, clipped to the map and to the table
    const int x0 = max(0, lo_x), x1 = min(min(hi_x, W), lo_x + side);
    const int y0 = max(0, lo_y), y1 = min(min(hi_y, H), lo_y + side);
    float* out = targets + (size_t)m * H * W;
    if (VEC4) {
        // W % 4 == 0: whole rows of 16-byte stores; (row, quad-in-row) advance incrementally, no division
        const int qpr = W >> 2, nq = H * qpr;
        const int step_y = 32 / qpr, step_x = 32 - step_y * qpr;
        int y = lane / qpr;
        int xq = lane - y * qpr;
        float4* o4 = reinterpret_cast<float4*>(out);
        for (int q = lane; q < nq; q += 32) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int x = 4 * xq;
            if (draw && y >= y0 && y < y1 && x + 3 >= x0 && x < x1) {
                const float* row = tab + (y - lo_y) * side - lo_x;
                if (x >= x0 && x < x1) v.x = row[x];
                if (x + 1 >= x0 && x + 1 < x1) v.y = row[x + 1];
                if (x + 2 >= x0 && x + 2 < x1) v.z = row[x + 2];
                if (x + 3 >= x0 && x + 3 < x1) v.w = row[x + 3];
            }
            o4[q] = v;
            xq += step_x;
            y += step_y;
            if (xq >= qpr) { xq -= qpr; ++y; }
        }
    } else {
        for (int i = lane; i < H * W; i += 32) {
            const int y = i / W, x = i - y * W;
            float v = 0.f;
            if (draw && x >= x0 && x < x1 && y >= y0 && y < y1) v = tab[(y - lo_y) * side + (x - lo_x)];
            out[i] = v;
        }
    }
}

}  // namespace

extern "C" int sp_encode_f32(const float* joints, float* targets, float* weights,
                             int B, int K, int H, int W, double sigma, void* stream) {
    SP_RETURN_IF(B < 0 || K <= 0 || H <= 0 || W <= 0 || !(sigma > 0.0), SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF(B > 0 && (!joints || !targets || !weights), SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF((long long)B * K > 0x7fffffffLL || (long long)H * W > (1 << 24), SP_ERR_UNSUPPORTED);
    if (B == 0) return 0;
    const int nmaps = B * K;
    const double reach_d = sigma * 3.0;            // tmp_size = sigma * 3 (Python float)
    const float reach = (float)reach_d;            // weak Python scalar -> float32 in the cull test
    const double denom = 2.0 * (sigma * sigma);    // 2 * sigma ** 2
    const int wpad = (W + 1) & ~1;
    // one map per warp, no persistence: the hardware backfills CTAs as they retire. Small CTAs keep
    // the per-SM load even: 2-warp CTAs put 29.4 +- 0.5 CTAs on an SM where 8-warp CTAs put 7 or 8
    // (96x72, 512 persons: 9 % of the launch was the SMs that drew 8). Measured 8 -> 2 warps per CTA:
    // 0.934 -> 0.954 of the HBM peak at 64x48 x 1024, 0.862 -> 0.900 at 96x72 x 512, 10.6 -> 7.9 us at 64x48 x 128.
    const SpTuning& tune = sp_tuning();
    int warps = sp_knob(tune.encode_warps, 2);
    if (warps != 2 && warps != 4 && warps != 8) warps = 2;
    const size_t smem = (size_t)warps * (wpad + ((H + 1) & ~1)) * sizeof(double);
    SP_RETURN_IF(smem > 200 * 1024, SP_ERR_UNSUPPORTED);
    const bool vec4 = (W % 4 == 0) && sp_aligned16(targets);
    // A unit of work is a range of ~32 rows of one map, one warp each (parts = H / 32 per map). Whole
    // maps per warp leave a launch of 512 persons at 96x72 with 29 CTAs per SM, all resident at once:
    // ONE wave in which every warp first evaluates its float64 factors and then stores, the two
    // phases never overlapping across CTAs. Shorter units give several waves and an even finish:
    // 96x72 x 512: 41.3 -> 39.0 us (0.89 -> 0.945 of the HBM peak), x 1024: 0.96 -> 1.00;
    // 64x48 x 1024: 35.0 -> 33.7 us (0.94 -> 0.97). 16-row units lose again (the W factors are
    // re-evaluated per unit: 64x48 x 1024 40.9 us).
    int parts = H / 32;
    if (parts < 1) parts = 1;
    parts = sp_knob(tune.encode_parts, parts);
    if (parts < 1) parts = 1;
    if (parts > H) parts = H;
    const long long units = (long long)nmaps * parts;
    SP_RETURN_IF((units + warps - 1) / warps > 0x7fffffffLL, SP_ERR_UNSUPPORTED);
    const int grid = (int)((units + warps - 1) / warps);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define SP_LAUNCH_ENCODE(V, NW)                                                                                         \
    do {                                                                                                                \
        SP_CUDA(sp_launch_smem(encode_refine_kernel<V, NW>, dim3(grid), dim3(NW * SP_WARP), smem, st, joints, targets, weights, nmaps, H, W, reach, denom, parts)); \
    } while (0)
    if (vec4) { if (warps == 2) SP_LAUNCH_ENCODE(true, 2); else if (warps == 4) SP_LAUNCH_ENCODE(true, 4); else SP_LAUNCH_ENCODE(true, 8); }
    else      { if (warps == 2) SP_LAUNCH_ENCODE(false, 2); else if (warps == 4) SP_LAUNCH_ENCODE(false, 4); else SP_LAUNCH_ENCODE(false, 8); }
#undef SP_LAUNCH_ENCODE
    return 0;
}

extern "C" int sp_encode_basic_f32(const float* joints, const float* table, float* targets, float* weights,
                                   int B, int K, int H, int W, double sigma, int stride, int side, void* stream) {
    SP_RETURN_IF(B < 0 || K <= 0 || H <= 0 || W <= 0 || !(sigma > 0.0) || stride <= 0 || side <= 0, SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF(B > 0 && (!joints || !table || !targets || !weights), SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF((long long)B * K > 0x7fffffffLL || (long long)H * W > (1 << 24) || side > 96, SP_ERR_UNSUPPORTED);
    if (B == 0) return 0;
    const int nmaps = B * K;
    const int grid = (nmaps + kWarpsPerCta - 1) / kWarpsPerCta;
    const size_t smem = (size_t)side * side * sizeof(float);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (W % 4 == 0 && sp_aligned16(targets))
        SP_CUDA(sp_launch(encode_basic_kernel<true>, dim3(grid), dim3(kWarpsPerCta * SP_WARP), smem, st, joints, table, targets, weights,
                          nmaps, H, W, sigma * 3.0, (float)stride, side));
    else
        SP_CUDA(sp_launch(encode_basic_kernel<false>, dim3(grid), dim3(kWarpsPerCta * SP_WARP), smem, st, joints, table, targets, weights,
                          nmaps, H, W, sigma * 3.0, (float)stride, side));
    return 0;
}
