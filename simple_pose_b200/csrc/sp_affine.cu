// Eval-side caller of the decoder: detection box -> (centre, scale, area, heatmap->image affine).
// Reference: BasicTransform.__call__ without the image warp (datasets/naive_data.py:44-56) =
// box_to_center_scale (commons/joint_utils.py:39-56) + get_affine_transform(c, s, 0, output_shape)
// (:115-152), whose cv.getAffineTransform is OpenCV's 6x6 partial-pivot LU in float64.
//
// 32 bytes in, 52 bytes out per person and ~300 flops: latency-bound, one thread per box. What
// matters is the arithmetic: every rounding of the NumPy/OpenCV path is reproduced (float64 box
// arithmetic, float32 centre/scale, float64->float32 rounding of the second triangle point,
// float32 third point, LU with the same pivot rule and operation order, no FMA contraction), so
// trans_inv is bit-identical to the reference including the ~1e-16 round-off it leaves in the two
// structurally-zero entries.
#include "sp_common.cuh"
#include <float.h>

namespace {

// cv::getAffineTransform(from, to): solve [x y 1 0 0 0; 0 0 0 x y 1] m = (x', y') by in-place LU
__device__ void solve_affine_lu(const float (&from)[3][2], const float (&to)[3][2], double (&m)[6]) {
    double a[6][6], b[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = 0; j < 6; ++j) a[i][j] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double px = (double)from[i][0], py = (double)from[i][1];
        a[2 * i][0] = px; a[2 * i][1] = py; a[2 * i][2] = 1.0;
        a[2 * i + 1][3] = px; a[2 * i + 1][4] = py; a[2 * i + 1][5] = 1.0;
        b[2 * i] = (double)to[i][0];
        b[2 * i + 1] = (double)to[i][1];
    }
    bool singular = false;
    for (int i = 0; i < 6; ++i) {
        int k = i;
        for (int j = i + 1; j < 6; ++j)
            if (fabs(a[j][i]) > fabs(a[k][i])) k = j;
        if (fabs(a[k][i]) < DBL_EPSILON * 100) { singular = true; break; }
        if (k != i) {
            for (int c = 0; c < 6; ++c) { const double t = a[i][c]; a[i][c] = a[k][c]; a[k][c] = t; }
            const double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        const double d = __ddiv_rn(-1.0, a[i][i]);
        for (int j = i + 1; j < 6; ++j) {
            const double alpha = __dmul_rn(a[j][i], d);
            for (int c = i + 1; c < 6; ++c) a[j][c] = __dadd_rn(a[j][c], __dmul_rn(alpha, a[i][c]));
            b[j] = __dadd_rn(b[j], __dmul_rn(alpha, b[i]));
        }
    }
    if (singular) {
        for (int i = 0; i < 6; ++i) m[i] = 0.0;
        return;
    }
    for (int i = 5; i >= 0; --i) {
        double s = b[i];
        for (int k = i + 1; k < 6; ++k) s = __dsub_rn(s, __dmul_rn(a[i][k], m[k]));
        m[i] = __ddiv_rn(s, a[i][i]);
    }
}

// third point of get_3rd_point(a, b) = b + (-(a-b).y, (a-b).x), float32 arithmetic
__device__ __forceinline__ void third_point(float (&tri)[3][2]) {
    const float dx = __fsub_rn(tri[0][0], tri[1][0]), dy = __fsub_rn(tri[0][1], tri[1][1]);
    tri[2][0] = __fadd_rn(tri[1][0], -dy);
    tri[2][1] = __fadd_rn(tri[1][1], dx);
}

// get_affine_transform(center, scale, rot = 0, (dst_w, dst_h)): inv = heatmap -> image, fwd = image -> heatmap
__device__ void affines_from_center_scale(float cx, float cy, float sw, double dst_w, double dst_h,
                                          double (&inv)[6], double* fwd /* 6 or nullptr */) {
    float src[3][2], dst[3][2];
    src[0][0] = cx; src[0][1] = cy;
    src[1][0] = __double2float_rn(__dadd_rn((double)cx, 0.0));
    src[1][1] = __double2float_rn(__dadd_rn((double)cy, (double)__fmul_rn(sw, -0.5f)));
    third_point(src);
    const double hw = __dmul_rn(dst_w, 0.5), hh = __dmul_rn(dst_h, 0.5);
    dst[0][0] = __double2float_rn(hw); dst[0][1] = __double2float_rn(hh);
    dst[1][0] = __double2float_rn(__dadd_rn(hw, 0.0));
    dst[1][1] = __double2float_rn(__dadd_rn(hh, (double)__double2float_rn(__dmul_rn(dst_w, -0.5))));
    third_point(dst);
    if (fwd) {
        double m[6];
        solve_affine_lu(src, dst, m);
#pragma unroll
        for (int e = 0; e < 6; ++e) fwd[e] = m[e];
    }
    solve_affine_lu(dst, src, inv);
}

__device__ __forceinline__ void store_affines(size_t i, const double (&inv)[6], float* trans_inv, double* trans_inv_f64) {
#pragma unroll
    for (int e = 0; e < 6; ++e) {
        if (trans_inv) trans_inv[6 * i + e] = __double2float_rn(inv[e]);
        if (trans_inv_f64) trans_inv_f64[6 * i + e] = inv[e];
    }
}

__global__ void __launch_bounds__(128)
box_affine_kernel(const double* __restrict__ boxes, int xywh, float* __restrict__ center, float* __restrict__ scale,
                  float* __restrict__ area, float* __restrict__ trans_inv, double* __restrict__ trans_inv_f64,
                  double* __restrict__ trans_f64, int P, double ratio, double dst_w, double dst_h, float scale_mult) {
    sp::grid_dep_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const double x1 = boxes[4 * (size_t)i + 0], y1 = boxes[4 * (size_t)i + 1];
    double w = boxes[4 * (size_t)i + 2], h = boxes[4 * (size_t)i + 3];
    if (!xywh) { w = __dsub_rn(w, x1); h = __dsub_rn(h, y1); }
    // box_to_center_scale
    const float cx = __double2float_rn(__dadd_rn(x1, __dmul_rn(w, 0.5)));
    const float cy = __double2float_rn(__dadd_rn(y1, __dmul_rn(h, 0.5)));
    const double rh = __dmul_rn(ratio, h);
    if (w > rh) h = __ddiv_rn(w, ratio);
    else if (w < rh) w = __dmul_rn(h, ratio);
    float sw = __double2float_rn(w), sh = __double2float_rn(h);
    if (cx != -1.0f) { sw = __fmul_rn(sw, scale_mult); sh = __fmul_rn(sh, scale_mult); }
    if (center) { center[2 * (size_t)i] = cx; center[2 * (size_t)i + 1] = cy; }
    if (scale) { scale[2 * (size_t)i] = sw; scale[2 * (size_t)i + 1] = sh; }
    if (area) area[i] = __fmul_rn(sw, sh);
    if (trans_inv || trans_inv_f64 || trans_f64) {
        double inv[6];
        affines_from_center_scale(cx, cy, sw, dst_w, dst_h, inv, trans_f64 ? trans_f64 + 6 * (size_t)i : nullptr);
        store_affines((size_t)i, inv, trans_inv, trans_inv_f64);
    }
}

__global__ void __launch_bounds__(128)
center_scale_affine_kernel(const float* __restrict__ center, const float* __restrict__ scale, float* __restrict__ trans_inv,
                           double* __restrict__ trans_inv_f64, double* __restrict__ trans_f64, int P, double dst_w,
                           double dst_h) {
    sp::grid_dep_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    double inv[6];
    affines_from_center_scale(center[2 * (size_t)i], center[2 * (size_t)i + 1], scale[2 * (size_t)i], dst_w, dst_h, inv,
                              trans_f64 ? trans_f64 + 6 * (size_t)i : nullptr);
    store_affines((size_t)i, inv, trans_inv, trans_inv_f64);
}

}  // namespace

extern "C" int sp_box_affine_f64(const double* boxes, int box_format, float* center, float* scale, float* area,
                                 float* trans_inv, double* trans_inv_f64, double* trans_f64, int P,
                                 double w_h_ratio, int out_w, int out_h, float scale_mult, void* stream) {
    SP_RETURN_IF(P < 0 || out_w <= 0 || out_h <= 0 || !(w_h_ratio > 0.0), SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF(box_format != SP_BOX_XYXY && box_format != SP_BOX_XYWH, SP_ERR_BAD_ARGUMENT);
    if (P == 0) return 0;                  /* empty batches carry null data pointers */
    SP_RETURN_IF(!boxes || (!trans_inv && !trans_inv_f64 && !trans_f64 && !center && !scale && !area), SP_ERR_BAD_ARGUMENT);
    SP_CUDA(sp_launch(box_affine_kernel, dim3((P + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream),
                      boxes, box_format == SP_BOX_XYWH ? 1 : 0, center, scale, area, trans_inv, trans_inv_f64, trans_f64, P,
                      w_h_ratio, (double)out_w, (double)out_h, scale_mult));
    return sp_launch_status();
}

extern "C" int sp_center_scale_affine_f64(const float* center, const float* scale, float* trans_inv,
                                          double* trans_inv_f64, double* trans_f64, int P, int out_w, int out_h,
                                          void* stream) {
    SP_RETURN_IF(P < 0 || out_w <= 0 || out_h <= 0, SP_ERR_BAD_ARGUMENT);
    if (P == 0) return 0;
    SP_RETURN_IF(!center || !scale || (!trans_inv && !trans_inv_f64 && !trans_f64), SP_ERR_BAD_ARGUMENT);
    SP_CUDA(sp_launch(center_scale_affine_kernel, dim3((P + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream),
                      center, scale, trans_inv, trans_inv_f64, trans_f64, P, (double)out_w, (double)out_h));
    return sp_launch_status();
}
