// Callers either side of the path. Eval side: detection box -> (centre, scale, area, heatmap->image
// affine). Train side (bottom of the file): box + augmentation draws + image-pixel joints ->
// heatmap-pixel joints for the encoder, trans_inv, input-pixel joints.
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

// get_affine_transform(center, scale, rot, (dst_w, dst_h)): inv = heatmap -> image, fwd = image -> heatmap.
// rot in degrees (float64): rot_rad = pi*rot/180, get_dir = (0*cs - p*sn, 0*sn + p*cs) in float64 with
// p = float32(src_w * -0.5) promoted (commons/joint_utils.py:78-85,137-138). rot == 0 gives sn = 0, cs = 1
// exactly, i.e. the eval-side transform. sincos() is CUDA's 1-ulp double routine: against NumPy's the
// direction can differ in the last float64 bit, which survives the float32 rounding of the triangle
// points with probability ~1e-8 (train-side caller).
__device__ void affines_from_center_scale(float cx, float cy, float sw, double rot_deg, double dst_w, double dst_h,
                                          double (&inv)[6], double* fwd /* 6 or nullptr */) {
    float src[3][2], dst[3][2];
    const double p = (double)__fmul_rn(sw, -0.5f);
    double sn = 0.0, cs = 1.0;
    if (rot_deg != 0.0) sincos(__ddiv_rn(__dmul_rn(3.141592653589793, rot_deg), 180.0), &sn, &cs);
    const double dir_x = __dsub_rn(__dmul_rn(0.0, cs), __dmul_rn(p, sn));
    const double dir_y = __dadd_rn(__dmul_rn(0.0, sn), __dmul_rn(p, cs));
    src[0][0] = cx; src[0][1] = cy;
    src[1][0] = __double2float_rn(__dadd_rn((double)cx, dir_x));
    src[1][1] = __double2float_rn(__dadd_rn((double)cy, dir_y));
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
        affines_from_center_scale(cx, cy, sw, 0.0, dst_w, dst_h, inv, trans_f64 ? trans_f64 + 6 * (size_t)i : nullptr);
        store_affines((size_t)i, inv, trans_inv, trans_inv_f64);
    }
}

__global__ void __launch_bounds__(128)
center_scale_affine_kernel(const float* __restrict__ center, const float* __restrict__ scale,
                           const double* __restrict__ rot_deg, float* __restrict__ trans_inv,
                           double* __restrict__ trans_inv_f64, double* __restrict__ trans_f64, int P, double dst_w,
                           double dst_h) {
    sp::grid_dep_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    double inv[6];
    affines_from_center_scale(center[2 * (size_t)i], center[2 * (size_t)i + 1], scale[2 * (size_t)i],
                              rot_deg ? rot_deg[i] : 0.0, dst_w, dst_h, inv, trans_f64 ? trans_f64 + 6 * (size_t)i : nullptr);
    store_affines((size_t)i, inv, trans_inv, trans_inv_f64);
}


// ---- train-side caller of the encoder --------------------------------------------------------
// RefineSimpleTransform.__call__ (commons/transforms.py:193-223) without the image work; the random
// draws (scale_ratio, rot, flip) are inputs. Phase 1: one thread per person derives centre/scale
// (box_to_center_scale, :200-201; scale * scale_ratio in float32, :203; centre mirror :209) and the
// affines of get_affine_transform for the heatmap size (:213) and, when asked, the input size (:212)
// into shared memory. Phase 2: the CTA walks its persons' joints coalesced: flip_joints
// (commons/joint_utils.py:102-112: x -> width - x - 1 in float32 for every row, rows permuted) and
// affine_transform_batch (:88-99: rows with vis > 0, float64 dot in dgemm order
// fma(1, t2, fma(y, t1, x*t0)), rounded to float32).
constexpr int kGeomPersons = 128;

__device__ __forceinline__ float affine_row(double x, double y, const double* t) {
    return __double2float_rn(__dadd_rn(__fma_rn(y, t[1], __dmul_rn(x, t[0])), t[2]));
}

// joints_in row (after the optional flip) of joint k of a person
__device__ __forceinline__ void load_joint(const float* __restrict__ person_joints, int k, bool flipped,
                                           const int* __restrict__ perm, float width, float& x, float& y, float& v) {
    const int src = flipped ? perm[k] : k;
    x = person_joints[3 * src + 0];
    y = person_joints[3 * src + 1];
    v = person_joints[3 * src + 2];
    if (flipped) x = __fsub_rn(__fsub_rn(width, x), 1.0f);
}

__global__ void __launch_bounds__(kGeomPersons)
train_geometry_kernel(const double* __restrict__ boxes, const int* __restrict__ img_w, const float* __restrict__ joints,
                      const double* __restrict__ scale_ratio, const double* __restrict__ rot_deg,
                      const unsigned char* __restrict__ flip, const int* __restrict__ perm,
                      float* __restrict__ joints_hm, float* __restrict__ joints_input, float* __restrict__ trans_inv,
                      double* __restrict__ trans_inv_f64, double* __restrict__ img_trans_f64,
                      float* __restrict__ center, float* __restrict__ scale, int P, int K, double ratio, double in_w,
                      double in_h, double out_w, double out_h, float scale_mult) {
    __shared__ double s_hm[kGeomPersons][6];     // image -> heatmap
    __shared__ double s_in[kGeomPersons][6];     // image -> network input
    sp::grid_dep_wait();
    const int base = blockIdx.x * kGeomPersons;
    const int i = base + threadIdx.x;
    const bool want_input = (joints_input != nullptr) || (img_trans_f64 != nullptr);
    if (i < P) {
        const double x1 = boxes[4 * (size_t)i + 0], y1 = boxes[4 * (size_t)i + 1];
        double w = __dsub_rn(boxes[4 * (size_t)i + 2], x1), h = __dsub_rn(boxes[4 * (size_t)i + 3], y1);
        float cx = __double2float_rn(__dadd_rn(x1, __dmul_rn(w, 0.5)));
        const float cy = __double2float_rn(__dadd_rn(y1, __dmul_rn(h, 0.5)));
        const double rh = __dmul_rn(ratio, h);
        if (w > rh) h = __ddiv_rn(w, ratio);
        else if (w < rh) w = __dmul_rn(h, ratio);
        float sw = __double2float_rn(w), sh = __double2float_rn(h);
        if (cx != -1.0f) { sw = __fmul_rn(sw, scale_mult); sh = __fmul_rn(sh, scale_mult); }
        if (scale_ratio) {
            const float r = __double2float_rn(scale_ratio[i]);
            sw = __fmul_rn(sw, r); sh = __fmul_rn(sh, r);
        }
        if (flip && flip[i]) cx = __fsub_rn(__fsub_rn((float)img_w[i], cx), 1.0f);
        const double rot = rot_deg ? rot_deg[i] : 0.0;
        double inv[6], fwd[6];
        affines_from_center_scale(cx, cy, sw, rot, out_w, out_h, inv, fwd);
#pragma unroll
        for (int e = 0; e < 6; ++e) s_hm[threadIdx.x][e] = fwd[e];
        store_affines((size_t)i, inv, trans_inv, trans_inv_f64);
        if (want_input) {
            affines_from_center_scale(cx, cy, sw, rot, in_w, in_h, inv, fwd);
#pragma unroll
            for (int e = 0; e < 6; ++e) {
                s_in[threadIdx.x][e] = fwd[e];
                if (img_trans_f64) img_trans_f64[6 * (size_t)i + e] = fwd[e];
            }
        }
        if (center) { center[2 * (size_t)i] = cx; center[2 * (size_t)i + 1] = cy; }
        if (scale) { scale[2 * (size_t)i] = sw; scale[2 * (size_t)i + 1] = sh; }
    }
    __syncthreads();
    const int persons = min(kGeomPersons, P - base);
    for (int j = threadIdx.x; j < persons * K; j += kGeomPersons) {
        const int lp = j / K, k = j - lp * K;
        const size_t person = (size_t)(base + lp);
        const bool flipped = flip && flip[person];
        float x, y, v;
        load_joint(joints + person * K * 3, k, flipped, perm, flipped ? (float)img_w[person] : 0.0f, x, y, v);
        const size_t o = (person * K + k) * 3;
        const bool vis = v > 0.0f;
        joints_hm[o + 0] = vis ? affine_row((double)x, (double)y, &s_hm[lp][0]) : x;
        joints_hm[o + 1] = vis ? affine_row((double)x, (double)y, &s_hm[lp][3]) : y;
        joints_hm[o + 2] = v;
        if (joints_input) {
            joints_input[o + 0] = vis ? affine_row((double)x, (double)y, &s_in[lp][0]) : x;
            joints_input[o + 1] = vis ? affine_row((double)x, (double)y, &s_in[lp][3]) : y;
            joints_input[o + 2] = v;
        }
    }
}

// flip_joints (joint half) and/or affine_transform_batch on [P,K,3] joints
__global__ void __launch_bounds__(128)
transform_joints_kernel(const float* __restrict__ joints, const double* __restrict__ trans,
                        const unsigned char* __restrict__ flip, const int* __restrict__ img_w,
                        const int* __restrict__ perm, float* __restrict__ out, long long rows, int K) {
    sp::grid_dep_wait();
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rows) return;
    const size_t person = (size_t)(j / K);
    const int k = (int)(j - (long long)person * K);
    const bool flipped = flip && flip[person];
    float x, y, v;
    load_joint(joints + person * K * 3, k, flipped, perm, flipped ? (float)img_w[person] : 0.0f, x, y, v);
    if (trans && v > 0.0f) {
        const double* t = trans + 6 * person;
        const float nx = affine_row((double)x, (double)y, t), ny = affine_row((double)x, (double)y, t + 3);
        x = nx; y = ny;
    }
    out[3 * (size_t)j + 0] = x;
    out[3 * (size_t)j + 1] = y;
    out[3 * (size_t)j + 2] = v;
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
    return 0;
}

static int center_scale_affine(const float* center, const float* scale, const double* rot_deg, float* trans_inv,
                               double* trans_inv_f64, double* trans_f64, int P, int out_w, int out_h, void* stream) {
    SP_RETURN_IF(P < 0 || out_w <= 0 || out_h <= 0, SP_ERR_BAD_ARGUMENT);
    if (P == 0) return 0;
    SP_RETURN_IF(!center || !scale || (!trans_inv && !trans_inv_f64 && !trans_f64), SP_ERR_BAD_ARGUMENT);
    SP_CUDA(sp_launch(center_scale_affine_kernel, dim3((P + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream),
                      center, scale, rot_deg, trans_inv, trans_inv_f64, trans_f64, P, (double)out_w, (double)out_h));
    return 0;
}

extern "C" int sp_center_scale_affine_f64(const float* center, const float* scale, float* trans_inv,
                                          double* trans_inv_f64, double* trans_f64, int P, int out_w, int out_h,
                                          void* stream) {
    return center_scale_affine(center, scale, nullptr, trans_inv, trans_inv_f64, trans_f64, P, out_w, out_h, stream);
}

extern "C" int sp_center_scale_rot_affine_f64(const float* center, const float* scale, const double* rot_deg,
                                              float* trans_inv, double* trans_inv_f64, double* trans_f64, int P,
                                              int out_w, int out_h, void* stream) {
    return center_scale_affine(center, scale, rot_deg, trans_inv, trans_inv_f64, trans_f64, P, out_w, out_h, stream);
}

extern "C" int sp_train_geometry_f32(const double* boxes, const int* img_w, const float* joints,
                                     const double* scale_ratio, const double* rot_deg, const unsigned char* flip,
                                     const int* perm, float* joints_hm, float* joints_input, float* trans_inv,
                                     double* trans_inv_f64, double* img_trans_f64, float* center, float* scale, int P,
                                     int K, int in_w, int in_h, int out_w, int out_h, float scale_mult, void* stream) {
    SP_RETURN_IF(P < 0 || K <= 0 || in_w <= 0 || in_h <= 0 || out_w <= 0 || out_h <= 0, SP_ERR_BAD_ARGUMENT);
    if (P == 0) return 0;
    SP_RETURN_IF(!boxes || !joints || !joints_hm, SP_ERR_BAD_ARGUMENT);
    SP_RETURN_IF(flip && (!img_w || !perm), SP_ERR_BAD_ARGUMENT);
    SP_CUDA(sp_launch(train_geometry_kernel, dim3((P + kGeomPersons - 1) / kGeomPersons), dim3(kGeomPersons), 0,
                      static_cast<cudaStream_t>(stream), boxes, img_w, joints, scale_ratio, rot_deg, flip, perm, joints_hm,
                      joints_input, trans_inv, trans_inv_f64, img_trans_f64, center, scale, P, K,
                      (double)in_w / (double)in_h, (double)in_w, (double)in_h, (double)out_w, (double)out_h, scale_mult));
    return 0;
}

extern "C" int sp_transform_joints_f32(const float* joints, const double* trans, const unsigned char* flip,
                                       const int* img_w, const int* perm, float* out, int P, int K, void* stream) {
    SP_RETURN_IF(P < 0 || K <= 0, SP_ERR_BAD_ARGUMENT);
    if (P == 0) return 0;
    SP_RETURN_IF(!joints || !out || joints == out, SP_ERR_BAD_ARGUMENT);     /* flipped rows are read across joints */
    SP_RETURN_IF(flip && (!img_w || !perm), SP_ERR_BAD_ARGUMENT);
    const long long rows = (long long)P * K;
    SP_CUDA(sp_launch(transform_joints_kernel, dim3((unsigned)((rows + 127) / 128)), dim3(128), 0,
                      static_cast<cudaStream_t>(stream), joints, trans, flip, img_w, perm, out, rows, K));
    return 0;
}
