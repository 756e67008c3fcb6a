// Pieces of the DarkPose target encoder shared by the stand-alone encoder (sp_encode.cu) and the
// fused encode+loss kernel (sp_train.cu). Reference: commons/transforms.py:167-191.
#pragma once
#include "sp_common.cuh"

namespace sp_gauss {

struct JointVerdict {
    float weight;   // value written to weights[b,k]
    bool draw;      // whether a Gaussian is rendered (else the map is zero)
    float mx, my;   // the centre the verdict was made for
};

// Cull test of transforms.py:180-185. NumPy 2 keeps float32 for float32-scalar (+,-) Python
// scalar, so the bounds are float32 sums truncated toward zero by int().
__device__ __forceinline__ JointVerdict judge_joint(float mx, float my, float vis, float reach, int H, int W) {
    const int lo_x = (int)__fsub_rn(mx, reach);
    const int lo_y = (int)__fsub_rn(my, reach);
    const int hi_x = (int)__fadd_rn(__fadd_rn(mx, reach), 1.0f);
    const int hi_y = (int)__fadd_rn(__fadd_rn(my, reach), 1.0f);
    JointVerdict v;
    v.mx = mx;
    v.my = my;
    if (lo_x >= W || lo_y >= H || hi_x < 0 || hi_y < 0) {
        v.weight = 0.0f;
        v.draw = false;
    } else {
        v.weight = vis;
        v.draw = vis > 0.5f;
    }
    return v;
}

__device__ __forceinline__ double gauss_factor(int p, float mu, double denom) {
    const double d = (double)p - (double)mu;
    return exp(__ddiv_rn(-__dmul_rn(d, d), denom));
}

}  // namespace sp_gauss
