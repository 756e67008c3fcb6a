// ABI bookkeeping: version, error strings, device query.
#include "sp_common.cuh"

extern "C" int sp_abi_version(void) { return SP_ABI_VERSION; }

extern "C" const char* sp_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case SP_ERR_BAD_ARGUMENT: return "simple_pose_b200: bad argument (null pointer, bad size or unsupported value)";
        case SP_ERR_BAD_ALIGNMENT: return "simple_pose_b200: base pointer is not 16-byte aligned";
        case SP_ERR_WORKSPACE: return "simple_pose_b200: workspace too small";
        case SP_ERR_UNSUPPORTED: return "simple_pose_b200: shape not supported by the kernels";
        default: break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "simple_pose_b200: unknown error code";
}

extern "C" int sp_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    SP_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    SP_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return 0;
}
