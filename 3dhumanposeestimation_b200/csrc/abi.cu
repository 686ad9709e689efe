// abi.cu -- version and error strings of the C ABI (include/pose_b200.h)
#include "common.cuh"

POSE_API int pose_b200_abi_version(void) { return 1; }

POSE_API const char *pose_b200_error_string(int code) {
    switch (code) {
        case POSE_OK: return "ok";
        case POSE_E_NULL: return "required pointer is NULL";
        case POSE_E_SHAPE: return "unsupported or inconsistent shape";
        case POSE_E_WORKSPACE: return "workspace too small";
        case POSE_E_UNSUPPORTED: return "request not covered by the sm_100a kernels";
        case POSE_E_ALIGN: return "pointer alignment";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown pose_b200 error";
}
