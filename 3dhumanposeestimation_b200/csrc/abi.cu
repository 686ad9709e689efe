// abi.cu -- version and error strings of the C ABI (include/pose_b200.h)
#include "common.cuh"

POSE_API int pose_b200_abi_version(void) { return 2; }

namespace pose {
pose_step_state *g_step_state = nullptr;

__global__ void step_tick_kernel(pose_step_state *st) {
    const uint32_t c = st->counter + 1u;
    st->counter = c;
    st->adam_step += 1;
    uint32_t x = c * 0x9E3779B9u;                       // lowbias32 of a Weyl sequence: a fresh 32-bit key per step
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    st->drop_key = x;
}
}  // namespace pose

POSE_API int pose_step_state_bind(pose_step_state *dev_state) {
    pose::g_step_state = dev_state;
    return POSE_OK;
}

POSE_API int pose_step_tick(pose_step_state *dev_state, pose_stream_t stream) {
    if (!dev_state) return POSE_E_NULL;
    pose::step_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dev_state);
    return pose::launch_status();
}

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
