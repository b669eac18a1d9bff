// api.cu — library-level entry points of libgraingnn_b200.
#include "common.cuh"

extern "C" const char* gg_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case GG_EINVAL: return "graingnn_b200: invalid argument";
        case GG_ERANGE: return "graingnn_b200: edge endpoint out of range";
        case GG_EALIGN: return "graingnn_b200: pointer or leading dimension not 16-byte aligned";
        case GG_ENOSPC: return "graingnn_b200: workspace too small";
        case GG_EARCH: return "graingnn_b200: device is not sm_100";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "graingnn_b200: unknown error";
    }
}

extern "C" int gg_version(void) { return 100; }

extern "C" int gg_device_is_sm100(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}
