#include "gs_common.cuh"

extern "C" int gs_abi_version(void) { return GS_ABI_VERSION; }

extern "C" const char* gs_strerror(int code) {
    switch (code) {
        case GS_OK: return "ok";
        case GS_EINVAL: return "gsage: invalid argument (null pointer or bad size)";
        case GS_EALIGN: return "gsage: pointer or leading dimension is not 16-byte aligned";
        case GS_ENOSUP: return "gsage: shape not supported by this kernel";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "gsage: unknown error";
    }
}
