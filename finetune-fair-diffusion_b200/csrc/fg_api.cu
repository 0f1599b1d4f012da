// ABI version and error strings.
#include "fg_common.cuh"

extern "C" int fg_abi_version(void) { return FG_ABI_VERSION; }

extern "C" const char* fg_error_string(int code) {
    switch (code) {
        case FG_OK: return "ok";
        case FG_ERR_INVALID_ARG: return "fairguide: invalid argument";
        case FG_ERR_DTYPE: return "fairguide: unsupported dtype";
        case FG_ERR_LIMIT: return "fairguide: size limit exceeded";
        case FG_ERR_WORKSPACE: return "fairguide: workspace missing or too small";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "fairguide: unknown error";
}
