// ABI version and error strings.
#include "common.cuh"


UAPS_API int uaps_abi_version(void) { return UAPS_ABI_VERSION; }

UAPS_API const char* uaps_error_string(int code) {
    switch (code) {
        case UAPS_OK: return "ok";
        case UAPS_EINVAL: return "invalid argument (null pointer or non-positive size)";
        case UAPS_ERANGE: return "K, C or pixel count outside the supported range";
        case UAPS_EALIGN: return "pointer not aligned to its element type";
        case UAPS_ENODEV: return "no sm_100 device";
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "unknown uaps error";
}
