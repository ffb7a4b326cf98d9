// Version / error-string entry points of libpcbridge.
#include "pcb_common.cuh"

PCB_API int pcb_version(void) { return PCB_VERSION; }

PCB_API const char *pcb_error_string(int code)
{
    switch (code) {
        case 0: return "success";
        case PCB_EINVAL: return "pcbridge: null pointer or non-positive size";
        case PCB_ERANGE: return "pcbridge: size outside the supported envelope";
        case PCB_EALIGN: return "pcbridge: pointer not aligned as required";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "pcbridge: unknown error";
}
