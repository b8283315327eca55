// ABI bookkeeping for libepnet_b200.so (include/epnet_b200.h).
#include "common.cuh"

EPNET_API int epnet_abi_version(void) { return 1; }

EPNET_API const char *epnet_error_string(int code)
{
    if (code == EPNET_OK) return "ok";
    if (code == EPNET_ERR_BAD_ARG) return "epnet_b200: bad argument (null pointer, negative size or misaligned buffer)";
    return cudaGetErrorString((cudaError_t)code);
}
