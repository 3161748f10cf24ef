// ABI bookkeeping entry points.
#include "common.cuh"
#include "ctrlora_b200.h"

extern "C" int ctrlora_abi_version(void) { return CTRLORA_ABI_VERSION; }

extern "C" const char* ctrlora_last_cuda_error(void) { return cudaGetErrorString(cudaPeekAtLastError()); }

// cudaMemsetAsync behind the ABI: zero-fills become memset nodes (no kernel) -- used for the key-padding columns of V^T
extern "C" int ctrlora_memset_zero(void* ptr, long long bytes, void* stream) {
    if (!ptr || bytes < 0) return CTRLORA_STATUS_BAD_ARGUMENT;
    return cudaMemsetAsync(ptr, 0, static_cast<size_t>(bytes), reinterpret_cast<cudaStream_t>(stream)) == cudaSuccess
               ? CTRLORA_STATUS_OK : CTRLORA_STATUS_CUDA_ERROR;
}
