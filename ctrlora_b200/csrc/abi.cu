// ABI bookkeeping entry points.
#include "common.cuh"
#include "ctrlora_b200.h"

extern "C" int ctrlora_abi_version(void) { return CTRLORA_ABI_VERSION; }

extern "C" const char* ctrlora_last_cuda_error(void) { return cudaGetErrorString(cudaPeekAtLastError()); }
