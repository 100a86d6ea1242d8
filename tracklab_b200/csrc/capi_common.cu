// Library-wide C-ABI helpers.
#include "tk_common.cuh"
#include "trackkern.h"

static int g_last_cuda_error = 0;
void tk_set_last_cuda_error(int e) { g_last_cuda_error = e; }

extern "C" {
int tk_abi_version(void) { return TK_ABI_VERSION; }
int tk_last_cuda_error(void) { return g_last_cuda_error; }
}
