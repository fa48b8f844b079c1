#include "b2n_common.cuh"

namespace b2n {
thread_local char g_err[512] = {0};
}

extern "C" int b2n_abi_version(void) { return B2N_ABI_VERSION; }
extern "C" const char* b2n_last_error(void) { return b2n::g_err; }
