// C-ABI plumbing shared by every entry point: thread-local error text, version, device info.
#include "common.cuh"
#include "fsrl_b200.h"
#include <stdarg.h>
#include <string.h>

namespace fsrl {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}
}  // namespace fsrl

extern "C" const char* fsrl_last_error(void) { return fsrl::g_err; }
extern "C" int fsrl_abi_version(void) { return 1; }
namespace fsrl { unsigned long long g_launches = 0; }
extern "C" unsigned long long fsrl_launch_count(void) { return fsrl::g_launches; }
extern "C" int fsrl_sm_count(void) { return fsrl::sm_count(); }

// sizes of the descriptor structs, checked against the ctypes mirrors at import time
extern "C" size_t fsrl_abi_sizeof(int which) {
    switch (which) {
        case 0: return sizeof(fsrl_mlp3_t);
        case 1: return sizeof(fsrl_collect_stats_t);
        case 2: return sizeof(fsrl_rollout_t);
        case 3: return sizeof(fsrl_ppo_update_t);
        case 4: return sizeof(fsrl_netref_t);
        case 5: return sizeof(fsrl_netlist_t);
        case 6: return sizeof(fsrl_engine_t);
        case 7: return sizeof(fsrl_eng_input_t);
        case 8: return sizeof(fsrl_offpolicy_t);
        case 9: return sizeof(fsrl_cpo_t);
        default: return 0;
    }
}
