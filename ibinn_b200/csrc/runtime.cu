// Device gate: the library serves sm_100 (B200) only and has no other path.
#include "common.cuh"

#ifndef IBINN_EMU
#include <stdlib.h>
unsigned long long ibinn_launch_counter = 0;
static int pdl_from_env() { const char* e = getenv("IBINN_PDL"); return !(e && e[0] == '0'); }
int ibinn_pdl_enabled = pdl_from_env();
#endif

char ibinn_errbuf[256] = "";

extern "C" int ibinn_abi_version(void) { return 1; }

// text of the most recent failure reported by this library ("" if none)
extern "C" const char* ibinn_last_error(void) { return ibinn_errbuf; }

// number of kernels this library has enqueued in this process (host-side counter; bench.py reports it)
extern "C" int64_t ibinn_launch_count(void) {
#ifdef IBINN_EMU
    return 0;
#else
    return (int64_t)ibinn_launch_counter;
#endif
}

extern "C" int ibinn_check_device(void) {
#ifdef IBINN_EMU
    return 0;
#else
    static int cached[64] = {0};   // 0 unknown, 1 ok, 2 wrong arch
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev] == 1 ? 0 : IBINN_EARCH;
    int major = 0, minor = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (e != cudaSuccess) return (int)e;
    const bool ok = (major == 10 && minor == 0);
    if (dev >= 0 && dev < 64) cached[dev] = ok ? 1 : 2;
    return ok ? 0 : IBINN_EARCH;
#endif
}
