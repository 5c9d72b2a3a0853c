// libb200pt: ABI version, thread-local error reporting, one-time init.
#include <stdlib.h>

#include "api.h"

namespace b200 {

static thread_local char g_err[512] = "";

char* err_buf() { return g_err; }

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static int g_num_sms = 0;
int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            g_num_sms = n;
        else
            g_num_sms = 148;  // B200
    }
    return g_num_sms;
}

bool pdl_enabled() {
    // Off by default: on the power-capped B200s of this pool (1 kW, ~1.3 GHz under load) closing the inter-kernel gaps did
    // not raise throughput (Pythia-1b step, same box, alternating runs: 178.4 / 178.6 k tokens/s without, 177.7 / 177.7 k with
    // — the step is energy-limited, and idle gaps are what lets the clock boost between them). B200_PDL=1 enables it.
    static const bool on = getenv("B200_PDL") && atoi(getenv("B200_PDL")) != 0;
    return on;
}

int resolve_driver();  // gemm.cu

}  // namespace b200

extern "C" int b200_abi_version(void) { return B200PT_ABI_VERSION; }
#ifdef B200_ELEM_FP16
extern "C" int b200_elem_dtype(void) { return 1; }  // this build reads every 16-bit tensor as IEEE fp16
#else
extern "C" int b200_elem_dtype(void) { return 0; }  // bf16
#endif
extern "C" const char* b200_last_error(void) { return b200::err_buf(); }

extern "C" int b200_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return b200::fail(-4, "b200_init: no CUDA device available (%s)", cudaGetErrorString(e));
    if (device < 0 || device >= n) return b200::fail(-1, "b200_init: device %d out of range (%d devices)", device, n);
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
    if (major != 10) return b200::fail(-4, "b200_init: device %d is sm_%d%d; libb200pt is built for sm_100a only", device, major, minor);
    return b200::resolve_driver();
}
