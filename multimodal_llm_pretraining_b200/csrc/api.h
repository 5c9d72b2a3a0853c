// Host-side helpers for the C ABI: thread-local error string, launch checks. Internal to libb200pt.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/b200pt.h"

namespace b200 {

char* err_buf();  // thread-local, 512 bytes
int fail(int code, const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(-2, "%s: launch failed: %s", what, cudaGetErrorString(e));
    return 0;
}

#define B200_REQUIRE(cond, ...)                         \
    do {                                                \
        if (!(cond)) return ::b200::fail(-1, __VA_ARGS__); \
    } while (0)

// Every kernel goes through launch_k: same as kern<<<grid, block, smem, st>>>(args...) plus the programmatic-dependent-launch
// attribute (common.cuh: pdl_trigger / pdl_wait) when B200_PDL=1 (see api.cu: pdl_enabled for the measurement).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline cudaStream_t as_stream(b200_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int num_sms();  // cached multiprocessor count of the current device (148 on B200)

}  // namespace b200
