// Shared helpers for libidgrec_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/idgrec.h"

// symmetric peer slabs of the row-partitioned multi-GPU path (csrc/peers.cu)
struct idg_peers {
    char* local_base = nullptr;
    int64_t bytes = 0;
    int rank = 0, world = 1;
    char* bases[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    char* mc_base = nullptr;  // NVSwitch multicast mapping of the same slab (NVLS): one store reaches every GPU
    unsigned long long timeout_ns = 20000000000ull;  // bound of one flag-barrier wait (idg_peers_set_timeout_ms)
};

namespace idg {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
    snprintf(g_err, sizeof(g_err), fmt, a, b, c);
    return code;
}
inline int cuda_fail(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
}

#define IDG_CUDA(expr)                                          \
    do {                                                        \
        cudaError_t _e = (expr);                                \
        if (_e != cudaSuccess) return idg::cuda_fail(_e, #expr); \
    } while (0)

#define IDG_LAUNCH_CHECK(name)                                   \
    do {                                                        \
        idg::g_launches.fetch_add(1, std::memory_order_relaxed); \
        cudaError_t _e = cudaGetLastError();                    \
        if (_e != cudaSuccess) return idg::cuda_fail(_e, name);  \
    } while (0)

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldcs4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void stcs4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ void stcg4(float* p, float4 v) { __stcg(reinterpret_cast<float4*>(p), v); }
// store through an NVSwitch multicast address: the switch replicates it into every GPU's copy of the slab
__device__ __forceinline__ void multimem_st4(float* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4fma(float w, float4 x, float4 a) {
    return make_float4(fmaf(w, x.x, a.x), fmaf(w, x.y, a.y), fmaf(w, x.z, a.z), fmaf(w, x.w, a.w));
}
__device__ __forceinline__ float4 f4shfl_xor(float4 v, int m) {
    return make_float4(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m),
                       __shfl_xor_sync(0xffffffffu, v.z, m), __shfl_xor_sync(0xffffffffu, v.w, m));
}
__device__ __forceinline__ float sgn(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

}  // namespace idg
