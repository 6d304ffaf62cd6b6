// Shared declarations for the B200 (sm_100a) noLZSS factorizer kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

namespace nlz {

// status codes shared with include/nolzss_b200.h
enum : int {
    OK = 0,
    ERR_RUNTIME = 1,   // maps to std::runtime_error  -> Python RuntimeError
    ERR_INVALID = 2,   // maps to std::invalid_argument -> Python ValueError
    ERR_CUDA = 3,      // CUDA failure (no device, launch error, OOM) -> RuntimeError
};

void set_error(const char* fmt, ...);

#define NLZ_CK(call)                                                                       \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            nlz::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__,   \
                           __LINE__, cudaGetErrorString(e__));                             \
            return nlz::ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

#define NLZ_TRY(expr)                   \
    do {                                \
        int rc__ = (expr);              \
        if (rc__ != nlz::OK) return rc__; \
    } while (0)

static inline u32 ceil_div_u32(u64 a, u64 b) { return (u32)((a + b - 1) / b); }

constexpr int kNumSM = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// ---- small device helpers -------------------------------------------------------------------
__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 lanemask_lt() {
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

}  // namespace nlz
