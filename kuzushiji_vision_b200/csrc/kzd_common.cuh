// Shared helpers for libkzd (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/kzd.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libkzd is written for sm_100a (B200) only"
#endif

namespace kzd {

void set_error(const char* fmt, ...);

#define KZD_CHECK_ARG(cond, ...)            \
    do {                                    \
        if (!(cond)) {                      \
            kzd::set_error(__VA_ARGS__);    \
            return KZD_EINVAL;              \
        }                                   \
    } while (0)

#define KZD_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            kzd::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return (int)e__;                                                             \
        }                                                                                \
    } while (0)

#define KZD_LAUNCH_CHECK(name)                                                           \
    do {                                                                                 \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) {                                                        \
            kzd::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));    \
            return (int)e__;                                                             \
        }                                                                                \
    } while (0)

int sm_count();

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- device helpers --------------------------------------------------------------------------

__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    // streaming 128-bit load: read once, do not keep in L1
    return __ldcs(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

// ---- half-precision words (DT: 1 fp16, 2 bf16) ------------------------------------------------------

// eight consecutive half-precision elements (one 16-byte word) <-> fp32
template <int DT>
__device__ __forceinline__ void unpack8(const uint4 w, float (&v)[8]) {
    const uint32_t u[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (DT == 1) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u[k]));
            v[2 * k] = f.x, v[2 * k + 1] = f.y;
        } else {
            v[2 * k] = __uint_as_float(u[k] << 16), v[2 * k + 1] = __uint_as_float(u[k] & 0xffff0000u);
        }
    }
}
template <int DT>
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    uint32_t u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (DT == 1) {
            const __half2 h = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
            u[k] = *reinterpret_cast<const uint32_t*>(&h);
        } else {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
            u[k] = *reinterpret_cast<const uint32_t*>(&h);
        }
    }
    return make_uint4(u[0], u[1], u[2], u[3]);
}

// ---- mbarrier / TMA bulk-copy PTX (cp.async.bulk: SASS UBLKCP) -------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, 16 B aligned).
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- NMS candidate pool (shared by nms_filter_kernel and the filtering decode kernel) ---------------

typedef unsigned long long u64;

// key = score_bits<<32 | ~order: a plain 64-bit descending order of the keys reproduces torchvision's
// stable descending score sort (order = anchor index, or anchor*nc+class for multi-label rows)
__device__ __forceinline__ void write_candidate(int slot, int cap, float x1, float y1, float x2, float y2, float conf,
                                                int cls, uint32_t order, unsigned long long* __restrict__ keys,
                                                float* __restrict__ rows) {
    if (slot < cap) {
        keys[slot] = ((unsigned long long)__float_as_uint(conf) << 32) | (unsigned long long)(0xffffffffu - order);
        float2* r = reinterpret_cast<float2*>(rows + (size_t)slot * 6);  // 24-byte rows are 8-byte aligned
        r[0] = make_float2(x1, y1), r[1] = make_float2(x2, y2), r[2] = make_float2(conf, (float)cls);
    }
}

// Warp-aggregated slot allocation: every lane asks for `want` (0..VEC) consecutive slots.
__device__ __forceinline__ int warp_alloc(int want, int32_t* __restrict__ count) {
    const int lane = threadIdx.x & 31;
    int incl = want;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return 0;
    int base = 0;
    if (lane == 31) base = atomicAdd(count, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    return base + incl - want;
}

}  // namespace kzd
