// DFL decode kernels (sm_100a).
//
//  * decode_infer : Detect._inference + DFL.forward + dist2bbox(xywh) + sigmoid, reading the
//                   per-level NCHW maps directly (reference: nn/modules/head.py:100-131,
//                   nn/modules/block.py:71-74, utils/tal.py:348-357).
//                   Optional extras from the same read of the maps: the training-side boxes (DUAL) and
//                   the filter stage of non_max_suppression, utils/ops.py:234-286 (FILTER).
//  * decode_train : v8DetectionLoss.bbox_decode + dist2bbox(xyxy) on the anchor-major (b,A,64)
//                   layout (reference: utils/loss.py:197-204), and its backward.
//
// Both are pure streaming kernels bound by HBM bandwidth (272 B per anchor).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <type_traits>

#include "kzd_common.cuh"

namespace kzd {

// ------------------------------------------------------------------------------------------
// inference decode
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct InferParams {
    const void* ptr[KZD_MAX_LEVELS];
    int hw[KZD_MAX_LEVELS];
    int w[KZD_MAX_LEVELS];
    int off[KZD_MAX_LEVELS + 1];
    float stride[KZD_MAX_LEVELS];
    int n_levels, b, nc, A;
};

// VEC consecutive elements of a head map -> fp32 registers (128-bit loads for fp32 x4, 64-bit for half x4)
template <int VEC, typename T>
struct In;
template <>
struct In<4, float> {
    static __device__ __forceinline__ void load(float (&v)[4], const float* p) {
        const float4 t = __ldcs(reinterpret_cast<const float4*>(p));
        v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
    }
};
template <>
struct In<4, __half> {
    static __device__ __forceinline__ void load(float (&v)[4], const __half* p) {
        const uint2 t = __ldcs(reinterpret_cast<const uint2*>(p));
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
        v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y;
    }
};
template <>
struct In<4, __nv_bfloat16> {
    static __device__ __forceinline__ void load(float (&v)[4], const __nv_bfloat16* p) {
        const uint2 t = __ldcs(reinterpret_cast<const uint2*>(p));
        v[0] = __uint_as_float(t.x << 16), v[1] = __uint_as_float(t.x & 0xffff0000u);
        v[2] = __uint_as_float(t.y << 16), v[3] = __uint_as_float(t.y & 0xffff0000u);
    }
};
template <typename T>
struct In<1, T> {
    static __device__ __forceinline__ void load(float (&v)[1], const T* p) { v[0] = (float)*p; }
};

template <int VEC>
struct Out;
template <>
struct Out<4> {
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    }
};
template <>
struct Out<1> {
    static __device__ __forceinline__ void store(float* p, const float (&v)[1]) { __stcs(p, v[0]); }
};

// One thread decodes VEC consecutive anchors of one image: 4 sides x 16 bins are read as 64
// coalesced VEC-wide loads (anchors are the fastest dimension of the NCHW maps), the softmax
// expectation is evaluated in registers, and the 4+nc output rows are written VEC-wide.
// DUAL: additionally emit the training-side decode of the same maps - xyxy in grid units (b,A,4) and
// its pixel-space copy (utils/loss.py:197-204,236) - so that validation-during-training
// (engine/validator.py:180-185: eval-mode Detect, then model.loss on the same maps) reads the head
// outputs once instead of twice.
// FILTER: additionally do the first stage of non_max_suppression (utils/ops.py:234-286, best-class path)
// on the values just computed: y's box rows are written as xyxy (the state ops.py:243-244 leaves the
// caller's tensor in) and every anchor above conf is appended to the candidate pool, exactly as
// nms_filter_kernel would after re-reading y - which saves that pass (a read of y and the write-back).
struct FilterArgs {
    float conf;
    u64* keys;
    float* rows;
    int32_t* counts;
    int cap;
};

// WIDE: heads with many classes (the kuzushiji recogniser's ~4,300 Unicode classes). The class loop then carries
// nearly all of the traffic AND of the instructions: with the IEEE division of the sigmoid the kernel is
// issue-bound (4.11 ms for 4 pages at nc = 4300, 0.70 of the HBM peak); the approximate division (MUFU.RCP + one
// multiply, <= 2 ulp) brings it to 3.11 ms = 0.93. A separate instantiation, because the extra registers of
// that loop cost the single-class kernel a resident CTA per SM (107.7 -> 110 us).
template <int VEC, typename T, bool DUAL, bool FILTER, bool WIDE>
__global__ void __launch_bounds__(128) decode_infer_kernel(const InferParams p, float* __restrict__ y,
                                                           float* __restrict__ out_xyxy,
                                                           float* __restrict__ out_px, const FilterArgs f) {
    const int bi = blockIdx.y;
    const int a0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    const bool live = a0 < p.A;
    if (!FILTER && !live) return;  // FILTER: whole warps stay alive for the slot allocation
    float ox[VEC], oy[VEC], ow[VEC], oh[VEC], best[VEC];
    int cls[VEC];
    if (live) {
        int l = 0;
#pragma unroll
        for (int k = 1; k < KZD_MAX_LEVELS; ++k)
            if (k < p.n_levels && a0 >= p.off[k]) l = k;
        const int hw = p.hw[l];
        const int pos = a0 - p.off[l];
        const int no = 4 * KZD_REG_MAX + p.nc;
        const T* __restrict__ src = reinterpret_cast<const T*>(p.ptr[l]) + (size_t)bi * no * hw + pos;

        float d[4][VEC];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            float v[KZD_REG_MAX][VEC];
#pragma unroll
            for (int k = 0; k < KZD_REG_MAX; ++k) In<VEC, T>::load(v[k], src + (size_t)(s * KZD_REG_MAX + k) * hw);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                float m = v[0][i];
#pragma unroll
                for (int k = 1; k < KZD_REG_MAX; ++k) m = fmaxf(m, v[k][i]);
                // exp(v - m) = ex2(v*log2e - m*log2e): one FFMA + one MUFU per bin; weights per 4-bin chunk
                const float L2E = 1.4426950408889634f;
                const float ml = -m * L2E;
                float sum = 0.f, wsum = 0.f;
#pragma unroll
                for (int c = 0; c < KZD_REG_MAX; c += 4) {
                    const float e0 = ex2_approx(fmaf(v[c][i], L2E, ml)), e1 = ex2_approx(fmaf(v[c + 1][i], L2E, ml));
                    const float e2 = ex2_approx(fmaf(v[c + 2][i], L2E, ml)), e3 = ex2_approx(fmaf(v[c + 3][i], L2E, ml));
                    const float sc = (e0 + e1) + (e2 + e3);
                    sum += sc;
                    wsum += fmaf((float)c, sc, fmaf(3.f, e3, fmaf(2.f, e2, e1)));
                }
                d[s][i] = wsum / sum;
            }
        }
        const float st = p.stride[l];
        const int w = p.w[l];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const int q = pos + i;
            const float ax = (float)(q % w) + 0.5f;
            const float ay = (float)(q / w) + 0.5f;
            const float x1 = ax - d[0][i], y1 = ay - d[1][i];
            const float x2 = ax + d[2][i], y2 = ay + d[3][i];
            ox[i] = ((x1 + x2) * 0.5f) * st;
            oy[i] = ((y1 + y2) * 0.5f) * st;
            ow[i] = (x2 - x1) * st;
            oh[i] = (y2 - y1) * st;
            if (DUAL) {
                const size_t o = ((size_t)bi * p.A + a0 + i) * 4;
                __stcs(reinterpret_cast<float4*>(out_xyxy + o), make_float4(x1, y1, x2, y2));
                if (out_px)
                    __stcs(reinterpret_cast<float4*>(out_px + o),
                           make_float4(__fmul_rn(x1, st), __fmul_rn(y1, st), __fmul_rn(x2, st), __fmul_rn(y2, st)));
            }
        }
        if (FILTER) {  // xywh -> xyxy from the fp32 values y would have held (utils/ops.py:429-432)
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float hw2 = ow[i] * 0.5f, hh2 = oh[i] * 0.5f;
                const float cx = ox[i], cy = oy[i];
                ox[i] = cx - hw2, oy[i] = cy - hh2, ow[i] = cx + hw2, oh[i] = cy + hh2;
            }
        }
        float* __restrict__ dst = y + (size_t)bi * (4 + p.nc) * p.A + a0;
        Out<VEC>::store(dst, ox);
        Out<VEC>::store(dst + (size_t)p.A, oy);
        Out<VEC>::store(dst + 2 * (size_t)p.A, ow);
        Out<VEC>::store(dst + 3 * (size_t)p.A, oh);
        const T* __restrict__ csrc = src + (size_t)(4 * KZD_REG_MAX) * hw;
        float* __restrict__ cdst = dst + 4 * (size_t)p.A;
        // class rows in groups of CG: all loads of a group are issued before any is used (a multi-class
        // kuzushiji head has thousands of rows, so this loop then carries almost all of the traffic)
        constexpr int CG = 8;
        for (int c0 = 0; c0 < p.nc; c0 += CG) {
            float v[CG][VEC];
#pragma unroll
            for (int k = 0; k < CG; ++k)
                if (c0 + k < p.nc) In<VEC, T>::load(v[k], csrc + (size_t)(c0 + k) * hw);
#pragma unroll
            for (int k = 0; k < CG; ++k) {
                const int c = c0 + k;
                if (c >= p.nc) break;
                float o[VEC];
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    // |err| < 2e-7: ex2.approx + IEEE division; WIDE: approximate division, |err| < 4e-7
                    o[i] = WIDE ? __fdividef(1.f, 1.f + __expf(-v[k][i])) : 1.f / (1.f + __expf(-v[k][i]));
                    if (FILTER && (c == 0 || o[i] > best[i])) best[i] = o[i], cls[i] = c;  // first maximum wins (ops.py:274)
                }
                Out<VEC>::store(cdst + (size_t)c * p.A, o);
            }
        }
    }
    if (FILTER) {
        int want = 0;
        bool is_cand[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            is_cand[i] = live && best[i] > f.conf;
            want += is_cand[i] ? 1 : 0;
        }
        int slot = warp_alloc(want, f.counts + bi);
        if (want) {
            u64* k_img = f.keys + (size_t)bi * f.cap;
            float* r_img = f.rows + (size_t)bi * f.cap * 6;
#pragma unroll
            for (int i = 0; i < VEC; ++i)
                if (is_cand[i]) {
                    write_candidate(slot, f.cap, ox[i], oy[i], ow[i], oh[i], best[i], cls[i], (uint32_t)(a0 + i), k_img,
                                    r_img);
                    ++slot;
                }
        }
    }
}

// ------------------------------------------------------------------------------------------
// training decode: TMA bulk-copy pipeline
// ------------------------------------------------------------------------------------------

constexpr int TR_ANCH = 64;                     // anchors per tile (backward kernel)
constexpr int TR_FLOATS = TR_ANCH * 64;         // 4096 floats = 16 KB
constexpr int TR_STAGES = 4;                    // 64 KB of loads in flight per CTA
constexpr int TR_THREADS = TR_ANCH * 4;         // one thread per (anchor, side)
constexpr int TR_SMEM = TR_STAGES * TR_FLOATS * 4;

// The 16 bins of one (anchor, side) from the shared-memory tile, as fp32, plus the bin index of each
// register. fp32 rows are 64 B (four 16-byte chunks), half rows 32 B (two chunks); the chunks are visited
// in a per-thread rotated order so that every quarter-warp hits 8 distinct bank groups.
template <typename T>
struct Bins;
template <>
struct Bins<float> {
    static __device__ __forceinline__ void load(const void* tile, int tid, float (&v)[16], int (&first)[4]) {
        const float* row = reinterpret_cast<const float*>(tile) + tid * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int q = (j + (tid >> 1)) & 3;
            const float4 t = *reinterpret_cast<const float4*>(row + q * 4);
            v[4 * j] = t.x, v[4 * j + 1] = t.y, v[4 * j + 2] = t.z, v[4 * j + 3] = t.w;
            first[j] = q * 4;
        }
    }
};
template <typename H>
struct BinsHalf {
    static __device__ __forceinline__ void load(const void* tile, int tid, float (&v)[16], int (&first)[4]) {
        const uint4* row = reinterpret_cast<const uint4*>(tile) + tid * 2;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int q = (j + (tid >> 2)) & 1;
            const uint4 t = row[q];
            const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (std::is_same<H, __half>::value) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
                    v[8 * j + 2 * k] = f.x, v[8 * j + 2 * k + 1] = f.y;
                } else {
                    v[8 * j + 2 * k] = __uint_as_float(w[k] << 16), v[8 * j + 2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
                }
            }
            first[2 * j] = q * 8, first[2 * j + 1] = q * 8 + 4;
        }
    }
};
template <>
struct Bins<__half> : BinsHalf<__half> {};
template <>
struct Bins<__nv_bfloat16> : BinsHalf<__nv_bfloat16> {};

// Softmax expectation of the 16 bins of one (anchor, side).
// exp(v - m) = ex2(v*log2e - m*log2e): one FFMA + one MUFU per bin; the bin weights are applied per
// 4-bin group (sum_k (f+k) e_k = f * sum_k e_k + (e_1 + 2 e_2 + 3 e_3)).
template <typename T>
__device__ __forceinline__ void side_stats(const void* tile, int tid, float& d, float& m_out, float& sum_out) {
    float v[16];
    int first[4];
    Bins<T>::load(tile, tid, v, first);
    float m = v[0];
#pragma unroll
    for (int k = 1; k < 16; ++k) m = fmaxf(m, v[k]);
    const float L2E = 1.4426950408889634f;
    const float ml = -m * L2E;
    float sum = 0.f, wsum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float base = __uint_as_float(0x4b000000u | (uint32_t)first[j]) - 8388608.f;  // exact int->float
        const float e0 = ex2_approx(fmaf(v[4 * j], L2E, ml)), e1 = ex2_approx(fmaf(v[4 * j + 1], L2E, ml));
        const float e2 = ex2_approx(fmaf(v[4 * j + 2], L2E, ml)), e3 = ex2_approx(fmaf(v[4 * j + 3], L2E, ml));
        const float sc = (e0 + e1) + (e2 + e3);
        const float tc = fmaf(3.f, e3, fmaf(2.f, e2, e1));
        sum += sc;
        wsum += fmaf(base, sc, tc);
    }
    d = wsum / sum;
    m_out = m;
    sum_out = sum;
}

// Persistent CTAs; thread 0 is the TMA producer (cp.async.bulk of one ANCH*256 B tile per stage, signalled
// on an mbarrier), all ANCH*4 threads consume: thread t owns (anchor t/4, side t%4) of the tile and
// writes output element t of the tile's ANCH*4 contiguous floats. IDX = int when b*A < 2^31.
template <int ANCH, int STAGES, typename IDX, typename T>
__global__ void __launch_bounds__(ANCH * 4) decode_train_kernel(const T* __restrict__ pred,
                                                                const float* __restrict__ anc,
                                                                const float* __restrict__ stride, int A, IDX N,
                                                                float* __restrict__ out,
                                                                float* __restrict__ out_px) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* buf = reinterpret_cast<T*>(smem_raw);
    __shared__ uint64_t full[STAGES];
    constexpr int FLOATS = ANCH * 64;  // elements per tile
    const int tid = threadIdx.x;
    const IDX n_tiles = (N + ANCH - 1) / ANCH;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](IDX tile, int stage) {
        const IDX first = tile * ANCH;
        const IDX cnt = (N - first < ANCH) ? (N - first) : (IDX)ANCH;
        const uint32_t bytes = (uint32_t)cnt * (uint32_t)(64 * sizeof(T));
        mbar_expect_tx(&full[stage], bytes);
        tma_load_1d(buf + stage * FLOATS, pred + (size_t)first * 64, bytes, &full[stage]);
    };
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            const IDX tile = blockIdx.x + (IDX)s * gridDim.x;
            if (tile < n_tiles) issue(tile, s);
        }
    }
    const int a_local = tid >> 2, side = tid & 3;
    int stage = 0;
    uint32_t parity = 0;
    for (IDX tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const IDX n = tile * ANCH + a_local;
        // anchor (and stride) fetched before waiting on the tile so their latency hides behind the TMA
        const int a_img = (int)((typename std::make_unsigned<IDX>::type)n % (typename std::make_unsigned<IDX>::type)A);
        float ap = 0.f, st = 0.f;
        if (n < N) {
            ap = __ldg(anc + a_img * 2 + (side & 1));
            if (out_px) st = __ldg(stride + a_img);
        }
        mbar_wait(&full[stage], parity);
        if (n < N) {
            float d, m, s;
            side_stats<T>(buf + stage * FLOATS, tid, d, m, s);
            const float v = (side < 2) ? (ap - d) : (ap + d);
            out[(size_t)n * 4 + side] = v;
            if (out_px) out_px[(size_t)n * 4 + side] = __fmul_rn(v, st);  // utils/loss.py:236
        }
        __syncthreads();  // every thread is done with this stage before it is refilled
        if (tid == 0) {
            const IDX next = tile + (IDX)STAGES * gridDim.x;
            if (next < n_tiles) issue(next, stage);
        }
        if (++stage == STAGES) stage = 0, parity ^= 1u;
    }
}

// Backward: d = sum_j j p_j  =>  dd/dx_j = p_j (j - d); x1y1 = a - lt, x2y2 = a + rb.
// T = float, or the AMP flow's fp16 / bf16 pred_dist: read natively, the gradient rounded once into the same dtype.
template <typename T>
__global__ void __launch_bounds__(TR_THREADS) decode_train_bwd_kernel(const T* __restrict__ pred,
                                                                      const float* __restrict__ gout,
                                                                      long long N, T* __restrict__ gin) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* buf = reinterpret_cast<T*>(smem_raw);
    __shared__ uint64_t full[TR_STAGES];
    const int tid = threadIdx.x;
    const long long n_tiles = (N + TR_ANCH - 1) / TR_ANCH;
    if (tid == 0) {
        for (int s = 0; s < TR_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](long long tile, int stage) {
        const long long first = tile * TR_ANCH;
        const long long cnt = (N - first < TR_ANCH) ? (N - first) : TR_ANCH;
        const uint32_t bytes = (uint32_t)cnt * (uint32_t)(64 * sizeof(T));
        mbar_expect_tx(&full[stage], bytes);
        tma_load_1d(buf + stage * TR_FLOATS, pred + first * 64, bytes, &full[stage]);
    };
    if (tid == 0) {
        for (int s = 0; s < TR_STAGES; ++s) {
            const long long tile = blockIdx.x + (long long)s * gridDim.x;
            if (tile < n_tiles) issue(tile, s);
        }
    }
    const int a_local = tid >> 2, side = tid & 3;
    for (long long it = 0;; ++it) {
        const long long tile = blockIdx.x + it * gridDim.x;
        if (tile >= n_tiles) break;
        const int stage = (int)(it % TR_STAGES);
        mbar_wait(&full[stage], (uint32_t)((it / TR_STAGES) & 1));
        const long long n = tile * TR_ANCH + a_local;
        if (n < N) {
            const T* tile_ptr = buf + stage * TR_FLOATS;
            float d, m, s;
            side_stats<T>(tile_ptr, tid, d, m, s);
            float g = gout[n * 4 + side];
            g = (side < 2) ? -g : g;
            const float inv = g / s;
            // the row again, in the registers' (rotated) order: first[j] = bin of register 4 j
            float v[16];
            int first[4];
            Bins<T>::load(tile_ptr, tid, v, first);
            float o[16];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k) o[4 * j + k] = inv * __expf(v[4 * j + k] - m) * ((float)(first[j] + k) - d);
            T* dst = gin + n * 64 + side * 16;
            if (std::is_same<T, float>::value) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + first[j]) =
                        make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            } else {  // registers 8 jj .. 8 jj + 7 hold eight consecutive bins starting at first[2 jj]
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    float w8[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) w8[k] = o[8 * jj + k];
                    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(dst) + first[2 * jj]) =
                        std::is_same<T, __half>::value ? pack8<1>(w8) : pack8<2>(w8);
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            const long long next = tile + (long long)TR_STAGES * gridDim.x;
            if (next < n_tiles) issue(next, stage);
        }
    }
}

}  // namespace kzd

using namespace kzd;

constexpr int WIDE_NC = 32;  // class count from which the WIDE instantiation is used

template <int VEC, typename T, bool WIDE>
static void launch_infer_vec(const InferParams& p, dim3 grid, float* y, float* out_xyxy, float* out_px,
                             const FilterArgs* f, cudaStream_t st) {
    const FilterArgs none{0.f, nullptr, nullptr, nullptr, 0};
    if (f) {
        if (out_xyxy) decode_infer_kernel<VEC, T, true, true, WIDE><<<grid, 128, 0, st>>>(p, y, out_xyxy, out_px, *f);
        else decode_infer_kernel<VEC, T, false, true, WIDE><<<grid, 128, 0, st>>>(p, y, nullptr, nullptr, *f);
    } else {
        if (out_xyxy) decode_infer_kernel<VEC, T, true, false, WIDE><<<grid, 128, 0, st>>>(p, y, out_xyxy, out_px, none);
        else decode_infer_kernel<VEC, T, false, false, WIDE><<<grid, 128, 0, st>>>(p, y, nullptr, nullptr, none);
    }
}

template <typename T>
static void launch_infer(const InferParams& p, int total, int b, bool vec4, float* y, float* out_xyxy, float* out_px,
                         const FilterArgs* f, cudaStream_t st) {
    if (vec4 && p.nc >= WIDE_NC) launch_infer_vec<4, T, true>(p, dim3((total / 4 + 127) / 128, b), y, out_xyxy, out_px, f, st);
    else if (vec4) launch_infer_vec<4, T, false>(p, dim3((total / 4 + 127) / 128, b), y, out_xyxy, out_px, f, st);
    else launch_infer_vec<1, T, false>(p, dim3((total + 127) / 128, b), y, out_xyxy, out_px, f, st);
}

static int decode_maps(const void* const* h_level_ptrs, int in_dtype, const kzd_level* h_levels, int n_levels, int b,
                       int nc, float* y, float* out_xyxy, float* out_px, const FilterArgs* f, void* stream,
                       const char* who);

extern "C" __attribute__((visibility("default"))) int kzd_decode_infer(const void* const* h_level_ptrs, int in_dtype, const kzd_level* h_levels,
                                int n_levels, int b, int nc, float* y, void* stream) {
    return decode_maps(h_level_ptrs, in_dtype, h_levels, n_levels, b, nc, y, nullptr, nullptr, nullptr, stream,
                       "kzd_decode_infer");
}

extern "C" __attribute__((visibility("default"))) int kzd_decode_dual(const void* const* h_level_ptrs, int in_dtype, const kzd_level* h_levels,
                               int n_levels, int b, int nc, float* y, float* out_xyxy, float* out_xyxy_px,
                               void* stream) {
    if (!out_xyxy) {
        set_error("kzd_decode_dual: null out_xyxy");
        return KZD_EINVAL;
    }
    return decode_maps(h_level_ptrs, in_dtype, h_levels, n_levels, b, nc, y, out_xyxy, out_xyxy_px, nullptr, stream,
                       "kzd_decode_dual");
}

extern "C" __attribute__((visibility("default"))) int kzd_decode_filter(const void* const* h_level_ptrs, int in_dtype, const kzd_level* h_levels,
                                 int n_levels, int b, int nc, float* y, float* out_xyxy, float* out_xyxy_px,
                                 float conf_thres, uint64_t* cand_keys, float* cand_rows, int32_t* cand_counts,
                                 int cand_cap, void* stream) {
    KZD_CHECK_ARG(cand_keys && cand_rows && cand_counts, "kzd_decode_filter: null candidate pool");
    KZD_CHECK_ARG(nc >= 1 && cand_cap >= 1, "kzd_decode_filter: bad sizes nc=%d cap=%d", nc, cand_cap);
    KZD_CHECK_ARG(out_xyxy || !out_xyxy_px, "kzd_decode_filter: out_xyxy_px without out_xyxy");
    if (b <= 0) return b == 0 ? 0 : KZD_EINVAL;
    KZD_CUDA(cudaMemsetAsync(cand_counts, 0, (size_t)b * sizeof(int32_t), (cudaStream_t)stream));
    const FilterArgs f{conf_thres, (u64*)cand_keys, cand_rows, cand_counts, cand_cap};
    return decode_maps(h_level_ptrs, in_dtype, h_levels, n_levels, b, nc, y, out_xyxy, out_xyxy_px, &f, stream,
                       "kzd_decode_filter");
}

static int decode_maps(const void* const* h_level_ptrs, int in_dtype, const kzd_level* h_levels, int n_levels, int b,
                       int nc, float* y, float* out_xyxy, float* out_px, const FilterArgs* f, void* stream,
                       const char* who) {
    KZD_CHECK_ARG(h_level_ptrs && h_levels && y, "%s: null pointer", who);
    KZD_CHECK_ARG(n_levels >= 1 && n_levels <= KZD_MAX_LEVELS, "%s: n_levels=%d not in [1,%d]", who, n_levels,
                  KZD_MAX_LEVELS);
    KZD_CHECK_ARG(b >= 0 && nc >= 0, "%s: negative size", who);
    KZD_CHECK_ARG(in_dtype >= 0 && in_dtype <= 2, "%s: in_dtype=%d not in {0 fp32, 1 fp16, 2 bf16}", who, in_dtype);
    KZD_CHECK_ARG((!out_xyxy || ((uintptr_t)out_xyxy) % 16 == 0) && (!out_px || ((uintptr_t)out_px) % 16 == 0),
                  "%s: box outputs must be 16-byte aligned", who);
    InferParams p;
    int off = 0;
    bool vec4 = true;
    const uintptr_t in_align = in_dtype == 0 ? 16 : 8;  // 4 elements per load
    for (int l = 0; l < n_levels; ++l) {
        KZD_CHECK_ARG(h_levels[l].h > 0 && h_levels[l].w > 0, "%s: empty level %d", who, l);
        KZD_CHECK_ARG(h_levels[l].offset == off, "%s: level %d offset %d, expected %d", who, l, h_levels[l].offset, off);
        p.ptr[l] = h_level_ptrs[l];
        p.hw[l] = h_levels[l].h * h_levels[l].w;
        p.w[l] = h_levels[l].w;
        p.off[l] = off;
        p.stride[l] = h_levels[l].stride;
        off += p.hw[l];
        vec4 = vec4 && (p.hw[l] % 4 == 0) && (((uintptr_t)p.ptr[l]) % in_align == 0);
    }
    for (int l = n_levels; l <= KZD_MAX_LEVELS; ++l) p.off[l] = off;
    for (int l = n_levels; l < KZD_MAX_LEVELS; ++l) p.ptr[l] = nullptr, p.hw[l] = 1, p.w[l] = 1, p.stride[l] = 0.f;
    p.n_levels = n_levels, p.b = b, p.nc = nc, p.A = off;
    vec4 = vec4 && (((uintptr_t)y) % 16 == 0);
    if (b == 0) return 0;
    KZD_CHECK_ARG(b <= 65535, "%s: b=%d > 65535", who, b);
    cudaStream_t st = (cudaStream_t)stream;
    if (in_dtype == 1) launch_infer<__half>(p, off, b, vec4, y, out_xyxy, out_px, f, st);
    else if (in_dtype == 2) launch_infer<__nv_bfloat16>(p, off, b, vec4, y, out_xyxy, out_px, f, st);
    else launch_infer<float>(p, off, b, vec4, y, out_xyxy, out_px, f, st);
    KZD_LAUNCH_CHECK("decode_infer_kernel");
    return 0;
}

static int train_grid(long long n_tiles, const void* fn, int threads, int smem) {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem);
    if (per_sm < 1) per_sm = 1;
    long long g = (long long)sm_count() * per_sm;
    return (int)(n_tiles < g ? n_tiles : g);
}

template <int ANCH, int STAGES, typename IDX, typename T>
static int launch_train(const void* pred_dist, const float* anchor_points, const float* stride_tensor, int A,
                        long long N, float* out_xyxy, float* out_xyxy_px, cudaStream_t st) {
    constexpr int SMEM = STAGES * ANCH * 64 * (int)sizeof(T);
    auto fn = decode_train_kernel<ANCH, STAGES, IDX, T>;
    KZD_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    const long long n_tiles = (N + ANCH - 1) / ANCH;
    const int grid = train_grid(n_tiles, (const void*)fn, ANCH * 4, SMEM);
    fn<<<grid, ANCH * 4, SMEM, st>>>(reinterpret_cast<const T*>(pred_dist), anchor_points, stride_tensor, A, (IDX)N,
                                     out_xyxy, out_xyxy_px);
    KZD_LAUNCH_CHECK("decode_train_kernel");
    return 0;
}

template <typename T>
static int launch_train_idx(const void* pred_dist, const float* anchor_points, const float* stride_tensor, int A,
                            long long N, float* out_xyxy, float* out_xyxy_px, cudaStream_t st) {
    // 64-anchor tiles, 3 stages (4 CTAs/SM for fp32): tile/stage variants (32..128 anchors, 2..6 stages) were
    // swept on B200 and are all within 1 % of each other - the kernel is HBM-bound.
    if (N + 4096 >= 0x7fffffffLL)
        return launch_train<64, 3, long long, T>(pred_dist, anchor_points, stride_tensor, A, N, out_xyxy, out_xyxy_px, st);
    return launch_train<64, 3, int, T>(pred_dist, anchor_points, stride_tensor, A, N, out_xyxy, out_xyxy_px, st);
}

extern "C" __attribute__((visibility("default"))) int kzd_decode_train(const void* pred_dist, int in_dtype, const float* anchor_points,
                                const float* stride_tensor, int b, int A, float* out_xyxy, float* out_xyxy_px,
                                void* stream) {
    KZD_CHECK_ARG(pred_dist && anchor_points && out_xyxy, "kzd_decode_train: null pointer");
    KZD_CHECK_ARG(b >= 0 && A >= 0, "kzd_decode_train: negative size");
    KZD_CHECK_ARG(in_dtype >= 0 && in_dtype <= 2, "kzd_decode_train: in_dtype=%d not in {0 fp32, 1 fp16, 2 bf16}", in_dtype);
    KZD_CHECK_ARG(((uintptr_t)pred_dist) % 16 == 0, "kzd_decode_train: pred_dist must be 16-byte aligned");
    KZD_CHECK_ARG(!out_xyxy_px || stride_tensor, "kzd_decode_train: out_xyxy_px needs stride_tensor");
    const long long N = (long long)b * A;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (in_dtype == 1)
        return launch_train_idx<__half>(pred_dist, anchor_points, stride_tensor, A, N, out_xyxy, out_xyxy_px, st);
    if (in_dtype == 2)
        return launch_train_idx<__nv_bfloat16>(pred_dist, anchor_points, stride_tensor, A, N, out_xyxy, out_xyxy_px, st);
    return launch_train_idx<float>(pred_dist, anchor_points, stride_tensor, A, N, out_xyxy, out_xyxy_px, st);
}

template <typename T>
static int launch_train_bwd(const void* pred_dist, const float* grad_xyxy, long long N, void* grad_pred_dist, cudaStream_t st) {
    auto fn = decode_train_bwd_kernel<T>;
    KZD_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, TR_SMEM));
    const long long n_tiles = (N + TR_ANCH - 1) / TR_ANCH;
    const int grid = train_grid(n_tiles, (const void*)fn, TR_THREADS, TR_SMEM);
    fn<<<grid, TR_THREADS, TR_SMEM, st>>>(reinterpret_cast<const T*>(pred_dist), grad_xyxy, N, reinterpret_cast<T*>(grad_pred_dist));
    KZD_LAUNCH_CHECK("decode_train_bwd_kernel");
    return 0;
}

extern "C" __attribute__((visibility("default"))) int kzd_decode_train_bwd(const void* pred_dist, int in_dtype, const float* grad_xyxy, int b,
                                    int A, void* grad_pred_dist, void* stream) {
    KZD_CHECK_ARG(pred_dist && grad_xyxy && grad_pred_dist, "kzd_decode_train_bwd: null pointer");
    KZD_CHECK_ARG(b >= 0 && A >= 0, "kzd_decode_train_bwd: negative size");
    KZD_CHECK_ARG(in_dtype >= 0 && in_dtype <= 2, "kzd_decode_train_bwd: in_dtype=%d not in {0 fp32, 1 fp16, 2 bf16}", in_dtype);
    KZD_CHECK_ARG(((uintptr_t)pred_dist) % 16 == 0 && ((uintptr_t)grad_pred_dist) % 16 == 0,
                  "kzd_decode_train_bwd: tensors must be 16-byte aligned");
    const long long N = (long long)b * A;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (in_dtype == 1) return launch_train_bwd<__half>(pred_dist, grad_xyxy, N, grad_pred_dist, st);
    if (in_dtype == 2) return launch_train_bwd<__nv_bfloat16>(pred_dist, grad_xyxy, N, grad_pred_dist, st);
    return launch_train_bwd<float>(pred_dist, grad_xyxy, N, grad_pred_dist, st);
}
