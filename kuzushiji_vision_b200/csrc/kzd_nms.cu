// Batched NMS (sm_100a): ops.non_max_suppression + torchvision.ops.nms for all images at once.
//
// reference: utils/ops.py:167-316 (filter :234-286, class offset :289-295, nms call :296, [:max_det]
// :297, row gather :311) and torchvision's greedy NMS (pinned 0.21.0 in the reference's uv.lock;
// semantics restated in-tree at examples/YOLOv8-LibTorch-CPP-Inference/main.cc:80-136).
//
// Pipeline (no host synchronisation inside; candidate counts live on the device):
//   nms_filter_kernel  conf filter + best class / multi-label + xywh->xyxy, candidates appended to
//                      a per-image pool with a 64-bit key  score_bits<<32 | ~order
//   nms_order_bucket_kernel   the usual case (suppression capacity <= 4096): one CTA per image ranks the
//                      candidates by score buckets in shared memory (histogram + prefix + in-bucket compares)
//   nms_chunk_sort / nms_rank   larger capacities: 512-key chunk sorts (registers + warp shuffles + shared memory) + rank by
//                      binary search = torchvision's stable descending score sort; writes score-ordered,
//                      class-offset boxes. nms_order_cluster_kernel does both in one launch (one 8-CTA
//                      thread-block cluster per image, rank through distributed shared memory) when the
//                      candidate pool is at most 4096 per image
//   nms_mask_kernel    persistent CTAs over 64x64 tiles of the upper triangle: bit (i,j) = IoU > thr
//                      (4-compare superset test for all pairs, exact IoU only for intersecting ones)
//   nms_scan_kernel    per image: greedy walk in 64-row chunks; the diagonal tile is resolved by one
//                      warp with ballot/shuffle, helper warps OR the kept rows into later chunks with
//                      sparse shared-memory atomics; the first max_det kept rows are gathered in score order
//   mask + scan run in two phases of column ranges: the walk stops once max_det boxes are kept, so the
//   second phase is needed only when some image's walk is not over after the first; finished images skip it
//   on the device (kzd_nms_suppress)
#include <cooperative_groups.h>

#include "kzd_common.cuh"

namespace cg = cooperative_groups;

namespace kzd {

// ------------------------------------------------------------------------------------------
// stage 1: filter
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ bool class_allowed(int c, const int32_t* __restrict__ classes, int n_classes) {
    if (!classes) return true;
    for (int k = 0; k < n_classes; ++k)
        if (classes[k] == c) return true;
    return false;
}

// Best-class path, VEC consecutive anchors per thread (128-bit loads/stores when VEC == 4).
template <int VEC>
__global__ void __launch_bounds__(256) nms_filter_kernel(float* pred, int A, int nc, float conf_thres,
                                                         const int32_t* __restrict__ classes, int n_classes,
                                                         int in_place, u64* __restrict__ keys,
                                                         float* __restrict__ rows, int32_t* __restrict__ counts,
                                                         int cap) {
    const int bi = blockIdx.y;
    const int a0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;  // whole warps stay alive for the shuffles
    const bool live = a0 < A;
    float* base = pred + (size_t)bi * (4 + nc) * A;
    float x1[VEC], y1[VEC], x2[VEC], y2[VEC], best[VEC];
    int cls[VEC];
    int want = 0;
    bool is_cand[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) is_cand[i] = false;
    if (live) {
        float cx[VEC], cy[VEC], bw[VEC], bh[VEC];
        if (VEC == 4) {
            const float4 t0 = *reinterpret_cast<const float4*>(base + a0);
            const float4 t1 = *reinterpret_cast<const float4*>(base + (size_t)A + a0);
            const float4 t2 = *reinterpret_cast<const float4*>(base + 2 * (size_t)A + a0);
            const float4 t3 = *reinterpret_cast<const float4*>(base + 3 * (size_t)A + a0);
            cx[0] = t0.x, cx[1 % VEC] = t0.y, cx[2 % VEC] = t0.z, cx[3 % VEC] = t0.w;
            cy[0] = t1.x, cy[1 % VEC] = t1.y, cy[2 % VEC] = t1.z, cy[3 % VEC] = t1.w;
            bw[0] = t2.x, bw[1 % VEC] = t2.y, bw[2 % VEC] = t2.z, bw[3 % VEC] = t2.w;
            bh[0] = t3.x, bh[1 % VEC] = t3.y, bh[2 % VEC] = t3.z, bh[3 % VEC] = t3.w;
        } else {
            cx[0] = base[a0], cy[0] = base[(size_t)A + a0], bw[0] = base[2 * (size_t)A + a0],
            bh[0] = base[3 * (size_t)A + a0];
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            if (in_place == 2) {  // rows already hold xyxy (second pass after a pool overflow)
                x1[i] = cx[i], y1[i] = cy[i], x2[i] = bw[i], y2[i] = bh[i];
            } else {              // utils/ops.py:429-432
                const float hw = bw[i] * 0.5f, hh = bh[i] * 0.5f;
                x1[i] = cx[i] - hw, y1[i] = cy[i] - hh, x2[i] = cx[i] + hw, y2[i] = cy[i] + hh;
            }
        }
        if (in_place == 1) {
            if (VEC == 4) {
                *reinterpret_cast<float4*>(base + a0) = make_float4(x1[0], x1[1 % VEC], x1[2 % VEC], x1[3 % VEC]);
                *reinterpret_cast<float4*>(base + (size_t)A + a0) =
                    make_float4(y1[0], y1[1 % VEC], y1[2 % VEC], y1[3 % VEC]);
                *reinterpret_cast<float4*>(base + 2 * (size_t)A + a0) =
                    make_float4(x2[0], x2[1 % VEC], x2[2 % VEC], x2[3 % VEC]);
                *reinterpret_cast<float4*>(base + 3 * (size_t)A + a0) =
                    make_float4(y2[0], y2[1 % VEC], y2[2 % VEC], y2[3 % VEC]);
            } else {
                base[a0] = x1[0], base[(size_t)A + a0] = y1[0], base[2 * (size_t)A + a0] = x2[0],
                base[3 * (size_t)A + a0] = y2[0];
            }
        }
        const float* csrc = base + 4 * (size_t)A + a0;
#pragma unroll
        for (int i = 0; i < VEC; ++i) best[i] = -INFINITY, cls[i] = 0;
        for (int c = 0; c < nc; ++c) {
            float s[VEC];
            if (VEC == 4) {
                const float4 t = *reinterpret_cast<const float4*>(csrc + (size_t)c * A);
                s[0] = t.x, s[1 % VEC] = t.y, s[2 % VEC] = t.z, s[3 % VEC] = t.w;
            } else {
                s[0] = csrc[(size_t)c * A];
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i)
                if (c == 0 || s[i] > best[i]) best[i] = s[i], cls[i] = c;  // first maximum wins (utils/ops.py:274)
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            is_cand[i] = best[i] > conf_thres && class_allowed(cls[i], classes, n_classes);
            want += is_cand[i] ? 1 : 0;
        }
    }
    int slot = warp_alloc(want, counts + bi);
    if (want) {
        u64* k_img = keys + (size_t)bi * cap;
        float* r_img = rows + (size_t)bi * cap * 6;
#pragma unroll
        for (int i = 0; i < VEC; ++i)
            if (is_cand[i]) {
                write_candidate(slot, cap, x1[i], y1[i], x2[i], y2[i], best[i], cls[i], (uint32_t)(a0 + i), k_img,
                                r_img);
                ++slot;
            }
    }
}

// Multi-label path (utils/ops.py:271-272): one row per (anchor, class) above conf; order = anchor*nc+class.
__global__ void __launch_bounds__(256) nms_filter_multilabel_kernel(float* pred, int A, int nc, float conf_thres,
                                                                    const int32_t* __restrict__ classes,
                                                                    int n_classes, int in_place,
                                                                    u64* __restrict__ keys, float* __restrict__ rows,
                                                                    int32_t* __restrict__ counts, int cap) {
    const int bi = blockIdx.y;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = a < A;
    float* base = pred + (size_t)bi * (4 + nc) * A;
    float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f;
    if (live) {
        const float cx = base[a], cy = base[(size_t)A + a];
        const float bw = base[2 * (size_t)A + a], bh = base[3 * (size_t)A + a];
        if (in_place == 2) {
            x1 = cx, y1 = cy, x2 = bw, y2 = bh;
        } else {
            const float hw = bw * 0.5f, hh = bh * 0.5f;
            x1 = cx - hw, y1 = cy - hh, x2 = cx + hw, y2 = cy + hh;
            if (in_place == 1)
                base[a] = x1, base[(size_t)A + a] = y1, base[2 * (size_t)A + a] = x2, base[3 * (size_t)A + a] = y2;
        }
    }
    const float* cls = base + 4 * (size_t)A;
    u64* k_img = keys + (size_t)bi * cap;
    float* r_img = rows + (size_t)bi * cap * 6;
    for (int c = 0; c < nc; ++c) {
        const float s = live ? cls[(size_t)c * A + a] : 0.f;
        const bool is_cand = live && s > conf_thres && class_allowed(c, classes, n_classes);
        const int slot = warp_alloc(is_cand ? 1 : 0, counts + bi);
        if (is_cand)
            write_candidate(slot, cap, x1, y1, x2, y2, s, c, (uint32_t)a * (uint32_t)nc + (uint32_t)c, k_img, r_img);
    }
}

// Apriori labels (autolabelling, utils/ops.py:256-261): every label row [cls, cx, cy, w, h] (pixels) becomes a
// candidate with confidence 1.0 for its class, appended after the image's predictions (order = order_base + k:
// the reference concatenates them below the prediction rows, so they come last among equal scores).
__global__ void __launch_bounds__(256) nms_append_labels_kernel(const float* __restrict__ labels,
                                                                const int32_t* __restrict__ label_counts, int L, int nc,
                                                                float conf_thres, const int32_t* __restrict__ classes,
                                                                int n_classes, uint32_t order_base, u64* __restrict__ keys,
                                                                float* __restrict__ rows, int32_t* __restrict__ counts,
                                                                int cap) {
    const int bi = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;  // whole warps stay alive for the shuffles
    bool is_cand = false;
    float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f;
    int cls = 0;
    if (k < L && k < label_counts[bi]) {
        const float* lb = labels + ((size_t)bi * L + k) * 5;
        cls = (int)(long long)lb[0];
        const float hw = lb[3] * 0.5f, hh = lb[4] * 0.5f;  // utils/ops.py:429-432
        x1 = lb[1] - hw, y1 = lb[2] - hh, x2 = lb[1] + hw, y2 = lb[2] + hh;
        is_cand = cls >= 0 && cls < nc && 1.0f > conf_thres && class_allowed(cls, classes, n_classes);
    }
    const int slot = warp_alloc(is_cand ? 1 : 0, counts + bi);
    if (is_cand)
        write_candidate(slot, cap, x1, y1, x2, y2, 1.0f, cls, order_base + (uint32_t)k, keys + (size_t)bi * cap,
                        rows + (size_t)bi * cap * 6);
}

// ------------------------------------------------------------------------------------------
// stage 2a: ordering.  chunk sort (512 keys per CTA, all SMs busy) + rank by binary search
// ------------------------------------------------------------------------------------------

constexpr int CHUNK = 512;
constexpr int CHUNK_THREADS = 256;

struct SuppressWs {
    float4* sbox;       // (b, n_cap)   score-ordered boxes, class offset applied
    uint32_t* sslot;    // (b, n_cap)   pool slot of each ordered box
    int32_t* n_sorted;  // (b)          boxes entering suppression, or -1 on overflow
    u64* mask;          // (b, T_cap*T_cap*64)
    u64* ckeys;         // (b, C_cap*CHUNK)  keys, sorted descending inside each chunk (0 = padding)
    uint32_t* cvals;    // (b, C_cap*CHUNK)  pool slot of each key
    int32_t* state;     // (b, 4)       scan state carried between phases: done, kept, chunks walked
    u64* keepmask;      // (b, T_cap)   keep mask of every walked chunk
    int32_t* kept_before;  // (b, T_cap)
    int32_t* pending;      // [1] images whose walk is not over after the first phase (0: later phases return at once)
    int T_cap, C_cap;
};

// Bitonic sort of one 512-key chunk by 256 threads, descending (keys are unique apart from the zero
// padding). Thread t holds elements t and t + 256 in registers: compare-exchange distance 256 is inside the
// thread, distances 16..1 go through warp shuffles, only distances 128, 64, 32 need shared memory and a CTA
// barrier (9 of the 45 stages). On return the sorted chunk is also in keys[] / vals[].
__device__ __forceinline__ void order_cx_shfl(u64& key, uint32_t& val, int i, int k, int j) {
    const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)key, j);
    const uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(key >> 32), j);
    const uint32_t pv = __shfl_xor_sync(0xffffffffu, val, j);
    const u64 pk = ((u64)hi << 32) | lo;
    const bool want_max = (((i & k) == 0) == ((i & j) == 0));  // descending overall: the lower index keeps the larger key
    if (want_max ? (pk > key) : (pk < key)) key = pk, val = pv;
}

__device__ __forceinline__ void chunk_sort_desc(u64& k0, uint32_t& v0, u64& k1, uint32_t& v1, u64* keys, uint32_t* vals) {
    const int t = threadIdx.x, i0 = t, i1 = t + CHUNK_THREADS;
    for (int k = 2; k <= CHUNK; k <<= 1) {
        int j = k >> 1;
        if (j == CHUNK_THREADS) {  // distance 256: both elements live in this thread (only k == 512: descending)
            if (k0 < k1) {
                const u64 tk = k0; k0 = k1, k1 = tk;
                const uint32_t tv = v0; v0 = v1, v1 = tv;
            }
            j >>= 1;
        }
        if (j >= 32) {  // distances 128, 64, 32 through shared memory
            __syncthreads();
            keys[i0] = k0, keys[i1] = k1, vals[i0] = v0, vals[i1] = v1;
            for (; j >= 32; j >>= 1) {
                __syncthreads();
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), q = i | j;
                const u64 ki = keys[i], kq = keys[q];
                if ((ki < kq) == ((i & k) == 0)) {
                    keys[i] = kq, keys[q] = ki;
                    const uint32_t vi = vals[i];
                    vals[i] = vals[q], vals[q] = vi;
                }
            }
            __syncthreads();
            k0 = keys[i0], k1 = keys[i1], v0 = vals[i0], v1 = vals[i1];
        }
        for (; j > 0; j >>= 1) {  // distances 16..1 inside the warp
            order_cx_shfl(k0, v0, i0, k, j);
            order_cx_shfl(k1, v1, i1, k, j);
        }
    }
    __syncthreads();
    keys[i0] = k0, keys[i1] = k1, vals[i0] = v0, vals[i1] = v1;
}

// CTA (chunk, image): sort of the chunk's 512 keys, descending; padding keys are 0.
__global__ void __launch_bounds__(CHUNK_THREADS) nms_chunk_sort_kernel(const u64* __restrict__ cand_keys,
                                                                       const int32_t* __restrict__ cand_counts,
                                                                       int cand_cap, int n_cap, int max_nms,
                                                                       SuppressWs ws) {
    __shared__ u64 keys[CHUNK];
    __shared__ uint32_t vals[CHUNK];
    const int bi = blockIdx.y, chunk = blockIdx.x;
    const int count = cand_counts[bi];
    const int n = count < max_nms ? count : max_nms;
    const bool overflow = count > cand_cap || n > n_cap;  // pool or suppression stage too small
    if (chunk == 0 && threadIdx.x == 0) ws.n_sorted[bi] = overflow ? -1 : n;
    if (chunk == 0 && bi == 0 && threadIdx.x == 0) *ws.pending = 0;
    if (overflow || chunk * CHUNK >= count) return;
    const u64* __restrict__ src = cand_keys + (size_t)bi * cand_cap;
    const int t = threadIdx.x;
    uint32_t v0 = (uint32_t)(chunk * CHUNK + t), v1 = (uint32_t)(chunk * CHUNK + t + CHUNK_THREADS);
    u64 k0 = (int)v0 < count ? src[v0] : 0ull, k1 = (int)v1 < count ? src[v1] : 0ull;
    chunk_sort_desc(k0, v0, k1, v1, keys, vals);
    u64* __restrict__ dk = ws.ckeys + ((size_t)bi * ws.C_cap + chunk) * CHUNK;
    uint32_t* __restrict__ dv = ws.cvals + ((size_t)bi * ws.C_cap + chunk) * CHUNK;
    dk[t] = k0, dk[t + CHUNK_THREADS] = k1, dv[t] = v0, dv[t + CHUNK_THREADS] = v1;
}

// CTA (chunk, image): global rank of each of the chunk's keys = position inside its own chunk + number of
// larger keys in every other chunk (binary search; keys are unique; one 4 KB staging buffer - staging all
// other chunks at once, 28 KB per CTA, was 2 us faster alone and 20 us per step slower with several
// pipelines in flight: the fat CTAs no longer fit next to the assigner's). The rank is the box's position in
// torchvision's stable descending score order; the first n boxes are emitted with their class offset.
__global__ void __launch_bounds__(CHUNK_THREADS) nms_rank_kernel(const float* __restrict__ cand_rows,
                                                                 const int32_t* __restrict__ cand_counts,
                                                                 int cand_cap, int n_cap, int agnostic, float max_wh,
                                                                 SuppressWs ws) {
    __shared__ u64 other[CHUNK];
    constexpr int STAGE = 7;  // other chunks whose keys are in flight / in registers together
    const int bi = blockIdx.y, chunk = blockIdx.x;
    const int n = ws.n_sorted[bi];
    const int count = cand_counts[bi];
    if (n <= 0 || chunk * CHUNK >= count) return;
    const int n_chunks = (count + CHUNK - 1) / CHUNK;
    const u64* __restrict__ ck = ws.ckeys + (size_t)bi * ws.C_cap * CHUNK;
    const int t0 = threadIdx.x, t1 = threadIdx.x + CHUNK_THREADS;
    const u64 k0 = ck[(size_t)chunk * CHUNK + t0], k1 = ck[(size_t)chunk * CHUNK + t1];
    // own slot / row loads do not depend on the rank: requested now, used at the end
    const uint32_t* __restrict__ cv = ws.cvals + ((size_t)bi * ws.C_cap + chunk) * CHUNK;
    const uint32_t slot0 = cv[t0], slot1 = cv[t1];
    int r0 = t0, r1 = t1;
    for (int base = 0; base < n_chunks; base += STAGE) {
        u64 st0[STAGE], st1[STAGE];
#pragma unroll
        for (int k = 0; k < STAGE; ++k) {  // all loads of the group in flight at once (one memory round trip)
            const int c = base + k;
            st0[k] = st1[k] = 0ull;
            if (c < n_chunks && c != chunk) st0[k] = ck[(size_t)c * CHUNK + t0], st1[k] = ck[(size_t)c * CHUNK + t1];
        }
#pragma unroll
        for (int k = 0; k < STAGE; ++k) {
            const int c = base + k;
            if (c >= n_chunks || c == chunk) continue;  // uniform
            __syncthreads();
            other[t0] = st0[k], other[t1] = st1[k];
            __syncthreads();
            int lo0 = 0, hi0 = CHUNK, lo1 = 0, hi1 = CHUNK;
#pragma unroll
            for (int s = 0; s < 10; ++s) {  // 2^9 = 512: 10 halvings always converge
                const int m0 = (lo0 + hi0) >> 1, m1 = (lo1 + hi1) >> 1;
                const bool g0 = lo0 < hi0 && other[m0 < CHUNK ? m0 : CHUNK - 1] > k0;
                const bool g1 = lo1 < hi1 && other[m1 < CHUNK ? m1 : CHUNK - 1] > k1;
                if (lo0 < hi0) { if (g0) lo0 = m0 + 1; else hi0 = m0; }
                if (lo1 < hi1) { if (g1) lo1 = m1 + 1; else hi1 = m1; }
            }
            r0 += lo0, r1 += lo1;
        }
    }
    const float* __restrict__ rows = cand_rows + (size_t)bi * cand_cap * 6;
    float4* __restrict__ sbox = ws.sbox + (size_t)bi * n_cap;
    uint32_t* __restrict__ sslot = ws.sslot + (size_t)bi * n_cap;
    auto emit = [&](u64 key, int rank, uint32_t slot) {
        if (key == 0ull || rank >= n) return;
        const float2* r = reinterpret_cast<const float2*>(rows + (size_t)slot * 6);
        const float2 p0 = r[0], p1 = r[1], p2 = r[2];
        const float off = agnostic ? 0.f : __fmul_rn(p2.y, max_wh);                      // utils/ops.py:289
        sbox[rank] = make_float4(__fadd_rn(p0.x, off), __fadd_rn(p0.y, off), __fadd_rn(p1.x, off),
                                 __fadd_rn(p1.y, off));                                   // utils/ops.py:295
        sslot[rank] = slot;
    };
    emit(k0, r0, slot0);
    emit(k1, r1, slot1);
}

// Ordering in ONE launch for the usual suppression capacity (n_cap <= 4096): one CTA per image ranks the
// image's candidates by bucketing instead of sorting. Scores are floats in a narrow range, so their bit
// patterns - rebased to the image's smallest score and shifted to 11 bits - spread the candidates over 2048
// buckets with one or two members each: a shared-memory histogram + one descending prefix sum give every key
// the number of keys in higher buckets, the keys of the buckets that reach into the first n ranks are
// scattered bucket by bucket into shared memory, and a key's exact rank is its bucket's start plus the number
// of larger keys inside its bucket (a handful of compares). No sort stages, no second kernel, no global
// round trip for the keys: the chunk sort + rank pair takes 22 us on the bench workload, this takes a few.
// Equal scores (ties) share a bucket and cost compares quadratic in the tie count - still exact.
constexpr int BR_THREADS = 512;
constexpr int BR_BUCKETS = 2048;
constexpr int BR_KEYS = 4096 + 512;  // keys the shared-memory stage holds: n_cap plus the overhang of the last bucket
constexpr int BR_REG = 8;            // keys per thread kept in registers between the passes

__global__ void __launch_bounds__(BR_THREADS) nms_order_bucket_kernel(const u64* __restrict__ cand_keys,
                                                                      const float* __restrict__ cand_rows,
                                                                      const int32_t* __restrict__ cand_counts,
                                                                      int cand_cap, int n_cap, int max_nms, int agnostic,
                                                                      float max_wh, SuppressWs ws) {
    extern __shared__ unsigned char br_smem[];
    u64* skey = reinterpret_cast<u64*>(br_smem);                         // BR_KEYS
    uint32_t* fill = reinterpret_cast<uint32_t*>(skey + BR_KEYS);         // BR_BUCKETS: members scattered so far
    uint16_t* hist = reinterpret_cast<uint16_t*>(fill + BR_BUCKETS);      // BR_BUCKETS: members
    uint16_t* start = hist + BR_BUCKETS;                                  // BR_BUCKETS: keys in higher buckets
    uint16_t* sval = start + BR_BUCKETS;                                  // BR_KEYS: pool slot of the key
    __shared__ uint32_t s_min, s_max, s_warp[BR_THREADS / 32];
    __shared__ int s_stored, s_over;
    const int bi = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int count = cand_counts[bi];
    const int n = count < max_nms ? count : max_nms;
    const bool overflow = count > cand_cap || n > n_cap;  // pool or suppression stage too small
    if (t == 0) ws.n_sorted[bi] = overflow ? -1 : n;
    if (bi == 0 && t == 0) *ws.pending = 0;
    if (overflow || count <= 0) return;
    const u64* __restrict__ src = cand_keys + (size_t)bi * cand_cap;
    u64 kreg[BR_REG];
#pragma unroll
    for (int i = 0; i < BR_REG; ++i) {
        const int q = i * BR_THREADS + t;
        kreg[i] = q < count ? src[q] : 0ull;
    }
    auto for_each_key = [&](auto&& f) {  // f(key, pool slot): the first BR_REG keys of a thread live in registers
#pragma unroll
        for (int i = 0; i < BR_REG; ++i) {
            const int q = i * BR_THREADS + t;
            if (q < count) f(kreg[i], q);
        }
        for (int q = BR_REG * BR_THREADS + t; q < count; q += BR_THREADS) f(src[q], q);
    };
    for (int b = t; b < BR_BUCKETS; b += BR_THREADS) hist[b] = 0, fill[b] = 0u;
    if (t == 0) s_min = 0xffffffffu, s_max = 0u, s_stored = 0, s_over = 0;
    __syncthreads();
    // pass 0: score range of the image
    uint32_t lo = 0xffffffffu, hi = 0u;
    for_each_key([&](u64 key, int) {
        const uint32_t sb = (uint32_t)(key >> 32);
        lo = min(lo, sb), hi = max(hi, sb);
    });
    lo = __reduce_min_sync(0xffffffffu, lo), hi = __reduce_max_sync(0xffffffffu, hi);
    if (lane == 0) atomicMin(&s_min, lo), atomicMax(&s_max, hi);
    __syncthreads();
    const uint32_t base = s_min, range = s_max - s_min;
    const int shift = range < (uint32_t)BR_BUCKETS ? 0 : (32 - __clz(range)) - 11;  // (score - base) >> shift < 2048
    // pass 1: histogram (bucket 0 = lowest scores)
    for_each_key([&](u64 key, int) {
        const uint32_t b = ((uint32_t)(key >> 32) - base) >> shift;
        atomicAdd(reinterpret_cast<unsigned int*>(fill) + b, 1u);  // 32-bit counters first, narrowed below
    });
    __syncthreads();
    // descending exclusive prefix: start[b] = number of keys in buckets above b. Thread t owns the four buckets
    // 2047-4t .. 2044-4t, in that (descending) order.
    uint32_t c4[4], sum = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c4[k] = fill[BR_BUCKETS - 1 - (4 * t + k)];
        sum += c4[k];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < BR_THREADS / 32 ? s_warp[lane] : 0u, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += v;
        }
        if (lane < BR_THREADS / 32) s_warp[lane] = wi - w;  // exclusive
    }
    __syncthreads();
    uint32_t run = s_warp[warp] + incl - sum;  // keys above this thread's first bucket
    int stored_hi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int b = BR_BUCKETS - 1 - (4 * t + k);
        start[b] = (uint16_t)min(run, 65535u);
        hist[b] = (uint16_t)min(c4[k], 65535u);
        if (run < (uint32_t)n && c4[k]) stored_hi = (int)min(run + c4[k], 0x7fffffffu);  // this bucket reaches into the first n ranks
        run += c4[k];
    }
    __syncthreads();  // everyone has read fill[] as counters
#pragma unroll
    for (int k = 0; k < 4; ++k) fill[BR_BUCKETS - 1 - (4 * t + k)] = 0u;
    stored_hi = __reduce_max_sync(0xffffffffu, stored_hi);
    if (lane == 0 && stored_hi) atomicMax(&s_stored, stored_hi);
    __syncthreads();
    // pass 2: scatter the keys of the buckets that reach into the first n ranks, bucket by bucket
    for_each_key([&](u64 key, int q) {
        const uint32_t b = ((uint32_t)(key >> 32) - base) >> shift;
        const uint32_t st = start[b];
        if (st < (uint32_t)n) {
            const uint32_t pos = st + atomicAdd(reinterpret_cast<unsigned int*>(fill) + b, 1u);
            if (pos < (uint32_t)BR_KEYS) skey[pos] = key, sval[pos] = (uint16_t)q;
            else s_over = 1;  // a tie group larger than the shared-memory overhang
        }
    });
    __syncthreads();
    if (s_over || count > 65535) {  // (pool slots are kept as 16-bit values here)
        if (t == 0) ws.n_sorted[bi] = -1;
        return;
    }
    // pass 3: exact rank inside the bucket, emit the first n in score order with their class offset
    const int stored = s_stored;
    const float* __restrict__ rows = cand_rows + (size_t)bi * cand_cap * 6;
    float4* __restrict__ sbox = ws.sbox + (size_t)bi * n_cap;
    uint32_t* __restrict__ sslot = ws.sslot + (size_t)bi * n_cap;
    for (int p = t; p < stored; p += BR_THREADS) {
        const u64 key = skey[p];
        const uint32_t b = ((uint32_t)(key >> 32) - base) >> shift;
        const int s0 = start[b], m = hist[b];
        int rank = s0;
        for (int q = s0; q < s0 + m; ++q) rank += skey[q] > key ? 1 : 0;
        if (rank >= n) continue;
        const uint32_t slot = sval[p];
        const float2* r = reinterpret_cast<const float2*>(rows + (size_t)slot * 6);
        const float2 p0 = r[0], p1 = r[1], p2 = r[2];
        const float off = agnostic ? 0.f : __fmul_rn(p2.y, max_wh);                      // utils/ops.py:289
        sbox[rank] = make_float4(__fadd_rn(p0.x, off), __fadd_rn(p0.y, off), __fadd_rn(p1.x, off),
                                 __fadd_rn(p1.y, off));                                   // utils/ops.py:295
        sslot[rank] = slot;
    }
}

constexpr size_t BR_SMEM = (size_t)BR_KEYS * 8 + (size_t)BR_BUCKETS * 4 + (size_t)BR_BUCKETS * 2 * 2 + (size_t)BR_KEYS * 2;

// One thread-block cluster per image does the whole ordering when the candidate pool is at most 8 chunks
// (the first-try case of the eager drop-in): CTA c sorts chunk c (chunk_sort_desc), the cluster
// synchronises, and every CTA ranks its keys against the other CTAs' sorted chunks read through distributed shared memory - no
// round trip through global memory and one launch instead of two.
constexpr int ORDER_CL = 8;

__global__ void __cluster_dims__(ORDER_CL, 1, 1) __launch_bounds__(CHUNK_THREADS)
    nms_order_cluster_kernel(const u64* __restrict__ cand_keys, const float* __restrict__ cand_rows,
                             const int32_t* __restrict__ cand_counts, int cand_cap, int n_cap, int max_nms, int agnostic,
                             float max_wh, SuppressWs ws) {
    __shared__ u64 keys[CHUNK];
    __shared__ uint32_t vals[CHUNK];
    __shared__ u64 other[ORDER_CL - 1][CHUNK];
    cg::cluster_group cluster = cg::this_cluster();
    const int bi = blockIdx.y, chunk = (int)cluster.block_rank();
    const int t = threadIdx.x;
    const int count = cand_counts[bi];
    const int n = count < max_nms ? count : max_nms;
    const bool overflow = count > cand_cap || n > n_cap;  // pool or suppression stage too small
    if (chunk == 0 && t == 0) ws.n_sorted[bi] = overflow ? -1 : n;
    if (chunk == 0 && bi == 0 && t == 0) *ws.pending = 0;
    if (overflow || count <= 0) return;  // uniform over the cluster
    const bool active = chunk * CHUNK < count;
    const int i0 = t, i1 = t + CHUNK_THREADS;
    u64 k0 = 0ull, k1 = 0ull;
    uint32_t v0 = (uint32_t)(chunk * CHUNK + i0), v1 = (uint32_t)(chunk * CHUNK + i1);
    if (active) {
        const u64* __restrict__ src = cand_keys + (size_t)bi * cand_cap;
        if ((int)v0 < count) k0 = src[v0];
        if ((int)v1 < count) k1 = src[v1];
        chunk_sort_desc(k0, v0, k1, v1, keys, vals);
    }
    cluster.sync();  // every chunk of the image is sorted and visible cluster-wide
    if (active) {
        const int n_chunks = (count + CHUNK - 1) / CHUNK;
        int slot = 0;
        for (int c = 0; c < n_chunks; ++c) {
            if (c == chunk) continue;
            const u64* remote = cluster.map_shared_rank(keys, c);
            other[slot][i0] = remote[i0], other[slot][i1] = remote[i1];
            ++slot;
        }
        __syncthreads();
        int r0 = i0, r1 = i1;
        for (int sidx = 0; sidx < slot; ++sidx) {
            const u64* __restrict__ oc = other[sidx];
            int lo0 = 0, hi0 = CHUNK, lo1 = 0, hi1 = CHUNK;
#pragma unroll
            for (int step = 0; step < 10; ++step) {  // 2^9 = 512: 10 halvings always converge
                const int m0 = (lo0 + hi0) >> 1, m1 = (lo1 + hi1) >> 1;
                const bool g0 = lo0 < hi0 && oc[m0 < CHUNK ? m0 : CHUNK - 1] > k0;
                const bool g1 = lo1 < hi1 && oc[m1 < CHUNK ? m1 : CHUNK - 1] > k1;
                if (lo0 < hi0) { if (g0) lo0 = m0 + 1; else hi0 = m0; }
                if (lo1 < hi1) { if (g1) lo1 = m1 + 1; else hi1 = m1; }
            }
            r0 += lo0, r1 += lo1;
        }
        const float* __restrict__ rows = cand_rows + (size_t)bi * cand_cap * 6;
        float4* __restrict__ sbox = ws.sbox + (size_t)bi * n_cap;
        uint32_t* __restrict__ sslot = ws.sslot + (size_t)bi * n_cap;
        auto emit = [&](u64 key, int rank, uint32_t slot_) {
            if (key == 0ull || rank >= n) return;
            const float2* r = reinterpret_cast<const float2*>(rows + (size_t)slot_ * 6);
            const float2 p0 = r[0], p1 = r[1], p2 = r[2];
            const float off = agnostic ? 0.f : __fmul_rn(p2.y, max_wh);                      // utils/ops.py:289
            sbox[rank] = make_float4(__fadd_rn(p0.x, off), __fadd_rn(p0.y, off), __fadd_rn(p1.x, off),
                                     __fadd_rn(p1.y, off));                                   // utils/ops.py:295
            sslot[rank] = slot_;
        };
        emit(k0, r0, v0);
        emit(k1, r1, v1);
    }
    cluster.sync();  // no CTA leaves while its shared memory may still be read by the others
}

// ------------------------------------------------------------------------------------------
// stage 2b: IoU bitmask
// ------------------------------------------------------------------------------------------

constexpr int MASK_THREADS = 64;

// torchvision's devIoU as compiled for sm_100 in torchvision 0.26.0+cu128 (read from its SASS):
// the row box's area is a rounded product, the column box's area is contracted into an FMA with it.
__device__ __forceinline__ bool iou_gt(const float4 a, const float sa, const float4 b, const float thr) {
    const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
    const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
    const float inter = __fmul_rn(w, h);
    const float u = __fsub_rn(__fmaf_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y), sa), inter);
    // inter / u > thr, decided without the division whenever the quotient is not within a few ulps of thr (the
    // product thr * u carries one rounding, the quotient another: 2^-20 of slack covers both many times over);
    // only the rare borderline pair pays for the IEEE division that fixes torchvision's exact answer
    if (u > 0.f && thr > 0.f) {
        const float tu = __fmul_rn(thr, u);
        if (inter > __fmul_rn(tu, 1.000001f)) return true;
        if (inter < __fmul_rn(tu, 0.999999f)) return false;
    }
    return __fdiv_rn(inter, u) > thr;
}

// Work item = (image, row tile r, group of MASK_CG consecutive column tiles) restricted to the upper
// triangle; items of all images are striped over a persistent grid. Thread i of the CTA owns row
// r*64+i and writes one 64-bit word of the mask per column tile.
// Pass 1 (branch-free, fully unrolled): which of the 64 column boxes can intersect the row box at all
// (4 compares; a superset of the true intersections). Pass 2 (rare): exact IoU > thr only for those.
// Mask layout: [image][r*T_cap + c][i]  (64 consecutive words per tile => coalesced 512 B stores,
// and the scan reads whole tiles).
#ifndef KZD_MASK_CG
#define KZD_MASK_CG 1
#endif
constexpr int MASK_CG = KZD_MASK_CG;

// items of column groups [0, g) for an image of T tiles: group k holds rows 0..min(T, CG*(k+1))-1
__device__ __forceinline__ int items_before(int T, int g) {
    const int Gc = (T + MASK_CG - 1) / MASK_CG;
    if (g <= 0 || T <= 0) return 0;
    if (g >= Gc) return MASK_CG * (Gc - 1) * Gc / 2 + T;  // CG * (1 + ... + (Gc-1)) + T
    return MASK_CG * g * (g + 1) / 2;
}

static long long items_before_host(int T, int g) {  // host twin of items_before
    const int Gc = (T + MASK_CG - 1) / MASK_CG;
    if (g <= 0 || T <= 0) return 0;
    if (g >= Gc) return (long long)MASK_CG * (Gc - 1) * Gc / 2 + T;
    return (long long)MASK_CG * g * (g + 1) / 2;
}

// Phase = column groups [g_lo, g_hi) of every image that is not finished yet (the greedy walk stops as
// soon as max_det boxes are kept, so later columns are often never needed: see kzd_nms_suppress).
__global__ void __launch_bounds__(MASK_THREADS) nms_mask_kernel(int b, int n_cap, float iou_thres, int phase,
                                                                int g_lo, int g_hi, SuppressWs ws) {
    extern __shared__ int pre[];  // (b+1) prefix of item counts
    __shared__ float4 colbox[MASK_CG * 64];
    if (phase > 0 && *ws.pending == 0) return;  // every image's walk ended in the first phase (the usual case)
    for (int i = threadIdx.x; i < b; i += MASK_THREADS) {  // item counts: the loads of all images in flight together
        const int n = ws.n_sorted[i];
        const int done = phase > 0 ? ws.state[i * 4] : 0;
        int items = 0;
        if (n > 0 && !done) {
            const int T = (n + 63) >> 6;
            items = items_before(T, g_hi) - items_before(T, g_lo);
        }
        pre[i + 1] = items;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        pre[0] = 0;
        for (int i = 1; i <= b; ++i) acc += pre[i], pre[i] = acc;
    }
    __syncthreads();
    const int total = pre[b];
    const int i = threadIdx.x;
    int bi = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        while (pre[bi + 1] <= w) ++bi;  // w only grows
        const int n = ws.n_sorted[bi];
        const int T = (n + 63) >> 6;
        const int q = w - pre[bi] + items_before(T, g_lo);
        // items are ordered by column group g, then row r:  offset(g) = CG g (g+1) / 2
        int g = (int)((sqrtf(1.f + (8.f / MASK_CG) * (float)q) - 1.f) * 0.5f);
        g = max(g, 0);
        while (MASK_CG * g * (g + 1) / 2 > q) --g;
        while (MASK_CG * (g + 1) * (g + 2) / 2 <= q) ++g;
        const int r = q - MASK_CG * g * (g + 1) / 2;
        const int c_first = g * MASK_CG;
        const float4* __restrict__ sbox = ws.sbox + (size_t)bi * n_cap;
        const int row = r * 64 + i;
        const float4 a = (row < n) ? sbox[row] : make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();  // previous item's colbox readers are done
#pragma unroll
        for (int k = 0; k < MASK_CG; ++k) {
            const int col = (c_first + k) * 64 + i;
            colbox[k * 64 + i] = (col < n) ? sbox[col] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
#pragma unroll 1
        for (int k = 0; k < MASK_CG; ++k) {
            const int c = c_first + k;
            if (c < r || c >= T) continue;  // below the diagonal / beyond the last tile (uniform)
            const float4* cb_tile = colbox + k * 64;
            uint32_t lo = 0u, hi = 0u;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float4 cb = cb_tile[j];
                if (a.z > cb.x && cb.z > a.x && a.w > cb.y && cb.w > a.y) lo |= (1u << j);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float4 cb = cb_tile[32 + j];
                if (a.z > cb.x && cb.z > a.x && a.w > cb.y && cb.w > a.y) hi |= (1u << j);
            }
            u64 cand = ((u64)hi << 32) | lo;
            const int jn = n - c * 64;  // valid columns of this tile
            if (jn < 64) cand &= (1ull << jn) - 1ull;
            if (r == c) cand &= (i < 63) ? (~0ull << (i + 1)) : 0ull;  // only later boxes can be suppressed
            if (row >= n) cand = 0ull;
            u64 word = 0ull;
            while (cand) {
                const int j = __ffsll((long long)cand) - 1;
                cand &= cand - 1ull;
                if (iou_gt(a, sa, cb_tile[j], iou_thres)) word |= 1ull << j;
            }
            ws.mask[((size_t)bi * ws.T_cap * ws.T_cap + (size_t)r * ws.T_cap + c) * 64 + i] = word;
        }
    }
}

// ------------------------------------------------------------------------------------------
// stage 2c: greedy scan + row gather
// ------------------------------------------------------------------------------------------

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
// Warp 0 walks the diagonal tiles. Warps 4, 8, 12 share its scheduler (warp id mod 4) and only follow the
// barriers: clock64 traces showed the walker - the critical path - waiting hundreds of cycles for issue
// slots behind helper warps that the scheduler runs greedily after every barrier release.
constexpr int SCAN_HELPERS = SCAN_WARPS - SCAN_WARPS / 4;
#ifndef KZD_SCAN_BATCH
#define KZD_SCAN_BATCH 4
#endif
#ifndef KZD_SCAN_DEPTH
#define KZD_SCAN_DEPTH 2
#endif
constexpr int SCAN_BATCH = KZD_SCAN_BATCH;  // column tiles per helper warp held in registers (48 columns ahead; more: from global)
constexpr int SCAN_DEPTH = KZD_SCAN_DEPTH;  // chunks a mask load is issued ahead of its use (2, 4 and 6 measured the same)

__device__ __forceinline__ u64 shfl64(u64 v, int src) {
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
    const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
    return ((u64)hi << 32) | lo;
}

// One CTA per image walks the 64-row chunks in score order. Per chunk the only serial work is the
// diagonal tile (warp 0: ballot of the rows that suppress anything inside the chunk, shuffle their
// words, test/OR); the helper warps then OR the kept rows into the removed words of the later chunks.
// A chunk costs what the slowest warp's instruction stream costs (a lone warp issues one dependent
// instruction every ~4 cycles; clock64 traces showed ~350 instructions per chunk in the first version of
// this kernel), so the two roles run separate, minimal loops: every tile pointer advances by the same
// stride per chunk (row and column both move by one tile), tile validity is one compare against a
// per-warp constant, keep masks and kept counts stay in shared memory until the walk is over, and all
// global loads are issued SCAN_DEPTH chunks ahead into rotating registers (they do not depend on the keep mask).
// Both loops meet at two named barriers per chunk and leave together (the exit tests are CTA-uniform).
__device__ __forceinline__ void scan_bar() { asm volatile("bar.sync 1, %0;" ::"n"(SCAN_THREADS) : "memory"); }

__global__ void __launch_bounds__(SCAN_THREADS) nms_scan_kernel(const float* __restrict__ cand_rows, int cand_cap,
                                                                int n_cap, int max_det, int phase, int t_lo,
                                                                int t_hi, SuppressWs ws,
                                                                float* __restrict__ out_rows,
                                                                int32_t* __restrict__ out_counts) {
    extern __shared__ u64 scan_smem[];
    u64* removed = scan_smem;                                               // T_cap: bits suppressed by a kept row
    u64* keep_s = scan_smem + ws.T_cap;                                     // T_cap: keep mask of every chunk walked here
    int* kb_s = reinterpret_cast<int*>(scan_smem + 2 * (size_t)ws.T_cap);   // T_cap: boxes kept before the chunk
    const int bi = blockIdx.x;
    if (phase > 0 && *ws.pending == 0) return;  // every image's walk ended in the first phase (the usual case)
    const int n = ws.n_sorted[bi];
    int32_t* state = ws.state + bi * 4;  // [0] done, [1] kept, [2] chunks walked
    if (n < 0) {
        if (threadIdx.x == 0 && phase == 0) out_counts[bi] = -1, state[0] = 1;
        return;
    }
    if (phase > 0 && state[0]) return;
    const int T = (n + 63) >> 6;
    const int hi = t_hi < T ? t_hi : T;  // this phase walks chunks [t_lo, hi); mask columns < hi exist
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64* __restrict__ keepmask = ws.keepmask + (size_t)bi * ws.T_cap;
    int32_t* __restrict__ kept_before = ws.kept_before + (size_t)bi * ws.T_cap;
    const u64* __restrict__ mask = ws.mask + (size_t)bi * ws.T_cap * ws.T_cap * 64;
    for (int c = threadIdx.x; c < hi; c += SCAN_THREADS) removed[c] = 0ull;
    __syncthreads();
    int kept = 0, chunks = 0;
    if (phase > 0) {
        kept = state[1], chunks = state[2];
        // columns [t_lo, hi) are new: apply the kept rows of every chunk walked so far to them
        const int width = hi - t_lo;
        for (int item = warp; item < t_lo * width; item += SCAN_WARPS) {
            const int r = item / width, c = t_lo + item % width;
            const u64 keep = keepmask[r];
            const u64* t = mask + ((size_t)r * ws.T_cap + c) * 64;
            const u64 v = (((keep >> lane) & 1ull) ? t[lane] : 0ull) | (((keep >> (lane + 32)) & 1ull) ? t[lane + 32] : 0ull);
            const uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)v);
            const uint32_t hi32 = __reduce_or_sync(0xffffffffu, (uint32_t)(v >> 32));
            if (lane == 0 && (lo | hi32)) atomicOr(reinterpret_cast<unsigned long long*>(removed + c), ((u64)hi32 << 32) | lo);
        }
        __syncthreads();
    }
    bool finished = (hi <= t_lo);  // nothing new to walk: the image ended in an earlier phase's range
    const size_t stride = ((size_t)ws.T_cap + 1) * 64;  // tile (r, c) -> tile (r+1, c+1)
    if (!finished) {
        if (warp == 0) {
            // ---- walker: diagonal tiles ----
            const u64* dp = mask + ((size_t)t_lo * ws.T_cap + t_lo) * 64 + lane;  // this lane's words of tile (r, r)
            u64 q_lo[SCAN_DEPTH + 1], q_hi[SCAN_DEPTH + 1];  // [0]: chunk r ... [SCAN_DEPTH]: chunk r + SCAN_DEPTH
#pragma unroll
            for (int d = 0; d < SCAN_DEPTH; ++d) {
                q_lo[d] = q_hi[d] = 0ull;
                if (t_lo + d < hi) q_lo[d] = dp[d * stride], q_hi[d] = dp[d * stride + 32];
            }
            q_lo[SCAN_DEPTH] = q_hi[SCAN_DEPTH] = 0ull;
            dp += SCAN_DEPTH * stride;  // next tile to fetch
            for (int r = t_lo;; ++r) {
                if (r + SCAN_DEPTH < hi) q_lo[SCAN_DEPTH] = dp[0], q_hi[SCAN_DEPTH] = dp[32];  // in flight for SCAN_DEPTH chunks
                dp += stride;
                const u64 a_lo = q_lo[0], a_hi = q_hi[0];
                u64 rem = removed[r];
                const int valid = n - r * 64;  // rows of this chunk that exist
                if (valid < 64) rem |= ~0ull << valid;
                u64 nz = (u64)__ballot_sync(0xffffffffu, a_lo != 0ull) | ((u64)__ballot_sync(0xffffffffu, a_hi != 0ull) << 32);
                while (nz) {  // only rows that suppress something inside the chunk need the serial step
                    const int i = __ffsll((long long)nz) - 1;
                    nz &= nz - 1;
                    const u64 di = (i < 32) ? shfl64(a_lo, i) : shfl64(a_hi, i - 32);
                    if (!((rem >> i) & 1ull)) rem |= di;
                }
                const u64 keep = ~rem;
                if (lane == 0) keep_s[r] = keep, kb_s[r] = kept;
                kept += __popcll(keep);
                chunks = r + 1;
                scan_bar();  // keep_s[r] is published
                if (kept >= max_det || r + 1 == T) {  // later boxes cannot enter the first max_det
                    finished = true;
                    break;
                }
                if (r + 1 == hi) break;  // end of this phase's columns
#pragma unroll
                for (int d = 0; d < SCAN_DEPTH; ++d) q_lo[d] = q_lo[d + 1], q_hi[d] = q_hi[d + 1];
                scan_bar();  // the helpers have applied chunk r's kept rows
            }
        } else if ((warp & 3) == 0) {
            // ---- warps on the walker's scheduler: follow the barriers and the exit test, nothing else ----
            for (int r = t_lo;; ++r) {
                scan_bar();
                kept += __popcll(keep_s[r]);
                chunks = r + 1;
                if (kept >= max_det || r + 1 == T) {
                    finished = true;
                    break;
                }
                if (r + 1 == hi) break;
                scan_bar();
            }
        } else {
            // ---- helpers: tiles (r, r + 1 + hw + k * SCAN_HELPERS), k < SCAN_BATCH, and the rare columns beyond ----
            const int hw = warp - 1 - (warp >> 2);  // 0..11 over the warps with warp % 4 != 0
            const u64* hp = mask + ((size_t)t_lo * ws.T_cap + (t_lo + 1 + hw)) * 64 + lane;
            // tile k of chunk r exists while r < hi - 1 - hw - 15 k
            int rmax[SCAN_BATCH];
#pragma unroll
            for (int k = 0; k < SCAN_BATCH; ++k) rmax[k] = hi - 1 - hw - k * SCAN_HELPERS;
            u64 q_lo[SCAN_DEPTH + 1][SCAN_BATCH], q_hi[SCAN_DEPTH + 1][SCAN_BATCH];  // [d]: this warp's tiles of chunk r + d
#pragma unroll
            for (int d = 0; d <= SCAN_DEPTH; ++d)
#pragma unroll
                for (int k = 0; k < SCAN_BATCH; ++k) {
                    q_lo[d][k] = q_hi[d][k] = 0ull;
                    if (d < SCAN_DEPTH && t_lo + d < rmax[k])
                        q_lo[d][k] = hp[d * stride + (size_t)k * SCAN_HELPERS * 64],
                        q_hi[d][k] = hp[d * stride + (size_t)k * SCAN_HELPERS * 64 + 32];
                }
            hp += SCAN_DEPTH * stride;
            for (int r = t_lo;; ++r) {
#pragma unroll
                for (int k = 0; k < SCAN_BATCH; ++k)
                    if (r + SCAN_DEPTH < rmax[k])
                        q_lo[SCAN_DEPTH][k] = hp[(size_t)k * SCAN_HELPERS * 64], q_hi[SCAN_DEPTH][k] = hp[(size_t)k * SCAN_HELPERS * 64 + 32];
                hp += stride;
                scan_bar();  // keep_s[r] is published
                const u64 keep = keep_s[r];
                kept += __popcll(keep);
                chunks = r + 1;
                if (kept >= max_det || r + 1 == T) {
                    finished = true;
                    break;
                }
                if (r + 1 == hi) break;
                const bool k_lo = (keep >> lane) & 1ull, k_hi = (keep >> (lane + 32)) & 1ull;
#pragma unroll
                for (int k = 0; k < SCAN_BATCH; ++k)
                    if (r < rmax[k]) {
                        // suppression words are sparse: the few lanes that hold a non-zero word of a kept row OR
                        // it in with a shared-memory atomic; no warp-wide reduction sits on the chunk's path
                        const u64 v = (k_lo ? q_lo[0][k] : 0ull) | (k_hi ? q_hi[0][k] : 0ull);
                        if (v) atomicOr(reinterpret_cast<unsigned long long*>(removed + r + 1 + hw + k * SCAN_HELPERS), v);
                    }
                if (r + 1 + hw + SCAN_BATCH * SCAN_HELPERS < hi) {  // very long rows: straight from global memory
                    const u64* __restrict__ tile_row = mask + (size_t)r * ws.T_cap * 64;
                    for (int c = r + 1 + hw + SCAN_BATCH * SCAN_HELPERS; c < hi; c += SCAN_HELPERS) {
                        const u64* t = tile_row + (size_t)c * 64;
                        const u64 v = (k_lo ? t[lane] : 0ull) | (k_hi ? t[lane + 32] : 0ull);
                        if (v) atomicOr(reinterpret_cast<unsigned long long*>(removed + c), v);
                    }
                }
#pragma unroll
                for (int d = 0; d < SCAN_DEPTH; ++d)
#pragma unroll
                    for (int k = 0; k < SCAN_BATCH; ++k) q_lo[d][k] = q_lo[d + 1][k], q_hi[d][k] = q_hi[d + 1][k];
                scan_bar();  // chunk r's kept rows are applied
            }
        }
    }
    __syncthreads();
    if (hi > t_lo) {  // park the walked chunks' keep masks for the gather and for later phases
        for (int c = t_lo + threadIdx.x; c < chunks; c += SCAN_THREADS) keepmask[c] = keep_s[c], kept_before[c] = kb_s[c];
        __syncthreads();
    }
    if (n == 0) finished = true;
    if (!finished) {  // more columns are needed: the next phase continues from here
        if (threadIdx.x == 0) {
            state[0] = 0, state[1] = kept, state[2] = chunks;
            if (phase == 0) atomicAdd(ws.pending, 1);
        }
        return;
    }
    // gather the kept rows in score order, all threads
    const uint32_t* __restrict__ sslot = ws.sslot + (size_t)bi * n_cap;
    const float* __restrict__ rows = cand_rows + (size_t)bi * cand_cap * 6;
    float* __restrict__ out = out_rows + (size_t)bi * max_det * 6;
    for (int t = threadIdx.x; t < chunks * 64; t += SCAN_THREADS) {
        const int r = t >> 6, i = t & 63;
        const u64 keep = keepmask[r];
        if ((keep >> i) & 1ull) {
            const int pos = kept_before[r] + __popcll(keep & ((1ull << i) - 1ull));
            if (pos < max_det) {
                const float2* src = reinterpret_cast<const float2*>(rows + (size_t)sslot[t] * 6);
                float2* dst = reinterpret_cast<float2*>(out + (size_t)pos * 6);
                dst[0] = src[0], dst[1] = src[1], dst[2] = src[2];
            }
        }
    }
    if (threadIdx.x == 0) out_counts[bi] = kept < max_det ? kept : max_det, state[0] = 1;
}

static int split_ws(void* workspace, int b, int cand_cap, int n_cap, SuppressWs& ws, size_t& total) {
    const int T_cap = (n_cap + 63) / 64;
    const int C_cap = (cand_cap + CHUNK - 1) / CHUNK;
    unsigned char* p = (unsigned char*)workspace;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        unsigned char* r = p ? p + o : nullptr;
        o += align_up(bytes, 256);
        return r;
    };
    ws.sbox = (float4*)take((size_t)b * n_cap * sizeof(float4));
    ws.sslot = (uint32_t*)take((size_t)b * n_cap * sizeof(uint32_t));
    ws.n_sorted = (int32_t*)take((size_t)b * sizeof(int32_t));
    ws.mask = (u64*)take((size_t)b * T_cap * T_cap * 64 * sizeof(u64));
    ws.ckeys = (u64*)take((size_t)b * C_cap * CHUNK * sizeof(u64));
    ws.cvals = (uint32_t*)take((size_t)b * C_cap * CHUNK * sizeof(uint32_t));
    ws.state = (int32_t*)take((size_t)b * 4 * sizeof(int32_t));
    ws.keepmask = (u64*)take((size_t)b * T_cap * sizeof(u64));
    ws.kept_before = (int32_t*)take((size_t)b * T_cap * sizeof(int32_t));
    ws.pending = (int32_t*)take(sizeof(int32_t));
    ws.T_cap = T_cap;
    ws.C_cap = C_cap;
    total = o;
    return 0;
}

}  // namespace kzd

using namespace kzd;

extern "C" __attribute__((visibility("default"))) int kzd_nms_filter(float* prediction, int b, int A, int nc, float conf_thres, int multi_label,
                              const int32_t* classes, int n_classes, int in_place, uint64_t* cand_keys,
                              float* cand_rows, int32_t* cand_counts, int cand_cap, void* stream) {
    KZD_CHECK_ARG(prediction && cand_keys && cand_rows && cand_counts, "kzd_nms_filter: null pointer");
    KZD_CHECK_ARG(b >= 0 && A >= 0 && nc >= 1 && cand_cap >= 1, "kzd_nms_filter: bad sizes b=%d A=%d nc=%d cap=%d", b,
                  A, nc, cand_cap);
    KZD_CHECK_ARG(b <= 65535, "kzd_nms_filter: b=%d > 65535", b);
    KZD_CHECK_ARG(in_place >= 0 && in_place <= 2, "kzd_nms_filter: in_place=%d not in {0,1,2}", in_place);
    KZD_CHECK_ARG(!multi_label || (double)A * nc < 4294967295.0, "kzd_nms_filter: A*nc overflows the order field");
    KZD_CHECK_ARG(n_classes == 0 || classes, "kzd_nms_filter: n_classes > 0 with null classes");
    cudaStream_t st = (cudaStream_t)stream;
    if (b == 0) return 0;
    KZD_CUDA(cudaMemsetAsync(cand_counts, 0, (size_t)b * sizeof(int32_t), st));
    if (A == 0) return 0;
    const int32_t* cls = n_classes ? classes : nullptr;
    if (multi_label && nc > 1) {
        dim3 grid((A + 255) / 256, b);
        nms_filter_multilabel_kernel<<<grid, 256, 0, st>>>(prediction, A, nc, conf_thres, cls, n_classes, in_place,
                                                           (u64*)cand_keys, cand_rows, cand_counts, cand_cap);
    } else if (A % 4 == 0 && ((uintptr_t)prediction) % 16 == 0) {
        dim3 grid((A / 4 + 255) / 256, b);
        nms_filter_kernel<4><<<grid, 256, 0, st>>>(prediction, A, nc, conf_thres, cls, n_classes, in_place,
                                                   (u64*)cand_keys, cand_rows, cand_counts, cand_cap);
    } else {
        dim3 grid((A + 255) / 256, b);
        nms_filter_kernel<1><<<grid, 256, 0, st>>>(prediction, A, nc, conf_thres, cls, n_classes, in_place,
                                                   (u64*)cand_keys, cand_rows, cand_counts, cand_cap);
    }
    KZD_LAUNCH_CHECK("nms_filter_kernel");
    return 0;
}

extern "C" __attribute__((visibility("default"))) int kzd_nms_append_labels(const float* labels, const int32_t* label_counts, int b, int L, int A,
                                     int nc, float conf_thres, int multi_label, const int32_t* classes, int n_classes,
                                     uint64_t* cand_keys, float* cand_rows, int32_t* cand_counts, int cand_cap,
                                     void* stream) {
    KZD_CHECK_ARG(label_counts && cand_keys && cand_rows && cand_counts, "kzd_nms_append_labels: null pointer");
    KZD_CHECK_ARG(labels || L == 0, "kzd_nms_append_labels: null labels with L=%d", L);
    KZD_CHECK_ARG(b >= 0 && L >= 0 && A >= 0 && nc >= 1 && cand_cap >= 1, "kzd_nms_append_labels: bad sizes");
    KZD_CHECK_ARG(b <= 65535, "kzd_nms_append_labels: b=%d > 65535", b);
    KZD_CHECK_ARG(n_classes == 0 || classes, "kzd_nms_append_labels: n_classes > 0 with null classes");
    const double base = (multi_label && nc > 1) ? (double)A * nc : (double)A;
    KZD_CHECK_ARG(base + L < 4294967295.0, "kzd_nms_append_labels: order field overflows");
    if (b == 0 || L == 0) return 0;
    dim3 grid((L + 255) / 256, b);
    nms_append_labels_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(labels, label_counts, L, nc, conf_thres,
                                                                   n_classes ? classes : nullptr, n_classes,
                                                                   (uint32_t)base, (u64*)cand_keys, cand_rows, cand_counts,
                                                                   cand_cap);
    KZD_LAUNCH_CHECK("nms_append_labels_kernel");
    return 0;
}

extern "C" __attribute__((visibility("default"))) size_t kzd_nms_workspace_bytes(int b, int cand_cap, int n_cap) {
    if (b < 0 || cand_cap < 1 || n_cap < 1) return 0;
    SuppressWs ws;
    size_t total = 0;
    split_ws(nullptr, b, cand_cap, n_cap, ws, total);
    return total;
}

extern "C" __attribute__((visibility("default"))) int kzd_nms_suppress(const uint64_t* cand_keys, const float* cand_rows, const int32_t* cand_counts, int b,
                                int cand_cap, int n_cap, float iou_thres, int agnostic, int max_det, int max_nms,
                                float max_wh, float* out_rows, int32_t* out_counts, void* workspace,
                                size_t workspace_bytes, void* stream) {
    KZD_CHECK_ARG(cand_keys && cand_rows && cand_counts && out_rows && out_counts && workspace,
                  "kzd_nms_suppress: null pointer");
    KZD_CHECK_ARG(b >= 0 && cand_cap >= 1 && n_cap >= 1 && max_det >= 1 && max_nms >= 1,
                  "kzd_nms_suppress: bad sizes");
    KZD_CHECK_ARG(b <= 65535, "kzd_nms_suppress: b=%d > 65535", b);
    KZD_CHECK_ARG(((uintptr_t)workspace) % 256 == 0, "kzd_nms_suppress: workspace must be 256-byte aligned");
    if (b == 0) return 0;
    SuppressWs ws;
    size_t need = 0;
    split_ws(workspace, b, cand_cap, n_cap, ws, need);
    if (workspace_bytes < need) {
        set_error("kzd_nms_suppress: workspace %zu B < required %zu B", workspace_bytes, need);
        return KZD_ENOSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
#ifndef KZD_NO_CLUSTER
#define KZD_NO_CLUSTER 0
#endif
#ifndef KZD_NO_BUCKET
#define KZD_NO_BUCKET 0
#endif
    if (!KZD_NO_BUCKET && n_cap <= 4096 && cand_cap <= 65535) {  // the usual capacity: one CTA per image ranks by buckets
        // (per call: the attribute belongs to the current device's copy of the kernel)
        KZD_CUDA(cudaFuncSetAttribute(nms_order_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BR_SMEM));
        nms_order_bucket_kernel<<<b, BR_THREADS, BR_SMEM, st>>>((const u64*)cand_keys, cand_rows, cand_counts, cand_cap, n_cap,
                                                               max_nms, agnostic, max_wh, ws);
        KZD_LAUNCH_CHECK("nms_order_bucket_kernel");
    } else if (!KZD_NO_CLUSTER && ws.C_cap <= ORDER_CL) {  // pool of at most 8 chunks: one cluster per image orders it in one launch
        nms_order_cluster_kernel<<<dim3(ORDER_CL, b), CHUNK_THREADS, 0, st>>>((const u64*)cand_keys, cand_rows, cand_counts,
                                                                            cand_cap, n_cap, max_nms, agnostic, max_wh, ws);
        KZD_LAUNCH_CHECK("nms_order_cluster_kernel");
    } else {
        dim3 cgrid(ws.C_cap, b);
        nms_chunk_sort_kernel<<<cgrid, CHUNK_THREADS, 0, st>>>((const u64*)cand_keys, cand_counts, cand_cap, n_cap,
                                                               max_nms, ws);
        KZD_LAUNCH_CHECK("nms_chunk_sort_kernel");
        nms_rank_kernel<<<cgrid, CHUNK_THREADS, 0, st>>>(cand_rows, cand_counts, cand_cap, n_cap, agnostic, max_wh, ws);
        KZD_LAUNCH_CHECK("nms_rank_kernel");
    }
    // The greedy walk stops as soon as max_det boxes are kept, so the mask is built and walked in two phases
    // of column ranges: first enough tiles for max_det kept boxes plus 25 %, then all the others. Images
    // whose walk ended in the first range skip the second on the device (no host round trip): when every
    // image did - the usual case when max_det cuts the walk - the second pair of launches costs ~4.5 us.
    // (A CUDA-graph conditional node around the second phase was measured: correct, but evaluating the node
    // costs ~14 us per replay, three times what the two idle launches do. Not kept.)
#ifndef KZD_MASK_GRID
#define KZD_MASK_GRID 16
#endif
#ifndef KZD_MASK_GRID0
#define KZD_MASK_GRID0 16
#endif
    const int grid = sm_count() * KZD_MASK_GRID;
    int t0 = ((max_det + max_det / 4 + 64 + 63) / 64 + MASK_CG - 1) / MASK_CG * MASK_CG;
    if (t0 > ws.T_cap) t0 = ws.T_cap;
    for (int phase = 0, t_lo = 0, t_hi = t0; t_lo < ws.T_cap; ++phase, t_lo = t_hi, t_hi = ws.T_cap) {
        const int g_lo = t_lo / MASK_CG, g_hi = (t_hi + MASK_CG - 1) / MASK_CG;
        // First phase: one CTA per tile when the machine can hold them all at once (32 resident 64-thread CTAs per
        // SM), so no CTA builds two tiles while others are done. The second phase usually finds every image
        // finished: a quarter of the persistent grid keeps that exit cheap (the kernel strides over its items, so a
        // phase with real work is only spread over fewer, longer CTAs).
        long long items0 = (long long)b * items_before_host(t_hi, g_hi);
        const long long cap0 = (long long)sm_count() * KZD_MASK_GRID0;
        const int grid0 = (int)(items0 < cap0 ? (items0 > grid ? items0 : grid) : cap0);
        nms_mask_kernel<<<phase == 0 ? grid0 : grid / 4, MASK_THREADS, (size_t)(b + 1) * sizeof(int), st>>>(
            b, n_cap, iou_thres, phase, g_lo, g_hi, ws);
        KZD_LAUNCH_CHECK("nms_mask_kernel");
        nms_scan_kernel<<<b, SCAN_THREADS, (size_t)ws.T_cap * 20, st>>>(cand_rows, cand_cap, n_cap, max_det, phase, t_lo, t_hi,
                                                                        ws, out_rows, out_counts);
        KZD_LAUNCH_CHECK("nms_scan_kernel");
    }
    return 0;
}
