// Batched NMS (sm_100a): ops.non_max_suppression + torchvision.ops.nms for all images at once.
//
// reference: utils/ops.py:167-316 (filter :234-286, class offset :289-295, nms call :296, [:max_det]
// :297, row gather :311) and torchvision's greedy NMS (pinned 0.21.0 in the reference's uv.lock;
// semantics restated in-tree at examples/YOLOv8-LibTorch-CPP-Inference/main.cc:80-136).
//
// Pipeline (no host synchronisation inside; candidate counts live on the device):
//   nms_filter_kernel  conf filter + best class / multi-label + xywh->xyxy, candidates appended to
//                      a per-image pool with a 64-bit key  score_bits<<32 | ~order
//   nms_sort_kernel    per image: bitonic sort of the keys (descending) = torchvision's stable
//                      descending score sort; writes score-ordered, class-offset boxes
//   nms_mask_kernel    persistent CTAs over 64x64 tiles of the upper triangle: bit (i,j) = IoU > thr
//   nms_scan_kernel    per image: greedy scan in 64-row chunks; the diagonal chunk is resolved by
//                      one warp with ballot/shuffle, the rest of the row OR-reduced by all warps;
//                      the first max_det kept rows are gathered in score order
#include "kzd_common.cuh"

namespace kzd {

typedef unsigned long long u64;

// ------------------------------------------------------------------------------------------
// stage 1: filter
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ bool class_allowed(int c, const int32_t* __restrict__ classes, int n_classes) {
    if (!classes) return true;
    for (int k = 0; k < n_classes; ++k)
        if (classes[k] == c) return true;
    return false;
}

// Append one candidate per set lane; slots are handed out with one atomic per warp.
__device__ __forceinline__ void append_candidate(bool is_cand, float x1, float y1, float x2, float y2, float conf,
                                                 int cls, uint32_t order, u64* __restrict__ keys,
                                                 float* __restrict__ rows, int32_t* __restrict__ count, int cap) {
    const unsigned m = __ballot_sync(0xffffffffu, is_cand);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (is_cand) {
        const int slot = base + __popc(m & ((1u << lane) - 1u));
        if (slot < cap) {
            keys[slot] = ((u64)__float_as_uint(conf) << 32) | (u64)(0xffffffffu - order);
            float* r = rows + (size_t)slot * 6;
            r[0] = x1, r[1] = y1, r[2] = x2, r[3] = y2, r[4] = conf, r[5] = (float)cls;
        }
    }
}

__global__ void __launch_bounds__(256) nms_filter_kernel(float* pred, int A, int nc, float conf_thres,
                                                         int multi_label, const int32_t* __restrict__ classes,
                                                         int n_classes, int in_place, u64* __restrict__ keys,
                                                         float* __restrict__ rows, int32_t* __restrict__ counts,
                                                         int cap) {
    const int bi = blockIdx.y;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;  // whole warps stay alive for the ballots
    const bool live = a < A;
    float* base = pred + (size_t)bi * (4 + nc) * A;
    u64* k_img = keys + (size_t)bi * cap;
    float* r_img = rows + (size_t)bi * cap * 6;
    int32_t* cnt = counts + bi;

    float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f;
    if (live) {
        const float cx = base[a], cy = base[(size_t)A + a];
        const float bw = base[2 * (size_t)A + a], bh = base[3 * (size_t)A + a];
        if (in_place == 2) {  // rows already hold xyxy (second pass after a pool overflow)
            x1 = cx, y1 = cy, x2 = bw, y2 = bh;
        } else {
            const float hw = bw * 0.5f, hh = bh * 0.5f;
            x1 = cx - hw, y1 = cy - hh, x2 = cx + hw, y2 = cy + hh;  // utils/ops.py:429-432
            if (in_place == 1) {
                base[a] = x1, base[(size_t)A + a] = y1, base[2 * (size_t)A + a] = x2, base[3 * (size_t)A + a] = y2;
            }
        }
    }
    const float* cls = base + 4 * (size_t)A;
    if (!multi_label) {
        float best = 0.f;
        int j = 0;
        if (live) {
            best = cls[a];
            for (int c = 1; c < nc; ++c) {
                const float s = cls[(size_t)c * A + a];
                if (s > best) best = s, j = c;  // first maximum wins (utils/ops.py:274)
            }
        }
        const bool is_cand = live && best > conf_thres && class_allowed(j, classes, n_classes);
        append_candidate(is_cand, x1, y1, x2, y2, best, j, (uint32_t)a, k_img, r_img, cnt, cap);
    } else {
        for (int c = 0; c < nc; ++c) {  // utils/ops.py:271-272: one row per (anchor, class) above conf
            const float s = live ? cls[(size_t)c * A + a] : 0.f;
            const bool is_cand = live && s > conf_thres && class_allowed(c, classes, n_classes);
            append_candidate(is_cand, x1, y1, x2, y2, s, c, (uint32_t)a * (uint32_t)nc + (uint32_t)c, k_img, r_img,
                             cnt, cap);
        }
    }
}

// ------------------------------------------------------------------------------------------
// stage 2a: per-image sort
// ------------------------------------------------------------------------------------------

constexpr int SORT_THREADS = 1024;
constexpr int SORT_SMEM_ELEMS = 8192;  // 8192 * (8 + 4) B = 96 KB

__device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// descending bitonic network over P = 2^k elements; padding keys are 0 and sink to the tail
template <typename KeyPtr, typename ValPtr>
__device__ __forceinline__ void bitonic_desc(KeyPtr keys, ValPtr vals, int P) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const u64 ki = keys[i], kp = keys[p];
                const bool desc = (i & k) == 0;
                if ((ki < kp) == desc) {
                    keys[i] = kp, keys[p] = ki;
                    const uint32_t vi = vals[i];
                    vals[i] = vals[p], vals[p] = vi;
                }
            }
            __syncthreads();
        }
    }
}

struct SuppressWs {
    float4* sbox;     // (b, n_cap)   score-ordered boxes, class offset applied
    uint32_t* sslot;  // (b, n_cap)   pool slot of each ordered box
    int32_t* n_sorted;  // (b)        boxes entering suppression, or -1 on overflow
    u64* mask;        // (b, T_cap*T_cap*64)
    u64* gkeys;       // (b, P_cap)   global-memory sort buffers (only used above SORT_SMEM_ELEMS)
    uint32_t* gvals;  // (b, P_cap)
    int T_cap, P_cap;
};

__global__ void __launch_bounds__(SORT_THREADS) nms_sort_kernel(const u64* __restrict__ cand_keys,
                                                                const float* __restrict__ cand_rows,
                                                                const int32_t* __restrict__ cand_counts, int cand_cap,
                                                                int n_cap, int max_nms, int agnostic, float max_wh,
                                                                SuppressWs ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int bi = blockIdx.x;
    const int count = cand_counts[bi];
    int n_all = count;
    int n = n_all < max_nms ? n_all : max_nms;
    if (count > cand_cap || n > n_cap) {  // pool or suppression stage too small: report, caller retries
        if (threadIdx.x == 0) ws.n_sorted[bi] = -1;
        return;
    }
    if (threadIdx.x == 0) ws.n_sorted[bi] = n;
    if (n == 0) return;
    const u64* __restrict__ src = cand_keys + (size_t)bi * cand_cap;
    const int P = next_pow2(n_all);
    float4* __restrict__ sbox = ws.sbox + (size_t)bi * n_cap;
    uint32_t* __restrict__ sslot = ws.sslot + (size_t)bi * n_cap;
    const float* __restrict__ rows = cand_rows + (size_t)bi * cand_cap * 6;

    auto emit = [&](auto keys, auto vals) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t slot = vals[i];
            const float* r = rows + (size_t)slot * 6;
            const float off = agnostic ? 0.f : __fmul_rn(r[5], max_wh);  // utils/ops.py:289
            sbox[i] = make_float4(__fadd_rn(r[0], off), __fadd_rn(r[1], off), __fadd_rn(r[2], off),
                                  __fadd_rn(r[3], off));                 // utils/ops.py:295
            sslot[i] = slot;
        }
    };
    if (P <= SORT_SMEM_ELEMS) {
        u64* keys = reinterpret_cast<u64*>(smem_raw);
        uint32_t* vals = reinterpret_cast<uint32_t*>(smem_raw + (size_t)SORT_SMEM_ELEMS * 8);
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            keys[i] = i < n_all ? src[i] : 0ull;
            vals[i] = (uint32_t)i;
        }
        __syncthreads();
        bitonic_desc(keys, vals, P);
        emit(keys, vals);
    } else {
        u64* keys = ws.gkeys + (size_t)bi * ws.P_cap;
        uint32_t* vals = ws.gvals + (size_t)bi * ws.P_cap;
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            keys[i] = i < n_all ? src[i] : 0ull;
            vals[i] = (uint32_t)i;
        }
        __syncthreads();
        bitonic_desc(keys, vals, P);
        emit(keys, vals);
    }
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
    if (!(w > 0.f && h > 0.f)) return false;  // inter == 0  =>  0/u (or NaN) is never > thr >= 0
    const float inter = __fmul_rn(w, h);
    const float u = __fsub_rn(__fmaf_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y), sa), inter);
    return __fdiv_rn(inter, u) > thr;
}

// Work item = (image, row tile r, column tile c >= r) over all images; items are striped over a
// persistent grid. Thread i of the CTA owns row r*64+i and writes one 64-bit word of the mask.
// Mask layout: [image][r*T_cap + c][i]  (64 consecutive words per tile => coalesced 512 B stores,
// and the scan reads whole tiles).
__global__ void __launch_bounds__(MASK_THREADS) nms_mask_kernel(int b, int n_cap, float iou_thres, SuppressWs ws) {
    extern __shared__ int pre[];  // (b+1) prefix of tile-pair counts
    __shared__ float4 colbox[64];
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int i = 0; i < b; ++i) {
            pre[i] = acc;
            const int n = ws.n_sorted[i];
            const int T = n > 0 ? (n + 63) >> 6 : 0;
            acc += T * (T + 1) / 2;
        }
        pre[b] = acc;
    }
    __syncthreads();
    const int total = pre[b];
    const int i = threadIdx.x;
    int bi = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        while (pre[bi + 1] <= w) ++bi;  // w only grows
        const int n = ws.n_sorted[bi];
        const int T = (n + 63) >> 6;
        const int q = w - pre[bi];
        // invert q = r*T - r(r-1)/2 + (c - r)
        int r = (int)(((2.0 * T + 1.0) - sqrt((2.0 * T + 1.0) * (2.0 * T + 1.0) - 8.0 * (double)q)) * 0.5);
        if (r < 0) r = 0;
        if (r > T - 1) r = T - 1;
        while (r > 0 && r * T - r * (r - 1) / 2 > q) --r;
        while (r + 1 < T && (r + 1) * T - (r + 1) * r / 2 <= q) ++r;
        const int c = r + (q - (r * T - r * (r - 1) / 2));
        const float4* __restrict__ sbox = ws.sbox + (size_t)bi * n_cap;
        const int row = r * 64 + i, col0 = c * 64;
        __syncthreads();  // previous item's colbox readers are done
        colbox[i] = (col0 + i < n) ? sbox[col0 + i] : make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        u64 word = 0;
        if (row < n) {
            const float4 a = sbox[row];
            const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
            const int jn = min(64, n - col0);
            const int j0 = (r == c) ? i + 1 : 0;
            for (int j = j0; j < jn; ++j)
                if (iou_gt(a, sa, colbox[j], iou_thres)) word |= 1ull << j;
        }
        ws.mask[((size_t)bi * ws.T_cap * ws.T_cap + (size_t)r * ws.T_cap + c) * 64 + i] = word;
    }
}

// ------------------------------------------------------------------------------------------
// stage 2c: greedy scan + row gather
// ------------------------------------------------------------------------------------------

constexpr int SCAN_THREADS = 256;

__device__ __forceinline__ u64 shfl64(u64 v, int src) {
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
    const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
    return ((u64)hi << 32) | lo;
}

__global__ void __launch_bounds__(SCAN_THREADS) nms_scan_kernel(const float* __restrict__ cand_rows, int cand_cap,
                                                                int n_cap, int max_det, SuppressWs ws,
                                                                float* __restrict__ out_rows,
                                                                int32_t* __restrict__ out_counts) {
    extern __shared__ u64 removed[];  // T_cap words
    __shared__ u64 keep_sh;
    const int bi = blockIdx.x;
    const int n = ws.n_sorted[bi];
    if (n < 0) {
        if (threadIdx.x == 0) out_counts[bi] = -1;
        return;
    }
    const int T = (n + 63) >> 6;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = SCAN_THREADS / 32;
    for (int c = threadIdx.x; c < T; c += SCAN_THREADS) removed[c] = 0ull;
    __syncthreads();
    const u64* __restrict__ mask = ws.mask + (size_t)bi * ws.T_cap * ws.T_cap * 64;
    const uint32_t* __restrict__ sslot = ws.sslot + (size_t)bi * n_cap;
    const float* __restrict__ rows = cand_rows + (size_t)bi * cand_cap * 6;
    float* __restrict__ out = out_rows + (size_t)bi * max_det * 6;
    int kept = 0;
    for (int r = 0; r < T; ++r) {
        const u64* __restrict__ tile_row = mask + (size_t)r * ws.T_cap * 64;
        if (warp == 0) {
            // diagonal tile: rows r*64..r*64+63 against themselves (bits j > i only)
            const u64* d = tile_row + (size_t)r * 64;
            const u64 d_lo = d[lane], d_hi = d[lane + 32];
            u64 rem = removed[r];
            const int valid = n - r * 64;  // rows of this chunk that exist
            if (valid < 64) rem |= ~0ull << valid;
            u64 nz = (u64)__ballot_sync(0xffffffffu, d_lo != 0ull) |
                     ((u64)__ballot_sync(0xffffffffu, d_hi != 0ull) << 32);
            while (nz) {  // only rows that suppress something inside the chunk need the serial step
                const int i = __ffsll((long long)nz) - 1;
                nz &= nz - 1;
                const u64 di = (i < 32) ? shfl64(d_lo, i) : shfl64(d_hi, i - 32);
                if (!((rem >> i) & 1ull)) rem |= di;
            }
            if (lane == 0) keep_sh = ~rem;
        }
        __syncthreads();
        const u64 keep = keep_sh;
        // gather kept rows of this chunk, in order
        if (threadIdx.x < 64 && ((keep >> threadIdx.x) & 1ull)) {
            const int pos = kept + __popcll(keep & ((1ull << threadIdx.x) - 1ull));
            if (pos < max_det) {
                const float* src = rows + (size_t)sslot[r * 64 + threadIdx.x] * 6;
                float* dst = out + (size_t)pos * 6;
#pragma unroll
                for (int k = 0; k < 6; ++k) dst[k] = src[k];
            }
        }
        kept += __popcll(keep);
        if (kept >= max_det) break;  // uniform: later boxes cannot enter the first max_det
        // OR the kept rows into the removed words of the later chunks
        const bool k_lo = (keep >> lane) & 1ull, k_hi = (keep >> (lane + 32)) & 1ull;
        for (int c = r + 1 + warp; c < T; c += n_warps) {
            const u64* t = tile_row + (size_t)c * 64;
            u64 v = (k_lo ? t[lane] : 0ull) | (k_hi ? t[lane + 32] : 0ull);
            const uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)v);
            const uint32_t hi = __reduce_or_sync(0xffffffffu, (uint32_t)(v >> 32));
            if (lane == 0) removed[c] |= ((u64)hi << 32) | lo;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out_counts[bi] = kept < max_det ? kept : max_det;
}

static int split_ws(void* workspace, int b, int cand_cap, int n_cap, SuppressWs& ws, size_t& total) {
    const int T_cap = (n_cap + 63) / 64;
    int P_cap = 1;
    while (P_cap < cand_cap) P_cap <<= 1;
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
    const bool need_global_sort = P_cap > SORT_SMEM_ELEMS;
    ws.gkeys = (u64*)take(need_global_sort ? (size_t)b * P_cap * sizeof(u64) : 0);
    ws.gvals = (uint32_t*)take(need_global_sort ? (size_t)b * P_cap * sizeof(uint32_t) : 0);
    ws.T_cap = T_cap;
    ws.P_cap = P_cap;
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
    KZD_CHECK_ARG(!multi_label || (double)A * nc < 4294967295.0, "kzd_nms_filter: A*nc overflows the order field");
    KZD_CHECK_ARG(n_classes == 0 || classes, "kzd_nms_filter: n_classes > 0 with null classes");
    cudaStream_t st = (cudaStream_t)stream;
    if (b == 0) return 0;
    KZD_CUDA(cudaMemsetAsync(cand_counts, 0, (size_t)b * sizeof(int32_t), st));
    if (A == 0) return 0;
    dim3 grid((A + 255) / 256, b);
    nms_filter_kernel<<<grid, 256, 0, st>>>(prediction, A, nc, conf_thres, multi_label && nc > 1,
                                            n_classes ? classes : nullptr, n_classes, in_place,
                                            (u64*)cand_keys, cand_rows, cand_counts, cand_cap);
    KZD_LAUNCH_CHECK("nms_filter_kernel");
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
    const size_t sort_smem = (size_t)SORT_SMEM_ELEMS * 12;
    KZD_CUDA(cudaFuncSetAttribute(nms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem));
    nms_sort_kernel<<<b, SORT_THREADS, sort_smem, st>>>((const u64*)cand_keys, cand_rows, cand_counts, cand_cap, n_cap,
                                                        max_nms, agnostic, max_wh, ws);
    KZD_LAUNCH_CHECK("nms_sort_kernel");
    const int grid = sm_count() * 16;
    nms_mask_kernel<<<grid, MASK_THREADS, (size_t)(b + 1) * sizeof(int), st>>>(b, n_cap, iou_thres, ws);
    KZD_LAUNCH_CHECK("nms_mask_kernel");
    nms_scan_kernel<<<b, SCAN_THREADS, (size_t)ws.T_cap * sizeof(u64), st>>>(cand_rows, cand_cap, n_cap, max_det, ws,
                                                                            out_rows, out_counts);
    KZD_LAUNCH_CHECK("nms_scan_kernel");
    return 0;
}
