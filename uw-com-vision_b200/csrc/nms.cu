// Kernel 4: batched score filter + per-class greedy NMS + top-k on the box list.
//
// Replaces detectron2/modeling/roi_heads/fast_rcnn.py::fast_rcnn_inference_single_image's
// "scores > thresh -> batched_nms -> keep[:topk]" (thresholds set at
// nn_inference.py:226) with the per-class "vanilla" semantics of
// torchvision.ops.boxes._batched_nms_vanilla.
//
// float32 arithmetic of torchvision's nms kernel, each op rounded separately:
//   area = (x2 - x1) * (y2 - y1);  w = max(0, min(x2) - max(x1));  h likewise;
//   inter = w * h;  iou = inter / ((area_i + area_j) - inter);  suppress iff (double)iou > thr
// Within a class candidates are visited in (score descending, index ascending) order.
//
// Launches for the whole batch:
//   A  per image: candidates above the score threshold are COMPACTED first (the cost of every
//      later step follows the survivors, as upstream's "scores > thresh" before batched_nms
//      does), key = (class, ~orderable(score), index), bitonic sort of the survivors (shared-
//      memory passes for strides < 4096), gather sorted boxes, class segment table
//   C  per (image, class): greedy sweep over the class segment, 64 candidates per step: the
//      64 x 64 IoU bits of the step are computed in place, the step's keepers resolved by one
//      thread in registers, and every later candidate of the segment is tested against those
//      keepers on the fly (early exit at the first hit).  No suppression matrix is stored:
//      the workspace is O(R), not O(R^2 / 64) (r01: 800 MB per image at R = 1000 x 80 classes)
//   D  per image: merge the per-class keep lists by score (rank = own rank + lower bounds
//      in the other classes' lists), top-k
#include "uwcv_common.cuh"

namespace uwcv {

constexpr int kNmsThreads = 1024;
constexpr int kIdxBits = 18;            // <= 262144 candidates per image
constexpr int kClsShift = 50;           // key = class << 50 | ~score << 18 | index
constexpr int kSortChunk = 4096;        // keys per shared-memory sort chunk

struct NmsWorkspace {
  int64_t* off;        // [B + 1] device copy of the per-image candidate offsets
  int32_t* nvalid;     // [B]
  uint64_t* keys;      // [2 R + B]
  float4* sbox;        // [R] boxes in sorted order (per image segment)
  int32_t* scls;       // [R]
  int32_t* seg;        // [B][C][2] class segments (lo, hi) in sorted positions
  int32_t* ckeep;      // [R] per-class keep lists (sorted positions), stored at the segment start
  int32_t* ccount;     // [B][C]
};

__host__ __device__ inline size_t nms_fixed_bytes(int64_t R, int B, int C) {
  size_t s = 0;
  s += align_up((size_t)(B + 1) * 8, 256);
  s += align_up((size_t)B * 4, 256);
  s += align_up((size_t)(2 * R + B) * 8, 256);
  s += align_up((size_t)R * 16, 256);
  s += align_up((size_t)R * 4, 256);
  s += align_up((size_t)B * C * 8, 256);
  s += align_up((size_t)R * 4, 256);
  s += align_up((size_t)B * C * 4, 256);
  return s;
}

inline NmsWorkspace nms_carve(void* ws, int64_t R, int B, int C) {
  NmsWorkspace w;
  char* p = (char*)ws;
  w.off = (int64_t*)p;      p += align_up((size_t)(B + 1) * 8, 256);
  w.nvalid = (int32_t*)p;   p += align_up((size_t)B * 4, 256);
  w.keys = (uint64_t*)p;    p += align_up((size_t)(2 * R + B) * 8, 256);
  w.sbox = (float4*)p;      p += align_up((size_t)R * 16, 256);
  w.scls = (int32_t*)p;     p += align_up((size_t)R * 4, 256);
  w.seg = (int32_t*)p;      p += align_up((size_t)B * C * 8, 256);
  w.ckeep = (int32_t*)p;    p += align_up((size_t)R * 4, 256);
  w.ccount = (int32_t*)p;
  return w;
}

__device__ __forceinline__ uint32_t orderable(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// one bitonic compare-exchange pass with stride j of the stage k on keys[0, count)
// (global index of keys[0] is `base`; count and base are multiples of 2 j)
__device__ __forceinline__ void bitonic_pass(uint64_t* keys, int count, int base, int k, int j,
                                             int tid) {
  for (int t = tid; t < count / 2; t += kNmsThreads) {
    const int i = ((t / j) * 2 * j) + (t % j);          // lower element of the pair
    const int p = i + j;
    const uint64_t a = keys[i], c = keys[p];
    const bool asc = ((base + i) & k) == 0;
    if ((a > c) == asc) { keys[i] = c; keys[p] = a; }
  }
}

// ---- A: filter + sort + class segments --------------------------------------------------
__global__ void __launch_bounds__(kNmsThreads)
nms_sort_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                const int64_t* __restrict__ cls, float score_thr, int C, NmsWorkspace w) {
  __shared__ uint64_t s_keys[kSortChunk];
  __shared__ int s_count;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int64_t lo = w.off[b];
  const int n = (int)(w.off[b + 1] - lo);
  uint64_t* keys = w.keys + 2 * lo + b;
  if (tid == 0) s_count = 0;
  for (int k = tid; k < C; k += kNmsThreads) {
    w.seg[((int64_t)b * C + k) * 2] = 0;
    w.seg[((int64_t)b * C + k) * 2 + 1] = 0;
    w.ccount[(int64_t)b * C + k] = 0;
  }
  __syncthreads();
  // score filter first: the survivors are appended (warp-aggregated) to keys[0, nv); the order
  // of the appends does not matter, the sort below fixes it
  for (int i0 = 0; i0 < n; i0 += kNmsThreads) {
    const int i = i0 + tid;
    uint64_t k = ~0ull;
    bool ok = false;
    if (i < n) {
      const float s = scores[lo + i];
      const float4 bx = reinterpret_cast<const float4*>(boxes)[lo + i];
      const long long c = cls[lo + i];
      const bool fin = isfinite(bx.x) && isfinite(bx.y) && isfinite(bx.z) && isfinite(bx.w) &&
                       isfinite(s);
      if (fin && s > score_thr && c >= 0 && c < C) {
        k = ((uint64_t)c << kClsShift) | ((uint64_t)(~orderable(s)) << kIdxBits) | (uint32_t)i;
        ok = true;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    int base = 0;
    if ((tid & 31) == 0 && m) base = atomicAdd(&s_count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (ok) keys[base + __popc(m & ((1u << (tid & 31)) - 1u))] = k;
  }
  __syncthreads();
  const int nvalid_ = s_count;
  int P = 1;
  while (P < nvalid_) P <<= 1;
  for (int i = nvalid_ + tid; i < P; i += kNmsThreads) keys[i] = ~0ull;      // padding sorts last
  __syncthreads();
  // stages k <= chunk: every chunk is sorted entirely in shared memory
  const int chunk = P < kSortChunk ? P : kSortChunk;
  if (chunk >= 2) {
    for (int c0 = 0; c0 < P; c0 += chunk) {
      for (int i = tid; i < chunk; i += kNmsThreads) s_keys[i] = keys[c0 + i];
      __syncthreads();
      for (int k = 2; k <= chunk; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
          bitonic_pass(s_keys, chunk, c0, k, j, tid);
          __syncthreads();
        }
      for (int i = tid; i < chunk; i += kNmsThreads) keys[c0 + i] = s_keys[i];
      __syncthreads();
    }
    // stages k > chunk: strides >= chunk in global memory, the rest per chunk in shared memory
    for (int k = 2 * chunk; k <= P; k <<= 1) {
      for (int j = k >> 1; j >= chunk; j >>= 1) {
        bitonic_pass(keys, P, 0, k, j, tid);
        __syncthreads();
      }
      for (int c0 = 0; c0 < P; c0 += chunk) {
        for (int i = tid; i < chunk; i += kNmsThreads) s_keys[i] = keys[c0 + i];
        __syncthreads();
        for (int j = chunk >> 1; j > 0; j >>= 1) {
          bitonic_pass(s_keys, chunk, c0, k, j, tid);
          __syncthreads();
        }
        for (int i = tid; i < chunk; i += kNmsThreads) keys[c0 + i] = s_keys[i];
        __syncthreads();
      }
    }
  }
  const int nv = s_count;
  if (tid == 0) w.nvalid[b] = nv;
  for (int r = tid; r < nv; r += kNmsThreads) {
    const uint64_t key = keys[r];
    const int li = (int)(key & ((1u << kIdxBits) - 1u));
    const int c = (int)(key >> kClsShift);
    w.sbox[lo + r] = reinterpret_cast<const float4*>(boxes)[lo + li];
    w.scls[lo + r] = c;
    const int cp = r > 0 ? (int)(keys[r - 1] >> kClsShift) : -1;
    const int cn = r + 1 < nv ? (int)(keys[r + 1] >> kClsShift) : -1;
    if (cp != c) w.seg[((int64_t)b * C + c) * 2] = r;
    if (cn != c) w.seg[((int64_t)b * C + c) * 2 + 1] = r + 1;
  }
}

// ---- C: greedy sweep of one class segment, IoUs computed on the fly -------------------------
constexpr int kMaxBlocks = 4096;     // 262144 candidates per image
constexpr int kSweepThreads = 1024;
constexpr int kSweepSmemBoxes = 8192;            // class segments up to this size are staged in shared memory

// suppress iff iou > thr, with torchvision's float32 arithmetic; boxes that do not intersect are
// decided without the division (inter == 0 -> iou == 0 or NaN, never above a threshold >= 0)
__device__ __forceinline__ bool iou_above(const float4& bi, float area_i, const float4& bj,
                                          double iou_thr) {
  const float ww = fmaxf(0.f, __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)));
  const float hh = fmaxf(0.f, __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y)));
  const float inter = __fmul_rn(ww, hh);
  if (!(inter > 0.f) && iou_thr >= 0.0) return false;
  const float area_j = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
  const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_i, area_j), inter));
  return (double)iou > iou_thr;
}

__global__ void __launch_bounds__(kSweepThreads)
nms_sweep_kernel(int topk, int C, double iou_thr, NmsWorkspace w) {
  extern __shared__ float4 s_seg[];          // the segment's boxes (when it fits)
  const int b = blockIdx.y, cls = blockIdx.x, tid = threadIdx.x;
  const int s = w.seg[((int64_t)b * C + cls) * 2], e = w.seg[((int64_t)b * C + cls) * 2 + 1];
  if (e <= s) return;
  const int64_t lo = w.off[b];
  const int c_first = s >> 6, c_last = (e - 1) >> 6;
  __shared__ uint64_t s_removed[kMaxBlocks];
  __shared__ uint64_t s_diag[64];
  __shared__ float4 s_kbox[64];              // boxes kept in the current step
  __shared__ uint64_t s_keepbits;
  __shared__ int s_kept;
  const bool staged = (e - s) <= kSweepSmemBoxes;
  if (staged)
    for (int k = s + tid; k < e; k += kSweepThreads) s_seg[k - s] = w.sbox[lo + k];
  // box at sorted position p of this image (p in [s, e))
  const float4* bx = staged ? (s_seg - s) : (w.sbox + lo);
  for (int k = c_first + tid; k <= c_last; k += kSweepThreads) s_removed[k] = 0ull;
  if (tid == 0) s_kept = 0;
  __syncthreads();
  for (int c = c_first; c <= c_last; ++c) {
    const int ifirst = c == c_first ? (s & 63) : 0;          // positions < s: previous class
    const int nin = min(64, e - c * 64);                     // positions >= e belong to the next class
    const int kept_before = s_kept;          // stable: last written before the previous barrier
    // the step's 64 x 64 suppression bits, upper triangle: 16 threads per row i, each testing
    // four of the candidates j > i (rows of candidates already removed are never consulted)
    if (tid < 64) s_diag[tid] = 0ull;
    __syncthreads();
    {
      const int i = tid >> 4, j0 = (tid & 15) * 4;
      if (i >= ifirst && i < nin && !((s_removed[c] >> i) & 1ull)) {
        const float4 bi = bx[c * 64 + i];
        const float area_i = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
        uint64_t bits = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = j0 + k;
          if (j > i && j < nin && iou_above(bi, area_i, bx[c * 64 + j], iou_thr)) bits |= 1ull << j;
        }
        if (bits) atomicOr((unsigned long long*)&s_diag[i], (unsigned long long)bits);
      }
    }
    __syncthreads();
    if (tid == 0) {
      uint64_t cur = s_removed[c], kb = 0;
      int kept = kept_before;
      for (int i = ifirst; i < nin && kept < topk; ++i) {
        if (!((cur >> i) & 1ull)) { kb |= 1ull << i; ++kept; cur |= s_diag[i]; }
      }
      s_keepbits = kb;
      s_kept = kept;
    }
    __syncthreads();
    const uint64_t kb = s_keepbits;
    const int nkept = __popcll(kb);
    if (tid < 64 && ((kb >> tid) & 1ull)) {
      const int rank = __popcll(kb & ((1ull << tid) - 1ull));
      w.ckeep[lo + s + kept_before + rank] = c * 64 + tid;      // sorted position of a kept box
      s_kbox[rank] = bx[c * 64 + tid];
    }
    if (s_kept >= topk) break;
    __syncthreads();
    // every later candidate of the segment that is still alive against this step's keepers
    // (score order: a keeper always precedes the candidates it suppresses; torchvision
    // evaluates iou(keeper, candidate) with the keeper first: same operand order here)
    if (nkept) {
      for (int j = (c + 1) * 64 + tid; j < e; j += kSweepThreads) {
        if ((s_removed[j >> 6] >> (j & 63)) & 1ull) continue;
        const float4 bj = bx[j];
        bool hit = false;
        for (int r = 0; r < nkept && !hit; ++r) {
          const float4 bi = s_kbox[r];
          hit = iou_above(bi, __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y)), bj, iou_thr);
        }
        if (hit) atomicOr((unsigned long long*)&s_removed[j >> 6], 1ull << (j & 63));
      }
    }
    __syncthreads();
  }
  __syncthreads();
  if (tid == 0) w.ccount[(int64_t)b * C + cls] = s_kept;
}

// ---- D: merge the per-class keep lists by score, top-k -------------------------------------
__global__ void __launch_bounds__(kNmsThreads)
nms_merge_kernel(int topk, int C, NmsWorkspace w, int64_t* __restrict__ keep,
                 int32_t* __restrict__ keep_count) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const int64_t lo = w.off[b];
  const uint64_t* keys = w.keys + 2 * lo + b;
  const uint64_t kScoreMask = (1ull << kClsShift) - 1ull;      // (~score, index): global order
  __shared__ int s_total;
  if (tid == 0) {
    int t = 0;
    for (int k = 0; k < C; ++k) t += w.ccount[(int64_t)b * C + k];
    s_total = t;
    keep_count[b] = t < topk ? t : topk;
  }
  __syncthreads();
  if (s_total == 0) return;
  for (int k = 0; k < C; ++k) {
    const int cnt = w.ccount[(int64_t)b * C + k];
    const int s = w.seg[((int64_t)b * C + k) * 2];
    for (int j = tid; j < cnt; j += kNmsThreads) {
      const uint64_t key = keys[w.ckeep[lo + s + j]];
      const uint64_t mine = key & kScoreMask;
      int rank = j;
      for (int k2 = 0; k2 < C; ++k2) {
        if (k2 == k) continue;
        const int cnt2 = w.ccount[(int64_t)b * C + k2];
        if (!cnt2) continue;
        const int s2 = w.seg[((int64_t)b * C + k2) * 2];
        int a = 0, z = cnt2;                               // lower bound of `mine` in class k2
        while (a < z) {
          const int mid = (a + z) >> 1;
          if ((keys[w.ckeep[lo + s2 + mid]] & kScoreMask) < mine) a = mid + 1; else z = mid;
        }
        rank += a;
      }
      if (rank < topk) keep[lo + rank] = lo + (int64_t)(key & ((1u << kIdxBits) - 1u));
    }
  }
}

// The per-image candidate offsets reach the device as KERNEL ARGUMENTS (32 per launch): the host
// array is only read during the call, nothing is copied asynchronously from pageable memory, the
// call never synchronises and can be captured in a CUDA graph.
struct OffsetChunk { int64_t v[32]; };
__global__ void nms_set_offsets_kernel(int64_t* __restrict__ off, int first, int count, OffsetChunk c) {
  const int t = threadIdx.x;
  if (t < count) off[first + t] = c.v[t];
}

size_t nms_workspace_bytes_host(const int64_t* image_off_host, int B, int C) {
  return nms_fixed_bytes(image_off_host[B], B, C) + 256;
}

cudaError_t launch_nms(const float* boxes, const float* scores, const int64_t* cls,
                       const int64_t* image_off_host, int B, int C, float score_thr,
                       double iou_thr, int topk, int64_t* keep, int32_t* keep_count, void* ws,
                       cudaStream_t stream) {
  const int64_t R = image_off_host[B];
  NmsWorkspace w = nms_carve(ws, R, B, C);
  for (int f = 0; f <= B; f += 32) {
    OffsetChunk c;
    const int cnt = (B + 1 - f) < 32 ? (B + 1 - f) : 32;
    for (int k = 0; k < 32; ++k) c.v[k] = k < cnt ? image_off_host[f + k] : 0;
    nms_set_offsets_kernel<<<1, 32, 0, stream>>>(w.off, f, cnt, c);
  }
  nms_sort_kernel<<<B, kNmsThreads, 0, stream>>>(boxes, scores, cls, score_thr, C, w);
  if (R > 0) {
    dim3 sgrid((unsigned)C, (unsigned)B);
    int64_t maxn = 0;
    for (int b = 0; b < B; ++b) {
      const int64_t nb = image_off_host[b + 1] - image_off_host[b];
      if (nb > maxn) maxn = nb;
    }
    // room for the largest class segment that can occur, capped at kSweepSmemBoxes boxes
    const size_t dyn = (size_t)(maxn < kSweepSmemBoxes ? maxn : kSweepSmemBoxes) * sizeof(float4);
    if (dyn > 12 * 1024)                               // (per device and context: set on every call)
      cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           kSweepSmemBoxes * (int)sizeof(float4));
    nms_sweep_kernel<<<sgrid, kSweepThreads, dyn, stream>>>(topk, C, iou_thr, w);
  }
  nms_merge_kernel<<<B, kNmsThreads, 0, stream>>>(topk, C, w, keep, keep_count);
  return cudaPeekAtLastError();
}

}  // namespace uwcv
