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
//   A  per image: the survivors of the score filter (the cost of every later step follows them,
//      as upstream's "scores > thresh" before batched_nms does) are bucketed by class
//      (shared-memory histogram, scan, scatter: the class segment table); per (image, class): the
//      segment's keys (~orderable(score), index) are sorted by their own CTA, in shared memory
//      up to 8 192 keys, and the boxes gathered in sorted order
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

// ---- A: filter + class buckets (per image), then one sort per (image, class) ------------------
// key of a candidate: class << 50 | ~orderable(score) << 18 | index.  The survivors of the score
// filter are counted per class (shared-memory histogram), the class segments laid out by an
// exclusive scan, and the keys scattered into their segment (unordered); every segment is then
// sorted by its own CTA -- the serial depth of the sort follows the largest class, not the image,
// and the sorts of all classes and images run side by side.
constexpr int kMaxClasses = 8192;
constexpr int kSegSortSmem = 8192;       // keys sorted entirely in shared memory (64 KB)

__device__ __forceinline__ bool nms_key_of(const float* __restrict__ boxes, const float* __restrict__ scores,
                                           const int64_t* __restrict__ cls, int64_t lo, int i, float score_thr,
                                           int C, uint64_t& key, int& c_out) {
  const float s = scores[lo + i];
  const float4 bx = reinterpret_cast<const float4*>(boxes)[lo + i];
  const long long c = cls[lo + i];
  const bool fin = isfinite(bx.x) && isfinite(bx.y) && isfinite(bx.z) && isfinite(bx.w) && isfinite(s);
  if (!(fin && s > score_thr && c >= 0 && c < C)) return false;
  key = ((uint64_t)c << kClsShift) | ((uint64_t)(~orderable(s)) << kIdxBits) | (uint32_t)i;
  c_out = (int)c;
  return true;
}

__global__ void __launch_bounds__(kNmsThreads)
nms_bucket_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                  const int64_t* __restrict__ cls, float score_thr, int C, NmsWorkspace w) {
  __shared__ int s_hist[kMaxClasses];      // count per class, then the scatter cursor
  __shared__ int s_part[kNmsThreads];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int64_t lo = w.off[b];
  const int n = (int)(w.off[b + 1] - lo);
  uint64_t* keys = w.keys + 2 * lo + b;
  for (int k = tid; k < C; k += kNmsThreads) { s_hist[k] = 0; w.ccount[(int64_t)b * C + k] = 0; }
  __syncthreads();
  for (int i = tid; i < n; i += kNmsThreads) {
    uint64_t key; int c;
    if (nms_key_of(boxes, scores, cls, lo, i, score_thr, C, key, c)) atomicAdd(&s_hist[c], 1);
  }
  __syncthreads();
  // exclusive scan over the classes: thread t owns classes [t * per, (t + 1) * per)
  const int per = (C + kNmsThreads - 1) / kNmsThreads;
  int mine = 0;
  for (int k = tid * per; k < min(C, (tid + 1) * per); ++k) mine += s_hist[k];
  s_part[tid] = mine;
  __syncthreads();
  for (int off = 1; off < kNmsThreads; off <<= 1) {
    int v = 0;
    if (tid >= off) v = s_part[tid - off];
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  int run = s_part[tid] - mine;
  for (int k = tid * per; k < min(C, (tid + 1) * per); ++k) {
    const int cnt = s_hist[k];
    w.seg[((int64_t)b * C + k) * 2] = cnt ? run : 0;
    w.seg[((int64_t)b * C + k) * 2 + 1] = cnt ? run + cnt : 0;
    s_hist[k] = run;                                   // from now on: the class's scatter cursor
    run += cnt;
  }
  if (tid == kNmsThreads - 1) w.nvalid[b] = s_part[kNmsThreads - 1];
  __syncthreads();
  for (int i = tid; i < n; i += kNmsThreads) {
    uint64_t key; int c;
    if (nms_key_of(boxes, scores, cls, lo, i, score_thr, C, key, c)) keys[atomicAdd(&s_hist[c], 1)] = key;
  }
}

// Bitonic network for ANY length m (no padding stored): the first pass of every stage pairs i with
// i ^ (k - 1), the following ones with i ^ j, all comparisons ascending; partners >= m are skipped.
__device__ __forceinline__ void bitonic_any(uint64_t* keys, int m, int tid, int nthreads) {
  int P = 1;
  while (P < m) P <<= 1;
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      const bool first = j == (k >> 1);
      for (int i = tid; i < m; i += nthreads) {
        const int l = first ? (i ^ (k - 1)) : (i ^ j);
        if (l > i && l < m) {
          const uint64_t a = keys[i], c = keys[l];
          if (a > c) { keys[i] = c; keys[l] = a; }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(kNmsThreads)
nms_segment_sort_kernel(const float* __restrict__ boxes, int C, NmsWorkspace w) {
  extern __shared__ uint64_t s_sort[];
  const int b = blockIdx.y, cls = blockIdx.x, tid = threadIdx.x;
  const int s = w.seg[((int64_t)b * C + cls) * 2], e = w.seg[((int64_t)b * C + cls) * 2 + 1];
  const int m = e - s;
  if (m <= 0) return;
  const int64_t lo = w.off[b];
  uint64_t* keys = w.keys + 2 * lo + b + s;
  if (m <= kSegSortSmem) {
    for (int i = tid; i < m; i += kNmsThreads) s_sort[i] = keys[i];
    __syncthreads();
    bitonic_any(s_sort, m, tid, kNmsThreads);
    for (int i = tid; i < m; i += kNmsThreads) keys[i] = s_sort[i];
  } else {
    __syncthreads();
    bitonic_any(keys, m, tid, kNmsThreads);            // (a class with more than 8 192 survivors)
  }
  __syncthreads();
  for (int r = tid; r < m; r += kNmsThreads) {
    const uint64_t key = m <= kSegSortSmem ? s_sort[r] : keys[r];
    const int li = (int)(key & ((1u << kIdxBits) - 1u));
    w.sbox[lo + s + r] = reinterpret_cast<const float4*>(boxes)[lo + li];
    w.scls[lo + s + r] = cls;
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
// The non-empty classes of the image (count, first kept slot, prefix of the counts) are staged in
// shared memory; one thread per kept box: its rank in the global score order is its rank in its own
// class plus, for every other non-empty class, the number of that class's kept boxes that precede
// it (binary search in the class's keep list, which is in score order).
constexpr int kMergeClasses = 2048;      // non-empty classes staged in shared memory

__global__ void __launch_bounds__(kNmsThreads)
nms_merge_kernel(int topk, int C, NmsWorkspace w, int64_t* __restrict__ keep,
                 int32_t* __restrict__ keep_count) {
  __shared__ int s_cnt[kMergeClasses], s_seg[kMergeClasses], s_pre[kMergeClasses + 1];
  __shared__ int s_ne, s_total;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int64_t lo = w.off[b];
  const uint64_t* keys = w.keys + 2 * lo + b;
  const uint64_t kScoreMask = (1ull << kClsShift) - 1ull;      // (~score, index): global order
  if (tid == 0) {                                              // compact the non-empty classes
    int ne = 0, tot = 0;
    for (int k = 0; k < C; ++k) {
      const int c = w.ccount[(int64_t)b * C + k];
      if (!c) continue;
      if (ne < kMergeClasses) { s_cnt[ne] = c; s_seg[ne] = w.seg[((int64_t)b * C + k) * 2]; s_pre[ne] = tot; }
      ++ne; tot += c;
    }
    s_ne = ne; s_total = tot;
    if (ne <= kMergeClasses) s_pre[ne] = tot;
    keep_count[b] = tot < topk ? tot : topk;
  }
  __syncthreads();
  const int ne = s_ne, total = s_total;
  if (total == 0) return;
  if (ne <= kMergeClasses) {
    for (int q = tid; q < total; q += kNmsThreads) {
      int a = 0, z = ne;                                       // class slot of kept box q: s_pre[e] <= q
      while (z - a > 1) { const int m = (a + z) >> 1; if (s_pre[m] <= q) a = m; else z = m; }
      const int e = a, j = q - s_pre[e];
      const uint64_t key = keys[w.ckeep[lo + s_seg[e] + j]];
      const uint64_t mine = key & kScoreMask;
      int rank = j;
      for (int e2 = 0; e2 < ne; ++e2) {
        if (e2 == e) continue;
        const int s2 = s_seg[e2];
        int x = 0, y = s_cnt[e2];                              // lower bound of `mine` in class e2
        while (x < y) {
          const int mid = (x + y) >> 1;
          if ((keys[w.ckeep[lo + s2 + mid]] & kScoreMask) < mine) x = mid + 1; else y = mid;
        }
        rank += x;
      }
      if (rank < topk) keep[lo + rank] = lo + (int64_t)(key & ((1u << kIdxBits) - 1u));
    }
    return;
  }
  // more non-empty classes than the staging area holds: the same ranking straight from global memory
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
        int a = 0, z = cnt2;
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
  nms_bucket_kernel<<<B, kNmsThreads, 0, stream>>>(boxes, scores, cls, score_thr, C, w);
  if (R > 0) {
    dim3 sgrid((unsigned)C, (unsigned)B);
    cudaFuncSetAttribute(nms_segment_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         kSegSortSmem * (int)sizeof(uint64_t));
    nms_segment_sort_kernel<<<sgrid, kNmsThreads, kSegSortSmem * sizeof(uint64_t), stream>>>(boxes, C, w);
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
