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
//   A  per image: key = (class, ~orderable(score), index), bitonic sort (shared-memory
//      passes for strides < 4096), gather sorted boxes, class segment table
//   B  2-D grid of 64x64 IoU blocks over the sorted list (blocks without a common class are
//      skipped) -> upper-triangular suppression bit matrix
//   C  per (image, class): greedy sweep over the class segment, 64 candidates per step --
//      the serial part of NMS runs once per class in parallel
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
  int64_t* mask_off;   // [B + 1] offsets into mask (u64 words)
  uint64_t* mask;      // [sum n_b * ceil(n_b / 64)]
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
  s += align_up((size_t)(B + 1) * 8, 256);
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
  w.ccount = (int32_t*)p;   p += align_up((size_t)B * C * 4, 256);
  w.mask_off = (int64_t*)p; p += align_up((size_t)(B + 1) * 8, 256);
  w.mask = (uint64_t*)p;
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
  int P = 1;
  while (P < n) P <<= 1;
  uint64_t* keys = w.keys + 2 * lo + b;
  if (tid == 0) s_count = 0;
  for (int k = tid; k < C; k += kNmsThreads) {
    w.seg[((int64_t)b * C + k) * 2] = 0;
    w.seg[((int64_t)b * C + k) * 2 + 1] = 0;
    w.ccount[(int64_t)b * C + k] = 0;
  }
  __syncthreads();
  int local = 0;
  for (int i = tid; i < P; i += kNmsThreads) {
    uint64_t k = ~0ull;
    if (i < n) {
      const float s = scores[lo + i];
      const float4 bx = reinterpret_cast<const float4*>(boxes)[lo + i];
      const long long c = cls[lo + i];
      const bool fin = isfinite(bx.x) && isfinite(bx.y) && isfinite(bx.z) && isfinite(bx.w) &&
                       isfinite(s);
      if (fin && s > score_thr && c >= 0 && c < C) {
        k = ((uint64_t)c << kClsShift) | ((uint64_t)(~orderable(s)) << kIdxBits) | (uint32_t)i;
        ++local;
      }
    }
    keys[i] = k;
  }
  if (local) atomicAdd(&s_count, local);
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

// ---- B: suppression bit matrix ------------------------------------------------------
__global__ void __launch_bounds__(64)
nms_mask_kernel(double iou_thr, NmsWorkspace w) {
  const int b = blockIdx.z;
  const int nv = w.nvalid[b];
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (rb * 64 >= nv || cb * 64 >= nv || cb < rb) return;
  const int64_t lo = w.off[b];
  // the list is sorted by class: no pair to test when the column block starts in a later class
  // than the row block ends
  if (w.scls[lo + cb * 64] > w.scls[lo + min(rb * 64 + 63, nv - 1)]) return;
  const int nblk = (nv + 63) >> 6;
  __shared__ float4 s_box[64];
  __shared__ int s_cls[64];
  const int t = threadIdx.x;
  const int ncol = min(64, nv - cb * 64);
  if (t < ncol) {
    s_box[t] = w.sbox[lo + cb * 64 + t];
    s_cls[t] = w.scls[lo + cb * 64 + t];
  }
  __syncthreads();
  const int i = rb * 64 + t;
  if (i >= nv) return;
  const float4 bi = w.sbox[lo + i];
  const int ci = w.scls[lo + i];
  const float area_i = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
  uint64_t bits = 0;
  const int jstart = (rb == cb) ? t + 1 : 0;
  for (int j = jstart; j < ncol; ++j) {
    if (s_cls[j] != ci) continue;
    const float4 bj = s_box[j];
    const float area_j = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
    const float ww = fmaxf(0.f, __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)));
    const float hh = fmaxf(0.f, __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y)));
    const float inter = __fmul_rn(ww, hh);
    const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_i, area_j), inter));
    if ((double)iou > iou_thr) bits |= 1ull << j;
  }
  w.mask[w.mask_off[b] + (int64_t)i * nblk + cb] = bits;
}

// ---- C: greedy sweep of one class segment -------------------------------------------------
constexpr int kMaxBlocks = 4096;     // 262144 candidates per image
constexpr int kSweepThreads = 256;

__global__ void __launch_bounds__(kSweepThreads)
nms_sweep_kernel(int topk, int C, NmsWorkspace w) {
  const int b = blockIdx.y, cls = blockIdx.x, tid = threadIdx.x;
  const int s = w.seg[((int64_t)b * C + cls) * 2], e = w.seg[((int64_t)b * C + cls) * 2 + 1];
  if (e <= s) return;
  const int nv = w.nvalid[b];
  const int64_t lo = w.off[b];
  const int nblk = (nv + 63) >> 6;
  const int c_first = s >> 6, c_last = (e - 1) >> 6;
  const uint64_t* mask = w.mask + w.mask_off[b];
  __shared__ uint64_t s_removed[kMaxBlocks];
  __shared__ uint64_t s_diag[64];
  __shared__ uint64_t s_keepbits;
  __shared__ int s_kept;
  __shared__ int s_rows[64];
  for (int k = c_first + tid; k <= c_last; k += kSweepThreads) s_removed[k] = 0ull;
  if (tid == 0) s_kept = 0;
  __syncthreads();
  for (int c = c_first; c <= c_last; ++c) {
    const int nin = min(64, e - c * 64);     // positions >= e belong to the next class
    const int kept_before = s_kept;          // stable: last written before the previous barrier
    if (tid < nin) s_diag[tid] = mask[(int64_t)(c * 64 + tid) * nblk + c];
    __syncthreads();
    if (tid == 0) {
      uint64_t cur = s_removed[c], kb = 0;
      int kept = kept_before;
      const int ifirst = c == c_first ? (s & 63) : 0;          // positions < s: previous class
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
      s_rows[rank] = c * 64 + tid;
    }
    if (s_kept >= topk) break;
    __syncthreads();
    // OR the suppression rows of the boxes kept in this step into the words to the right:
    // (kept rows) x (remaining words of the segment) independent loads
    const int rem = c_last - c;
    const int items = nkept * rem;
    for (int it = tid; it < items; it += kSweepThreads) {
      const int r = it / rem, wd = c + 1 + (it - r * rem);
      const uint64_t m = mask[(int64_t)s_rows[r] * nblk + wd];
      if (m) atomicOr((unsigned long long*)&s_removed[wd], (unsigned long long)m);
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

// mask_off[b] = sum_{b' < b} n_b' * ceil(n_b' / 64)  (upper bound using all candidates)
__global__ void nms_offsets_kernel(int B, NmsWorkspace w) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int64_t acc = 0;
    for (int b = 0; b < B; ++b) {
      w.mask_off[b] = acc;
      const int64_t n = w.off[b + 1] - w.off[b];
      acc += n * ((n + 63) >> 6);
    }
    w.mask_off[B] = acc;
  }
}

size_t nms_workspace_bytes_host(const int64_t* image_off_host, int B, int C) {
  const int64_t R = image_off_host[B];
  size_t s = nms_fixed_bytes(R, B, C);
  size_t m = 0;
  for (int b = 0; b < B; ++b) {
    const int64_t n = image_off_host[b + 1] - image_off_host[b];
    m += (size_t)n * ((n + 63) >> 6);
  }
  return s + m * 8 + 256;
}

cudaError_t launch_nms(const float* boxes, const float* scores, const int64_t* cls,
                       const int64_t* image_off_host, int B, int C, float score_thr,
                       double iou_thr, int topk, int64_t* keep, int32_t* keep_count, void* ws,
                       cudaStream_t stream) {
  const int64_t R = image_off_host[B];
  NmsWorkspace w = nms_carve(ws, R, B, C);
  cudaError_t e = cudaMemcpyAsync(w.off, image_off_host, (size_t)(B + 1) * 8,
                                  cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return e;
  int64_t maxn = 0;
  for (int b = 0; b < B; ++b) {
    const int64_t n = image_off_host[b + 1] - image_off_host[b];
    if (n > maxn) maxn = n;
  }
  nms_offsets_kernel<<<1, 32, 0, stream>>>(B, w);
  nms_sort_kernel<<<B, kNmsThreads, 0, stream>>>(boxes, scores, cls, score_thr, C, w);
  if (maxn > 0) {
    const unsigned nb = (unsigned)((maxn + 63) / 64);
    dim3 grid(nb, nb, (unsigned)B);
    nms_mask_kernel<<<grid, 64, 0, stream>>>(iou_thr, w);
    dim3 sgrid((unsigned)C, (unsigned)B);
    nms_sweep_kernel<<<sgrid, kSweepThreads, 0, stream>>>(topk, C, w);
  }
  nms_merge_kernel<<<B, kNmsThreads, 0, stream>>>(topk, C, w, keep, keep_count);
  return cudaPeekAtLastError();
}

}  // namespace uwcv
