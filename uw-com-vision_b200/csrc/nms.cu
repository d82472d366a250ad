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
// Candidates are visited in (score descending, index ascending) order.
//
// Three launches for the whole batch:
//   A  per image: key = (~orderable(score), index), bitonic sort, gather sorted boxes
//   B  2-D grid of 64x64 IoU blocks -> upper-triangular suppression bit matrix
//   C  per image: greedy sweep over the bit matrix, 64 candidates per step
#include "uwcv_common.cuh"

namespace uwcv {

constexpr int kNmsThreads = 1024;

struct NmsWorkspace {
  int64_t* off;        // [B + 1] device copy of the per-image candidate offsets
  int32_t* nvalid;     // [B]
  uint64_t* keys;      // [2 R + B]
  float4* sbox;        // [R] boxes in sorted order (per image segment)
  int32_t* scls;       // [R]
  int32_t* sidx;       // [R] local candidate index in sorted order
  int64_t* mask_off;   // [B + 1] offsets into mask (u64 words)
  uint64_t* mask;      // [sum n_b * ceil(n_b / 64)]
};

__host__ __device__ inline size_t nms_fixed_bytes(int64_t R, int B) {
  size_t s = 0;
  s += align_up((size_t)(B + 1) * 8, 256);
  s += align_up((size_t)B * 4, 256);
  s += align_up((size_t)(2 * R + B) * 8, 256);
  s += align_up((size_t)R * 16, 256);
  s += align_up((size_t)R * 4, 256);
  s += align_up((size_t)R * 4, 256);
  s += align_up((size_t)(B + 1) * 8, 256);
  return s;
}

inline NmsWorkspace nms_carve(void* ws, int64_t R, int B) {
  NmsWorkspace w;
  char* p = (char*)ws;
  w.off = (int64_t*)p;      p += align_up((size_t)(B + 1) * 8, 256);
  w.nvalid = (int32_t*)p;   p += align_up((size_t)B * 4, 256);
  w.keys = (uint64_t*)p;    p += align_up((size_t)(2 * R + B) * 8, 256);
  w.sbox = (float4*)p;      p += align_up((size_t)R * 16, 256);
  w.scls = (int32_t*)p;     p += align_up((size_t)R * 4, 256);
  w.sidx = (int32_t*)p;     p += align_up((size_t)R * 4, 256);
  w.mask_off = (int64_t*)p; p += align_up((size_t)(B + 1) * 8, 256);
  w.mask = (uint64_t*)p;
  return w;
}

__device__ __forceinline__ uint32_t orderable(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ---- A: filter + sort ---------------------------------------------------------------
__global__ void __launch_bounds__(kNmsThreads)
nms_sort_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                const int64_t* __restrict__ cls, float score_thr, NmsWorkspace w) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const int64_t lo = w.off[b];
  const int n = (int)(w.off[b + 1] - lo);
  int P = 1;
  while (P < n) P <<= 1;
  uint64_t* keys = w.keys + 2 * lo + b;
  __shared__ int s_count;
  if (tid == 0) s_count = 0;
  __syncthreads();
  int local = 0;
  for (int i = tid; i < P; i += kNmsThreads) {
    uint64_t k = ~0ull;
    if (i < n) {
      const float s = scores[lo + i];
      const float4 bx = reinterpret_cast<const float4*>(boxes)[lo + i];
      const bool fin = isfinite(bx.x) && isfinite(bx.y) && isfinite(bx.z) && isfinite(bx.w) &&
                       isfinite(s);
      if (fin && s > score_thr) {
        k = ((uint64_t)(~orderable(s)) << 32) | (uint32_t)i;
        ++local;
      }
    }
    keys[i] = k;
  }
  if (local) atomicAdd(&s_count, local);
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P; i += kNmsThreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint64_t a = keys[i], c = keys[ixj];
          const bool asc = (i & k) == 0;
          if ((a > c) == asc) { keys[i] = c; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  const int nv = s_count;
  if (tid == 0) w.nvalid[b] = nv;
  for (int r = tid; r < nv; r += kNmsThreads) {
    const int li = (int)(keys[r] & 0xffffffffu);
    w.sidx[lo + r] = li;
    w.sbox[lo + r] = reinterpret_cast<const float4*>(boxes)[lo + li];
    w.scls[lo + r] = (int32_t)cls[lo + li];
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

// ---- C: greedy sweep ----------------------------------------------------------------
constexpr int kMaxBlocks = 4096;     // 262144 candidates per image

__global__ void __launch_bounds__(kNmsThreads)
nms_sweep_kernel(int topk, NmsWorkspace w, int64_t* __restrict__ keep,
                 int32_t* __restrict__ keep_count) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const int nv = w.nvalid[b];
  const int64_t lo = w.off[b];
  const int nblk = (nv + 63) >> 6;
  const uint64_t* mask = w.mask + w.mask_off[b];
  __shared__ uint64_t s_removed[kMaxBlocks];
  __shared__ uint64_t s_diag[64];
  __shared__ uint64_t s_keepbits;
  __shared__ int s_kept;
  for (int k = tid; k < nblk; k += kNmsThreads) s_removed[k] = 0ull;
  if (tid == 0) s_kept = 0;
  __syncthreads();
  for (int c = 0; c < nblk; ++c) {
    const int nin = min(64, nv - c * 64);
    const int kept_before = s_kept;          // stable: last written before the previous barrier
    if (tid < nin) s_diag[tid] = mask[(int64_t)(c * 64 + tid) * nblk + c];
    __syncthreads();
    if (tid == 0) {
      uint64_t cur = s_removed[c], kb = 0;
      int kept = kept_before;
      for (int i = 0; i < nin && kept < topk; ++i) {
        if (!((cur >> i) & 1ull)) { kb |= 1ull << i; ++kept; cur |= s_diag[i]; }
      }
      s_keepbits = kb;
      s_kept = kept;
    }
    __syncthreads();
    const uint64_t kb = s_keepbits;
    if (tid < 64 && ((kb >> tid) & 1ull)) {
      const int pos = kept_before + __popcll(kb & ((1ull << tid) - 1ull));
      keep[lo + pos] = lo + w.sidx[lo + c * 64 + tid];
    }
    if (s_kept >= topk) break;
    for (int wd = c + 1 + tid; wd < nblk; wd += kNmsThreads) {
      uint64_t acc = 0, bits = kb;
      while (bits) {
        const int i = __ffsll((long long)bits) - 1;
        bits &= bits - 1;
        acc |= mask[(int64_t)(c * 64 + i) * nblk + wd];
      }
      s_removed[wd] |= acc;
    }
    __syncthreads();
  }
  __syncthreads();
  if (tid == 0) keep_count[b] = s_kept;
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

size_t nms_workspace_bytes_host(const int64_t* image_off_host, int B) {
  const int64_t R = image_off_host[B];
  size_t s = nms_fixed_bytes(R, B);
  size_t m = 0;
  for (int b = 0; b < B; ++b) {
    const int64_t n = image_off_host[b + 1] - image_off_host[b];
    m += (size_t)n * ((n + 63) >> 6);
  }
  return s + m * 8 + 256;
}

cudaError_t launch_nms(const float* boxes, const float* scores, const int64_t* cls,
                       const int64_t* image_off_host, int B, float score_thr, double iou_thr,
                       int topk, int64_t* keep, int32_t* keep_count, void* ws,
                       cudaStream_t stream) {
  const int64_t R = image_off_host[B];
  NmsWorkspace w = nms_carve(ws, R, B);
  cudaError_t e = cudaMemcpyAsync(w.off, image_off_host, (size_t)(B + 1) * 8,
                                  cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return e;
  int64_t maxn = 0;
  for (int b = 0; b < B; ++b) {
    const int64_t n = image_off_host[b + 1] - image_off_host[b];
    if (n > maxn) maxn = n;
  }
  nms_offsets_kernel<<<1, 32, 0, stream>>>(B, w);
  nms_sort_kernel<<<B, kNmsThreads, 0, stream>>>(boxes, scores, cls, score_thr, w);
  if (maxn > 0) {
    const unsigned nb = (unsigned)((maxn + 63) / 64);
    dim3 grid(nb, nb, (unsigned)B);
    nms_mask_kernel<<<grid, 64, 0, stream>>>(iou_thr, w);
  }
  nms_sweep_kernel<<<B, kNmsThreads, 0, stream>>>(topk, w, keep, keep_count);
  return cudaPeekAtLastError();
}

}  // namespace uwcv
