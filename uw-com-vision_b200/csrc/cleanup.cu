// Mask clean-up + RLE export on the bit tiles.
//
// Replaces postprocess_masks (nn_inference.py:265-306) and rle_encoding (:253-263):
//   for every instance of an image, in list (= score) order:
//     mask = binary_fill_holes(mask)                       scipy: background 4-connected to the border stays
//     mask = erosion(dilation(mask))                       skimage defaults: 3 x 3 cross, reflected border
//     overlap += mask; mask[overlap > 1] = 0               pixels an earlier cleaned mask covers are cut
//     if label(mask).max() > 1: mask[()] = 0               8-connected pieces; more than one -> emptied
//   EncodedPixels = column-major 1-based (start, length) runs of the result.
//
// Input: the tile workspace left by stages 1+2 of uwcv_paste_measure (plane M = thresholded
// mask bits of every instance, one word-aligned tile each).  Planes V and G are scratch here
// (the border-trace stage must not be expected to run on this workspace afterwards); the
// cleaned mask replaces M.  Everything is word-parallel bit arithmetic, one warp per instance:
//   phase A  reach = flood of the background from the tile's outer ring (4-connected, in place,
//            alternating sweep direction until nothing changes); filled = ~reach;
//            closing = erode(dilate(filled)) with the image border ignored       -> plane G
//   phase B  cut = G_i & ~(G_j of every earlier instance j of the image whose tile intersects);
//            flood of one 8-connected piece from the first set pixel; anything left over means
//            several pieces -> emptied; result -> plane M; run count of the column-major RLE
//   phase C  runs written at the offsets of an exclusive scan over the run counts.
// The tile of an instance has at least one background pixel of margin on every side that is
// not the image border (tile_geometry), so "outside the tile" is background connected to the
// image border: the tile ring is a complete seed set for the hole fill, and dilation never
// leaves the tile.
#include <climits>
#include "uwcv_common.cuh"

namespace uwcv {

constexpr unsigned kAll = 0xffffffffu;
constexpr int kCleanThreads = 128;                 // 4 instances (warps) per CTA

struct TileBits {
  int tw, th, wx0, y0, W, H;
  // bits of word column w that are pixels of the image (x < W)
  __device__ __forceinline__ uint32_t valid(int w) const {
    const int x = (wx0 + w) * 32;
    if (x + 32 <= W) return kAll;
    if (x >= W) return 0u;
    return (1u << (W - x)) - 1u;
  }
  __device__ __forceinline__ bool left_is_border() const { return wx0 == 0; }
  __device__ __forceinline__ bool right_is_border() const { return (wx0 + tw) * 32 >= W; }
  __device__ __forceinline__ bool top_is_border() const { return y0 == 0; }
  __device__ __forceinline__ bool bottom_is_border() const { return y0 + th >= H; }
};

// spread `x` over the runs of `free` that it touches (both directions), inside one word
__device__ __forceinline__ uint32_t hfill(uint32_t x, uint32_t free) {
  x &= free;
  uint32_t m = free, u = x;
  u |= m & (u << 1); m &= m << 1;
  u |= m & (u << 2); m &= m << 2;
  u |= m & (u << 4); m &= m << 4;
  u |= m & (u << 8); m &= m << 8;
  u |= m & (u << 16);
  m = free;
  uint32_t d = x;
  d |= m & (d >> 1); m &= m >> 1;
  d |= m & (d >> 2); m &= m >> 2;
  d |= m & (d >> 4); m &= m >> 4;
  d |= m & (d >> 8); m &= m >> 8;
  d |= m & (d >> 16);
  return u | d;
}

// words of the (volatile, in-place updated) reach plane: bypass L1 so that every lane of the
// warp sees the words the others stored in the previous step
__device__ __forceinline__ uint32_t ldr(const uint32_t* p) { return __ldcg(p); }
__device__ __forceinline__ void str(uint32_t* p, uint32_t v) { __stcg(p, v); }

// In-place flood of `reach` through `free` (kConn8: 8-connected, else 4-connected); kInvFree:
// free = ~src & valid (the background of src), else free = src.  One warp, Gauss-Seidel sweeps
// in alternating raster direction until a whole sweep changes nothing.
template <bool kConn8, bool kInvFree>
__device__ void flood(const TileBits& t, const uint32_t* __restrict__ src, uint32_t* reach, int lane) {
  const int total = t.tw * t.th;
  int dir = 0;
  for (;;) {
    bool changed = false;
    for (int base = 0; base < total; base += 32) {
      const int kk = base + lane;
      if (kk < total) {
        const int k = dir == 0 ? kk : total - 1 - kk;
        const int r = k / t.tw, w = k - r * t.tw;
        const uint32_t s = ldr(src + k);              // (may have been written by this kernel)
        const uint32_t free = kInvFree ? (~s & t.valid(w)) : s;
        if (free) {
          const uint32_t cur = ldr(reach + k);
          const bool hl = w > 0, hr = w + 1 < t.tw, hu = r > 0, hd = r + 1 < t.th;
          const uint32_t l = hl ? ldr(reach + k - 1) : 0u, rt = hr ? ldr(reach + k + 1) : 0u;
          const uint32_t up = hu ? ldr(reach + k - t.tw) : 0u, dn = hd ? ldr(reach + k + t.tw) : 0u;
          uint32_t nb = up | dn | (cur << 1) | (l >> 31) | (cur >> 1) | (rt << 31);
          if (kConn8) {
            const uint32_t ul = (hu && hl) ? ldr(reach + k - t.tw - 1) : 0u;
            const uint32_t ur = (hu && hr) ? ldr(reach + k - t.tw + 1) : 0u;
            const uint32_t dl = (hd && hl) ? ldr(reach + k + t.tw - 1) : 0u;
            const uint32_t dr = (hd && hr) ? ldr(reach + k + t.tw + 1) : 0u;
            nb |= (up << 1) | (ul >> 31) | (up >> 1) | (ur << 31);
            nb |= (dn << 1) | (dl >> 31) | (dn >> 1) | (dr << 31);
          }
          const uint32_t x = hfill(cur | (free & nb), free);
          if (x != cur) { str(reach + k, x); changed = true; }
        }
      }
      __syncwarp();
    }
    if (!__any_sync(kAll, changed)) break;
    dir ^= 1;
  }
}

__device__ __forceinline__ TileBits tile_of(const TileDesc& d, int H, int W) {
  TileBits t;
  t.tw = d.tw; t.th = d.th; t.wx0 = d.wx0; t.y0 = d.y0; t.W = W; t.H = H;
  return t;
}

// ---- column totals of the thresholded masks (the reference's keep_ind quirk, :277) ----------
__global__ void __launch_bounds__(kCleanThreads)
column_totals_kernel(int64_t n, int W, Workspace ws, const int32_t* __restrict__ image_idx,
                     int32_t* __restrict__ coltot) {
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * kCleanThreads + threadIdx.x) >> 5;
  if (i >= n) return;
  const TileDesc d = ws.desc[i];
  const uint32_t* M = ws.M + d.word_off;
  const int64_t img = image_idx ? image_idx[i] : 0;
  for (int w = 0; w < d.tw; ++w) {
    int cnt = 0;
    for (int r = 0; r < d.th; ++r) cnt += (__ldg(M + r * d.tw + w) >> lane) & 1u;
    const int x = (d.wx0 + w) * 32 + lane;
    if (cnt && x < W) atomicAdd(coltot + img * W + x, cnt);
  }
}

// ---- phase A: fill holes, close -----------------------------------------------------------------
__global__ void __launch_bounds__(kCleanThreads)
clean_close_kernel(int64_t n, int H, int W, Workspace ws, const int32_t* __restrict__ image_slot,
                   const int32_t* __restrict__ inst_idx, const int32_t* __restrict__ limit) {
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * kCleanThreads + threadIdx.x) >> 5;
  if (i >= n) return;
  const TileDesc d = ws.desc[i];
  const int total = d.tw * d.th;
  if (total == 0) return;
  if (limit && inst_idx[i] >= limit[image_slot[i]]) return;        // truncated away (:280)
  const TileBits t = tile_of(d, H, W);
  const uint32_t* M = ws.M + d.word_off;
  uint32_t* V = ws.V + d.word_off;
  uint32_t* G = ws.G + d.word_off;
  // seeds: background pixels on the tile's outer ring
  for (int k = lane; k < total; k += 32) {
    const int r = k / d.tw, w = k - r * d.tw;
    const uint32_t v = t.valid(w);
    uint32_t ring = 0u;
    if (r == 0 || r == d.th - 1) ring = v;
    if (w == 0) ring |= 1u;
    // last valid column of the tile
    if (v && (w == d.tw - 1 || t.valid(w + 1) == 0u)) ring |= 1u << (31 - __clz(v));
    str(V + k, ring & ~__ldg(M + k) & v);
  }
  __syncwarp();
  flood<false, true>(t, M, V, lane);
  // filled = everything the outside did not reach
  for (int k = lane; k < total; k += 32) {
    const int w = k % d.tw;
    G[k] = t.valid(w) & ~ldr(V + k);
  }
  __syncwarp();
  // dilation by the cross (never leaves the tile; neighbours outside the image do not exist)
  for (int k = lane; k < total; k += 32) {
    const int r = k / d.tw, w = k - r * d.tw;
    const uint32_t g = G[k];
    const uint32_t l = w > 0 ? G[k - 1] : 0u, rt = w + 1 < d.tw ? G[k + 1] : 0u;
    const uint32_t up = r > 0 ? G[k - d.tw] : 0u, dn = r + 1 < d.th ? G[k + d.tw] : 0u;
    str(V + k, (g | (g << 1) | (l >> 31) | (g >> 1) | (rt << 31) | up | dn) & t.valid(w));
  }
  __syncwarp();
  // erosion by the cross: a neighbour outside the IMAGE is ignored (reflected border repeats the
  // pixel itself), a neighbour outside the tile but inside the image is background
  const uint32_t out_l = t.left_is_border() ? kAll : 0u, out_r = t.right_is_border() ? kAll : 0u;
  const uint32_t out_u = t.top_is_border() ? kAll : 0u, out_d = t.bottom_is_border() ? kAll : 0u;
  for (int k = lane; k < total; k += 32) {
    const int r = k / d.tw, w = k - r * d.tw;
    const uint32_t v = t.valid(w);
    const uint32_t dd = ldr(V + k) | ~v;                          // columns past W: "ignored"
    const uint32_t l = w > 0 ? (ldr(V + k - 1) | ~t.valid(w - 1)) : out_l;
    const uint32_t rt = w + 1 < d.tw ? (ldr(V + k + 1) | ~t.valid(w + 1)) : out_r;
    const uint32_t up = r > 0 ? ldr(V + k - d.tw) : out_u;
    const uint32_t dn = r + 1 < d.th ? ldr(V + k + d.tw) : out_d;
    G[k] = dd & ((dd << 1) | (l >> 31)) & ((dd >> 1) | (rt << 31)) & up & dn & v;
  }
}

// ---- phase B: cut what earlier masks cover, keep single pieces, count runs --------------------
__global__ void __launch_bounds__(kCleanThreads)
clean_cut_kernel(int64_t n, int H, int W, Workspace ws, const int32_t* __restrict__ image_slot,
                 const int32_t* __restrict__ inst_idx, const int32_t* __restrict__ limit,
                 int32_t* __restrict__ flags, int64_t* __restrict__ area,
                 int64_t* __restrict__ nruns) {
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * kCleanThreads + threadIdx.x) >> 5;
  if (i >= n) return;
  const TileDesc d = ws.desc[i];
  const int total = d.tw * d.th;
  const int my_idx = inst_idx[i];
  const bool dropped = limit && my_idx >= limit[image_slot[i]];
  uint32_t* M = ws.M + d.word_off;
  if (total == 0 || dropped) {
    for (int k = lane; k < total; k += 32) M[k] = 0u;
    if (lane == 0) { flags[i] = dropped ? 2 : 0; area[i] = 0; nruns[i] = 0; }
    return;
  }
  const TileBits t = tile_of(d, H, W);
  uint32_t* V = ws.V + d.word_off;
  const uint32_t* G = ws.G + d.word_off;
  for (int k = lane; k < total; k += 32) { str(V + k, G[k]); M[k] = 0u; }
  __syncwarp();
  // earlier instances of the same image (list order = score order): i - my_idx .. i - 1
  const int64_t first = i - my_idx;
  for (int64_t jb = first; jb < i; jb += 32) {
    const int64_t j = jb + lane;
    bool hit = false;
    TileDesc e;
    e.wx0 = e.y0 = e.tw = e.th = 0; e.word_off = 0; e.row_off = 0;
    if (j < i) {
      e = ws.desc[j];
      hit = e.tw > 0 && e.wx0 < d.wx0 + d.tw && d.wx0 < e.wx0 + e.tw &&
            e.y0 < d.y0 + d.th && d.y0 < e.y0 + e.th;
    }
    unsigned hits = __ballot_sync(kAll, hit);
    while (hits) {
      const int src = __ffs(hits) - 1;
      hits &= hits - 1;
      const int ewx0 = __shfl_sync(kAll, e.wx0, src), ey0 = __shfl_sync(kAll, e.y0, src);
      const int etw = __shfl_sync(kAll, e.tw, src), eth = __shfl_sync(kAll, e.th, src);
      const int64_t eoff = __shfl_sync(kAll, e.word_off, src);
      const int wa = max(d.wx0, ewx0), wb = min(d.wx0 + d.tw, ewx0 + etw);
      const int ya = max(d.y0, ey0), yb = min(d.y0 + d.th, ey0 + eth);
      const int iw = wb - wa, cnt = iw * (yb - ya);
      const uint32_t* Ge = ws.G + eoff;
      for (int q = lane; q < cnt; q += 32) {
        const int rr = q / iw, ww = q - rr * iw;
        const uint32_t other = Ge[(ya + rr - ey0) * etw + (wa + ww - ewx0)];
        if (other) {
          uint32_t* p = V + (ya + rr - d.y0) * d.tw + (wa + ww - d.wx0);
          str(p, ldr(p) & ~other);
        }
      }
      __syncwarp();
    }
  }
  // one 8-connected piece from the raster-first pixel; leftovers mean several pieces
  int first_k = total;
  for (int base = 0; base < total && first_k == total; base += 32) {
    const int k = base + lane;
    const uint32_t v = k < total ? ldr(V + k) : 0u;
    const unsigned nz = __ballot_sync(kAll, v != 0u);
    if (nz) first_k = base + __ffs(nz) - 1;
  }
  bool multi = false;
  long long cnt_px = 0;
  if (first_k < total) {
    if (lane == 0) { const uint32_t v = ldr(V + first_k); str(M + first_k, v & (0u - v)); }
    __syncwarp();
    flood<true, false>(t, V, M, lane);
    bool left_over = false;
    for (int k = lane; k < total; k += 32) {
      const uint32_t v = ldr(V + k), m = ldr(M + k);
      left_over |= (v & ~m) != 0u;
      cnt_px += __popc(m);
    }
    multi = __any_sync(kAll, left_over);
  }
  __syncwarp();
  if (multi) {
    for (int k = lane; k < total; k += 32) str(M + k, 0u);
    cnt_px = 0;
  }
  __syncwarp();
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) cnt_px += __shfl_xor_sync(kAll, cnt_px, off);
  // column-major run count: lane = pixel column of word column w
  long long runs = 0;
  if (!multi && first_k < total) {
    for (int w = 0; w < d.tw; ++w) {
      uint32_t prev = 0u;
      for (int r = 0; r < d.th; ++r) {
        const uint32_t b = (ldr(M + r * d.tw + w) >> lane) & 1u;
        runs += b & ~prev;
        prev = b;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) runs += __shfl_xor_sync(kAll, runs, off);
  }
  if (lane == 0) { flags[i] = multi ? 1 : 0; area[i] = cnt_px; nruns[i] = runs; }
}

// ---- phase C: write the runs ---------------------------------------------------------------------
__global__ void __launch_bounds__(kCleanThreads)
rle_write_kernel(int64_t n, int H, int W, Workspace ws, const int64_t* __restrict__ run_off,
                 int64_t* __restrict__ runs) {
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * kCleanThreads + threadIdx.x) >> 5;
  if (i >= n) return;
  if (run_off[i + 1] == run_off[i]) return;
  const TileDesc d = ws.desc[i];
  const uint32_t* M = ws.M + d.word_off;
  int64_t base = run_off[i];
  for (int w = 0; w < d.tw; ++w) {
    // pass 1: runs in my column -> exclusive offsets over the 32 columns of this word column
    int mine = 0;
    uint32_t prev = 0u;
    for (int r = 0; r < d.th; ++r) {
      const uint32_t b = (__ldg(M + r * d.tw + w) >> lane) & 1u;
      mine += b & ~prev;
      prev = b;
    }
    int incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(kAll, incl, off);
      if (lane >= off) incl += v;
    }
    const int total_w = __shfl_sync(kAll, incl, 31);
    if (total_w == 0) continue;
    int64_t o = base + incl - mine;
    // pass 2: emit (start, length); flat index of pixel (x, y) is x * H + y, starts are 1-based
    const int64_t x = (int64_t)(d.wx0 + w) * 32 + lane;
    prev = 0u;
    int64_t start = 0;
    for (int r = 0; r < d.th; ++r) {
      const uint32_t b = (__ldg(M + r * d.tw + w) >> lane) & 1u;
      if (b & ~prev) start = x * H + d.y0 + r + 1;
      if (prev & ~b) { runs[2 * o] = start; runs[2 * o + 1] = x * H + d.y0 + r + 1 - start; ++o; }
      prev = b;
    }
    if (prev) { runs[2 * o] = start; runs[2 * o + 1] = x * H + d.y0 + d.th + 1 - start; ++o; }
    base += total_w;
  }
}

// ---- phase D: the EncodedPixels text on the device --------------------------------------------------
// conv = ' '.join(map(str, rle_encoding(mask)))  (nn_inference.py:317, :253-263).  A run that ends on
// the last image row and the next one starting on row 0 of the following column are consecutive
// flat indices, i.e. ONE run for the reference: run k is a continuation when it starts where run
// k - 1 of the same instance ended.  prep: characters each run contributes ("start length " for a
// head, nothing for a continuation); the caller scans them; write: decimal digits at those offsets.
__device__ __forceinline__ int dec_digits(long long v) {
  int d = 1;
  while (v >= 10) { v /= 10; ++d; }
  return d;
}
__device__ __forceinline__ bool rle_is_cont(const int64_t* runs, const int32_t* run_inst, int64_t k) {
  return k > 0 && run_inst[k] == run_inst[k - 1] && runs[2 * k] == runs[2 * (k - 1)] + runs[2 * (k - 1) + 1];
}
__device__ __forceinline__ long long rle_merged_len(const int64_t* runs, const int32_t* run_inst,
                                                    int64_t k, int64_t total) {
  long long len = runs[2 * k + 1];
  for (int64_t j = k + 1; j < total && rle_is_cont(runs, run_inst, j); ++j) len += runs[2 * j + 1];
  return len;
}

__global__ void __launch_bounds__(256)
rle_text_prep_kernel(int64_t total, const int64_t* __restrict__ runs, const int32_t* __restrict__ run_inst,
                     int64_t* __restrict__ chars) {
  const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (k >= total) return;
  long long c = 0;
  if (!rle_is_cont(runs, run_inst, k))
    c = dec_digits(runs[2 * k]) + 1 + dec_digits(rle_merged_len(runs, run_inst, k, total)) + 1;
  chars[k] = c;
}

__global__ void __launch_bounds__(256)
rle_text_write_kernel(int64_t total, const int64_t* __restrict__ runs, const int32_t* __restrict__ run_inst,
                      const int64_t* __restrict__ text_off, uint8_t* __restrict__ text) {
  const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (k >= total || rle_is_cont(runs, run_inst, k)) return;
  uint8_t* o = text + text_off[k];
  long long v[2] = {runs[2 * k], rle_merged_len(runs, run_inst, k, total)};
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int nd = dec_digits(v[q]);
    long long x = v[q];
    for (int p = nd - 1; p >= 0; --p) { o[p] = (uint8_t)('0' + (int)(x % 10)); x /= 10; }
    o[nd] = ' ';
    o += nd + 1;
  }
}

// ---- launchers -----------------------------------------------------------------------------------
static inline unsigned warps_grid(int64_t n) {
  return (unsigned)((n * 32 + kCleanThreads - 1) / kCleanThreads);
}

cudaError_t launch_column_totals(int64_t n, int W, const Workspace& ws, const int32_t* image_slot,
                                 int32_t* coltot, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  column_totals_kernel<<<warps_grid(n), kCleanThreads, 0, stream>>>(n, W, ws, image_slot, coltot);
  return cudaPeekAtLastError();
}

cudaError_t launch_clean(int64_t n, int H, int W, const Workspace& ws, const int32_t* image_slot,
                         const int32_t* inst_idx, const int32_t* limit, int32_t* flags,
                         int64_t* area, int64_t* nruns, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  clean_close_kernel<<<warps_grid(n), kCleanThreads, 0, stream>>>(n, H, W, ws, image_slot, inst_idx,
                                                                 limit);
  clean_cut_kernel<<<warps_grid(n), kCleanThreads, 0, stream>>>(n, H, W, ws, image_slot, inst_idx,
                                                               limit, flags, area, nruns);
  return cudaPeekAtLastError();
}

cudaError_t launch_rle_write(int64_t n, int H, int W, const Workspace& ws, const int64_t* run_off,
                             int64_t* runs, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  rle_write_kernel<<<warps_grid(n), kCleanThreads, 0, stream>>>(n, H, W, ws, run_off, runs);
  return cudaPeekAtLastError();
}

cudaError_t launch_rle_text_prep(int64_t total, const int64_t* runs, const int32_t* run_inst,
                                 int64_t* chars, cudaStream_t stream) {
  if (total == 0) return cudaSuccess;
  rle_text_prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(total, runs, run_inst, chars);
  return cudaPeekAtLastError();
}
cudaError_t launch_rle_text_write(int64_t total, const int64_t* runs, const int32_t* run_inst,
                                  const int64_t* text_off, uint8_t* text, cudaStream_t stream) {
  if (total == 0) return cudaSuccess;
  rle_text_write_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(total, runs, run_inst, text_off, text);
  return cudaPeekAtLastError();
}

}  // namespace uwcv

// ---- literal postprocess_masks entry: N x H x W bool masks in ---------------------------------
// (for callers that keep Detectron2's own paste and hand over pred_masks as the reference does,
//  nn_inference.py:325-327).  pixel_boxes: per instance the box [xmin, ymin, xmax + 1, ymax + 1]
//  of its set pixels as float32 XYXY (all zeros for an empty mask), which the ordinary layout
//  stage turns into a tile with a margin; pack_tiles then fills the tile plane M from the bytes.
namespace uwcv {

__global__ void __launch_bounds__(256)
pixel_boxes_kernel(const uint8_t* __restrict__ masks, int H, int W, float* __restrict__ boxes) {
  __shared__ int s_box[4];
  const int64_t i = blockIdx.x;
  if (threadIdx.x == 0) { s_box[0] = INT_MAX; s_box[1] = INT_MAX; s_box[2] = -1; s_box[3] = -1; }
  __syncthreads();
  const uint8_t* m = masks + i * (int64_t)H * W;
  int x0 = INT_MAX, y0 = INT_MAX, x1 = -1, y1 = -1;
  const int64_t total = (int64_t)H * W;
  const bool vec = (W % 16 == 0) && ((reinterpret_cast<uintptr_t>(m) & 15) == 0);
  if (vec) {
    const uint4* v = reinterpret_cast<const uint4*>(m);
    const int wq = W / 16;
    for (int64_t k = threadIdx.x; k < total / 16; k += blockDim.x) {
      const uint4 q = __ldg(v + k);
      if (q.x | q.y | q.z | q.w) {
        const int y = (int)(k / wq), xb = (int)(k - (int64_t)y * wq) * 16;
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
          if (w[a]) {
            const int lo = (__ffs(w[a]) - 1) >> 3, hi = (31 - __clz(w[a])) >> 3;
            x0 = min(x0, xb + 4 * a + lo); x1 = max(x1, xb + 4 * a + hi);
          }
        y0 = min(y0, y); y1 = max(y1, y);
      }
    }
  } else {
    for (int64_t k = threadIdx.x; k < total; k += blockDim.x)
      if (m[k]) {
        const int y = (int)(k / W), x = (int)(k - (int64_t)y * W);
        x0 = min(x0, x); x1 = max(x1, x); y0 = min(y0, y); y1 = max(y1, y);
      }
  }
  if (x1 >= 0) {
    atomicMin(&s_box[0], x0); atomicMin(&s_box[1], y0);
    atomicMax(&s_box[2], x1); atomicMax(&s_box[3], y1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const bool any = s_box[2] >= 0;
    boxes[4 * i + 0] = any ? (float)s_box[0] : 0.f;
    boxes[4 * i + 1] = any ? (float)s_box[1] : 0.f;
    boxes[4 * i + 2] = any ? (float)(s_box[2] + 1) : 0.f;
    boxes[4 * i + 3] = any ? (float)(s_box[3] + 1) : 0.f;
  }
}

// one warp per instance: lane = pixel of a 32-pixel tile word
__global__ void __launch_bounds__(kCleanThreads)
pack_tiles_kernel(int64_t n, const uint8_t* __restrict__ masks, int H, int W, Workspace ws) {
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * kCleanThreads + threadIdx.x) >> 5;
  if (i >= n) return;
  const TileDesc d = ws.desc[i];
  uint32_t* M = ws.M + d.word_off;
  const uint8_t* m = masks + i * (int64_t)H * W;
  const int total = d.tw * d.th;
  for (int k = 0; k < total; ++k) {
    const int r = k / d.tw, w = k - r * d.tw;
    const int x = (d.wx0 + w) * 32 + lane, y = d.y0 + r;
    const bool bit = x < W && m[(int64_t)y * W + x] != 0;
    const uint32_t word = __ballot_sync(kAll, bit);
    if (lane == 0) M[k] = word;
  }
}

// cleaned tiles -> N x H x W uint8 (0 / 1), the list the reference's postprocess_masks returns
__global__ void __launch_bounds__(kCleanThreads)
tiles_to_masks_kernel(int64_t n, int H, int W, Workspace ws, uint8_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * kCleanThreads + threadIdx.x) >> 5;
  if (i >= n) return;
  const TileDesc d = ws.desc[i];
  const uint32_t* M = ws.M + d.word_off;
  uint8_t* o = out + i * (int64_t)H * W;
  const int total = d.tw * d.th;
  for (int k = 0; k < total; ++k) {
    const uint32_t word = __ldg(M + k);
    if (!word) continue;
    const int r = k / d.tw, w = k - r * d.tw;
    const int x = (d.wx0 + w) * 32 + lane;
    if (x < W && ((word >> lane) & 1u)) o[(int64_t)(d.y0 + r) * W + x] = 1;
  }
}

cudaError_t launch_pixel_boxes(const uint8_t* masks, int64_t n, int H, int W, float* boxes,
                               cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  pixel_boxes_kernel<<<(unsigned)n, 256, 0, stream>>>(masks, H, W, boxes);
  return cudaPeekAtLastError();
}

cudaError_t launch_pack_tiles(const uint8_t* masks, int64_t n, int H, int W, const Workspace& ws,
                              cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  pack_tiles_kernel<<<warps_grid(n), kCleanThreads, 0, stream>>>(n, masks, H, W, ws);
  return cudaPeekAtLastError();
}

cudaError_t launch_tiles_to_masks(int64_t n, int H, int W, const Workspace& ws, uint8_t* out,
                                  cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  tiles_to_masks_kernel<<<warps_grid(n), kCleanThreads, 0, stream>>>(n, H, W, ws, out);
  return cudaPeekAtLastError();
}

}  // namespace uwcv
