// Kernel 0 (tile layout) and kernel 1+2 (fused paste / threshold / bit-pack / moments).
//
// Replaces, for the whole batch at once:
//   * detectron2/layers/mask_ops.py::paste_masks_in_image + _do_paste_mask
//     (F.grid_sample bilinear, zeros padding, align_corners=False, ">= threshold"),
//     reached from nn_inference.py:372 via detector_postprocess;
//   * the pixel-set reductions the reference gets from NumPy/OpenCV afterwards
//     (nn_inference.py:394-401 union paint; cv2.moments / bbox of each instance).
//
// Arithmetic contract (bit-exact vs torch 2.11 CPU grid_sample):
//   g  = ((p + 0.5f) - x0) / (x1 - x0) * 2 - 1          every op rounded separately
//   ix = fmaf(g + 1, 14, -0.5);  fx = floor(ix);  w = ix - fx;  e = 1 - w   (same for y: n, s)
//   out = fmaf(v_se, n*w, fmaf(v_sw, n*e, fmaf(v_ne, s*w, v_nw * (s*e))));  bit = out >= thr
// The file is compiled with -fmad=false; every FMA below is explicit.
//
// Data movement: the full-frame bit-plane of an instance is H rows of `wpr` 32-bit
// words.  Rows above / below the instance's tile are zero and are written with TMA
// bulk stores (cp.async.bulk shared->global) from a zeroed shared-memory buffer;
// the rows of the tile band are written with ordinary stores.  Algorithmic bytes per
// instance: 3136 (probabilities) + 16 (box) read, H*W/8 + rows written.
#include <cstdlib>
#include "contour_common.cuh"

namespace uwcv {

// rows-only contract: tiles up to this many words are pasted one instance per WARP
// (tile_measure_kernel), larger ones by a whole CTA (paste_measure_kernel<false>, only_big)
constexpr int kBigTileWords = 2048;

// ---------------------------------------------------------------------------------
// kernel 0: tile geometry + exclusive prefix sums (two passes over ceil(N / 1024) CTAs)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void tile_geometry(const float* __restrict__ box, int H, int W,
                                              int& wx0, int& y0, int& tw, int& th) {
  float x0 = box[0], yy0 = box[1], x1 = box[2], y1 = box[3];
  float bw = x1 - x0, bh = y1 - yy0;
  wx0 = y0 = tw = th = 0;
  if (!(bw > 0.f) || !(bh > 0.f)) return;                    // also rejects NaN
  if (!isfinite(x0) || !isfinite(x1) || !isfinite(yy0) || !isfinite(y1)) return;
  // support of the zero-padded bilinear kernel: p + 0.5 in (x0 - bw/56, x1 + bw/56);
  // one mask cell + one pixel of slack covers it for any positive threshold.
  double mx = (double)bw / kMaskSide + 1.0, my = (double)bh / kMaskSide + 1.0;
  double fa = floor((double)x0 - mx), fb = ceil((double)x1 + mx);
  double ga = floor((double)yy0 - my), gb = ceil((double)y1 + my);
  if (fb < 0 || ga > H - 1 || fa > W - 1 || gb < 0) return;
  int pxa = fa < 0 ? 0 : (int)fa;
  int pxb = fb > W - 1 ? W - 1 : (int)fb;
  int pya = ga < 0 ? 0 : (int)ga;
  int pyb = gb > H - 1 ? H - 1 : (int)gb;
  if (pxa > pxb || pya > pyb) return;
  wx0 = pxa >> 5;
  tw = (pxb >> 5) - wx0 + 1;
  y0 = pya;
  th = pyb - pya + 1;
}

// block-wide exclusive scan of two int64 values (Hillis-Steele over 1024 threads)
__device__ __forceinline__ void block_scan2(int64_t& a, int64_t& b, int64_t& total_a,
                                            int64_t& total_b, int64_t* sa, int64_t* sb) {
  const int t = threadIdx.x;
  const int64_t va = a, vb = b;
  sa[t] = va; sb[t] = vb;
  __syncthreads();
  for (int off = 1; off < kLayoutThreads; off <<= 1) {
    int64_t x = 0, y = 0;
    if (t >= off) { x = sa[t - off]; y = sb[t - off]; }
    __syncthreads();
    sa[t] += x; sb[t] += y;
    __syncthreads();
  }
  a = sa[t] - va; b = sb[t] - vb;
  total_a = sa[kLayoutThreads - 1]; total_b = sb[kLayoutThreads - 1];
}

// pass A: one instance per thread -- geometry, CTA-local exclusive offsets, CTA totals
__global__ void __launch_bounds__(kLayoutThreads)
layout_local_kernel(const float* __restrict__ boxes, int64_t n, int H, int W,
                    TileDesc* __restrict__ desc, int64_t* __restrict__ block_sums,
                    int64_t* __restrict__ big_tiles) {
  __shared__ int64_t s_words[kLayoutThreads];
  __shared__ int64_t s_rows[kLayoutThreads];
  const int64_t i = (int64_t)blockIdx.x * kLayoutThreads + threadIdx.x;
  int wx0 = 0, y0 = 0, tw = 0, th = 0;
  if (i < n) tile_geometry(boxes + 4 * i, H, W, wx0, y0, tw, th);
  int64_t words = (int64_t)tw * th, rows = th, tw_total, tr_total;
  block_scan2(words, rows, tw_total, tr_total, s_words, s_rows);
  if (i < n) {
    TileDesc d;
    d.wx0 = wx0; d.y0 = y0; d.tw = tw; d.th = th;
    d.word_off = words; d.row_off = rows;
    desc[i] = d;
  }
  if (threadIdx.x == 0) {
    block_sums[2 * blockIdx.x] = tw_total;
    block_sums[2 * blockIdx.x + 1] = tr_total;
  }
  // the raw moments are exact int64 sums: a tile whose worst case (every pixel set) could pass
  // 2^63 in m30 / m03 is refused instead of wrapping silently (full-frame masks above ~6 000 px)
  if (i < n && th > 0) {
    const double px = (double)tw * 32.0 * (double)th;
    const double xm = (double)(wx0 + tw) * 32.0, ym = (double)(y0 + th);
    const double worst = px * fmax(xm * xm * xm, ym * ym * ym);
    if (worst >= 9.0e18) atomicOr(reinterpret_cast<unsigned long long*>(block_sums + 2 * gridDim.x), 1ull);
  }
  // tiles too large for the one-instance-per-warp kernel of the rows-only contract
  const int big = __syncthreads_count(i < n && (int64_t)tw * th > kBigTileWords);
  if (threadIdx.x == 0 && big) atomicAdd(reinterpret_cast<unsigned long long*>(big_tiles),
                                         (unsigned long long)big);
}

// pass B: every CTA sums the totals of the CTAs before it and rebases its descriptors;
// CTA 0 also publishes the grand totals / overflow flag in the status word
__global__ void __launch_bounds__(kLayoutThreads)
layout_rebase_kernel(int64_t n, int nblk, TileDesc* __restrict__ desc,
                     const int64_t* __restrict__ block_sums, int64_t cap_words,
                     int64_t* __restrict__ status, unsigned int* __restrict__ sched) {
  __shared__ int64_t s_words[kLayoutThreads];
  __shared__ int64_t s_rows[kLayoutThreads];
  const int t = threadIdx.x;
  const int upto = blockIdx.x == 0 ? nblk : (int)blockIdx.x;     // CTA 0 needs the grand total
  int64_t a = 0, b = 0;
  for (int k = t; k < upto; k += kLayoutThreads) { a += block_sums[2 * k]; b += block_sums[2 * k + 1]; }
  s_words[t] = a; s_rows[t] = b;
  __syncthreads();
  for (int off = kLayoutThreads / 2; off > 0; off >>= 1) {
    if (t < off) { s_words[t] += s_words[t + off]; s_rows[t] += s_rows[t + off]; }
    __syncthreads();
  }
  const int64_t base_w = s_words[0], base_r = s_rows[0];
  if (blockIdx.x == 0) {
    if (t == 0) {
      status[1] = base_w;
      status[2] = base_r;
      status[0] = block_sums[2 * nblk] ? (int64_t)E_TOO_LARGE
                                       : ((base_w > cap_words) ? (int64_t)E_CAPACITY : 0);
      for (int k = 0; k < 8; ++k) sched[k] = 0u;
    }
    return;                                              // offsets of CTA 0 need no rebase
  }
  const int64_t i = (int64_t)blockIdx.x * kLayoutThreads + t;
  if (i < n) {
    desc[i].word_off += base_w;
    desc[i].row_off += base_r;
  }
}

// ---------------------------------------------------------------------------------
// kernel 1+2: paste + threshold + bit-pack + raw moments / bbox
// ---------------------------------------------------------------------------------
constexpr int kPasteThreads = 256;
constexpr int kPasteWarps = 7;                       // compute warps; warp 7 only issues TMA fills
constexpr int kComputeThreads = kPasteWarps * 32;    // 224
__device__ __forceinline__ void compute_barrier() {  // named barrier over the compute warps
  asm volatile("bar.sync 1, %0;" ::"n"(kComputeThreads) : "memory");
}
constexpr int kZeroBytesDefault = 16384;     // shared zero source for the bulk stores

__device__ __forceinline__ void bulk_store_zero(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_zero_hint(void* gdst, uint32_t smem_src, uint32_t bytes,
                                                     uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
               :: "l"(gdst), "r"(smem_src), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_all() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// coordinate of pixel index p along one axis -> (padded tap base, weight hi, weight lo)
// base indexes the zero-framed 32-wide mask copy: taps are base and base + 1.
__device__ __forceinline__ void axis_coord(int p, float lo, float hi, int& base, float& w_hi,
                                           float& w_lo) {
  float g = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(__fadd_rn((float)p, 0.5f), lo),
                                          __fsub_rn(hi, lo)), 2.0f), 1.0f);
  float i = __fmaf_rn(__fadd_rn(g, 1.0f), (float)kMaskSide * 0.5f, -0.5f);
  float f = floorf(i);
  w_hi = __fsub_rn(i, f);          // weight of the tap at f + 1  ("w" / "n")
  w_lo = __fsub_rn(1.0f, w_hi);    // weight of the tap at f      ("e" / "s")
  // taps outside [0, 27] read zeros from the frame; clamp in the float domain so that
  // huge / non-finite coordinates never reach a float->int conversion out of range.
  float fc = fminf(fmaxf(f, -(float)kPad), (float)kMaskSide);
  base = (int)fc + kPad;           // in [0, kMaskSide + kPad] = [0, 30]; taps base, base+1 <= 31
}

// torch's CUDA sigmoid (ATen UnaryOpsKernel: 1 / (1 + std::exp(-a)) in float): accurate expf,
// IEEE add and divide
__device__ __forceinline__ float sigmoid_as_torch(float a) {
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-a)));
}

// Raw moments / pixel bbox accumulated by one lane (reduced over the warp afterwards)
struct TileAcc {
  long long m00 = 0, m10 = 0, m01 = 0, m20 = 0, m11 = 0, m02 = 0, m30 = 0, m21 = 0, m12 = 0, m03 = 0;
  int xmin = INT_MAX, xmax = -1, ymin = INT_MAX, ymax = -1;
};

// One warp pastes the 32 tile rows starting at rbase: lane = column inside each 32-pixel strip,
// rows broadcast by shuffle, __ballot_sync packs a row of 32 pixels into a word; the same bits
// feed the per-lane column sums that fold into the raw moments.
template <bool kPlanes>
__device__ __forceinline__ void paste_rows(const TileDesc& d, int rbase, const float* __restrict__ mk,
                                           float bx0, float by0, float bx1, float by1, float thr,
                                           int W, int wpr, uint32_t* __restrict__ plane,
                                           uint32_t* __restrict__ tM, int lane, TileAcc& a) {
  int rowb; float rn, rs;
  axis_coord(d.y0 + rbase + lane, by0, by1, rowb, rn, rs);
  rowb *= kMaskPitch;
  const int nrows = min(32, d.th - rbase);
  for (int strip = 0; strip < d.tw; ++strip) {
    const int px = (d.wx0 + strip) * 32 + lane;
    int colb; float cw, ce;
    axis_coord(px, bx0, bx1, colb, cw, ce);
    const bool col_ok = px < W;
    int S0 = 0, S1 = 0, S2 = 0, S3 = 0;
    uint32_t myword = 0;
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
      const int rb = __shfl_sync(0xffffffffu, rowb, j);
      const float n_ = __shfl_sync(0xffffffffu, rn, j);
      const float s_ = __shfl_sync(0xffffffffu, rs, j);
      const float* mp = mk + rb + colb;
      const float v_nw = mp[0], v_ne = mp[1], v_sw = mp[kMaskPitch], v_se = mp[kMaskPitch + 1];
      const float nw = __fmul_rn(s_, ce), ne = __fmul_rn(s_, cw);
      const float sw = __fmul_rn(n_, ce), se = __fmul_rn(n_, cw);
      float out = __fmul_rn(v_nw, nw);
      out = __fmaf_rn(v_ne, ne, out);
      out = __fmaf_rn(v_sw, sw, out);
      out = __fmaf_rn(v_se, se, out);
      const bool bit = (out >= thr) && col_ok && (j < nrows);
      const uint32_t word = __ballot_sync(0xffffffffu, bit);
      if (lane == j) myword = word;
      const int b = bit ? 1 : 0;
      S0 += b; S1 += b * j; S2 += b * j * j; S3 += b * j * j * j;
    }
    // lane r holds the word of row rbase + r
    if (lane < nrows) {
      const int64_t o = (int64_t)(rbase + lane) * d.tw + strip;
      UWCV_BOUND(o, (int64_t)d.tw * d.th);
      UWCV_BOUND(d.y0 + rbase + lane, 32768); UWCV_BOUND(d.wx0 + strip, wpr);
      tM[o] = myword;
      if (kPlanes) plane[(int64_t)(d.y0 + rbase + lane) * wpr + d.wx0 + strip] = myword;
      if (myword) { a.ymin = min(a.ymin, d.y0 + rbase + lane); a.ymax = max(a.ymax, d.y0 + rbase + lane); }
    }
    if (S0) {
      // column sums over the 32 rows, shifted to frame coordinates (exact integers)
      const long long yb = d.y0 + rbase, x = px;
      const long long T0 = S0;
      const long long T1 = S1 + yb * S0;
      const long long T2 = S2 + 2 * yb * S1 + yb * yb * S0;
      const long long T3 = S3 + 3 * yb * S2 + 3 * yb * yb * S1 + yb * yb * yb * S0;
      const long long x2 = x * x, x3 = x2 * x;
      a.m00 += T0; a.m10 += x * T0; a.m20 += x2 * T0; a.m30 += x3 * T0;
      a.m01 += T1; a.m11 += x * T1; a.m21 += x2 * T1;
      a.m02 += T2; a.m12 += x * T2;
      a.m03 += T3;
      a.xmin = min(a.xmin, px); a.xmax = max(a.xmax, px);
    }
  }
}

// mask source of one instance staged into a zero-framed 32 x 32 copy by `nthreads` threads
__device__ __forceinline__ void stage_mask(const float* __restrict__ masks, const int64_t* __restrict__ classes,
                                           int64_t inst, const MaskSource& src, float* __restrict__ mk,
                                           int t, int nthreads) {
  // probabilities as given, or (single-forward path) the predicted class's channel of the
  // mask head's logits with the sigmoid of mask_rcnn_inference applied on the way in
  int64_t chan = 0;
  if (src.channels > 1) chan = (classes ? classes[inst] : 0) + src.channel_offset;
  const bool chan_ok = chan >= 0 && chan < src.channels;       // bad class id: empty mask
  const float* msrc = masks + inst * src.stride + (chan_ok ? chan : 0) * (kMaskSide * kMaskSide);
  for (int k = t; k < kMaskSide * kMaskSide; k += nthreads) {
    int r = k / kMaskSide, c = k - r * kMaskSide;
    float v = chan_ok ? __ldg(msrc + k) : 0.f;
    if (src.logits) v = chan_ok ? sigmoid_as_torch(v) : 0.f;
    mk[(r + kPad) * kMaskPitch + c + kPad] = v;
  }
}

template <bool kPlanes>
__global__ void __launch_bounds__(kPasteThreads, 3)
paste_measure_kernel(const float* __restrict__ masks, const float* __restrict__ boxes,
                     const int32_t* __restrict__ image_idx, const int32_t* __restrict__ inst_idx,
                     const int64_t* __restrict__ classes, int64_t n, int H, int W, float thr,
                     uint32_t* __restrict__ planes, int64_t* __restrict__ rows_i,
                     Workspace ws, const int64_t* __restrict__ status, int zero_bytes,
                     int rot_mul, int fill_mode, int debug_skip, int64_t first, MaskSource src,
                     int only_big, int band_bytes_cap, const int32_t* __restrict__ order) {
  extern __shared__ __align__(128) unsigned char s_zero[];     // zero_bytes + band buffer (planes only)
  __shared__ __align__(16) float s_mask[kMaskPitch * kMaskPitch];
  __shared__ unsigned long long s_acc[10];
  __shared__ int s_bbox[4];
  __shared__ long long s_next[2];
  constexpr int kThreads = kPasteThreads;

  if (status[0] != 0) return;                       // layout overflowed the workspace
  if (only_big && status[3] == 0) return;           // nothing left for the whole-CTA pass

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wpr = plane_row_words(W);
  const int64_t plane_words = (int64_t)H * wpr;

  if (kPlanes) {
    for (int k = tid; k < zero_bytes / 16; k += kThreads)
      reinterpret_cast<uint4*>(s_zero)[k] = make_uint4(0, 0, 0, 0);
    // band buffer behind the zero source: the tile band of one instance is composed here
    // (zeros + tile words) and leaves through a TMA bulk store as well
    for (int k = tid; k < band_bytes_cap / 16; k += kThreads)
      reinterpret_cast<uint4*>(s_zero + zero_bytes)[k] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();                       // generic-proxy zeros -> visible to TMA
  }
  uint32_t* s_band = reinterpret_cast<uint32_t*>(s_zero + zero_bytes);
  // the frame of the padded mask stays zero for the whole kernel
  for (int k = tid; k < kMaskPitch * kMaskPitch; k += kThreads) s_mask[k] = 0.f;
  __syncthreads();
  const uint32_t zero_smem = (uint32_t)__cvta_generic_to_shared(s_zero);

  // ---- warp 7: zero rows above / below every band of this CTA's instances ---------------
  // One thread streams TMA bulk stores (shared zero buffer -> plane) for the whole instance
  // list, decoupled from the compute warps: the two never touch the same bytes, so no
  // ordering is needed and the fill runs at the speed of the TMA / HBM write path while the
  // other seven warps paste, pack and reduce.
  if (warp == kPasteWarps) {
    if (kPlanes && lane == 0 && (fill_mode == 0 || fill_mode == 2)) {
      // fill_mode 2: the zero rows are marked evict-first in L2, so that the stream of plane
      // bytes does not push the tile workspace (read by the border trace) out of the cache
      uint64_t policy = 0;
      if (fill_mode == 2)
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      // planes are claimed one at a time from a device-wide counter: under a saturated write
      // path the SMs do not drain at the same rate, and a static split leaves the fast ones
      // idle at the end (tools/fill_bench2.cu: 6.3 TB/s static vs 7.6 TB/s dynamic)
      int64_t inst = first + atomicAdd(&ws.sched[0], 1u);
      while (inst < n) {
        const int64_t next = first + atomicAdd(&ws.sched[0], 1u);   // in flight under the fill
        const TileDesc d = ws.desc[inst];
        char* base = reinterpret_cast<char*>(planes + inst * plane_words);
        const int band_lo = d.th > 0 ? d.y0 : H;            // empty tile: whole plane is zero
        const int band_hi = d.th > 0 ? d.y0 + d.th : H;
        const int64_t seg_lo[2] = {0, (int64_t)band_hi * wpr * 4};
        const int64_t seg_hi[2] = {(int64_t)band_lo * wpr * 4, plane_words * 4};
        for (int sgi = 0; sgi < 2; ++sgi)
          for (int64_t o = seg_lo[sgi]; o < seg_hi[sgi]; o += zero_bytes) {
            const int64_t rem = seg_hi[sgi] - o;
            const uint32_t nbytes = (uint32_t)(rem < zero_bytes ? rem : zero_bytes);
            UWCV_BOUND(o + nbytes, plane_words * 4 + 1); UWCV_BOUND(inst, n);
            if (fill_mode == 2) bulk_store_zero_hint(base + o, zero_smem, nbytes, policy);
            else bulk_store_zero(base + o, zero_smem, nbytes);
            if (rot_mul > 0) {                       // bounded number of bulk stores in flight
              bulk_commit();
              if (rot_mul == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              else if (rot_mul == 2) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
              else if (rot_mul == 4) asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
            }
          }
        bulk_commit();
        inst = next;
      }
      bulk_wait_read_all();                            // the zero source must outlive the copies
    }
  } else {

  auto claim_next = [&]() -> long long {
    const long long k = (long long)atomicAdd(&ws.sched[1], 1u);
    return (order && first + k < n) ? (long long)order[first + k] : first + k;
  };
  if (tid == 0) s_next[0] = claim_next();
  compute_barrier();
  for (int it = 0;; ++it) {
    const int64_t inst = s_next[it & 1];
    if (inst >= n) break;
    // the claim of the next instance travels under this one's work: the counter's reply is
    // only consumed (stored to shared memory) before the last barrier of the iteration
    long long claim = 0;
    if (tid == 0) claim = claim_next();
    const TileDesc d = ws.desc[inst];
    if (only_big && (int64_t)d.tw * d.th <= kBigTileWords) {     // done by tile_measure_kernel
      if (tid == 0) s_next[(it + 1) & 1] = claim;
      compute_barrier();
      continue;
    }
    const float bx0 = boxes[4 * inst + 0], by0 = boxes[4 * inst + 1];
    const float bx1 = boxes[4 * inst + 2], by1 = boxes[4 * inst + 3];
    uint32_t* plane = kPlanes ? planes + inst * plane_words : nullptr;
    // fill_mode 1 (profiling): zero rows by 16-byte LSU stores from the compute warps
    if (kPlanes && fill_mode == 1) {
      const int band_lo = d.th > 0 ? d.y0 : H;
      const int band_hi = d.th > 0 ? d.y0 + d.th : H;
      uint4* base = reinterpret_cast<uint4*>(plane);
      const int64_t hi0 = (int64_t)band_lo * wpr / 4;
      const int64_t lo1 = (int64_t)band_hi * wpr / 4, hi1 = plane_words / 4;
      const uint4 z = make_uint4(0, 0, 0, 0);
      for (int64_t o = tid; o < hi0; o += kComputeThreads) __stcs(base + o, z);
      for (int64_t o = lo1 + tid; o < hi1; o += kComputeThreads) __stcs(base + o, z);
    }

    // ---- zero the band rows: whole rows as 16-byte stores (complete sectors); the tile words
    //      are stored over them after the barrier below
    const int band_bytes = d.th * wpr * 4;
    const bool band_tma = kPlanes && d.th > 0 && band_bytes <= band_bytes_cap;
    if (kPlanes && d.th > 0 && !band_tma && !(debug_skip & 2)) {
      uint4* band = reinterpret_cast<uint4*>(plane + (int64_t)d.y0 * wpr);
      const int total = d.th * (wpr / 4);
      for (int k = tid; k < total; k += kComputeThreads) band[k] = make_uint4(0, 0, 0, 0);
    }

    // ---- stage the 28x28 probabilities into the zero-framed copy --------------------
    stage_mask(masks, classes, inst, src, s_mask, tid, kComputeThreads);
    if (tid < 10) s_acc[tid] = 0ull;
    if (tid == 0) { s_bbox[0] = INT_MAX; s_bbox[1] = INT_MAX; s_bbox[2] = -1; s_bbox[3] = -1; }
    compute_barrier();

    // ---- the tile: each warp takes groups of 32 rows --------------------------------
    UWCV_BOUND(d.word_off + (int64_t)d.tw * d.th, ws.cap_words + 1);
    UWCV_BOUND(d.y0 + d.th, H + 1); UWCV_BOUND(d.wx0 + d.tw, wpr + 1);
    if (band_tma) UWCV_BOUND((int64_t)d.th * wpr * 4, (int64_t)band_bytes_cap + 1);
    uint32_t* tM = ws.M + d.word_off;

    TileAcc ta;
    for (int g = warp; g * 32 < d.th && !(debug_skip & 1); g += kPasteWarps)
      paste_rows<kPlanes>(d, g * 32, s_mask, bx0, by0, bx1, by1, thr, W, wpr,
                          // band composed in shared memory: row y of the plane is row y - y0 of it
                          band_tma ? s_band - (int64_t)d.y0 * wpr : plane, tM, lane, ta);
    if (band_tma) fence_proxy_async_smem();           // my tile words -> visible to the TMA
    const long long m00 = ta.m00, m10 = ta.m10, m01 = ta.m01, m20 = ta.m20, m11 = ta.m11, m02 = ta.m02,
                    m30 = ta.m30, m21 = ta.m21, m12 = ta.m12, m03 = ta.m03;
    int xmin = ta.xmin, xmax = ta.xmax, ymin = ta.ymin, ymax = ta.ymax;
    // ---- warp-shuffle reduction, one shared-memory atomic per warp and moment -------
    long long acc[10] = {m00, m10, m01, m20, m11, m02, m30, m21, m12, m03};
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      long long v = acc[k];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0 && v != 0) atomicAdd(&s_acc[k], (unsigned long long)v);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, off));
      ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, off));
      xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, off));
      ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, off));
    }
    if (lane == 0 && xmax >= 0) {
      atomicMin(&s_bbox[0], xmin); atomicMin(&s_bbox[1], ymin);
      atomicMax(&s_bbox[2], xmax); atomicMax(&s_bbox[3], ymax);
    }
    compute_barrier();
    // ---- integer part of the row --------------------------------------------------------
    if (tid < kNumInt) {
      long long v = 0;
      const long long area = (long long)s_acc[0];
      switch (tid) {
        case I_IMAGE: v = image_idx ? image_idx[inst] : 0; break;
        case I_INST:  v = inst_idx ? inst_idx[inst] : inst; break;
        case I_CLASS: v = classes ? classes[inst] : 0; break;
        case I_VALID: v = area > 0; break;
        case I_NCONT: v = 0; break;
        case I_AREA:  v = area; break;
        case I_BX0: v = area > 0 ? s_bbox[0] : -1; break;
        case I_BY0: v = area > 0 ? s_bbox[1] : -1; break;
        case I_BX1: v = area > 0 ? s_bbox[2] : -1; break;
        case I_BY1: v = area > 0 ? s_bbox[3] : -1; break;
        case I_NPTS: v = 0; break;
        default: v = (long long)s_acc[tid - I_M10 + 1]; break;     // m10 .. m03
      }
      rows_i[inst * kNumInt + tid] = v;
    }
    if (tid == 0) s_next[(it + 1) & 1] = claim;
    if (band_tma) {
      // (the barrier after the reduction ordered every warp's tile words before this point)
      if (tid == 0) {
        bulk_store_zero(plane + (int64_t)d.y0 * wpr, (uint32_t)__cvta_generic_to_shared(s_band),
                        (uint32_t)band_bytes);
        bulk_commit();
        bulk_wait_read_all();                          // the buffer is reused for the next band
      }
      compute_barrier();
      for (int k = tid; k < d.th * d.tw; k += kComputeThreads) {     // back to all zeros
        const int r = k / d.tw, c = k - r * d.tw;
        UWCV_BOUND((r * wpr + d.wx0 + c) * 4, band_bytes_cap);
        s_band[r * wpr + d.wx0 + c] = 0u;
      }
      fence_proxy_async_smem();
    }
    compute_barrier();                                 // s_mask / s_acc are rewritten next round
  }
  }
  // the last CTA out re-arms the counters for the next launch on this workspace
  __syncthreads();
  if (tid == 0) {
    const unsigned int done = atomicAdd(&ws.sched[2], 1u);
    if (done == gridDim.x - 1) { ws.sched[0] = 0u; ws.sched[1] = 0u; ws.sched[2] = 0u; }
  }
}

// ---------------------------------------------------------------------------------
// rows-only contract (no full-frame planes): one instance per WARP
// ---------------------------------------------------------------------------------
// With a whole CTA per instance only ceil(th / 32) of the seven compute warps have rows to work
// on (two for a 60-row tile) and every instance pays three CTA barriers: ~12 us of latency per
// CTA and instance, 1.7 ms per 64 000 instances.  Here every warp claims instances on its own
// (device-wide counter), stages the mask in its own zero-framed copy, pastes / packs / reduces and
// writes the integer row without any CTA-level synchronisation: 0.75 ms.  Tiles above
// kBigTileWords are left to the whole-CTA kernel.
constexpr int kTileWarps = 8;

__global__ void __launch_bounds__(kTileWarps * 32, 3)
tile_measure_kernel(const float* __restrict__ masks, const float* __restrict__ boxes,
                    const int32_t* __restrict__ image_idx, const int32_t* __restrict__ inst_idx,
                    const int64_t* __restrict__ classes, int64_t n, int H, int W, float thr,
                    int64_t* __restrict__ rows_i, Workspace ws, const int64_t* __restrict__ status,
                    int64_t first, MaskSource src) {
  __shared__ __align__(16) float s_mask[kTileWarps * kMaskPitch * kMaskPitch];
  if (status[0] != 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wpr = plane_row_words(W);
  for (int k = tid; k < kTileWarps * kMaskPitch * kMaskPitch; k += kTileWarps * 32) s_mask[k] = 0.f;
  __syncthreads();
  float* mk = s_mask + warp * (kMaskPitch * kMaskPitch);
  long long c0 = 0;
  if (lane == 0) c0 = first + atomicAdd(&ws.sched[1], 1u);
  int64_t inst = __shfl_sync(0xffffffffu, c0, 0);
  while (inst < n) {
    long long claim = 0;                               // travels under this instance's work
    if (lane == 0) claim = first + atomicAdd(&ws.sched[1], 1u);
    const TileDesc d = ws.desc[inst];
    if ((int64_t)d.tw * d.th <= kBigTileWords) {
      const float bx0 = boxes[4 * inst + 0], by0 = boxes[4 * inst + 1];
      const float bx1 = boxes[4 * inst + 2], by1 = boxes[4 * inst + 3];
      stage_mask(masks, classes, inst, src, mk, lane, 32);
      __syncwarp();
      TileAcc a;
      uint32_t* tM = ws.M + d.word_off;
      for (int rbase = 0; rbase < d.th; rbase += 32)
        paste_rows<false>(d, rbase, mk, bx0, by0, bx1, by1, thr, W, wpr, nullptr, tM, lane, a);
      long long acc[10] = {a.m00, a.m10, a.m01, a.m20, a.m11, a.m02, a.m30, a.m21, a.m12, a.m03};
#pragma unroll
      for (int k = 0; k < 10; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
      }
      int xmin = a.xmin, xmax = a.xmax, ymin = a.ymin, ymax = a.ymax;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, off));
        ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, off));
        xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, off));
        ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, off));
      }
      if (lane < kNumInt) {                            // lane = column of the integer row
        long long v = 0;
        const long long area = acc[0];
        switch (lane) {
          case I_IMAGE: v = image_idx ? image_idx[inst] : 0; break;
          case I_INST:  v = inst_idx ? inst_idx[inst] : inst; break;
          case I_CLASS: v = classes ? classes[inst] : 0; break;
          case I_VALID: v = area > 0; break;
          case I_NCONT: v = 0; break;
          case I_AREA:  v = area; break;
          case I_BX0: v = area > 0 ? xmin : -1; break;
          case I_BY0: v = area > 0 ? ymin : -1; break;
          case I_BX1: v = area > 0 ? xmax : -1; break;
          case I_BY1: v = area > 0 ? ymax : -1; break;
          case I_NPTS: v = 0; break;
          case I_M10: v = acc[1]; break;
          case I_M01: v = acc[2]; break;
          case I_M20: v = acc[3]; break;
          case I_M11: v = acc[4]; break;
          case I_M02: v = acc[5]; break;
          case I_M30: v = acc[6]; break;
          case I_M21: v = acc[7]; break;
          case I_M12: v = acc[8]; break;
          default: v = acc[9]; break;                  // I_M03
        }
        rows_i[inst * kNumInt + lane] = v;
      }
      __syncwarp();                                    // the mask copy is rewritten next round
    }
    inst = __shfl_sync(0xffffffffu, claim, 0);
  }
  // the last CTA out re-arms the counters for the next launch on this workspace
  __syncthreads();
  if (tid == 0) {
    const unsigned int done = atomicAdd(&ws.sched[2], 1u);
    if (done == gridDim.x - 1) { ws.sched[0] = 0u; ws.sched[1] = 0u; ws.sched[2] = 0u; }
  }
}

// pass C: the visited / sign planes of the border trace start at zero (only the words the
// layout handed out: status[1], known on the device)
__global__ void __launch_bounds__(256)
zero_marks_kernel(uint32_t* __restrict__ V, uint32_t* __restrict__ G,
                  const int64_t* __restrict__ status) {
  if (status[0] != 0) return;
  const int64_t n16 = (status[1] + 3) / 4;
  uint4* v = reinterpret_cast<uint4*>(V);
  uint4* g = reinterpret_cast<uint4*>(G);
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n16;
       k += (int64_t)gridDim.x * blockDim.x) {
    v[k] = z; g[k] = z;
  }
}

// ---------------------------------------------------------------------------------
// host-side launchers (called from the C ABI)
// ---------------------------------------------------------------------------------
cudaError_t launch_layout(const float* boxes, int64_t n, int H, int W, const Workspace& ws,
                          int64_t* status, int num_sms, cudaStream_t stream) {
  const int nblk = (int)layout_blocks(n);
  if (cudaMemsetAsync(status + 3, 0, sizeof(int64_t), stream) != cudaSuccess) return cudaGetLastError();
  if (cudaMemsetAsync(ws.block_sums + 2 * nblk, 0, sizeof(int64_t), stream) != cudaSuccess)   // moment-overflow flag
    return cudaGetLastError();
  share_carveout(layout_local_kernel);
  share_carveout(layout_rebase_kernel);
  share_carveout(zero_marks_kernel);
  layout_local_kernel<<<nblk, kLayoutThreads, 0, stream>>>(boxes, n, H, W, ws.desc, ws.block_sums,
                                                           status + 3);
  layout_rebase_kernel<<<nblk, kLayoutThreads, 0, stream>>>(n, nblk, ws.desc, ws.block_sums,
                                                            ws.cap_words, status, ws.sched);
  zero_marks_kernel<<<num_sms * 4, 256, 0, stream>>>(ws.V, ws.G, status);
  return cudaPeekAtLastError();
}

// instances [first, first + count) of the call (kernel argument n = first + count: the end)
cudaError_t launch_paste_measure(const float* masks, const float* boxes, const int32_t* image_idx,
                                 const int32_t* inst_idx, const int64_t* classes, int64_t first,
                                 int64_t count, int H, int W, float thr, uint32_t* planes,
                                 int64_t* rows_i, const Workspace& ws, const int64_t* status,
                                 int num_sms, cudaStream_t stream, const MaskSource& src) {
  if (count == 0) return cudaSuccess;
  const int64_t n = first + count;
  int per_sm = 0;
  cudaError_t e;
  // defaults chosen on B200 (profiles/): 16 KB zero source, unbounded bulk stores in flight, TMA
  // fill, tile band composed in shared memory up to 36 KB.  The release library reads NO
  // environment variable; the sweep knobs exist only in the -DUWCV_TUNING build
  // (lib/libuwcv_tuning.so, used by tools/*.sh).
  int zero_bytes = kZeroBytesDefault, rot_mul = 0, fill_mode = 0, debug_skip = 0;
  int band_cap = 36 * 1024;
#ifdef UWCV_TUNING
  if (const char* v = getenv("UWCV_DEBUG_SKIP")) debug_skip = atoi(v) & 15;
  if (const char* v = getenv("UWCV_FILL")) { fill_mode = atoi(v); if (fill_mode < 0 || fill_mode > 2) fill_mode = 0; }
  if (const char* v = getenv("UWCV_ZERO_KB")) zero_bytes = atoi(v) * 1024;
  if (const char* v = getenv("UWCV_PASTE_ROT")) rot_mul = atoi(v);
  if (zero_bytes < 1024 || zero_bytes > 160 * 1024 || (zero_bytes & 1023)) zero_bytes = kZeroBytesDefault;
  if (const char* v = getenv("UWCV_BAND_KB")) band_cap = atoi(v) * 1024;       // 0: generic stores
#endif
  if (band_cap < 0 || band_cap > 96 * 1024 || !planes) band_cap = 0;
  const size_t dyn = planes ? (size_t)zero_bytes + band_cap : 0;
  const int threads = kPasteThreads;
  share_carveout(paste_measure_kernel<true>);
  share_carveout(paste_measure_kernel<false>);
  share_carveout(tile_measure_kernel);
  if (planes) {
    if (dyn > 40 * 1024)
      cudaFuncSetAttribute(paste_measure_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, paste_measure_kernel<true>, threads, dyn);
  } else {
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, paste_measure_kernel<false>, kPasteThreads, dyn);
  }
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
#ifdef UWCV_TUNING
  // caps the CTAs per SM: 1 -> 4 780, 2 -> 5 590, 3 -> 5 640 GB/s (static split, r01)
  if (const char* v = getenv("UWCV_PASTE_CTAS")) {
    const int cap = atoi(v);
    if (cap >= 1 && cap < per_sm) per_sm = cap;
  }
#endif
  int64_t grid = (int64_t)num_sms * per_sm;           // persistent: a whole number of waves
  if (grid > count) grid = count;
  if (planes) {
    paste_measure_kernel<true><<<(unsigned)grid, threads, dyn, stream>>>(
          masks, boxes, image_idx, inst_idx, classes, n, H, W, thr, planes, rows_i, ws, status,
          zero_bytes, rot_mul, fill_mode, debug_skip, first, src, 0, band_cap, nullptr);
    return cudaPeekAtLastError();
  }
  // rows only: one instance per warp for ordinary tiles, then the whole-CTA kernel for the
  // few giant ones (it returns at once when the layout counted none)
  bool per_warp = true;
#ifdef UWCV_TUNING
  per_warp = !getenv("UWCV_TILE_PER_CTA");
#endif
  if (per_warp) {
    int tper = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&tper, tile_measure_kernel, kTileWarps * 32,
                                                      0) != cudaSuccess || tper < 1)
      tper = 2;
    int64_t tgrid = (int64_t)num_sms * tper;
    const int64_t need = (count + kTileWarps - 1) / kTileWarps;
    if (tgrid > need) tgrid = need;
    tile_measure_kernel<<<(unsigned)tgrid, kTileWarps * 32, 0, stream>>>(
        masks, boxes, image_idx, inst_idx, classes, n, H, W, thr, rows_i, ws, status, first, src);
  }
  paste_measure_kernel<false><<<(unsigned)grid, kPasteThreads, dyn, stream>>>(
      masks, boxes, image_idx, inst_idx, classes, n, H, W, thr, planes, rows_i, ws, status,
      zero_bytes, rot_mul, fill_mode, debug_skip, first, src, per_warp ? 1 : 0, 0, nullptr);
  return cudaPeekAtLastError();
}

}  // namespace uwcv

