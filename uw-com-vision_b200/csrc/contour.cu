// Kernel 3: external-border trace, contour descriptors, min-area rectangle / Feret,
// and the floating-point columns of the measurement row.  One thread per instance, written
// warp-converged: every loop with a data-dependent trip count is a warp-uniform loop
// (`while (__any_sync(...))`) in which each lane advances its own instance by one
// micro-step, so the 32 instances of a warp run the scan / border-following / hull /
// calipers code in lock step instead of serialising 32 divergent paths.
//
// Replaces nn_inference.py:405-459 (cvtColor + cv2.findContours(RETR_EXTERNAL,
// CHAIN_APPROX_SIMPLE) + contourArea / arcLength / minAreaRect / boxPoints + the
// descriptor block) applied to each instance mask, and cv2.moments' central moments.
//
// The arithmetic mirrors the library code the reference runs (pinned bit-for-bit
// against OpenCV 4.13 by the oracle tests):
//   * border following: Suzuki-Abe as in OpenCV's contour scanner -- raster scan, an
//     unvisited foreground pixel with background on its left starts an outer border
//     unless the last marked pixel to its left on the row carries a positive mark;
//     first move counter-clockwise on screen; a pixel gets the negative mark when its
//     right neighbour was examined and found empty.
//   * contourArea: shoelace over the steps (exact integers).
//   * arcLength: float32 sqrt per CHAIN_APPROX_SIMPLE segment, summed in double.
//   * minAreaRect: strict convex hull (clockwise on screen, the component's raster-first
//     pixel last) -> float32 rotating calipers (exact cross-product side selection,
//     "area <= minarea" keeps the last minimum) -> centre / sides / angle in [-90, 0).
//   * boxPoints -> int truncation -> imutils order_points -> midpoints -> dA, dB -> the
//     nine descriptors of nn_inference.py:434-449.
// Compiled with -fmad=false: no contraction anywhere in this file.
#include <cstdlib>
#include "contour_common.cuh"

namespace uwcv {

constexpr int kContourThreads = 64;
#ifdef UWCV_TUNING
// per-instance iteration counts of the state machine (scan steps, border steps, contours, loop
// iteration at which the lane finished): tools/trace_stats.py
__device__ int32_t* g_trace_stats = nullptr;
#endif
static_assert((kNumInt * 8) % 16 == 0 && (kNumFloat * 8) % 16 == 0, "rows are copied as 16-byte words");
static_assert(F_CHORDS - F_CAREA == 15, "describe_contour writes 16 contiguous float columns");

#ifdef UWCV_CONTOUR_MINBLOCKS                               // (tuning sweep: register cap vs co-residency)
__global__ void __launch_bounds__(kContourThreads, UWCV_CONTOUR_MINBLOCKS)
#else
__global__ void __launch_bounds__(kContourThreads)
#endif
contour_measure_kernel(int64_t first, int64_t n, int lanes, const float* __restrict__ scores,
                       double pixels_per_metric, int64_t* __restrict__ rows_i,
                       double* __restrict__ rows_f, Workspace ws,
                       const int64_t* __restrict__ status, GatherDst g) {
  if (status[0] != 0) return;                            // uniform over the grid
  // `lanes` instances per warp: the kernel is bound by the latency of each lane's serial
  // chain, not by issue slots, so small batches are spread over more warps
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * kContourThreads + threadIdx.x) >> 5;
  const int64_t inst = first + warp * lanes + lane;
  const bool live = lane < lanes && inst < n;
  const double kPi = 3.141592653589793;
  TileDesc d;
  d.wx0 = d.y0 = d.tw = d.th = 0; d.word_off = d.row_off = 0;
  if (live) d = ws.desc[inst];
  int64_t* ri = rows_i + (live ? inst : 0) * kNumInt;
  double* rf = rows_f + (live ? inst : 0) * kNumFloat;
  bool work = false;                                     // this lane has a non-empty mask
  if (live) {
    for (int k = 0; k < kNumFloat; ++k) rf[k] = 0.0;
    rf[F_SCORE] = scores ? (double)scores[inst] : 0.0;
    work = ri[I_AREA] > 0 && d.th > 0;
  }

  // ---- central moments (OpenCV completeMomentState) ---------------------------------
  if (work) {
    const double m00 = (double)ri[I_AREA], m10 = (double)ri[I_M10], m01 = (double)ri[I_M01];
    const double m20 = (double)ri[I_M20], m11 = (double)ri[I_M11], m02 = (double)ri[I_M02];
    const double m30 = (double)ri[I_M30], m21 = (double)ri[I_M21], m12 = (double)ri[I_M12];
    const double m03 = (double)ri[I_M03];
    const double inv_m00 = 1.0 / m00;
    const double cx = m10 * inv_m00, cy = m01 * inv_m00;
    const double mu20 = m20 - m10 * cx;
    double mu11 = m11 - m10 * cy;
    const double mu02 = m02 - m01 * cy;
    rf[F_CX] = cx; rf[F_CY] = cy;
    rf[F_MU20] = mu20; rf[F_MU11] = mu11; rf[F_MU02] = mu02;
    rf[F_MU30] = m30 - cx * (3 * mu20 + cx * m10);
    const double a = mu20 / m00, b = mu11 / m00, c = mu02 / m00;
    mu11 += mu11;
    rf[F_MU21] = m21 - cx * (mu11 + cx * m01) - cy * mu20;
    rf[F_MU12] = m12 - cy * (mu11 + cy * m10) - cx * mu02;
    rf[F_MU03] = m03 - cy * (3 * mu02 + cy * m01);
    rf[F_EQD] = sqrt(4.0 * m00 / kPi);
    // moment ellipse (the build's definition)
    const double common = sqrt(((a - c) * 0.5) * ((a - c) * 0.5) + b * b);
    const double lp = (a + c) * 0.5 + common, lm = (a + c) * 0.5 - common;
    rf[F_ELL_MAJOR] = 4.0 * sqrt(lp > 0.0 ? lp : 0.0);
    rf[F_ELL_MINOR] = 4.0 * sqrt(lm > 0.0 ? lm : 0.0);
    rf[F_ELL_THETA] = 0.5 * atan2(2.0 * b, a - c);
  }

  // ---- raster scan + border following; keep the largest contour and its row extremes ---
  // Lane per instance: each iteration of the warp-uniform loop advances a lane by one 64-pixel
  // scan step or one border step (LaneTracer), so no lane waits for another lane's contour.
  TileView t;
  t.M = ws.M + d.word_off; t.V = ws.V + d.word_off; t.G = ws.G + d.word_off;
  t.tw = d.tw; t.th = d.th;
  // two sets of per-row extremes (left | right): the contour being traced and the best so far
  uint32_t* ext0 = ws.scratch + 4 * d.row_off;
  if (live) {
    UWCV_BOUND(d.word_off + (int64_t)d.tw * d.th, ws.cap_words + 1);
    UWCV_BOUND(4 * (d.row_off + d.th), 4 * ws.cap_words + 1);     // extremes: 4 words per tile row
  }
  LaneTracer<GlobalMem> T;
  T.idle();
  // rows outside the pixel bbox cannot hold a start pixel
  if (work) T.begin(t, ext0, ext0 + 2 * d.th, d.th, (int)ri[I_BY0] - d.y0, (int)ri[I_BY1] - d.y0);
#ifdef UWCV_TUNING
  int st_scan = 0, st_trace = 0, st_iter = 0, st_end = 0;
#endif
  while (__any_sync(kFull, !T.done())) {
#ifdef UWCV_TUNING
    ++st_iter;
    if (T.state == LaneTracer<GlobalMem>::kScan) ++st_scan;
    else if (T.state == LaneTracer<GlobalMem>::kTrace) ++st_trace;
    if (!T.done()) st_end = st_iter;
#endif
    if (!T.done()) T.step();
  }
  const int ncont = T.ncont, best_npts = T.best_npts, best_y = T.best_y, best_ymax = T.best_ymax;
  const long long best_a2 = T.best_a2;
  const double best_perim = T.best_perim;
  uint32_t* best = T.best;
  if (work) { ri[I_NCONT] = ncont; ri[I_NPTS] = best_npts; }
  const bool have = work && ncont > 0;
#ifdef UWCV_TUNING
  if (g_trace_stats && live) {
    int32_t* o = g_trace_stats + 4 * inst;
    o[0] = st_scan; o[1] = st_trace; o[2] = ncont; o[3] = st_end;
  }
#endif

  // ---- hull, min-area rectangle and descriptor block of the best contour ----------------
  describe_contour(have, best, best + d.th, best_y, best_ymax, d.wx0 * 32, d.y0, best_a2,
                   best_perim, pixels_per_metric, rf + F_CAREA);

  // ---- fused all-gather: publish this warp's rows into every rank's table ------------------
  // The `lanes` instances of a warp are consecutive, so its rows are one contiguous block of
  // each table (160 B and 240 B per row): the warp copies them with coalesced 16-byte stores
  // into the same rows (shifted by this rank's offset) of every peer, over NVLink for the
  // remote ones.  The tables become consistent at the caller's barrier after this kernel.
  if (g.world > 0) {
    __syncwarp();
    const int64_t w0 = first + warp * lanes;
    if (w0 < n) {
      const int cnt = (int)((n - w0) < (int64_t)lanes ? (n - w0) : (int64_t)lanes);
      const uint4* si = reinterpret_cast<const uint4*>(rows_i + w0 * kNumInt);
      const uint4* sf = reinterpret_cast<const uint4*>(rows_f + w0 * kNumFloat);
      const int ni = cnt * (kNumInt * 8 / 16), nf = cnt * (kNumFloat * 8 / 16);
      for (int p = 0; p < g.world; ++p) {
        if (g.only >= 0 && p != g.only) continue;      // gather to one destination
        uint4* di = reinterpret_cast<uint4*>(g.rows_i[p] + (g.row_base + w0) * kNumInt);
        uint4* df = reinterpret_cast<uint4*>(g.rows_f[p] + (g.row_base + w0) * kNumFloat);
        for (int k = lane; k < ni; k += 32) di[k] = si[k];
        for (int k = lane; k < nf; k += 32) df[k] = sf[k];
      }
    }
  }
}

// Dynamic shared memory the trace of the split pipeline asks for and never touches, so that the
// driver honours its carve-out hint (share_carveout, uwcv_common.cuh).  Streamed call, one fill per
// call, two calls in flight -- the trace lands a few microseconds before the fill of its own call:
// 6.05 ms per 64 000 instances with 0 B, 5.50 with 1 KB, 4.91 with 16 KB, 5.26 with 40 KB (fewer
// trace CTAs fit); three calls in flight 4.60 / 4.57 / 4.58 / 5.12 (profiles/r02_carveout_probe.txt).
constexpr size_t kContourCarveoutBytes = 16 * 1024;

cudaError_t launch_contour_measure(int64_t first, int64_t count, const float* scores, double ppm,
                                   int64_t* rows_i, double* rows_f, const Workspace& ws,
                                   const int64_t* status, int num_sms, cudaStream_t stream,
                                   const GatherDst& gather, int reserve_ctas) {
  if (count == 0) return cudaSuccess;
  const int64_t n = count;                               // sizing below is per launched range
  const bool beside_fill = reserve_ctas > 0;
  share_carveout(contour_measure_kernel, beside_fill);
  size_t dyn = beside_fill ? kContourCarveoutBytes : 0;
#ifdef UWCV_TUNING
  if (const char* v = getenv("UWCV_TRACE_SMEM")) {
    const int b = atoi(v);
    if (beside_fill && b >= 0 && b <= 48 * 1024) dyn = (size_t)b;
  }
#endif
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, contour_measure_kernel,
                                                    kContourThreads, dyn) != cudaSuccess || per_sm < 1)
    per_sm = 8;
  // split pipeline: the plane-fill CTAs of the same call are resident beside this kernel; keep the
  // trace grid within ONE wave of what is left (a second wave starts behind the shortest warps
  // and ends a whole serial chain later)
  if (reserve_ctas > 0 && per_sm > reserve_ctas + 1) per_sm -= reserve_ctas;
  // as many warps as can be resident at once, each carrying ceil(n / warps) instances
  const int64_t resident_warps = (int64_t)num_sms * per_sm * (kContourThreads / 32);
  int lanes = (int)((n + resident_warps - 1) / resident_warps);
  lanes = lanes < 1 ? 1 : (lanes > 32 ? 32 : lanes);
#ifdef UWCV_TUNING
  if (const char* v = getenv("UWCV_TRACE_LANES")) { const int l = atoi(v); if (l >= 1 && l <= 32) lanes = l; }
#endif
  const int64_t warps = (n + lanes - 1) / lanes;
  const unsigned grid = (unsigned)((warps * 32 + kContourThreads - 1) / kContourThreads);
  contour_measure_kernel<<<grid, kContourThreads, dyn, stream>>>(first, first + count, lanes, scores,
                                                               ppm, rows_i, rows_f, ws, status,
                                                               gather);
  return cudaPeekAtLastError();
}

}  // namespace uwcv

#ifdef UWCV_TUNING
extern "C" int uwcv_tuning_set_trace_stats(int32_t* dev_ptr) {
  return cudaMemcpyToSymbol(uwcv::g_trace_stats, &dev_ptr, sizeof(dev_ptr)) == cudaSuccess ? 0 : -6;
}
#endif
