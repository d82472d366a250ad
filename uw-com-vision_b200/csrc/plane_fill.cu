// Split pipeline, stage 8: the full-frame bit-planes written FROM THE TILES by a kernel that is
// nothing but data movement -- so that the HBM-bound part of the step (512 KiB of plane per
// instance at 2048^2, 97 % of it zeros) no longer owns the SMs while it waits for the write path.
//
// The fused paste kernel (paste_measure.cu) holds 3 CTAs x 256 threads x 80 registers per SM for
// the 4.9 ms it needs to push 33.8 GB through the TMA, although its issue slots are ~15 % busy;
// the border trace (latency-bound, 1.7 ms) and the tile arithmetic (compute-bound, 0.65 ms) cannot
// run beside it, so ~1.2 ms of every step stay exposed.  Here the same bytes leave through the same
// instructions -- cp.async.bulk shared -> global from a zeroed buffer for the rows above / below a
// tile band, the band itself composed in shared memory (zeros + the tile's words, read back from
// the tile workspace: 0.7 KB per instance) and sent with one bulk store -- but from CTAs of 96
// threads and ~30 registers, two per SM: one issuer warp streaming the zero rows, two warps
// composing bands.  What is left of every SM (9/10 of its registers, all of its issue slots) runs
// the tile kernel of the NEXT call and the border trace of THIS one at full occupancy, on their own
// streams (uwcv/api.py::Engine.run_overlapped(split=True)).
//
// Same output as paste_measure_kernel<true>: plane row y = zeros, or for y in [y0, y0 + th) the
// tile's words at word columns [wx0, wx0 + tw) and zeros elsewhere (Detectron2's
// paste_masks_in_image, reached from nn_inference.py:372).
#include <cstdlib>
#include "uwcv_common.cuh"

namespace uwcv {

constexpr int kFillComposeWarps = 2;
constexpr int kFillComposeThreads = kFillComposeWarps * 32;
constexpr int kFillThreads = 32 + kFillComposeThreads;     // warp 0 issues the zero rows
constexpr int kFillZeroBytes = 16384;
constexpr int kFillBandBytes = 36 * 1024;

namespace {
__device__ __forceinline__ void fill_bulk_store(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fill_bulk_store_hint(void* gdst, uint32_t smem_src, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
               :: "l"(gdst), "r"(smem_src), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void fill_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void fill_wait_read_all() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fill_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void compose_barrier() {
  asm volatile("bar.sync 1, %0;" ::"n"(kFillComposeThreads) : "memory");
}
}  // namespace

__global__ void __launch_bounds__(kFillThreads)
plane_fill_kernel(uint32_t* __restrict__ planes, int64_t first, int64_t n, int H, int W, Workspace ws,
                  const int64_t* __restrict__ status, int policy_mode) {
  extern __shared__ __align__(128) unsigned char s_fill[];   // zero source | band image
  __shared__ long long s_next[2];
  if (status[0] != 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wpr = plane_row_words(W);
  const int64_t plane_words = (int64_t)H * wpr;
  for (int k = tid; k < (kFillZeroBytes + kFillBandBytes) / 16; k += kFillThreads)
    reinterpret_cast<uint4*>(s_fill)[k] = make_uint4(0, 0, 0, 0);
  fill_fence_async();                                       // generic-proxy zeros -> visible to the TMA
  __syncthreads();
  const uint32_t zero_smem = (uint32_t)__cvta_generic_to_shared(s_fill);
  uint32_t* s_band = reinterpret_cast<uint32_t*>(s_fill + kFillZeroBytes);
  const uint32_t band_smem = (uint32_t)__cvta_generic_to_shared(s_band);

  if (warp == 0) {
    // ---- zero rows above / below every band: planes claimed one at a time from a device-wide
    //      counter (a static split ends with the slowest SM, profiles/README.md r01 v9) ---------
    if (lane == 0) {
      uint64_t policy = 0;
      if (policy_mode & 3)
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      int64_t inst = first + atomicAdd(&ws.sched[4], 1u);
      while (inst < n) {
        const int64_t next = first + atomicAdd(&ws.sched[4], 1u);     // in flight under the fill
        const TileDesc d = ws.desc[inst];
        char* base = reinterpret_cast<char*>(planes + inst * plane_words);
        const int band_lo = d.th > 0 ? d.y0 : H;              // empty tile: the whole plane is zero
        const int band_hi = d.th > 0 ? d.y0 + d.th : H;
        const int64_t seg_lo[2] = {0, (int64_t)band_hi * wpr * 4};
        const int64_t seg_hi[2] = {(int64_t)band_lo * wpr * 4, plane_words * 4};
        for (int sgi = 0; sgi < 2; ++sgi)
          for (int64_t o = seg_lo[sgi]; o < seg_hi[sgi]; o += kFillZeroBytes) {
            const int64_t rem = seg_hi[sgi] - o;
            const uint32_t nbytes = (uint32_t)(rem < kFillZeroBytes ? rem : kFillZeroBytes);
            UWCV_BOUND(o + nbytes, plane_words * 4 + 1); UWCV_BOUND(inst, n);
            if (policy_mode & 3) fill_bulk_store_hint(base + o, zero_smem, nbytes, policy);
            else fill_bulk_store(base + o, zero_smem, nbytes);
          }
        fill_commit();
        inst = next;
      }
      fill_wait_read_all();                                   // the zero source must outlive the copies
    }
  } else if (!(policy_mode & 4)) {                         // (4: tuning, no bands)
    // ---- bands: zeros + the tile's words, composed in shared memory, one bulk store per chunk of
    //      rows that fits the band image (a whole band for tiles up to 144 rows at 2048 px) --------
    const int ct = tid - 32;
    uint64_t policy = 0;
    if ((policy_mode & 3) > 1)
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    const int rows_cap = kFillBandBytes / (wpr * 4);          // >= 9 (wpr <= 1024 words)
    if (ct == 0) s_next[0] = first + (long long)atomicAdd(&ws.sched[5], 1u);
    compose_barrier();
    for (int it = 0;; ++it) {
      const int64_t inst = s_next[it & 1];
      if (inst >= n) break;
      long long claim = 0;                                    // travels under this instance's work
      if (ct == 0) claim = first + (long long)atomicAdd(&ws.sched[5], 1u);
      const TileDesc d = ws.desc[inst];
      const uint32_t* __restrict__ tM = ws.M + d.word_off;
      uint32_t* plane = planes + inst * plane_words;
      UWCV_BOUND(d.word_off + (int64_t)d.tw * d.th, ws.cap_words + 1);
      UWCV_BOUND(d.y0 + d.th, H + 1); UWCV_BOUND(d.wx0 + d.tw, wpr + 1);
      for (int r0 = 0; r0 < d.th; r0 += rows_cap) {
        const int nr = min(rows_cap, d.th - r0);
        const int nw = nr * d.tw;
        for (int k = ct; k < nw; k += kFillComposeThreads) {
          const int r = k / d.tw, c = k - r * d.tw;
          UWCV_BOUND((r * wpr + d.wx0 + c) * 4, kFillBandBytes);
          s_band[r * wpr + d.wx0 + c] = __ldg(tM + (int64_t)r0 * d.tw + k);
        }
        fill_fence_async();                                   // my words -> visible to the TMA
        compose_barrier();
        if (ct == 0) {
          if ((policy_mode & 3) > 1)
            fill_bulk_store_hint(plane + (int64_t)(d.y0 + r0) * wpr, band_smem, (uint32_t)(nr * wpr * 4), policy);
          else
            fill_bulk_store(plane + (int64_t)(d.y0 + r0) * wpr, band_smem, (uint32_t)(nr * wpr * 4));
          fill_commit();
          fill_wait_read_all();                               // the image is reused for the next chunk
        }
        compose_barrier();
        for (int k = ct; k < nw; k += kFillComposeThreads) {  // back to all zeros
          const int r = k / d.tw, c = k - r * d.tw;
          s_band[r * wpr + d.wx0 + c] = 0u;
        }
      }
      if (ct == 0) s_next[(it + 1) & 1] = claim;
      fill_fence_async();
      compose_barrier();
    }
  }
  // the last CTA out re-arms the counters for the next launch on this workspace
  __syncthreads();
  if (tid == 0) {
    const unsigned int done = atomicAdd(&ws.sched[6], 1u);
    if (done == gridDim.x - 1) { ws.sched[4] = 0u; ws.sched[5] = 0u; ws.sched[6] = 0u; }
  }
}

// planes of instances [first, first + count) from their tiles (which the paste stage has written)
cudaError_t launch_plane_fill(uint32_t* planes, int64_t first, int64_t count, int H, int W,
                              const Workspace& ws, const int64_t* status, int num_sms,
                              cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  const size_t dyn = (size_t)kFillZeroBytes + kFillBandBytes;
  if (cudaFuncSetAttribute(plane_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) !=
      cudaSuccess)
    return cudaGetLastError();
  share_carveout(plane_fill_kernel);
  int per_sm = 2;                      // two issuer threads per SM saturate the write path (r01 v9)
#ifdef UWCV_TUNING
  if (const char* v = getenv("UWCV_FILL_CTAS")) { const int c = atoi(v); if (c >= 1 && c <= 4) per_sm = c; }
#endif
  // every plane byte is marked evict-first in L2: the 33 GB stream then recycles a small part of the
  // cache instead of pushing out the tiles, marks and extremes the co-running border trace works on
  // (fill next to the trace 5.52 -> 5.24 ms per 64 000 instances, the trace itself 5.0 -> 4.2 ms;
  // profiles/r02_split_contention.txt)
  int policy_mode = 2;
#ifdef UWCV_TUNING
  if (const char* v = getenv("UWCV_FILL_POLICY")) { policy_mode = atoi(v); if (policy_mode < 0 || policy_mode > 7) policy_mode = 0; }
#endif
  int64_t grid = (int64_t)num_sms * per_sm;
  if (grid > count) grid = count;
  plane_fill_kernel<<<(unsigned)grid, kFillThreads, dyn, stream>>>(planes, first, first + count, H, W, ws,
                                                                  status, policy_mode);
  return cudaPeekAtLastError();
}

}  // namespace uwcv
