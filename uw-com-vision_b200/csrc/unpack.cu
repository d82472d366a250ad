// Bit-plane -> Detectron2-literal N x H x W bool (uint8 0/1), on demand.
//
// detectron2's paste_masks_in_image returns a bool tensor (one byte per pixel); the
// product keeps 1 bit per pixel and expands only when a caller asks for the literal
// layout (nn_inference.py:376 reads pred_masks as such).  One thread expands one
// 32-pixel word into two 16-byte stores; rows are W bytes, words past W are skipped.
#include "uwcv_common.cuh"

namespace uwcv {

__global__ void __launch_bounds__(256)
unpack_planes_kernel(const uint32_t* __restrict__ planes, int64_t n, int H, int W, int wpr,
                     uint8_t* __restrict__ out) {
  const int wreal = (W + 31) >> 5;
  const int64_t total = n * H * wreal;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total;
       k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = k / wreal;
    const int wi = (int)(k - row * wreal);
    const uint32_t bits = __ldg(planes + row * wpr + wi);
    uint8_t* dst = out + row * W + (int64_t)wi * 32;
    const int npx = min(32, W - wi * 32);
    if (npx == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
      uint32_t v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t nib = (bits >> (4 * q)) & 0xFu;
        // spread 4 bits into 4 bytes: bit i -> byte i
        v[q] = ((nib & 1u)) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
      }
      reinterpret_cast<uint4*>(dst)[0] = make_uint4(v[0], v[1], v[2], v[3]);
      reinterpret_cast<uint4*>(dst)[1] = make_uint4(v[4], v[5], v[6], v[7]);
    } else {
      for (int i = 0; i < npx; ++i) dst[i] = (bits >> i) & 1u;
    }
  }
}

cudaError_t launch_unpack(const uint32_t* planes, int64_t n, int H, int W, uint8_t* out,
                          int num_sms, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const int wpr = plane_row_words(W);
  const int64_t total = n * H * ((W + 31) >> 5);
  // one word per thread in a one-shot grid: the hardware hands CTAs to whichever SM drains
  // first (a static grid-stride split stops at 6.3 TB/s of stores, tools/fill_bench2.cu)
  (void)num_sms;
  int64_t grid = (total + 255) / 256;
  if (grid > 0x7fffffffLL) grid = 0x7fffffffLL;
  unpack_planes_kernel<<<(unsigned)grid, 256, 0, stream>>>(planes, n, H, W, wpr, out);
  return cudaPeekAtLastError();
}

// Small host arrays (boxes, scores, classes, image / instance indices: ~1.5 MB per 64 000
// instances) are pulled from pinned, device-mapped host memory by the SMs instead of being
// queued on the copy engine: the engine serves copies in issue order, so a 1 MB copy issued
// behind 200 MB of mask probabilities lands 4 ms later and holds the layout kernel back.
__global__ void __launch_bounds__(256)
ingest_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n16) dst[i] = src[i];
}

cudaError_t launch_ingest(const void* src_host_mapped, void* dst, size_t bytes, cudaStream_t stream) {
  const size_t n16 = (bytes + 15) / 16;
  if (n16 == 0) return cudaSuccess;
  share_carveout(ingest_kernel);
  ingest_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(src_host_mapped), reinterpret_cast<uint4*>(dst), n16);
  return cudaPeekAtLastError();
}

}  // namespace uwcv
