// Micro-benchmark: how fast can a B200 be filled with zeros through (a) TMA bulk stores
// from shared memory, (b) 16-byte LSU stores, (c) cudaMemsetAsync.  Sets the ceiling for the
// paste kernel's zero rows.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fill_bench fill_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) fill_tma(char* dst, size_t bytes, int chunk, int issuers) {
  extern __shared__ __align__(128) unsigned char z[];
  for (int k = threadIdx.x; k < chunk / 16; k += blockDim.x) reinterpret_cast<uint4*>(z)[k] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint32_t zs = (uint32_t)__cvta_generic_to_shared(z);
  const size_t nchunks = bytes / chunk;
  // `issuers` threads of the CTA (one per warp) issue interleaved chunks
  if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < issuers) {
    const size_t stride = (size_t)gridDim.x * issuers;
    for (size_t c = (size_t)blockIdx.x * issuers + (threadIdx.x >> 5); c < nchunks; c += stride) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * chunk), "r"(zs), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

// every CTA streams its own private contiguous regions of `region` bytes (as a per-instance
// plane fill does), instead of interleaving chunks with the other CTAs
__global__ void __launch_bounds__(256) fill_tma_private(char* dst, size_t bytes, int chunk, size_t region) {
  extern __shared__ __align__(128) unsigned char z[];
  for (int k = threadIdx.x; k < chunk / 16; k += blockDim.x) reinterpret_cast<uint4*>(z)[k] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint32_t zs = (uint32_t)__cvta_generic_to_shared(z);
  if (threadIdx.x == 0) {
    const size_t nreg = bytes / region;
    for (size_t r = blockIdx.x; r < nreg; r += gridDim.x) {
      char* base = dst + r * region;
      for (size_t o = 0; o < region; o += chunk)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + o), "r"(zs), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

__global__ void __launch_bounds__(256) fill_lsu(uint4* dst, size_t n16) {
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = z;
}

int main() {
  const size_t bytes = (size_t)16 << 30;
  char* d; cudaMalloc(&d, bytes);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto time = [&](const char* name, auto fn) {
    fn(); cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < 5; ++r) { cudaEventRecord(a); fn(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    printf("%-44s %8.3f ms  %8.1f GB/s  (%s)\n", name, best, bytes / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  };
  time("cudaMemsetAsync", [&] { cudaMemsetAsync(d, 0, bytes); });
  for (int per_sm : {1, 2, 3, 4, 8}) {
    char nm[96]; snprintf(nm, 96, "LSU 16B stores, %d CTAs/SM x 256 thr", per_sm);
    time(nm, [&] { fill_lsu<<<148 * per_sm, 256>>>((uint4*)d, bytes / 16); });
  }
  cudaFuncSetAttribute(fill_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  for (int chunk : {4096, 16384, 65536}) for (int per_sm : {1, 3, 6}) for (int iss : {1, 4}) {
    if ((size_t)chunk * per_sm > 200 * 1024) continue;
    char nm[96]; snprintf(nm, 96, "TMA bulk %3d KB, %d CTAs/SM, %d issuers/CTA", chunk / 1024, per_sm, iss);
    time(nm, [&] { fill_tma<<<148 * per_sm, 256, chunk>>>(d, bytes, chunk, iss); });
  }
  for (size_t region : {(size_t)128 << 10, (size_t)512 << 10, (size_t)2 << 20}) for (int per_sm : {1, 3}) {
    char nm[96]; snprintf(nm, 96, "TMA private %4zu KB regions, 16 KB, %d CTAs/SM", region >> 10, per_sm);
    time(nm, [&] { fill_tma_private<<<148 * per_sm, 256, 16384>>>(d, bytes, 16384, region); });
  }
  return 0;
}
