// Does a zero-dominated plane fill get faster in COMPRESSIBLE device memory (compute data compression:
// cuMemCreate with CU_MEM_ALLOCATION_COMP_GENERIC)?  The planes of the hot path are 97 % zeros.
// Persistent TMA fill (the shape of plane_fill_kernel: 16 KB bulk stores, dynamic claims) of a 16 GB buffer that
// is (a) ordinary cudaMalloc memory, (b) a compressible VMM allocation; fill value zero, and a plane-like pattern
// (97 % zero rows + a band of random words) so that the incompressible part is there too; cudaMemset for reference;
// read-back (sum kernel) speed of both buffers and a check that the contents are what was written.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fill_compress tools/fill_compress.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
#define CU(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char* s; cuGetErrorString(e, &s); printf("%s: %s\n", #x, s); return 1; } } while (0)

// one claim = one 512 KB "plane": rows of 256 B; rows [band0, band0 + 48) get pattern words, the rest zeros
__global__ void __launch_bounds__(64) fill_kernel(char* dst, size_t planes, unsigned* counter, int with_band) {
  extern __shared__ __align__(128) unsigned char sm[];        // 16 KB zero source | 12 KB band image
  for (int k = threadIdx.x; k < 16384 / 16; k += blockDim.x) reinterpret_cast<uint4*>(sm)[k] = make_uint4(0, 0, 0, 0);
  uint32_t* band = reinterpret_cast<uint32_t*>(sm + 16384);
  for (int k = threadIdx.x; k < 12288 / 4; k += blockDim.x) {
    const int col = k & 63;                                   // 64 words per 256-byte row: a 4-word tile at words 20..23
    band[k] = (col >= 20 && col < 24) ? (0x9E3779B9u * (k + 1)) : 0u;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t zsrc = (uint32_t)__cvta_generic_to_shared(sm), bsrc = zsrc + 16384;
    for (;;) {
      const size_t p = atomicAdd(counter, 1u);
      if (p >= planes) break;
      char* base = dst + p * (512 * 1024);
      const int band0 = with_band ? (int)((p * 37) % 1900) : 2048;      // first band row
      const size_t lo = (size_t)band0 * 256, hi = with_band ? lo + 12288 : lo;
      for (size_t o = 0; o < lo; o += 16384) {
        const uint32_t nb = (uint32_t)((lo - o) < 16384 ? (lo - o) : 16384);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(base + o), "r"(zsrc), "r"(nb) : "memory");
      }
      if (with_band)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(base + lo), "r"(bsrc), "r"(12288) : "memory");
      for (size_t o = hi; o < 512 * 1024; o += 16384) {
        const uint32_t nb = (uint32_t)((512 * 1024 - o) < 16384 ? (512 * 1024 - o) : 16384);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(base + o), "r"(zsrc), "r"(nb) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

__global__ void sum_kernel(const uint4* p, size_t n16, unsigned long long* out) {
  unsigned long long s = 0;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n16; k += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = p[k];
    s += (unsigned long long)v.x + v.y + v.z + v.w;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

int main() {
  const size_t bytes = (size_t)16 << 30, planes = bytes / (512 * 1024);
  CK(cudaFree(0));
  CUdevice dev; CU(cuDeviceGet(&dev, 0));
  int comp = 0; CU(cuDeviceGetAttribute(&comp, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, dev));
  printf("generic compression supported: %d\n", comp);
  char* plain; CK(cudaMalloc(&plain, bytes));
  char* cmem = nullptr;
  if (comp) {
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
    prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
    size_t gran = 0; CU(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    const size_t sz = (bytes + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h; CU(cuMemCreate(&h, sz, &prop, 0));
    CUmemAllocationProp got = {}; CU(cuMemGetAllocationPropertiesFromHandle(&got, h));
    printf("granularity %zu, compression granted: %d\n", gran, (int)got.allocFlags.compressionType);
    CUdeviceptr va; CU(cuMemAddressReserve(&va, sz, 0, 0, 0)); CU(cuMemMap(va, sz, 0, h, 0));
    CUmemAccessDesc acc = {}; acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CU(cuMemSetAccess(va, sz, &acc, 1));
    cmem = (char*)va;
  }
  unsigned* counter; CK(cudaMalloc(&counter, 4));
  unsigned long long* out; CK(cudaMalloc(&out, 8));
  CK(cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 12288));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto timed_fill = [&](char* dst, int with_band, const char* name) {
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaMemset(counter, 0, 4); cudaDeviceSynchronize();
      cudaEventRecord(a); fill_kernel<<<148 * 2, 64, 16384 + 12288>>>(dst, planes, counter, with_band); cudaEventRecord(b);
      cudaDeviceSynchronize(); float t; cudaEventElapsedTime(&t, a, b); if (t < best) best = t;
    }
    printf("%-44s %.3f ms  %.0f GB/s\n", name, best, bytes / best / 1e6);
  };
  auto timed_memset = [&](char* dst, const char* name) {
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(a); cudaMemsetAsync(dst, 0, bytes); cudaEventRecord(b); cudaDeviceSynchronize();
      float t; cudaEventElapsedTime(&t, a, b); if (t < best) best = t; }
    printf("%-44s %.3f ms  %.0f GB/s\n", name, best, bytes / best / 1e6);
  };
  auto timed_sum = [&](char* src, const char* name) {
    float best = 1e9f; unsigned long long h = 0;
    for (int rep = 0; rep < 3; ++rep) { cudaMemset(out, 0, 8); cudaEventRecord(a); sum_kernel<<<148 * 16, 256>>>((const uint4*)src, bytes / 16, out); cudaEventRecord(b);
      cudaDeviceSynchronize(); float t; cudaEventElapsedTime(&t, a, b); if (t < best) best = t; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost); }
    printf("%-44s %.3f ms  %.0f GB/s  sum %llu\n", name, best, bytes / best / 1e6, h);
  };
  timed_memset(plain, "cudaMemset, ordinary memory");
  if (cmem) timed_memset(cmem, "cudaMemset, compressible memory");
  timed_fill(plain, 0, "TMA zero fill, ordinary memory");
  if (cmem) timed_fill(cmem, 0, "TMA zero fill, compressible memory");
  timed_fill(plain, 1, "TMA plane-like fill (band), ordinary");
  timed_sum(plain, "read back, ordinary");
  if (cmem) { timed_fill(cmem, 1, "TMA plane-like fill (band), compressible"); timed_sum(cmem, "read back, compressible"); }
  return 0;
}
