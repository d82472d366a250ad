// What does a latency-bound co-runner cost a saturated HBM write stream on B200?
// Stream A: persistent TMA zero fill of a large buffer (16 KB bulk stores, dynamic claims: the shape of
// plane_fill_kernel).  Stream B: a grid of lanes each running a serial chain of memory operations of ONE kind
// (RED.OR, ST, LD through L1, LD past L1) over a buffer that is L2-resident (32 MB) or not (4 GB), with
// ~compute between the operations.  Prints A alone, A next to every B, and B's operation rate.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fill_contention tools/fill_contention.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(64) fill_kernel(char* dst, size_t chunks, unsigned* counter) {
  extern __shared__ __align__(128) unsigned char z[];
  for (int k = threadIdx.x; k < 16384 / 16; k += blockDim.x) reinterpret_cast<uint4*>(z)[k] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(z);
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    // a claim = 32 chunks of 16 KB (512 KB: one plane)
    for (;;) {
      const size_t c = (size_t)atomicAdd(counter, 1u) * 32;
      if (c >= chunks) break;
      for (int k = 0; k < 32 && c + k < chunks; ++k)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                     :: "l"(dst + (c + k) * 16384), "r"(src), "r"(16384), "l"(pol) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

// kind 0: red.or  1: st  2: ld (L1)  3: ld.cg (L2)  4: no memory operation (ALU only)
__global__ void __launch_bounds__(64) chain_kernel(uint32_t* buf, size_t words, int kind, int ops, int alu, int active_lanes,
                                                   unsigned long long* sink) {
  const int lane = threadIdx.x & 31;
  if (lane >= active_lanes) return;
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 7919u % words;
  uint32_t acc = (uint32_t)i;
  for (int k = 0; k < ops; ++k) {
    for (int a = 0; a < alu; ++a) acc = acc * 1664525u + 1013904223u;       // dependent ALU work between operations
    // the next address stays near the previous one (a border walk touches neighbouring rows)
    i = (i + 5 + (acc & 7)) % words;
    if (kind == 0) atomicOr(buf + i, 1u << (acc & 31));
    else if (kind == 1) buf[i] = acc;
    else if (kind == 2) acc += buf[i];
    else if (kind == 3) acc += __ldcg(buf + i);
  }
  if (acc == 0x12345678u) *sink = acc;
}

int main() {
  const size_t fill_bytes = (size_t)16 << 30, chunks = fill_bytes / 16384;
  char* dst; unsigned* counter; uint32_t* buf; unsigned long long* sink;
  CK(cudaMalloc(&dst, fill_bytes)); CK(cudaMalloc(&counter, 4)); CK(cudaMalloc(&sink, 8));
  const size_t big_words = (size_t)1 << 30;                 // 4 GB
  CK(cudaMalloc(&buf, big_words * 4)); CK(cudaMemset(buf, 0, big_words * 4));
  cudaStream_t sa, sb; int lo, hi; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CK(cudaStreamCreateWithPriority(&sa, cudaStreamNonBlocking, hi)); CK(cudaStreamCreateWithPriority(&sb, cudaStreamNonBlocking, lo));
  CK(cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  cudaEvent_t a0, a1, b0, b1; cudaEventCreate(&a0); cudaEventCreate(&a1); cudaEventCreate(&b0); cudaEventCreate(&b1);
  auto run = [&](int kind, size_t words, int ops, int alu, int lanes, int ctas, const char* name) {
    float best_a = 1e9f, best_b = 0.f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaMemsetAsync(counter, 0, 4, sa); cudaDeviceSynchronize();
      cudaEventRecord(a0, sa); fill_kernel<<<148 * 2, 64, 16384, sa>>>(dst, chunks, counter); cudaEventRecord(a1, sa);
      if (kind >= 0) { cudaEventRecord(b0, sb); chain_kernel<<<ctas, 64, 0, sb>>>(buf, words, kind, ops, alu, lanes, sink); cudaEventRecord(b1, sb); }
      cudaDeviceSynchronize();
      float ta, tb = 0.f; cudaEventElapsedTime(&ta, a0, a1); if (kind >= 0) cudaEventElapsedTime(&tb, b0, b1);
      if (ta < best_a) { best_a = ta; best_b = tb; }
    }
    const double nops = kind >= 0 ? (double)ctas * 2 * lanes * ops : 0.0;
    printf("%-44s fill %.3f ms (%.0f GB/s)  co-runner %.3f ms  %.1f M ops  %.1f G ops/s\n", name, best_a,
           fill_bytes / best_a / 1e6, best_b, nops / 1e6, best_b > 0 ? nops / best_b / 1e6 : 0.0);
  };
  const size_t small_words = (size_t)8 << 20;               // 32 MB: L2-resident
  run(-1, 0, 0, 0, 0, 0, "fill alone");
  const int ctas = 148 * 9;                                 // the trace grid: 9 CTAs of 2 warps per SM
  for (int lanes : {6, 22}) {
    char nm[128];
    const int ops = lanes == 6 ? 12000 : 3300;              // ~19 M operations per run either way
    const char* kn[5] = {"red.or", "st", "ld (L1)", "ld.cg", "alu only"};
    for (int kind = 0; kind < 5; ++kind) {
      snprintf(nm, sizeof nm, "%s, 32 MB buffer, %d lanes/warp", kn[kind], lanes); run(kind, small_words, ops, 40, lanes, ctas, nm);
      if (kind < 4) { snprintf(nm, sizeof nm, "%s, 4 GB buffer, %d lanes/warp", kn[kind], lanes); run(kind, big_words, ops, 40, lanes, ctas, nm); }
    }
  }
  // the same 19 M red.or operations with less / more ALU work between them
  run(0, small_words, 12000, 0, 6, ctas, "red.or, 32 MB, 6 lanes, no ALU between");
  run(0, small_words, 12000, 200, 6, ctas, "red.or, 32 MB, 6 lanes, 200 ALU between");
  run(4, small_words, 12000, 200, 6, ctas, "alu only, 6 lanes, 200 ALU between");
  return 0;
}
