// Device memory for the full-frame bit-planes as a COMPRESSIBLE allocation.
//
// The planes of this path are zeros almost everywhere (97 % at 2048^2 with ~100-pixel instances).
// B200 compresses such data on the way from L2 to HBM when the PAGES are allocated compressible
// (CUDA virtual memory management, CU_MEM_ALLOCATION_COMP_GENERIC; "compute data compression"):
// the same TMA bulk stores then fill 16 GB at 8 470 instead of 7 420 GB/s and a read-back runs at
// 9 770 instead of 6 960 GB/s (tools/fill_compress.cu, profiles/r02_fill_compress_microbench.txt);
// any kernel or copy sees ordinary memory.  cudaMalloc / torch cannot hand out such pages, so the
// library offers the two calls below -- the only entry points that allocate; every kernel keeps
// taking caller-owned pointers and works on ordinary memory just the same.
//
// Stands in for the allocation behind Detectron2's paste_masks_in_image output
// (`img_masks = torch.zeros(N, img_h, img_w, ...)`, layers/mask_ops.py), reached from
// nn_inference.py:372.  The driver entry points are looked up through the runtime
// (cudaGetDriverEntryPoint), so the library does not link against libcuda and still loads on a
// machine without a driver (the CPU test tier).
#include <cstdint>
#include <mutex>
#include <unordered_map>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../include/uwcv.h"

namespace {

struct Mapping {
  CUmemGenericAllocationHandle handle;
  size_t size;
};
std::mutex g_mu;
std::unordered_map<void*, Mapping> g_maps;

struct Driver {
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemGetAllocationPropertiesFromHandle)(CUmemAllocationProp*, CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  bool ok = false;
};

template <class F>
bool entry(const char* name, F& fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  fn = reinterpret_cast<F>(p);
  return true;
}

const Driver& driver() {
  static Driver d = [] {
    Driver x;
    x.ok = entry("cuDeviceGet", x.DeviceGet) && entry("cuDeviceGetAttribute", x.DeviceGetAttribute) &&
           entry("cuMemGetAllocationGranularity", x.MemGetAllocationGranularity) &&
           entry("cuMemCreate", x.MemCreate) &&
           entry("cuMemGetAllocationPropertiesFromHandle", x.MemGetAllocationPropertiesFromHandle) &&
           entry("cuMemAddressReserve", x.MemAddressReserve) && entry("cuMemMap", x.MemMap) &&
           entry("cuMemSetAccess", x.MemSetAccess) && entry("cuMemUnmap", x.MemUnmap) &&
           entry("cuMemRelease", x.MemRelease) && entry("cuMemAddressFree", x.MemAddressFree);
    return x;
  }();
  return d;
}

}  // namespace

extern "C" {

int uwcv_planes_alloc(size_t bytes, void** ptr, int* compressed) {
  if (!ptr) return UWCV_E_NULL;
  *ptr = nullptr;
  if (compressed) *compressed = 0;
  if (bytes == 0) return UWCV_E_SHAPE;
  int ordinal = 0;
  if (cudaGetDevice(&ordinal) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) {   // (primary context up)
    cudaGetLastError();
    return UWCV_E_LAUNCH;
  }
  const Driver& d = driver();
  if (!d.ok) return UWCV_E_LAUNCH;
  CUdevice dev;
  if (d.DeviceGet(&dev, ordinal) != CUDA_SUCCESS) return UWCV_E_LAUNCH;
  int can = 0;
  if (d.DeviceGetAttribute(&can, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, dev) != CUDA_SUCCESS) can = 0;
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = ordinal;
  prop.allocFlags.compressionType = can ? CU_MEM_ALLOCATION_COMP_GENERIC : CU_MEM_ALLOCATION_COMP_NONE;
  size_t gran = 0;
  if (d.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0)
    return UWCV_E_LAUNCH;
  const size_t size = (bytes + gran - 1) / gran * gran;
  CUmemGenericAllocationHandle h;
  if (d.MemCreate(&h, size, &prop, 0) != CUDA_SUCCESS) return UWCV_E_WORKSPACE;      // out of memory
  CUmemAllocationProp got = {};
  const bool granted = d.MemGetAllocationPropertiesFromHandle(&got, h) == CUDA_SUCCESS &&
                       got.allocFlags.compressionType == CU_MEM_ALLOCATION_COMP_GENERIC;
  CUdeviceptr va = 0;
  if (d.MemAddressReserve(&va, size, 0, 0, 0) != CUDA_SUCCESS) { d.MemRelease(h); return UWCV_E_LAUNCH; }
  if (d.MemMap(va, size, 0, h, 0) != CUDA_SUCCESS) { d.MemAddressFree(va, size); d.MemRelease(h); return UWCV_E_LAUNCH; }
  CUmemAccessDesc acc = {};
  acc.location = prop.location;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  if (d.MemSetAccess(va, size, &acc, 1) != CUDA_SUCCESS) {
    d.MemUnmap(va, size); d.MemAddressFree(va, size); d.MemRelease(h);
    return UWCV_E_LAUNCH;
  }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_maps[reinterpret_cast<void*>(va)] = Mapping{h, size};
  }
  *ptr = reinterpret_cast<void*>(va);
  if (compressed) *compressed = granted ? 1 : 0;
  return UWCV_OK;
}

int uwcv_planes_free(void* ptr) {
  if (!ptr) return UWCV_OK;
  Mapping m;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_maps.find(ptr);
    if (it == g_maps.end()) return UWCV_E_SHAPE;           // not from uwcv_planes_alloc
    m = it->second;
    g_maps.erase(it);
  }
  const Driver& d = driver();
  if (!d.ok) return UWCV_E_LAUNCH;
  // (the caller has made sure no kernel or copy still uses the range, as with cudaFree)
  const CUdeviceptr va = reinterpret_cast<CUdeviceptr>(ptr);
  bool ok = d.MemUnmap(va, m.size) == CUDA_SUCCESS;
  ok = (d.MemRelease(m.handle) == CUDA_SUCCESS) && ok;
  ok = (d.MemAddressFree(va, m.size) == CUDA_SUCCESS) && ok;
  return ok ? UWCV_OK : UWCV_E_LAUNCH;
}

}  // extern "C"
