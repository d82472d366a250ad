// Union / connected-component mode: the reference-literal GetMask_Contours.
//
// nn_inference.py:394-459 ORs the masks of the selected class into ONE image, takes the
// external contours of that union and measures every contour with contourArea >= 100, left
// to right.  Touching instances merge and an instance with several blobs yields several rows.
//
// Instances whose (1-pixel dilated) pixel boxes do not overlap cannot share a component nor
// nest inside one another's holes, so the host groups the instances of an image by box
// overlap and every group is processed independently:
//   1. union_or_kernel      OR the member tiles (paste kernel output) into the group tile
//   2. union_trace_kernel   one thread per group: raster scan + outer-border following of the
//                           group tile with OpenCV's RETR_EXTERNAL rule; one record per contour
//   3. union_scan_kernel    exclusive scan of the rows each contour spans -> scratch offsets
//   4. union_describe_kernel one thread per contour: re-trace collecting per-row extremes,
//                           bounding rect, hull, min-area rectangle, descriptor block
// The device code of steps 2 and 4 is the same as the per-instance kernel's
// (contour_common.cuh); ordering and the area cut are host work on the small row table.
#include "contour_common.cuh"

namespace uwcv {

constexpr int kUnionThreads = 64;
constexpr int kURowI = 10;      // int64 columns of a union row
constexpr int kURowF = 16;      // float64 columns of a union row

struct UnionRecords {
  int32_t* rec;        // [cap][8]  group, sx, sy, ymax, npts, -, -, -
  int64_t* a2;         // [cap]     |twice-area|
  double* perim;       // [cap]
  int64_t* ext_off;    // [cap]     offset of the contour's extremes in the pool (u32 units)
  uint32_t* pool;      // [2 * ext_rows_cap]
  int64_t cap, ext_rows_cap;
};

__host__ __device__ inline size_t union_ws_bytes(int64_t cap, int64_t ext_rows_cap) {
  return align_up((size_t)cap * 32, 256) + 3 * align_up((size_t)cap * 8, 256) +
         align_up((size_t)ext_rows_cap * 8, 256) + 256;
}

inline UnionRecords union_carve(void* ws, int64_t cap, int64_t ext_rows_cap) {
  UnionRecords r;
  char* p = (char*)ws;
  r.rec = (int32_t*)p;      p += align_up((size_t)cap * 32, 256);
  r.a2 = (int64_t*)p;       p += align_up((size_t)cap * 8, 256);
  r.perim = (double*)p;     p += align_up((size_t)cap * 8, 256);
  r.ext_off = (int64_t*)p;  p += align_up((size_t)cap * 8, 256);
  r.pool = (uint32_t*)p;
  r.cap = cap; r.ext_rows_cap = ext_rows_cap;
  return r;
}

// ---- 1. OR member tiles into group tiles ---------------------------------------------------
__global__ void __launch_bounds__(256)
union_or_kernel(int64_t n, const TileDesc* __restrict__ desc, const uint32_t* __restrict__ M,
                const int32_t* __restrict__ member_group, const TileDesc* __restrict__ gdesc,
                uint32_t* __restrict__ gM, const int64_t* __restrict__ gcount) {
  if (gcount && gcount[3] != 0) return;                  // group planes too small: the caller retries
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp0; i < n; i += nwarps) {
    const int g = member_group[i];
    if (g < 0) continue;
    const TileDesc d = desc[i];
    const TileDesc gd = gdesc[g];
    const int words = d.tw * d.th;
    for (int k = lane; k < words; k += 32) {
      const uint32_t w = __ldg(M + d.word_off + k);
      if (!w) continue;
      const int r = k / d.tw, c = k - r * d.tw;
      const int dr = d.y0 + r - gd.y0, dc = d.wx0 + c - gd.wx0;
      if ((unsigned)dr < (unsigned)gd.th && (unsigned)dc < (unsigned)gd.tw)
        atomicOr(gM + gd.word_off + (int64_t)dr * gd.tw + dc, w);
    }
  }
}

// ---- 2. trace every external contour of every group -------------------------------------------
__global__ void __launch_bounds__(kUnionThreads)
union_trace_kernel(int64_t G, const int64_t* __restrict__ gcount, int lanes,
                   const TileDesc* __restrict__ gdesc,
                   const uint32_t* __restrict__ gM, uint32_t* __restrict__ gV,
                   uint32_t* __restrict__ gG, UnionRecords R, int64_t* __restrict__ counters) {
  if (gcount) {                                          // groups were formed on the device
    if (gcount[3] != 0) { if (blockIdx.x == 0 && threadIdx.x == 0) counters[1] = E_CAPACITY; return; }
    G = gcount[0];
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * kUnionThreads + threadIdx.x) >> 5;
  const int64_t g = warp * lanes + lane;
  const bool live = lane < lanes && g < G;
  TileDesc d;
  d.wx0 = d.y0 = d.tw = d.th = 0; d.word_off = d.row_off = 0;
  if (live) d = gdesc[g];
  TileView t;
  t.M = gM + d.word_off; t.V = gV + d.word_off; t.G = gG + d.word_off;
  t.tw = d.tw; t.th = d.th;
  const bool work = live && d.th > 0;
  int y = 0;
  const int yhi = work ? d.th - 1 : -1;
  int wi = 0;
  uint64_t carry = 0, start_mask = 0, cand = 0, vpair = 0, gpair = 0;
  int cy_ = 0, cwi = 0, sy = 0, sx = 0;
  enum { kScan = 0, kTrace = 1, kDone = 2 };
  int state = work ? kScan : kDone;
  Trace tr;
  tr.active = false;
  auto load_pair = [&](const uint32_t* plane, int yy, int w0) -> uint64_t {
    const uint32_t* row = plane + yy * t.tw;
    const uint32_t lo = row[w0];
    const uint32_t hi = (w0 + 1 < t.tw) ? row[w0 + 1] : 0u;
    return (uint64_t)lo | ((uint64_t)hi << 32);
  };
  auto sign_of_top = [](uint64_t v, uint64_t gg) -> int {
    const int top = 63 - __clzll((long long)v);
    return ((gg >> top) & 1ull) ? -1 : +1;
  };
  uint64_t m_next = (state == kScan && y <= yhi) ? load_pair(t.M, y, 0) : 0ull;
  bool fresh = true;
  while (__any_sync(kFull, state != kDone)) {
    bool finished = false;
    if (state == kScan) {
      if (fresh) {
        if (y > yhi) {
          state = kDone;
        } else {
          const uint64_t m = m_next;
          start_mask = m & ~((m << 1) | carry);
          carry = m >> 63;
          cy_ = y; cwi = wi;
          wi += 2;
          if (wi >= t.tw) { wi = 0; ++y; carry = 0; }
          if (y <= yhi) m_next = load_pair(t.M, y, wi);
          vpair = start_mask ? load_pair(t.V, cy_, cwi) : 0ull;
          cand = start_mask & ~vpair;
          gpair = cand ? load_pair(t.G, cy_, cwi) : 0ull;
          fresh = cand == 0;
        }
      }
      if (state == kScan && !fresh) {
        bool start = false;
        int b = 0;
        while (cand) {
          b = __ffsll((long long)cand) - 1;
          cand &= cand - 1;
          const uint64_t below = vpair & ((1ull << b) - 1ull);
          const int sgn = below ? sign_of_top(below, gpair) : last_mark_before(t, cwi, cy_);
          if (sgn <= 0) { start = true; break; }
        }
        if (start) {
          sy = cy_; sx = cwi * 32 + b;
          trace_begin<true, false>(t, tr, sx, sy, nullptr, nullptr);
          if (tr.active) state = kTrace; else finished = true;
        } else {
          fresh = true;
        }
      }
    } else if (state == kTrace) {
      trace_step<true, false>(t, tr, nullptr, nullptr);
      if (!tr.active) { finished = true; state = kScan; }
    }
    if (finished) {
      const long long slot = (long long)atomicAdd((unsigned long long*)&counters[0], 1ull);
      if (slot < R.cap) {
        int32_t* rec = R.rec + slot * 8;
        rec[0] = (int32_t)g; rec[1] = sx; rec[2] = sy; rec[3] = tr.ymax; rec[4] = tr.npts;
        R.a2[slot] = tr.area2 < 0 ? -tr.area2 : tr.area2;
        R.perim[slot] = tr.perim;
      } else {
        counters[1] = E_CAPACITY;
      }
      vpair = load_pair(t.V, cy_, cwi); gpair = load_pair(t.G, cy_, cwi);
      cand &= ~vpair;
    }
  }
}

// ---- 3. scratch offsets of the contours' per-row extremes (single CTA) ---------------------------
__global__ void __launch_bounds__(1024)
union_scan_kernel(UnionRecords R, int64_t* __restrict__ counters) {
  __shared__ int64_t s[1024];
  const int t = threadIdx.x;
  int64_t nrec = counters[0];
  if (nrec > R.cap) nrec = R.cap;
  const int64_t per = (nrec + 1023) / 1024;
  const int64_t lo = (int64_t)t * per, hi = lo + per < nrec ? lo + per : nrec;
  int64_t sum = 0;
  for (int64_t k = lo; k < hi; ++k) sum += 2 * (int64_t)(R.rec[k * 8 + 3] - R.rec[k * 8 + 2] + 1);
  s[t] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    int64_t a = 0;
    if (t >= off) a = s[t - off];
    __syncthreads();
    s[t] += a;
    __syncthreads();
  }
  int64_t run = s[t] - sum;
  for (int64_t k = lo; k < hi; ++k) {
    R.ext_off[k] = run;
    run += 2 * (int64_t)(R.rec[k * 8 + 3] - R.rec[k * 8 + 2] + 1);
  }
  if (t == 0) {
    counters[2] = s[1023] / 2;                           // extreme rows needed
    if (s[1023] > 2 * R.ext_rows_cap) counters[1] = E_CAPACITY;
  }
}

// ---- 4. descriptors of every contour --------------------------------------------------------------
__global__ void __launch_bounds__(kUnionThreads)
union_describe_kernel(int lanes, const TileDesc* __restrict__ gdesc,
                      const int32_t* __restrict__ group_image, const uint32_t* __restrict__ gM,
                      UnionRecords R, double ppm, int64_t* __restrict__ rows_i,
                      double* __restrict__ rows_f, const int64_t* __restrict__ counters) {
  if (counters[1] != 0) return;                        // capacity overflow: the host retries
  int64_t nrec = counters[0];
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * kUnionThreads + threadIdx.x) >> 5;
  const int64_t k = warp * lanes + lane;
  const bool live = lane < lanes && k < nrec;
  int g = 0, sx = 0, sy = 0, ymax = -1, npts = 0;
  if (live) {
    const int32_t* rec = R.rec + k * 8;
    g = rec[0]; sx = rec[1]; sy = rec[2]; ymax = rec[3]; npts = rec[4];
  }
  TileDesc d;
  d.wx0 = d.y0 = d.tw = d.th = 0; d.word_off = d.row_off = 0;
  if (live) d = gdesc[g];
  TileView t;
  t.M = gM + d.word_off; t.V = nullptr; t.G = nullptr;
  t.tw = d.tw; t.th = d.th;
  const int rows = ymax - sy + 1;
  // extremes indexed by tile row: shift the pool pointers so that [sy] is the first entry
  uint32_t* ext_l = R.pool + (live ? R.ext_off[k] : 0) - sy;
  uint32_t* ext_r = ext_l + (live ? rows : 0);
  Trace tr;
  tr.active = false;
  if (live) trace_begin<false, true>(t, tr, sx, sy, ext_l, ext_r);
  while (__any_sync(kFull, tr.active)) {
    if (tr.active) trace_step<false, true>(t, tr, ext_l, ext_r);
  }
  // bounding rect of the contour points
  int xmin = 0x7fffffff, xmax = -1;
  for (int yy = sy; __any_sync(kFull, yy <= ymax); ++yy) {
    if (yy <= ymax) {
      xmin = min(xmin, (int)ext_l[yy]);
      xmax = max(xmax, (int)ext_r[yy]);
    }
  }
  double* rf = rows_f + (live ? k : 0) * kURowF;
  describe_contour(live, ext_l, ext_r, sy, ymax, d.wx0 * 32, d.y0, live ? R.a2[k] : 0,
                   live ? R.perim[k] : 0.0, ppm, rf);
  if (!live) return;
  int64_t* ri = rows_i + k * kURowI;
  ri[0] = group_image[g];
  ri[1] = g;
  ri[2] = d.wx0 * 32 + sx;           // start pixel (raster-first pixel of the component)
  ri[3] = d.y0 + sy;
  ri[4] = d.wx0 * 32 + xmin;         // cv2.boundingRect of the contour
  ri[5] = d.y0 + sy;
  ri[6] = xmax - xmin + 1;
  ri[7] = rows;
  ri[8] = npts;
  ri[9] = 1;
}

// ---- 0. groups of instances on the device --------------------------------------------------------
// Instances of one image whose 1-pixel-dilated pixel boxes overlap (transitively) form a group.
// One CTA per image: minimum-label propagation over the overlap graph (every instance looks at
// every other instance of its image: N^2 / 1024 box tests per thread and round, boxes in shared
// memory) with pointer jumping, until no label changes; the roots become the groups, numbered
// image-major in instance order (deterministic), each with the union of its members' boxes as
// its tile.  A single-CTA scan turns the per-image counts into global group ids / word offsets.
constexpr int kGroupThreads = 1024;
constexpr int kGroupSmemBoxes = 6144;                  // boxes (16 B) + labels (4 B) per image in shared memory

struct GroupScratch {
  int32_t* label;       // [N]   propagation labels of images too large for shared memory
  int32_t* local_gid;   // [N]   group of the instance inside its image (-1: none)
  TileDesc* tmp;        // [N]   per image: tiles of its groups, local word offsets
  int64_t* per_image;   // [B][4] groups, words, first instance, -
};

__host__ __device__ inline size_t group_scratch_bytes(int64_t n, int B) {
  return 2 * align_up((size_t)n * 4, 256) + align_up((size_t)n * sizeof(TileDesc), 256) +
         align_up((size_t)B * 32, 256) + 256;
}
inline GroupScratch group_carve(void* ws, int64_t n, int B) {
  GroupScratch g;
  char* p = (char*)ws;
  g.label = (int32_t*)p;      p += align_up((size_t)n * 4, 256);
  g.local_gid = (int32_t*)p;  p += align_up((size_t)n * 4, 256);
  g.tmp = (TileDesc*)p;       p += align_up((size_t)n * sizeof(TileDesc), 256);
  g.per_image = (int64_t*)p;
  return g;
}

__global__ void __launch_bounds__(kGroupThreads)
union_group_kernel(const int64_t* __restrict__ rows_i, int64_t n, int64_t image_base,
                   GroupScratch gs) {
  extern __shared__ int4 s_box[];                        // dilated pixel boxes (x0-1, y0-1, x1+1, y1+1)
  __shared__ int s_scan[kGroupThreads];
  const int b = blockIdx.x, tid = threadIdx.x;
  // instances of image b are consecutive: first row with image index >= base + b (binary search)
  auto lower = [&](int64_t img) {
    int64_t a = 0, z = n;
    while (a < z) { const int64_t m = (a + z) >> 1; if (rows_i[m * kNumInt + I_IMAGE] < img) a = m + 1; else z = m; }
    return a;
  };
  const int64_t lo = lower(image_base + b), hi = lower(image_base + b + 1);
  const int nb = (int)(hi - lo);
  int* s_label = reinterpret_cast<int*>(s_box + (nb <= kGroupSmemBoxes ? nb : 0));
  const bool staged = nb <= kGroupSmemBoxes;
  auto box_of = [&](int i) -> int4 {
    const int64_t* r = rows_i + (lo + i) * kNumInt;
    if (r[I_VALID] != 1) return make_int4(1, 1, -1, -1);          // empty: overlaps nothing
    return make_int4((int)r[I_BX0] - 1, (int)r[I_BY0] - 1, (int)r[I_BX1] + 1, (int)r[I_BY1] + 1);
  };
  int32_t* g_label = gs.label + lo;                      // labels live here when the image is too big
  for (int i = tid; i < nb; i += kGroupThreads) {
    const int4 bx = box_of(i);
    if (staged) { s_box[i] = bx; s_label[i] = bx.z >= bx.x ? i : -1; }
    else g_label[i] = bx.z >= bx.x ? i : -1;
  }
  __syncthreads();
  volatile int* label = staged ? (volatile int*)s_label : (volatile int*)g_label;
  for (;;) {
    int changed = 0;
    for (int i = tid; i < nb; i += kGroupThreads) {
      int m = label[i];
      if (m < 0) continue;
      const int4 a = staged ? s_box[i] : box_of(i);
      for (int j = 0; j < nb; ++j) {
        const int4 c = staged ? s_box[j] : box_of(j);
        // [x0-1, x1+1] x [y0-1, y1+1] intersect  <=>  8-adjacent or overlapping pixel boxes
        if (a.x <= c.z && c.x <= a.z && a.y <= c.w && c.y <= a.w && c.z >= c.x) {
          const int lj = label[j];
          if (lj < m) m = lj;
        }
      }
      // pointer jumping: follow the labels down to a root
      for (int k = 0; k < 8; ++k) { const int up = label[m]; if (up >= m) break; m = up; }
      if (m < label[i]) { label[i] = m; changed = 1; }
    }
    if (!__syncthreads_or(changed)) break;
  }
  // At the fixed point every label is the smallest instance index of its group, and that
  // instance (the root) carries its own index.  Roots -> dense local group ids in instance
  // order (block scan of the root flags); the group's box starts empty.
  int base = 0;
  for (int i0 = 0; i0 < nb; i0 += kGroupThreads) {
    const int i = i0 + tid;
    const int is_root = (i < nb && label[i] == i) ? 1 : 0;
    s_scan[tid] = is_root;
    __syncthreads();
    for (int off = 1; off < kGroupThreads; off <<= 1) {
      int v = 0;
      if (tid >= off) v = s_scan[tid - off];
      __syncthreads();
      s_scan[tid] += v;
      __syncthreads();
    }
    if (is_root) {
      const int gid = base + s_scan[tid] - 1;
      TileDesc d;                                        // pixel extents first, words further down
      d.wx0 = 0x7fffffff; d.y0 = 0x7fffffff; d.tw = -1; d.th = -1; d.word_off = 0; d.row_off = 0;
      gs.tmp[lo + gid] = d;
      gs.local_gid[lo + i] = gid;
    }
    base += s_scan[kGroupThreads - 1];
    __syncthreads();
  }
  const int ngroups = base;
  // members take their root's id; the group box is the union of the members' pixel boxes
  for (int i = tid; i < nb; i += kGroupThreads) {
    const int root = label[i];
    if (root < 0) { gs.local_gid[lo + i] = -1; continue; }
    const int gid = gs.local_gid[lo + root];             // (roots were written before the barrier)
    if (root != i) gs.local_gid[lo + i] = gid;
    const int64_t* r = rows_i + (lo + i) * kNumInt;
    TileDesc* d = gs.tmp + lo + gid;
    atomicMin(&d->wx0, (int)r[I_BX0]);
    atomicMin(&d->y0, (int)r[I_BY0]);
    atomicMax(&d->tw, (int)r[I_BX1]);
    atomicMax(&d->th, (int)r[I_BY1]);
  }
  __syncthreads();
  // tiles in words + local word offsets (serial over the image's groups: a few hundred)
  __syncthreads();
  if (tid == 0) {
    int64_t words = 0;
    for (int g = 0; g < ngroups; ++g) {
      TileDesc d = gs.tmp[lo + g];
      const int x0 = d.wx0, y0 = d.y0, x1 = d.tw, y1 = d.th;
      d.wx0 = x0 >> 5; d.tw = (x1 >> 5) - d.wx0 + 1; d.y0 = y0; d.th = y1 - y0 + 1;
      d.word_off = words; d.row_off = 0;
      words += (int64_t)d.tw * d.th;
      gs.tmp[lo + g] = d;
    }
    gs.per_image[b * 4 + 0] = ngroups;
    gs.per_image[b * 4 + 1] = words;
    gs.per_image[b * 4 + 2] = lo;
    gs.per_image[b * 4 + 3] = nb;
  }
}

// global group ids / word offsets: exclusive scan over the images (single thread: B is small),
// then every image's groups and members are written at their final places
__global__ void __launch_bounds__(1024)
union_group_finish_kernel(int B, int64_t image_base, GroupScratch gs, int32_t* __restrict__ member_group,
                          TileDesc* __restrict__ gdesc, int32_t* __restrict__ group_image,
                          int64_t group_words_cap, int64_t* __restrict__ gcount) {
  const int tid = threadIdx.x;
  int64_t gbase = 0, wbase = 0;
  for (int b = 0; b < B; ++b) {
    const int64_t ng = gs.per_image[b * 4 + 0], words = gs.per_image[b * 4 + 1];
    const int64_t lo = gs.per_image[b * 4 + 2], nb = gs.per_image[b * 4 + 3];
    for (int64_t g = tid; g < ng; g += blockDim.x) {
      TileDesc d = gs.tmp[lo + g];
      d.word_off += wbase;
      gdesc[gbase + g] = d;
      group_image[gbase + g] = (int32_t)(image_base + b);
    }
    for (int64_t i = tid; i < nb; i += blockDim.x) {
      const int l = gs.local_gid[lo + i];
      member_group[lo + i] = l < 0 ? -1 : (int32_t)(gbase + l);
    }
    gbase += ng; wbase += words;
  }
  if (tid == 0) {
    gcount[0] = gbase;
    gcount[1] = (wbase + 3) & ~(int64_t)3;
    gcount[2] = 0;
    gcount[3] = (((wbase + 3) & ~(int64_t)3) > group_words_cap) ? (int64_t)E_CAPACITY : 0;
  }
}

size_t union_group_workspace_bytes_host(int64_t n, int B) { return group_scratch_bytes(n, B); }

cudaError_t launch_union_group(const int64_t* rows_i, int64_t n, int B, int64_t image_base, void* scratch,
                               int32_t* member_group, TileDesc* gdesc, int32_t* group_image,
                               int64_t group_words_cap, int64_t* gcount, cudaStream_t stream) {
  if (n == 0 || B == 0) return cudaMemsetAsync(gcount, 0, 4 * sizeof(int64_t), stream);
  GroupScratch gs = group_carve(scratch, n, B);
  const size_t dyn = (size_t)kGroupSmemBoxes * 20;
  cudaFuncSetAttribute(union_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  union_group_kernel<<<B, kGroupThreads, dyn, stream>>>(rows_i, n, image_base, gs);
  union_group_finish_kernel<<<1, 1024, 0, stream>>>(B, image_base, gs, member_group, gdesc, group_image,
                                                     group_words_cap, gcount);
  return cudaPeekAtLastError();
}

// gcount != NULL: the groups were formed on the device (launch_union_group); G is then the
// CAPACITY the grids are sized for (at most N groups) and gwords the capacity of the planes
cudaError_t launch_union(int64_t n, const Workspace& ws, const int32_t* member_group,
                         const TileDesc* gdesc, const int32_t* group_image, int64_t G,
                         uint32_t* gplanes, int64_t gwords, void* rec_ws, int64_t rec_cap,
                         int64_t ext_rows_cap, double ppm, int64_t* rows_i, double* rows_f,
                         int64_t* counters, int num_sms, cudaStream_t stream,
                         const int64_t* gcount) {
  cudaError_t e = cudaMemsetAsync(counters, 0, 4 * sizeof(int64_t), stream);
  if (e != cudaSuccess) return e;
  if (G == 0 || n == 0) return cudaSuccess;
  e = cudaMemsetAsync(gplanes, 0, (size_t)gwords * 12, stream);
  if (e != cudaSuccess) return e;
  uint32_t* gM = gplanes;
  uint32_t* gV = gplanes + gwords;
  uint32_t* gG = gplanes + 2 * gwords;
  UnionRecords R = union_carve(rec_ws, rec_cap, ext_rows_cap);
  {
    int64_t blocks = (n * 32 + 255) / 256;
    const int64_t cap = (int64_t)num_sms * 16;
    if (blocks > cap) blocks = cap;
    union_or_kernel<<<(unsigned)blocks, 256, 0, stream>>>(n, ws.desc, ws.M, member_group, gdesc, gM, gcount);
  }
  auto lanes_for = [&](int64_t items) {
    const int64_t resident = (int64_t)num_sms * 10 * (kUnionThreads / 32);
    int l = (int)((items + resident - 1) / resident);
    return l < 1 ? 1 : (l > 32 ? 32 : l);
  };
  {
    const int lanes = lanes_for(G);
    const int64_t warps = (G + lanes - 1) / lanes;
    const unsigned grid = (unsigned)((warps * 32 + kUnionThreads - 1) / kUnionThreads);
    union_trace_kernel<<<grid, kUnionThreads, 0, stream>>>(G, gcount, lanes, gdesc, gM, gV, gG, R, counters);
  }
  union_scan_kernel<<<1, 1024, 0, stream>>>(R, counters);
  {
    // the record count lives on the device: size the grid for the capacity
    const int lanes = lanes_for(rec_cap);
    const int64_t warps = (rec_cap + lanes - 1) / lanes;
    const unsigned grid = (unsigned)((warps * 32 + kUnionThreads - 1) / kUnionThreads);
    union_describe_kernel<<<grid, kUnionThreads, 0, stream>>>(lanes, gdesc, group_image, gM, R, ppm,
                                                              rows_i, rows_f, counters);
  }
  return cudaPeekAtLastError();
}

size_t union_workspace_bytes_host(int64_t rec_cap, int64_t ext_rows_cap) {
  return union_ws_bytes(rec_cap, ext_rows_cap);
}

}  // namespace uwcv
