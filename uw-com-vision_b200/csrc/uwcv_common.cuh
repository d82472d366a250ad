// Shared definitions for the uwcv sm_100a kernels.
//
// Hot path: the post-inference stage of uw-com-vision's Mask R-CNN pipeline
// (/root/reference/nn_inference.py:371-459 and the Detectron2 post-process it
// reaches through predictor(im), :372).  See DESIGN.md for the data layout.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

// -DUWCV_CHECK (lib/libuwcv_check.so): device-side bounds traps on every tile / plane / band /
// extremes / mark index the kernels compute -- the memory-safety substitute for compute-sanitizer
// (closed on the GPU pool).  The GPU test set and tools/parity_sweep.py are run against it
// (UWCV_TEST_VARIANT=check); a violated bound prints its source line and traps the kernel.
#ifdef UWCV_CHECK
#include <cstdio>
#define UWCV_BOUND(idx, limit)                                                                   \
  do {                                                                                           \
    if (!((long long)(idx) >= 0 && (long long)(idx) < (long long)(limit))) {                     \
      printf("UWCV_CHECK %s:%d: index %lld outside [0, %lld)\n", __FILE__, __LINE__,             \
             (long long)(idx), (long long)(limit));                                              \
      __trap();                                                                                  \
    }                                                                                            \
  } while (0)
#else
#define UWCV_BOUND(idx, limit) do { } while (0)
#endif

namespace uwcv {

constexpr int kMaskSide = 28;              // Detectron2 / torchvision mask head output side
constexpr int kPad = 2;                    // zero frame so that taps at -2..29 need no branch
constexpr int kMaskPitch = kMaskSide + 2 * kPad;   // 32
constexpr int kNumInt = 20;                // int64 columns per row  
constexpr int kNumFloat = 30;              // float64 columns per row

// int64 column indices
enum IntCol {
  I_IMAGE = 0, I_INST, I_CLASS, I_VALID, I_NCONT, I_AREA, I_BX0, I_BY0, I_BX1, I_BY1,
  I_M10, I_M01, I_M20, I_M11, I_M02, I_M30, I_M21, I_M12, I_M03, I_NPTS
};
// float64 column indices
enum FloatCol {
  F_SCORE = 0, F_CX, F_CY, F_MU20, F_MU11, F_MU02, F_MU30, F_MU21, F_MU12, F_MU03,
  F_EQD, F_ELL_MAJOR, F_ELL_MINOR, F_ELL_THETA, F_CAREA, F_PERIM,
  F_RCX, F_RCY, F_RW, F_RH, F_RANGLE,
  F_FERET, F_ASPECT, F_ROUND, F_CIRC, F_SPHER, F_LENGTH, F_WIDTH, F_CED, F_CHORDS
};

// error codes (C ABI returns these; 0 = ok)
enum Err {
  OK = 0, E_NULL = -1, E_SHAPE = -2, E_ALIGN = -3, E_THRESH = -4, E_WORKSPACE = -5,
  E_LAUNCH = -6, E_CAPACITY = -7, E_TOO_LARGE = -8
};

// Per-instance tile: the word-aligned window of the frame that can hold set pixels.
struct __align__(16) TileDesc {
  int32_t wx0;        // first 32-pixel word column of the tile (pixel x origin = wx0 * 32)
  int32_t y0;         // first pixel row
  int32_t tw;         // width in words
  int32_t th;         // height in rows  (0 => empty tile)
  int64_t word_off;   // offset of the tile's first word in the M / V / G planes
  int64_t row_off;    // offset of the tile's first row in the per-row trace scratch
};

// Where the 28 x 28 mask of instance i comes from: base + i * stride (+ channel * 784)
struct MaskSource {
  int64_t stride;        // floats between consecutive instances (784 for plain probabilities)
  int channels;          // 1: the only channel; > 1: channel classes[i] + channel_offset
  int channel_offset;
  int logits;            // 1: apply the sigmoid of mask_rcnn_inference while staging
};

// Fused all-gather of the measurement rows: every rank's trace kernel stores its finished rows
// into the tables of ALL ranks (peer memory over NVLink), at its own row offset.
constexpr int kMaxPeers = 16;
struct GatherDst {
  int world;                       // 0: no gather
  int only;                        // -1: every rank's table; p: the table of rank p alone
  int64_t row_base;                // first row of this rank in the gathered tables
  int64_t* rows_i[kMaxPeers];      // [total_rows, kNumInt]   per peer (this rank included)
  double* rows_f[kMaxPeers];       // [total_rows, kNumFloat] per peer
};

// (reserved header record per instance: 32 bytes, kept so that the workspace carve-up -- and with it
// uwcv_workspace_bytes -- does not change; the fused-trace experiment of round 2 handed its results
// over through it)
struct __align__(16) TraceRec {
  long long a2;
  double perim;
  int32_t best_y, best_ymax, npts, ncont;
};

// One shared-memory / L1 split for every kernel of the measure pipeline.
//
// The plane fill is a persistent kernel of 2 CTAs per SM with 53 KB of shared memory each.  An SM
// changes its shared-memory carve-out only when it is empty, and a persistent fill CTA keeps it
// occupied for the whole launch: when the fill's CTAs arrive on SMs that a co-running kernel with
// a small carve-out (layout, mark clearing, the border trace: no or little shared memory) reached a
// microsecond earlier, only ONE fill CTA fits per SM and the second never gets in -- the whole
// fill then runs at 60 % of its rate (6.6 instead of 4.3 ms per 64 000 instances), and which of
// the two happens is decided by the launch order of that microsecond (CUPTI timelines,
// profiles/r02_carveout_*.txt).  With every kernel that can be resident beside the fill asking for
// the same split (all of it shared memory), the carve-out never stands in the way.  The driver
// ignores the hint for a kernel that uses no shared memory at all, so the border trace of the split
// pipeline also asks for 16 KB it never touches (kContourCarveoutBytes, contour.cu).
// (`on` = false hands the choice back to the driver: the border trace of a call that writes no
// planes has no fill beside it and keeps the large L1 -- 1.74 against 1.90 ms per 64 000 instances.)
template <class Kernel>
inline void share_carveout(Kernel kernel, bool on = true) {
#ifdef UWCV_TUNING
  if (getenv("UWCV_NO_CARVEOUT")) return;
#endif
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                           on ? (int)cudaSharedmemCarveoutMaxShared
                              : (int)cudaSharedmemCarveoutDefault) != cudaSuccess)
    cudaGetLastError();                 // a hint: never fails a launch
}

// Workspace carve-up, computed identically on host and device.
constexpr int kLayoutThreads = 1024;   // instances per layout CTA

struct Workspace {
  TileDesc* desc;      // [N]
  TraceRec* rec;       // [N]  (reserved)
  int32_t* order;      // [N]  paste order of a launch's range: large tiles first (their traces are the long ones)
  int64_t* block_sums; // [2 * ceil(N / 1024)]  per-CTA (tile words, tile rows) of the layout
  unsigned int* sched; // [8]  work counters of the paste kernel (fill, tile, CTAs done, ...); zero between launches
  uint32_t* M;         // [cap_words]  mask bits
  uint32_t* V;         // [cap_words]  border-visited marks
  uint32_t* G;         // [cap_words]  "right neighbour was background" marks (negative marks)
  uint32_t* scratch;   // [4 * cap_words]  two sets of per-row extremes / hull chains (x | y << 16)
  int64_t cap_words;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline size_t layout_blocks(int64_t n) {
  return (size_t)((n + kLayoutThreads - 1) / kLayoutThreads);
}
__host__ __device__ inline size_t header_bytes(int64_t n) {
  return align_up((size_t)n * sizeof(TileDesc), 256) + align_up((size_t)n * sizeof(TraceRec), 256) +
         align_up((size_t)n * sizeof(int32_t), 256) + align_up(layout_blocks(n) * 16 + 16, 256) + 256;
}
__host__ __device__ inline size_t workspace_bytes(int64_t n, int64_t tile_words) {
  return header_bytes(n) + (size_t)tile_words * 28 + 256;
}

__host__ __device__ inline Workspace carve(void* ws, size_t ws_bytes, int64_t n) {
  Workspace w;
  char* p = (char*)ws;
  w.desc = (TileDesc*)p;
  size_t o = align_up((size_t)n * sizeof(TileDesc), 256);
  w.rec = (TraceRec*)(p + o);
  o += align_up((size_t)n * sizeof(TraceRec), 256);
  w.order = (int32_t*)(p + o);
  o += align_up((size_t)n * sizeof(int32_t), 256);
  w.block_sums = (int64_t*)(p + o);
  size_t d = header_bytes(n);
  w.sched = (unsigned int*)(p + d - 256);
  int64_t cap = ws_bytes > d + 256 ? (int64_t)((ws_bytes - d - 256) / 28) : 0;
  cap &= ~(int64_t)3;                         // keep every plane 16-byte aligned
  w.cap_words = cap;
  w.M = (uint32_t*)(p + d);
  w.V = w.M + cap;
  w.G = w.V + cap;
  w.scratch = w.G + cap;
  return w;
}

// plane row stride in 32-bit words: ceil(W / 32) rounded up to a multiple of 4 so that
// every row (and every plane) starts on a 16-byte boundary for the bulk stores.
__host__ __device__ inline int plane_row_words(int W) { return ((W + 31) / 32 + 3) & ~3; }

}  // namespace uwcv
