// Device code shared by the per-instance contour kernel (contour.cu) and the union /
// connected-component mode (union.cu): bit-tile access, the register neighbourhood window,
// Suzuki-Abe border following as resumable micro-steps, convex hull, OpenCV's float32
// rotating calipers, boxPoints / order_points / descriptor block.
#pragma once
#include "uwcv_common.cuh"

namespace uwcv {

struct TileView {
  const uint32_t* M;
  uint32_t* V;
  uint32_t* G;
  int tw, th;
};

// Where a tile (mask bits, marks, per-row extremes) lives, as a policy of the trace code.  GlobalMem:
// the tile workspace in HBM / L2.  (Round 2 also ran the walk on a shared-memory arena inside the
// paste kernel through a second policy; it lost -- profiles/r02_fused_trace_sweep.txt -- and is gone.)
struct GlobalMem {
  static __device__ __forceinline__ uint32_t ldm(const uint32_t* p) { return __ldg(p); }   // mask bits: read-only
  static __device__ __forceinline__ uint32_t ld(const uint32_t* p) { return *p; }
  // marks written by ANOTHER lane of the warp (result-less atomics, performed in L2): read past L1
  static __device__ __forceinline__ uint32_t ldc(const uint32_t* p) { return __ldcg(p); }
  static __device__ __forceinline__ void st(uint32_t* p, uint32_t v) { *p = v; }
  static __device__ __forceinline__ void or_(uint32_t* p, uint32_t v) { atomicOr(p, v); }
  static __device__ __forceinline__ void min_(uint32_t* p, uint32_t v) { atomicMin(p, v); }
  static __device__ __forceinline__ void max_(uint32_t* p, uint32_t v) { atomicMax(p, v); }
};
// direction s: 0 = east, then counter-clockwise on a y-up plane (1 = x+1, y-1 on screen)
__device__ __forceinline__ int dir_dx(int s) { return (int)((0x901Au >> (2 * s)) & 3u) - 1; }
__device__ __forceinline__ int dir_dy(int s) { return (int)((0xA901u >> (2 * s)) & 3u) - 1; }

// 64-pixel window of tile row y starting at word wb (words outside the tile read as zero)
template <class Mem = GlobalMem>
__device__ __forceinline__ uint64_t load_row64(const TileView& t, int y, int wb) {
  if ((unsigned)y >= (unsigned)t.th) return 0ull;
  const uint32_t* row = t.M + y * t.tw;
  const uint32_t lo = ((unsigned)wb < (unsigned)t.tw) ? Mem::ldm(row + wb) : 0u;
  const uint32_t hi = ((unsigned)(wb + 1) < (unsigned)t.tw) ? Mem::ldm(row + wb + 1) : 0u;
  return (uint64_t)lo | ((uint64_t)hi << 32);
}

// Rows y-1, y, y+1 of the mask around the current border pixel, kept in registers so that
// the 8-neighbour search is pure ALU work; one 64-bit row is fetched per vertical step.
struct Window {
  uint64_t r0, r1, r2;
  int wb;                                  // first word of the window
  template <class Mem = GlobalMem>
  __device__ __forceinline__ void load(const TileView& t, int x, int y) {
    wb = (x >> 5) - (((x & 31) < 16) ? 1 : 0);         // x - 32 wb in [16, 48)
    r0 = load_row64<Mem>(t, y - 1, wb);
    r1 = load_row64<Mem>(t, y, wb);
    r2 = load_row64<Mem>(t, y + 1, wb);
  }
  // bit s of the result = neighbour in direction s is foreground
  __device__ __forceinline__ uint32_t neighbours(int x) const {
    const int sh = x - wb * 32 - 1;                    // in [0, 61]
    const uint32_t up = (uint32_t)(r0 >> sh) & 7u, mid = (uint32_t)(r1 >> sh) & 7u,
                   dn = (uint32_t)(r2 >> sh) & 7u;
    return (mid >> 2) | ((up >> 2) << 1) | (((up >> 1) & 1u) << 2) | ((up & 1u) << 3) |
           ((mid & 1u) << 4) | ((dn & 1u) << 5) | (((dn >> 1) & 1u) << 6) | ((dn >> 2) << 7);
  }
  template <class Mem = GlobalMem>
  __device__ __forceinline__ void move(const TileView& t, int x, int y, int dy) {
    // (x, y) is the new position, dy the vertical part of the step just taken
    if (dy < 0) { r2 = r1; r1 = r0; r0 = load_row64<Mem>(t, y - 1, wb); }
    else if (dy > 0) { r0 = r1; r1 = r2; r2 = load_row64<Mem>(t, y + 1, wb); }
    const int sx = x - wb * 32;
    if (sx < 1 || sx > 62) load<Mem>(t, x, y);
  }
};

constexpr unsigned kFull = 0xffffffffu;

// State of one border-following run (Suzuki-Abe outer border from the raster-first pixel
// (x0, y0) of a component).  Marks (V / G) and per-row extremes are written with result-less
// atomics (RED): no load latency on the serial chain; the same thread's later loads observe
// them (same-address program order).
struct Trace {
  Window w;
  uint32_t nb;
  int x0, y0, x1, y1, x3, y3, s, prev_s;
  int fvx, fvy, lvx, lvy;                  // first / last emitted CHAIN_APPROX_SIMPLE vertex
  long long area2;                         // signed twice-area (shoelace)
  double perim;
  int npts, ymax;
  bool active;
};

template <bool kMark, bool kExt, class Mem = GlobalMem>
__device__ __forceinline__ void trace_begin(const TileView& t, Trace& c, int x0, int y0,
                                            uint32_t* ext_l, uint32_t* ext_r) {
  c.area2 = 0; c.perim = 0.0; c.npts = 0; c.ymax = y0;
  c.x0 = x0; c.y0 = y0; c.x3 = x0; c.y3 = y0;
  c.fvx = c.fvy = c.lvx = c.lvy = 0;
  UWCV_BOUND(y0, t.th); UWCV_BOUND(x0, 32 * t.tw);
  c.w.template load<Mem>(t, x0, y0);
  c.nb = c.w.neighbours(x0);
  if (kExt) { Mem::st(ext_l + y0, (uint32_t)x0); Mem::st(ext_r + y0, (uint32_t)x0); }
  // first search: clockwise from west (3, 2, 1, 0, 7, 6, 5)
  int s = -1;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const int d = (3 - i) & 7;
    if (s < 0 && ((c.nb >> d) & 1u)) s = d;
  }
  if (s < 0) {                                // isolated pixel
    const int o = y0 * t.tw + (x0 >> 5);
    const uint32_t b = 1u << (x0 & 31);
    if (kMark) { Mem::or_(t.V + o, b); Mem::or_(t.G + o, b); }
    c.npts = 1;
    c.active = false;
    c.s = 0; c.prev_s = 0; c.x1 = x0; c.y1 = y0;
    return;
  }
  c.x1 = x0 + dir_dx(s); c.y1 = y0 + dir_dy(s);
  c.s = s;
  c.prev_s = s ^ 4;
  c.active = true;
}

template <bool kMark, bool kExt, class Mem = GlobalMem>
__device__ __forceinline__ void trace_step(const TileView& t, Trace& c, uint32_t* ext_l,
                                           uint32_t* ext_r) {
  const int s_end = c.s;
  // first foreground neighbour counter-clockwise after s_end
  const uint32_t rot = ((c.nb | (c.nb << 8)) >> ((s_end + 1) & 7)) & 0xFFu;
  const int s = (s_end + __ffs(rot)) & 7;     // s_end + 1 + (ffs - 1)
  const int x3 = c.x3, y3 = c.y3;
  UWCV_BOUND(y3, t.th); UWCV_BOUND(x3, 32 * t.tw);
  if (kMark) {
    const int o = y3 * t.tw + (x3 >> 5);
    const uint32_t b = 1u << (x3 & 31);
    Mem::or_(t.V + o, b);
    if ((unsigned)(s - 1) < (unsigned)s_end) Mem::or_(t.G + o, b);
  }
  if (s != c.prev_s) {                        // CHAIN_APPROX_SIMPLE vertex
    if (c.npts == 0) { c.fvx = x3; c.fvy = y3; }
    else {
      const float dx = (float)(x3 - c.lvx), dy = (float)(y3 - c.lvy);
      c.perim += (double)sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    }
    c.lvx = x3; c.lvy = y3;
    ++c.npts;
    c.prev_s = s;
  }
  const int dy = dir_dy(s);
  const int x4 = x3 + dir_dx(s), y4 = y3 + dy;
  c.area2 += (long long)x3 * y4 - (long long)y3 * x4;
  if (x4 == c.x0 && y4 == c.y0 && x3 == c.x1 && y3 == c.y1) {
    if (c.npts >= 2) {                        // closing segment last vertex -> first vertex
      const float dx = (float)(c.fvx - c.lvx), dyy = (float)(c.fvy - c.lvy);
      c.perim += (double)sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dyy, dyy)));
    }
    c.active = false;
    return;
  }
  c.x3 = x4; c.y3 = y4;
  UWCV_BOUND(y4, t.th); UWCV_BOUND(x4, 32 * t.tw);     // the walk never leaves the tile (zero frame)
  if (y4 > c.ymax) {                          // rows are first reached in increasing order
    c.ymax = y4;
    if (kExt) { Mem::st(ext_l + y4, (uint32_t)x4); Mem::st(ext_r + y4, (uint32_t)x4); }
  } else if (kExt) {
    Mem::min_(ext_l + y4, (uint32_t)x4);
    Mem::max_(ext_r + y4, (uint32_t)x4);
  }
  c.w.template move<Mem>(t, x4, y4, dy);
  c.nb = c.w.neighbours(x4);
  c.s = (s + 4) & 7;
}

// sign of the last marked pixel in words [0, wi) of row y: 0 none, +1 positive, -1 negative
template <class Mem = GlobalMem>
static __device__ int last_mark_before(const TileView& t, int wi, int y) {
  UWCV_BOUND(y, t.th); UWCV_BOUND(wi, t.tw + 1);
  const int row = y * t.tw;
  for (--wi; wi >= 0; --wi) {
    const uint32_t v = Mem::ld(t.V + row + wi);
    if (v) {
      const int b = 31 - __clz(v);
      return ((Mem::ld(t.G + row + wi) >> b) & 1u) ? -1 : +1;
    }
  }
  return 0;
}

// Raster scan + border following of ONE instance as a resumable per-lane state machine: every
// call of step() advances the lane by one 64-pixel scan step (two tile words, the next pair
// prefetched) or by one border step; the largest external contour (by |area|, the raster-first
// one on ties) and its per-row extremes are kept.  Used lane-per-instance by the stand-alone
// trace kernel (GlobalMem) and by the tracer warp of the paste kernel (SharedMem).
template <class Mem>
struct LaneTracer {
  enum { kScan = 0, kTrace = 1, kDone = 2 };
  TileView t;
  uint32_t* cur;                           // extremes of the contour being traced: l = cur, r = cur + estride
  uint32_t* best;                          // ... of the best contour so far
  int estride;
  int ncont, best_y, best_npts, best_ymax;
  long long best_a2;
  double best_perim;
  int y, yhi, wi, cy_, cwi, sy, state;
  uint64_t carry, start_mask, cand, vpair, gpair, m_next;
  bool fresh;                              // the next scan step opens a new pair
  Trace tr;

  __device__ __forceinline__ uint64_t load_pair(const uint32_t* plane, int yy, int w0) const {
    UWCV_BOUND(yy, t.th); UWCV_BOUND(w0, t.tw);
    const uint32_t* row = plane + yy * t.tw;
    const uint32_t lo = Mem::ld(row + w0);
    const uint32_t hi = (w0 + 1 < t.tw) ? Mem::ld(row + w0 + 1) : 0u;
    return (uint64_t)lo | ((uint64_t)hi << 32);
  }
  __device__ __forceinline__ uint64_t load_pair_m(int yy, int w0) const {
    UWCV_BOUND(yy, t.th); UWCV_BOUND(w0, t.tw);
    const uint32_t* row = t.M + yy * t.tw;
    const uint32_t lo = Mem::ldm(row + w0);
    const uint32_t hi = (w0 + 1 < t.tw) ? Mem::ldm(row + w0 + 1) : 0u;
    return (uint64_t)lo | ((uint64_t)hi << 32);
  }
  static __device__ __forceinline__ int sign_of_top(uint64_t v, uint64_t g) {      // v != 0
    const int top = 63 - __clzll((long long)v);
    return ((g >> top) & 1ull) ? -1 : +1;
  }

  __device__ __forceinline__ void idle() {
    t.M = nullptr; t.V = t.G = nullptr; t.tw = t.th = 0;
    cur = best = nullptr; estride = 0;
    ncont = 0; best_a2 = -1; best_y = 0; best_npts = 0; best_ymax = -1; best_perim = 0.0;
    y = 0; yhi = -1; wi = 0; cy_ = cwi = sy = 0; state = kDone;
    carry = start_mask = cand = vpair = gpair = m_next = 0ull;
    fresh = true;
    tr.active = false;
  }
  // rows [y_first, y_last] (tile coordinates) are the only ones that can hold a start pixel
  __device__ __forceinline__ void begin(const TileView& tv, uint32_t* ext_a, uint32_t* ext_b,
                                        int ext_stride, int y_first, int y_last) {
    idle();
    t = tv; cur = ext_a; best = ext_b; estride = ext_stride;
    y = y_first; yhi = y_last;
    state = kScan;
    m_next = (y <= yhi) ? load_pair_m(y, 0) : 0ull;
  }
  __device__ __forceinline__ bool done() const { return state == kDone; }

  __device__ __forceinline__ void step() {
    bool finished = false;                               // a contour was completed this step
    if (state == kScan) {
      if (fresh) {
        if (y > yhi) {
          state = kDone;
        } else {
          const uint64_t m = m_next;
          start_mask = m & ~((m << 1) | carry);          // foreground with background on the left
          carry = m >> 63;
          cy_ = y; cwi = wi;
          wi += 2;
          if (wi >= t.tw) { wi = 0; ++y; carry = 0; }
          if (y <= yhi) m_next = load_pair_m(y, wi);     // prefetch the next pair
          // marks are only consulted where the pair holds start candidates (V), and their
          // signs only where a candidate is still unvisited (G)
          vpair = start_mask ? load_pair(t.V, cy_, cwi) : 0ull;
          cand = start_mask & ~vpair;
          gpair = cand ? load_pair(t.G, cy_, cwi) : 0ull;
          fresh = cand == 0;
        }
      }
      if (state == kScan && !fresh) {
        // resolve the candidates of this pair in registers: a candidate starts an external
        // border unless the last marked pixel to its left carries a positive mark
        bool start = false;
        int b = 0;
        while (cand) {
          b = __ffsll((long long)cand) - 1;
          cand &= cand - 1;
          const uint64_t below = vpair & ((1ull << b) - 1ull);
          // last marked pixel to the left: in this pair, else search the earlier words
          const int sgn = below ? sign_of_top(below, gpair) : last_mark_before<Mem>(t, cwi, cy_);
          if (sgn <= 0) { start = true; break; }
        }
        if (start) {
          sy = cy_;
          trace_begin<true, true, Mem>(t, tr, cwi * 32 + b, cy_, cur, cur + estride);
          if (tr.active) state = kTrace; else finished = true;
        } else {
          fresh = true;
        }
      }
    } else if (state == kTrace) {
      trace_step<true, true, Mem>(t, tr, cur, cur + estride);
      if (!tr.active) { finished = true; state = kScan; }
    }
    if (finished) {
      ++ncont;
      const long long a2 = tr.area2 < 0 ? -tr.area2 : tr.area2;
      if (a2 > best_a2) {
        best_a2 = a2; best_y = sy; best_npts = tr.npts; best_perim = tr.perim;
        best_ymax = tr.ymax;
        uint32_t* tmp = cur; cur = best; best = tmp;
      }
      // the trace marked pixels of this row (never to the left of its start): reload
      vpair = load_pair(t.V, cy_, cwi); gpair = load_pair(t.G, cy_, cwi);
      cand &= ~vpair;
    }
  }
};

__device__ __forceinline__ uint32_t pk(int x, int y) { return (uint32_t)x | ((uint32_t)y << 16); }
__device__ __forceinline__ int pkx(uint32_t p) { return (int)(p & 0xffffu); }
__device__ __forceinline__ int pky(uint32_t p) { return (int)(p >> 16); }
__device__ __forceinline__ long long cross3(uint32_t a, uint32_t b, uint32_t c) {
  return (long long)(pkx(b) - pkx(a)) * (pky(c) - pky(b)) -
         (long long)(pky(b) - pky(a)) * (pkx(c) - pkx(b));
}

struct Hull {
  const uint32_t* R; int nr; int skip_r0;   // right chain, top -> bottom
  const uint32_t* L; int nl; int skip_lb;   // left chain,  top -> bottom (walked backwards)
  int n;
  int ox, oy;                               // tile origin in frame pixels
  bool swap2;                               // n == 2: (x, y)-larger point first
  __device__ __forceinline__ uint32_t raw(int i) const {
    if (n == 2 && swap2) i ^= 1;
    const int nrr = nr - skip_r0;
    if (i < nrr) return R[i + skip_r0];
    return L[nl - 1 - skip_lb - (i - nrr)];
  }
  __device__ __forceinline__ float x(int i) const { return (float)(pkx(raw(i)) + ox); }
  __device__ __forceinline__ float y(int i) const { return (float)(pky(raw(i)) + oy); }
};

struct Rect { float cx, cy, w, h, angle; };

// OpenCV rotcalipers.cpp::rotatingCalipers(CALIPERS_MINAREARECT) + minAreaRect epilogue.
// Called by all 32 lanes of a warp (lanes without work pass n == 0): its loops are
// warp-uniform.
static __device__ Rect min_area_rect(const Hull& hl) {
  const double kPi = 3.1415926535897932384626433832795;
  Rect r;
  r.cx = r.cy = 0.f;
  const int n = hl.n;
  float o0x = 0, o0y = 0, o1x = 0, o1y = 0, o2x = 0, o2y = 0;
  float w = 0.f, h = 0.f;
  double ang = 0.0;
  const bool big = n > 2;
  const int nn = big ? n : 0;               // trip count of the calipers loops for this lane
  if (__any_sync(kFull, big)) {
    int left = 0, bottom = 0, right = 0, top = 0;
    float left_x = 0, right_x = 0, top_y = 0, bottom_y = 0;
    if (big) { left_x = right_x = hl.x(0); top_y = bottom_y = hl.y(0); }
    for (int i = 0; __any_sync(kFull, i < nn); ++i) {
      if (i < nn) {
        const float px = hl.x(i), py = hl.y(i);
        if (px < left_x) { left_x = px; left = i; }
        if (px > right_x) { right_x = px; right = i; }
        if (py > top_y) { top_y = py; top = i; }
        if (py < bottom_y) { bottom_y = py; bottom = i; }
      }
    }
    auto vx = [&](int i) { const int j = (i + 1 == n) ? 0 : i + 1; return hl.x(j) - hl.x(i); };
    auto vy = [&](int i) { const int j = (i + 1 == n) ? 0 : i + 1; return hl.y(j) - hl.y(i); };
    float orientation = 0.f;
    {
      double ax = 0, ay = 0;
      if (big) { ax = vx(n - 1); ay = vy(n - 1); }
      for (int i = 0; __any_sync(kFull, i < nn && orientation == 0.f); ++i) {
        if (i < nn && orientation == 0.f) {
          const double bx = vx(i), by = vy(i);
          const double convexity = ax * by - ay * bx;
          if (convexity != 0) orientation = convexity > 0 ? 1.f : -1.f;
          ax = bx; ay = by;
        }
      }
    }
    float base_a = orientation, base_b = 0.f;
    // caliper sides 0..3 = bottom, right, top, left: index, point and outgoing edge kept in
    // registers; only the side that advances fetches a new hull point (one load per step)
    int q0 = bottom, q1 = right, q2 = top, q3 = left;
    float p0x = 0, p0y = 0, p1x = 0, p1y = 0, p2x = 0, p2y = 0, p3x = 0, p3y = 0;
    float e0x = 0, e0y = 0, e1x = 0, e1y = 0, e2x = 0, e2y = 0, e3x = 0, e3y = 0;
    if (big) {
      p0x = hl.x(q0); p0y = hl.y(q0); p1x = hl.x(q1); p1y = hl.y(q1);
      p2x = hl.x(q2); p2y = hl.y(q2); p3x = hl.x(q3); p3y = hl.y(q3);
      e0x = vx(q0); e0y = vy(q0); e1x = vx(q1); e1y = vy(q1);
      e2x = vx(q2); e2y = vy(q2); e3x = vx(q3); e3y = vy(q3);
    }
    float minarea = 3.402823466e+38f;
    float bl_x = 0, bl_y = 0, bb_x = 0, bb_y = 0;       // "leftist" and "bottom" points of the best
    float b_a = 0, b_b = 0, b_w = 0, b_h = 0;
    for (int k = 0; __any_sync(kFull, k < nn); ++k) {
      if (k >= nn) continue;
      // edge of each caliper side rotated into side 0's frame
      const float rvx[4] = {e0x, e1y, -e2x, -e3y};
      const float rvy[4] = {e0y, -e1x, -e2y, e3x};
      int main_el = 0;
      float mx_ = rvx[0], my_ = rvy[0];
#pragma unroll
      for (int i = 1; i < 4; ++i) {
        // firstVecIsRight(rv[i], rv[main]): rotate90CW(rv[i]) . rv[main] < 0
        const float t0 = rvy[i], t1 = -rvx[i];
        if (__fadd_rn(__fmul_rn(t0, mx_), __fmul_rn(t1, my_)) < 0.f) {
          main_el = i; mx_ = rvx[i]; my_ = rvy[i];
        }
      }
      {
        const float lx = main_el == 0 ? e0x : main_el == 1 ? e1x : main_el == 2 ? e2x : e3x;
        const float ly = main_el == 0 ? e0y : main_el == 1 ? e1y : main_el == 2 ? e2y : e3y;
        const double dx = lx, dy = ly;
        const float inv_len = (float)(1.0 / sqrt(dx * dx + dy * dy));
        const float lead_x = __fmul_rn(lx, inv_len);
        const float lead_y = __fmul_rn(ly, inv_len);
        switch (main_el) {
          case 0: base_a = lead_x;  base_b = lead_y;  break;
          case 1: base_a = lead_y;  base_b = -lead_x; break;
          case 2: base_a = -lead_x; base_b = -lead_y; break;
          default: base_a = -lead_y; base_b = lead_x; break;
        }
      }
      // advance the chosen side: its point becomes the old edge's end, fetch the next edge
      {
        int q = main_el == 0 ? q0 : main_el == 1 ? q1 : main_el == 2 ? q2 : q3;
        q = (q + 1 == n) ? 0 : q + 1;
        const int qn = (q + 1 == n) ? 0 : q + 1;
        const float nx = hl.x(qn), ny = hl.y(qn);
        if (main_el == 0) { p0x += e0x; p0y += e0y; e0x = nx - p0x; e0y = ny - p0y; q0 = q; }
        else if (main_el == 1) { p1x += e1x; p1y += e1y; e1x = nx - p1x; e1y = ny - p1y; q1 = q; }
        else if (main_el == 2) { p2x += e2x; p2y += e2y; e2x = nx - p2x; e2y = ny - p2y; q2 = q; }
        else { p3x += e3x; p3y += e3y; e3x = nx - p3x; e3y = ny - p3y; q3 = q; }
      }
      float dx = p1x - p3x;
      float dy = p1y - p3y;
      const float width = __fadd_rn(__fmul_rn(dx, base_a), __fmul_rn(dy, base_b));
      dx = p2x - p0x;
      dy = p2y - p0y;
      const float height = __fadd_rn(__fmul_rn(-dx, base_b), __fmul_rn(dy, base_a));
      const float area = __fmul_rn(width, height);
      if (area <= minarea) {
        minarea = area;
        bl_x = p3x; bl_y = p3y; b_a = base_a; b_w = width; b_b = base_b; b_h = height;
        bb_x = p0x; bb_y = p0y;
      }
    }
    if (big) {
    const float A1 = b_a, B1 = b_b, A2 = -b_b, B2 = b_a;
    const float C1 = __fadd_rn(__fmul_rn(A1, bl_x), __fmul_rn(bl_y, B1));
    const float C2 = __fadd_rn(__fmul_rn(A2, bb_x), __fmul_rn(bb_y, B2));
    const float idet = __fdiv_rn(1.f, __fsub_rn(__fmul_rn(A1, B2), __fmul_rn(A2, B1)));
    o0x = __fmul_rn(__fsub_rn(__fmul_rn(C1, B2), __fmul_rn(C2, B1)), idet);
    o0y = __fmul_rn(__fsub_rn(__fmul_rn(A1, C2), __fmul_rn(A2, C1)), idet);
    o1x = __fmul_rn(A1, b_w); o1y = __fmul_rn(B1, b_w);
    o2x = __fmul_rn(A2, b_h); o2y = __fmul_rn(B2, b_h);
    r.cx = __fadd_rn(o0x, __fmul_rn(__fadd_rn(o1x, o2x), 0.5f));
    r.cy = __fadd_rn(o0y, __fmul_rn(__fadd_rn(o1y, o2y), 0.5f));
    w = (float)sqrt((double)o1x * o1x + (double)o1y * o1y);
    h = (float)sqrt((double)o2x * o2x + (double)o2y * o2y);
    if (o1y == 0.f) ang = o1x >= 0.f ? 0.0 : kPi;
    else if (o1x == 0.f) ang = o1y > 0.f ? kPi * 0.5 : -kPi * 0.5;
    else ang = atan2((double)o1y, (double)o1x);
    }
  }
  if (big) {
  } else if (n == 2) {
    r.cx = __fmul_rn(__fadd_rn(hl.x(0), hl.x(1)), 0.5f);
    r.cy = __fmul_rn(__fadd_rn(hl.y(0), hl.y(1)), 0.5f);
    const double dx = hl.x(1) - hl.x(0), dy = hl.y(1) - hl.y(0);
    w = (float)sqrt(dx * dx + dy * dy);
    h = 0.f;
    if (dy == 0.0) ang = dx >= 0.0 ? 0.0 : kPi;
    else if (dx == 0.0) ang = dy > 0.0 ? kPi * 0.5 : -kPi * 0.5;
    else ang = atan2(dy, dx);
  } else {
    r.cx = n == 1 ? hl.x(0) : 0.f;
    r.cy = n == 1 ? hl.y(0) : 0.f;
  }
  ang = ang * 180 / kPi;
  while (ang >= 0.0) { ang -= 90.0; const float t = w; w = h; h = t; }
  while (ang < -90.0) { ang += 90.0; const float t = w; w = h; h = t; }
  r.w = w; r.h = h; r.angle = (float)ang;
  return r;
}

// cv2.boxPoints -> np.array(dtype="int") -> imutils order_points -> midpoints -> (dA, dB)
static __device__ void feret_extents(const Rect& r, float& dA, float& dB) {
  const double kPi = 3.1415926535897932384626433832795;
  const double ang = (double)r.angle * kPi / 180.;
  const float b = __fmul_rn((float)cos(ang), 0.5f);
  const float a = __fmul_rn((float)sin(ang), 0.5f);
  float px[4], py[4];
  px[0] = __fsub_rn(__fsub_rn(r.cx, __fmul_rn(a, r.h)), __fmul_rn(b, r.w));
  py[0] = __fsub_rn(__fadd_rn(r.cy, __fmul_rn(b, r.h)), __fmul_rn(a, r.w));
  px[1] = __fsub_rn(__fadd_rn(r.cx, __fmul_rn(a, r.h)), __fmul_rn(b, r.w));
  py[1] = __fsub_rn(__fsub_rn(r.cy, __fmul_rn(b, r.h)), __fmul_rn(a, r.w));
  px[2] = __fsub_rn(__fmul_rn(2.f, r.cx), px[0]);
  py[2] = __fsub_rn(__fmul_rn(2.f, r.cy), py[0]);
  px[3] = __fsub_rn(__fmul_rn(2.f, r.cx), px[1]);
  py[3] = __fsub_rn(__fmul_rn(2.f, r.cy), py[1]);
  int X[4], Y[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { X[i] = (int)px[i]; Y[i] = (int)py[i]; }   // truncation toward zero
  // order_points: stable sort by x (NumPy argsort of 4 elements is an insertion sort)
  int idx[4] = {0, 1, 2, 3};
#pragma unroll
  for (int i = 1; i < 4; ++i) {
#pragma unroll
    for (int j = i; j > 0; --j) {
      if (X[idx[j]] < X[idx[j - 1]]) { const int t = idx[j]; idx[j] = idx[j - 1]; idx[j - 1] = t; }
    }
  }
  int l0 = idx[0], l1 = idx[1], r0 = idx[2], r1 = idx[3];
  if (Y[l1] < Y[l0]) { const int t = l0; l0 = l1; l1 = t; }      // (tl, bl) by y, stable
  const int tl = l0, bl = l1;
  const long long d0 = (long long)(X[r0] - X[tl]) * (X[r0] - X[tl]) +
                       (long long)(Y[r0] - Y[tl]) * (Y[r0] - Y[tl]);
  const long long d1 = (long long)(X[r1] - X[tl]) * (X[r1] - X[tl]) +
                       (long long)(Y[r1] - Y[tl]) * (Y[r1] - Y[tl]);
  // argsort(D)[::-1]: ascending stable then reversed -> (br, tr)
  int br, tr;
  if (d1 < d0) { br = r0; tr = r1; } else { br = r1; tr = r0; }
  // midpoints are float32 half-integers (order_points returns float32, exact); scipy's
  // dist.euclidean keeps float32 and reduces with snrm2 = sqrtf(fl(dx*dx) + fl(dy*dy))
  // (pinned by probe against scipy 1.18 / OpenBLAS in this image).
  const float ax = 0.5f * (float)((X[tl] + X[tr]) - (X[bl] + X[br]));
  const float ay = 0.5f * (float)((Y[tl] + Y[tr]) - (Y[bl] + Y[br]));
  const float bx = 0.5f * (float)((X[tl] + X[bl]) - (X[tr] + X[br]));
  const float by = 0.5f * (float)((Y[tl] + Y[bl]) - (Y[tr] + Y[br]));
  dA = sqrtf(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)));
  dB = sqrtf(__fadd_rn(__fmul_rn(bx, bx), __fmul_rn(by, by)));
}


// Hull -> min-area rectangle -> descriptor block for one traced contour, from its per-row
// extremes ext_l / ext_r (tile rows y0 .. ymax).  Warp-uniform: all 32 lanes call it, lanes
// without a contour pass have == false.  Writes the 16 doubles
// [contour_area, perimeter, rect cx, cy, w, h, angle, Feret, Aspect_Ratio, Roundness,
//  Circularity, Sphericity, Length, Width, CircularED, Chords] (nn_inference.py:434-449).
static __device__ void describe_contour(bool have, uint32_t* ext_l, uint32_t* ext_r, int y0, int ymax,
                                 int ox, int oy, long long a2, double perim, double ppm,
                                 double* out) {
  const double kPi = 3.141592653589793;
  const int ylast = have ? ymax : y0 - 1;
  // right chain (top -> bottom, clockwise on screen): pop while the turn is not strictly
  // convex; one pop or one push per iteration, the two topmost entries live in registers
  int nr = 0, nl = 0;
  {
    uint32_t a = 0, b = 0, p = 0;
    int yy = y0;
    bool need = true;
    while (__any_sync(kFull, yy <= ylast)) {
      if (yy <= ylast) {
        UWCV_BOUND(yy - y0, ymax - y0 + 1); UWCV_BOUND(nr, ymax - y0 + 1);
        if (need) { p = pk((int)ext_r[yy], yy); need = false; }
        if (nr >= 2 && cross3(a, b, p) <= 0) {
          --nr; b = a;
          if (nr >= 2) a = ext_r[y0 + nr - 2];
        } else {
          ext_r[y0 + nr] = p; ++nr;
          a = b; b = p;
          ++yy; need = true;
        }
      }
    }
  }
  // left chain, also top -> bottom (counter-clockwise on screen): mirrored turn test
  {
    uint32_t a = 0, b = 0, p = 0;
    int yy = y0;
    bool need = true;
    while (__any_sync(kFull, yy <= ylast)) {
      if (yy <= ylast) {
        UWCV_BOUND(yy - y0, ymax - y0 + 1); UWCV_BOUND(nl, ymax - y0 + 1);
        if (need) { p = pk((int)ext_l[yy], yy); need = false; }
        if (nl >= 2 && cross3(a, b, p) >= 0) {
          --nl; b = a;
          if (nl >= 2) a = ext_l[y0 + nl - 2];
        } else {
          ext_l[y0 + nl] = p; ++nl;
          a = b; b = p;
          ++yy; need = true;
        }
      }
    }
  }
  Hull hl;
  hl.R = ext_r + y0; hl.nr = nr;
  hl.L = ext_l + y0; hl.nl = nl;
  hl.skip_r0 = 0; hl.skip_lb = 0; hl.n = 0;
  hl.ox = ox; hl.oy = oy;
  hl.swap2 = false;
  if (have) {
    hl.skip_r0 = (hl.R[0] == hl.L[0]) ? 1 : 0;                      // single-pixel top row
    hl.skip_lb = (hl.R[nr - 1] == hl.L[nl - 1]) ? 1 : 0;            // single-pixel bottom row
    hl.n = (nr - hl.skip_r0) + (nl - hl.skip_lb);
    if (hl.n <= 0) {            // a single pixel: both chains hold the same point
      hl.skip_r0 = 0; hl.skip_lb = 1; hl.n = 1;
    }
    if (hl.n == 2) {
      const uint32_t p0 = hl.raw(0), p1 = hl.raw(1);
      const bool p0_larger = pkx(p0) > pkx(p1) || (pkx(p0) == pkx(p1) && pky(p0) > pky(p1));
      hl.swap2 = !p0_larger;
    }
  }
  const Rect rect = min_area_rect(hl);                   // warp-uniform loops inside
  if (!have) return;
  float dA, dB;
  feret_extents(rect, dA, dB);
  const double area = (double)a2 * 0.5;
  // dA, dB are numpy float32 scalars in the reference; float32 / Python float stays
  // float32 (NumPy >= 2 promotion), so Feret / Aspect_Ratio / Roundness / Length / Width
  // are float32 arithmetic, while area and perimeter (Python floats) stay float64.
  const float ppm32 = (float)ppm;
  const float dimA = __fdiv_rn(dA, ppm32), dimB = __fdiv_rn(dB, ppm32);
  const double dimArea = area / ppm, dimPerimeter = perim / ppm;
  const float mx = dimA > dimB ? dimA : dimB, mn = dimA < dimB ? dimA : dimB;
  const float aspect = (dimA != 0.f && dimB != 0.f) ? __fdiv_rn(mx, mn) : 0.f;
  out[0] = area;
  out[1] = perim;
  out[2] = rect.cx; out[3] = rect.cy; out[4] = rect.w; out[5] = rect.h; out[6] = rect.angle;
  out[7] = mx;
  out[8] = aspect;
  out[9] = aspect != 0.f ? __fdiv_rn(1.f, aspect) : 0.f;
  out[10] = 4 * kPi * (dimArea / (dimPerimeter * dimPerimeter));
  out[11] = (2 * sqrt(kPi * dimArea)) / dimPerimeter;
  out[12] = mn;
  out[13] = mx;
  out[14] = sqrt(4 * area / kPi);
  out[15] = perim;
}

}  // namespace uwcv
