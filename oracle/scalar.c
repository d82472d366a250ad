/*
 * oracle/scalar.c -- plain-C scalar restatement of the paste arithmetic and the raw
 * image moments (TEST INFRASTRUCTURE; never linked into the product).
 *
 * Follows the published algorithms the reference reaches through predictor(im)
 * (nn_inference.py:372): detectron2/layers/mask_ops.py::_do_paste_mask +
 * ATen GridSamplerKernel.cpp (bilinear, zeros padding, align_corners=False) as pinned
 * by the probe recorded in SURVEY.md section 8(c), and OpenCV moments.cpp raw moments.
 * Validated against torch.nn.functional.grid_sample / cv2.moments by
 * tests/test_oracle_c.py.  Built by oracle/Makefile into oracle/_build/liboracle.so
 * with -ffp-contract=off so that only the explicit fmaf() calls fuse.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MS 28

static void axis(int p, float lo, float hi, float *fl, float *w_hi, float *w_lo) {
  volatile float a = (float)p + 0.5f;
  volatile float b = a - lo;
  volatile float c = hi - lo;
  volatile float d = b / c;
  volatile float e = d * 2.0f;
  volatile float g = e - 1.0f;
  volatile float t = g + 1.0f;
  float i = fmaf(t, (float)MS * 0.5f, -0.5f);
  *fl = floorf(i);
  *w_hi = i - *fl;
  *w_lo = 1.0f - *w_hi;
}

static float tap(const float *m, float fy, float fx) {
  if (!(fy >= 0.f && fy <= (float)(MS - 1) && fx >= 0.f && fx <= (float)(MS - 1))) return 0.f;
  return m[(int)fy * MS + (int)fx];
}

/* Paste one instance over the pixel window [x_lo, x_hi) x [y_lo, y_hi); out is row-major
 * uint8 (0/1) of that window.  Returns the number of set pixels. */
int64_t oracle_paste_window(const float *mask, const float *box, int x_lo, int x_hi, int y_lo,
                            int y_hi, float thr, uint8_t *out) {
  const float x0 = box[0], y0 = box[1], x1 = box[2], y1 = box[3];
  const int w = x_hi - x_lo;
  int64_t count = 0;
  for (int py = y_lo; py < y_hi; ++py) {
    float fy, n, s;
    axis(py, y0, y1, &fy, &n, &s);
    for (int px = x_lo; px < x_hi; ++px) {
      float fx, ww, e;
      axis(px, x0, x1, &fx, &ww, &e);
      volatile float nw = s * e, ne = s * ww, sw = n * e, se = n * ww;
      volatile float acc = tap(mask, fy, fx) * nw;
      float o = fmaf(tap(mask, fy, fx + 1.f), ne, acc);
      o = fmaf(tap(mask, fy + 1.f, fx), sw, o);
      o = fmaf(tap(mask, fy + 1.f, fx + 1.f), se, o);
      const uint8_t bit = (o >= thr) ? 1 : 0;
      out[(int64_t)(py - y_lo) * w + (px - x_lo)] = bit;
      count += bit;
    }
  }
  return count;
}

/* Raw moments m00 m10 m01 m20 m11 m02 m30 m21 m12 m03 of a 0/1 window placed at
 * (x_off, y_off) of the frame, exact in int64. */
void oracle_raw_moments(const uint8_t *win, int w, int h, int x_off, int y_off, int64_t *m) {
  memset(m, 0, 10 * sizeof(int64_t));
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      if (win[(int64_t)y * w + x]) {
        const int64_t X = x + x_off, Y = y + y_off;
        m[0] += 1; m[1] += X; m[2] += Y; m[3] += X * X; m[4] += X * Y; m[5] += Y * Y;
        m[6] += X * X * X; m[7] += X * X * Y; m[8] += X * Y * Y; m[9] += Y * Y * Y;
      }
}
