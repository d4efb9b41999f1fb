/* oracle/vt_oracle.c — plain-C restatement of the integer/index-heavy parts of the hot path. TEST INFRASTRUCTURE ONLY
 * (same rule as vt_oracle.py: only tests/, smoke() and bench.py's cpu_baseline may call it; never the product path).
 *
 *   vo_upfirdn2d      : model/stylegan/op_cpu/upfirdn2d.py:19-60 / op/upfirdn2d_kernel.cu:49-105 — zero-stuff, pad/crop,
 *                       true convolution (flipped kernel), decimate; floor semantics for negative coordinates.
 *   vo_fused_bias_act : model/stylegan/op_cpu/fused_act.py:23-34
 *   vo_conv2d_nchw    : naive direct cross-correlation, the arithmetic of F.conv2d as called at
 *                       model/stylegan/op/conv2d_gradfix.py:34-42 (stride, zero padding, dilation), for small cases.
 *   vo_tensor2cv2     : util.py:190-192
 * Build: gcc -O2 -shared -fPIC -o oracle/_build/libvt_oracle.so oracle/vt_oracle.c   (see oracle/Makefile)
 */
#include <stdint.h>
#include <stddef.h>

static long floordiv(long a, long b) { long q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) --q; return q; }

int vo_upfirdn2d(const float* in, const float* k, float* out, long planes, int in_h, int in_w, int kh, int kw,
                 int up_x, int up_y, int down_x, int down_y, int px0, int px1, int py0, int py1) {
  long out_h = floordiv((long)in_h * up_y + py0 + py1 - kh + down_y, down_y);
  long out_w = floordiv((long)in_w * up_x + px0 + px1 - kw + down_x, down_x);
  if (out_h < 1 || out_w < 1) return 1;
  for (long p = 0; p < planes; ++p)
    for (long oy = 0; oy < out_h; ++oy)
      for (long ox = 0; ox < out_w; ++ox) {
        float acc = 0.f;
        for (int ky = 0; ky < kh; ++ky) {
          long u = oy * down_y + ky - py0;              /* row in the zero-stuffed signal */
          if (u < 0 || u % up_y != 0) continue;
          long iy = u / up_y;
          if (iy >= in_h) continue;
          for (int kx = 0; kx < kw; ++kx) {
            long v = ox * down_x + kx - px0;
            if (v < 0 || v % up_x != 0) continue;
            long ix = v / up_x;
            if (ix >= in_w) continue;
            acc += in[(p * in_h + iy) * in_w + ix] * k[(kh - 1 - ky) * kw + (kw - 1 - kx)];
          }
        }
        out[(p * out_h + oy) * out_w + ox] = acc;
      }
  return 0;
}

void vo_fused_bias_act(const float* in, const float* bias, float* out, long n, long step_b, int size_b, float slope,
                       float scale) {
  for (long i = 0; i < n; ++i) {
    float v = in[i] + (bias ? bias[(i / step_b) % size_b] : 0.f);
    out[i] = (v > 0.f ? v : v * slope) * scale;
  }
}

void vo_conv2d_nchw(const float* in, const float* w, const float* bias, float* out, int B, int Cin, int H, int W,
                    int Cout, int k, int stride, int pad, int dil) {
  int Ho = (H + 2 * pad - dil * (k - 1) - 1) / stride + 1, Wo = (W + 2 * pad - dil * (k - 1) - 1) / stride + 1;
  for (int b = 0; b < B; ++b)
    for (int n = 0; n < Cout; ++n)
      for (int oy = 0; oy < Ho; ++oy)
        for (int ox = 0; ox < Wo; ++ox) {
          double acc = bias ? bias[n] : 0.0;
          for (int c = 0; c < Cin; ++c)
            for (int ky = 0; ky < k; ++ky) {
              int iy = oy * stride + ky * dil - pad;
              if (iy < 0 || iy >= H) continue;
              for (int kx = 0; kx < k; ++kx) {
                int ix = ox * stride + kx * dil - pad;
                if (ix < 0 || ix >= W) continue;
                acc += (double)in[((size_t)(b * Cin + c) * H + iy) * W + ix] * w[((size_t)(n * Cin + c) * k + ky) * k + kx];
              }
            }
          out[((size_t)(b * Cout + n) * Ho + oy) * Wo + ox] = (float)acc;
        }
}

void vo_tensor2cv2(const float* in, uint8_t* out, int H, int W) {
  for (int p = 0; p < H * W; ++p)
    for (int c = 0; c < 3; ++c) {
      float x = in[(size_t)c * H * W + p];
      x = x < -1.f ? -1.f : (x > 1.f ? 1.f : x);
      out[(size_t)p * 3 + (2 - c)] = (uint8_t)(int)((x + 1.0f) * 127.5f); /* RGB -> BGR */
    }
}
