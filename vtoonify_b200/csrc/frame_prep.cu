// frame_prep.cu — f3: the frame loop's CPU pre-processing of high-resolution video frames on the device
// (style_transfer.py:124-130, 151-156): `cv2.sepFilter2D(frame, -1, k, k)` with k = [1, 3, 3, 1] / 8 applied 0, 1 or 2 times,
// then `cv2.resize(frame, (w, h))[top:bottom, left:right]` (bilinear).  uint8 in, uint8 out, BIT-EXACT with OpenCV 4.x:
//   * sepFilter2D on 8-bit data: anchor = ksize / 2 = 2 (taps at -2 .. +1), BORDER_REFLECT_101, the exact value
//     sum_{a,b} w_a w_b p / 64 rounded half-to-even (cvRound), saturated to [0, 255];
//   * resize INTER_LINEAR on 8-bit data: 11-bit fixed-point coefficients (tables built on the host exactly as
//     cv::resize does: float source coordinate, cvFloor, cvRound(coef * 2048), x clamped with unit weight, y rows clamped),
//     horizontal pass in int32, vertical pass ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.
// Both are HBM-bound byte kernels: one thread per output pixel (3 bytes), coalesced row-major, reads served by L1/L2 (every
// source byte is touched by <= 16 neighbouring threads of the same block).
#include "common.cuh"

namespace {

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

// grid (ceil(W*H / 256), B)
__global__ void __launch_bounds__(256)
frame_blur4_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int H, int W) {
  const int64_t HW = (int64_t)H * W;
  const uint8_t* ip = in + (int64_t)blockIdx.y * HW * 3;
  uint8_t* op = out + (int64_t)blockIdx.y * HW * 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / W), x = (int)(i - (int64_t)y * W);
    int xs[4], ys[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { xs[k] = reflect101(x + k - 2, W); ys[k] = reflect101(y + k - 2, H); }
    const int wt[4] = {1, 3, 3, 1};
    int acc[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const uint8_t* row = ip + (int64_t)ys[a] * W * 3;
      int r[3] = {0, 0, 0};
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const uint8_t* px = row + xs[b] * 3;
        r[0] += wt[b] * px[0]; r[1] += wt[b] * px[1]; r[2] += wt[b] * px[2];
      }
      acc[0] += wt[a] * r[0]; acc[1] += wt[a] * r[1]; acc[2] += wt[a] * r[2];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int s = acc[c];                               // 64 x the exact filtered value
      int v = (s + 31 + ((s >> 6) & 1)) >> 6;             // round half to even
      op[i * 3 + c] = (uint8_t)(v > 255 ? 255 : v);
    }
  }
}

// tables: xofs/xa0/xa1 [dw], yofs/yb0/yb1 [dh] (int32); output = resized[top : top+Ho, left : left+Wo]
__global__ void __launch_bounds__(256)
frame_resize_crop_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int Hs, int Ws, int Ho, int Wo, int top, int left,
                            const int* __restrict__ xofs, const int* __restrict__ xa0, const int* __restrict__ xa1,
                            const int* __restrict__ yofs, const int* __restrict__ yb0, const int* __restrict__ yb1) {
  const int64_t n = (int64_t)Ho * Wo;
  const uint8_t* ip = in + (int64_t)blockIdx.y * Hs * Ws * 3;
  uint8_t* op = out + (int64_t)blockIdx.y * n * 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int oy = (int)(i / Wo), ox = (int)(i - (int64_t)oy * Wo);
    const int dy = oy + top, dx = ox + left;
    const int sx0 = __ldg(xofs + dx), a0 = __ldg(xa0 + dx), a1 = __ldg(xa1 + dx);
    const int sx1 = sx0 + 1 < Ws ? sx0 + 1 : Ws - 1;
    const int sy = __ldg(yofs + dy), b0 = __ldg(yb0 + dy), b1 = __ldg(yb1 + dy);
    const int y0 = sy < 0 ? 0 : (sy > Hs - 1 ? Hs - 1 : sy);
    const int y1 = sy + 1 < 0 ? 0 : (sy + 1 > Hs - 1 ? Hs - 1 : sy + 1);
    const uint8_t* r0 = ip + (int64_t)y0 * Ws * 3;
    const uint8_t* r1 = ip + (int64_t)y1 * Ws * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int s0 = r0[sx0 * 3 + c] * a0 + r0[sx1 * 3 + c] * a1;
      const int s1 = r1[sx0 * 3 + c] * a0 + r1[sx1 * 3 + c] * a1;
      int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
      op[i * 3 + c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
  }
}

unsigned grid_px(int64_t n) {
  int64_t blocks = vt_cdiv(n, 256);
  const int64_t cap = (int64_t)vt_num_sms() * 32;
  return (unsigned)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

}  // namespace

extern "C" int vt_frame_blur4_u8(const uint8_t* in, uint8_t* out, int B, int H, int W, void* stream) {
  VT_CHECK(in && out && in != out && B >= 1 && B <= 65535 && H >= 1 && W >= 1, "frame_blur4_u8: bad args (in-place is not supported)");
  frame_blur4_u8_kernel<<<dim3(grid_px((int64_t)H * W), (unsigned)B), 256, 0, (cudaStream_t)stream>>>(in, out, H, W);
  VT_LAUNCH_CHECK();
  return 0;
}

extern "C" int vt_frame_resize_crop_u8(const uint8_t* in, uint8_t* out, int B, int Hs, int Ws, int dh, int dw, int top, int left,
                                       int Ho, int Wo, const int* xtab, const int* ytab, void* stream) {
  VT_CHECK(in && out && xtab && ytab && B >= 1 && B <= 65535 && Hs >= 1 && Ws >= 1 && dh >= 1 && dw >= 1, "frame_resize_crop_u8: bad args");
  VT_CHECK(top >= 0 && left >= 0 && Ho >= 1 && Wo >= 1 && top + Ho <= dh && left + Wo <= dw, "frame_resize_crop_u8: crop window outside the resized frame");
  frame_resize_crop_u8_kernel<<<dim3(grid_px((int64_t)Ho * Wo), (unsigned)B), 256, 0, (cudaStream_t)stream>>>(
      in, out, Hs, Ws, Ho, Wo, top, left, xtab, xtab + dw, xtab + 2 * dw, ytab, ytab + dh, ytab + 2 * dh);
  VT_LAUNCH_CHECK();
  return 0;
}
