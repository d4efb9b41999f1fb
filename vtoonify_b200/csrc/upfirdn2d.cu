// upfirdn2d.cu — planar (NCHW) upsample / pad-crop / 2-D FIR / downsample in one pass.
//
// Semantics follow the reference CPU spec model/stylegan/op_cpu/upfirdn2d.py:19-60 and the CUDA
// op model/stylegan/op/upfirdn2d_kernel.cu:49-105 (generic kernel): zero-stuff by `up`, pad (negative
// pad crops), true convolution (kernel flipped), keep every `down`-th sample.  Written from the maths:
//   out[p,oy,ox] = sum_{ky,kx} U[p, oy*down_y + ky - pad_y0, ox*down_x + kx - pad_x0] * k[kh-1-ky][kw-1-kx]
//   U[p,u,v] = in[p,u/up_y,v/up_x] if up_y|u, up_x|v and inside, else 0.
// All offsets are 64-bit (the reference kernels overflow 32-bit ints at 576x1024 B>=8).
//
// Two kernels:
//   * upfirdn2d_generic_kernel : any (up, down, pad, kh, kw); one output per thread.
//   * upfirdn2d_sep4_kernel    : the hot-path instances (4-tap separable [1,3,3,1]-style filters given as
//     their 2-D outer product, up in {1,2}, down in {1,2}); each lane owns one output column, horizontal taps
//     come from neighbouring lanes through warp shuffles, vertical taps are a register sliding window.
#include "common.cuh"

namespace {

constexpr int kMaxSmemTaps = 1024;

__global__ void __launch_bounds__(256)
upfirdn2d_generic_kernel(const float* __restrict__ in, const float* __restrict__ kernel, float* __restrict__ out,
                         int64_t planes, int in_h, int in_w, int out_h, int out_w, int kh, int kw,
                         int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_y0, int taps_in_smem) {
  __shared__ float sk[kMaxSmemTaps];
  if (taps_in_smem) {
    for (int i = threadIdx.x; i < kh * kw; i += blockDim.x) sk[i] = kernel[i];
    __syncthreads();
  }
  const float* kp = taps_in_smem ? sk : kernel;
  const int64_t total = planes * (int64_t)out_h * out_w;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % out_w);
    const int64_t t1 = idx / out_w;
    const int oy = (int)(t1 % out_h);
    const int64_t p = t1 / out_h;
    const int ty = oy * down_y - pad_y0;  // position of tap ky=0 in the zero-stuffed signal
    const int tx = ox * down_x - pad_x0;
    int ky0, kx0;
    if (ty >= 0) { int r = ty % up_y; ky0 = r ? up_y - r : 0; } else { ky0 = -ty; }
    if (tx >= 0) { int r = tx % up_x; kx0 = r ? up_x - r : 0; } else { kx0 = -tx; }
    const float* ip = in + p * (int64_t)in_h * in_w;
    float acc = 0.f;
    for (int ky = ky0; ky < kh; ky += up_y) {
      const int iy = (ty + ky) / up_y;
      if (iy >= in_h) break;
      const float* row = ip + (int64_t)iy * in_w;
      const float* krow = kp + (kh - 1 - ky) * kw;
      for (int kx = kx0; kx < kw; kx += up_x) {
        const int ix = (tx + kx) / up_x;
        if (ix >= in_w) break;
        acc += __ldg(row + ix) * krow[kw - 1 - kx];
      }
    }
    out[idx] = acc;
  }
}

}  // namespace

extern "C" int vt_upfirdn2d_out_size(int in_h, int in_w, int kh, int kw, int up_x, int up_y, int down_x, int down_y,
                                     int pad_x0, int pad_x1, int pad_y0, int pad_y1, int* out_h, int* out_w) {
  VT_CHECK(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1, "upfirdn2d: up/down must be >= 1");
  VT_CHECK(kh >= 1 && kw >= 1, "upfirdn2d: empty kernel");
  const int64_t nh = (int64_t)in_h * up_y + pad_y0 + pad_y1 - kh + down_y;
  const int64_t nw = (int64_t)in_w * up_x + pad_x0 + pad_x1 - kw + down_x;
  // floor division like Python's // (model/stylegan/op_cpu/upfirdn2d.py:58-59)
  auto fdiv = [](int64_t a, int64_t b) { int64_t q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) --q; return q; };
  *out_h = (int)fdiv(nh, down_y);
  *out_w = (int)fdiv(nw, down_x);
  return 0;
}

extern "C" int vt_upfirdn2d_f32(const float* in, const float* kernel, float* out, int64_t planes, int in_h, int in_w,
                                int kh, int kw, int up_x, int up_y, int down_x, int down_y,
                                int pad_x0, int pad_x1, int pad_y0, int pad_y1, void* stream) {
  int out_h = 0, out_w = 0;
  if (vt_upfirdn2d_out_size(in_h, in_w, kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1, &out_h, &out_w))
    return 1;
  VT_CHECK(in && kernel && out, "upfirdn2d: null pointer");
  VT_CHECK(planes >= 0 && in_h >= 1 && in_w >= 1, "upfirdn2d: bad input shape");
  VT_CHECK(out_h >= 1 && out_w >= 1, "upfirdn2d: empty output (%d x %d)", out_h, out_w);
  if (planes == 0) return 0;
  const int64_t total = planes * (int64_t)out_h * out_w;
  const int threads = 256;
  int64_t blocks = vt_cdiv(total, threads);
  const int64_t max_blocks = (int64_t)vt_num_sms() * 32;
  if (blocks > max_blocks) blocks = max_blocks;
  upfirdn2d_generic_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
      in, kernel, out, planes, in_h, in_w, out_h, out_w, kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_y0,
      (kh * kw <= kMaxSmemTaps) ? 1 : 0);
  VT_LAUNCH_CHECK();
  return 0;
}
