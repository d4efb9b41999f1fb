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
//   * upfirdn2d_tiled_kernel   : same semantics, the input window of a 32 x 64 output tile staged once in shared memory
//     (coalesced), division-free tap loops; used whenever the window fits (all hot-path instances).
#include "common.cuh"

int g_upfirdn_tiled = 1;   // 0 forces the generic kernel (tests compare both)

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

// ---- tiled fast path -----------------------------------------------------------------------------------------------
// A block owns a TOH x TOW output tile of one plane: the input window it needs (with zero halo) is staged once in
// shared memory by coalesced row reads, then each thread walks a column of the tile.  Tap index arithmetic is hoisted:
// the first contributing tap and its input column depend only on the thread's output column, (tx + kx) / up_x advances
// by exactly one input pixel per step of up_x taps, so there is no division in the tap loops.  HBM traffic is the
// algorithmic read-once / write-once (the generic kernel re-reads every input ~kh*kw/(up*up) times through L1/L2).
constexpr int T_OW = 64, T_OH = 32, T_ROWS_PER_THREAD = T_OH / 4;   // 256 threads = 64 columns x 4 row groups

__global__ void __launch_bounds__(256)
upfirdn2d_tiled_kernel(const float* __restrict__ in, const float* __restrict__ kernel, float* __restrict__ out,
                       int in_h, int in_w, int out_h, int out_w, int kh, int kw, int up_x, int up_y, int down_x, int down_y,
                       int pad_x0, int pad_y0, int tiles_x, int tile_rows, int tile_cols) {
  extern __shared__ float smem[];
  float* sk = smem;                        // flipped kernel [kh][kw]
  float* S = smem + kh * kw;               // input window [tile_rows][tile_cols]
  for (int i = threadIdx.x; i < kh * kw; i += blockDim.x) sk[i] = kernel[(kh - 1 - i / kw) * kw + (kw - 1 - i % kw)];
  const int64_t plane = blockIdx.y;
  const int oy0 = (blockIdx.x / tiles_x) * T_OH, ox0 = (blockIdx.x % tiles_x) * T_OW;
  // first input row / column any output of this tile can touch (floor division, may be negative -> zero halo)
  const int uy0 = oy0 * down_y - pad_y0, ux0 = ox0 * down_x - pad_x0;
  const int iy_base = (uy0 >= 0) ? uy0 / up_y : -((-uy0 + up_y - 1) / up_y);
  const int ix_base = (ux0 >= 0) ? ux0 / up_x : -((-ux0 + up_x - 1) / up_x);
  const float* ip = in + plane * (int64_t)in_h * in_w;
  for (int i = threadIdx.x; i < tile_rows * tile_cols; i += blockDim.x) {
    const int r = i / tile_cols, c = i - r * tile_cols;
    const int iy = iy_base + r, ix = ix_base + c;
    S[i] = (iy >= 0 && iy < in_h && ix >= 0 && ix < in_w) ? __ldg(ip + (int64_t)iy * in_w + ix) : 0.f;
  }
  __syncthreads();
  const int c = threadIdx.x & (T_OW - 1), rg = threadIdx.x / T_OW;
  const int ox = ox0 + c;
  if (ox >= out_w) return;
  // column taps: kx = kx0, kx0 + up_x, ... ; input column (tx + kx) / up_x = sx0, sx0 + 1, ...  (relative to the window)
  const int tx = ox * down_x - pad_x0;
  int kx0;
  if (tx >= 0) { const int r = tx % up_x; kx0 = r ? up_x - r : 0; } else { const int r = (-tx) % up_x; kx0 = r; }
  // (tx + kx0) is a multiple of up_x; it may be negative (zero halo covers it)
  const int sx0 = ((tx + kx0) >= 0 ? (tx + kx0) / up_x : -((-(tx + kx0)) / up_x)) - ix_base;
  float* op = out + plane * (int64_t)out_h * out_w;
  for (int j = 0; j < T_ROWS_PER_THREAD; ++j) {
    const int oy = oy0 + rg * T_ROWS_PER_THREAD + j;
    if (oy >= out_h) break;
    const int ty = oy * down_y - pad_y0;
    int ky0;
    if (ty >= 0) { const int r = ty % up_y; ky0 = r ? up_y - r : 0; } else { ky0 = (-ty) % up_y; }
    int sy = ((ty + ky0) >= 0 ? (ty + ky0) / up_y : -((-(ty + ky0)) / up_y)) - iy_base;
    float acc = 0.f;
    for (int ky = ky0; ky < kh; ky += up_y, ++sy) {
      const float* srow = S + sy * tile_cols + sx0;
      const float* krow = sk + ky * kw;
      int sx = 0;
      for (int kx = kx0; kx < kw; kx += up_x, ++sx) acc = fmaf(srow[sx], krow[kx], acc);
    }
    op[(int64_t)oy * out_w + ox] = acc;
  }
}

// ---- 4x4-kernel specialisation (the hot-path instances: Blur pad(1,1)/(2,2), Upsample up=2 pad(2,1), Downsample down=2) ----
// 32 x 128 output tile per block, each thread owns a 4 x 4 output micro-tile.  UP/DOWN are compile-time (one of them is 1),
// the input window is staged in shared memory row by row (one warp per row, coalesced, no index division) and pulled
// into registers once per micro-tile, so a thread issues ~20-30 shared loads and 64-256 FMAs for 16 outputs.
constexpr int K4_OW = 128, K4_OH = 32;

template <int UP, int DOWN>
__global__ void __launch_bounds__(256)
upfirdn2d_k4_kernel(const float* __restrict__ in, const float* __restrict__ kernel, float* __restrict__ out, int in_h,
                    int in_w, int out_h, int out_w, int pad_x0, int pad_y0, int tiles_x) {
  constexpr int WR = ((K4_OH - 1) * DOWN + 3) / UP + 2;      // window rows / cols (upper bounds)
  constexpr int WC = (((K4_OW - 1) * DOWN + 3) / UP + 2 + 3) / 4 * 4;
  extern __shared__ __align__(16) float S[];                  // [WR][WC]
  __shared__ float sk[16];                                    // flipped kernel
  if (threadIdx.x < 16) sk[threadIdx.x] = kernel[(3 - threadIdx.x / 4) * 4 + (3 - threadIdx.x % 4)];
  const int64_t plane = blockIdx.y;
  const int oy0 = (blockIdx.x / tiles_x) * K4_OH, ox0 = (blockIdx.x % tiles_x) * K4_OW;
  const int uy0 = oy0 * DOWN - pad_y0, ux0 = ox0 * DOWN - pad_x0;      // first zero-stuffed coordinate of the tile
  // floor(u / UP) for possibly negative u (UP is 1 or 2)
  const int iy_base = (UP == 1) ? uy0 : (uy0 >> 1), ix_base = (UP == 1) ? ux0 : (ux0 >> 1);
  const float* ip = in + plane * (int64_t)in_h * in_w;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < WR; r += 8) {
    const int iy = iy_base + r;
    const bool row_ok = iy >= 0 && iy < in_h;
    const float* row = ip + (int64_t)iy * in_w;
    for (int c = lane; c < WC; c += 32) {
      const int ix = ix_base + c;
      S[r * WC + c] = (row_ok && ix >= 0 && ix < in_w) ? __ldg(row + ix) : 0.f;
    }
  }
  __syncthreads();
  float k[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) k[i] = sk[i];
  const int cg = threadIdx.x & 31, rg = threadIdx.x >> 5;     // 32 column groups x 8 row groups, 4 x 4 outputs each
  const int oxl = cg * 4, oyl = rg * 4;
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
  if (UP == 1) {
    // u = o*DOWN + k - pad  ->  window row (oyl+j)*DOWN + ky, window col (oxl+i)*DOWN + kx
    constexpr int NR = 3 * DOWN + 4, NC = 3 * DOWN + 4;        // 7 (blur) or 10 (downsample)
    float w[NR][NC];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const float* sp = S + (oyl * DOWN + r) * WC + oxl * DOWN;   // oxl*DOWN is a multiple of 4: float4 loads
#pragma unroll
      for (int c4 = 0; c4 < NC / 4; ++c4) {
        const float4 v = *reinterpret_cast<const float4*>(sp + c4 * 4);
        w[r][c4 * 4] = v.x; w[r][c4 * 4 + 1] = v.y; w[r][c4 * 4 + 2] = v.z; w[r][c4 * 4 + 3] = v.w;
      }
#pragma unroll
      for (int c = NC / 4 * 4; c < NC; ++c) w[r][c] = sp[c];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int ky = 0; ky < 4; ++ky)
#pragma unroll
          for (int kx = 0; kx < 4; ++kx) acc[j][i] = fmaf(w[j * DOWN + ky][i * DOWN + kx], k[ky * 4 + kx], acc[j][i]);
  } else {
    // UP == 2, DOWN == 1: output o uses taps k = p, p + 2 with p = (o - pad) & 1 and inputs (o - pad + p) >> 1, +1
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ty = oy0 + oyl + j - pad_y0;
      const int py = ty & 1;
      const int sy = ((ty + py) >> 1) - iy_base;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int tx = ox0 + oxl + i - pad_x0;
        const int px = tx & 1;
        const int sx = ((tx + px) >> 1) - ix_base;
        const float* sp = S + sy * WC + sx;
        const float k00 = px ? (py ? k[5] : k[1]) : (py ? k[4] : k[0]);
        const float k01 = px ? (py ? k[7] : k[3]) : (py ? k[6] : k[2]);
        const float k10 = px ? (py ? k[13] : k[9]) : (py ? k[12] : k[8]);
        const float k11 = px ? (py ? k[15] : k[11]) : (py ? k[14] : k[10]);
        float a = sp[0] * k00;
        a = fmaf(sp[1], k01, a);
        a = fmaf(sp[WC], k10, a);
        a = fmaf(sp[WC + 1], k11, a);
        acc[j][i] = a;
      }
    }
  }
  float* op = out + plane * (int64_t)out_h * out_w;
  const int ox = ox0 + oxl;
  const bool vec = (out_w % 4 == 0) && ((reinterpret_cast<uintptr_t>(op) & 15) == 0) && (ox + 3 < out_w);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int oy = oy0 + oyl + j;
    if (oy >= out_h) break;
    float* orow = op + (int64_t)oy * out_w + ox;
    if (vec) {
      *reinterpret_cast<float4*>(orow) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (ox + i < out_w) orow[i] = acc[j][i];
    }
  }
}

template <int UP, int DOWN>
int launch_k4(const float* in, const float* kernel, float* out, int64_t planes, int in_h, int in_w, int out_h, int out_w,
              int pad_x0, int pad_y0, cudaStream_t st) {
  constexpr int WR = ((K4_OH - 1) * DOWN + 3) / UP + 2;
  constexpr int WC = (((K4_OW - 1) * DOWN + 3) / UP + 2 + 3) / 4 * 4;
  const size_t smem = (size_t)WR * WC * sizeof(float);
  if (smem > 48 * 1024)
    VT_CUDA(cudaFuncSetAttribute(upfirdn2d_k4_kernel<UP, DOWN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles_x = (int)vt_cdiv(out_w, K4_OW), tiles_y = (int)vt_cdiv(out_h, K4_OH);
  dim3 grid((unsigned)(tiles_x * tiles_y), (unsigned)planes);
  upfirdn2d_k4_kernel<UP, DOWN><<<grid, 256, smem, st>>>(in, kernel, out, in_h, in_w, out_h, out_w, pad_x0, pad_y0, tiles_x);
  return 0;
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
  if (g_upfirdn_tiled && kh == 4 && kw == 4 && up_x == up_y && down_x == down_y && planes <= 65535 &&
      ((up_x == 1 && down_x <= 2) || (up_x == 2 && down_x == 1))) {
    // the hot-path instances: compile-time up/down, 4 x 4 register micro-tiles
    int rc;
    if (up_x == 2) rc = launch_k4<2, 1>(in, kernel, out, planes, in_h, in_w, out_h, out_w, pad_x0, pad_y0, (cudaStream_t)stream);
    else if (down_x == 2) rc = launch_k4<1, 2>(in, kernel, out, planes, in_h, in_w, out_h, out_w, pad_x0, pad_y0, (cudaStream_t)stream);
    else rc = launch_k4<1, 1>(in, kernel, out, planes, in_h, in_w, out_h, out_w, pad_x0, pad_y0, (cudaStream_t)stream);
    if (rc) return rc;
    VT_LAUNCH_CHECK();
    return 0;
  }
  {
    // tiled fast path when the input window of a 32 x 64 output tile fits in shared memory
    const int tile_rows = ((T_OH - 1) * down_y + kh - 1) / up_y + 2;
    const int tile_cols = ((T_OW - 1) * down_x + kw - 1) / up_x + 2;
    const size_t smem = ((size_t)tile_rows * tile_cols + (size_t)kh * kw) * sizeof(float);
    if (g_upfirdn_tiled && smem <= 160 * 1024 && planes <= 65535) {
      const int tiles_x = (int)vt_cdiv(out_w, T_OW), tiles_y = (int)vt_cdiv(out_h, T_OH);
      if (smem > 48 * 1024)
        VT_CUDA(cudaFuncSetAttribute(upfirdn2d_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      dim3 grid((unsigned)(tiles_x * tiles_y), (unsigned)planes);
      upfirdn2d_tiled_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(in, kernel, out, in_h, in_w, out_h, out_w, kh, kw, up_x,
                                                                        up_y, down_x, down_y, pad_x0, pad_y0, tiles_x, tile_rows,
                                                                        tile_cols);
      VT_LAUNCH_CHECK();
      return 0;
    }
  }
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
