// modulate.cu — per-style constants of the modulated convolution and the small linears.
//
//   vt_linear_f32           : EqualLinear.forward (model/stylegan/model.py:153-162): F.linear(x, W*scale, b*lr_mul)
//                             [+ fused_leaky_relu].  One warp per output element, shuffle reduction.
//   vt_modulate_weights_f32 : ModulatedConv2d fused branch (model/stylegan/model.py:259-267):
//                             w'[b,n,c,t] = (scale*W[n,c,t]) * s[b,c];  demod[b,n] = rsqrt(sum_{c,t} w'^2 + 1e-8);
//                             written in the conv kernels' layout [b][t][n][c_pad] (K-major rows of one tap,
//                             128-byte multiples so a TMA box of 32 channels is one swizzle row).
//                             With style == NULL it is the plain re-layout used for nn.Conv2d / EqualConv2d weights.
#include "common.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace {

__global__ void __launch_bounds__(256)
linear_kernel(const float* __restrict__ in, const float* __restrict__ weight, const float* __restrict__ bias,
              float* __restrict__ out, int rows, int in_dim, int out_dim, float w_scale, float b_scale, int act) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows * out_dim) return;
  const int r = warp / out_dim, o = warp % out_dim;
  const float* x = in + (int64_t)r * in_dim;
  const float* w = weight + (int64_t)o * in_dim;
  float acc = 0.f;
  for (int i = lane; i < in_dim; i += 32) acc += x[i] * (w[i] * w_scale);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) {
    float v = acc;
    const float b = bias ? bias[o] * b_scale : 0.f;
    if (act == 1) v = vt_lrelu(v + b, 0.2f) * 1.4142135623730951f;
    else if (act == 2) v = vt_lrelu(v + b, 0.2f);
    else if (act == 3) v = fmaxf(v + b, 0.f);
    else if (act == 4) v = 1.0f / (1.0f + expf(-(v + b)));
    else v = v + b;
    out[(int64_t)r * out_dim + o] = v;
  }
}

// PixelNorm (model/stylegan/model.py:13-18): x * rsqrt(mean(x^2, dim=1) + 1e-8); one warp per row
__global__ void __launch_bounds__(256)
pixelnorm_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int dim) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* x = in + (int64_t)warp * dim;
  float ss = 0.f;
  for (int i = lane; i < dim; i += 32) ss = fmaf(x[i], x[i], ss);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
  const float r = 1.0f / sqrtf(ss / (float)dim + 1e-8f);
  for (int i = lane; i < dim; i += 32) out[(int64_t)warp * dim + i] = x[i] * r;
}

// grid (Cout, wB); block 256. Each block owns one (b, n) filter row of Cin*taps weights.
__global__ void __launch_bounds__(256)
modulate_weights_kernel(const float* __restrict__ W, const float* __restrict__ style, float* __restrict__ out,
                        int Cout, int Cin, int taps, int cin_pad, float scale, int demodulate, int round_tf32) {
  const int n = blockIdx.x, b = blockIdx.y;
  const float* wrow = W + (int64_t)n * Cin * taps;
  const float* srow = style ? style + (int64_t)b * Cin : nullptr;
  __shared__ float red[32];
  __shared__ float s_demod;
  float d = 1.f;
  if (demodulate) {
    float ss = 0.f;
    for (int i = threadIdx.x; i < Cin * taps; i += blockDim.x) {
      const int c = i / taps;
      float w = scale * wrow[i];
      if (srow) w = w * srow[c];
      ss += w * w;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
      float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
      if (threadIdx.x == 0) s_demod = 1.0f / sqrtf(v + 1e-8f);
    }
    __syncthreads();
    d = s_demod;
  }
  float* obase = out + ((int64_t)b * taps * Cout + n) * cin_pad;  // + t*Cout*cin_pad + c
  for (int i = threadIdx.x; i < cin_pad * taps; i += blockDim.x) {
    const int t = i / cin_pad, c = i % cin_pad;
    float v = 0.f;
    if (c < Cin) {
      v = scale * wrow[c * taps + t];
      if (srow) v = v * srow[c];
      if (demodulate) v = v * d;
      if (round_tf32) v = vt_round_tf32(v);
    }
    obase[(int64_t)t * Cout * cin_pad + c] = v;
  }
}

// Fold Blur o conv_transpose2d(stride 2) into 4 phase kernels: one thread per (b, n, c) reads the 9 modulated weights and
// writes the 36 folded ones.  blur: out[o] = sum_k kf[k] * T[o + k - 1], kf = flipped 4-tap kernel; T[2i + kap] += x[i] w[kap]
//   => G_r[dlt] = sum_kap w[kap] * kf[2*dlt + kap - r + 1]   (r = output parity, dlt = input offset in {-1,0,1}), per axis.
__global__ void __launch_bounds__(256)
fold_upconv_kernel(const float* __restrict__ w, const float* __restrict__ blur, float* __restrict__ out, int wB, int Cout,
                   int cpad, int round_tf32) {
  __shared__ float kf[4][4];
  if (threadIdx.x < 16) kf[threadIdx.x / 4][threadIdx.x % 4] = blur[(3 - threadIdx.x / 4) * 4 + (3 - threadIdx.x % 4)];
  __syncthreads();
  const int64_t per_b = (int64_t)Cout * cpad;
  const int64_t total = (int64_t)wB * per_b;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / per_b, nc = i % per_b;
    float wv[3][3];
#pragma unroll
    for (int t = 0; t < 9; ++t) wv[t / 3][t % 3] = w[(b * 9 + t) * per_b + nc];
#pragma unroll
    for (int ry = 0; ry < 2; ++ry)
#pragma unroll
      for (int rx = 0; rx < 2; ++rx)
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            float g = 0.f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const int iy = 2 * dy + ky - ry + 1;
              if (iy < 0 || iy > 3) continue;
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const int ix = 2 * dx + kx - rx + 1;
                if (ix < 0 || ix > 3) continue;
                g = fmaf(wv[ky][kx], kf[iy][ix], g);
              }
            }
            // out layout [wB][9 taps][4 phases][Cout][cpad]: the 4 phases are stacked along the GEMM N dimension
            const int slab = (dy + 1) * 3 + (dx + 1), ph = ry * 2 + rx;
            out[((b * 9 + slab) * 4 + ph) * per_b + nc] = round_tf32 ? vt_round_tf32(g) : g;
          }
  }
}

}  // namespace

extern "C" int vt_fold_upconv_weights_f32(const float* w, const float* blur, float* out, int wB, int Cout, int cpad,
                                          int round_tf32, void* stream) {
  VT_CHECK(w && blur && out && wB >= 1 && Cout >= 1 && cpad >= 1, "fold_upconv_weights: bad args");
  const int64_t total = (int64_t)wB * Cout * cpad;
  int64_t blocks = vt_cdiv(total, 256);
  if (blocks > (int64_t)vt_num_sms() * 8) blocks = (int64_t)vt_num_sms() * 8;
  fold_upconv_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(w, blur, out, wB, Cout, cpad, round_tf32);
  VT_LAUNCH_CHECK();
  return 0;
}

extern "C" int vt_linear_f32(const float* in, const float* weight, const float* bias, float* out, int rows, int in_dim,
                             int out_dim, float w_scale, float b_scale, int act, void* stream) {
  VT_CHECK(in && weight && out && rows >= 1 && in_dim >= 1 && out_dim >= 1, "linear: bad args");
  VT_CHECK(act >= 0 && act <= 4, "linear: act must be in [0, 4]");
  const int64_t warps = (int64_t)rows * out_dim;
  const int threads = 256;
  const int64_t blocks = vt_cdiv(warps * 32, threads);
  VT_CHECK(blocks < (1LL << 31), "linear: too large");
  linear_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(in, weight, bias, out, rows, in_dim, out_dim,
                                                                       w_scale, b_scale, act);
  VT_LAUNCH_CHECK();
  return 0;
}

extern "C" int vt_pixelnorm_f32(const float* in, float* out, int rows, int dim, void* stream) {
  VT_CHECK(in && out && rows >= 1 && dim >= 1, "pixelnorm: bad args");
  pixelnorm_kernel<<<(unsigned)vt_cdiv((int64_t)rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(in, out, rows, dim);
  VT_LAUNCH_CHECK();
  return 0;
}

extern "C" int vt_modulate_weights_f32(const float* W, const float* style, float* out, int wB, int Cout, int Cin,
                                       int kh, int kw, int cin_pad, float scale, int demodulate, int round_tf32,
                                       void* stream) {
  VT_CHECK(W && out, "modulate_weights: null pointer");
  VT_CHECK(wB >= 1 && Cout >= 1 && Cin >= 1 && kh >= 1 && kw >= 1 && cin_pad >= Cin, "modulate_weights: bad shape");
  VT_CHECK(style || wB == 1, "modulate_weights: style == NULL requires wB == 1");
  VT_CHECK(wB <= 65535, "modulate_weights: batch too large");
  dim3 grid((unsigned)Cout, (unsigned)wB);
  modulate_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(W, style, out, Cout, Cin, kh * kw, cin_pad, scale,
                                                                 demodulate, round_tf32);
  VT_LAUNCH_CHECK();
  return 0;
}

// ---- bf16x3 weight split -----------------------------------------------------------------------------------------
// one thread per (row, 32-channel chunk, 4-channel group): reads a float4, writes 2 bf16x4 halves.
// nstack_rows == 0: out row = in row, chunk = [hi(32) | lo(32)].  nstack_rows = R > 0: input rows are taken in groups of R;
// group g becomes R rows [hi|hi] followed by R rows [lo|lo] (output row stride unchanged, twice as many rows).
__global__ void __launch_bounds__(256)
split_bf16x3_kernel(const float* __restrict__ w, uint2* __restrict__ out, int64_t n_quads, int C, int nstack_rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_quads) return;
  const float4 v = __ldg(reinterpret_cast<const float4*>(w) + i);
  const float f[4] = {v.x, v.y, v.z, v.w};
  unsigned short h[4], l[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat16 hb = __float2bfloat16_rn(f[k]);
    h[k] = __bfloat16_as_ushort(hb);
    l[k] = __bfloat16_as_ushort(__float2bfloat16_rn(f[k] - __bfloat162float(hb)));
  }
  const uint2 hq = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
  const uint2 lq = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
  const int64_t chunk = i >> 3;          // 8 quads per 32-channel chunk
  const int q = (int)(i & 7);
  if (nstack_rows == 0) {
    uint2* base = out + chunk * 16;      // a chunk is 128 bytes = 16 uint2
    base[q] = hq;
    base[8 + q] = lq;
  } else {
    const int cpr = C / 32;              // chunks per row
    const int64_t row = chunk / cpr;
    const int cc = (int)(chunk - row * cpr);
    const int64_t grp = row / nstack_rows, rr = row - grp * nstack_rows;
    uint2* hi_row = out + (((grp * 2) * nstack_rows + rr) * cpr + cc) * 16;
    uint2* lo_row = out + (((grp * 2 + 1) * nstack_rows + rr) * cpr + cc) * 16;
    hi_row[q] = hq; hi_row[8 + q] = hq;
    lo_row[q] = lq; lo_row[8 + q] = lq;
  }
}

extern "C" int vt_split_weights_bf16x3(const float* w, void* out, int64_t rows, int C, int nstack_rows, void* stream) {
  VT_CHECK(w && out && rows >= 1 && C >= 32 && C % 32 == 0, "split_weights_bf16x3: bad args (rows=%lld C=%d)", (long long)rows, C);
  VT_CHECK(nstack_rows >= 0 && (nstack_rows == 0 || rows % nstack_rows == 0), "split_weights_bf16x3: rows must be a multiple of nstack_rows");
  VT_CHECK(((uintptr_t)w & 15) == 0 && ((uintptr_t)out & 15) == 0, "split_weights_bf16x3: pointers must be 16-byte aligned");
  const int64_t n_quads = rows * C / 4;
  split_bf16x3_kernel<<<(unsigned)vt_cdiv(n_quads, 256), 256, 0, (cudaStream_t)stream>>>(w, (uint2*)out, n_quads, C, nstack_rows);
  VT_LAUNCH_CHECK();
  return 0;
}

// ---- fp16 split of weight rows: out chunk = [half(w * scale) x 32 | half(w * scale - hi) x 32] ---------------------------------
// fp16 keeps 11 + 11 mantissa bits (bf16: 8 + 8), i.e. the split represents w to 2^-22 as long as both halves stay in fp16's
// normal range: `scale` (a power of two, undone by the consumer's acc_scale) lifts the small demodulated weights away from the
// subnormals; |w * scale| must stay below 65504.
__global__ void __launch_bounds__(256)
split_f16x3_kernel(const float* __restrict__ w, uint2* __restrict__ out, int64_t n_quads, float scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_quads) return;
  const float4 v = __ldg(reinterpret_cast<const float4*>(w) + i);
  const float f[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
  unsigned short h[4], l[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __half hb = __float2half_rn(f[k]);
    h[k] = __half_as_ushort(hb);
    l[k] = __half_as_ushort(__float2half_rn(f[k] - __half2float(hb)));
  }
  const uint2 hq = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
  const uint2 lq = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
  const int64_t chunk = i >> 3;
  const int q = (int)(i & 7);
  uint2* base = out + chunk * 16;
  base[q] = hq;
  base[8 + q] = lq;
}

extern "C" int vt_split_weights_f16x3(const float* w, void* out, int64_t rows, int C, float scale, void* stream) {
  VT_CHECK(w && out && rows >= 1 && C >= 32 && C % 32 == 0, "split_weights_f16x3: bad args (rows=%lld C=%d)", (long long)rows, C);
  VT_CHECK(scale > 0.f, "split_weights_f16x3: scale must be positive");
  VT_CHECK(((uintptr_t)w & 15) == 0 && ((uintptr_t)out & 15) == 0, "split_weights_f16x3: pointers must be 16-byte aligned");
  const int64_t n_quads = rows * C / 4;
  split_f16x3_kernel<<<(unsigned)vt_cdiv(n_quads, 256), 256, 0, (cudaStream_t)stream>>>(w, (uint2*)out, n_quads, scale);
  VT_LAUNCH_CHECK();
  return 0;
}
