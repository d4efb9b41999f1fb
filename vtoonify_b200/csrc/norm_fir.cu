// norm_fir.cu — instance-norm statistics, AdaIN apply, and the NHWC FIR (Blur) with fused StyledConv epilogue.
//
//   vt_instnorm_stats_nhwc / vt_adain_apply_nhwc : AdaptiveInstanceNorm.forward (model/dualstylegan.py:16-21):
//        nn.InstanceNorm2d(affine=False) = (x - mean) / sqrt(biased_var + 1e-5) per (b, c) plane, then gamma*x + beta.
//        mode 1 evaluates the virtual concat cat(f_G, |f_G - f_E|) of Fusion.forward (model/vtoonify.py:125-126)
//        without materialising it for the statistics pass.
//   vt_fir_nhwc_f32 : Blur.forward after the stride-2 transposed conv (model/stylegan/model.py:74-90,285) =
//        upfirdn2d(x, k, pad=(p0,p1)) with up=down=1, on NHWC, with NoiseInjection + FusedLeakyReLU
//        (model/stylegan/model.py:315-320,364-370) applied in the same pass. Each thread owns 4 channels (float4)
//        of a 1 x 4 vertical strip of outputs so every input row is loaded once per strip.
#include "common.cuh"

int g_fir4 = 1;   // 1: specialised 4x4 pad (1,1) FIR kernel; 0: generic kernel (tests / A-B)

namespace {

// Deterministic two-stage reduction (no atomics): stage 1 — grid (chunks, B): every thread owns one float4 channel
// group and every pstep-th pixel of its chunk, partials are combined across the pixel lanes in a fixed order through
// shared memory and written to ws[chunk][b][Cs][2]; stage 2 sums the chunks in index order in double.
// dyn smem: pstep * Cs * 2 floats.
__global__ void __launch_bounds__(256)
instnorm_partial_kernel(const float* __restrict__ in, const float* __restrict__ in2, int mode, int64_t HW, int C,
                        int c_stride, int64_t chunk, float* __restrict__ ws) {
  extern __shared__ float sacc[];  // [pstep][Cs][2]
  const int Cs = mode ? 2 * C : C;
  const int b = blockIdx.y;
  const int nvec = C / 4;
  const int64_t p_begin = (int64_t)blockIdx.x * chunk;
  const int64_t p_end = (p_begin + chunk < HW) ? p_begin + chunk : HW;
  const float* ip = in + (int64_t)b * HW * c_stride;
  const float* ip2 = mode ? in2 + (int64_t)b * HW * c_stride : nullptr;
  const int pstep = blockDim.x / nvec;           // >= 1 (nvec <= 256 checked by the host)
  const int v = threadIdx.x % nvec, lane_p = threadIdx.x / nvec;
  if (lane_p < pstep) {
    float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0}, q2[4] = {0, 0, 0, 0};
    // four pixels per trip: the loads are issued together (the kernel is latency-bound: one 16-byte load in flight per thread
    // gave 1.2 TB/s on L2-resident maps), the accumulation order stays the pixel order
    int64_t p = p_begin + lane_p;
    for (; p + 3 * (int64_t)pstep < p_end; p += 4 * (int64_t)pstep) {
      float4 a[4], e[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = __ldg(reinterpret_cast<const float4*>(ip + (p + u * (int64_t)pstep) * c_stride + v * 4));
      if (mode) {
#pragma unroll
        for (int u = 0; u < 4; ++u) e[u] = __ldg(reinterpret_cast<const float4*>(ip2 + (p + u * (int64_t)pstep) * c_stride + v * 4));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float av[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { s[i] += av[i]; q[i] = fmaf(av[i], av[i], q[i]); }
        if (mode) {
          const float ev[4] = {e[u].x, e[u].y, e[u].z, e[u].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) { const float d = fabsf(av[i] - ev[i]); s2[i] += d; q2[i] = fmaf(d, d, q2[i]); }
        }
      }
    }
    for (; p < p_end; p += pstep) {
      const float4 a = *reinterpret_cast<const float4*>(ip + p * c_stride + v * 4);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) { s[i] += av[i]; q[i] = fmaf(av[i], av[i], q[i]); }
      if (mode) {
        const float4 e = *reinterpret_cast<const float4*>(ip2 + p * c_stride + v * 4);
        const float ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float d = fabsf(av[i] - ev[i]); s2[i] += d; q2[i] = fmaf(d, d, q2[i]); }
      }
    }
    float* row = sacc + (size_t)lane_p * Cs * 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      row[(v * 4 + i) * 2 + 0] = s[i];
      row[(v * 4 + i) * 2 + 1] = q[i];
      if (mode) {
        row[(C + v * 4 + i) * 2 + 0] = s2[i];
        row[(C + v * 4 + i) * 2 + 1] = q2[i];
      }
    }
  }
  __syncthreads();
  float* wrow = ws + ((int64_t)blockIdx.x * gridDim.y + b) * Cs * 2;
  for (int i = threadIdx.x; i < Cs * 2; i += blockDim.x) {
    float t = 0.f;
    for (int l = 0; l < pstep; ++l) t += sacc[(size_t)l * Cs * 2 + i];
    wrow[i] = t;
  }
}

// block = 32 (b, c) entries x 8 chunk slices: slice ks adds chunks ks, ks + 8, ... in double, the 8 slice sums are then added in
// slice order (fixed order -> deterministic).  One thread per entry walking all chunks was a 140-step dependent chain of
// strided loads (57 us per call on the 72x128 maps once the partial kernel went to 144 chunks).
__global__ void __launch_bounds__(256)
instnorm_finalize_kernel(const float* __restrict__ ws, float* __restrict__ stats, int n, int chunks, double inv_hw, float eps) {
  __shared__ double red[2][8][32];
  const int li = threadIdx.x & 31, ks = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + li;   // i = b * Cs + c
  double s = 0.0, q = 0.0;
  if (i < n) {
    const float2* w2 = reinterpret_cast<const float2*>(ws);
#pragma unroll 4
    for (int k = ks; k < chunks; k += 8) {
      const float2 v = __ldg(w2 + (int64_t)k * n + i);
      s += (double)v.x;
      q += (double)v.y;
    }
  }
  red[0][ks][li] = s; red[1][ks][li] = q;
  __syncthreads();
  if (ks == 0 && i < n) {
    double ss = 0.0, qq = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { ss += red[0][k][li]; qq += red[1][k][li]; }
    const double mean = ss * inv_hw;
    double var = qq * inv_hw - mean * mean;
    if (var < 0) var = 0;
    stats[i * 2] = (float)mean;
    stats[i * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// one thread per (pixel, float4 of the OUTPUT channels Cs)
__global__ void __launch_bounds__(256)
adain_apply_kernel(const float* __restrict__ in, const float* __restrict__ in2, int mode, int64_t HW, int C, int c_stride,
                   const float* __restrict__ stats, const float* __restrict__ gb, float* __restrict__ out, int round_tf32) {
  const int Cs = mode ? 2 * C : C;
  const int nvec = Cs / 4;
  const int b = blockIdx.y;
  const int64_t total = HW * nvec;
  const float* st = stats + (int64_t)b * Cs * 2;
  const float* gamma = gb + (int64_t)b * 2 * Cs;
  const float* beta = gamma + Cs;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / nvec;
    const int c = (int)(i % nvec) * 4;
    float x[4];
    if (!mode || c < C) {
      const float4 a = *reinterpret_cast<const float4*>(in + ((int64_t)b * HW + p) * c_stride + c);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
    } else {
      const float4 a = *reinterpret_cast<const float4*>(in + ((int64_t)b * HW + p) * c_stride + (c - C));
      const float4 e = *reinterpret_cast<const float4*>(in2 + ((int64_t)b * HW + p) * c_stride + (c - C));
      x[0] = fabsf(a.x - e.x); x[1] = fabsf(a.y - e.y); x[2] = fabsf(a.z - e.z); x[3] = fabsf(a.w - e.w);
    }
    float y[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xn = (x[k] - st[(c + k) * 2]) * st[(c + k) * 2 + 1];
      float v = gamma[c + k] * xn + beta[c + k];
      y[k] = round_tf32 ? vt_round_tf32(v) : v;
    }
    *reinterpret_cast<float4*>(out + ((int64_t)b * HW + p) * Cs + c) = make_float4(y[0], y[1], y[2], y[3]);
  }
}

constexpr int FIR_R = 4;      // output rows per thread
constexpr int FIR_MAXK = 8;   // max kernel extent

__global__ void __launch_bounds__(256)
fir_nhwc_kernel(const float* __restrict__ in, const float* __restrict__ kernel, float* __restrict__ out, int H, int W,
                int C, int Ho, int Wo, int kh, int kw, int pad0, const float* __restrict__ bias,
                const float* __restrict__ noise, const float* __restrict__ noise_w, int act, float slope, float gain,
                int round_tf32) {
  __shared__ float sk[FIR_MAXK * FIR_MAXK];  // flipped kernel: sk[ky][kx] multiplies in[oy+ky-pad0][ox+kx-pad0]
  if (threadIdx.x < kh * kw) {
    const int ky = threadIdx.x / kw, kx = threadIdx.x % kw;
    sk[threadIdx.x] = kernel[(kh - 1 - ky) * kw + (kw - 1 - kx)];
  }
  __syncthreads();
  const int nvec = C / 4;
  const int b = blockIdx.z;
  const int strips = (Ho + FIR_R - 1) / FIR_R;
  const int64_t total = (int64_t)strips * Wo * nvec;
  const float nw = noise ? *noise_w : 0.f;
  const float* ip = in + (int64_t)b * H * W * C;
  float* op = out + (int64_t)b * Ho * Wo * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    const int64_t r = i / nvec;
    const int ox = (int)(r % Wo);
    const int oy0 = (int)(r / Wo) * FIR_R;
    float4 acc[FIR_R];
#pragma unroll
    for (int j = 0; j < FIR_R; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    // input rows oy0 - pad0 .. oy0 + FIR_R - 1 + kh - 1 - pad0
    for (int ry = 0; ry < FIR_R + kh - 1; ++ry) {
      const int iy = oy0 - pad0 + ry;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < kw; ++kx) {
        const int ix = ox - pad0 + kx;
        if (ix < 0 || ix >= W) continue;
        const float4 a = *reinterpret_cast<const float4*>(ip + ((int64_t)iy * W + ix) * C + v * 4);
#pragma unroll
        for (int j = 0; j < FIR_R; ++j) {
          const int ky = ry - j;
          if (ky >= 0 && ky < kh) {
            const float w = sk[ky * kw + kx];
            acc[j].x = fmaf(a.x, w, acc[j].x); acc[j].y = fmaf(a.y, w, acc[j].y);
            acc[j].z = fmaf(a.z, w, acc[j].z); acc[j].w = fmaf(a.w, w, acc[j].w);
          }
        }
      }
    }
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bv = *reinterpret_cast<const float4*>(bias + v * 4);
#pragma unroll
    for (int j = 0; j < FIR_R; ++j) {
      const int oy = oy0 + j;
      if (oy >= Ho) break;
      float o[4] = {acc[j].x, acc[j].y, acc[j].z, acc[j].w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
      const float nz = noise ? nw * noise[((int64_t)b * Ho + oy) * Wo + ox] : 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float t = o[k];
        if (noise) t += nz;
        if (act) t = vt_lrelu(t + bb[k], slope) * gain;
        else if (bias) t += bb[k];
        o[k] = round_tf32 ? vt_round_tf32(t) : t;
      }
      *reinterpret_cast<float4*>(op + ((int64_t)oy * Wo + ox) * C + v * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

// 4x4 FIR, pad (1,1) (the Blur after a stride-2 transposed conv, model/stylegan/model.py:284-285): a thread owns 2 adjacent
// output columns x FIR4_R rows of one 4-channel group.  Per input row it issues 5 independent 16-byte loads (clamped address +
// zero mask, no branches), so 3.4 loads per output instead of 7 and all of them in flight together.
constexpr int FIR4_R = 8;

__global__ void __launch_bounds__(256)
fir4_nhwc_kernel(const float* __restrict__ in, const float* __restrict__ kernel, float* __restrict__ out, int H, int W,
                 int C, int Ho, int Wo, const float* __restrict__ bias, const float* __restrict__ noise,
                 const float* __restrict__ noise_w, int act, float slope, float gain, int round_tf32) {
  __shared__ float sk[16];  // flipped kernel: sk[ky][kx] multiplies in[oy+ky-1][ox+kx-1]
  if (threadIdx.x < 16) sk[threadIdx.x] = kernel[(3 - threadIdx.x / 4) * 4 + (3 - threadIdx.x % 4)];
  __syncthreads();
  float kk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) kk[i] = sk[i];
  const int nvec = C / 4;
  const int b = blockIdx.z;
  const int strips = (Ho + FIR4_R - 1) / FIR4_R, pairs = (Wo + 1) / 2;
  const int64_t total = (int64_t)strips * pairs * nvec;
  const float nw = noise ? *noise_w : 0.f;
  const float* ip = in + (int64_t)b * H * W * C;
  float* op = out + (int64_t)b * Ho * Wo * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    const int64_t r = i / nvec;
    const int ox0 = (int)(r % pairs) * 2;
    const int oy0 = (int)(r / pairs) * FIR4_R;
    float4 acc[FIR4_R][2];
#pragma unroll
    for (int j = 0; j < FIR4_R; ++j) acc[j][0] = acc[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    int cx[5];
    float mx[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int ix = ox0 - 1 + k;
      mx[k] = (ix >= 0 && ix < W) ? 1.f : 0.f;
      cx[k] = ix < 0 ? 0 : (ix >= W ? W - 1 : ix);
    }
#pragma unroll
    for (int ry = 0; ry < FIR4_R + 3; ++ry) {
      const int iy = oy0 - 1 + ry;
      const float my = (iy >= 0 && iy < H) ? 1.f : 0.f;
      const int cy = iy < 0 ? 0 : (iy >= H ? H - 1 : iy);
      const float* rowp = ip + (int64_t)cy * W * C + v * 4;
      float4 a[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        a[k] = __ldg(reinterpret_cast<const float4*>(rowp + (int64_t)cx[k] * C));
        const float m = my * mx[k];
        a[k].x *= m; a[k].y *= m; a[k].z *= m; a[k].w *= m;
      }
#pragma unroll
      for (int j = 0; j < FIR4_R; ++j) {
        const int ky = ry - j;
        if (ky >= 0 && ky < 4) {
#pragma unroll
          for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
              const float w = kk[ky * 4 + kx];
              acc[j][o].x = fmaf(a[o + kx].x, w, acc[j][o].x); acc[j][o].y = fmaf(a[o + kx].y, w, acc[j][o].y);
              acc[j][o].z = fmaf(a[o + kx].z, w, acc[j][o].z); acc[j][o].w = fmaf(a[o + kx].w, w, acc[j][o].w);
            }
        }
      }
    }
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bv = *reinterpret_cast<const float4*>(bias + v * 4);
    const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
    for (int j = 0; j < FIR4_R; ++j) {
      const int oy = oy0 + j;
      if (oy >= Ho) break;
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        const int ox = ox0 + o;
        if (ox >= Wo) continue;
        float t4[4] = {acc[j][o].x, acc[j][o].y, acc[j][o].z, acc[j][o].w};
        const float nz = noise ? nw * noise[((int64_t)b * Ho + oy) * Wo + ox] : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float t = t4[k];
          if (noise) t += nz;
          if (act) t = vt_lrelu(t + bb[k], slope) * gain;
          else if (bias) t += bb[k];
          t4[k] = round_tf32 ? vt_round_tf32(t) : t;
        }
        *reinterpret_cast<float4*>(op + ((int64_t)oy * Wo + ox) * C + v * 4) = make_float4(t4[0], t4[1], t4[2], t4[3]);
      }
    }
  }
}

}  // namespace

int g_instnorm_chunks = 296;   // target number of chunks per sample on large maps (0: always the small chunks)

static void instnorm_plan(int64_t HW, int C, int64_t* chunk, int64_t* chunks) {
  const int nvec = C / 4;
  const int plan = 256 / nvec;                      // pixels processed concurrently by a block
  int64_t ch = (int64_t)plan * 32;                  // 32 pixels per thread: several blocks per SM even on the 72x128 maps
  if (ch < 64) ch = 64;
  // Large maps: ~2 chunks per SM and sample instead of thousands of 8-trip blocks.  The partial pass itself is already
  // HBM-bound (ncu, [4,128,576,1024] x 2 sources: 2.42 GB in 344 us = 86 % DRAM throughput); what shrinks is the partial-sum
  // buffer and with it the finalize pass (2304 -> 296 chunks per entry).  The plan depends on (HW, C) only, never on the batch
  // size: a frame's statistics must not depend on the batch it travels in.
  if (g_instnorm_chunks > 0) {
    const int64_t unit = (int64_t)plan * 4;         // one trip of the main loop
    const int64_t want = vt_cdiv(vt_cdiv(HW, g_instnorm_chunks), unit) * unit;
    if (want > ch) ch = want;
  }
  *chunk = ch;
  *chunks = vt_cdiv(HW, ch);
}

extern "C" int64_t vt_instnorm_ws_bytes(int B, int64_t HW, int C, int mode) {
  if (B < 1 || HW < 1 || C < 4 || C % 4 || C / 4 > 256) return -1;
  int64_t chunk, chunks;
  instnorm_plan(HW, C, &chunk, &chunks);
  return chunks * B * (mode ? 2 * C : C) * 2 * (int64_t)sizeof(float);
}

extern "C" int vt_instnorm_stats_nhwc(const float* in, const float* in2, int mode, int B, int64_t HW, int C, int c_stride,
                                      float eps, float* stats, void* ws, void* stream) {
  VT_CHECK(in && stats && ws && (mode == 0 || (mode == 1 && in2)), "instnorm_stats: bad pointers/mode");
  VT_CHECK(B >= 1 && B <= 65535 && HW >= 1 && C >= 4 && C % 4 == 0 && c_stride >= C && c_stride % 4 == 0, "instnorm_stats: bad shape");
  VT_CHECK(C / 4 <= 256, "instnorm_stats: C must be <= 1024");
  const int Cs = mode ? 2 * C : C;
  cudaStream_t st = (cudaStream_t)stream;
  int64_t chunk, chunks;
  instnorm_plan(HW, C, &chunk, &chunks);
  const int pstep = 256 / (C / 4);
  dim3 grid((unsigned)chunks, (unsigned)B);
  instnorm_partial_kernel<<<grid, 256, (size_t)pstep * Cs * 2 * sizeof(float), st>>>(in, in2, mode, HW, C, c_stride, chunk,
                                                                                    (float*)ws);
  VT_LAUNCH_CHECK();
  const int n = B * Cs;
  instnorm_finalize_kernel<<<(unsigned)vt_cdiv(n, 32), 256, 0, st>>>((const float*)ws, stats, n, (int)chunks,
                                                                    1.0 / (double)HW, eps);
  VT_LAUNCH_CHECK();
  return 0;
}

extern "C" int vt_instnorm_finalize_f32(const float* ws, float* stats, int B, int Cs, int chunks, int64_t HW, float eps, void* stream) {
  VT_CHECK(ws && stats && B >= 1 && Cs >= 1 && chunks >= 1 && HW >= 1, "instnorm_finalize: bad args");
  const int n = B * Cs;
  instnorm_finalize_kernel<<<(unsigned)vt_cdiv(n, 32), 256, 0, (cudaStream_t)stream>>>(ws, stats, n, chunks, 1.0 / (double)HW, eps);
  VT_LAUNCH_CHECK();
  return 0;
}

__global__ void adain_affine_kernel(const float* __restrict__ stats, const float* __restrict__ gb, float* __restrict__ aff, int B, int Cs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // i = b * Cs + c
  if (i >= B * Cs) return;
  const int b = i / Cs, c = i - b * Cs;
  const float gamma = gb[(int64_t)b * 2 * Cs + c], beta = gb[(int64_t)b * 2 * Cs + Cs + c];
  const float sc = gamma * stats[i * 2 + 1];
  aff[i * 2] = sc;
  aff[i * 2 + 1] = beta - sc * stats[i * 2];
}

extern "C" int vt_adain_affine_f32(const float* stats, const float* gamma_beta, float* affine, int B, int Cs, void* stream) {
  VT_CHECK(stats && gamma_beta && affine && B >= 1 && Cs >= 1, "adain_affine: bad args");
  adain_affine_kernel<<<(unsigned)vt_cdiv((int64_t)B * Cs, 128), 128, 0, (cudaStream_t)stream>>>(stats, gamma_beta, affine, B, Cs);
  VT_LAUNCH_CHECK();
  return 0;
}

extern "C" int vt_adain_apply_nhwc(const float* in, const float* in2, int mode, int B, int64_t HW, int C, int c_stride,
                                   const float* stats, const float* gamma_beta, float* out, int round_tf32, void* stream) {
  VT_CHECK(in && stats && gamma_beta && out && (mode == 0 || (mode == 1 && in2)), "adain_apply: bad pointers/mode");
  VT_CHECK(B >= 1 && B <= 65535 && HW >= 1 && C >= 4 && C % 4 == 0 && c_stride >= C && c_stride % 4 == 0, "adain_apply: bad shape");
  const int Cs = mode ? 2 * C : C;
  const int64_t total = HW * (Cs / 4);
  int64_t blocks = vt_cdiv(total, 256);
  const int64_t cap = (int64_t)vt_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  dim3 grid((unsigned)blocks, (unsigned)B);
  adain_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, in2, mode, HW, C, c_stride, stats, gamma_beta, out, round_tf32);
  VT_LAUNCH_CHECK();
  return 0;
}

extern "C" int vt_fir_nhwc_f32(const float* in, const float* kernel, float* out, int B, int H, int W, int C, int kh, int kw,
                               int pad0, int pad1, const float* bias, const float* noise, const float* noise_w, int act,
                               float slope, float gain, int round_tf32, void* stream) {
  VT_CHECK(in && kernel && out, "fir_nhwc: null pointer");
  VT_CHECK(B >= 1 && B <= 65535 && H >= 1 && W >= 1 && C >= 4 && C % 4 == 0, "fir_nhwc: bad shape (C must be a multiple of 4)");
  VT_CHECK(kh >= 1 && kw >= 1 && kh <= FIR_MAXK && kw <= FIR_MAXK, "fir_nhwc: kernel extent must be <= %d", FIR_MAXK);
  VT_CHECK(pad0 >= 0 && pad1 >= 0, "fir_nhwc: negative pad not supported on the NHWC path");
  VT_CHECK(!noise || noise_w, "fir_nhwc: noise without noise_w");
  const int Ho = H + pad0 + pad1 - kh + 1, Wo = W + pad0 + pad1 - kw + 1;
  VT_CHECK(Ho >= 1 && Wo >= 1, "fir_nhwc: empty output");
  if (kh == 4 && kw == 4 && pad0 == 1 && pad1 == 1 && g_fir4) {
    const int64_t total4 = (int64_t)vt_cdiv(Ho, FIR4_R) * vt_cdiv(Wo, 2) * (C / 4);
    int64_t blocks4 = vt_cdiv(total4, 256);
    const int64_t cap4 = (int64_t)vt_num_sms() * 16;
    if (blocks4 > cap4) blocks4 = cap4;
    fir4_nhwc_kernel<<<dim3((unsigned)blocks4, 1, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(
        in, kernel, out, H, W, C, Ho, Wo, bias, noise, noise_w, act, slope, gain, round_tf32);
    VT_LAUNCH_CHECK();
    return 0;
  }
  const int strips = (Ho + FIR_R - 1) / FIR_R;
  const int64_t total = (int64_t)strips * Wo * (C / 4);
  int64_t blocks = vt_cdiv(total, 256);
  const int64_t cap = (int64_t)vt_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  dim3 grid((unsigned)blocks, 1, (unsigned)B);
  fir_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, kernel, out, H, W, C, Ho, Wo, kh, kw, pad0, bias, noise,
                                                         noise_w, act, slope, gain, round_tf32);
  VT_LAUNCH_CHECK();
  return 0;
}
