// conv_rsu.cu — row-strip UP-convolution: Blur(4x4, pad (1,1)) o conv_transpose2d(stride 2, 3x3) of the StyledConv up-layers
// (model/stylegan/model.py:273-286) for the last two generator levels (64 -> 32 @ 2H x 2W -> 4H x 4W, 128 -> 64 @ H x W).
//
// The folded form used so far (one 3x3 convolution with N = 4 * Cout: vt_fold_upconv_weights_f32) issues 4x the algorithmic
// MACs and is bound by the tensor pipe.  Here only the HORIZONTAL half of the separable blur is folded into the weights
// (2x the MACs) and the vertical half runs on the accumulators:
//
//   t   = conv_transpose2d(x, W, stride 2)                       [(2H+1) x (2W+1)]
//   tx[p][2J+px] = sum_n g[n] t[p][2J+px+n-1]                   x-blurred rows p = 0 .. 2H  (g = flipped 1-D blur taps)
//               = sum_{(i,ky): 2i+ky=p} sum_{dj=-1..1} x[i][J+dj] . Gx[ky][px][dj],   Gx = sum_kx W[ky][kx] g[2dj+kx-px+1]
//   out[Y][X]   = sum_{m=0..3} g[m] tx[Y+m-1][X]                 Y = 0 .. 2H-1
//
// One MMA per (input row i, dx = dj, 32-channel chunk, product):  D[128 px of row i, (ky, px, co)] += X[row i, px+dj] * Gx,
// N = 3 * 2 * 32 = 192: an input row adds into the three tx rows 2i, 2i+1, 2i+2, which are slots of a TMEM ring (64 columns
// each: [px=0 | px=1] x 32 output channels; slot of row p = (p + 2) mod 6, one mirror slot behind the ring keeps every window
// contiguous).  The MMAs always accumulate; the epilogue forms the two output rows of an input row as 4-tap vertical blurs of five
// finished tx rows (TMEM -> registers, one pass per pixel phase, the two dying rows first), applies noise / bias / leaky-relu, writes
// the two interleaved pixel phases through a swizzled staging tile + TMA store, and zeroes a tx row once it has been read for the
// last time (its slot is what the MMA issuer's next input row is waiting for).  The slot of a row depends on
// its image row only, so results do not depend on the strip partition or the batch composition.
// A launch produces 32 output channels; Cout = 64 is two passes over the input (the second reads it from L2 / HBM again).
// Weight tiles (12 KB per (chunk, dx) and CTA of a pair) stream through a 6-deep ring like the activations (Cin = 128 needs 144 KB of
// them per row; at Cin = 64 the ring holds a whole row's tiles).
// Roles (10 warps): 0 TMA producer | 1 MMA issuer | 2,3,8,9 operand transform | 4-7 epilogue.
#include "tc_common.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <mutex>

using namespace vt_tc;

int vt_validate_conv_desc(const vt_conv_desc* d, const char* who);
extern unsigned long long* g_tc_dbg_export;

namespace {

constexpr int RU_PX = 128;
constexpr int RU_BOX = RU_PX + 2;
constexpr int RU_A_STAGE = 17 * 1024;
constexpr int RU_THREADS = 320;
constexpr int RU_XFORM_WARPS = 4;
constexpr int RU_MAX_SMEM = 227 * 1024;
constexpr int RU_N = 192;                      // 3 tx rows x 2 pixel phases x 32 output channels
constexpr int RU_SLOT = 64;                    // TMEM columns per tx row
constexpr int RU_S = 6;                        // logical slots (physical 7: slot 6 mirrors slot 0)
constexpr int RU_B_STAGES_MAX = 8;             // weight-tile ring: 6 stages by default (see vt_conv_up2_rs), 4 in the first version
constexpr int RU_STAGING = 4 * 2 * 8192;       // per epilogue warp: two 64 px x 128 B buffers

struct RuArgs {
  CUtensorMap in_map;      // fp32 (cstride, W, H, B), box (32, 130, 1, 1)
  CUtensorMap w_map;       // 16-bit (64, 192, n_half*KC*3, wB), box (64, 192/CG, 1, 1)
  CUtensorMap out_map;     // fp32 (Cout_total, 2W, 2H, B), box (32, 64, 1, 1)
  int B, H, W, KC, wB;
  int half, c_base;        // output-channel slice of this pass: tile base = half*KC*3, channels [c_base, c_base + 32)
  int rows_per_strip, strips_x, strips_y, total_strips;
  int a_stages, b_stages;
  int b_tile_bytes;        // (192 / CG) rows x 128 B
  const float* bias; const float* noise; const float* noise_w;
  int act; float slope, gain;
  float g[4];              // flipped 1-D blur taps (vertical pass)
  int fmt; float acc_scale;
  unsigned long long* dbg;
};

__device__ __forceinline__ void ru_zero32(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(z)
      : "memory");
}
__device__ __forceinline__ void ru_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ int ru_slot(int p) { return (p + 2) % RU_S; }   // p >= -2

#define RU_TWAIT(slot, stmt) do { if (p.dbg) { const long long t__ = clock64(); stmt; tw[slot] += clock64() - t__; } else { stmt; } } while (0)

// EPI = 0: one output row at a time (4 accumulator loads per row and pixel phase, each waited for on its own);
// EPI = 1: both output rows of an input row from ONE pass over the five tx rows per pixel phase (10 loads instead of 16, two in
//          flight per wait), the two dying rows first, so their slots return to the MMA issuer after 7 loads instead of 15.
template <int CG, int EPI>
__global__ void __launch_bounds__(RU_THREADS, 1)
conv_rsu_kernel(const __grid_constant__ RuArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + (uint32_t)p.a_stages * RU_A_STAGE;
  const uint32_t st_base = (b_base + (uint32_t)p.b_stages * (uint32_t)p.b_tile_bytes + 1023u) & ~1023u;
  const uint32_t cst_base = st_base + RU_STAGING;
  const uint32_t bar_base = cst_base + 32 * 4;
  // barriers: a_full[8] a_ready[8] a_empty[8] b_full[8] row_full[8] row_empty[8] | tmem slot | b_empty[8]
  auto a_full = [&](int i) { return bar_base + 8u * i; };
  auto a_ready = [&](int i) { return bar_base + 64u + 8u * i; };
  auto a_empty = [&](int i) { return bar_base + 128u + 8u * i; };
  auto b_full = [&](int i) { return bar_base + 192u + 8u * i; };
  auto b_empty = [&](int i) { return bar_base + 392u + 8u * i; };
  auto row_full = [&](int i) { return bar_base + 256u + 8u * i; };
  auto row_empty = [&](int i) { return bar_base + 320u + 8u * i; };
  const uint32_t tmem_slot = bar_base + 384u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  float* cst = reinterpret_cast<float*>(smem_gen + (cst_base - smem_base));   // bias of this pass's 32 channels

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int cta_i = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int cta_n = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  long long tw[4] = {0, 0, 0, 0};
  const long long t_begin = clock64();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.in_map); tma_prefetch_desc(&p.w_map); tma_prefetch_desc(&p.out_map);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.a_stages; ++i) { mbar_init(a_full(i), 1); mbar_init(a_ready(i), RU_XFORM_WARPS * CG); mbar_init(a_empty(i), 1); }
    for (int i = 0; i < p.b_stages; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
    for (int i = 0; i < 8; ++i) { mbar_init(row_full(i), 1); mbar_init(row_empty(i), 4 * CG); }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 2) {
    if (CG == 2) tmem_alloc_2sm(tmem_slot, 512u); else tmem_alloc(tmem_slot, 512u);
    tc_fence_before();
  }
  if (warp >= 4 && warp < 8) {
    const int r = (warp - 4) * 32 + lane;
    if (r < 32) cst[r] = p.bias ? __ldg(p.bias + p.c_base + r) : 0.f;
  }
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  if (warp >= 4 && warp < 8) {
    const uint32_t t0 = tmem_base + ((uint32_t)((warp - 4) * 32) << 16);
    for (int c = 0; c < 512; c += 32) ru_zero32(t0 + (uint32_t)c);
    ru_wait_st();
    tc_fence_before();
  }
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();

  const int strips_per_img = p.strips_y * p.strips_x;
  auto strip_geom = [&](int strip, int& b, int& i0, int& x0, int& R) {
    b = strip / strips_per_img;
    const int rem = strip - b * strips_per_img;
    const int ys = rem / p.strips_x, xs = rem - ys * p.strips_x;
    i0 = ys * p.rows_per_strip;                 // first input row whose two output rows this strip produces
    R = min(p.rows_per_strip, p.H - i0);
    x0 = xs * (RU_PX * CG) + (int)rank * RU_PX;
  };
  // A strip produces output rows 2*i0 .. 2*(i0+R)-1 from the R+2 input rows i0-1 .. i0+R (index k, i = i0-1+k).  Input row i adds
  // into tx rows 2i, 2i+1, 2i+2; tx rows 2i and 2i+1 are final once input i has been issued.  Output rows 2i-2 and 2i-1 need tx
  // rows 2i-3 .. 2i+1 and are produced after input i (k >= 2); tx rows 2i-3 and 2i-2 are dead afterwards.

  if (warp == 0) {
    // ================= TMA producer: per input row and chunk one activation box, then its three weight tiles =================
    int a_st = 0, b_st = 0; uint32_t a_par = 0, b_par = 0;
    for (int strip = cta_i; strip < p.total_strips; strip += cta_n) {
      int b, i0, x0, R;
      strip_geom(strip, b, i0, x0, R);
      const int wb = p.wB > 1 ? b : 0;
      for (int k = 0; k < R + 2; ++k) {
        const int y = i0 - 1 + k;
        for (int kc = 0; kc < p.KC; ++kc) {
          RU_TWAIT(0, mbar_wait(a_empty(a_st), a_par ^ 1, 31));
          if (elect_one()) {
            mbar_arrive_expect_tx(a_full(a_st), (uint32_t)(RU_BOX * 128));
            tma_load_4d(a_base + (uint32_t)a_st * RU_A_STAGE, &p.in_map, a_full(a_st), kc * 32, x0 - 1, y, b);
          }
          __syncwarp();
          if (++a_st == p.a_stages) { a_st = 0; a_par ^= 1; }
          for (int dxi = 0; dxi < 3; ++dxi) {
            RU_TWAIT(1, mbar_wait(b_empty(b_st), b_par ^ 1, 32));
            if (elect_one()) {
              const uint32_t dst = b_base + (uint32_t)b_st * (uint32_t)p.b_tile_bytes;
              const int tile = (p.half * p.KC + kc) * 3 + dxi;
              if (CG == 2) {
                if (rank == 0) mbar_arrive_expect_tx(b_full(b_st), 2u * (uint32_t)p.b_tile_bytes);
                tma_load_4d_2sm(dst, &p.w_map, b_full(b_st), 0, (int)rank * (RU_N / 2), tile, wb);
              } else {
                mbar_arrive_expect_tx(b_full(b_st), (uint32_t)p.b_tile_bytes);
                tma_load_4d(dst, &p.w_map, b_full(b_st), 0, 0, tile, wb);
              }
            }
            __syncwarp();
            if (++b_st == p.b_stages) { b_st = 0; b_par ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ================= MMA issuer =================
    const uint32_t idesc = p.fmt ? make_idesc_f16(RU_PX * CG, RU_N) : make_idesc_bf16(RU_PX * CG, RU_N);
    int a_st = 0, b_st = 0; uint32_t a_par = 0, b_par = 0;
    uint32_t acq = 0;
    auto acquire = [&](int q) {
      RU_TWAIT(2, mbar_wait(row_empty(q), ((acq >> q) & 1u) ^ 1u, 33));
      acq ^= 1u << q;
    };
    for (int strip = cta_i; strip < p.total_strips; strip += cta_n) {
      int b, i0, x0, R;
      strip_geom(strip, b, i0, x0, R);
      for (int k = 0; k < R + 2; ++k) {
        const int i = i0 - 1 + k;
        const int q0 = ru_slot(2 * i), q1 = ru_slot(2 * i + 1), q2 = ru_slot(2 * i + 2);
        if (k == 0) acquire(q0);
        acquire(q1);
        acquire(q2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(q0 * RU_SLOT);   // q0 in {0, 2, 4}: window q0 .. q0+2 (slot 6 mirrors 0)
        for (int kc = 0; kc < p.KC; ++kc) {
          RU_TWAIT(0, mbar_wait(a_ready(a_st), a_par, 34));
          const uint32_t a_addr = a_base + (uint32_t)a_st * RU_A_STAGE;
          for (int dxi = 0; dxi < 3; ++dxi) {
            RU_TWAIT(1, mbar_wait(b_full(b_st), b_par, 35));
            tc_fence_after();
            if (elect_one()) {
              const uint64_t adesc = make_smem_desc_sw128(a_addr + (uint32_t)dxi * 128u, 1024, 0);
              const uint64_t bdesc = make_smem_desc_sw128(b_base + (uint32_t)b_st * (uint32_t)p.b_tile_bytes, 1024, 0);
              if (CG == 2) {
                umma_bf16_2sm(d_tmem, adesc, bdesc, idesc, 1);
                umma_bf16_2sm(d_tmem, adesc + 2, bdesc + 2, idesc, 1);
                umma_bf16_2sm(d_tmem, adesc + 4, bdesc, idesc, 1);
                umma_bf16_2sm(d_tmem, adesc + 6, bdesc + 2, idesc, 1);
                umma_bf16_2sm(d_tmem, adesc, bdesc + 4, idesc, 1);
                umma_bf16_2sm(d_tmem, adesc + 2, bdesc + 6, idesc, 1);
                umma_commit_2sm(b_empty(b_st));
                if (dxi == 2) umma_commit_2sm(a_empty(a_st));
              } else {
                umma_bf16(d_tmem, adesc, bdesc, idesc, 1);
                umma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1);
                umma_bf16(d_tmem, adesc + 4, bdesc, idesc, 1);
                umma_bf16(d_tmem, adesc + 6, bdesc + 2, idesc, 1);
                umma_bf16(d_tmem, adesc, bdesc + 4, idesc, 1);
                umma_bf16(d_tmem, adesc + 2, bdesc + 6, idesc, 1);
                umma_commit(b_empty(b_st));
                if (dxi == 2) umma_commit(a_empty(a_st));
              }
            }
            __syncwarp();
            if (++b_st == p.b_stages) { b_st = 0; b_par ^= 1; }
          }
          if (++a_st == p.a_stages) { a_st = 0; a_par ^= 1; }
        }
        // tx rows 2i and 2i+1 are final; after the strip's last input row nothing else will touch 2i+2 either
        if (elect_one()) {
          if (CG == 2) { umma_commit_2sm(row_full(q0)); umma_commit_2sm(row_full(q1)); if (k == R + 1) umma_commit_2sm(row_full(q2)); }
          else { umma_commit(row_full(q0)); umma_commit(row_full(q1)); if (k == R + 1) umma_commit(row_full(q2)); }
        }
        __syncwarp();
      }
    }
  } else if (warp == 2 || warp == 3 || warp >= 8) {
    // ================= operand transform: fp32 rows -> [hi(32) | lo(32)] 16-bit rows, in place =================
    const int t = (warp < 4 ? warp - 2 : warp - 6) * 32 + lane;
    int a_st = 0; uint32_t a_par = 0;
    for (int strip = cta_i; strip < p.total_strips; strip += cta_n) {
      int b, i0, x0, R;
      strip_geom(strip, b, i0, x0, R);
      const int loads = (R + 2) * p.KC;
      for (int l = 0; l < loads; ++l) {
        RU_TWAIT(0, mbar_wait(a_full(a_st), a_par, 36));
        const uint32_t stage = a_base + (uint32_t)a_st * RU_A_STAGE;
        for (int r = t; r < RU_BOX; r += 32 * RU_XFORM_WARPS) {
          const uint32_t row = stage + (uint32_t)r * 128u;
          const uint32_t ph = (row >> 7) & 7u;
          float f[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(row + ((j ^ ph) << 4)));
            f[4 * j] = v.x; f[4 * j + 1] = v.y; f[4 * j + 2] = v.z; f[4 * j + 3] = v.w;
          }
          uint32_t hi[16], lo[16];
          if (p.fmt) {
#pragma unroll
            for (int i = 0; i < 16; ++i) split_f16x2(f[2 * i], f[2 * i + 1], hi[i], lo[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
              hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
              const float r0 = f[2 * i] - __uint_as_float(hi[i] << 16), r1 = f[2 * i + 1] - __uint_as_float(hi[i] & 0xffff0000u);
              const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0, r1);
              lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
            }
          }
#pragma unroll
          for (int m4 = 0; m4 < 4; ++m4) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((m4 ^ ph) << 4)), "r"(hi[4 * m4]), "r"(hi[4 * m4 + 1]), "r"(hi[4 * m4 + 2]), "r"(hi[4 * m4 + 3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((m4 + 4) ^ ph) << 4)), "r"(lo[4 * m4]), "r"(lo[4 * m4 + 1]), "r"(lo[4 * m4 + 2]), "r"(lo[4 * m4 + 3]) : "memory");
          }
        }
        fence_proxy_async_smem();     // (see conv_rs.cu: a plain remote arrive is sufficient after this fence)
        __syncwarp();
        if (lane == 0) { if (CG == 2) mbar_arrive_cta0(a_ready(a_st)); else mbar_arrive(a_ready(a_st)); }
        if (++a_st == p.a_stages) { a_st = 0; a_par ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ================= epilogue: vertical 4-tap blur of finished tx rows, noise / bias / activation, store =================
    const int q = warp - 4;
    const int r = q * 32 + lane;                 // accumulator lane == input column x0 + r
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t sbuf0 = st_base + (uint32_t)q * 16384u;
    const float nw = (p.noise && p.noise_w) ? *p.noise_w : 0.f;
    const int Ho = 2 * p.H, Wo = 2 * p.W;
    uint32_t n_store = 0;
    uint32_t full_par = 0;
    auto wait_full = [&](int s) {
      RU_TWAIT(0, mbar_wait(row_full(s), (full_par >> s) & 1u, 37));
      full_par ^= 1u << s;
    };
    // read 32 columns (one pixel phase) of tx row p, including the mirror of slot 0
    auto ld_row = [&](int pp, int half, float* v) {
      const int s = ru_slot(pp);
      tmem_ld_32x32(t_lane + (uint32_t)(s * RU_SLOT + half * 32), v);
      if (s == 0) {
        float m[32];
        tmem_ld_32x32(t_lane + (uint32_t)(RU_S * RU_SLOT + half * 32), m);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += m[i];
      }
    };
    auto free_row = [&](int pp) {    // zero the row's slot (and mirror) and hand it back to the MMA issuer
      const int s = ru_slot(pp);
      ru_zero32(t_lane + (uint32_t)(s * RU_SLOT)); ru_zero32(t_lane + (uint32_t)(s * RU_SLOT + 32));
      if (s == 0) { ru_zero32(t_lane + (uint32_t)(RU_S * RU_SLOT)); ru_zero32(t_lane + (uint32_t)(RU_S * RU_SLOT + 32)); }
      ru_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (CG == 2) mbar_arrive_cta0(row_empty(s)); else mbar_arrive(row_empty(s)); }
    };
    auto free_rows2 = [&](int pp) {   // rows pp and pp + 1: zero both slots (and the mirror), one store wait, two arrives
      const int sa = ru_slot(pp), sb = ru_slot(pp + 1);
      ru_zero32(t_lane + (uint32_t)(sa * RU_SLOT)); ru_zero32(t_lane + (uint32_t)(sa * RU_SLOT + 32));
      ru_zero32(t_lane + (uint32_t)(sb * RU_SLOT)); ru_zero32(t_lane + (uint32_t)(sb * RU_SLOT + 32));
      if (sa == 0 || sb == 0) { ru_zero32(t_lane + (uint32_t)(RU_S * RU_SLOT)); ru_zero32(t_lane + (uint32_t)(RU_S * RU_SLOT + 32)); }
      ru_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) { mbar_arrive_cta0(row_empty(sa)); mbar_arrive_cta0(row_empty(sb)); }
        else { mbar_arrive(row_empty(sa)); mbar_arrive(row_empty(sb)); }
      }
    };
    // noise / bias / activation on one output row's 32 channels of one pixel, then into its 128-byte row of the staging tile
    auto finish_row = [&](float* acc, float nzv, uint32_t row, int rr) {
      const float4* bp = reinterpret_cast<const float4*>(cst);
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 bq = bp[c4];
        acc[4 * c4 + 0] = fmaf(acc[4 * c4 + 0], p.acc_scale, bq.x + nzv); acc[4 * c4 + 1] = fmaf(acc[4 * c4 + 1], p.acc_scale, bq.y + nzv);
        acc[4 * c4 + 2] = fmaf(acc[4 * c4 + 2], p.acc_scale, bq.z + nzv); acc[4 * c4 + 3] = fmaf(acc[4 * c4 + 3], p.acc_scale, bq.w + nzv);
      }
      if (p.act == VT_ACT_LRELU) {
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = vt_lrelu(acc[c], p.slope) * p.gain;
      }
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t dst = row + (uint32_t)((kk ^ (rr & 7)) << 4);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(acc[4 * kk]), "f"(acc[4 * kk + 1]), "f"(acc[4 * kk + 2]), "f"(acc[4 * kk + 3]) : "memory");
      }
    };
    for (int strip = cta_i; strip < p.total_strips; strip += cta_n) {
      int b, i0, x0, R;
      strip_geom(strip, b, i0, x0, R);
      const int X0 = 2 * (x0 + r);                       // this thread's two output columns X0, X0 + 1
      const bool x_in = (x0 + r) < p.W;
      float nz_next[4] = {0.f, 0.f, 0.f, 0.f};          // noise of the next iteration's two output rows x two columns
      auto prefetch = [&](int Y) {
        if (p.noise && x_in) {
          const float* np_ = p.noise + ((int64_t)b * Ho + Y) * Wo + X0;
          nz_next[0] = nw * __ldg(np_); nz_next[1] = nw * __ldg(np_ + 1);
          nz_next[2] = nw * __ldg(np_ + Wo); nz_next[3] = nw * __ldg(np_ + Wo + 1);
        }
      };
      prefetch(2 * i0);
      for (int k = 0; k < R + 2; ++k) {
        const int i = i0 - 1 + k;
        wait_full(ru_slot(2 * i));
        wait_full(ru_slot(2 * i + 1));
        tc_fence_after();
        if (k == 1) free_row(2 * i0 - 2);                // the strip's leading row: only ever an incomplete sum
        if (k < 2) continue;
        float nz[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) nz[j] = nz_next[j];
        if (EPI == 1) {
          // The MMA issuer's next input row needs exactly the slots of tx rows 2i-3 and 2i-2: everything before free_rows2() is on
          // the critical path of the whole pipeline, everything after it overlaps the next row's MMAs.
          const int Yb = 2 * i - 2;                      // output rows Yb, Yb + 1 <- tx rows Yb-1 .. Yb+3
          if (lane == 0) tma_store_wait_read<0>();       // both staging buffers (stores of the previous step) are free
          __syncwarp();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            // Five tx rows j = 0..4 (row Yb-1+j) with weights (g[j], g[j-1]) for the two output rows; the logical slot 0 is the
            // sum of physical slot 0 and the mirror slot, so the mirror is a sixth entry with its row's weights.  Two register
            // buffers: a load is issued into a buffer right after its previous content has been consumed.
            uint32_t u0[32], u1[32];
            float a0[32], a1[32];
            const uint32_t tcol = t_lane + (uint32_t)(half * 32);
            const uint32_t t_mir = tcol + (uint32_t)(RU_S * RU_SLOT);
            const int s0 = ru_slot(Yb - 1);              // slots of rows j: (s0 + j) % 6
            const int jm = (6 - s0) % 6;                 // row whose slot is 0 (5: none of the five)
            auto taddr = [&](int j) { int sj = s0 + j; if (sj >= RU_S) sj -= RU_S; return tcol + (uint32_t)(sj * RU_SLOT); };
            tmem_ld_32x32_issue(taddr(0), u0);
            tmem_ld_32x32_issue(taddr(1), u1);
            if (jm <= 1) {                               // rows 0 / 1 carry the mirror: third load in flight (a0 / a1 are not live yet)
              uint32_t m[32];
              tmem_ld_32x32_issue(t_mir, m);
              tmem_ld_wait();
              tmem_ld_pin(u0); tmem_ld_pin(u1); tmem_ld_pin(m);
              const float ma = jm == 0 ? p.g[0] : p.g[1], mb = jm == 0 ? 0.f : p.g[0];
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                a0[c] = fmaf(p.g[0], __uint_as_float(u0[c]), 0.f);
                a0[c] = fmaf(p.g[1], __uint_as_float(u1[c]), a0[c]);
                a0[c] = fmaf(ma, __uint_as_float(m[c]), a0[c]);
                a1[c] = fmaf(p.g[0], __uint_as_float(u1[c]), 0.f);
                a1[c] = fmaf(mb, __uint_as_float(m[c]), a1[c]);
              }
            } else {
              tmem_ld_wait();
              tmem_ld_pin(u0); tmem_ld_pin(u1);
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                a0[c] = fmaf(p.g[0], __uint_as_float(u0[c]), 0.f);
                a0[c] = fmaf(p.g[1], __uint_as_float(u1[c]), a0[c]);
                a1[c] = fmaf(p.g[0], __uint_as_float(u1[c]), 0.f);
              }
            }
            if (half == 1) free_rows2(2 * i - 3);        // rows 2i-3, 2i-2 (and their mirror) have been read for the last time
            tmem_ld_32x32_issue(taddr(2), u0);
            tmem_ld_32x32_issue(taddr(3), u1);
            tmem_ld_wait();
            tmem_ld_pin(u0); tmem_ld_pin(u1);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              a0[c] = fmaf(p.g[2], __uint_as_float(u0[c]), a0[c]);
              a1[c] = fmaf(p.g[1], __uint_as_float(u0[c]), a1[c]);
            }
            // next into u0: the mirror of row 2 / 3 if one of them is slot 0, else row 4
            const bool mir23 = (jm == 2 || jm == 3);
            tmem_ld_32x32_issue(mir23 ? t_mir : taddr(4), u0);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              a0[c] = fmaf(p.g[3], __uint_as_float(u1[c]), a0[c]);
              a1[c] = fmaf(p.g[2], __uint_as_float(u1[c]), a1[c]);
            }
            // next into u1: row 4 (after the mirror of 2 / 3), or row 4's own mirror
            const bool second = mir23 || jm == 4;
            if (second) tmem_ld_32x32_issue(mir23 ? taddr(4) : t_mir, u1);
            tmem_ld_wait();
            tmem_ld_pin(u0);
            if (mir23) {
              const float ma = jm == 2 ? p.g[2] : p.g[3], mb = jm == 2 ? p.g[1] : p.g[2];
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                a0[c] = fmaf(ma, __uint_as_float(u0[c]), a0[c]);
                a1[c] = fmaf(mb, __uint_as_float(u0[c]), a1[c]);
              }
            } else {
#pragma unroll
              for (int c = 0; c < 32; ++c) a1[c] = fmaf(p.g[3], __uint_as_float(u0[c]), a1[c]);
            }
            if (second) {                                // row 4 or its mirror: weight g[3] on the second output row
              tmem_ld_pin(u1);
#pragma unroll
              for (int c = 0; c < 32; ++c) a1[c] = fmaf(p.g[3], __uint_as_float(u1[c]), a1[c]);
            }
            const int rr = 2 * lane + half;              // output pixel inside this warp's 64-pixel run
            finish_row(a0, nz[half], sbuf0 + (uint32_t)rr * 128u, rr);
            finish_row(a1, nz[2 + half], sbuf0 + 8192u + (uint32_t)rr * 128u, rr);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&p.out_map, sbuf0, p.c_base, 2 * (x0 + q * 32), Yb, b);
            tma_store_4d(&p.out_map, sbuf0 + 8192u, p.c_base, 2 * (x0 + q * 32), Yb + 1, b);
            tma_store_commit();
          }
        } else {
#pragma unroll
        for (int yi = 0; yi < 2; ++yi) {
          const int Y = 2 * i - 2 + yi;
          const uint32_t sbuf = sbuf0 + (n_store & 1u) * 8192u;
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float acc[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[c] = 0.f;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              float v[32];
              ld_row(Y + m - 1, half, v);
#pragma unroll
              for (int c = 0; c < 32; ++c) acc[c] = fmaf(p.g[m], v[c], acc[c]);
              // tx rows 2i-3 and 2i-2 have now been read for the last time (2i-3 by row Y = 2i-2, 2i-2 by both rows)
              if (yi == 1 && half == 1 && m == 0) { free_row(2 * i - 3); free_row(2 * i - 2); }
            }
            const float4* bp = reinterpret_cast<const float4*>(cst);
            const float nzv = nz[yi * 2 + half];
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              const float4 bq = bp[c4];
              acc[4 * c4 + 0] = fmaf(acc[4 * c4 + 0], p.acc_scale, bq.x + nzv); acc[4 * c4 + 1] = fmaf(acc[4 * c4 + 1], p.acc_scale, bq.y + nzv);
              acc[4 * c4 + 2] = fmaf(acc[4 * c4 + 2], p.acc_scale, bq.z + nzv); acc[4 * c4 + 3] = fmaf(acc[4 * c4 + 3], p.acc_scale, bq.w + nzv);
            }
            if (p.act == VT_ACT_LRELU) {
#pragma unroll
              for (int c = 0; c < 32; ++c) acc[c] = vt_lrelu(acc[c], p.slope) * p.gain;
            }
            const int rr = 2 * lane + half;              // output pixel inside this warp's 64-pixel run
            const uint32_t row = sbuf + (uint32_t)rr * 128u;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint32_t dst = row + (uint32_t)((kk ^ (rr & 7)) << 4);
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(acc[4 * kk]), "f"(acc[4 * kk + 1]), "f"(acc[4 * kk + 2]), "f"(acc[4 * kk + 3]) : "memory");
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&p.out_map, sbuf, p.c_base, 2 * (x0 + q * 32), Y, b);
            tma_store_commit();
          }
          ++n_store;
        }
        }
        if (k + 1 < R + 2) prefetch(2 * i);              // next iteration's rows 2(i+1)-2, 2(i+1)-1
      }
      // the strip's last rows: 2i-1, 2i, 2i+1 (i = i0+R) were inputs of the final output rows; 2i+2 is the trailing incomplete row
      const int il = i0 + R;
      wait_full(ru_slot(2 * il + 2));
      tc_fence_after();
      free_row(2 * il - 1); free_row(2 * il); free_row(2 * il + 1); free_row(2 * il + 2);
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  if (p.dbg && lane == 0 && (warp == 0 || warp == 1 || warp == 4)) {
    const int role = warp == 4 ? 2 : warp;
    unsigned long long* o = p.dbg + (size_t)blockIdx.x * 16 + role * 5;
    o[0] = (unsigned long long)(clock64() - t_begin);
    o[1] = (unsigned long long)tw[0]; o[2] = (unsigned long long)tw[1]; o[3] = (unsigned long long)tw[2]; o[4] = (unsigned long long)tw[3];
  }
  if (p.dbg && lane == 0 && warp == 2) p.dbg[(size_t)blockIdx.x * 16 + 15] = (unsigned long long)tw[0];
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2sm(tmem_base, 512u); else tmem_dealloc(tmem_base, 512u);
  }
}

// ---- weight fold: horizontal half of the blur into the transposed-conv taps ---------------------------------------------------
// w9: [wB][9 = ky*3+kx][Cout][Cin] (un-flipped conv_transpose taps, modulated); g: flipped 1-D blur taps [4];
// out: [wB][Cout/32][Cin/32][dj 3][ (ky*2 + px)*32 + co ][32 ch]
__global__ void __launch_bounds__(256)
fold_upconv_x_kernel(const float* __restrict__ w9, float g0, float g1, float g2, float g3, float* __restrict__ out, int Cout, int Cin, int64_t total) {
  const float g[4] = {g0, g1, g2, g3};
  const int KC = Cin / 32, NH = Cout / 32;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = idx;
    const int c = (int)(t % 32); t /= 32;
    const int co = (int)(t % 32); t /= 32;
    const int px = (int)(t % 2); t /= 2;
    const int ky = (int)(t % 3); t /= 3;
    const int dj = (int)(t % 3) - 1; t /= 3;
    const int kc = (int)(t % KC); t /= KC;
    const int h = (int)(t % NH); t /= NH;
    const int b = (int)t;
    float acc = 0.f;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int n = 2 * dj + kx - px + 1;
      if (n >= 0 && n <= 3) acc = fmaf(w9[(((int64_t)b * 9 + ky * 3 + kx) * Cout + h * 32 + co) * Cin + kc * 32 + c], g[n], acc);
    }
    out[idx] = acc;
  }
}

int g_rsu_cg = 0, g_rsu_rows = 0;
int g_rsu_bstages = 6;   // weight-tile ring depth: the MMA issuer consumes a tile per ~1 k cycles, 4 stages left it waiting for weights
                         // 5-11 % of the time (role timing); 6 stages = all tiles of a row at Cin = 64
int g_rsu_epi = 1;   // epilogue form (see the kernel template): 1 = both output rows per pass, dying rows first

}  // namespace

int vt_rsu_set_option(const char* key, int value, int* old) {
  if (key && strcmp(key, "rsu_cg") == 0) { *old = g_rsu_cg; g_rsu_cg = value; return 1; }
  if (key && strcmp(key, "rsu_rows") == 0) { *old = g_rsu_rows; g_rsu_rows = value; return 1; }
  if (key && strcmp(key, "rsu_bstages") == 0) { *old = g_rsu_bstages; g_rsu_bstages = value; return 1; }
  if (key && strcmp(key, "rsu_epi") == 0) { *old = g_rsu_epi; g_rsu_epi = value; return 1; }
  return 0;
}

extern "C" int vt_fold_upconv_x_weights_f32(const float* w9, const float* g_host4, float* out, int wB, int Cout, int Cin, void* stream) {
  VT_CHECK(w9 && g_host4 && out && wB >= 1 && Cout % 32 == 0 && Cin % 32 == 0 && Cout >= 32 && Cin >= 32, "fold_upconv_x_weights: bad args");
  const int64_t total = (int64_t)wB * (Cout / 32) * (Cin / 32) * 3 * 3 * 2 * 32 * 32;
  int64_t blocks = vt_cdiv(total, 256);
  if (blocks > (int64_t)vt_num_sms() * 16) blocks = (int64_t)vt_num_sms() * 16;
  fold_upconv_x_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(w9, g_host4[0], g_host4[1], g_host4[2], g_host4[3], out, Cout, Cin, total);
  VT_LAUNCH_CHECK();
  return 0;
}

/* in: NHWC [B,H,W,Cin] (Cin in {64,128} or any multiple of 32), weights: split (vt_split_weights_bf16x3 / _f16x3) layout of
 * vt_fold_upconv_x_weights_f32, out: NHWC [B,2H,2W,Cout] */
extern "C" int vt_conv_up2_rs(const float* in, const void* w_split, float* out, int B, int H, int W, int Cin, int Cout, int wB,
                              const float* g_host4, const float* bias, const float* noise, const float* noise_w, int act,
                              float slope, float gain, int fmt, float acc_scale, void* stream) {
  VT_CHECK(in && w_split && out && g_host4 && B >= 1 && H >= 1 && W >= 1, "conv_up2_rs: bad args");
  VT_CHECK(Cin % 32 == 0 && Cin >= 32 && Cin <= 256 && Cout % 32 == 0 && Cout >= 32 && Cout <= 128, "conv_up2_rs: Cin in [32,256], Cout in [32,128], multiples of 32");
  VT_CHECK(wB == 1 || wB == B, "conv_up2_rs: wB must be 1 or B");
  VT_CHECK(act == VT_ACT_NONE || act == VT_ACT_LRELU, "conv_up2_rs: activation must be none or leaky-relu");
  VT_CHECK(!noise || noise_w, "conv_up2_rs: noise without noise_w");
  VT_CHECK((((uintptr_t)in | (uintptr_t)w_split | (uintptr_t)out) & 15) == 0, "conv_up2_rs: pointers must be 16-byte aligned");
  VT_CHECK(acc_scale > 0.f && (fmt == 0 || fmt == 1), "conv_up2_rs: bad fmt / acc_scale");
  static thread_local RuArgs a;
  memset(&a, 0, sizeof(a));
  const int cg = g_rsu_cg ? g_rsu_cg : (W >= 2 * RU_PX ? 2 : 1);
  VT_CHECK(cg == 1 || cg == 2, "conv_up2_rs: rsu_cg must be 0, 1 or 2");
  a.B = B; a.H = H; a.W = W; a.KC = Cin / 32; a.wB = wB;
  a.b_tile_bytes = (RU_N / cg) * 128;
  a.bias = bias; a.noise = noise; a.noise_w = noise_w; a.act = act; a.slope = slope; a.gain = gain;
  for (int i = 0; i < 4; ++i) a.g[i] = g_host4[i];
  a.fmt = fmt; a.acc_scale = acc_scale;
  a.dbg = g_tc_dbg_export;
  a.strips_x = (int)vt_cdiv(W, RU_PX * cg);
  const int units = vt_num_sms() / cg;
  int best_rows = 0; double best_cost = 1e30;
  for (int rows = 8; rows <= 256; rows += 4) {
    const int64_t strips = (int64_t)B * a.strips_x * vt_cdiv(H, rows);
    const double cost = (double)vt_cdiv(strips, units) * (rows + 2 + 3);
    if (cost < best_cost - 1e-9) { best_cost = cost; best_rows = rows; }
  }
  a.rows_per_strip = g_rsu_rows > 0 ? g_rsu_rows : best_rows;
  if (a.rows_per_strip > H) a.rows_per_strip = H;
  a.strips_y = (int)vt_cdiv(H, a.rows_per_strip);
  const int64_t total = (int64_t)B * a.strips_x * a.strips_y;
  VT_CHECK(total < (1LL << 30), "conv_up2_rs: too many strips");
  a.total_strips = (int)total;
  a.b_stages = g_rsu_bstages;
  VT_CHECK(a.b_stages >= 2 && a.b_stages <= RU_B_STAGES_MAX, "conv_up2_rs: rsu_bstages must be in [2, 8]");
  const int fixed0 = 1024 + RU_STAGING + 128 + 512 + 1024;
  // keep at least three activation stages (single-CTA launches hold whole 24 KB weight tiles: 4 of them)
  while (a.b_stages > 2 && (RU_MAX_SMEM - fixed0 - a.b_stages * a.b_tile_bytes) / RU_A_STAGE < 3) --a.b_stages;
  const int fixed = a.b_stages * a.b_tile_bytes + fixed0;
  a.a_stages = (RU_MAX_SMEM - fixed) / RU_A_STAGE;
  if (a.a_stages > 6) a.a_stages = 6;
  VT_CHECK(a.a_stages >= 2, "conv_up2_rs: shared memory plan does not fit");
  const int smem_bytes = a.a_stages * RU_A_STAGE + fixed;
  {
    const uint64_t cs = (uint64_t)Cin;
    const uint64_t dims[4] = {cs, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {cs * 4, (uint64_t)W * cs * 4, (uint64_t)H * W * cs * 4};
    const uint32_t box[4] = {32, RU_BOX, 1, 1};
    if (vt_tc_make_map4(&a.in_map, in, dims, str, box, "rsu input", false)) return 1;
  }
  {
    const uint64_t tiles = (uint64_t)(Cout / 32) * a.KC * 3;
    const uint64_t dims[4] = {64, RU_N, tiles, (uint64_t)wB};
    const uint64_t str[3] = {128, (uint64_t)RU_N * 128, tiles * RU_N * 128};
    const uint32_t box[4] = {64, (uint32_t)(RU_N / cg), 1, 1};
    if (vt_tc_make_map4(&a.w_map, w_split, dims, str, box, "rsu weight", true)) return 1;
  }
  {
    const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)(2 * W), (uint64_t)(2 * H), (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)Cout * 4, (uint64_t)(2 * W) * Cout * 4, (uint64_t)(2 * H) * (2 * W) * Cout * 4};
    const uint32_t box[4] = {32, 64, 1, 1};
    if (vt_tc_make_map4(&a.out_map, out, dims, str, box, "rsu output", false)) return 1;
  }
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(conv_rsu_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, RU_MAX_SMEM);
    if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(conv_rsu_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, RU_MAX_SMEM);
    if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(conv_rsu_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, RU_MAX_SMEM);
    if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(conv_rsu_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, RU_MAX_SMEM);
  });
  VT_CHECK(attr_err == cudaSuccess, "conv_up2_rs: cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
  for (int half = 0; half < Cout / 32; ++half) {
    a.half = half; a.c_base = half * 32;
    if (cg == 1) {
      int grid = vt_num_sms();
      if (grid > a.total_strips) grid = a.total_strips;
      if (g_rsu_epi) conv_rsu_kernel<1, 1><<<grid, RU_THREADS, smem_bytes, (cudaStream_t)stream>>>(a);
      else conv_rsu_kernel<1, 0><<<grid, RU_THREADS, smem_bytes, (cudaStream_t)stream>>>(a);
    } else {
      int pairs = vt_num_sms() / 2;
      if (pairs > a.total_strips) pairs = a.total_strips;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(2 * pairs));
      cfg.blockDim = dim3(RU_THREADS);
      cfg.dynamicSmemBytes = (size_t)smem_bytes;
      cfg.stream = (cudaStream_t)stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (g_rsu_epi) VT_CUDA(cudaLaunchKernelEx(&cfg, conv_rsu_kernel<2, 1>, a));
      else VT_CUDA(cudaLaunchKernelEx(&cfg, conv_rsu_kernel<2, 0>, a));
    }
    VT_LAUNCH_CHECK();
  }
  return 0;
}
