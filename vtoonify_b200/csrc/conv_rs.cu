// conv_rs.cu — "row-strip" 3x3 stride-1 convolution for the small-channel, full-resolution tail of the generator
// (StyledConv conv2 of the last levels: 32->32 @ 4H x 4W, 64->64 @ 2H x 2W; model/stylegan/model.py:298-304, 364-392).
//
// Why a second tensor-core kernel.  With N = Cout = 32 the tap-by-tap implicit GEMM of conv_tc.cu issues 54 MMAs of N = 32
// per 256 pixels and every one of them re-reads its 128 x 32 B pixel operand from shared memory: the instruction cost is
// ~45 + 0.62 N cycles, i.e. the tensor pipe idles behind a fixed per-instruction cost (measured: MMA warp busy 98 %, tensor pipe
// 22 %).  Here the three vertical taps are STACKED ALONG N instead:
//
//   one MMA:  D[128 px of input row r, (dy, co)] += X[row r, px + dx, ci] * W[(dy, co), ci]          N = 3 * Cout
//
// so an input row contributes, in one instruction, to the three output rows r-1, r, r+1 (dy = +1, 0, -1) and 18 MMAs of N = 96
// replace 54 of N = 32.  The three partial sums of an output row come from three different input rows; they are added by the
// tensor core itself: the accumulators form a RING of output-row slots in TMEM (Cout columns each), an input row's MMA window
// covers three consecutive slots, every MMA accumulates (never overwrites), and the epilogue zeroes a slot after it has read
// the finished row.  Two "mirror" slots behind the ring make every window contiguous (slots S, S+1 alias slots 0, 1; the
// epilogue adds the alias when it reads rows 0 and 1 of a lap).
//
// Geometry: a CTA owns a 128-pixel wide column strip (cta_group::2: the pair owns 256 pixels, one MMA covers both), walks down
// `rows_per_strip` rows and streams each input row once: TMA box (32 ch, 130 px, 1 row) with a 1-pixel halo, zero-filled outside
// the image (= the conv's zero padding); dx = -1, 0, +1 are descriptor start addresses shifted by whole 128-byte rows.  The
// weights of all 9 taps stay resident in shared memory for the whole launch (36 KB for 32->32).
// Roles (14 warps): 0 TMA producer | 1 MMA issuer | 2,3,12,13 operand transform (fp32 -> [hi|lo] 16-bit split, in place) |
// 4-7 and 8-11: two epilogue groups that take finished rows alternately (TMEM -> bias / noise / leaky-relu / fused ToRGB + skip
// up-sampling -> per-warp swizzled staging -> per-warp TMA store); the epilogue is a chain of latencies (barrier, TMEM, shared
// memory constants, proxy fence), two rows in flight hide them.
#include "tc_common.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <mutex>

using namespace vt_tc;

int vt_validate_conv_desc(const vt_conv_desc* d, const char* who);

namespace {

constexpr int RS_PX = 128;                 // pixels per CTA row segment = MMA M per CTA
constexpr int RS_BOX = RS_PX + 2;          // with the 1-pixel halo on both sides
constexpr int RS_A_STAGE = 17 * 1024;      // 130 rows x 128 B, padded to the 1024-byte swizzle atom
constexpr int RS_THREADS = 448;
constexpr int RS_XFORM_WARPS = 4;
constexpr int RS_MAX_SMEM = 227 * 1024;
constexpr int RS_STAGING = 8 * 2 * 4096;   // per epilogue warp: two 32 px x 128 B buffers
constexpr int RS_CONST_FLOATS = 64 + 3 * 64 + 4 + 16;   // bias | rgb_w[3][64] | rgb_bias (3, padded) | skip kernel

struct RsArgs {
  CUtensorMap in_map;      // fp32 (cstride, W, H, B), box (32, 130, 1, 1)
  CUtensorMap w_map;       // 16-bit (64, 3*Cout, KC*3, wB), box (64, 3*Cout/CG, 1, 1): rows [hi(32) | lo(32)] per 32-channel chunk
  CUtensorMap out_map;     // fp32 (Cout, W, H, B), box (32, 32, 1, 1)
  int B, H, W, Cout, KC, wB;
  int rows_per_strip, strips_x, strips_y, total_strips;
  int S;                   // logical output-row slots of the TMEM ring (physical: S + 2)
  int a_stages;
  int w_tile_bytes;        // one (kc, dx) weight tile held by this CTA: (3*Cout/CG) rows x 128 B
  const float* bias; const float* noise; const float* noise_w;
  int act; float slope, gain;
  const float* rgb_w; const float* rgb_bias; const float* rgb_skip; const float* rgb_skip_kernel; float* rgb_out;
  int fmt;                 // operand split: 0 = bf16 hi/lo, 1 = fp16 hi/lo
  float acc_scale;         // accumulators are multiplied by this before the bias (undoes a power-of-two weight scale)
  int strict_release;      // 1: cluster-scope release on the transform warps' remote arrive (A/B switch for tests)
  int store_out;           // 0: only the fused ToRGB image is produced (last layer of the generator: the activation has no reader)
  unsigned long long* dbg;
};

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) { tmem_ld_32x32(taddr, v); }

__device__ __forceinline__ void tmem_zero32(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(z)
      : "memory");
}
// this warp's 32 lanes x 16 consecutive columns, added into acc[0..15]
__device__ __forceinline__ void tmem_ld16_add(uint32_t taddr, float* acc) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] += __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }


#define RS_TWAIT(slot, stmt) do { if (p.dbg) { const long long t__ = clock64(); stmt; tw[slot] += clock64() - t__; } else { stmt; } } while (0)

template <int CG>
__global__ void __launch_bounds__(RS_THREADS, 1)
conv_rs_kernel(const __grid_constant__ RsArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t w_base = a_base + (uint32_t)p.a_stages * RS_A_STAGE;
  const uint32_t w_bytes = (uint32_t)(p.KC * 3) * (uint32_t)p.w_tile_bytes;
  const uint32_t st_base = (w_base + w_bytes + 1023u) & ~1023u;
  const uint32_t cst_base = st_base + RS_STAGING;
  const uint32_t bar_base = cst_base + RS_CONST_FLOATS * 4;
  // barriers (8 bytes each): a_full[8] a_ready[8] a_empty[8] row_full[16] row_empty[16] w_full w_empty | tmem slot
  auto a_full = [&](int i) { return bar_base + 8u * i; };
  auto a_ready = [&](int i) { return bar_base + 64u + 8u * i; };
  auto a_empty = [&](int i) { return bar_base + 128u + 8u * i; };
  auto row_full = [&](int i) { return bar_base + 192u + 8u * i; };
  auto row_empty = [&](int i) { return bar_base + 320u + 8u * i; };
  const uint32_t w_full = bar_base + 448u, w_empty = bar_base + 456u, tmem_slot = bar_base + 464u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  float* cst = reinterpret_cast<float*>(smem_gen + (cst_base - smem_base));   // [0,64) bias, [64,256) rgb_w, [256,260) rgb_bias, [260,276) kernel

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int cta_i = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int cta_n = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int S = p.S, Cout = p.Cout, N = 3 * p.Cout;
  long long tw[4] = {0, 0, 0, 0};
  const long long t_begin = clock64();
  const bool is_epi = warp >= 4 && warp < 12;
  const bool is_xform = warp == 2 || warp == 3 || warp >= 12;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.in_map); tma_prefetch_desc(&p.w_map);
    if (p.store_out) tma_prefetch_desc(&p.out_map);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.a_stages; ++i) { mbar_init(a_full(i), 1); mbar_init(a_ready(i), RS_XFORM_WARPS * CG); mbar_init(a_empty(i), 1); }
    for (int i = 0; i < 16; ++i) { mbar_init(row_full(i), 1); mbar_init(row_empty(i), 4 * CG); }
    mbar_init(w_full, 1); mbar_init(w_empty, 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 2) {
    if (CG == 2) tmem_alloc_2sm(tmem_slot, 512u); else tmem_alloc(tmem_slot, 512u);
    tc_fence_before();
  }
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  if (warp >= 4 && warp < 8) {
    // every accumulator slot starts at zero: MMAs only ever accumulate
    const uint32_t t0 = tmem_base + ((uint32_t)((warp - 4) * 32) << 16);
    for (int c = 0; c < 512; c += 32) tmem_zero32(t0 + (uint32_t)c);
    tmem_wait_st();
    tc_fence_before();
  }
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();

  // strip -> (b, ys, xs); a pair (CG == 2) owns 256 pixels, this CTA the half selected by its rank
  const int strips_per_img = p.strips_y * p.strips_x;
  auto strip_geom = [&](int strip, int& b, int& y0, int& x0, int& R) {
    b = strip / strips_per_img;
    const int rem = strip - b * strips_per_img;
    const int ys = rem / p.strips_x, xs = rem - ys * p.strips_x;
    y0 = ys * p.rows_per_strip;
    R = min(p.rows_per_strip, p.H - y0);
    x0 = xs * (RS_PX * CG) + (int)rank * RS_PX;
  };
  // A strip of R output rows y0 .. y0+R-1 streams the R+2 input rows y0-1 .. y0+R (index i).  Input row i adds into the output
  // rows k = i, i+1, i+2 where k counts from y0-2: rows k = 0, 1 and k = R+2, R+3 only ever hold incomplete sums (they belong to
  // the neighbouring strips) and are just drained; k = 2 .. R+1 are produced.  Output row y lives in slot (y + 2) mod S, i.e.
  // row k of the strip in slot (y0 + k) mod S.

  if (warp == 0) {
    // ================= TMA producer =================
    int a_st = 0; uint32_t a_par = 0, w_par = 0;
    int cur_wb = -1;
    for (int strip = cta_i; strip < p.total_strips; strip += cta_n) {
      int b, y0, x0, R;
      strip_geom(strip, b, y0, x0, R);
      const int wb = p.wB > 1 ? b : 0;
      if (wb != cur_wb) {
        if (cur_wb >= 0) { RS_TWAIT(1, mbar_wait(w_empty, w_par, 21)); w_par ^= 1; }   // every MMA that read the old weights is done
        if (elect_one()) {
          if (CG == 2) { if (rank == 0) mbar_arrive_expect_tx(w_full, 2u * w_bytes); }
          else mbar_arrive_expect_tx(w_full, w_bytes);
          for (int t = 0; t < p.KC * 3; ++t) {
            const uint32_t dst = w_base + (uint32_t)t * (uint32_t)p.w_tile_bytes;
            if (CG == 2) tma_load_4d_2sm(dst, &p.w_map, w_full, 0, (int)rank * (N / 2), t, wb);
            else tma_load_4d(dst, &p.w_map, w_full, 0, 0, t, wb);
          }
        }
        __syncwarp();
        cur_wb = wb;
      }
      for (int i = 0; i < R + 2; ++i) {
        const int y = y0 - 1 + i;
        for (int kc = 0; kc < p.KC; ++kc) {
          RS_TWAIT(0, mbar_wait(a_empty(a_st), a_par ^ 1, 22));
          if (elect_one()) {
            mbar_arrive_expect_tx(a_full(a_st), (uint32_t)(RS_BOX * 128));
            tma_load_4d(a_base + (uint32_t)a_st * RS_A_STAGE, &p.in_map, a_full(a_st), kc * 32, x0 - 1, y, b);
          }
          __syncwarp();
          if (++a_st == p.a_stages) { a_st = 0; a_par ^= 1; }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ================= MMA issuer (CTA rank 0 of the pair) =================
    const uint32_t idesc = p.fmt ? make_idesc_f16(RS_PX * CG, N) : make_idesc_bf16(RS_PX * CG, N);
    int a_st = 0; uint32_t a_par = 0, w_par = 0;
    int cur_wb = -1;
    uint32_t acq = 0;       // bit q: parity of the number of times slot q has been handed to a new output row
    auto acquire = [&](int q) {
      // the slot's previous output row has been read and zeroed by both CTAs' epilogues (first use: a fresh barrier passes)
      RS_TWAIT(2, mbar_wait(row_empty(q), ((acq >> q) & 1u) ^ 1u, 24));
      acq ^= 1u << q;
    };
    for (int strip = cta_i; strip < p.total_strips; strip += cta_n) {
      int b, y0, x0, R;
      strip_geom(strip, b, y0, x0, R);
      const int wb = p.wB > 1 ? b : 0;
      if (wb != cur_wb) { RS_TWAIT(1, mbar_wait(w_full, w_par, 23)); w_par ^= 1; cur_wb = wb; }
      int next_wb = wb;
      if (strip + cta_n < p.total_strips && p.wB > 1) next_wb = (strip + cta_n) / strips_per_img;
      int qs = y0 % S;      // slot of output row k = i (the first row of input i's window)
      for (int i = 0; i < R + 2; ++i) {
        int q1 = qs + 1; if (q1 >= S) q1 -= S;
        int q2 = q1 + 1; if (q2 >= S) q2 -= S;
        if (i == 0) { acquire(qs); acquire(q1); }
        acquire(q2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(qs * Cout);   // physical window qs .. qs+2 (slots S, S+1 mirror 0, 1)
        for (int kc = 0; kc < p.KC; ++kc) {
          RS_TWAIT(0, mbar_wait(a_ready(a_st), a_par, 25));
          tc_fence_after();
          const uint32_t a_addr = a_base + (uint32_t)a_st * RS_A_STAGE;
          if (elect_one()) {
#pragma unroll
            for (int dxi = 0; dxi < 3; ++dxi) {
              // dx = dxi - 1: box row (1 + dx) is the first pixel of the tap; rows are consecutive pixels, 8-row groups 1024 B apart
              const uint64_t adesc = make_smem_desc_sw128(a_addr + (uint32_t)dxi * 128u, 1024, 0);
              const uint64_t bdesc = make_smem_desc_sw128(w_base + (uint32_t)(kc * 3 + dxi) * (uint32_t)p.w_tile_bytes, 1024, 0);
              // a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo; rows are [hi(32) | lo(32)] 16-bit, +2 on a descriptor = +32 B = 16 elements of K
              if (CG == 2) {
                umma_bf16_2sm(d_tmem, adesc, bdesc, idesc, 1);
                umma_bf16_2sm(d_tmem, adesc + 2, bdesc + 2, idesc, 1);
                umma_bf16_2sm(d_tmem, adesc + 4, bdesc, idesc, 1);
                umma_bf16_2sm(d_tmem, adesc + 6, bdesc + 2, idesc, 1);
                umma_bf16_2sm(d_tmem, adesc, bdesc + 4, idesc, 1);
                umma_bf16_2sm(d_tmem, adesc + 2, bdesc + 6, idesc, 1);
              } else {
                umma_bf16(d_tmem, adesc, bdesc, idesc, 1);
                umma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1);
                umma_bf16(d_tmem, adesc + 4, bdesc, idesc, 1);
                umma_bf16(d_tmem, adesc + 6, bdesc + 2, idesc, 1);
                umma_bf16(d_tmem, adesc, bdesc + 4, idesc, 1);
                umma_bf16(d_tmem, adesc + 2, bdesc + 6, idesc, 1);
              }
            }
            if (CG == 2) umma_commit_2sm(a_empty(a_st)); else umma_commit(a_empty(a_st));
          }
          __syncwarp();
          if (++a_st == p.a_stages) { a_st = 0; a_par ^= 1; }
        }
        // output row k = i has received its last contribution; after the last input row so have the two trailing rows
        if (elect_one()) {
          if (CG == 2) umma_commit_2sm(row_full(qs)); else umma_commit(row_full(qs));
          if (i == R + 1) {
            if (CG == 2) { umma_commit_2sm(row_full(q1)); umma_commit_2sm(row_full(q2)); }
            else { umma_commit(row_full(q1)); umma_commit(row_full(q2)); }
          }
        }
        __syncwarp();
        qs = q1;
      }
      if (next_wb != wb) {
        if (elect_one()) { if (CG == 2) umma_commit_2sm(w_empty); else umma_commit(w_empty); }
        __syncwarp();
      }
    }
  } else if (is_xform) {
    // ================= operand transform: fp32 rows -> [hi(32) | lo(32)] 16-bit rows, in place =================
    const int t = (warp < 4 ? warp - 2 : warp - 10) * 32 + lane;   // 0 .. 127
    int a_st = 0; uint32_t a_par = 0;
    for (int strip = cta_i; strip < p.total_strips; strip += cta_n) {
      int b, y0, x0, R;
      strip_geom(strip, b, y0, x0, R);
      const int loads = (R + 2) * p.KC;
      for (int l = 0; l < loads; ++l) {
        RS_TWAIT(0, mbar_wait(a_full(a_st), a_par, 26));
        const uint32_t stage = a_base + (uint32_t)a_st * RS_A_STAGE;
        for (int r = t; r < RS_BOX; r += 32 * RS_XFORM_WARPS) {
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
        // Every lane has made its own rows visible to the async proxy (the fence waits until the stores are performed in this
        // SM's shared memory, which has a single copy), so a plain remote arrive is enough for the pair's issuing thread: a
        // cluster-scope release (MEMBAR.ALL.GPU) costs ~2k cycles per stage and made this warp role the bottleneck (measured).
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) { if (CG == 2) { if (p.strict_release) mbar_arrive_cta0_release(a_ready(a_st)); else mbar_arrive_cta0(a_ready(a_st)); } else mbar_arrive(a_ready(a_st)); }
        if (++a_st == p.a_stages) { a_st = 0; a_par ^= 1; }
      }
    }
  } else if (is_epi) {
    // ================= epilogue: two groups of four warps take finished rows alternately =================
    const int grp = (warp - 4) >> 2;             // 0: warps 4-7, 1: warps 8-11
    const int q = (warp - 4) & 3;                // TMEM lane quadrant (== warp % 4)
    const int r = q * 32 + lane;                 // accumulator lane == pixel of the row segment
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t sbuf0 = st_base + (uint32_t)(warp - 4) * 8192u;
    const float nw = (p.noise && p.noise_w) ? *p.noise_w : 0.f;
    const int nchunks = Cout / 32;
    const int64_t HW = (int64_t)p.H * p.W;
    const int hs = p.H >> 1, ws = p.W >> 1;
    const bool timed = p.dbg != nullptr && grp == 0;
    uint32_t n_store = 0;
    uint32_t full_par = 0;   // bit q: parity of the number of finished rows slot q has delivered so far
    uint32_t turn = 0;       // finished rows seen so far (both groups count all of them; a group works on those with its parity)
    int cur_wb = -1;
    for (int strip = cta_i; strip < p.total_strips; strip += cta_n) {
      int b, y0, x0, R;
      strip_geom(strip, b, y0, x0, R);
      const int wb = p.wB > 1 ? b : 0;
      const int x = x0 + r;
      const bool x_in = x < p.W;
      if (wb != cur_wb) {
        // epilogue constants of this sample -> shared memory (bias, ToRGB weights / bias, skip kernel)
        named_bar_sync(1, 256);                  // nobody still reads the previous sample's constants
        for (int i = grp * 128 + r; i < RS_CONST_FLOATS; i += 256) {
          float v = 0.f;
          if (i < 64) v = (p.bias && i < Cout) ? __ldg(p.bias + i) : 0.f;
          else if (i < 256) { const int c = (i - 64) / 64, n = (i - 64) % 64; v = (p.rgb_w && n < Cout) ? __ldg(p.rgb_w + ((int64_t)wb * 3 + c) * Cout + n) : 0.f; }
          else if (i < 260) v = (p.rgb_w && i < 259) ? __ldg(p.rgb_bias + (i - 256)) : 0.f;
          else v = p.rgb_skip ? __ldg(p.rgb_skip_kernel + (i - 260)) : 0.f;
          cst[i] = v;
        }
        named_bar_sync(1, 256);
        cur_wb = wb;
      }
      // software pipeline: the global loads of this group's NEXT produced row (noise, skip pixels) are issued one turn ahead,
      // right after the proxy fence of the current row (a fence would otherwise wait for them)
      float nz_next = 0.f, sk_next[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) sk_next[i] = 0.f;
      auto prefetch = [&](int y) {
        if (p.noise) nz_next = x_in ? nw * __ldg(p.noise + (int64_t)b * HW + (int64_t)y * p.W + x) : 0.f;
        if (p.rgb_skip && x_in) {
          const int ky0 = (y - 2) & 1, kx0 = (x - 2) & 1;
          const int iy0 = (y - 2 + ky0) >> 1, ix0 = (x - 2 + kx0) >> 1;
          const int cy0 = iy0 < 0 ? 0 : iy0, cy1 = (iy0 + 1) < hs ? iy0 + 1 : hs - 1;
          const int cx0 = ix0 < 0 ? 0 : ix0, cx1 = (ix0 + 1) < ws ? ix0 + 1 : ws - 1;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float* sp = p.rgb_skip + ((int64_t)b * 3 + c) * (int64_t)hs * ws;
            sk_next[4 * c + 0] = __ldg(sp + (int64_t)cy0 * ws + cx0); sk_next[4 * c + 1] = __ldg(sp + (int64_t)cy0 * ws + cx1);
            sk_next[4 * c + 2] = __ldg(sp + (int64_t)cy1 * ws + cx0); sk_next[4 * c + 3] = __ldg(sp + (int64_t)cy1 * ws + cx1);
          }
        }
      };
      {   // this group's first produced row of the strip: k = 2 or 3
        const int kf = 2 + (int)(((turn + 2u) & 1u) != (uint32_t)grp);
        if (kf <= R + 1) prefetch(y0 - 2 + kf);
      }
      int slot = y0 % S;     // slot of row k = 0
      for (int k = 0; k < R + 4; ++k, ++turn) {
        const int cur = slot;
        if (++slot == S) slot = 0;
        const uint32_t par = (full_par >> cur) & 1u;
        full_par ^= 1u << cur;                     // both groups track every row's barrier phase
        if ((turn & 1u) != (uint32_t)grp) continue;
        const int y = y0 - 2 + k;
        const bool valid = k >= 2 && k <= R + 1;
        float nz = 0.f, sk[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) sk[i] = 0.f;
        if (valid) {
          nz = nz_next;
#pragma unroll
          for (int i = 0; i < 12; ++i) sk[i] = sk_next[i];
        }
        if (timed) { const long long t__ = clock64(); mbar_wait(row_full(cur), par, 27); tw[0] += clock64() - t__; }
        else mbar_wait(row_full(cur), par, 27);
        const long long t_e0 = timed ? clock64() : 0;
        tc_fence_after();
        float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
        for (int c = 0; c < nchunks; ++c) {
          float v[32];
          const uint32_t col = (uint32_t)(cur * Cout + c * 32);
          tmem_ld32(t_lane + col, v);
          tmem_zero32(t_lane + col);
          if (cur < 2) {   // rows in slots 0 and 1 also collected partial sums in the mirror slots S, S+1
            const uint32_t mcol = (uint32_t)((S + cur) * Cout + c * 32);
            tmem_ld16_add(t_lane + mcol, v);
            tmem_ld16_add(t_lane + mcol + 16u, v + 16);
            tmem_zero32(t_lane + mcol);
          }
          if (c == nchunks - 1) {
            // the slot is clean again: hand it back to the MMA issuer before the math
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CG == 2) mbar_arrive_cta0(row_empty(cur)); else mbar_arrive(row_empty(cur)); }
          }
          if (!valid) continue;
          const long long t_e1 = timed ? clock64() : 0;
          if (timed) tw[1] += t_e1 - t_e0;
          const float4* bp = reinterpret_cast<const float4*>(cst + 32 * c);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bq = bp[i];
            v[4 * i + 0] = fmaf(v[4 * i + 0], p.acc_scale, bq.x + nz); v[4 * i + 1] = fmaf(v[4 * i + 1], p.acc_scale, bq.y + nz);
            v[4 * i + 2] = fmaf(v[4 * i + 2], p.acc_scale, bq.z + nz); v[4 * i + 3] = fmaf(v[4 * i + 3], p.acc_scale, bq.w + nz);
          }
          if (p.act == VT_ACT_LRELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = vt_lrelu(v[i], p.slope) * p.gain;
          }
          if (p.rgb_w) {
            const float4* w0 = reinterpret_cast<const float4*>(cst + 64 + 32 * c);
            const float4* w1 = reinterpret_cast<const float4*>(cst + 128 + 32 * c);
            const float4* w2 = reinterpret_cast<const float4*>(cst + 192 + 32 * c);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 a0 = w0[i], a1 = w1[i], a2 = w2[i];
              rgb0 = fmaf(v[4 * i], a0.x, rgb0); rgb0 = fmaf(v[4 * i + 1], a0.y, rgb0); rgb0 = fmaf(v[4 * i + 2], a0.z, rgb0); rgb0 = fmaf(v[4 * i + 3], a0.w, rgb0);
              rgb1 = fmaf(v[4 * i], a1.x, rgb1); rgb1 = fmaf(v[4 * i + 1], a1.y, rgb1); rgb1 = fmaf(v[4 * i + 2], a1.z, rgb1); rgb1 = fmaf(v[4 * i + 3], a1.w, rgb1);
              rgb2 = fmaf(v[4 * i], a2.x, rgb2); rgb2 = fmaf(v[4 * i + 1], a2.y, rgb2); rgb2 = fmaf(v[4 * i + 2], a2.z, rgb2); rgb2 = fmaf(v[4 * i + 3], a2.w, rgb2);
            }
          }
          const long long t_e2 = timed ? clock64() : 0;
          if (timed) tw[2] += t_e2 - t_e1;
          if (!p.store_out) continue;
          // per-warp staging (32 pixels x 128 B, 128B-swizzled) + per-warp TMA store: no CTA-wide barrier in the epilogue
          const uint32_t sbuf = sbuf0 + (n_store & 1u) * 4096u;
          if (lane == 0) tma_store_wait_read<1>();   // the store that used this buffer two stores ago has read it
          __syncwarp();
          const uint32_t row = sbuf + (uint32_t)lane * 128u;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t dst = row + (uint32_t)((kk ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(v[4 * kk]), "f"(v[4 * kk + 1]), "f"(v[4 * kk + 2]), "f"(v[4 * kk + 3]) : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&p.out_map, sbuf, c * 32, x0 + q * 32, y, b);
            tma_store_commit();
          }
          ++n_store;
          if (timed) tw[3] += clock64() - t_e2;
        }
        if (!valid) continue;
        if (k + 2 <= R + 1) prefetch(y + 2);     // this group's next row; its loads have a whole turn to land
        if (p.rgb_w && x_in) {
          // + bias + Upsample(skip): upfirdn2d(up=2, pad=(2,1), 4x4) touches 2x2 skip pixels per output pixel (model.py:385-391)
          float o3[3] = {rgb0 + cst[256], rgb1 + cst[257], rgb2 + cst[258]};
          if (p.rgb_skip) {
            const int ky0 = (y - 2) & 1, kx0 = (x - 2) & 1;
            const int iy0 = (y - 2 + ky0) >> 1, ix0 = (x - 2 + kx0) >> 1;
            const float my0 = iy0 >= 0 ? 1.f : 0.f, my1 = (iy0 + 1) < hs ? 1.f : 0.f;
            const float mx0 = ix0 >= 0 ? 1.f : 0.f, mx1 = (ix0 + 1) < ws ? 1.f : 0.f;
            const float* kk = cst + 260;
            const float w00 = kk[(3 - ky0) * 4 + (3 - kx0)] * my0 * mx0;
            const float w01 = kk[(3 - ky0) * 4 + (1 - kx0)] * my0 * mx1;
            const float w10 = kk[(1 - ky0) * 4 + (3 - kx0)] * my1 * mx0;
            const float w11 = kk[(1 - ky0) * 4 + (1 - kx0)] * my1 * mx1;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              float u = sk[4 * c] * w00;                 // same accumulation order as the reference loop (ky outer, kx inner)
              u = fmaf(sk[4 * c + 1], w01, u);
              u = fmaf(sk[4 * c + 2], w10, u);
              u = fmaf(sk[4 * c + 3], w11, u);
              o3[c] += u;
            }
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) p.rgb_out[((int64_t)b * 3 + c) * HW + (int64_t)y * p.W + x] = o3[c];
        }
      }
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

int g_rs_cg = 0;          // 0: automatic (pairs when the image is at least 256 pixels wide); 1 / 2: forced (tests)
int g_rs_rows = 0;        // 0: automatic rows per strip; > 0: forced (tests)
int g_rs_strict = 0;      // 1: cluster-scope release arrive in the transform warps (A/B timing / paranoia switch)

int rs_check(const vt_conv_desc* d, bool set_err) {
#define RS_SUP(cond, ...) do { if (!(cond)) { if (set_err) vt_set_error(__VA_ARGS__); return 0; } } while (0)
  RS_SUP(d->n_src == 1 && d->stride == 1 && d->n_phase == 1 && d->taps == 9, "conv_rs: one source, stride 1, one phase, 9 taps");
  for (int t = 0; t < 9; ++t)
    RS_SUP(d->tap_dy[t] == t / 3 - 1 && d->tap_dx[t] == t % 3 - 1 && d->tap_w[t] == t, "conv_rs: taps must be the 3x3 / padding 1 pattern");
  RS_SUP(d->Ho == d->H && d->Wo == d->W, "conv_rs: output size must equal input size");
  RS_SUP((d->src_c[0] == 32 || d->src_c[0] == 64) && (d->Cout == 32 || d->Cout == 64), "conv_rs: Cin and Cout must be 32 or 64");
  RS_SUP(d->src_cstride[0] % 4 == 0 && d->src_cstride[0] >= d->src_c[0], "conv_rs: bad channel stride");
  RS_SUP(d->weight_bf16x3 != nullptr && (d->bf16x3_nstack == 2 || d->bf16x3_nstack == 3),
         "conv_rs: weight_bf16x3 must hold the row-strip layout (bf16x3_nstack = 2: bf16 split, 3: fp16 split)");
  RS_SUP(d->wB == 1 || d->wB == d->B, "conv_rs: wB must be 1 or B");
  RS_SUP(d->out != nullptr || (d->rgb_w != nullptr && d->rgb_out != nullptr), "conv_rs: out may be NULL only with the fused ToRGB tail (RGB-only launch)");
  RS_SUP(d->out == nullptr || (d->out_sx == d->Cout && d->out_sy == (int64_t)d->Wo * d->Cout && d->out_sb == (int64_t)d->Ho * d->Wo * d->Cout && d->phase_off[0] == 0),
         "conv_rs: dense NHWC output only");
  RS_SUP(!d->res && !d->slope_vec && !d->src_scale[0] && !d->src_affine[0] && !d->round_tf32, "conv_rs: no residual / PReLU / source transforms / TF32 rounding");
  RS_SUP(d->act == VT_ACT_NONE || d->act == VT_ACT_LRELU, "conv_rs: activation must be none or leaky-relu");
  RS_SUP(((uintptr_t)d->out & 15) == 0 && ((uintptr_t)d->src[0] & 15) == 0 && ((uintptr_t)d->weight_bf16x3 & 15) == 0, "conv_rs: pointers must be 16-byte aligned");
  RS_SUP(!d->rgb_w || (!d->rgb_skip || (d->Ho % 2 == 0 && d->Wo % 2 == 0)), "conv_rs: the fused skip needs even Ho/Wo");
  RS_SUP(d->alpha == 1.f, "conv_rs: alpha must be 1");
  return 1;
#undef RS_SUP
}

// cta group and shared-memory plan of a launch; returns 0 when the layer does not fit (e.g. 64 -> 64 on a single CTA)
int rs_plan(const vt_conv_desc* d, int* cg_out, int* a_stages_out, int* smem_out) {
  const int cg = g_rs_cg ? g_rs_cg : (d->W >= 2 * RS_PX ? 2 : 1);
  if (cg != 1 && cg != 2) return 0;
  const int w_bytes = (d->src_c[0] / 32) * 3 * (3 * d->Cout / cg) * 128;
  const int fixed = w_bytes + 1024 + RS_STAGING + RS_CONST_FLOATS * 4 + 512 /*barriers*/ + 1024 /*alignment*/;
  int a_stages = (RS_MAX_SMEM - fixed) / RS_A_STAGE;
  if (a_stages > 8) a_stages = 8;
  if (a_stages < 3) return 0;
  *cg_out = cg; *a_stages_out = a_stages; *smem_out = a_stages * RS_A_STAGE + fixed;
  return 1;
}

}  // namespace

extern unsigned long long* g_tc_dbg_export;

extern "C" int vt_conv2d_rs_supported(const vt_conv_desc* d) {
  if (!d || d->struct_size != (int)sizeof(vt_conv_desc)) return 0;
  if (!rs_check(d, false)) return 0;
  int cg, st, sm;
  return rs_plan(d, &cg, &st, &sm);
}

int vt_rs_set_option(const char* key, int value, int* old) {
  if (key && strcmp(key, "rs_cg") == 0) { *old = g_rs_cg; g_rs_cg = value; return 1; }
  if (key && strcmp(key, "rs_rows") == 0) { *old = g_rs_rows; g_rs_rows = value; return 1; }
  if (key && strcmp(key, "rs_strict") == 0) { *old = g_rs_strict; g_rs_strict = value; return 1; }
  return 0;
}

extern "C" int vt_conv2d_rs(const vt_conv_desc* d, float acc_scale, void* stream) {
  if (vt_validate_conv_desc(d, "conv2d_rs")) return 1;
  if (!rs_check(d, true)) return 1;
  VT_CHECK(acc_scale > 0.f, "conv_rs: acc_scale must be positive");
  static thread_local RsArgs a;
  memset(&a, 0, sizeof(a));
  const int Cin = d->src_c[0], Cout = d->Cout;
  int cg = 0, plan_stages = 0, smem_bytes = 0;
  VT_CHECK(rs_plan(d, &cg, &plan_stages, &smem_bytes), "conv_rs: shared memory plan does not fit (Cin=%d Cout=%d)", Cin, Cout);
  a.B = d->B; a.H = d->H; a.W = d->W; a.Cout = Cout; a.KC = Cin / 32; a.wB = d->wB;
  a.S = 512 / Cout - 2;
  a.w_tile_bytes = (3 * Cout / cg) * 128;
  a.bias = d->bias; a.noise = d->noise; a.noise_w = d->noise_w;
  a.act = d->act; a.slope = d->slope; a.gain = d->gain;
  a.rgb_w = d->rgb_w; a.rgb_bias = d->rgb_bias; a.rgb_skip = d->rgb_skip; a.rgb_skip_kernel = d->rgb_skip_kernel; a.rgb_out = d->rgb_out;
  a.fmt = d->bf16x3_nstack == 3 ? 1 : 0;
  a.acc_scale = acc_scale;
  a.dbg = g_tc_dbg_export;
  a.strict_release = g_rs_strict;
  // strips: 128*cg pixels wide; rows per strip trade the 2 halo rows per strip against the tail of the last wave
  a.strips_x = (int)vt_cdiv(d->W, RS_PX * cg);
  const int units = vt_num_sms() / cg;
  int best_rows = 0; double best_cost = 1e30;
  for (int rows = 16; rows <= 256; rows += 8) {
    const int64_t sy = vt_cdiv(d->H, rows);
    const int64_t strips = (int64_t)d->B * a.strips_x * sy;
    const double cost = (double)vt_cdiv(strips, units) * (rows + 2 + 6);    // waves x (rows + halo + pipeline fill), in row times
    if (cost < best_cost - 1e-9) { best_cost = cost; best_rows = rows; }
  }
  a.rows_per_strip = g_rs_rows > 0 ? g_rs_rows : best_rows;
  if (a.rows_per_strip > d->H) a.rows_per_strip = d->H;
  a.strips_y = (int)vt_cdiv(d->H, a.rows_per_strip);
  const int64_t total = (int64_t)d->B * a.strips_x * a.strips_y;
  VT_CHECK(total < (1LL << 30), "conv_rs: too many strips");
  a.total_strips = (int)total;
  a.a_stages = plan_stages;
  // tensor maps
  {
    const uint64_t cs = (uint64_t)d->src_cstride[0];
    const uint64_t dims[4] = {cs, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    const uint64_t str[3] = {cs * 4, (uint64_t)d->W * cs * 4, (uint64_t)d->H * d->W * cs * 4};
    const uint32_t box[4] = {32, RS_BOX, 1, 1};
    if (vt_tc_make_map4(&a.in_map, d->src[0], dims, str, box, "rs input", false)) return 1;
  }
  {
    const uint64_t rows = (uint64_t)(3 * Cout), tiles = (uint64_t)(a.KC * 3);
    const uint64_t dims[4] = {64, rows, tiles, (uint64_t)d->wB};
    const uint64_t str[3] = {128, rows * 128, tiles * rows * 128};
    const uint32_t box[4] = {64, (uint32_t)(3 * Cout / cg), 1, 1};
    if (vt_tc_make_map4(&a.w_map, d->weight_bf16x3, dims, str, box, "rs weight", true)) return 1;
  }
  a.store_out = d->out != nullptr;
  if (a.store_out) {
    const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    const uint64_t str[3] = {(uint64_t)Cout * 4, (uint64_t)d->W * Cout * 4, (uint64_t)d->H * d->W * Cout * 4};
    const uint32_t box[4] = {32, 32, 1, 1};
    if (vt_tc_make_map4(&a.out_map, d->out, dims, str, box, "rs output", false)) return 1;
  }
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(conv_rs_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_MAX_SMEM);
    if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(conv_rs_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_MAX_SMEM);
  });
  VT_CHECK(attr_err == cudaSuccess, "conv_rs: cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
  if (cg == 1) {
    int grid = vt_num_sms();
    if (grid > a.total_strips) grid = a.total_strips;
    conv_rs_kernel<1><<<grid, RS_THREADS, smem_bytes, (cudaStream_t)stream>>>(a);
  } else {
    int pairs = vt_num_sms() / 2;
    if (pairs > a.total_strips) pairs = a.total_strips;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(RS_THREADS);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VT_CUDA(cudaLaunchKernelEx(&cfg, conv_rs_kernel<2>, a));
  }
  VT_LAUNCH_CHECK();
  return 0;
}
