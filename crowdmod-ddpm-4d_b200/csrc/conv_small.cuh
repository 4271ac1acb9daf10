// Fused GroupNorm(+SiLU) -> 3x3x3 conv (+ fused 1x1x1 match_input, bias, time-embedding add, residual) for the
// COARSEST UNet level of the sampling path, tcgen05 / TMEM, sm_100a.
//
// Reference ops replaced (one launch instead of two or three): nn.GroupNorm(8) + nn.SiLU + nn.Conv3d k3 p1 of a
// ResnetBlock (models/backbones/layers.py:57-58 and :70-73, 75) at the level where a sample is a few dozen pixels
// (ATC: 3x9 grid x 2 frames = 54 pixels, 128 channels).  Measured in round 1 (profiles/r1_launchlist_summary.csv):
// that level took 38 % of the denoiser step for 10 % of its FLOPs -- 13 split-K conv launches of ~17 us (each CTA
// issues dozens of dependent TMA im2col loads for a handful of MMAs) plus 13 GroupNorm launches of ~5.5 us.
//
// Design
//   * The level has D = 2 frames.  They are folded into the GEMM: K = (tap_h, tap_w, d_in, ci), N = (d_out, co), with
//     the combined weights W'[(d_out,co)][(th,tw,d_in,ci)] = w[co][ci][td = d_in - d_out + 1][th][tw] (every (d_in,
//     d_out) pair is a valid tap, so nothing is multiplied by padding along time); 9 spatial taps remain.
//   * M tile = 128 "positions": spt whole samples, each laid out as (H+1) x (W+1) positions whose last row / column
//     are zeros (the padding between rows and between samples).  The activation operand of the tile lives in shared
//     memory ONCE, K-major with the 128-byte swizzle, as [chunk of 64 (d_in,ci) channels][lead + 128 + lead rows];
//     the nine taps are nine row-offset descriptor views of it (+-(W+1) rows, +-1 row).  No im2col loads at all.
//   * The CTA builds that operand itself: it reads the fp32 source(s) of the GroupNorm (two sources = the decoder's
//     skip concat, unet.py:160), computes the group statistics of its samples (shifted sums + Chan merge), and writes
//     silu(GN(x)) as fp16 straight into the swizzled layout.  The 1x1x1 match_input source is built the same way
//     (raw cast) BEFORE the normalised operand, multiplied into the same accumulator, and its buffer is reused.
//   * grid = (tiles, cout/32): a CTA owns 32 output channels -> accumulator columns (term, d_out, co32) = 128, one
//     128x128x16 MMA per k16 step covers the hi and the lo weight term and both output frames.  Weights stream from L2
//     through a TMA ring (the only bulk traffic: 128 rows x K per CTA); warp 0 starts that stream at kernel entry, so
//     it overlaps the operand build.
//   * warps: 0 TMA producer, 1 MMA issuer, 2..9 operand builders, then epilogue (TMEM -> +bias +temb +residual ->
//     fp32 / fp16 channels-last stores, full 128-byte lines per pixel).
// Numerics: the same operand classes as the unfused path (fp16 activations, hi+lo fp16 weights, fp32 accumulate,
// fp32 GroupNorm statistics).
#pragma once
#include "common.cuh"

namespace cm {

constexpr int SC_THREADS = 320;
constexpr int SC_WORKERS = 256;
constexpr int SC_MAX_STAGES = 6;
constexpr int SC_BK = 64;              // K elements per k-block: 128-byte rows, SWIZZLE_128B
constexpr int SC_NB = 128;             // accumulator columns per CTA: (term, d_out, 32 output channels)
constexpr int SC_STAGE_BYTES = SC_NB * SC_BK * 2;

struct SmallConvParams {
  CUtensorMap bmap;      // packed weights [(cout/32)*128 rows][Ktot] K-major, box {64, 128}
  const float* src0;     // GroupNorm sources, fp32 channels-last [B][2][H][W][c0 | c1]
  const float* src1;
  int c0, c1;
  const float* gamma;
  const float* beta;
  float eps;
  int silu;
  const float* xs0;      // fused 1x1x1 source(s): raw fp32 [B][2][H][W][cx0 | cx1], or nullptr
  const float* xs1;
  int cx0, cx1;
  int B, H, W;
  int PP;                // positions per sample (H+1)*(W+1)
  int spt;               // samples per tile
  int lead;              // zero rows before / after the 128 tile rows
  int RA;                // rows per operand chunk = 2*lead + 128 (multiple of 8)
  int cin, cx, cout;
  int nchunk, nchunk_x;  // 64-channel chunks of the folded K: 2*cin/64, 2*cx/64
  int stages;
  const float* bias;
  const float* bias2;
  const float* temb;
  const int* t_dev;
  int temb_ld, temb_bstride;
  const float* resid;    // fp32 [B][2][H][W][cout] or nullptr
  float* out32;
  __half* out16;
  int* err_flag;
};

struct ScChan { float n, mean, m2; };
__device__ __forceinline__ ScChan sc_merge(const ScChan a, const ScChan b) {
  if (b.n == 0.f) return a;
  if (a.n == 0.f) return b;
  const float nn = a.n + b.n;
  const float delta = b.mean - a.mean;
  const float w = __fdividef(b.n, nn);
  ScChan r;
  r.n = nn;
  r.mean = a.mean + delta * w;
  r.m2 = a.m2 + b.m2 + delta * delta * (a.n * w);
  return r;
}
__device__ __forceinline__ float sc_silu(float y) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return y * r;
}
__device__ __forceinline__ void sc_workers_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// GroupNorm statistics of the tile's samples: stat[s][g] = {mean, rstd}.  Every worker keeps one channel quad.
__device__ __forceinline__ void sc_stats(const SmallConvParams& P, int tile, int wt, float* tri, float* stat) {
  const int C = P.c0 + P.c1, Q = C >> 2, vpp = Q >> 3;
  const int RPP = SC_WORKERS / Q, T = RPP * Q;
  const int q = wt % Q, rg = wt / Q, c = q * 4;
  const bool from0 = c < P.c0;
  const float* src = from0 ? P.src0 + c : P.src1 + (c - P.c0);
  const int ld = from0 ? P.c0 : P.c1;
  const int npix = 2 * P.H * P.W;
  const int warp = wt >> 5, lane = wt & 31;
  for (int s = 0; s < P.spt; ++s) {
    const int b = tile * P.spt + s;
    if (b >= P.B) break;                                   // uniform over the CTA
    float s1 = 0.f, s2 = 0.f, k = 0.f, cnt = 0.f;
    if (wt < T) {
      bool first = true;
      for (int i = rg; i < npix; i += RPP) {
        const float4 v = *reinterpret_cast<const float4*>(src + (static_cast<size_t>(b) * npix + i) * ld);
        if (first) { k = v.x; first = false; }
        const float dx = v.x - k, dy = v.y - k, dz = v.z - k, dw = v.w - k;
        s1 += (dx + dy) + (dz + dw);
        s2 += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        cnt += 4.f;
      }
    }
    {
      const float d = cnt > 0.f ? s1 / cnt : 0.f;
      tri[wt * 3 + 0] = cnt;
      tri[wt * 3 + 1] = k + d;
      tri[wt * 3 + 2] = fmaxf(s2 - s1 * d, 0.f);
    }
    sc_workers_sync();
    {
      // worker warp g merges the threads of group g: t = rr*Q + g*vpp + o  (rr < RPP, o < vpp), <= 32 members
      const int members = RPP * vpp;
      ScChan a{0.f, 0.f, 0.f};
      if (lane < members) {
        const int rr = lane / vpp, o = lane - rr * vpp;
        const int t = rr * Q + warp * vpp + o;
        a = ScChan{tri[t * 3], tri[t * 3 + 1], tri[t * 3 + 2]};
      }
#pragma unroll
      for (int w = 1; w < 32; w <<= 1) {
        ScChan o;
        o.n = __shfl_xor_sync(0xffffffffu, a.n, w);
        o.mean = __shfl_xor_sync(0xffffffffu, a.mean, w);
        o.m2 = __shfl_xor_sync(0xffffffffu, a.m2, w);
        a = (lane & w) ? sc_merge(o, a) : sc_merge(a, o);   // same order on both sides: identical results
      }
      if (lane == 0) {
        stat[(s * 8 + warp) * 2 + 0] = a.mean;
        stat[(s * 8 + warp) * 2 + 1] = 1.0f / sqrtf(a.m2 / fmaxf(a.n, 1.f) + P.eps);
      }
    }
    sc_workers_sync();
  }
}

// Writes one operand into the swizzled K-major chunk layout: NORM = silu?(GroupNorm(src0 || src1)), else the raw
// (xs0 || xs1) cast.  Separator / lead rows are never touched (zero).
template <bool NORM>
__device__ __forceinline__ void sc_build(const SmallConvParams& P, uint8_t* A, int tile, int wt, const float* stat) {
  const float* s0 = NORM ? P.src0 : P.xs0;
  const float* s1 = NORM ? P.src1 : P.xs1;
  const int c0 = NORM ? P.c0 : P.cx0, c1 = NORM ? P.c1 : P.cx1;
  const int C = c0 + c1, Q = C >> 2;
  const int RPP = SC_WORKERS / Q, T = RPP * Q;
  if (wt >= T) return;
  const int q = wt % Q, rg = wt / Q, c = q * 4;
  const bool from0 = c < c0;
  const float* src = from0 ? s0 + c : s1 + (c - c0);
  const int ld = from0 ? c0 : c1;
  const int HW = P.H * P.W, npix = 2 * HW;
  const int g = c / (C >> 3);
  float4 ga = make_float4(1.f, 1.f, 1.f, 1.f), be = make_float4(0.f, 0.f, 0.f, 0.f);
  if (NORM) {
    ga = *reinterpret_cast<const float4*>(P.gamma + c);
    be = *reinterpret_cast<const float4*>(P.beta + c);
  }
  const int cpd = C >> 6;                                   // 64-channel chunks per input frame
  const uint32_t cc = static_cast<uint32_t>(c & 63);
  for (int s = 0; s < P.spt; ++s) {
    const int b = tile * P.spt + s;
    if (b >= P.B) break;
    float4 sc = ga, sh = be;
    if (NORM) {
      const float mean = stat[(s * 8 + g) * 2], rstd = stat[(s * 8 + g) * 2 + 1];
      sc = make_float4(rstd * ga.x, rstd * ga.y, rstd * ga.z, rstd * ga.w);
      sh = make_float4(be.x - mean * sc.x, be.y - mean * sc.y, be.z - mean * sc.z, be.w - mean * sc.w);
    }
    for (int i = rg; i < npix; i += RPP) {
      const int d = i >= HW ? 1 : 0;
      const int r = i - d * HW;
      const int h = r / P.W, w = r - h * P.W;
      const float4 v = *reinterpret_cast<const float4*>(src + (static_cast<size_t>(b) * npix + i) * ld);
      float y0 = v.x, y1 = v.y, y2 = v.z, y3 = v.w;
      if (NORM) {
        y0 = fmaf(v.x, sc.x, sh.x); y1 = fmaf(v.y, sc.y, sh.y); y2 = fmaf(v.z, sc.z, sh.z); y3 = fmaf(v.w, sc.w, sh.w);
        if (P.silu) { y0 = sc_silu(y0); y1 = sc_silu(y1); y2 = sc_silu(y2); y3 = sc_silu(y3); }
      }
      const uint32_t p = static_cast<uint32_t>(P.lead + s * P.PP + h * (P.W + 1) + w);
      const uint32_t kc = static_cast<uint32_t>(d * cpd + (c >> 6));
      const uint32_t off = kc * static_cast<uint32_t>(P.RA) * 128u + p * 128u + ((((cc >> 3) ^ (p & 7u)) << 4) | ((cc & 7u) * 2u));
      __half2 h0 = __floats2half2_rn(y0, y1), h1 = __floats2half2_rn(y2, y3);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(A + off) = u;
    }
  }
}

__global__ void __launch_bounds__(SC_THREADS, 1) conv_small_kernel(const __grid_constant__ SmallConvParams P) {
  constexpr uint32_t IDESC = make_idesc_f16(128, SC_NB);
  constexpr uint32_t DESC_HI = kmajor_desc_hi(128);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int ncmax = P.nchunk > P.nchunk_x ? P.nchunk : P.nchunk_x;
  const uint32_t a_bytes = static_cast<uint32_t>(ncmax) * P.RA * 128u;      // multiple of 1024
  uint8_t* A = smem;
  uint8_t* Bst = smem + a_bytes;
  uint8_t* tail = Bst + P.stages * SC_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + SC_MAX_STAGES;
  uint64_t* a_ready = empty_bar + SC_MAX_STAGES;           // [2]: raw slab operand, normalised operand
  uint64_t* slab_done = a_ready + 2;
  uint64_t* tmem_full = slab_done + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* colv = reinterpret_cast<float*>(tail + 256);       // [32]
  float* stat = colv + 32;                                  // [spt][8][2]  (spt <= 32)
  float* tri = stat + 32 * 16;                              // [256][3]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, slice = blockIdx.y;
  const int nkb_x = P.nchunk_x, nkb_main = 9 * P.nchunk;
  const int S = P.stages;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&P.bmap);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&a_ready[0], 1);
    mbar_init(&a_ready[1], 1);
    mbar_init(slab_done, 1);
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, SC_NB);
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer: the weight stream (slab k-blocks first, then 9 taps x chunks) ==========
    int s = 0, git = 0;
    uint32_t ph = 0;
    const int kx0 = 18 * P.cin;                            // first packed-K column of the 1x1x1 slab
    for (int kb = 0; kb < nkb_x + nkb_main; ++kb, ++git) {
      if (git >= S && !mbar_wait(&empty_bar[s], ph ^ 1, P.err_flag, 501)) break;
      if (elect_one()) {
        mbar_expect_tx(&full_bar[s], SC_STAGE_BYTES);
        const int kcol = kb < nkb_x ? kx0 + kb * SC_BK : (kb - nkb_x) * SC_BK;
        tma_load_2d(&P.bmap, &full_bar[s], Bst + s * SC_STAGE_BYTES, kcol, slice * SC_NB);
      }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t a_lo_base = kmajor_desc_lo(smem_u32(A));
    const uint32_t chunk_lo = (static_cast<uint32_t>(P.RA) * 128u) >> 4;
    const uint32_t b_lo_base = kmajor_desc_lo(smem_u32(Bst));
    int s = 0;
    uint32_t ph = 0, acc = 0;
    bool alive = true;
    if (nkb_x) {
      alive = mbar_wait(&a_ready[0], 0, P.err_flag, 502);
      tc_fence_after();
      for (int kb = 0; kb < nkb_x && alive; ++kb) {
        if (!mbar_wait(&full_bar[s], ph, P.err_flag, 503)) { alive = false; break; }
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = a_lo_base + kb * chunk_lo + ((static_cast<uint32_t>(P.lead) * 128u) >> 4);
          const uint32_t b_lo = b_lo_base + ((s * SC_STAGE_BYTES) >> 4);
#pragma unroll
          for (int j = 0; j < SC_BK / 16; ++j) {
            umma_f16_lohi(tmem_base, a_lo + 2 * j, b_lo + 2 * j, DESC_HI, IDESC, acc);
            acc = 1;
          }
          umma_commit(&empty_bar[s]);
          if (kb == nkb_x - 1) umma_commit(slab_done);      // the raw operand may be overwritten
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
    if (alive) alive = mbar_wait(&a_ready[1], 0, P.err_flag, 504);
    tc_fence_after();
    for (int kb = 0; kb < nkb_main && alive; ++kb) {
      if (!mbar_wait(&full_bar[s], ph, P.err_flag, 505)) { alive = false; break; }
      tc_fence_after();
      if (elect_one()) {
        const int tap = kb / P.nchunk, kc = kb - tap * P.nchunk;
        const int th = tap / 3, tw = tap - th * 3;
        const int row = P.lead + (th - 1) * (P.W + 1) + (tw - 1);
        const uint32_t a_lo = a_lo_base + kc * chunk_lo + ((static_cast<uint32_t>(row) * 128u) >> 4);
        const uint32_t b_lo = b_lo_base + ((s * SC_STAGE_BYTES) >> 4);
#pragma unroll
        for (int j = 0; j < SC_BK / 16; ++j) {
          umma_f16_lohi(tmem_base, a_lo + 2 * j, b_lo + 2 * j, DESC_HI, IDESC, acc);
          acc = 1;
        }
        umma_commit(&empty_bar[s]);
        if (kb == nkb_main - 1) umma_commit(tmem_full);
      }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else {
    // ===================== operand builders, then epilogue (warps 2..9) =====================
    const int wt = threadIdx.x - 64;
    {
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      uint4* a4 = reinterpret_cast<uint4*>(A);
      for (uint32_t i = wt; i < a_bytes / 16; i += SC_WORKERS) a4[i] = z;
      // per-column epilogue constants of this CTA's 32 output channels
      if (wt < 32) {
        const int n = slice * 32 + wt;
        float v = P.bias ? P.bias[n] : 0.f;
        if (P.bias2) v += P.bias2[n];
        if (P.temb && P.temb_bstride == 0) v += P.temb[static_cast<size_t>(P.t_dev ? *P.t_dev : 0) * P.temb_ld + n];
        colv[wt] = v;
      }
    }
    sc_workers_sync();
    if (nkb_x) {
      sc_build<false>(P, A, tile, wt, nullptr);
      fence_proxy_async();
      sc_workers_sync();
      if (wt == 0) mbar_arrive(&a_ready[0]);
    }
    sc_stats(P, tile, wt, tri, stat);                       // global reads only: overlaps the slab MMAs
    if (nkb_x) {
      mbar_wait(slab_done, 0, P.err_flag, 506);             // the slab MMAs have read the raw operand
      tc_fence_after();
    }
    sc_build<true>(P, A, tile, wt, stat);
    fence_proxy_async();
    sc_workers_sync();
    if (wt == 0) mbar_arrive(&a_ready[1]);

    // ---- epilogue: lane quarter = warp % 4, output frame = (warp - 2) / 4 ----
    const int quarter = warp & 3;
    const int dout = (warp - 2) >> 2;
    const int prow = quarter * 32 + lane;                  // tile position of this thread
    const int s = prow / P.PP;
    const int r = prow - s * P.PP;
    const int h = r / (P.W + 1), w = r - h * (P.W + 1);
    const int b = tile * P.spt + s;
    const bool valid = s < P.spt && b < P.B && h < P.H && w < P.W;
    const size_t m = valid ? ((static_cast<size_t>(b) * 2 + dout) * P.H + h) * P.W + w : 0;
    const int n0 = slice * 32;
    float4 rv[8];
    if (valid && P.resid) {
#pragma unroll
      for (int i = 0; i < 8; ++i) rv[i] = *reinterpret_cast<const float4*>(P.resid + m * P.cout + n0 + 4 * i);
    }
    const float* trow = (P.temb && P.temb_bstride != 0 && valid)
                            ? P.temb + static_cast<size_t>(b) * P.temb_bstride + n0 : nullptr;
    if (!mbar_wait(tmem_full, 0, P.err_flag, 507)) return;
    tc_fence_after();
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    float v[32];
    {
      float a0[16], a1[16], l0[16], l1[16];
      tmem_ld16_async(t_lane + dout * 32, a0);
      tmem_ld16_async(t_lane + dout * 32 + 16, a1);
      tmem_ld16_async(t_lane + 64 + dout * 32, l0);
      tmem_ld16_async(t_lane + 64 + dout * 32 + 16, l1);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[i] = a0[i] + l0[i];
        v[16 + i] = a1[i] + l1[i];
      }
    }
    if (valid) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 c4 = *reinterpret_cast<const float4*>(colv + i);
        v[i] += c4.x; v[i + 1] += c4.y; v[i + 2] += c4.z; v[i + 3] += c4.w;
      }
      if (trow) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 t4 = *reinterpret_cast<const float4*>(trow + i);
          v[i] += t4.x; v[i + 1] += t4.y; v[i + 2] += t4.z; v[i + 3] += t4.w;
        }
      }
      if (P.resid) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[4 * i] += rv[i].x; v[4 * i + 1] += rv[i].y; v[4 * i + 2] += rv[i].z; v[4 * i + 3] += rv[i].w;
        }
      }
      if (P.out32) {
        float* op = P.out32 + m * P.cout + n0;
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(op + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
      if (P.out16) {
        __half* op = P.out16 + m * P.cout + n0;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          __half2 h0 = __floats2half2_rn(v[i], v[i + 1]), h1 = __floats2half2_rn(v[i + 2], v[i + 3]);
          __half2 h2 = __floats2half2_rn(v[i + 4], v[i + 5]), h3 = __floats2half2_rn(v[i + 6], v[i + 7]);
          uint4 u;
          u.x = *reinterpret_cast<uint32_t*>(&h0);
          u.y = *reinterpret_cast<uint32_t*>(&h1);
          u.z = *reinterpret_cast<uint32_t*>(&h2);
          u.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(op + i) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, SC_NB);
}

struct SmallLaunch {
  SmallConvParams p;
  dim3 grid;
  size_t smem;
  bool ok;
};

// Elements of the packed d-folded weights of one conv: (cout/32)*128 rows x (18*cin + 2*cx) columns.
size_t small_packed_elems(int cin, int cx, int cout);
// True when the fused kernel covers the geometry (D == 2 frames, a sample's padded grid fits one tile, channel
// counts in 64-channel chunks, hi+lo weights).
bool small_supported(int D, int H, int W, int cin, int cx, int cout, int terms);
// Fills L (ok = false when unsupported).  Pointers other than the weights are set by the caller.
int small_prepare(SmallLaunch* L, int B, int D, int H, int W, int c0, int c1, int cx0, int cx1, const __half* wpacked,
                  int cout, int terms);
int small_enqueue(const SmallLaunch& L, cudaStream_t st);
int small_init();

}  // namespace cm
