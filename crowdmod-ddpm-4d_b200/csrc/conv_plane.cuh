// "Plane-tile" implicit-GEMM 3-D convolution (k3 s1 p1) on tcgen05 / TMEM, sm_100a.
//
// Why a second conv kernel.  Measured on B200 (tools/plane_dbg.py, profiles/), conv_umma_kernel
// (conv_umma.cuh) is bound by what one SM can pull out of L2 (~32-45 B/clk): it re-fetches its A
// tile for each of the 27 taps (TMA im2col) and its weight tile for every 128-pixel M tile; and a
// 128 x 32 tcgen05.mma re-reads its 4 KB A operand from shared memory for only 32 output columns,
// so narrow layers are shared-memory-bandwidth bound inside the tensor pipe as well.  This kernel
//   * gives a work unit R whole planes (or a block of HB rows of one plane) of one sample: with W
//     padded to W+2 the unit is a flat array of P = R*HB*(W+2) "positions" and ONE tiled TMA box
//     {BK ch, W+2, HB, R} (zero-filled halo) per (td, th, channel chunk) serves all three tw taps:
//     a tap shift is a +1/+2 row offset of the UMMA shared-memory descriptor (positions on the two
//     pad columns compute garbage that the epilogue drops) -> 9 A loads per channel chunk, not 27.
//     (The UMMA swizzle is a function of the absolute shared-memory address: an offset start
//     address needs no descriptor base_offset -- verified on hardware, tools/plane_probe.py.)
//   * keeps ceil(P/128) accumulators in TMEM, so every weight tile is loaded once per unit and
//     reused by all of its M tiles;
//   * stacks the hi and lo weight terms along N: one 128 x (2*BN) MMA instead of two 128 x BN
//     ones, halving the A-operand shared-memory reads; the epilogue adds the two halves;
//   * is persistent: one CTA per SM walks the units; TMEM is double-buffered, so the epilogue of
//     unit i (coalesced through a shared-memory transpose) overlaps the main loop of unit i+1.
// Same operands, same packed weights, same epilogue semantics as conv_umma_kernel (mode 0 with
// the optional fused 1x1x1 match_input K-slab).
#pragma once
#include "common.cuh"

namespace cm {

constexpr int PL_THREADS = 192;      // warp0 TMA, warp1 MMA, warps2-5 epilogue
constexpr int PL_MAX_STAGES = 6;

struct PlaneParams {
  CUtensorMap amap;      // main source, tiled 5-D (C, W, H, D, N), box {BK, W+2, HB, R, 1}
  CUtensorMap xmap;      // optional 1x1x1 source over the same grid (same box)
  CUtensorMap bmap;      // packed weights [terms*cout][Ktot], K-major (as conv_umma)
  int H, W, D, Wp;       // plane geometry, Wp = W + 2
  int R, HB;             // unit = R planes x HB rows (R > 1 only with HB == H)
  int P;                 // positions per unit = R*HB*Wp
  int ntiles;            // ceil(P / 128) accumulators per unit
  int units_per_sample;  // (D / R) * (H / HB)
  int n_units;           // B * units_per_sample * (cout / BN)
  int n_ntiles;          // cout / BN
  int a_stage_bytes;     // bytes reserved for the A box per stage (covers ntiles*128 + 2 rows)
  int cin_main, cin_extra, cout, terms, stages;
  int dbg;               // bring-up knobs (CM_PLANE_DBG): 1 no A loads, 2 no B loads, 4 no MMA, 8 no stores
  const float* bias;
  const float* bias2;
  const float* temb;
  const int* t_dev;
  int temb_ld, temb_bstride;
  const float* resid;
  float* out32;
  __half* out16;
  int out_ld;
  int* err_flag;
};

__device__ __forceinline__ void tma_load_tile_5d(const void* desc, uint64_t* bar, void* smem, int c0,
                                                 int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

struct PlaneUnit {
  int n, d0, h0, n_tile;
};
__device__ __forceinline__ PlaneUnit plane_unit(const PlaneParams& P, int u) {
  PlaneUnit r;
  r.n_tile = u % P.n_ntiles;
  int v = u / P.n_ntiles;
  r.n = v / P.units_per_sample;
  v -= r.n * P.units_per_sample;
  const int hblocks = P.H / P.HB;
  const int db = v / hblocks;
  r.d0 = db * P.R;
  r.h0 = (v - db * hblocks) * P.HB;
  return r;
}

template <int BN, int BK, int TERMS>
__global__ void __launch_bounds__(PL_THREADS, 1)
conv_plane_kernel(const __grid_constant__ PlaneParams P) {
  constexpr int ROWB = BK * 2;
  constexpr int NST = BN * TERMS;                    // stacked N of one MMA (hi | lo)
  constexpr int B_TAP = NST * ROWB;                  // weight bytes per tap per stage
  constexpr int B_STAGE = 3 * B_TAP;
  constexpr uint32_t IDESC = make_idesc_f16(128, NST);
  constexpr int TLD = BN + 4;                        // transpose-buffer row (floats)

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int S = P.stages;
  const int stage_bytes = P.a_stage_bytes + B_STAGE;       // multiple of 1024
  uint8_t* tail = smem + S * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + PL_MAX_STAGES;
  uint64_t* tmem_full = empty_bar + PL_MAX_STAGES;         // [2]
  uint64_t* tmem_empty = tmem_full + 2;                    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* colv = reinterpret_cast<float*>(tail + 256);      // [BN]
  float* tbuf_all = reinterpret_cast<float*>(tail + 256 + 512);   // [4 warps][32][TLD]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncm = P.cin_main / BK;
  const int nks_main = 9 * ncm;                       // (td, th, chunk) steps
  const int nks = nks_main + P.cin_extra / BK;
  const uint32_t buf_cols = static_cast<uint32_t>(P.ntiles * NST);
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2 * buf_cols) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.amap);
    tma_prefetch_desc(&P.bmap);
    if (P.cin_extra) tma_prefetch_desc(&P.xmap);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 4);                   // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    const uint32_t a_bytes = static_cast<uint32_t>(P.P) * ROWB;
    int s = 0;
    uint32_t ph = 0;
    bool alive = true;
    for (int u = blockIdx.x; u < P.n_units && alive; u += gridDim.x) {
      const PlaneUnit U = plane_unit(P, u);
      int td = 0, th = 0, cc = 0;
      for (int ks = 0; ks < nks; ++ks) {
        if (!mbar_wait(&empty_bar[s], ph ^ 1, P.err_flag, 401)) { alive = false; break; }
        uint8_t* sa = smem + s * stage_bytes;
        uint8_t* sb = sa + P.a_stage_bytes;
        if (elect_one()) {
          if (ks < nks_main) {
            const uint32_t tx = ((P.dbg & 1) ? 0u : a_bytes) + ((P.dbg & 2) ? 0u : (uint32_t)B_STAGE);
            if (tx) mbar_expect_tx(&full_bar[s], tx); else mbar_arrive(&full_bar[s]);
            if (!(P.dbg & 1))
              tma_load_tile_5d(&P.amap, &full_bar[s], sa, cc * BK, -1, U.h0 + th - 1, U.d0 + td - 1, U.n);
            const int kcol = ((td * 3 + th) * 3) * P.cin_main + cc * BK;
            if (!(P.dbg & 2)) {
#pragma unroll
              for (int tw = 0; tw < 3; ++tw)
#pragma unroll
                for (int t = 0; t < TERMS; ++t)
                  tma_load_2d(&P.bmap, &full_bar[s], sb + tw * B_TAP + t * (BN * ROWB), kcol + tw * P.cin_main,
                              U.n_tile * BN + t * P.cout);
            }
          } else {
            // fused 1x1x1 source: centre tap only (tw = 1 against a box that starts at w = -1)
            const int xc = ks - nks_main;
            mbar_expect_tx(&full_bar[s], a_bytes + B_TAP);
            tma_load_tile_5d(&P.xmap, &full_bar[s], sa, xc * BK, -1, U.h0, U.d0, U.n);
#pragma unroll
            for (int t = 0; t < TERMS; ++t)
              tma_load_2d(&P.bmap, &full_bar[s], sb + 1 * B_TAP + t * (BN * ROWB), 27 * P.cin_main + xc * BK,
                          U.n_tile * BN + t * P.cout);
          }
        }
        if (++cc == ncm) {
          cc = 0;
          if (++th == 3) { th = 0; ++td; }
        }
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t DESC_HI = kmajor_desc_hi(ROWB);
    constexpr uint32_t TILE_LO = (128 * ROWB) >> 4;          // descriptor step between M tiles
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    bool alive = true;
    for (int u = blockIdx.x; u < P.n_units && alive; u += gridDim.x, ++it) {
      const int buf = it & 1;
      // wait until the epilogue has drained this accumulator buffer (first two units: free)
      if (it >= 2 && !mbar_wait(&tmem_empty[buf], ((it >> 1) - 1) & 1, P.err_flag, 404)) break;
      tc_fence_after();
      const uint32_t d_base = tmem_base + buf * buf_cols;
      for (int ks = 0; ks < nks; ++ks) {
        if (!mbar_wait(&full_bar[s], ph, P.err_flag, 402)) { alive = false; break; }
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
          const uint32_t b_lo0 = kmajor_desc_lo(a_addr + P.a_stage_bytes);
          const int tw0 = ks < nks_main ? 0 : 1, tw1 = ks < nks_main ? 3 : 2;
          for (int rep = 0; rep < ((P.dbg & 128) ? 2 : 1); ++rep)
          for (int tw = tw0; tw < ((P.dbg & 4) ? tw0 : tw1); ++tw) {
            const uint32_t a_lo0 = kmajor_desc_lo(a_addr + ((P.dbg & 16) ? 0 : tw) * ROWB + ((P.dbg & 32) ? tw * 8 * ROWB : 0));      // tap shift = +tw rows
            uint32_t first = (ks == 0 && tw == tw0) ? 0u : 1u;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint32_t b_lo = b_lo0 + ((tw * B_TAP) >> 4) + 2 * k;
              uint32_t a_lo = a_lo0 + 2 * k;
              uint32_t d = d_base;
#pragma unroll 4
              for (int r = 0; r < P.ntiles; ++r) {
                umma_f16_lohi(d, a_lo, b_lo, DESC_HI, IDESC, first);
                a_lo += TILE_LO;
                d += NST;
              }
              first = 1u;
            }
          }
          umma_commit(&empty_bar[s]);
          if (ks == nks - 1) umma_commit(&tmem_full[buf]);     // accumulators of this unit complete
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;
    const int et = threadIdx.x - 64;
    const bool temb_uniform = P.temb != nullptr && P.temb_bstride == 0;
    constexpr int LPR = BN / 4;                            // lanes per row in the store phase
    constexpr int RPI = 32 / LPR;                          // rows per store instruction
    constexpr int NIT = 32 / RPI;                          // store iterations per 32-row block
    float* tbuf = tbuf_all + static_cast<size_t>(quarter) * 32 * TLD;
    const int plane = P.HB * P.Wp;
    const int sub_r = lane / LPR, sub_c = (lane % LPR) * 4;
    int it = 0;
    bool alive = true;
    for (int u = blockIdx.x; u < P.n_units && alive; u += gridDim.x, ++it) {
      const PlaneUnit U = plane_unit(P, u);
      const int buf = it & 1;
      const int nn0 = U.n_tile * BN;
      asm volatile("bar.sync 1, 128;" ::: "memory");       // previous unit's colv no longer read
      {
        const int trow = (P.temb && P.t_dev) ? *P.t_dev : 0;
        for (int c = et; c < BN; c += 128) {
          const int nn = nn0 + c;
          float v = P.bias ? P.bias[nn] : 0.f;
          if (P.bias2) v += P.bias2[nn];
          if (temb_uniform) v += P.temb[static_cast<size_t>(trow) * P.temb_ld + nn];
          if (P.temb && !temb_uniform) v += P.temb[static_cast<size_t>(U.n) * P.temb_bstride + nn];
          colv[c] = v;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const float4 cv = *reinterpret_cast<const float4*>(colv + sub_c);
      if (!mbar_wait(&tmem_full[buf], (it >> 1) & 1, P.err_flag, 403)) { alive = false; break; }
      tc_fence_after();
      if (P.dbg & 64) {            // bring-up: no epilogue work at all
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        continue;
      }
#pragma unroll 1
      for (int r = 0; r < P.ntiles; ++r) {
        // global coordinates of this thread's rows of the store phase + residual prefetch
        const int q0 = r * 128 + quarter * 32;
        size_t mrow[NIT];
        float4 rv[NIT];
#pragma unroll
        for (int j = 0; j < NIT; ++j) {
          const int q = q0 + sub_r + j * RPI;
          mrow[j] = ~static_cast<size_t>(0);
          rv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (q < P.P) {
            const int dl = q / plane;
            const int rem = q - dl * plane;
            const int hl = rem / P.Wp;
            const int w = rem - hl * P.Wp;
            if (w < P.W && !(P.dbg & 8)) {
              mrow[j] = ((static_cast<size_t>(U.n) * P.D + U.d0 + dl) * P.H + U.h0 + hl) * P.W + w;
              if (P.resid) rv[j] = *reinterpret_cast<const float4*>(P.resid + mrow[j] * P.cout + nn0 + sub_c);
            }
          }
        }
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * buf_cols + r * NST;
#pragma unroll
        for (int c = 0; c < BN / 16; ++c) {
          float v[16];
          tmem_ld16(t_lane + c * 16, v);
          if (TERMS == 2) {
            float v2[16];
            tmem_ld16(t_lane + BN + c * 16, v2);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += v2[i];
          }
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(tbuf + lane * TLD + c * 16 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
        if (r == P.ntiles - 1) {
          // every accumulator of this buffer has been read by this warp: hand it back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < NIT; ++j) {
          if (mrow[j] == ~static_cast<size_t>(0)) continue;
          float4 v = *reinterpret_cast<const float4*>(tbuf + (sub_r + j * RPI) * TLD + sub_c);
          v.x += cv.x + rv[j].x; v.y += cv.y + rv[j].y; v.z += cv.z + rv[j].z; v.w += cv.w + rv[j].w;
          if (P.out32) *reinterpret_cast<float4*>(P.out32 + mrow[j] * P.out_ld + nn0 + sub_c) = v;
          if (P.out16) {
            __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
            uint2 uu;
            uu.x = *reinterpret_cast<uint32_t*>(&h0);
            uu.y = *reinterpret_cast<uint32_t*>(&h1);
            *reinterpret_cast<uint2*>(P.out16 + mrow[j] * P.out_ld + nn0 + sub_c) = uu;
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

struct PlaneLaunch {
  PlaneParams p;
  dim3 grid;
  int bn, bk;
  size_t smem;
  double flops;
  bool ok;               // false: geometry not covered, use conv_umma_kernel
};

// Fills L for a k3 s1 p1 conv (+ optional 1x1x1 extra source).  Returns 0 with L->ok = false when
// the geometry is outside what this kernel covers (caller falls back to conv_umma).
int plane_prepare(PlaneLaunch* L, const __half* act, int B, int D, int H, int W, int cin,
                  const __half* extra, int cin_extra, const __half* wpacked, int cout, int terms);
int plane_enqueue(const PlaneLaunch& L, cudaStream_t st);
int plane_init();

}  // namespace cm
