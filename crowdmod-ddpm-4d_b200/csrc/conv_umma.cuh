// Implicit-GEMM 3-D convolution on tcgen05 / TMEM fed by TMA (im2col mode), sm_100a.
//
// Replaces every nn.Conv3d the reference UNet dispatches to cuDNN except the two 3-channel
// ends (reference: models/backbones/layers.py:32,43,46,84,92-94; unet.py:32,121) and the two
// attention projections (layers.py:10,16 — a 1x1x1 "conv").
//
//   D[m, n] = sum_{tap, ci} A[pixel(m) + tap, ci] * W[n, tap, ci]      m = output pixel (NDHWC order)
//
// A (activations, fp16, channels-last [B, H, W, L, C]) is never materialised as an im2col
// matrix: each k-block is ONE cp.async.bulk.tensor.5d...im2col load of 128 consecutive output
// pixels x BK channels for one filter tap (hardware does the halo / zero padding / stride).
// W is pre-packed fp16 [terms*Cout, K] (K = taps*Cin, tap-major) and loaded with tiled TMA.
// Accumulation is fp32 in TMEM; the epilogue (4 warps, one TMEM lane quarter each) fuses
// bias, the time-embedding projection add, the residual add, fp32/fp16 stores and the
// nearest-x2-upsample scatter.
#pragma once
#include "common.cuh"

namespace cm {

constexpr int CONV_BM = 128;        // UMMA M (cta_group::1)
constexpr int CONV_THREADS = 192;   // warp0 TMA, warp1 MMA, warps2-5 epilogue
constexpr int CONV_MAX_STAGES = 6;

struct ConvParams {
  CUtensorMap amap[8];   // main source, one map per output phase (1 normally, 8 for upsample)
  CUtensorMap xmap;      // optional extra 1x1x1 source (match_input fused as a K-slab)
  CUtensorMap bmap;      // packed weights [terms*cout rows][Ktot], K-major
  int M;                 // rows of this launch (per phase)
  int od, oh, ow;        // GEMM-M traversal extents per sample: m = ((n*od+z)*oh+p)*ow+q
  int pps;               // od*oh*ow
  int conv_stride;       // traversal stride of the main source
  int kd, kh, kw;        // taps per dim of the main source
  int nphase;
  signed char lower[8][4];  // per phase lower corner {w,h,d}
  int cin_main, cin_extra;
  int kphase;            // packed-K elements per phase
  int terms;             // 1 = fp16 weights, 2 = hi+lo split weights
  int cout;
  int stages;
  const float* bias;
  const float* bias2;
  const float* temb;     // [rows][temb_ld] projection table, nullptr = none
  const int* t_dev;      // device int: row of `temb` for the whole batch (sampling); nullptr = 0
  int temb_ld;
  int temb_bstride;      // per-sample row stride (training: temb_ld, sampling: 0)
  const float* resid;    // fp32 [M][cout] or nullptr
  float* out32;          // fp32 [rows][out_ld] or nullptr
  __half* out16;         // fp16 [rows][out_ld] or nullptr
  int out_ld;
  int scatter;           // 1: rows are low-res pixels, written to (2z+pz, 2p+pp, 2q+pq)
  int* err_flag;
};

template <int BN, int BK>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_umma_kernel(const __grid_constant__ ConvParams P) {
  constexpr int ROWB = BK * 2;                  // bytes per smem row (swizzle span)
  constexpr int A_BYTES = CONV_BM * ROWB;
  constexpr int B_BYTES = BN * ROWB;
  constexpr uint32_t IDESC = make_idesc_f16(CONV_BM, BN);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  const int S = P.stages;
  const int terms = P.terms;
  const int stage_bytes = A_BYTES + terms * B_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * stage_bytes);
  uint64_t* empty_bar = full_bar + CONV_MAX_STAGES;
  uint64_t* tmem_full = empty_bar + CONV_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x;
  const int n_tile = blockIdx.y;
  const int phase = blockIdx.z;

  const int ncm = P.cin_main / BK;
  const int nkb_main = P.kd * P.kh * P.kw * ncm;
  const int nkb = nkb_main + P.cin_extra / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.amap[phase]);
    tma_prefetch_desc(&P.bmap);
    if (P.cin_extra) tma_prefetch_desc(&P.xmap);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int m0 = m_tile * CONV_BM;
      const int n0 = m0 / P.pps;
      int r = m0 - n0 * P.pps;
      const int z0 = r / (P.oh * P.ow);
      r -= z0 * P.oh * P.ow;
      const int p0 = r / P.ow;
      const int q0 = r - p0 * P.ow;
      const int w0 = q0 * P.conv_stride + P.lower[phase][0];
      const int h0 = p0 * P.conv_stride + P.lower[phase][1];
      const int d0 = z0 * P.conv_stride + P.lower[phase][2];
      const int kbase = phase * P.kphase;
      const uint32_t tx = A_BYTES + terms * B_BYTES;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (kb / S) & 1;
        if (!mbar_wait(&empty_bar[s], ph ^ 1, P.err_flag, 101)) break;
        uint8_t* sa = smem + s * stage_bytes;
        mbar_expect_tx(&full_bar[s], tx);
        if (kb < nkb_main) {
          const int tap = kb / ncm;
          const int cc = kb - tap * ncm;
          const int tw = tap % P.kw;
          const int th = (tap / P.kw) % P.kh;
          const int td = tap / (P.kw * P.kh);
          tma_load_im2col_5d(&P.amap[phase], &full_bar[s], sa, cc * BK, w0, h0, d0, n0,
                             (uint16_t)tw, (uint16_t)th, (uint16_t)td);
        } else {
          const int cc = kb - nkb_main;
          tma_load_im2col_5d(&P.xmap, &full_bar[s], sa, cc * BK, q0, p0, z0, n0, 0, 0, 0);
        }
        for (int t = 0; t < terms; ++t)
          tma_load_2d(&P.bmap, &full_bar[s], sa + A_BYTES + t * B_BYTES, kbase + kb * BK,
                      n_tile * BN + t * P.cout);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      bool ok = true;
      for (int kb = 0; kb < nkb && ok; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (kb / S) & 1;
        ok = mbar_wait(&full_bar[s], ph, P.err_flag, 102);
        if (!ok) break;
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + s * stage_bytes);
        const uint32_t b_base = a_base + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t adesc = make_kmajor_desc(a_base + k * 32, ROWB);
          for (int t = 0; t < terms; ++t) {
            const uint64_t bdesc = make_kmajor_desc(b_base + t * B_BYTES + k * 32, ROWB);
            umma_f16(tmem_base, adesc, bdesc, IDESC, (kb | k | t) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[s]);   // frees the smem slot once these MMAs retire
      }
      umma_commit(tmem_full);         // accumulator complete
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;     // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const int m = m_tile * CONV_BM + row;
    const bool valid = m < P.M;
    mbar_wait(tmem_full, 0, P.err_flag, 103);
    tc_fence_after();

    int b = 0;
    size_t orow = 0;
    if (valid) {
      b = m / P.pps;
      if (P.scatter) {
        int r = m - b * P.pps;
        const int z = r / (P.oh * P.ow);
        r -= z * P.oh * P.ow;
        const int p = r / P.ow;
        const int q = r - p * P.ow;
        const int pq = phase & 1, pp = (phase >> 1) & 1, pz = (phase >> 2) & 1;
        orow = ((static_cast<size_t>(b) * (2 * P.od) + (2 * z + pz)) * (2 * P.oh) + (2 * p + pp)) *
                   (2 * P.ow) + (2 * q + pq);
      } else {
        orow = static_cast<size_t>(m);
      }
    }
    const float* temb_row = nullptr;
    if (P.temb) {
      const int trow = P.t_dev ? *P.t_dev : 0;
      temb_row = P.temb + static_cast<size_t>(trow) * P.temb_ld +
                 static_cast<size_t>(b) * P.temb_bstride;
    }
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN / 16; ++c) {
      float v[16];
      tmem_ld16(t_lane + c * 16, v);
      if (!valid) continue;
      const int n = n_tile * BN + c * 16;
      if (P.bias) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 t4 = *reinterpret_cast<const float4*>(P.bias + n + i);
          v[i] += t4.x; v[i + 1] += t4.y; v[i + 2] += t4.z; v[i + 3] += t4.w;
        }
      }
      if (P.bias2) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 t4 = *reinterpret_cast<const float4*>(P.bias2 + n + i);
          v[i] += t4.x; v[i + 1] += t4.y; v[i + 2] += t4.z; v[i + 3] += t4.w;
        }
      }
      if (temb_row) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 t4 = *reinterpret_cast<const float4*>(temb_row + n + i);
          v[i] += t4.x; v[i + 1] += t4.y; v[i + 2] += t4.z; v[i + 3] += t4.w;
        }
      }
      if (P.resid) {
        const float* rp = P.resid + static_cast<size_t>(m) * P.cout + n;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 t4 = *reinterpret_cast<const float4*>(rp + i);
          v[i] += t4.x; v[i + 1] += t4.y; v[i + 2] += t4.z; v[i + 3] += t4.w;
        }
      }
      if (P.out32) {
        float* op = P.out32 + orow * P.out_ld + n;
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(op + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
      if (P.out16) {
        __half* op = P.out16 + orow * P.out_ld + n;
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          __half2 h0 = __floats2half2_rn(v[i], v[i + 1]);
          __half2 h1 = __floats2half2_rn(v[i + 2], v[i + 3]);
          __half2 h2 = __floats2half2_rn(v[i + 4], v[i + 5]);
          __half2 h3 = __floats2half2_rn(v[i + 6], v[i + 7]);
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
  if (warp == 2) tmem_dealloc(tmem_base, BN);
}

// -------------------------------- host side --------------------------------
struct ConvGeom {
  // source activation tensor (fp16, channels-last)
  int B, D, H, W;       // D=grid rows, H=grid cols, W=time frames (fastest spatial)
  int C;                // channels of the main source
};

struct ConvLaunch {
  ConvParams p;
  dim3 grid;
  int bn, bk;
  size_t smem;
  double flops;         // algorithmic 2*M*N*K of this launch (all phases)
};

// Fill amap/xmap/bmap + geometry.  `mode`: 0 = k3 s1 p1 ("same"), 1 = k3 s2 p1 (DownSample),
// 2 = nearest-x2 + k3 p1 decomposed into 8 phase convs with 2x2x2 combined taps (UpSample),
// 3 = 1x1x1.  Extra source: 1x1x1 over the OUTPUT grid with cin_extra channels.
int conv_prepare(ConvLaunch* L, int mode, const __half* act, int B, int D, int H, int W, int cin,
                 const __half* extra, int cin_extra, const __half* wpacked, int cout, int terms);
int conv_enqueue(const ConvLaunch& L, cudaStream_t st);
size_t conv_packed_k(int mode, int cin, int cin_extra);   // K elements of the packed weight rows

}  // namespace cm
