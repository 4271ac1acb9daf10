// Bandwidth / small-op kernels of the UNet hot path.  See kernels.cuh for the contracts and
// the reference lines each one replaces.
#include "kernels.cuh"
#include "conv_umma.cuh"

namespace cm {

// =============================================================================================
// weight packing
// =============================================================================================
__global__ void pack_conv_weights_kernel(const float* __restrict__ w, const float* __restrict__ wx,
                                         __half* __restrict__ dst, int cout, int cin, int cinx,
                                         int taps, int terms) {
  const size_t ktot = (size_t)taps * cin + cinx;
  const size_t total = (size_t)cout * ktot;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx / ktot);
    const size_t k = idx - (size_t)n * ktot;
    float v;
    if (k < (size_t)taps * cin) {
      const int tap = (int)(k / cin);
      const int ci = (int)(k - (size_t)tap * cin);
      v = w[((size_t)n * cin + ci) * taps + tap];
    } else {
      v = wx[(size_t)n * cinx + (k - (size_t)taps * cin)];
    }
    const __half hi = __float2half_rn(v);
    dst[(size_t)n * ktot + k] = hi;
    if (terms == 2) dst[((size_t)cout + n) * ktot + k] = __float2half_rn(v - __half2float(hi));
  }
}

int pack_conv_weights(const float* w, const float* wx, __half* dst, int cout, int cin, int cinx,
                      int taps, int terms, cudaStream_t st) {
  const size_t total = (size_t)cout * ((size_t)taps * cin + cinx);
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  pack_conv_weights_kernel<<<blocks, 256, 0, st>>>(w, wx, dst, cout, cin, cinx, taps, terms);
  CM_CUDA(cudaGetLastError());
  return 0;
}

__global__ void pack_upsample_weights_kernel(const float* __restrict__ w, __half* __restrict__ dst,
                                             int cout, int cin, int terms) {
  const size_t ktot = (size_t)64 * cin;
  const size_t total = (size_t)cout * ktot;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx / ktot);
    const int k = (int)(idx - (size_t)n * ktot);
    const int phase = k / (8 * cin);
    const int r = k - phase * 8 * cin;
    const int tap8 = r / cin;
    const int ci = r - tap8 * cin;
    // per dim: phase bit p, tap bit a -> contributing original taps [lo, hi]
    //   p=0 (even output): a=0 -> {0},   a=1 -> {1,2}
    //   p=1 (odd  output): a=0 -> {0,1}, a=1 -> {2}
    int lo[3], hi[3];
    const int pbit[3] = {phase & 1, (phase >> 1) & 1, (phase >> 2) & 1};   // w, h, d
    const int abit[3] = {tap8 & 1, (tap8 >> 1) & 1, (tap8 >> 2) & 1};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (pbit[d] == 0) { lo[d] = abit[d] ? 1 : 0; hi[d] = abit[d] ? 2 : 0; }
      else              { lo[d] = abit[d] ? 2 : 0; hi[d] = abit[d] ? 2 : 1; }
    }
    const float* wp = w + ((size_t)n * cin + ci) * 27;
    float v = 0.f;
    for (int kd = lo[2]; kd <= hi[2]; ++kd)
      for (int kh = lo[1]; kh <= hi[1]; ++kh)
        for (int kw = lo[0]; kw <= hi[0]; ++kw) v += wp[(kd * 3 + kh) * 3 + kw];
    const __half h = __float2half_rn(v);
    dst[(size_t)n * ktot + k] = h;
    if (terms == 2) dst[((size_t)cout + n) * ktot + k] = __float2half_rn(v - __half2float(h));
  }
}

int pack_upsample_weights(const float* w, __half* dst, int cout, int cin, int terms,
                          cudaStream_t st) {
  const size_t total = (size_t)cout * 64 * cin;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  pack_upsample_weights_kernel<<<blocks, 256, 0, st>>>(w, dst, cout, cin, terms);
  CM_CUDA(cudaGetLastError());
  return 0;
}

__global__ void cast_f32_to_f16_kernel(const float* __restrict__ s, __half* __restrict__ d, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    d[i] = __float2half_rn(s[i]);
}
int cast_f32_to_f16(const float* src, __half* dst, size_t n, cudaStream_t st) {
  const int blocks = (int)((n + 255) / 256 < 8192 ? (n + 255) / 256 : 8192);
  cast_f32_to_f16_kernel<<<blocks ? blocks : 1, 256, 0, st>>>(src, dst, n);
  CM_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// GroupNorm(8) + SiLU -> fp16
// One CTA per (sample, group).  Exact two-pass statistics (mean, then centred second moment)
// with the slab cached in registers (first GN_CACHE float4 per thread; the tail, if any, is
// re-read from L2).
// =============================================================================================
constexpr int GN_CACHE = 8;

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();   // protect `red` reuse
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;   // every thread holds the block total
}

__global__ void __launch_bounds__(1024) gn_silu_kernel(const GnParams p) {
  __shared__ float red[32];
  const int b = blockIdx.x >> 3, g = blockIdx.x & 7;
  const int C = p.c0 + p.c1;
  const int cg = C >> 3;
  const int vpp = cg >> 2;                       // float4 per pixel of this group
  const int nvec = p.pixels * vpp;
  const int nt = blockDim.x;
  const size_t pix0 = (size_t)b * p.pixels;

  auto load = [&](int i) -> float4 {
    const int px = i / vpp;
    const int c = g * cg + (i - px * vpp) * 4;
    const float* src = (c < p.c0) ? p.src0 + (pix0 + px) * p.c0 + c
                                  : p.src1 + (pix0 + px) * p.c1 + (c - p.c0);
    return *reinterpret_cast<const float4*>(src);
  };

  float4 cache[GN_CACHE];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < GN_CACHE; ++j) {
    const int i = threadIdx.x + j * nt;
    if (i < nvec) {
      cache[j] = load(i);
      s += (cache[j].x + cache[j].y) + (cache[j].z + cache[j].w);
    }
  }
  for (int i = threadIdx.x + GN_CACHE * nt; i < nvec; i += nt) {
    const float4 v = load(i);
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float inv_n = 1.0f / (float)(nvec * 4);
  const float mean = block_sum(s, red) * inv_n;

  float q = 0.f;
#pragma unroll
  for (int j = 0; j < GN_CACHE; ++j) {
    const int i = threadIdx.x + j * nt;
    if (i < nvec) {
      const float dx = cache[j].x - mean, dy = cache[j].y - mean, dz = cache[j].z - mean,
                  dw = cache[j].w - mean;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
  }
  for (int i = threadIdx.x + GN_CACHE * nt; i < nvec; i += nt) {
    const float4 v = load(i);
    const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
    q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
  }
  const float var = block_sum(q, red) * inv_n;
  const float rstd = 1.0f / sqrtf(var + p.eps);
  if (p.stats && threadIdx.x == 0) {
    p.stats[(b * 8 + g) * 2 + 0] = mean;
    p.stats[(b * 8 + g) * 2 + 1] = rstd;
  }

  auto emit = [&](int i, const float4& v) {
    const int px = i / vpp;
    const int c = g * cg + (i - px * vpp) * 4;
    const float4 ga = *reinterpret_cast<const float4*>(p.gamma + c);
    const float4 be = *reinterpret_cast<const float4*>(p.beta + c);
    float y0 = (v.x - mean) * rstd * ga.x + be.x;
    float y1 = (v.y - mean) * rstd * ga.y + be.y;
    float y2 = (v.z - mean) * rstd * ga.z + be.z;
    float y3 = (v.w - mean) * rstd * ga.w + be.w;
    if (p.silu) { y0 = silu_f(y0); y1 = silu_f(y1); y2 = silu_f(y2); y3 = silu_f(y3); }
    if (p.drop_scale) {
      const float4 ds = *reinterpret_cast<const float4*>(p.drop_scale + (size_t)b * C + c);
      y0 *= ds.x; y1 *= ds.y; y2 *= ds.z; y3 *= ds.w;
    }
    const size_t o = (pix0 + px) * C + c;
    __half2 h0 = __floats2half2_rn(y0, y1), h1 = __floats2half2_rn(y2, y3);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(p.out_norm + o) = u;
    if (p.out_raw) {
      __half2 r0 = __floats2half2_rn(v.x, v.y), r1 = __floats2half2_rn(v.z, v.w);
      u.x = *reinterpret_cast<uint32_t*>(&r0);
      u.y = *reinterpret_cast<uint32_t*>(&r1);
      *reinterpret_cast<uint2*>(p.out_raw + o) = u;
    }
  };
#pragma unroll
  for (int j = 0; j < GN_CACHE; ++j) {
    const int i = threadIdx.x + j * nt;
    if (i < nvec) emit(i, cache[j]);
  }
  for (int i = threadIdx.x + GN_CACHE * nt; i < nvec; i += nt) emit(i, load(i));
}

int gn_silu_enqueue(const GnParams& p, cudaStream_t st) {
  const int C = p.c0 + p.c1;
  CM_CHECK(C % 32 == 0 && p.c0 % 4 == 0, "GroupNorm channels must be a multiple of 32 (C=%d)", C);
  const int nvec = p.pixels * (C / 32);
  int threads = ((nvec + GN_CACHE - 1) / GN_CACHE + 31) / 32 * 32;
  if (threads < 64) threads = 64;
  if (threads > 1024) threads = 1024;
  gn_silu_kernel<<<p.B * 8, threads, 0, st>>>(p);
  CM_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// first conv (K = 27*cin is tiny; memory-bound): one thread per output pixel, all couts.
// =============================================================================================
template <int CIN>
__global__ void __launch_bounds__(128)
first_conv_kernel(const float* __restrict__ x, const float* __restrict__ past,
                  const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ out, int B, int H, int W, int P, int F, int cout) {
  extern __shared__ float ws[];   // [27*CIN][cout], k = ci*27 + tap
  constexpr int K = 27 * CIN;
  for (int idx = threadIdx.x; idx < K * cout; idx += blockDim.x) {
    const int k = idx / cout, co = idx - k * cout;
    ws[idx] = w[(size_t)co * K + k];
  }
  __syncthreads();
  const int L = P + F;
  const size_t total = (size_t)B * H * W * L;
  const size_t m = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (m >= total) return;
  int l = (int)(m % L);
  size_t r = m / L;
  const int wc = (int)(r % W);
  r /= W;
  const int h = (int)(r % H);
  const int b = (int)(r / H);

  float in[K];
#pragma unroll
  for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
    for (int td = 0; td < 3; ++td) {
#pragma unroll
      for (int th = 0; th < 3; ++th) {
#pragma unroll
        for (int tw = 0; tw < 3; ++tw) {
          const int hh = h + td - 1, ww = wc + th - 1, ll = l + tw - 1;
          float v = 0.f;
          if (hh >= 0 && hh < H && ww >= 0 && ww < W && ll >= 0 && ll < L) {
            const size_t plane = ((size_t)(b * CIN + ci) * H + hh) * W + ww;
            v = (ll < P) ? past[plane * P + ll] : x[plane * F + (ll - P)];
          }
          in[ci * 27 + (td * 3 + th) * 3 + tw] = v;
        }
      }
    }
  }
  for (int co0 = 0; co0 < cout; co0 += 32) {
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = bias[co0 + j];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float a = in[k];
      const float4* wr = reinterpret_cast<const float4*>(ws + k * cout + co0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 w4 = wr[j];
        acc[4 * j + 0] = fmaf(a, w4.x, acc[4 * j + 0]);
        acc[4 * j + 1] = fmaf(a, w4.y, acc[4 * j + 1]);
        acc[4 * j + 2] = fmaf(a, w4.z, acc[4 * j + 2]);
        acc[4 * j + 3] = fmaf(a, w4.w, acc[4 * j + 3]);
      }
    }
    float4* op = reinterpret_cast<float4*>(out + m * cout + co0);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      op[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
  }
}

int first_conv_enqueue(const float* x, const float* past, const float* w, const float* bias,
                       float* out, int B, int H, int W, int P, int F, int cin, int cout,
                       cudaStream_t st) {
  CM_CHECK(cin >= 1 && cin <= 4, "first conv supports 1..4 input channels (got %d)", cin);
  CM_CHECK(cout % 32 == 0, "first conv cout must be a multiple of 32");
  const size_t total = (size_t)B * H * W * (P + F);
  const int blocks = (int)((total + 127) / 128);
  const size_t smem = (size_t)27 * cin * cout * sizeof(float);
#define CM_FIRST(CI)                                                                           \
  case CI: {                                                                                   \
    first_conv_kernel<CI><<<blocks, 128, smem, st>>>(x, past, w, bias, out, B, H, W, P, F, cout); \
  } break;
  switch (cin) {
    CM_FIRST(1) CM_FIRST(2) CM_FIRST(3) CM_FIRST(4)
  }
#undef CM_FIRST
  CM_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// final conv (N = 3 is tiny; memory-bound) + reverse-step update.
// 4 lanes per output pixel, each lane owns 8-channel slices; only future frames are produced.
// =============================================================================================
template <int COUT>
__global__ void __launch_bounds__(256) final_conv_kernel(const FinalParams p) {
  extern __shared__ float ws[];   // [27][cin][COUT]
  const int cin = p.cin;
  for (int idx = threadIdx.x; idx < 27 * cin * COUT; idx += blockDim.x) {
    const int co = idx % COUT;
    const int r = idx / COUT;
    const int ci = r % cin, tap = r / cin;
    ws[idx] = p.w[((size_t)co * cin + ci) * 27 + tap];
  }
  __syncthreads();
  const int F = p.L - p.P;
  const size_t total = (size_t)p.B * p.H * p.W * F;
  const size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t pix = gid >> 2;
  const int sub = (int)(gid & 3);
  const bool active = pix < total;
  int f = 0, wc = 0, h = 0, b = 0;
  if (active) {
    f = (int)(pix % F);
    size_t r = pix / F;
    wc = (int)(r % p.W);
    r /= p.W;
    h = (int)(r % p.H);
    b = (int)(r / p.H);
  }
  const int l = p.P + f;
  float acc[COUT];
#pragma unroll
  for (int co = 0; co < COUT; ++co) acc[co] = 0.f;
  if (active) {
    for (int td = 0; td < 3; ++td) {
      const int hh = h + td - 1;
      if (hh < 0 || hh >= p.H) continue;
      for (int th = 0; th < 3; ++th) {
        const int ww = wc + th - 1;
        if (ww < 0 || ww >= p.W) continue;
        for (int tw = 0; tw < 3; ++tw) {
          const int ll = l + tw - 1;
          if (ll < 0 || ll >= p.L) continue;
          const int tap = (td * 3 + th) * 3 + tw;
          const __half* ap = p.act + ((((size_t)b * p.H + hh) * p.W + ww) * p.L + ll) * cin;
          for (int c0 = sub * 8; c0 < cin; c0 += 32) {
            const uint4 u = *reinterpret_cast<const uint4*>(ap + c0);
            const __half2* h2 = reinterpret_cast<const __half2*>(&u);
            const float* wp = ws + ((size_t)tap * cin + c0) * COUT;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 a2 = __half22float2(h2[e]);
#pragma unroll
              for (int co = 0; co < COUT; ++co) {
                acc[co] = fmaf(a2.x, wp[(2 * e) * COUT + co], acc[co]);
                acc[co] = fmaf(a2.y, wp[(2 * e + 1) * COUT + co], acc[co]);
              }
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int co = 0; co < COUT; ++co) {
    acc[co] += __shfl_xor_sync(0xffffffffu, acc[co], 1);
    acc[co] += __shfl_xor_sync(0xffffffffu, acc[co], 2);
  }
  if (!active || sub != 0) return;

  const size_t plane = (size_t)p.H * p.W * F;                       // elements per (b, c)
  const size_t e0 = ((size_t)b * COUT) * plane + ((size_t)h * p.W + wc) * F + f;
  float eps[COUT];
#pragma unroll
  for (int co = 0; co < COUT; ++co) eps[co] = acc[co] + p.bias[co];
  if (p.eps_out) {
#pragma unroll
    for (int co = 0; co < COUT; ++co) p.eps_out[e0 + co * plane] = eps[co];
  }
  if (!p.x) return;

  const int step = *p.step_dev;
  const float* cf = p.coef + (size_t)step * 8;
  const size_t nelem = (size_t)p.B * COUT * plane;
  float z[4] = {0.f, 0.f, 0.f, 0.f};
  const float zc = (p.mode == 0) ? cf[2] : cf[4];
  if (zc != 0.f) {
    if (p.noise) {
#pragma unroll
      for (int co = 0; co < COUT; ++co) z[co] = p.noise[(size_t)step * nelem + e0 + co * plane];
    } else {
      const unsigned long long gp =
          ((unsigned long long)(p.sample_offset + b)) * (unsigned long long)(p.H * p.W * F) +
          ((size_t)h * p.W + wc) * F + f;
      const uint4 rnd = philox4x32_10(
          make_uint4((uint32_t)gp, (uint32_t)(gp >> 32), (uint32_t)step, 0u),
          make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
      const float2 n0 = box_muller(rnd.x, rnd.y), n1 = box_muller(rnd.z, rnd.w);
      z[0] = n0.x; z[1] = n0.y; z[2] = n1.x; z[3] = n1.y;
    }
  }
#pragma unroll
  for (int co = 0; co < COUT; ++co) {
    const size_t e = e0 + co * plane;
    const float xv = p.x[e];
    float xn;
    if (p.mode == 0) {
      // ddpm.py:33-37: one_by_sqrt_alpha * (x - (beta/sqrt(1-abar)) * eps) + sqrt(beta) * z
      xn = cf[0] * (xv - cf[1] * eps[co]) + cf[2] * z[co];
    } else {
      // ddpm.py:262-265 (DDIM eq. 12)
      const float x0 = (xv - cf[0] * eps[co]) / cf[1];
      xn = cf[2] * x0 + cf[3] * eps[co] + cf[4] * z[co];
    }
    if (co == 0 && cf[5] != 0.f) {
      // guidance.py:4-8 + ddpm.py:223-226: x[:,0] -= lambda*sigma*sign(x[:,0])
      const float sg = (xn > 0.f) ? 1.f : ((xn < 0.f) ? -1.f : 0.f);
      xn -= cf[5] * sg;
    }
    p.x[e] = xn;
    if (p.history) p.history[(size_t)(step + 1) * nelem + e] = xn;
  }
}

int final_conv_enqueue(const FinalParams& p, cudaStream_t st) {
  CM_CHECK(p.cout >= 1 && p.cout <= 4, "final conv supports 1..4 output channels (got %d)", p.cout);
  CM_CHECK(p.cin % 32 == 0, "final conv cin must be a multiple of 32");
  const size_t total = (size_t)p.B * p.H * p.W * (p.L - p.P) * 4;
  const int blocks = (int)((total + 255) / 256);
  const size_t smem = (size_t)27 * p.cin * p.cout * sizeof(float);
#define CM_FINAL(CO)                                                                           \
  case CO: {                                                                                   \
    final_conv_kernel<CO><<<blocks, 256, smem, st>>>(p);                                       \
  } break;
  switch (p.cout) {
    CM_FINAL(1) CM_FINAL(2) CM_FINAL(3) CM_FINAL(4)
  }
#undef CM_FINAL
  CM_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// time embedding: table row -> Linear -> SiLU -> Linear -> (SiLU -> dense_1 of every block)
// =============================================================================================
__global__ void __launch_bounds__(256) temb_kernel(const TembParams p) {
  extern __shared__ float sm[];   // e[base] | s1[E] | s2[E]
  float* e = sm;
  float* s1 = sm + p.base;
  float* s2 = s1 + p.E;
  const int row = blockIdx.x;
  const long long t = p.t ? p.t[row] : (long long)row;
  for (int i = threadIdx.x; i < p.base; i += blockDim.x) e[i] = p.table[(size_t)t * p.base + i];
  __syncthreads();
  for (int j = threadIdx.x; j < p.E; j += blockDim.x) {
    float a = p.b1[j];
    const float* wr = p.w1 + (size_t)j * p.base;
    for (int i = 0; i < p.base; ++i) a = fmaf(wr[i], e[i], a);
    s1[j] = silu_f(a);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < p.E; j += blockDim.x) {
    float a = p.b2[j];
    const float* wr = p.w2 + (size_t)j * p.E;
    for (int i = 0; i < p.E; ++i) a = fmaf(wr[i], s1[i], a);
    s2[j] = silu_f(a);   // every consumer applies SiLU first (layers.py:62)
  }
  __syncthreads();
  for (int k = 0; k < p.nblocks; ++k) {
    const float* wd = p.wd[k];
    const float* bd = p.bd[k];
    const int co_n = p.couts[k];
    float* o = p.out + (size_t)row * p.ld + p.offs[k];
    for (int c = threadIdx.x; c < co_n; c += blockDim.x) {
      float a = bd[c];
      const float* wr = wd + (size_t)c * p.E;
      for (int i = 0; i < p.E; ++i) a = fmaf(wr[i], s2[i], a);
      o[c] = a;
    }
  }
}

int temb_enqueue(const TembParams& p, cudaStream_t st) {
  const size_t smem = (size_t)(p.base + 2 * p.E) * sizeof(float);
  temb_kernel<<<p.rows, 256, smem, st>>>(p);
  CM_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// attention core (S <= a few hundred tokens: whole K/V of one (sample, head) lives in smem)
// =============================================================================================
__global__ void __launch_bounds__(128)
attn_core_kernel(const float* __restrict__ qkv, __half* __restrict__ ctx, int S, int C, int heads) {
  extern __shared__ float sm[];
  const int dh = C / heads;
  const int b = blockIdx.x / heads, hd = blockIdx.x % heads;
  float* Ks = sm;                       // [S][dh+1]
  float* Vs = Ks + (size_t)S * (dh + 1);   // [S][dh]
  float* Qs = Vs + (size_t)S * dh;      // [4][dh]
  float* Pw = Qs + 4 * dh;              // [4][S]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* base = qkv + (size_t)b * S * 3 * C;
  for (int idx = threadIdx.x; idx < S * dh; idx += blockDim.x) {
    const int j = idx / dh, d = idx - j * dh;
    Ks[j * (dh + 1) + d] = base[(size_t)j * 3 * C + C + hd * dh + d];
    Vs[j * dh + d] = base[(size_t)j * 3 * C + 2 * C + hd * dh + d];
  }
  __syncthreads();
  const float scale = rsqrtf((float)dh);
  float* q = Qs + warp * dh;
  float* pw = Pw + warp * S;
  for (int i = warp; i < S; i += 4) {
    for (int d = lane; d < dh; d += 32) q[d] = base[(size_t)i * 3 * C + hd * dh + d] * scale;
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) {
      const float* kr = Ks + j * (dh + 1);
      float a = 0.f;
      for (int d = 0; d < dh; ++d) a = fmaf(q[d], kr[d], a);
      pw[j] = a;
      mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) {
      const float ev = expf(pw[j] - mx);
      pw[j] = ev;
      sum += ev;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.0f / sum;
    for (int d = lane; d < dh; d += 32) {
      float a = 0.f;
      for (int j = 0; j < S; ++j) a = fmaf(pw[j], Vs[j * dh + d], a);
      ctx[((size_t)b * S + i) * C + hd * dh + d] = __float2half_rn(a * inv);
    }
    __syncwarp();
  }
}

int attn_core_enqueue(const float* qkv, __half* ctx, int B, int S, int C, int heads,
                      cudaStream_t st) {
  CM_CHECK(C % heads == 0, "embed dim %d not divisible by heads %d", C, heads);
  const int dh = C / heads;
  const size_t smem = ((size_t)S * (dh + 1) + (size_t)S * dh + 4 * dh + 4 * (size_t)S) * sizeof(float);
  CM_CHECK(smem <= 200 * 1024, "attention tile too large for shared memory (S=%d dh=%d)", S, dh);
  attn_core_kernel<<<B * heads, 128, smem, st>>>(qkv, ctx, S, C, heads);
  CM_CUDA(cudaGetLastError());
  return 0;
}

int kernels_init() {
  static bool done = false;
  if (done) return 0;
  const int big = 160 * 1024;
  CM_CUDA(cudaFuncSetAttribute(first_conv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CM_CUDA(cudaFuncSetAttribute(first_conv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CM_CUDA(cudaFuncSetAttribute(first_conv_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CM_CUDA(cudaFuncSetAttribute(first_conv_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CM_CUDA(cudaFuncSetAttribute(final_conv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CM_CUDA(cudaFuncSetAttribute(final_conv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CM_CUDA(cudaFuncSetAttribute(final_conv_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CM_CUDA(cudaFuncSetAttribute(final_conv_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CM_CUDA(cudaFuncSetAttribute(attn_core_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  if (int rc = conv_init()) return rc;
  done = true;
  return 0;
}

// =============================================================================================
// chain bookkeeping
// =============================================================================================
__global__ void advance_step_kernel(int* step_dev, int* t_dev, const int* tsteps, int nsteps) {
  const int s = *step_dev + 1;
  *step_dev = s;
  *t_dev = tsteps[s < nsteps ? s : nsteps - 1];
}
int advance_step_enqueue(int* step_dev, int* t_dev, const int* tsteps, int nsteps, cudaStream_t st) {
  advance_step_kernel<<<1, 1, 0, st>>>(step_dev, t_dev, tsteps, nsteps);
  CM_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// test-only scalar restatement of conv_umma (same packed fp16 operands, fp32 accumulate)
// =============================================================================================
__global__ void conv_ref_kernel(const ConvParams p, const __half* __restrict__ act,
                                const __half* __restrict__ extra, const __half* __restrict__ wp,
                                int B, int D, int H, int W) {
  const size_t total = (size_t)p.nphase * p.M * p.cout;
  const int taps = p.kd * p.kh * p.kw;
  const size_t ktot = (p.nphase > 1) ? (size_t)p.nphase * p.kphase
                                     : (size_t)taps * p.cin_main + p.cin_extra;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx % p.cout);
    size_t r = idx / p.cout;
    const int m = (int)(r % p.M);
    const int phase = (int)(r / p.M);
    const int b = m / p.pps;
    int rr = m - b * p.pps;
    const int z = rr / (p.oh * p.ow);
    rr -= z * p.oh * p.ow;
    const int pp_ = rr / p.ow;
    const int q = rr - pp_ * p.ow;
    float acc = 0.f;
    for (int t = 0; t < p.terms; ++t) {
      const __half* wrow = wp + ((size_t)t * p.cout + n) * ktot + (size_t)phase * p.kphase;
      for (int tap = 0; tap < taps; ++tap) {
        const int tw = tap % p.kw, th = (tap / p.kw) % p.kh, td = tap / (p.kw * p.kh);
        const int w = q * p.conv_stride + p.lower[phase][0] + tw;
        const int h = pp_ * p.conv_stride + p.lower[phase][1] + th;
        const int d = z * p.conv_stride + p.lower[phase][2] + td;
        if (w < 0 || w >= W || h < 0 || h >= H || d < 0 || d >= D) continue;
        const __half* ap = act + ((((size_t)b * D + d) * H + h) * W + w) * p.cin_main;
        const __half* wk = wrow + (size_t)tap * p.cin_main;
        for (int c = 0; c < p.cin_main; ++c) acc = fmaf(__half2float(ap[c]), __half2float(wk[c]), acc);
      }
      if (p.cin_extra) {
        const __half* ap = extra + (size_t)m * p.cin_extra;
        const __half* wk = wrow + (size_t)taps * p.cin_main;
        for (int c = 0; c < p.cin_extra; ++c) acc = fmaf(__half2float(ap[c]), __half2float(wk[c]), acc);
      }
    }
    if (p.bias) acc += p.bias[n];
    if (p.bias2) acc += p.bias2[n];
    if (p.temb) {
      const int trow = p.t_dev ? *p.t_dev : 0;
      acc += p.temb[(size_t)trow * p.temb_ld + (size_t)b * p.temb_bstride + n];
    }
    if (p.resid) acc += p.resid[(size_t)m * p.cout + n];
    size_t orow = m;
    if (p.scatter) {
      const int pq = phase & 1, pp = (phase >> 1) & 1, pz = (phase >> 2) & 1;
      orow = (((size_t)b * (2 * p.od) + (2 * z + pz)) * (2 * p.oh) + (2 * pp_ + pp)) * (2 * p.ow) +
             (2 * q + pq);
    }
    if (p.out32) p.out32[orow * p.out_ld + n] = acc;
    if (p.out16) p.out16[orow * p.out_ld + n] = __float2half_rn(acc);
  }
}

int conv_ref_enqueue(const ConvParams& p, const __half* act, const __half* extra,
                     const __half* wpacked, int B, int D, int H, int W, cudaStream_t st) {
  conv_ref_kernel<<<1024, 256, 0, st>>>(p, act, extra, wpacked, B, D, H, W);
  CM_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace cm
