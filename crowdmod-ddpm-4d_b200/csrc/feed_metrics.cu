// The two callers either side of the hot path that SURVEY.md section 8 (f3, f4) marks "next":
//   * cm_window_gather  -- the data feed.  The reference keeps the raw sequences [N, C, ROWS, COLS, RAW_SEQ_LEN] on the
//     host, cuts PAST_LEN + FUTURE_LEN windows per sample in worker processes (utils/dataset.py:46-53) and copies every
//     batch to the GPU (models/diffusion/ddpm.py:136-137).  Here the sequences are resident in HBM and a batch is one
//     gather launch: HBM-bound index work, algorithmic bytes = B * C*ROWS*COLS * (P + F) * 4 * 2 (read + write).
//   * cm_metrics_reduce -- the reduction part of the metrics tail.  The reference loops over samples calling
//     .cpu().numpy() and numpy reductions per frame (utils/metrics/metricsGenerator.py:43-92,120-186,293-339); here one
//     launch reads prediction and ground truth once (algorithmic bytes = 2 * n * 3*ROWS*COLS*F * 4) and emits per
//     (sample, frame) the sums every one of those metrics is a closed form of.  fp64 accumulation, fixed order.
#include "../../include/crowdmod_b200.h"
#include "common.cuh"

namespace cm {
namespace {

// grid = (ceil(chw / 256), batch): one thread per (window, channel-row-col) copies the P + F consecutive frames of
// its cell -- no index divisions; reads are 4*(P+F)-byte runs (whole 32-byte sectors for the ATC 5 + 3 window), the
// past / future writes of a warp are contiguous
__global__ void __launch_bounds__(256) window_gather_kernel(const float* __restrict__ seq, int chw, int T,
                                                            const int* __restrict__ seq_idx, const int* __restrict__ t0,
                                                            const long long* __restrict__ ids, int P, int F,
                                                            float* __restrict__ past, float* __restrict__ future) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= chw) return;
  const long long wdw = ids ? ids[b] : b;          // window number -> (sequence, first frame) through the index table
  const float* src = seq + (static_cast<size_t>(seq_idx[wdw]) * chw + p) * T + t0[wdw];
  float* dp = past + (static_cast<size_t>(b) * chw + p) * P;
  float* df = future + (static_cast<size_t>(b) * chw + p) * F;
  for (int l = 0; l < P; ++l) dp[l] = __ldg(src + l);
  for (int l = 0; l < F; ++l) df[l] = __ldg(src + P + l);
}

constexpr int MET_THREADS = 128;
constexpr int MET_VALS = CM_METRICS_PER_FRAME;      // 21 doubles per (sample, frame)
constexpr int MET_SUMS = 15;                        // [0..14] are sums, [15..20] min / max pairs

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One CTA per sample.  The sample's prediction and ground truth (3 properties x rows x cols x F floats each, 15.5 KB for
// ATC) are staged in shared memory with coalesced 128-bit loads -- every byte is read from HBM once -- and the F
// frames are then reduced one after the other from shared memory (frame-strided reads, the two TV neighbours).
// fp64 sums (float32 differences / squares, as numpy forms them), float min / max (exact), fixed reduction order.
__global__ void __launch_bounds__(MET_THREADS) metrics_reduce_kernel(const float* __restrict__ pred,
                                                                     const float* __restrict__ gt, int C, int rows,
                                                                     int cols, int F, double* __restrict__ out) {
  extern __shared__ float msm[];
  __shared__ double red[MET_THREADS / 32][MET_VALS];
  const int n = blockIdx.x;
  const int npix = rows * cols;
  const int per = 3 * npix * F;                       // floats of the three properties of one sample
  float* sp = msm;
  float* sg = msm + per;
  const size_t sample = static_cast<size_t>(n) * C * npix * F;
  if (((sample | static_cast<size_t>(per)) & 3) == 0) {
    const float4* p4 = reinterpret_cast<const float4*>(pred + sample);
    const float4* g4 = reinterpret_cast<const float4*>(gt + sample);
    for (int i = threadIdx.x; i < per / 4; i += MET_THREADS) {
      reinterpret_cast<float4*>(sp)[i] = __ldg(p4 + i);
      reinterpret_cast<float4*>(sg)[i] = __ldg(g4 + i);
    }
  } else {
    for (int i = threadIdx.x; i < per; i += MET_THREADS) {
      sp[i] = __ldg(pred + sample + i);
      sg[i] = __ldg(gt + sample + i);
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = 0; j < F; ++j) {
    double acc[MET_SUMS];
    float mn[3], mx[3];
#pragma unroll
    for (int k = 0; k < MET_SUMS; ++k) acc[k] = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) { mn[c] = INFINITY; mx[c] = -INFINITY; }
    for (int p = threadIdx.x; p < npix; p += MET_THREADS) {
      const int r = p / cols, q = p - r * cols;
      const bool masked = sg[p * F + j] > 0.00001f;   // rho mask (metricsGenerator.py:144)
      if (masked) acc[6] += 1.0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int o = (c * npix + p) * F + j;
        const float g = sg[o], y = sp[o];
        const float d = g - y;                       // float32 difference and square, fp64 mean: as numpy does
        const float sq = d * d;
        acc[c] += static_cast<double>(sq);
        if (masked) acc[3 + c] += static_cast<double>(sq);
        float tvp = 0.f, tvg = 0.f;                  // total variation: |down - here| + |right - here|
        if (r + 1 < rows) {
          tvp += fabsf(sp[o + cols * F] - y);
          tvg += fabsf(sg[o + cols * F] - g);
        }
        if (q + 1 < cols) {
          tvp += fabsf(sp[o + F] - y);
          tvg += fabsf(sg[o + F] - g);
        }
        acc[7 + c] += static_cast<double>(tvp);
        acc[10 + c] += static_cast<double>(tvg);
        mn[c] = fminf(mn[c], g);
        mx[c] = fmaxf(mx[c], g);
        if (c == 0) {
          acc[13] += static_cast<double>(y);
          acc[14] += static_cast<double>(g);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < MET_SUMS; ++k) acc[k] = warp_sum_d(acc[k]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      mn[c] = -warp_max(-mn[c]);
      mx[c] = warp_max(mx[c]);
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < MET_SUMS; ++k) red[warp][k] = acc[k];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        red[warp][15 + 2 * c] = static_cast<double>(mn[c]);
        red[warp][16 + 2 * c] = static_cast<double>(mx[c]);
      }
    }
    __syncthreads();
    if (threadIdx.x < MET_VALS) {
      const int k = threadIdx.x;
      double v = red[0][k];
      for (int w = 1; w < MET_THREADS / 32; ++w) {     // fixed order over the warps
        if (k < MET_SUMS) v += red[w][k];
        else if ((k - MET_SUMS) % 2 == 0) v = fmin(v, red[w][k]);
        else v = fmax(v, red[w][k]);
      }
      out[(static_cast<size_t>(n) * F + j) * MET_VALS + k] = v;
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace cm

using namespace cm;

extern "C" {

int cm_window_gather(const float* seq, int64_t n_seq, int channels, int rows, int cols, int raw_len, const int32_t* seq_idx,
                     const int32_t* t0, const int64_t* ids, int batch, int past_len, int future_len, float* past, float* future,
                     void* stream) {
  CM_CHECK(seq && seq_idx && t0 && past && future, "null pointer");
  CM_CHECK(n_seq >= 1 && channels >= 1 && rows >= 1 && cols >= 1 && batch >= 1, "bad geometry");
  CM_CHECK(past_len >= 1 && future_len >= 1 && past_len + future_len <= raw_len, "window (%d + %d) does not fit the raw length %d",
           past_len, future_len, raw_len);
  const int chw = channels * rows * cols;
  CM_CHECK(batch <= 65535, "batch %d > 65535 windows per gather", batch);
  window_gather_kernel<<<dim3((chw + 255) / 256, batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      seq, chw, raw_len, seq_idx, t0, reinterpret_cast<const long long*>(ids), past_len, future_len, past, future);
  CM_CUDA(cudaGetLastError());
  return 0;
}

int cm_metrics_reduce(const float* pred, const float* gt, int n, int channels, int rows, int cols, int frames, double* out,
                      void* stream) {
  CM_CHECK(pred && gt && out, "null pointer");
  CM_CHECK(n >= 1 && channels >= 3 && rows >= 1 && cols >= 1 && frames >= 1, "bad geometry (needs rho, vx, vy channels)");
  const size_t smem = static_cast<size_t>(2) * 3 * rows * cols * frames * sizeof(float);
  CM_CHECK(smem <= 200 * 1024, "one sample (%zu bytes of rho, vx, vy for prediction + ground truth) does not fit shared memory", smem);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    CM_CUDA(cudaFuncSetAttribute(metrics_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    smem_set = 200 * 1024;
  }
  metrics_reduce_kernel<<<n, MET_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(pred, gt, channels, rows, cols, frames, out);
  CM_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
