// Micro-benchmark of the MMA-issuing thread's costs on sm_100a (one CTA, clock64 around each pattern):
//   * pace of back-to-back SS-mode tcgen05.mma (128 x N x 16, K-major SW64 / SW128 operands) and whether issue blocks,
//   * cost of tcgen05.commit, tcgen05.fence::after_thread_sync, mbarrier.try_wait on a completed phase,
//   * what a "stage boundary" (commit + wait + fence) costs between two groups of MMAs, with 1 and 2 issuing warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_microbench tools/umma_microbench.cu
// (operands are whatever is in shared memory: results are not checked, only timed)
#include <cstdio>
#include <cstdlib>
#include "../crowdmod-ddpm-4d_b200/csrc/common.cuh"

namespace cm {
void set_error(const std::string&) {}
const char* get_error() { return ""; }
}  // namespace cm
using namespace cm;

__device__ __forceinline__ long long clk() { return clock64(); }

struct Res {
  long long t[64];
};

// mode 0: G groups of NM MMAs back to back, one commit at the end, wait.
// mode 1: G groups of NM MMAs, after each group: commit(bar_g) only.
// mode 2: after each group: commit + try_wait on an already-completed barrier.
// mode 3: after each group: commit + try_wait(done) + tcgen05.fence::after_thread_sync.
// mode 4: like 3 + elect_one + __syncwarp (whole warp runs the loop).
// mode 5: two issuing warps (warp 1: accumulator 0, warp 2: accumulator 1), each G groups of NM/2 MMAs with the mode-3 boundary.
// mode 6: like 5 but warp 2 starts half a group late (phase offset).
template <int N, int ROWB>
__global__ void __launch_bounds__(128, 1) bench(int mode, int G, int NM, Res* out, int a_off = 0) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  __shared__ uint64_t bars[40];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 40; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  constexpr uint32_t IDESC = make_idesc_f16(128, N);
  constexpr uint32_t DESC_HI = kmajor_desc_hi(ROWB);
  const uint32_t a_lo = kmajor_desc_lo(smem_u32(smem) + a_off);   // a_off: row-offset view of the A tile (rows of ROWB bytes)
  const uint32_t b_lo = kmajor_desc_lo(smem_u32(smem) + 32 * 1024);
  uint64_t* done = &bars[32];            // a barrier whose phase 0 is completed up front
  if (threadIdx.x == 0) mbar_arrive(done);
  __syncthreads();

  if (mode <= 4) {
    if (warp == 1) {
      const bool whole_warp = mode == 4;
      if (lane == 0 || whole_warp) {
        long long t0 = clk(), t_issue = 0;
        for (int g = 0; g < G; ++g) {
          const bool me = whole_warp ? elect_one() : true;
          if (me) {
            for (int i = 0; i < NM; ++i) umma_f16_lohi(tmem_base, a_lo + 2 * (i & 1), b_lo + 2 * (i & 1), DESC_HI, IDESC, 1u);
            if (mode >= 1 || g == G - 1) umma_commit(&bars[g & 15]);
          }
          if (whole_warp) __syncwarp();
          if (mode >= 2) mbar_wait(done, 0, nullptr, 0);
          if (mode >= 3) tc_fence_after();
        }
        t_issue = clk();
        // wait for the last commit (barrier (G-1)&15 has been arrived on ceil(G/16) times in modes >= 1)
        if (mode == 0) mbar_wait(&bars[(G - 1) & 15], 0, nullptr, 0);
        else {
          const int idx = (G - 1) & 15;
          const int arrivals = (G - 1) / 16 + 1;
          mbar_wait(&bars[idx], (arrivals - 1) & 1, nullptr, 0);
        }
        long long t1 = clk();
        if (lane == 0) { out->t[0] = t_issue - t0; out->t[1] = t1 - t0; }
      }
    }
  } else {
    if ((warp == 1 || warp == 2) && lane == 0) {
      const int w = warp - 1;
      const uint32_t d = tmem_base + w * 256;
      if (mode == 6 && w == 1) {            // phase offset: burn ~ half a group
        long long s = clk();
        while (clk() - s < (long long)NM * 17) {}
      }
      long long t0 = clk();
      for (int g = 0; g < G; ++g) {
        for (int i = 0; i < NM / 2; ++i) umma_f16_lohi(d, a_lo + 2 * (i & 1), b_lo + 2 * (i & 1), DESC_HI, IDESC, 1u);
        umma_commit(&bars[w * 16 + (g & 15)]);
        mbar_wait(done, 0, nullptr, 0);
        tc_fence_after();
      }
      long long t_issue = clk();
      const int idx = (G - 1) & 15;
      const int arrivals = (G - 1) / 16 + 1;
      mbar_wait(&bars[w * 16 + idx], (arrivals - 1) & 1, nullptr, 0);
      long long t1 = clk();
      out->t[2 * w] = t_issue - t0;
      out->t[2 * w + 1] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// pace of 64 back-to-back MMAs whose operands are (re)used in different patterns: vary = 0 same A and B every time,
// 1 a different A tile per MMA (B fixed), 2 a different B tile per MMA (A fixed), 3 both different.  Distinct tiles are
// 4 KB (A) / N*ROWB bytes (B) apart inside a 96 KB window, so nothing can be served from an operand cache.
template <int N, int ROWB>
__global__ void __launch_bounds__(128, 1) bench_fresh(int vary, Res* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  constexpr uint32_t IDESC = make_idesc_f16(128, N);
  constexpr uint32_t DESC_HI = kmajor_desc_hi(ROWB);
  const uint32_t a0 = kmajor_desc_lo(smem_u32(smem));
  const uint32_t b0 = kmajor_desc_lo(smem_u32(smem) + 96 * 1024);
  constexpr uint32_t A_STEP = (128 * ROWB) >> 4, B_STEP = (N * ROWB) >> 4;     // one whole tile further
  constexpr int NA = 96 * 1024 / (128 * ROWB), NB = 96 * 1024 / (N * ROWB);
  if (warp == 1 && lane == 0) {
    long long t0 = clk();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
      const uint32_t a = a0 + ((vary & 1) ? (i % NA) * A_STEP : 0) + 2 * (i & 1);
      const uint32_t b = b0 + ((vary & 2) ? (i % NB) * B_STEP : 0) + 2 * (i & 1);
      umma_f16_lohi(tmem_base, a, b, DESC_HI, IDESC, 1u);
    }
    umma_commit(&bar);
    long long t1 = clk();
    mbar_wait(&bar, 0, nullptr, 0);
    long long t2 = clk();
    out->t[0] = t1 - t0;
    out->t[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int N, int ROWB>
void run_fresh(const char* name, Res* d) {
  cudaFuncSetAttribute(bench_fresh<N, ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  static const char* what[4] = {"same A, same B", "fresh A, same B", "same A, fresh B", "fresh A, fresh B"};
  for (int vary = 0; vary < 4; ++vary) {
    Res h{};
    cudaMemset(d, 0, sizeof(Res));
    bench_fresh<N, ROWB><<<1, 128, 200 * 1024>>>(vary, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s fresh %d: %s\n", name, vary, cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
    printf("%s %-18s: 64 MMAs issue %lld, complete %lld (%.1f / MMA)\n", name, what[vary], h.t[0], h.t[1], (h.t[1] - 752) / 63.0);
  }
}

// cost of single instructions, averaged over 64 repetitions (lane 0 of warp 1)
__global__ void __launch_bounds__(128, 1) instr_costs(Res* out) {
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) mbar_arrive(&bars[0]);
  __syncthreads();
  if (warp == 1 && lane == 0) {
    long long t0 = clk();
    for (int i = 0; i < 64; ++i) mbar_wait(&bars[0], 0, nullptr, 0);
    long long t1 = clk();
    for (int i = 0; i < 64; ++i) tc_fence_after();
    long long t2 = clk();
    for (int i = 0; i < 64; ++i) tc_fence_before();
    long long t3 = clk();
    // commit with nothing outstanding, then wait for its arrival (round trip), 32 times alternating phases
    for (int i = 0; i < 32; ++i) {
      umma_commit(&bars[1]);
      mbar_wait(&bars[1], i & 1, nullptr, 0);
    }
    long long t4 = clk();
    for (int i = 0; i < 64; ++i) asm volatile("" ::: "memory");
    long long t5 = clk();
    // commit issue cost alone (arrivals pile up on bars[2] with a huge expected count: never waited on)
    for (int i = 0; i < 32; ++i) umma_commit(&bars[2]);
    long long t6 = clk();
    // mbarrier.test_wait (non-blocking) on a completed phase, and three independent try_waits issued back to back
    long long t7 = clk();
    uint32_t acc = 0;
    for (int i = 0; i < 64; ++i) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(&bars[0])), "r"(0u) : "memory");
      if (!ok) break;
      acc += ok;
    }
    long long t8 = clk();
    for (int i = 0; i < 64; ++i) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred P, Q, R;\n\t"
                   "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %4;\n\t"
                   "mbarrier.try_wait.parity.shared::cta.b64 Q, [%2], %4;\n\t"
                   "mbarrier.try_wait.parity.shared::cta.b64 R, [%3], %4;\n\t"
                   "and.pred P, P, Q;\n\tand.pred P, P, R;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(&bars[0])), "r"(smem_u32(&bars[0])), "r"(smem_u32(&bars[0])), "r"(0u) : "memory");
      if (!ok) break;
      acc += ok;
    }
    long long t9 = clk();
    out->t[7] = (t8 - t7) / 64 + (acc == 12345);
    out->t[8] = (t9 - t8) / 64;
    out->t[0] = (t1 - t0) / 64;
    out->t[1] = (t2 - t1) / 64;
    out->t[2] = (t3 - t2) / 64;
    out->t[3] = (t4 - t3) / 32;
    out->t[4] = (t5 - t4);
    out->t[5] = (t6 - t5) / 32;
  }
  if (warp == 2) {   // whole-warp elect + syncwarp
    long long t0 = clk();
    int acc = 0;
    for (int i = 0; i < 64; ++i) {
      if (elect_one()) acc += i;
      __syncwarp();
    }
    long long t1 = clk();
    if (lane == 0) out->t[6] = (t1 - t0) / 64 + (acc == -1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot, 32);
}

template <int N, int ROWB>
void run(const char* name, Res* d) {
  cudaFuncSetAttribute(bench<N, ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  const int G = 12, NM = 24;
  for (int mode = 0; mode <= 6; ++mode) {
    Res h{};
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(d, 0, sizeof(Res));
      bench<N, ROWB><<<1, 128, 80 * 1024>>>(mode, G, NM, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s mode %d: %s\n", name, mode, cudaGetErrorString(e)); exit(1); }
      cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
    }
    const double n = (double)G * NM;
    if (mode <= 4)
      printf("%s mode %d: issue %lld cyc (%.1f/MMA), complete %lld cyc (%.1f/MMA)\n", name, mode, h.t[0], h.t[0] / n, h.t[1], h.t[1] / n);
    else
      printf("%s mode %d: warp1 issue %lld complete %lld | warp2 issue %lld complete %lld  (%.1f cyc per MMA of both, by the slower)\n", name,
             mode, h.t[0], h.t[1], h.t[2], h.t[3], (double)(h.t[1] > h.t[3] ? h.t[1] : h.t[3]) / n);
  }
  // pace vs group size in mode 0 (one group)
  for (int nm : {1, 2, 4, 8, 16, 32, 64}) {
    Res h{};
    cudaMemset(d, 0, sizeof(Res));
    bench<N, ROWB><<<1, 128, 80 * 1024>>>(0, 1, nm, d);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
    printf("%s one group of %2d MMAs: issue %lld, complete %lld\n", name, nm, h.t[0], h.t[1]);
  }
}

// pace of a group of 64 back-to-back MMAs when the A tile starts `rows` rows into the swizzle pattern
template <int N, int ROWB>
void run_offsets(const char* name, Res* d) {
  cudaFuncSetAttribute(bench<N, ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  for (int rows : {0, 1, 2, 4, 8, 16, 38, 40, 76, 80}) {
    Res h{};
    cudaMemset(d, 0, sizeof(Res));
    bench<N, ROWB><<<1, 128, 80 * 1024>>>(0, 1, 64, d, rows * ROWB);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
    printf("%s A offset %2d rows: 64 MMAs issue %lld, complete %lld (%.1f / MMA)\n", name, rows, h.t[0], h.t[1], (h.t[1] - 752) / 63.0);
  }
}

int main() {
  Res* d;
  cudaMalloc(&d, sizeof(Res));
  {
    Res h{};
    instr_costs<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("instr_costs: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
    printf("try_wait(completed) %lld cyc | fence::after %lld | fence::before %lld | commit+wait round trip %lld | empty loop(64) %lld | commit issue %lld | elect+syncwarp %lld | test_wait(completed) %lld | 3 try_waits back to back %lld\n",
           h.t[0], h.t[1], h.t[2], h.t[3], h.t[4], h.t[5], h.t[6], h.t[7], h.t[8]);
  }
  run_fresh<192, 64>("N=192 SW64 ", d);
  run_fresh<96, 64>("N=96  SW64 ", d);
  run_fresh<192, 128>("N=192 SW128", d);
  run_fresh<256, 128>("N=256 SW128", d);
  run_offsets<192, 64>("N=192 SW64 ", d);
  run_offsets<96, 64>("N=96  SW64 ", d);
  run_offsets<192, 128>("N=192 SW128", d);
  run<96, 64>("N=96  SW64 ", d);
  run<96, 128>("N=96  SW128", d);
  run<192, 64>("N=192 SW64 ", d);
  run<256, 128>("N=256 SW128", d);
  run<32, 64>("N=32  SW64 ", d);
  return 0;
}
