// A operand in tensor memory (tcgen05.cp smem -> TMEM, then TS-mode tcgen05.mma): layout check and pace on sm_100a.
//   1. layout: a logical [256 rows][32 fp16] K-major SW64 tile (as TMA writes it) is copied with tcgen05.cp.128x256b using the
//      MMA's own A descriptor (start = tile + row_off*64 + kslice*32) and read back with tcgen05.ld: lane i / column j
//      must hold elements (row_off + i, kslice*16 + 2j, 2j+1).
//   2. pace: 64 x { cp of a fresh A slice ; TS MMA 128 x N x 16 } against 64 SS MMAs with fresh A (the conv kernels' pattern),
//      with 1 / 2 / 4 rotating A buffers in TMEM, cp only, TS MMA only.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o utccp_microbench tools/utccp_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../crowdmod-ddpm-4d_b200/csrc/common.cuh"

namespace cm {
void set_error(const std::string&) {}
const char* get_error() { return ""; }
}  // namespace cm
using namespace cm;

__device__ __forceinline__ long long clk() { return clock64(); }

__global__ void __launch_bounds__(128, 1) layout_check(int row_off, int kslice, uint32_t* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  // logical X[r][k] = (r << 5) | k as 16-bit patterns, SW64 placement: chunk c of row r at r*64 + ((c ^ ((r >> 1) & 3)) << 4)
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) {
    const int r = i >> 5, k = i & 31, c = k >> 3;
    const uint32_t off = r * 64 + ((c ^ ((r >> 1) & 3)) << 4) + (k & 7) * 2;
    *reinterpret_cast<uint16_t*>(smem + off) = static_cast<uint16_t>((r << 5) | k);
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 32);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 1 && lane == 0) {
    const uint32_t lo = kmajor_desc_lo(smem_u32(smem) + row_off * 64) + 2 * kslice;
    tmem_cp_128x256b(tmem_base, lo, kmajor_desc_hi(64));
    umma_commit(&bar);
    mbar_wait(&bar, 0, nullptr, 0);
  }
  __syncthreads();
  tc_fence_after();
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(tmem_base + (static_cast<uint32_t>(warp * 32) << 16))
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 8; ++j) out[threadIdx.x * 8 + j] = r[j];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 32);
}

struct Res { long long t[8]; };

// mode 0: SS MMAs with a fresh A tile each (baseline) | 1: cp + TS MMA, NBUF rotating A buffers | 2: cp only | 3: TS MMA only
// (A resident) | 4: two cp + two TS MMAs per fresh tile (both k16 slices: the conv kernels' K = 32 pattern)
template <int N, int mode, int nbuf>
__global__ void __launch_bounds__(128, 1) pace(Res* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  constexpr uint32_t IDESC = make_idesc_f16(128, N);
  constexpr uint32_t DESC_HI = kmajor_desc_hi(64);
  const uint32_t a0 = kmajor_desc_lo(smem_u32(smem));
  const uint32_t b0 = kmajor_desc_lo(smem_u32(smem) + 96 * 1024);
  constexpr uint32_t A_STEP = (128 * 64) >> 4;
  const uint32_t a_tm = tmem_base + 256;             // A buffers: 8 columns per k16 slice
  if (warp == 1 && lane == 0) {
    long long t0 = clk();
#pragma unroll 8
    for (int i = 0; i < 64; ++i) {
      const uint32_t a = a0 + (i & 7) * A_STEP;
      const uint32_t ab = a_tm + (i & (nbuf - 1)) * 16;
      if (mode == 0) {
        umma_f16_lohi(tmem_base, a + 2 * (i & 1), b0 + 2 * (i & 1), DESC_HI, IDESC, 1u);
      } else if (mode == 1) {
        tmem_cp_128x256b(ab, a + 2 * (i & 1), DESC_HI);
        umma_f16_ts(tmem_base, ab, b0 + 2 * (i & 1), DESC_HI, IDESC, 1u);
      } else if (mode == 2) {
        tmem_cp_128x256b(ab, a + 2 * (i & 1), DESC_HI);
      } else if (mode == 3) {
        umma_f16_ts(tmem_base, a_tm + 8 * (i & 1), b0 + 2 * (i & 1), DESC_HI, IDESC, 1u);
      } else if (mode == 5) {        // SS, fresh A and fresh B (B tiles of N rows x 64 B, 8 of them in the upper 96 KB)
        umma_f16_lohi(tmem_base, a + 2 * (i & 1), b0 + (i & 7) * ((N * 64) >> 4) + 2 * (i & 1), DESC_HI, IDESC, 1u);
      } else if (mode == 6) {        // SS, same A, fresh B
        umma_f16_lohi(tmem_base, a0 + 2 * (i & 1), b0 + (i & 7) * ((N * 64) >> 4) + 2 * (i & 1), DESC_HI, IDESC, 1u);
      } else if (mode == 7) {        // the resident kernel's pattern: A views th*38 rows apart of one box, both k16 slices, fresh B each
        const uint32_t av = a0 + ((i >> 1) % 3) * ((38 * 64) >> 4) + ((i / 6) & 3) * ((200 * 64) >> 4);
        umma_f16_lohi(tmem_base, av + 2 * (i & 1), b0 + (i & 7) * ((N * 64) >> 4) + 2 * (i & 1), DESC_HI, IDESC, 1u);
      } else {
        tmem_cp_128x256b(ab, a, DESC_HI);
        tmem_cp_128x256b(ab + 8, a + 2, DESC_HI);
        umma_f16_ts(tmem_base, ab, b0, DESC_HI, IDESC, 1u);
        umma_f16_ts(tmem_base, ab + 8, b0 + 2, DESC_HI, IDESC, 1u);
      }
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

template <int N, int mode, int nbuf>
void run_one(Res* d) {
  static const char* what[8] = {"SS MMA, fresh A", "cp + TS MMA", "cp only", "TS MMA only (A resident)", "2 cp + 2 TS MMA (K = 32)", "SS MMA, fresh A, fresh B", "SS MMA, same A, fresh B", "SS MMA, res32 pattern"};
  cudaFuncSetAttribute(pace<N, mode, nbuf>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  Res h{};
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(d, 0, sizeof(Res));
    pace<N, mode, nbuf><<<1, 128, 200 * 1024>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("pace N=%d mode %d: %s\n", N, mode, cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
  }
  const double per = mode == 4 ? 128.0 : 64.0;
  printf("N=%-3d %-26s nbuf %d: issue %6lld, complete %6lld (%.1f cycles per %s)\n", N, what[mode], nbuf, h.t[0], h.t[1],
         h.t[1] / per, mode == 2 ? "cp" : "MMA");
}
template <int N>
void run_pace(Res* d) {
  run_one<N, 0, 1>(d);
  run_one<N, 1, 1>(d); run_one<N, 1, 2>(d); run_one<N, 1, 4>(d);
  run_one<N, 2, 1>(d); run_one<N, 2, 4>(d);
  run_one<N, 3, 1>(d);
  run_one<N, 4, 1>(d); run_one<N, 4, 4>(d);
  run_one<N, 5, 1>(d); run_one<N, 6, 1>(d); run_one<N, 7, 1>(d);
}

int main() {
  uint32_t* d_out;
  cudaMalloc(&d_out, 128 * 8 * 4);
  cudaFuncSetAttribute(layout_check, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  for (int row_off : {0, 1, 2, 38, 76})
    for (int ks : {0, 1}) {
      cudaMemset(d_out, 0xff, 128 * 8 * 4);
      layout_check<<<1, 128, 40 * 1024>>>(row_off, ks, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("layout_check: %s\n", cudaGetErrorString(e)); return 1; }
      std::vector<uint32_t> h(128 * 8);
      cudaMemcpy(h.data(), d_out, 128 * 8 * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int i = 0; i < 128; ++i)
        for (int j = 0; j < 8; ++j) {
          const uint32_t r = row_off + i, k = ks * 16 + 2 * j;
          const uint32_t want = ((r << 5) | k) | (((r << 5) | (k + 1)) << 16);
          if (h[i * 8 + j] != want) {
            if (bad < 4) printf("  row_off %d ks %d lane %d col %d: got %08x (row %u k %u | row %u k %u) want %08x\n", row_off, ks, i, j,
                                h[i * 8 + j], (h[i * 8 + j] & 0xffff) >> 5, h[i * 8 + j] & 31, h[i * 8 + j] >> 21, (h[i * 8 + j] >> 16) & 31, want);
            ++bad;
          }
        }
      printf("layout row_off %2d kslice %d: %s (%d mismatches)\n", row_off, ks, bad ? "MISMATCH" : "ok", bad);
    }
  Res* d;
  cudaMalloc(&d, sizeof(Res));
  run_pace<192>(d);
  run_pace<96>(d);
  run_pace<256>(d);
  return 0;
}
