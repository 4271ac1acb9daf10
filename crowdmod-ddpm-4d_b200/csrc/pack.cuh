// Weight packing bodies: fp32 nn.Conv3d / nn.Linear layouts -> the fp16 K-major rows the tcgen05 kernels
// load by TMA (optional hi+lo split: w = hi + lo exactly to 2^-22).  Shared by the per-tensor launchers
// (op-level tests) and by pack_all_kernel, which re-derives EVERY cache of a plan -- forward weights, the
// 8-phase UpSample folds, the zero-padded first conv, and the flipped / transposed / phase-folded dgrad
// weights -- in ONE table-driven launch after an optimizer step (the reference re-reads its fp32 weights
// in every cuDNN call; here the re-pack is the only per-step cost of keeping fp32 masters).
#pragma once
#include "common.cuh"

namespace cm {

__device__ __forceinline__ int src_tap_of(int tap, int perm) {
  // tap = (d*3 + h)*3 + w in activation-dim order (time, rows, cols); the reference's nn.Conv3d weight
  // is over (rows, cols, time): source index = (h*3 + w)*3 + d.
  return perm ? (((tap / 3) % 3) * 3 + (tap % 3)) * 3 + tap / 9 : tap;
}

// dst: [terms*cout][taps*cin + cinx], column = tap*cin + ci, then taps*cin + cx; channels ci >= cin_src
// are zero (first conv: 3 API channels padded to 32).
// dup = 2: the activation operand is a K-concatenated hi|lo fp16 pair per pixel ([0,cin) = hi, [cin,2cin) = lo):
// every tap slab becomes [W(ci) for the hi half | W(ci) for the lo half]; the lo WEIGHT term carries zeros against
// the lo half (see pack_dgrad_body).  cin / cinx are the LOGICAL channel counts; K per row = dup*(taps*cin + cinx).
__device__ __forceinline__ void pack_conv_body(const float* __restrict__ w, const float* __restrict__ wx,
                                               __half* __restrict__ dst, int cout, int cin, int cinx, int taps,
                                               int terms, int perm, int cin_src, int dup, size_t i0, size_t istep) {
  const size_t kmain = (size_t)taps * cin * dup;
  const size_t ktot = kmain + (size_t)cinx * dup;
  const size_t total = (size_t)cout * ktot;
  for (size_t idx = i0; idx < total; idx += istep) {
    const int n = (int)(idx / ktot);
    const size_t k = idx - (size_t)n * ktot;
    float v;
    bool lo_half;
    if (k < kmain) {
      const int tap = (int)(k / ((size_t)cin * dup));
      const int c2 = (int)(k - (size_t)tap * cin * dup);
      lo_half = c2 >= cin;
      const int ci = lo_half ? c2 - cin : c2;
      const int st = (taps == 27) ? src_tap_of(tap, perm) : tap;
      v = ci < cin_src ? w[((size_t)n * cin_src + ci) * taps + st] : 0.f;
    } else {
      const int c2 = (int)(k - kmain);
      lo_half = c2 >= cinx;
      v = wx[(size_t)n * cinx + (lo_half ? c2 - cinx : c2)];
    }
    const __half hi = __float2half_rn(v);
    dst[(size_t)n * ktot + k] = hi;
    if (terms == 2)
      dst[((size_t)cout + n) * ktot + k] = lo_half ? __float2half_rn(0.f) : __float2half_rn(v - __half2float(hi));
  }
}

// nearest-x2 + k3 p1 (layers.py:92-94) folded into 8 phase convs with 2x2x2 combined taps.
// dst: [terms*cout][64*dup*cin], column = phase*8*dup*cin + tap8*dup*cin + c2 (c2 < cin: hi half, else lo half).
__device__ __forceinline__ void pack_upsample_body(const float* __restrict__ w, __half* __restrict__ dst, int cout,
                                                   int cin, int terms, int perm, int dup, size_t i0, size_t istep) {
  const int cw = cin * dup;                       // packed channels per tap slab (hi | lo halves when dup = 2)
  const size_t ktot = (size_t)64 * cw;
  const size_t total = (size_t)cout * ktot;
  for (size_t idx = i0; idx < total; idx += istep) {
    const int n = (int)(idx / ktot);
    const int k = (int)(idx - (size_t)n * ktot);
    const int phase = k / (8 * cw);
    const int r = k - phase * 8 * cw;
    const int tap8 = r / cw;
    const int c2 = r - tap8 * cw;
    const bool lo_half = c2 >= cin;
    const int ci = lo_half ? c2 - cin : c2;
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
        for (int kw = lo[0]; kw <= hi[0]; ++kw) v += wp[src_tap_of((kd * 3 + kh) * 3 + kw, perm)];
    const __half h = __float2half_rn(v);
    dst[(size_t)n * ktot + k] = h;
    if (terms == 2)
      dst[((size_t)cout + n) * ktot + k] = lo_half ? __float2half_rn(0.f) : __float2half_rn(v - __half2float(h));
  }
}

// dgrad weights (see backward.cuh): rows n = forward input channel (x terms), K-major columns of the conv
// mode that computes the data gradient.  dup = 2: the dOut operand is a K-concatenated hi|lo fp16 pair
// ([pixel][2*cout_f]: channels [0, cout_f) = hi, [cout_f, 2*cout_f) = lo = fp16(v - hi)); every tap slab is
// then [W(co) for the hi half | W(co) for the lo half], and the lo WEIGHT term carries zeros against the lo
// half (the lo*lo product is below fp32 resolution), so dX = W_hi*d_hi + W_hi*d_lo + W_lo*d_hi.
__device__ __forceinline__ void pack_dgrad_body(int fwd_mode, const float* __restrict__ w, __half* __restrict__ dst,
                                                int cout_f, int cin_f, int terms, int perm, int dup, size_t ktot,
                                                size_t i0, size_t istep) {
  const size_t total = (size_t)cin_f * ktot;
  const int cw = dup * cout_f;                    // packed channels per tap slab
  for (size_t idx = i0; idx < total; idx += istep) {
    const int n = (int)(idx / ktot);              // row = forward input channel
    const int kq = (int)(idx - (size_t)n * ktot);
    const int slab = kq / cw;
    const int c2 = kq - slab * cw;
    const bool lo_half = c2 >= cout_f;
    const int co = lo_half ? c2 - cout_f : c2;
    float v = 0.f;
    if (fwd_mode == 0) {
      v = w[((size_t)co * cin_f + n) * 27 + src_tap_of(26 - slab, perm)];
    } else if (fwd_mode == 3) {
      v = w[(size_t)co * cin_f + n];
    } else if (fwd_mode == 1) {
      // k3 s2 p1 forward: dX[2j]   = W[1]^T dY[j]
      //                   dX[2j+1] = W[2]^T dY[j] + W[0]^T dY[j+1]
      // expressed in the 8-phase / 2x2x2-tap scheme of conv mode 2 (phase bit p: taps at
      // j-1, j when p = 0; j, j+1 when p = 1).
      const int phase = slab >> 3, tap8 = slab & 7;
      int t[3];
      bool live = true;
#pragma unroll
      for (int d = 0; d < 3; ++d) {             // d = 0: w, 1: h, 2: d
        const int p = (phase >> d) & 1, a = (tap8 >> d) & 1;
        if (p == 0) { t[d] = 1; live = live && (a == 1); }
        else t[d] = a ? 0 : 2;
      }
      if (live) v = w[((size_t)co * cin_f + n) * 27 + src_tap_of((t[2] * 3 + t[1]) * 3 + t[0], perm)];
    } else {
      // nearest-x2 + k3 p1 forward: dX[j] = sum_s Ws[s]^T dY[2j - 1 + s], s = 0..3 with
      // Ws = {W2, W1+W2, W0+W1, W0} per dimension (k4 s2 conv, conv mode 4).
      const int s[3] = {slab & 3, (slab >> 2) & 3, slab >> 4};   // w, h, d
      int lo[3], hi[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        lo[d] = (s[d] == 0) ? 2 : (s[d] == 1 ? 1 : 0);
        hi[d] = (s[d] == 0) ? 2 : (s[d] == 1 ? 2 : (s[d] == 2 ? 1 : 0));
      }
      const float* wp = w + ((size_t)co * cin_f + n) * 27;
      for (int kd = lo[2]; kd <= hi[2]; ++kd)
        for (int kh = lo[1]; kh <= hi[1]; ++kh)
          for (int kw = lo[0]; kw <= hi[0]; ++kw) v += wp[src_tap_of((kd * 3 + kh) * 3 + kw, perm)];
    }
    const __half h = __float2half_rn(v);
    dst[(size_t)n * ktot + kq] = h;
    if (terms == 2)
      dst[((size_t)cin_f + n) * ktot + kq] = lo_half ? __float2half_rn(0.f) : __float2half_rn(v - __half2float(h));
  }
}

// One job of pack_all_kernel (blockIdx.y = job).
struct PackJob {
  const float* w;
  const float* wx;
  __half* dst;
  int kind;       // 0 conv (taps 27 | 1, optional 1x1 slab, zero-padded channels), 1 UpSample fold, 2 dgrad
  int mode;       // kind 2: forward conv mode
  int cout, cin, cinx, taps, terms, perm, cin_src, dup;
  unsigned long long ktot;   // kind 2: packed K per row
};
int pack_all_enqueue(const PackJob* d_jobs, int njobs, cudaStream_t st);

}  // namespace cm
