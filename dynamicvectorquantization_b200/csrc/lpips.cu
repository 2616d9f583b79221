// LPIPS level head, fused (modules/losses/lpips.py:44-55,116-122 of the reference):
//
//   val[n] = mean_pixels sum_c w_c * drop_c * ( f0_c / (|f0| + eps)  -  f1_c / (|f1| + eps) )^2
//
// for one VGG level, f0 / f1 = NHWC bf16 feature maps [N, HW, C] of the two images.  The reference
// materialises both normalised maps, their difference, its square, the dropout mask and the 1x1 conv
// output in fp32 NCHW (ten full-size tensors per level and direction); here a pixel's C channels are read
// once by a group of lanes, reduced with shuffles, and only per-CTA partial sums leave the SM.  The
// backward pass recomputes the normalisation from the same two reads and writes the gradient w.r.t. either
// feature map in bf16.  HBM-bound: 2*C bf16 read per pixel forward, plus C..2C written backward.
//
// Dropout (the nn.Dropout in front of every lin head is live in training mode - LPIPS().eval() does not
// survive the LightningModule's .train()): a counter-based Bernoulli mask, hash(seed, element index),
// identical in forward and backward; `seed` is read from device memory so a captured CUDA graph draws a
// fresh mask every replay.
#include "common.cuh"

namespace b2 {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {   // lowbias32 (Wellons): full-avalanche integer hash
  x ^= x >> 16; x *= 0x21f0aaadu;
  x ^= x >> 15; x *= 0x735a2d97u;
  x ^= x >> 15;
  return x;
}
// keep-probability 1 - p: returns the multiplier of element `idx` (0 or 1/(1-p)); seed == null -> 1
__device__ __forceinline__ float drop_scale(const unsigned long long* seed, unsigned long long s,
                                            unsigned long long idx, uint32_t thresh, float inv_keep) {
  if (!seed) return 1.f;
  // hash the element index first, then key it: masks of two seeds are neither shifted copies nor
  // index permutations of one another
  const uint32_t e = mix32(static_cast<uint32_t>(idx) + 0x9e3779b9u * static_cast<uint32_t>(idx >> 32));
  const uint32_t h = mix32(mix32(e ^ static_cast<uint32_t>(s)) + static_cast<uint32_t>(s >> 32));
  return h >= thresh ? inv_keep : 0.f;
}

template <int TG, int VPL>   // lanes per pixel, 8-channel vectors per lane: C = 8 * TG * VPL
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = TG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int TG, int VPL>
__global__ void __launch_bounds__(256)
lpips_head_fwd_kernel(const uint4* __restrict__ f0, const uint4* __restrict__ f1, const float* __restrict__ w,
                      float* part, int HW, int pix_per_block, const unsigned long long* seed, uint32_t thresh,
                      float inv_keep) {
  constexpr int C = 8 * TG * VPL, PPW = 32 / TG;            // pixels a warp handles per iteration
  __shared__ float sh[8];
  const int n = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gl = lane % TG, gi = lane / TG;
  float wc[VPL][8];
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int k = 0; k < 8; ++k) wc[v][k] = w[(gl + v * TG) * 8 + k];
  const unsigned long long s = seed ? *seed : 0ull;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
  float acc = 0.f;
  // every lane of a warp must reach the shuffles: iterate on the warp-uniform pixel base
  for (int pb = p0 + warp * PPW; pb < p1; pb += 8 * PPW) {
    const int p = pb + gi;
    const bool live = p < p1;
    const long long base = (static_cast<long long>(n) * HW + (live ? p : p1 - 1)) * (C / 8);
    float a[VPL][8], b[VPL][8];
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      unpack8(__ldg(f0 + base + gl + v * TG), a[v]);
      unpack8(__ldg(f1 + base + gl + v * TG), b[v]);
#pragma unroll
      for (int k = 0; k < 8; ++k) { s0 = fmaf(a[v][k], a[v][k], s0); s1 = fmaf(b[v][k], b[v][k], s1); }
    }
    s0 = group_sum<TG, VPL>(s0);
    s1 = group_sum<TG, VPL>(s1);
    const float i0 = 1.f / (sqrtf(s0) + 1e-10f), i1 = 1.f / (sqrtf(s1) + 1e-10f);
    float t = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = a[v][k] * i0 - b[v][k] * i1;
        const unsigned long long idx = (static_cast<unsigned long long>(base) + gl + v * TG) * 8 + k;
        t = fmaf(wc[v][k] * drop_scale(seed, s, idx, thresh, inv_keep), d * d, t);
      }
    if (live) acc += t;                                     // lanes of a group hold disjoint channel slices
  }
  // every lane holds a partial: fixed-order reduction
  acc = warp_sum(acc);
  if (lane == 0) sh[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += sh[k];
    part[static_cast<long long>(n) * gridDim.x + blockIdx.x] = tot;
  }
}

// g[n] = d loss / d val[n]; G = g[n] / HW.  df0 / df1 may be null.
template <int TG, int VPL>
__global__ void __launch_bounds__(256)
lpips_head_bwd_kernel(const uint4* __restrict__ f0, const uint4* __restrict__ f1, const float* __restrict__ w,
                      const float* __restrict__ g, uint4* df0, uint4* df1, int HW, int pix_per_block,
                      const unsigned long long* seed, uint32_t thresh, float inv_keep) {
  constexpr int C = 8 * TG * VPL, PPW = 32 / TG;
  const int n = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gl = lane % TG, gi = lane / TG;
  float wc[VPL][8];
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int k = 0; k < 8; ++k) wc[v][k] = w[(gl + v * TG) * 8 + k];
  const unsigned long long s = seed ? *seed : 0ull;
  const float G = g[n] / static_cast<float>(HW);
  const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
  // every lane of a warp must reach the shuffles: iterate on the warp-uniform pixel base
  for (int pb = p0 + warp * PPW; pb < p1; pb += 8 * PPW) {
    const int p = pb + gi;
    const bool live = p < p1;
    const long long base = (static_cast<long long>(n) * HW + (live ? p : p1 - 1)) * (C / 8);
    float a[VPL][8], b[VPL][8], da[VPL][8];
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      unpack8(__ldg(f0 + base + gl + v * TG), a[v]);
      unpack8(__ldg(f1 + base + gl + v * TG), b[v]);
#pragma unroll
      for (int k = 0; k < 8; ++k) { s0 = fmaf(a[v][k], a[v][k], s0); s1 = fmaf(b[v][k], b[v][k], s1); }
    }
    s0 = group_sum<TG, VPL>(s0);
    s1 = group_sum<TG, VPL>(s1);
    const float n0 = sqrtf(s0), n1 = sqrtf(s1);
    const float i0 = 1.f / (n0 + 1e-10f), i1 = 1.f / (n1 + 1e-10f);
    float dot0 = 0.f, dot1 = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = a[v][k] * i0 - b[v][k] * i1;
        const unsigned long long idx = (static_cast<unsigned long long>(base) + gl + v * TG) * 8 + k;
        const float gd = 2.f * G * wc[v][k] * drop_scale(seed, s, idx, thresh, inv_keep) * d;   // d val / d a_c
        da[v][k] = gd;                                       // d/d b_c = -gd
        dot0 = fmaf(gd, a[v][k], dot0);
        dot1 = fmaf(gd, b[v][k], dot1);
      }
    dot0 = group_sum<TG, VPL>(dot0);
    dot1 = group_sum<TG, VPL>(dot1);
    if (!live) continue;
    // a = f / (|f| + eps):  d a_c / d f_j = delta_cj / (|f|+eps) - f_c f_j / (|f| (|f|+eps)^2)
    const float k0 = n0 > 0.f ? i0 * i0 / n0 : 0.f, k1 = n1 > 0.f ? i1 * i1 / n1 : 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      if (df0) {
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = da[v][k] * i0 - a[v][k] * dot0 * k0;
        df0[base + gl + v * TG] = pack8(o);
      }
      if (df1) {
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = -(da[v][k] * i1 - b[v][k] * dot1 * k1);
        df1[base + gl + v * TG] = pack8(o);
      }
    }
  }
}

}  // namespace b2

using namespace b2;

static int lpips_ppb(int N, int HW) {
  // enough CTAs for ~8 per SM over the batch, at least 64 pixels each
  long long want = (148LL * 8 + N - 1) / (N > 0 ? N : 1);
  long long ppb = (HW + want - 1) / want;
  if (ppb < 64) ppb = 64;
  return (int)ppb;
}

extern "C" {

// Number of per-image partial sums b2dq_lpips_head_fwd writes (part = N * chunks floats).
int b2dq_lpips_head_chunks(int N, int HW) {
  if (N <= 0 || HW <= 0) return 0;
  const int ppb = lpips_ppb(N, HW);
  return (HW + ppb - 1) / ppb;
}

static inline uint32_t drop_thresh(float p) {
  double t = (double)p * 4294967296.0;
  if (t < 0) t = 0;
  if (t > 4294967295.0) t = 4294967295.0;
  return (uint32_t)t;
}

int b2dq_lpips_head_fwd(const void* f0, const void* f1, const float* w, float* part, int N, int HW, int C,
                        const unsigned long long* seed, float p_drop, cudaStream_t st) {
  if (N <= 0 || HW <= 0) return 0;
  if (seed && !(p_drop >= 0.f && p_drop < 1.f)) return -2;
  const int ppb = lpips_ppb(N, HW);
  dim3 grid((HW + ppb - 1) / ppb, N);
  const uint32_t th = drop_thresh(p_drop);
  const float ik = 1.f / (1.f - p_drop);
  const uint4* a = reinterpret_cast<const uint4*>(f0);
  const uint4* b = reinterpret_cast<const uint4*>(f1);
  switch (C) {
    case 64: lpips_head_fwd_kernel<8, 1><<<grid, 256, 0, st>>>(a, b, w, part, HW, ppb, seed, th, ik); break;
    case 128: lpips_head_fwd_kernel<16, 1><<<grid, 256, 0, st>>>(a, b, w, part, HW, ppb, seed, th, ik); break;
    case 256: lpips_head_fwd_kernel<32, 1><<<grid, 256, 0, st>>>(a, b, w, part, HW, ppb, seed, th, ik); break;
    case 512: lpips_head_fwd_kernel<32, 2><<<grid, 256, 0, st>>>(a, b, w, part, HW, ppb, seed, th, ik); break;
    default: return -1;
  }
  return (int)cudaGetLastError();
}

int b2dq_lpips_head_bwd(const void* f0, const void* f1, const float* w, const float* g, void* df0, void* df1,
                        int N, int HW, int C, const unsigned long long* seed, float p_drop, cudaStream_t st) {
  if (N <= 0 || HW <= 0) return 0;
  if (seed && !(p_drop >= 0.f && p_drop < 1.f)) return -2;
  const int ppb = lpips_ppb(N, HW);
  dim3 grid((HW + ppb - 1) / ppb, N);
  const uint32_t th = drop_thresh(p_drop);
  const float ik = 1.f / (1.f - p_drop);
  const uint4* a = reinterpret_cast<const uint4*>(f0);
  const uint4* b = reinterpret_cast<const uint4*>(f1);
  uint4* d0 = reinterpret_cast<uint4*>(df0);
  uint4* d1 = reinterpret_cast<uint4*>(df1);
  switch (C) {
    case 64: lpips_head_bwd_kernel<8, 1><<<grid, 256, 0, st>>>(a, b, w, g, d0, d1, HW, ppb, seed, th, ik); break;
    case 128: lpips_head_bwd_kernel<16, 1><<<grid, 256, 0, st>>>(a, b, w, g, d0, d1, HW, ppb, seed, th, ik); break;
    case 256: lpips_head_bwd_kernel<32, 1><<<grid, 256, 0, st>>>(a, b, w, g, d0, d1, HW, ppb, seed, th, ik); break;
    case 512: lpips_head_bwd_kernel<32, 2><<<grid, 256, 0, st>>>(a, b, w, g, d0, d1, HW, ppb, seed, th, ik); break;
    default: return -1;
  }
  return (int)cudaGetLastError();
}

}  // extern "C"
