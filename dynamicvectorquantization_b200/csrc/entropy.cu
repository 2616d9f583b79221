// Per-patch grey-level entropy of the entropy-routed DQ-VAE
// (reference: models/stage1_dynamic/dqvae_dual_entropy.py:25-63).
//
// One CTA per patch.  A warp takes 32 pixels at a time; lane l owns histogram bin l and receives
// the 32 grey values by shuffle, so the soft histogram needs no atomics and no [pixels x bins]
// intermediate (the reference materialises 268 MB of it at B=32).  HBM traffic is the image read
// once (12 B / pixel) plus 4 B per patch.
//
// Arithmetic follows the reference op by op in fp32 with no FMA contraction and no flush-to-zero:
// eps = 1e-40 is a subnormal, and log(pdf + eps) of the empty bins relies on it.
#include "common.cuh"

namespace b2 {

constexpr int kEntropyBins = 32;

__global__ void __launch_bounds__(256)
patch_entropy_kernel(const float* __restrict__ x, const float* __restrict__ bins,
                     float* __restrict__ out, int H, int W, int p, float sigma) {
  const int pw = W / p, ph = H / p;
  const int patch = blockIdx.x;
  const int b = patch / (ph * pw);
  const int rem = patch - b * (ph * pw);
  const int py = rem / pw, px = rem - py * pw;
  const size_t plane = (size_t)H * W;
  const float* img = x + (size_t)b * 3 * plane;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float bin = bins[lane];
  const int P = p * p;
  float acc = 0.f;
  for (int base = warp * 32; base < P; base += nwarps * 32) {
    const int i = base + lane;
    const bool ok = i < P;
    float v = 0.f;
    if (ok) {
      const int yy = py * p + i / p, xx = px * p + i % p;
      const size_t o = (size_t)yy * W + xx;
      // gray = 0.2989 R + 0.5870 G + 0.1140 B, each product and sum rounded separately (:51)
      v = __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, img[o]), __fmul_rn(0.5870f, img[o + plane])),
                    __fmul_rn(0.1140f, img[o + 2 * plane]));
    }
    const unsigned valid = __ballot_sync(0xffffffffu, ok);
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      const float vj = __shfl_sync(0xffffffffu, v, j);
      if ((valid >> j) & 1u) {
        const float t = __fdiv_rn(__fsub_rn(vj, bin), sigma);           // residuals / sigma (:32-33)
        acc += expf(__fmul_rn(-0.5f, __fmul_rn(t, t)));
      }
    }
  }
  __shared__ float s_part[8][kEntropyBins];
  s_part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    float tot = 0.f;
    for (int w = 0; w < nwarps; ++w) tot += s_part[w][lane];
    float pdf = __fdiv_rn(tot, (float)P);                                // mean over the patch (:35)
    const float norm = __fadd_rn(warp_sum(pdf), 1e-40f);                 // (:36)
    pdf = __fadd_rn(__fdiv_rn(pdf, norm), 1e-40f);                       // (:37)
    const float e = warp_sum(__fmul_rn(pdf, logf(pdf)));                 // (:39)
    if (lane == 0) out[patch] = -e;
  }
}

}  // namespace b2

extern "C" int b2dq_patch_entropy(const float* x_nchw, const float* bins, float* out, int B, int H,
                                  int W, int patch, int nbins, float sigma, cudaStream_t stream) {
  if (nbins != b2::kEntropyBins || patch <= 0 || H % patch || W % patch) return -1;
  const long long patches = (long long)B * (H / patch) * (W / patch);
  if (patches <= 0) return 0;
  b2::patch_entropy_kernel<<<(unsigned)patches, 256, 0, stream>>>(x_nchw, bins, out, H, W, patch, sigma);
  return (int)cudaGetLastError();
}
