// GroupNorm(32 groups, eps 1e-6, affine) fused with swish, NHWC bf16, forward and backward.
// Reference: modules/diffusionmodules/model.py:29-35 (nonlinearity, Normalize) and their uses at
// :119-127 (ResnetBlock), :170 (AttnBlock, no swish), EncoderDual.py:116-117,126-127,
// DecoderPositional.py:142-143.  HBM-bound: forward = 2 reads + 1 write of the tensor
// (statistics pass + apply pass), backward = 2 reads of (dy, x) + 1 write.
// Statistics are accumulated per CTA in fp32 and combined across CTAs in a fixed order in fp64 (no atomics), so the
// E[x^2]-E[x]^2 form does not lose the variance.
#include "common.cuh"

namespace b2 {

__device__ __forceinline__ float sigmoidf_(float z) { return __fdividef(1.f, 1.f + __expf(-z)); }
// `swish` argument of every entry point below = activation after the affine normalisation:
//   0 none, 1 swish (x * sigmoid(x), model.py:29-31), 2 LeakyReLU(0.2) (the PatchGAN stages,
//   modules/discriminator/model.py:41,52,62: BatchNorm over [N,H,W] = these kernels with N = 1, G = C)
constexpr float LRELU_SLOPE = 0.2f;
__device__ __forceinline__ float act_fwd(float z, int act) {
  if (act == 1) return z * sigmoidf_(z);
  if (act == 2) return z > 0.f ? z : LRELU_SLOPE * z;
  return z;
}
// d act(z) / dz
__device__ __forceinline__ float act_bwd(float z, int act) {
  if (act == 1) {
    const float sg = sigmoidf_(z);
    return sg * (1.f + z * (1.f - sg));
  }
  if (act == 2) return z > 0.f ? 1.f : LRELU_SLOPE;
  return 1.f;
}
// ------------------------------------------------------------------ forward statistics
// grid (chunks, N); block 256.  Thread t owns channel vector (8 ch) v = t % (C/8) and walks rows.
// Deterministic: per-thread sums are combined in a fixed order in shared memory, every CTA writes its
// (sum, sum of squares) per group to part[n][chunk][g], and gn_finalize adds the chunks in order.
__global__ void gn_partial_stats_kernel(const __nv_bfloat16* __restrict__ x, float* part, int HW,
                                        int C, int G, int rows_per_block) {
  extern __shared__ float sh[];  // [2][rstep][C]
  const int n = blockIdx.y;
  const int vecs = C >> 3;
  const int v = threadIdx.x % vecs;
  const int rlane = threadIdx.x / vecs;
  const int rstep = blockDim.x / vecs;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(HW, r0 + rows_per_block);
  float s[8], ss[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; ss[i] = 0.f; }
  const __nv_bfloat16* base = x + (static_cast<long long>(n) * HW) * C + v * 8;
  int r = r0 + rlane;
  for (; r + 3 * rstep < r1; r += 4 * rstep) {
    uint4 u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      u[j] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(r + j * rstep) * C));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f[8];
      unpack8(u[j], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += f[i]; ss[i] = fmaf(f[i], f[i], ss[i]); }
    }
  }
  for (; r < r1; r += rstep) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(r) * C));
    float f[8];
    unpack8(u, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] += f[i]; ss[i] = fmaf(f[i], f[i], ss[i]); }
  }
  float* sh_s = sh;
  float* sh_q = sh + rstep * C;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sh_s[rlane * C + v * 8 + i] = s[i];
    sh_q[rlane * C + v * 8 + i] = ss[i];
  }
  __syncthreads();
  const int cg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int rl = 0; rl < rstep; ++rl)
      for (int c = 0; c < cg; ++c) {
        a += sh_s[rl * C + g * cg + c];
        b += sh_q[rl * C + g * cg + c];
      }
    float* o = part + ((static_cast<long long>(n) * gridDim.x + blockIdx.x) * G + g) * 2;
    o[0] = a;
    o[1] = b;
  }
}
// One warp per (image, group): lanes take the chunks round-robin (fp64 partial sums), then a fixed xor tree -
// deterministic, and O(chunks / 32) deep (BatchNorm = one "image" of 1184 chunks would otherwise be a serial loop).
__global__ void gn_finalize_kernel(const float* __restrict__ part, float* stats, int N, int G,
                                   int chunks, double inv_count, float eps) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= N * G) return;
  const int n = i / G, g = i % G;
  double a = 0.0, b = 0.0;
  for (int c = lane; c < chunks; c += 32) {
    const float2 v = *reinterpret_cast<const float2*>(part + ((static_cast<long long>(n) * chunks + c) * G + g) * 2);
    a += v.x;
    b += v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane) return;
  const double mean = a * inv_count;
  double var = b * inv_count - mean * mean;
  if (var < 0) var = 0;
  stats[2 * i] = static_cast<float>(mean);
  stats[2 * i + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

// ------------------------------------------------------------------ forward apply
// grid (chunks, N); thread = (channel vector v, row lane): y = swish(x * a + b) with the per-channel
// a = rstd*gamma, b = beta - mean*rstd*gamma held in registers.
__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                __nv_bfloat16* __restrict__ y, int HW, int C, int G, int swish,
                                int rows_per_block) {
  const int n = blockIdx.y;
  const int vecs = C >> 3;
  const int v = threadIdx.x % vecs;
  const int rlane = threadIdx.x / vecs;
  const int rstep = blockDim.x / vecs;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(HW, r0 + rows_per_block);
  const int cg = C / G;
  float a[8], b[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = v * 8 + k;
    const int g = c / cg;
    const float mean = stats[(n * G + g) * 2], rstd = stats[(n * G + g) * 2 + 1];
    a[k] = rstd * gamma[c];
    b[k] = beta[c] - mean * a[k];
  }
  const long long off = (static_cast<long long>(n) * HW) * C + v * 8;
  int r = r0 + rlane;
  for (; r + 3 * rstep < r1; r += 4 * rstep) {
    uint4 u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      u[j] = __ldg(reinterpret_cast<const uint4*>(x + off + static_cast<long long>(r + j * rstep) * C));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f[8];
      unpack8(u[j], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        f[k] = act_fwd(fmaf(f[k], a[k], b[k]), swish);
      }
      *reinterpret_cast<uint4*>(y + off + static_cast<long long>(r + j * rstep) * C) = pack8(f);
    }
  }
  for (; r < r1; r += rstep) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + off + static_cast<long long>(r) * C));
    float f[8];
    unpack8(u, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      f[k] = act_fwd(fmaf(f[k], a[k], b[k]), swish);
    }
    *reinterpret_cast<uint4*>(y + off + static_cast<long long>(r) * C) = pack8(f);
  }
}

// ------------------------------------------------------------------ backward statistics
// part[n][chunk][c] = (sum_rows dz, sum_rows dz * xhat) over the chunk's rows (dz = grad wrt GN output),
// combined in a fixed order; gn_bwd_reduce adds the chunks in order into ws_nc[n][c].
__global__ void gn_bwd_partial_kernel(const __nv_bfloat16* __restrict__ dy,
                                      const __nv_bfloat16* __restrict__ x,
                                      const float* __restrict__ stats,
                                      const float* __restrict__ gamma,
                                      const float* __restrict__ beta, float* part, int HW, int C,
                                      int G, int swish, int rows_per_block) {
  extern __shared__ float sh[];  // [2][rstep][C]
  const int n = blockIdx.y;
  const int vecs = C >> 3;
  const int v = threadIdx.x % vecs;
  const int rlane = threadIdx.x / vecs;
  const int rstep = blockDim.x / vecs;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(HW, r0 + rows_per_block);
  const int cg = C / G;
  float a[8], b[8], mean[8], rstd[8], gm[8], bt[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = v * 8 + k;
    const int g = c / cg;
    a[k] = 0.f; b[k] = 0.f;
    mean[k] = stats[(n * G + g) * 2]; rstd[k] = stats[(n * G + g) * 2 + 1];
    gm[k] = gamma[c]; bt[k] = beta[c];
  }
  const long long off = (static_cast<long long>(n) * HW) * C + v * 8;
  auto accum = [&](const uint4& ux, const uint4& ud) {
    float fx[8], fd[8];
    unpack8(ux, fx); unpack8(ud, fd);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (fx[k] - mean[k]) * rstd[k];
      float dz = fd[k];
      if (swish) dz *= act_bwd(fmaf(xh, gm[k], bt[k]), swish);
      a[k] += dz;
      b[k] = fmaf(dz, xh, b[k]);
    }
  };
  int r = r0 + rlane;
  for (; r + rstep < r1; r += 2 * rstep) {
    const uint4 ux0 = __ldg(reinterpret_cast<const uint4*>(x + off + static_cast<long long>(r) * C));
    const uint4 ud0 = __ldg(reinterpret_cast<const uint4*>(dy + off + static_cast<long long>(r) * C));
    const uint4 ux1 = __ldg(reinterpret_cast<const uint4*>(x + off + static_cast<long long>(r + rstep) * C));
    const uint4 ud1 = __ldg(reinterpret_cast<const uint4*>(dy + off + static_cast<long long>(r + rstep) * C));
    accum(ux0, ud0);
    accum(ux1, ud1);
  }
  for (; r < r1; r += rstep) {
    const uint4 ux = __ldg(reinterpret_cast<const uint4*>(x + off + static_cast<long long>(r) * C));
    const uint4 ud = __ldg(reinterpret_cast<const uint4*>(dy + off + static_cast<long long>(r) * C));
    accum(ux, ud);
  }
  float* sh_a = sh;
  float* sh_b = sh + rstep * C;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sh_a[rlane * C + v * 8 + k] = a[k];
    sh_b[rlane * C + v * 8 + k] = b[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float sa = 0.f, sb = 0.f;
    for (int rl = 0; rl < rstep; ++rl) { sa += sh_a[rl * C + c]; sb += sh_b[rl * C + c]; }
    float* o = part + ((static_cast<long long>(n) * gridDim.x + blockIdx.x) * C + c) * 2;
    o[0] = sa;
    o[1] = sb;
  }
}
// one warp per (image, channel): lanes take the chunks round-robin, fixed xor tree (deterministic)
__global__ void gn_bwd_reduce_kernel(const float* __restrict__ part, float* ws_nc, int N, int C,
                                     int chunks) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= N * C) return;
  const int n = i / C, c = i % C;
  float a = 0.f, b = 0.f;
  for (int k = lane; k < chunks; k += 32) {
    const float2 v = *reinterpret_cast<const float2*>(part + ((static_cast<long long>(n) * chunks + k) * C + c) * 2);
    a += v.x;
    b += v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) {
    ws_nc[2 * i] = a;
    ws_nc[2 * i + 1] = b;
  }
}
// dgamma[c] = sum_n ws[n][c][1], dbeta[c] = sum_n ws[n][c][0]
__global__ void gn_bwd_param_kernel(const float* __restrict__ ws_nc, float* dgb, int N, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int n = 0; n < N; ++n) {
    a += ws_nc[(static_cast<long long>(n) * C + c) * 2 + 0];
    b += ws_nc[(static_cast<long long>(n) * C + c) * 2 + 1];
  }
  dgb[c] = b;       // dgamma
  dgb[C + c] = a;   // dbeta
}
// dx = rstd * (dz*gamma - S1/cnt - xhat * S2/cnt),  S1 = sum_g dz*gamma, S2 = sum_g dz*gamma*xhat
// Same thread mapping as the forward apply; per-channel constants live in registers.
__global__ void gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy,
                                    const __nv_bfloat16* __restrict__ x,
                                    const float* __restrict__ stats,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ ws_nc, __nv_bfloat16* __restrict__ dx,
                                    const __nv_bfloat16* __restrict__ add, int HW, int C, int G,
                                    int swish, int rows_per_block) {
  const int n = blockIdx.y;
  const int vecs = C >> 3;
  const int v = threadIdx.x % vecs;
  const int rlane = threadIdx.x / vecs;
  const int rstep = blockDim.x / vecs;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(HW, r0 + rows_per_block);
  const int cg = C / G;
  const float inv_cnt = 1.f / (static_cast<float>(HW) * cg);
  // per-group constants once per CTA (thread g sums its group's channels in channel order), not once
  // per thread: for the small late-stage tensors the old per-thread loop cost more than the rows
  __shared__ float s_k[512][2];
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float S1 = 0.f, S2 = 0.f;
    for (int cc = g * cg; cc < (g + 1) * cg; ++cc) {
      const float gmm = gamma[cc];
      S1 = fmaf(gmm, ws_nc[(static_cast<long long>(n) * C + cc) * 2 + 0], S1);
      S2 = fmaf(gmm, ws_nc[(static_cast<long long>(n) * C + cc) * 2 + 1], S2);
    }
    s_k[g][0] = S1 * inv_cnt;
    s_k[g][1] = S2 * inv_cnt;
  }
  __syncthreads();
  float mean[8], rstd[8], gm[8], bt[8], k1[8], k2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = v * 8 + k;
    const int g = c / cg;
    mean[k] = stats[(n * G + g) * 2]; rstd[k] = stats[(n * G + g) * 2 + 1];
    gm[k] = gamma[c]; bt[k] = beta[c];
    k1[k] = s_k[g][0]; k2[k] = s_k[g][1];
  }
  const long long off = (static_cast<long long>(n) * HW) * C + v * 8;
  int r = r0 + rlane;
  for (; r + rstep < r1; r += 2 * rstep) {
    uint4 ux[2], ud[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      ux[j] = __ldg(reinterpret_cast<const uint4*>(x + off + static_cast<long long>(r + j * rstep) * C));
      ud[j] = __ldg(reinterpret_cast<const uint4*>(dy + off + static_cast<long long>(r + j * rstep) * C));
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float fx[8], fd[8], o[8];
      unpack8(ux[j], fx); unpack8(ud[j], fd);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xh = (fx[k] - mean[k]) * rstd[k];
        float dz = fd[k];
        if (swish) dz *= act_bwd(fmaf(xh, gm[k], bt[k]), swish);
        o[k] = rstd[k] * (dz * gm[k] - k1[k] - xh * k2[k]);
      }
      if (add) {
        float fa[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(add + off + static_cast<long long>(r + j * rstep) * C)), fa);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] += fa[k];
      }
      *reinterpret_cast<uint4*>(dx + off + static_cast<long long>(r + j * rstep) * C) = pack8(o);
    }
  }
  for (; r < r1; r += rstep) {
    const uint4 ux = __ldg(reinterpret_cast<const uint4*>(x + off + static_cast<long long>(r) * C));
    const uint4 ud = __ldg(reinterpret_cast<const uint4*>(dy + off + static_cast<long long>(r) * C));
    float fx[8], fd[8], o[8];
    unpack8(ux, fx); unpack8(ud, fd);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (fx[k] - mean[k]) * rstd[k];
      float dz = fd[k];
      if (swish) dz *= act_bwd(fmaf(xh, gm[k], bt[k]), swish);
      o[k] = rstd[k] * (dz * gm[k] - k1[k] - xh * k2[k]);
    }
    if (add) {
      float fa[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(add + off + static_cast<long long>(r) * C)), fa);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] += fa[k];
    }
    *reinterpret_cast<uint4*>(dx + off + static_cast<long long>(r) * C) = pack8(o);
  }
}

}  // namespace b2

using namespace b2;

static int pick_rows_per_block(int HW, int N) {
  // aim for >= ~4 waves of 148 SMs without making per-block work tiny
  int chunks = (148 * 8 + N - 1) / N;
  if (chunks < 1) chunks = 1;
  int rpb = (HW + chunks - 1) / chunks;
  if (rpb < 32) rpb = 32;
  return rpb;
}

extern "C" {

// Number of row chunks (CTAs per image) the statistics kernels use for an [N, HW, C] tensor; callers
// size their scratch with it: gn_stats ws = N*chunks*G*2 floats, gn_bwd_stats part = N*chunks*C*2 floats.
int b2dq_gn_chunks(int N, int HW) {
  if (N <= 0 || HW <= 0) return 0;
  const int rpb = pick_rows_per_block(HW, N);
  return (HW + rpb - 1) / rpb;
}

// stats[n][g] = (mean, rstd) fp32; ws = [N * chunks * G * 2] floats of scratch (b2dq_gn_chunks).
int b2dq_gn_stats(const void* x, float* stats, float* ws, int N, int HW, int C, int G, float eps,
                  cudaStream_t stream) {
  if (N <= 0 || HW <= 0) return 0;
  if (C % 8 || C % G || 256 % (C / 8)) return -1;
  const int rpb = pick_rows_per_block(HW, N);
  const int chunks = (HW + rpb - 1) / rpb;
  dim3 grid(chunks, N);
  gn_partial_stats_kernel<<<grid, 256, 2 * 256 * 8 * sizeof(float), stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ws, HW, C, G, rpb);
  const int NG = N * G;
  gn_finalize_kernel<<<(NG + 3) / 4, 128, 0, stream>>>(
      ws, stats, N, G, chunks, 1.0 / (static_cast<double>(HW) * (C / G)), eps);
  return (int)cudaGetLastError();
}

int b2dq_gn_apply(const void* x, const float* stats, const float* gamma, const float* beta, void* y,
                  int N, int HW, int C, int G, int swish, cudaStream_t stream) {
  if (N <= 0 || HW <= 0) return 0;
  if (C % 8 || C % G) return -1;
  if (256 % (C / 8)) return -1;
  const int rpb = pick_rows_per_block(HW, N);
  dim3 grid((HW + rpb - 1) / rpb, N);
  gn_apply_kernel<<<grid, 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), stats, gamma, beta,
      reinterpret_cast<__nv_bfloat16*>(y), HW, C, G, swish, rpb);
  return (int)cudaGetLastError();
}

// part: [N*chunks*C*2] floats of scratch (b2dq_gn_chunks); ws_nc: [N*C*2] floats (output of the pass).
int b2dq_gn_bwd_stats(const void* dy, const void* x, const float* stats, const float* gamma,
                      const float* beta, float* part, float* ws_nc, int N, int HW, int C, int G,
                      int swish, cudaStream_t stream) {
  if (N <= 0 || HW <= 0) return 0;
  if (C % 8 || C % G || 256 % (C / 8)) return -1;
  const int rpb = pick_rows_per_block(HW, N);
  const int chunks = (HW + rpb - 1) / rpb;
  dim3 grid(chunks, N);
  gn_bwd_partial_kernel<<<grid, 256, 2 * 256 * 8 * sizeof(float), stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(x), stats,
      gamma, beta, part, HW, C, G, swish, rpb);
  gn_bwd_reduce_kernel<<<(N * C + 7) / 8, 256, 0, stream>>>(part, ws_nc, N, C, chunks);
  return (int)cudaGetLastError();
}

// dgb: [2*C] floats: dgamma then dbeta (overwritten).
int b2dq_gn_bwd_apply(const void* dy, const void* x, const float* stats, const float* gamma,
                      const float* beta, const float* ws_nc, void* dx, float* dgb, const void* add,
                      int N, int HW, int C, int G, int swish, cudaStream_t stream) {
  if (N <= 0 || HW <= 0) return 0;
  if (G > 512 || C % 8 || C % G || 256 % (C / 8)) return -1;
  if (dgb) gn_bwd_param_kernel<<<(C + 127) / 128, 128, 0, stream>>>(ws_nc, dgb, N, C);
  const int rpb = pick_rows_per_block(HW, N);
  dim3 grid((HW + rpb - 1) / rpb, N);
  gn_bwd_apply_kernel<<<grid, 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(x), stats,
      gamma, beta, ws_nc, reinterpret_cast<__nv_bfloat16*>(dx),
      reinterpret_cast<const __nv_bfloat16*>(add), HW, C, G, swish, rpb);
  return (int)cudaGetLastError();
}

// dgb [2*C] = (dgamma, dbeta) from ws_nc [N*C*2] (for callers that run the backward in image groups and
// pass dgb = NULL to b2dq_gn_bwd_apply).
int b2dq_gn_bwd_param(const float* ws_nc, float* dgb, int N, int C, cudaStream_t stream) {
  if (N <= 0 || C <= 0) return 0;
  gn_bwd_param_kernel<<<(C + 127) / 128, 128, 0, stream>>>(ws_nc, dgb, N, C);
  return (int)cudaGetLastError();
}

}  // extern "C"
