// Dual-grain code permuter of the stage-2 tokenisation path
// (reference: modules/dynamic_modules/permuter.py:50-132).
//
// forward : the [B,F,F] code map + [B,Hc,Hc] grain map -> padded coarse / fine content, position and
//           segment sequences.  The reference builds them per sample with boolean masking, torch.cat and
//           pad_sequence; here one CTA per sample does an order-preserving stream compaction (ballot +
//           warp-count scan) and writes the eos / pad tail itself.
// backward: sequences -> [B,F,F] code map.  The reference walks every sequence element in Python
//           (B*(Lc+Lf) device round trips); here one CTA per sample finds the eos, resolves duplicate
//           positions "last writer wins" with an atomicMax on the sequence index in shared memory, and
//           writes the map once.
// Pure int64 index work: results are bit-exact by construction; HBM traffic is the inputs once and the
// outputs once.
#include "common.cuh"

namespace b2 {

// Exclusive prefix of `flag` over the CTA (blockDim.x a multiple of 32, <= 1024); total in `total`.
__device__ __forceinline__ int block_rank(bool flag, int* s_warp, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const unsigned b = __ballot_sync(0xffffffffu, flag);
  const int prefix = __popc(b & ((1u << lane) - 1u));
  if (lane == 0) s_warp[warp] = __popc(b);
  __syncthreads();
  if (warp == 0) {
    const int v = lane < nwarps ? s_warp[lane] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    s_warp[lane] = incl - v;
    if (lane == 31) s_warp[32] = incl;
  }
  __syncthreads();
  const int r = s_warp[warp] + prefix;
  total = s_warp[32];
  __syncthreads();
  return r;
}

struct PermuterCodes {
  long long content_pad, content_eos, cpos_pad, cpos_eos, fpos_pad, fpos_eos;
};

__global__ void __launch_bounds__(1024)
permuter_forward_kernel(const long long* __restrict__ indices, const long long* __restrict__ grain,
                        long long* __restrict__ c_content, long long* __restrict__ c_position,
                        long long* __restrict__ c_segment, long long* __restrict__ f_content,
                        long long* __restrict__ f_position, long long* __restrict__ f_segment,
                        int hw1, int hw2, int Lc, int Lf, int region_first, PermuterCodes codes) {
  __shared__ int s_warp[33];
  const int b = blockIdx.x;
  const int F = hw1 * hw2, cells = hw1 * hw1, fines = F * F, sub = hw2 * hw2;
  const long long* idx = indices + (size_t)b * fines;
  const long long* gr = grain + (size_t)b * cells;
  long long* cc = c_content + (size_t)b * Lc;
  long long* cp = c_position + (size_t)b * Lc;
  long long* cs = c_segment + (size_t)b * Lc;
  long long* fc = f_content + (size_t)b * Lf;
  long long* fp = f_position + (size_t)b * Lf;
  long long* fs = f_segment + (size_t)b * Lf;

  // ---- coarse sequence: cells with grain == 0 in raster order; content = top-left code of the cell (:58-60)
  int base = 0;
  for (int t0 = 0; t0 < cells; t0 += blockDim.x) {
    const int t = t0 + threadIdx.x;
    const bool flag = t < cells && gr[t] == 0;
    int total;
    const int r = block_rank(flag, s_warp, total);
    if (flag && base + r < Lc) {
      const int h1 = t / hw1, w1 = t - h1 * hw1;
      cc[base + r] = idx[(size_t)(h1 * hw2) * F + w1 * hw2];
      cp[base + r] = t;
    }
    base += total;
  }
  for (int j = threadIdx.x; j < Lc; j += blockDim.x) {
    cs[j] = 0;
    if (j == base) {
      cc[j] = codes.content_eos;
      cp[j] = codes.cpos_eos;
    } else if (j > base) {
      cc[j] = codes.content_pad;
      cp[j] = codes.cpos_pad;
    }
  }
  // ---- fine sequence: codes of the cells with grain == 1; region-first walks cell by cell (h2 w2 inside),
  //      row-first walks the fine map in raster order (:78-96); positions are always raster ids
  base = 0;
  for (int t0 = 0; t0 < fines; t0 += blockDim.x) {
    const int t = t0 + threadIdx.x;
    int y = 0, x = 0;
    if (t < fines) {
      if (region_first) {
        const int cell = t / sub, s = t - cell * sub;
        const int h1 = cell / hw1, w1 = cell - h1 * hw1, h2 = s / hw2, w2 = s - h2 * hw2;
        y = h1 * hw2 + h2;
        x = w1 * hw2 + w2;
      } else {
        y = t / F;
        x = t - y * F;
      }
    }
    const bool flag = t < fines && gr[(y / hw2) * hw1 + x / hw2] == 1;
    int total;
    const int r = block_rank(flag, s_warp, total);
    if (flag && base + r < Lf) {
      fc[base + r] = idx[(size_t)y * F + x];
      fp[base + r] = (long long)y * F + x;
    }
    base += total;
  }
  for (int j = threadIdx.x; j < Lf; j += blockDim.x) {
    fs[j] = 1;
    if (j == base) {
      fc[j] = codes.content_eos;
      fp[j] = codes.fpos_eos;
    } else if (j > base) {
      fc[j] = codes.content_pad;
      fp[j] = codes.fpos_pad;
    }
  }
}

__global__ void __launch_bounds__(1024)
permuter_backward_kernel(const long long* __restrict__ c_content, const long long* __restrict__ f_content,
                         const long long* __restrict__ c_position, const long long* __restrict__ f_position,
                         long long* __restrict__ target, int hw1, int hw2, int Lc, int Lf,
                         long long cpos_eos, long long fpos_eos) {
  extern __shared__ int s_last[];          // [cells] last coarse writer, [fines] last fine writer
  __shared__ int s_eos[2];
  const int b = blockIdx.x;
  const int F = hw1 * hw2, cells = hw1 * hw1, fines = F * F;
  int* last_c = s_last;
  int* last_f = s_last + cells;
  const long long* cc = c_content + (size_t)b * Lc;
  const long long* cp = c_position + (size_t)b * Lc;
  const long long* fc = f_content + (size_t)b * Lf;
  const long long* fp = f_position + (size_t)b * Lf;
  for (int i = threadIdx.x; i < cells + fines; i += blockDim.x) s_last[i] = -1;
  if (threadIdx.x == 0) {
    s_eos[0] = Lc;
    s_eos[1] = Lf;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < Lc; j += blockDim.x)
    if (cp[j] == cpos_eos) atomicMin(&s_eos[0], j);
  for (int j = threadIdx.x; j < Lf; j += blockDim.x)
    if (fp[j] == fpos_eos) atomicMin(&s_eos[1], j);
  __syncthreads();
  const int eos_c = s_eos[0], eos_f = s_eos[1];
  // elements before the eos write their slot in sequence order: the highest sequence index wins (:121,:127)
  for (int j = threadIdx.x; j < eos_c; j += blockDim.x) {
    const long long p = cp[j];
    if (p >= 0 && p < cells) atomicMax(&last_c[(int)p], j);
  }
  for (int j = threadIdx.x; j < eos_f; j += blockDim.x) {
    const long long p = fp[j];
    if (p >= 0 && p < fines) atomicMax(&last_f[(int)p], j);
  }
  __syncthreads();
  // the coarse map is only spread over the fine map when its eos was reached (:117-120)
  const bool spread = eos_c < Lc;
  long long* out = target + (size_t)b * fines;
  for (int t = threadIdx.x; t < fines; t += blockDim.x) {
    const int y = t / F, x = t - y * F;
    long long v = 0;
    if (spread) {
      const int lc = last_c[(y / hw2) * hw1 + x / hw2];
      if (lc >= 0) v = cc[lc];
    }
    const int lf = last_f[t];
    if (lf >= 0) v = fc[lf];
    out[t] = v;
  }
}

}  // namespace b2

extern "C" {

int b2dq_permuter_forward(const long long* indices, const long long* grain, long long* coarse_content,
                          long long* coarse_position, long long* coarse_segment, long long* fine_content,
                          long long* fine_position, long long* fine_segment, int B, int coarse_hw,
                          int fine_hw, int coarse_len, int fine_len, int region_first,
                          const long long* codes6, cudaStream_t stream) {
  if (B <= 0) return 0;
  if (coarse_hw <= 0 || fine_hw % coarse_hw || coarse_len <= 0 || fine_len <= 0) return -1;
  b2::PermuterCodes c{codes6[0], codes6[1], codes6[2], codes6[3], codes6[4], codes6[5]};
  b2::permuter_forward_kernel<<<B, 1024, 0, stream>>>(
      indices, grain, coarse_content, coarse_position, coarse_segment, fine_content, fine_position,
      fine_segment, coarse_hw, fine_hw / coarse_hw, coarse_len, fine_len, region_first, c);
  return (int)cudaGetLastError();
}

int b2dq_permuter_backward(const long long* coarse_content, const long long* fine_content,
                           const long long* coarse_position, const long long* fine_position,
                           long long* target, int B, int coarse_hw, int fine_hw, int coarse_len,
                           int fine_len, long long coarse_position_eos, long long fine_position_eos,
                           cudaStream_t stream) {
  if (B <= 0) return 0;
  if (coarse_hw <= 0 || fine_hw % coarse_hw) return -1;
  const size_t smem = (size_t)(coarse_hw * coarse_hw + fine_hw * fine_hw) * sizeof(int);
  if (smem > 200 * 1024) return -2;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(b2::permuter_backward_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  b2::permuter_backward_kernel<<<B, 1024, smem, stream>>>(
      coarse_content, fine_content, coarse_position, fine_position, target, coarse_hw,
      fine_hw / coarse_hw, coarse_len, fine_len, coarse_position_eos, fine_position_eos);
  return (int)cudaGetLastError();
}

}  // extern "C"
