// General batched / split-K GEMM on tcgen05 with per-operand major-ness, used for
//   * convolution weight gradients (reference: autograd of nn.Conv2d in
//     modules/diffusionmodules/model.py:43-47,62-66,88-115,146-165):
//       dW[co, tap, ci] = sum_pixels dY[pixel, co] * X[pixel shifted by tap, ci]
//     A = dY (MN-major: channels contiguous, pixels are the contraction), B = X (MN-major),
//     up to 4 taps share one dY tile and accumulate into separate TMEM column ranges;
//     the pixel range is split across CTAs and fp32 partials are reduced by wgrad_reduce.
//   * the single-head attention contractions (model.py:176-188) and their gradients:
//     any combination of K-major / MN-major A and B, batched over images.
// Operand tiles are 128 B-swizzled TMA boxes; MN-major tiles are stacks of {64 mn, 64 k} boxes
// (8 KB each, LBO = 8 KB between 64-wide mn groups, 2 KB per K=16 step).
#include <cstdlib>
#include "common.cuh"
#include "tmap.h"

namespace b2 {

struct MmParams {
  int a_mn, b_mn;                 // 0 = K-major ([rows][K], K contiguous), 1 = MN-major ([K][rows])
  int ntaps;                      // total B-operand shifts (filter taps), 1..12
  int taps_per_cta;               // accumulators per CTA (1..3); tap groups are folded into grid.x
  int tap_c[12], tap_w[12], tap_p[12], tap_h[12];
  int KW, KH, KN;                 // MN-major k-block box extents (KW*KH*KN == 64)
  int ktiles_w, ktiles_h;         // k-block -> (kw, kh, kn) decomposition for MN-major operands
  int kblocks;                    // total k-blocks (of 64) in the contraction
  int splits;                     // split-K factor; blockIdx.z = batch*splits + split
  int M, N;                       // valid output extents
  void* out;
  long long oZ, oT, oM;           // element strides: per blockIdx.z, per tap, per output row
  float alpha;
  int out_f32;
  float* colsum;                  // optional [splits][M] fp32: sum over the contraction index of A[m,k]
                                  // (bias gradient = column sums of dY), computed by the tensor core as
                                  // A x ones; written by the tap-group-0 CTAs (MAXTAPS == 3 kernels only)
  int tma_out;                    // 1: output tiles leave through shared memory and bulk tensor stores (tmO)
};

// STRIP: the (up to 3) taps of a CTA are horizontal 1-pixel shifts of each other (one 3x3 filter row):
// the B operand is loaded ONCE per k-block as a strip of KW+2 = 66 pixels and tap t is read from pixel
// row t of the strip (descriptor start + t*128 B; the 128 B swizzle is an absolute-address function).
template <int BN, int STAGES, int MAXTAPS, bool STRIP = false>
struct MmCfg {
  static constexpr uint32_t A_BYTES = 128 * 128;
  static constexpr uint32_t B_BYTES = BN * 128;
  static constexpr uint32_t STRIP_ROWS = 66;
  static constexpr uint32_t STRIP_SLOT = 9 * 1024;                   // 66 x 128 B padded to 1 KB multiple
  static constexpr uint32_t STAGE_BYTES = STRIP ? (A_BYTES + (BN / 64) * STRIP_SLOT) : (A_BYTES + MAXTAPS * B_BYTES);
  static constexpr uint32_t ONES_BYTES = 16 * 128;                   // K-major [16 n][64 k] tile of bf16 1.0
  static constexpr uint32_t SMEM = STAGES * STAGE_BYTES + 1024 + 256 + ONES_BYTES + 1024;
  static constexpr uint32_t TMEM_COLS = (BN * MAXTAPS <= 128) ? 128 : (BN * MAXTAPS <= 256 ? 256 : 512);
};

// CL > 1 (strip variant, launched as clusters of CL = 3 CTAs along x): the three tap groups (filter rows) of one
// pixel range share the dY tile.  It is fetched from L2 ONCE per cluster - rank 0 loads channel box 0, rank 1 box 1,
// both multicast to all three CTAs - instead of once per CTA: 21.8 KB instead of 32.5 KB of L2->SM traffic per CTA
// and k-block (the unicast version ran at the ~42 B/clk/SM L2 limit, tensor pipe 61 %).  A stage is released to
// the producers by the tcgen05.commit of all three MMA issuers (multicast arrive on every CTA's empty barrier).
template <int BN, int STAGES, int MAXTAPS, bool STRIP = false, int CL = 1>
__global__ void __launch_bounds__(192, 1)
mmgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
              const __grid_constant__ CUtensorMap tmO, const __grid_constant__ MmParams p) {
  using Cfg = MmCfg<BN, STAGES, MAXTAPS, STRIP>;
  constexpr uint16_t CL_MASK = static_cast<uint16_t>((1u << CL) - 1u);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sBar = base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * STAGES, bar_tfull = sBar + 16 * STAGES;
  uint32_t* tmem_slot =
      reinterpret_cast<uint32_t*>(smem_raw + (sBar + 16 * STAGES + 16 - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tap group = fastest grid index: the CTAs working on the same pixel range with different
  // filter rows run at the same time, so dY / X are fetched from HBM once and hit in L2 after.
  const int ngroups = (p.ntaps + p.taps_per_cta - 1) / p.taps_per_cta;
  const int group = blockIdx.x % ngroups;
  const int tap0 = group * p.taps_per_cta;
  const int ntl = min(p.taps_per_cta, p.ntaps - tap0);          // taps handled by this CTA
  const int m0 = (blockIdx.x / ngroups) * 128, n0 = blockIdx.y * BN;
  const int batch = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
  const int per = (p.kblocks + p.splits - 1) / p.splits;
  const int kb0 = split * per;
  const int kb1 = min(p.kblocks, kb0 + per);
  const int kiters = max(kb1 - kb0, 0);
  // tile of ones (B operand of the column-sum MMA), after the barrier block, 1 KB aligned
  const uint32_t sOnes = (sBar + 256 + 1023u) & ~1023u;
  const bool do_colsum = (MAXTAPS == 3) && p.colsum != nullptr && group == 0 && p.a_mn;
  if (do_colsum) {
    uint32_t* o = reinterpret_cast<uint32_t*>(smem_raw + (sOnes - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < (int)(Cfg::ONES_BYTES / 4); i += blockDim.x) o[i] = 0x3F803F80u;
    fence_proxy_async_smem();
  }

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, CL);
    }
    mbar_init(bar_tfull, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_out) tma_prefetch_desc(&tmO);
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();        // every CTA's barriers exist before a peer multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t tx = STRIP ? (Cfg::A_BYTES + (BN / 64) * Cfg::STRIP_ROWS * 128)
                                : (Cfg::A_BYTES + ntl * Cfg::B_BYTES);
      for (int kb = kb0; kb < kb1; ++kb) {
        const int kw = kb % p.ktiles_w, kh = (kb / p.ktiles_w) % p.ktiles_h,
                  kn = kb / (p.ktiles_w * p.ktiles_h);
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        const uint32_t sa = base + stage * Cfg::STAGE_BYTES;
        const uint32_t fb = bar_full + 8 * stage;
        mbar_arrive_expect_tx(fb, tx);
        if constexpr (CL > 1) {
          // dY tile of the cluster: channel box `group` (ranks 0 and 1) for everybody
          if (group < 2)
            tma_load_5d_mc(sa + group * 8192, &tmA, fb, m0 + 64 * group, kw * p.KW, 0, kh * p.KH, kn * p.KN + batch,
                           CL_MASK);
        } else if (!p.a_mn) {
          tma_load_5d(sa, &tmA, fb, kb * 64, m0, 0, 0, batch);
        } else {
#pragma unroll
          for (int g = 0; g < 2; ++g)
            tma_load_5d(sa + g * 8192, &tmA, fb, m0 + 64 * g, kw * p.KW, 0, kh * p.KH,
                        kn * p.KN + batch);
        }
        if constexpr (STRIP) {
#pragma unroll
          for (int g = 0; g < BN / 64; ++g)
            tma_load_5d(sa + Cfg::A_BYTES + g * Cfg::STRIP_SLOT, &tmB, fb, n0 + 64 * g + p.tap_c[tap0],
                        kw * p.KW + p.tap_w[tap0], p.tap_p[tap0], kh * p.KH + p.tap_h[tap0],
                        kn * p.KN + batch);
        } else
        for (int t = 0; t < ntl; ++t) {
          const uint32_t sb = sa + Cfg::A_BYTES + t * Cfg::B_BYTES;
          const int tg = tap0 + t;
          if (!p.b_mn) {
            tma_load_5d(sb, &tmB, fb, kb * 64, n0, 0, 0, batch);
          } else {
#pragma unroll
            for (int g = 0; g < BN / 64; ++g)
              tma_load_5d(sb + g * 8192, &tmB, fb, n0 + 64 * g + p.tap_c[tg],
                          kw * p.KW + p.tap_w[tg], p.tap_p[tg], kh * p.KH + p.tap_h[tg],
                          kn * p.KN + batch);
          }
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, BN, p.a_mn, p.b_mn);
      const uint32_t a_step = p.a_mn ? 2048u : 32u, b_step = p.b_mn ? 2048u : 32u;
      const uint32_t a_lbo = p.a_mn ? 8192u : 0u, b_lbo = p.b_mn ? 8192u : 0u;
      uint32_t stage = 0, phase = 0;
      for (int it = 0; it < kiters; ++it) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        const uint32_t sa = base + stage * Cfg::STAGE_BYTES;
        for (int t = 0; t < ntl; ++t) {
          const uint32_t sb = sa + Cfg::A_BYTES + (STRIP ? t * 128u : t * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = make_smem_desc(sa + k * a_step, a_lbo, 1024);
            const uint64_t db = make_smem_desc(sb + k * b_step, STRIP ? Cfg::STRIP_SLOT : b_lbo, 1024);
            umma_bf16(tmem_base + t * BN, da, db, idesc, (it | k) ? 1u : 0u);
          }
        }
        if (do_colsum) {
          const uint32_t idesc1 = make_idesc_bf16(128, 16, 1, 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + MAXTAPS * BN, make_smem_desc(sa + k * a_step, a_lbo, 1024),
                      make_smem_desc(sOnes + k * 32, 0, 1024), idesc1, (it | k) ? 1u : 0u);
        }
        if constexpr (CL > 1) umma_commit_mc(bar_empty + 8 * stage, CL_MASK);
        else umma_commit(bar_empty + 8 * stage);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(bar_tfull);
    }
  } else {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    const bool valid = m < p.M;
    if (kiters > 0) {
      mbar_wait(bar_tfull, 0);
      tc_fence_after();
    }
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    if (do_colsum && blockIdx.y == 0) {
      uint32_t r[16];
      if (kiters > 0) {
        tmem_ld_32x16(trow + MAXTAPS * BN, r);
        tmem_ld_wait();
      } else {
        r[0] = 0u;
      }
      if (valid) p.colsum[static_cast<long long>(blockIdx.z) * p.M + m] = __uint_as_float(r[0]);
    }
    if (p.tma_out) {
      // A thread owns one output row, so its 16 B stores hit 32 different lines per warp instruction (32 LSU
      // wavefronts each): for the short contractions of the attention block (K = 256: 4 k-blocks) the 2048 store
      // wavefronts per warp of a 128 x 256 fp32 tile cost twice the main loop.  Instead each tap's tile is staged
      // in the idle operand ring as 16 KB chunks in the 128 B-swizzled layout of a TMA box {128 B of columns,
      // 128 rows} and leaves as one bulk tensor store per chunk (rows / columns past M / N are clipped by the map).
      const int ml = q * 32 + lane;
      const uint32_t swz = static_cast<uint32_t>(ml & 7);
      const bool leader = threadIdx.x == 64;
      uint8_t* stage0 = smem_raw + (base - smem_u32(smem_raw));
      const int cpc = p.out_f32 ? 32 : 64;                       // columns per 16 KB chunk
      for (int t = 0; t < ntl; ++t) {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          if (kiters > 0) {
            tmem_ld_32x32(trow + t * BN + c0, r);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = 0u;
          }
          if (p.out_f32) {
            uint8_t* row = stage0 + (c0 >> 5) * 16384 + ml * 128;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<float4*>(row + ((static_cast<uint32_t>(i) ^ swz) << 4)) =
                  make_float4(__uint_as_float(r[4 * i]) * p.alpha, __uint_as_float(r[4 * i + 1]) * p.alpha,
                              __uint_as_float(r[4 * i + 2]) * p.alpha, __uint_as_float(r[4 * i + 3]) * p.alpha);
          } else {
            uint8_t* row = stage0 + (c0 >> 6) * 16384 + ml * 128;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(r[8 * i + 0]) * p.alpha, __uint_as_float(r[8 * i + 1]) * p.alpha);
              u.y = pack_bf16x2(__uint_as_float(r[8 * i + 2]) * p.alpha, __uint_as_float(r[8 * i + 3]) * p.alpha);
              u.z = pack_bf16x2(__uint_as_float(r[8 * i + 4]) * p.alpha, __uint_as_float(r[8 * i + 5]) * p.alpha);
              u.w = pack_bf16x2(__uint_as_float(r[8 * i + 6]) * p.alpha, __uint_as_float(r[8 * i + 7]) * p.alpha);
              *reinterpret_cast<uint4*>(row + (((static_cast<uint32_t>((c0 & 63) >> 3) + i) ^ swz) << 4)) = u;
            }
          }
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (leader) {
          for (int ch = 0; ch * cpc < BN; ++ch)
            if (n0 + ch * cpc < p.N)
              tma_store_4d(&tmO, base + ch * 16384, n0 + ch * cpc, m0, tap0 + t, static_cast<int>(blockIdx.z));
          tma_store_commit();
          tma_store_wait_read<0>();                              // the staging chunks are rewritten by the next tap
        }
        if (t + 1 < ntl) asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    } else
    for (int t = 0; t < ntl; ++t) {
      const long long ooff = blockIdx.z * p.oZ + (tap0 + t) * p.oT + static_cast<long long>(m) * p.oM;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        if (kiters > 0) {
          tmem_ld_32x32(trow + t * BN + c0, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
        const int col = n0 + c0;
        if (!valid || col >= p.N) continue;
        if (col + 32 > p.N) {                       // ragged last columns: element-wise
          for (int i = 0; i < 32 && col + i < p.N; ++i) {
            const float o = __uint_as_float(r[i]) * p.alpha;
            if (p.out_f32) static_cast<float*>(p.out)[ooff + col + i] = o;
            else static_cast<__nv_bfloat16*>(p.out)[ooff + col + i] = __float2bfloat16_rn(o);
          }
          continue;
        }
        if (p.out_f32) {
          float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + ooff + col);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            op[i] = make_float4(__uint_as_float(r[4 * i]) * p.alpha,
                                __uint_as_float(r[4 * i + 1]) * p.alpha,
                                __uint_as_float(r[4 * i + 2]) * p.alpha,
                                __uint_as_float(r[4 * i + 3]) * p.alpha);
        } else {
          uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + ooff + col);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(r[8 * i + 0]) * p.alpha, __uint_as_float(r[8 * i + 1]) * p.alpha);
            u.y = pack_bf16x2(__uint_as_float(r[8 * i + 2]) * p.alpha, __uint_as_float(r[8 * i + 3]) * p.alpha);
            u.z = pack_bf16x2(__uint_as_float(r[8 * i + 4]) * p.alpha, __uint_as_float(r[8 * i + 5]) * p.alpha);
            u.w = pack_bf16x2(__uint_as_float(r[8 * i + 6]) * p.alpha, __uint_as_float(r[8 * i + 7]) * p.alpha);
            op[i] = u;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();        // no CTA leaves while a peer may still arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// partial[split][tap][co][ci] (fp32)  ->  dW[co][ci][tap] (fp32, OIHW with tap = r*S+s), +=
// The splits are added in index order (deterministic); eight loads are in flight per thread before the first add.
// Blocks past the weight range (colsum != null) reduce the per-split column sums of dY into the bias gradient db, so
// a convolution's weight and bias gradients are finished by ONE launch.
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                    int splits, int taps, int cout, int cin, int accumulate,
                                    const float* __restrict__ colsum, float* __restrict__ db, int wblocks) {
  if (static_cast<int>(blockIdx.x) >= wblocks) {
    const int m = (blockIdx.x - wblocks) * blockDim.x + threadIdx.x;
    if (m >= cout) return;
    float a = 0.f;
    for (int sp = 0; sp < splits; ++sp) a += colsum[static_cast<long long>(sp) * cout + m];
    db[m] = a;
    return;
  }
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long per = static_cast<long long>(taps) * cout * cin;
  if (i >= per) return;
  const int ci = static_cast<int>(i % cin);
  const int co = static_cast<int>((i / cin) % cout);
  const int t = static_cast<int>(i / (static_cast<long long>(cin) * cout));
  float s = 0.f;
  int sp = 0;
  for (; sp + 8 <= splits; sp += 8) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldcs(partial + (sp + k) * per + i);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
  }
  for (; sp < splits; ++sp) s += __ldcs(partial + sp * per + i);
  float* o = dw + (static_cast<long long>(co) * cin + ci) * taps + t;
  *o = accumulate ? (*o + s) : s;
}

// out[m] = sum_splits colsum[split][m]  (ordered: deterministic)
__global__ void colsum_reduce_kernel(const float* __restrict__ part, float* out, int splits, int M) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float a = 0.f;
  for (int s = 0; s < splits; ++s) a += part[static_cast<long long>(s) * M + m];
  out[m] = a;
}

template <int BN, int STAGES, int MAXTAPS, bool STRIP = false, int CL = 1>
static int launch_mm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, MmParams p, dim3 grid,
                     cudaStream_t stream) {
  using Cfg = MmCfg<BN, STAGES, MAXTAPS, STRIP>;
  static unsigned long long attr_mask = 0;
  if (int e = set_max_smem_once(mmgemm_kernel<BN, STAGES, MAXTAPS, STRIP, CL>, Cfg::SMEM, attr_mask)) return e;
  if constexpr (CL > 1) {
    p.tma_out = 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = Cfg::SMEM; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, mmgemm_kernel<BN, STAGES, MAXTAPS, STRIP, CL>, tmA, tmB, tmO, p);
  }
  mmgemm_kernel<BN, STAGES, MAXTAPS, STRIP, CL><<<grid, 192, Cfg::SMEM, stream>>>(tmA, tmB, tmO, p);
  return (int)cudaGetLastError();
}

}  // namespace b2

using namespace b2;

extern "C" {

struct b2dq_mm_desc {
  // 5-D bf16 views (innermost first) of the two operands; strides in elements (stride[0] ignored)
  const void* a_ptr; long long a_dims[5]; long long a_strides[5];
  const void* b_ptr; long long b_dims[5]; long long b_strides[5];
  int a_mn, b_mn;
  int ntaps;          // total taps (1..12)
  int taps_per_cta;   // accumulators per CTA (1..3)
  int tap_c[12], tap_w[12], tap_p[12], tap_h[12];
  int KW, KH, KN;
  int ktiles_w, ktiles_h, kblocks;
  int splits, batches;
  int M, N;
  void* out; long long oZ, oT, oM;
  float alpha;
  int out_f32;
  int block_n;   // 128 or 256 (0 = auto)
  int b_strip;   // 1: the taps of a CTA are 1-pixel horizontal shifts -> one 66-pixel strip load per k-block
  float* colsum; // optional [splits][M] scratch: per-split sums of A over the contraction (bias gradient)
};

int b2dq_mmgemm(const b2dq_mm_desc* d, cudaStream_t stream) {
  if (!d || d->ntaps < 1 || d->ntaps > 12 || d->splits < 1 || d->batches < 1) return -1;
  const int tpc = d->taps_per_cta > 0 ? d->taps_per_cta : (d->ntaps < 3 ? d->ntaps : 3);
  if (tpc > 3) return -1;
  if (d->M <= 0 || d->N <= 0) return 0;
  int bn = d->block_n ? d->block_n : ((d->N % 256 == 0 && tpc == 1) ? 256 : 128);
  if (bn * tpc > 512) return -2;
  if ((d->a_mn || d->b_mn) && d->KW * d->KH * d->KN != 64) return -3;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[5], str[5];
    for (int i = 0; i < 5; ++i) { dims[i] = (uint64_t)d->a_dims[i]; str[i] = (uint64_t)d->a_strides[i]; }
    uint32_t box_k[5] = {64, 128, 1, 1, 1};
    uint32_t box_mn[5] = {64, (uint32_t)d->KW, 1, (uint32_t)d->KH, (uint32_t)d->KN};
    int r = make_tmap_bf16(&tmA, d->a_ptr, 5, dims, str, d->a_mn ? box_mn : box_k);
    if (r) return r;
  }
  {
    uint64_t dims[5], str[5];
    for (int i = 0; i < 5; ++i) { dims[i] = (uint64_t)d->b_dims[i]; str[i] = (uint64_t)d->b_strides[i]; }
    uint32_t box_k[5] = {64, (uint32_t)bn, 1, 1, 1};
    uint32_t box_mn[5] = {64, (uint32_t)d->KW + (d->b_strip ? 2u : 0u), 1, (uint32_t)d->KH, (uint32_t)d->KN};
    int r = make_tmap_bf16(&tmB, d->b_ptr, 5, dims, str, d->b_mn ? box_mn : box_k);
    if (r) return r - 1000;
  }
  MmParams p;
  p.a_mn = d->a_mn; p.b_mn = d->b_mn; p.ntaps = d->ntaps; p.taps_per_cta = tpc;
  const int ngroups = (d->ntaps + tpc - 1) / tpc;
  for (int i = 0; i < 12; ++i) {
    p.tap_c[i] = d->tap_c[i]; p.tap_w[i] = d->tap_w[i]; p.tap_p[i] = d->tap_p[i]; p.tap_h[i] = d->tap_h[i];
  }
  p.KW = d->KW ? d->KW : 64; p.KH = d->KH ? d->KH : 1; p.KN = d->KN ? d->KN : 1;
  p.ktiles_w = d->ktiles_w > 0 ? d->ktiles_w : (1 << 30);
  p.ktiles_h = d->ktiles_h > 0 ? d->ktiles_h : 1;
  p.kblocks = d->kblocks; p.splits = d->splits;
  p.M = d->M; p.N = d->N;
  p.out = d->out; p.oZ = d->oZ; p.oT = d->oT; p.oM = d->oM;
  p.alpha = d->alpha; p.out_f32 = d->out_f32;
  p.colsum = (tpc == 3 && d->a_mn) ? d->colsum : nullptr;
  if (d->colsum && !p.colsum) return -7;
  dim3 grid((unsigned)(((d->M + 127) / 128) * ngroups), (unsigned)((d->N + bn - 1) / bn),
            (unsigned)(d->batches * d->splits));
  // output through shared memory + bulk tensor stores when every stride is 16 B aligned
  static const bool tma_out_enabled = [] {
    const char* e = getenv("B2DQ_MMGEMM_TMA_OUT");
    return !(e && e[0] == '0');
  }();
  CUtensorMap tmO = tmA;
  p.tma_out = 0;
  {
    const long long esz = d->out_f32 ? 4 : 2, al = 16 / esz;
    const long long nz = (long long)d->batches * d->splits;
    const long long sT = d->ntaps > 1 ? d->oT : d->oM * d->M, sZ = nz > 1 ? d->oZ : d->oM * d->M;
    if (tma_out_enabled && d->oM % al == 0 && sT % al == 0 && sZ % al == 0 && sT > 0 && sZ > 0 &&
        (reinterpret_cast<uintptr_t>(d->out) & 15) == 0) {
      uint64_t dims[4] = {(uint64_t)d->N, (uint64_t)d->M, (uint64_t)d->ntaps, (uint64_t)nz};
      uint64_t str[4] = {1, (uint64_t)d->oM, (uint64_t)sT, (uint64_t)sZ};
      uint32_t box[4] = {d->out_f32 ? 32u : 64u, 128, 1, 1};
      int r = make_tmap_elem(&tmO, d->out, 4, dims, str, box, d->out_f32 != 0);
      if (r) return r - 2000;
      p.tma_out = 1;
    }
  }
  if (d->b_strip) {
    if (!(d->a_mn && d->b_mn) || d->KW != 64 || d->KH != 1 || d->KN != 1 || bn != 128 || tpc != 3 ||
        d->ntaps % 3 != 0)
      return -6;
    // opt-in (B2DQ_WGRAD_CLUSTER=1): clusters of 3 (the three filter rows of a pixel range) share the dY tile by TMA
    // multicast.  Correct, but SLOWER than three unicast loads: L2 already serves the near-simultaneous identical
    // requests of the three CTAs once, and the lock-step release of a stage by three MMA issuers adds stalls.
    static int use_cluster = -1;
    if (use_cluster < 0) {
      const char* e = getenv("B2DQ_WGRAD_CLUSTER");
      use_cluster = e ? atoi(e) : 0;   // measured on B200: 0.845 ms clustered vs 0.609 ms unicast
    }
    if (use_cluster && ngroups == 3) return launch_mm<128, 5, 3, true, 3>(tmA, tmB, tmO, p, grid, stream);
    return launch_mm<128, 5, 3, true>(tmA, tmB, tmO, p, grid, stream);
  }
  if (bn == 128) {
    if (tpc == 1) return launch_mm<128, 4, 1>(tmA, tmB, tmO, p, grid, stream);
    return launch_mm<128, 3, 3>(tmA, tmB, tmO, p, grid, stream) ;
  } else if (bn == 256) {
    if (tpc != 1) return -4;
    return launch_mm<256, 4, 1>(tmA, tmB, tmO, p, grid, stream);
  }
  return -5;
}

int b2dq_colsum_reduce(const float* part, float* out, int splits, int M, cudaStream_t stream) {
  if (M <= 0) return 0;
  colsum_reduce_kernel<<<(M + 127) / 128, 128, 0, stream>>>(part, out, splits, M);
  return (int)cudaGetLastError();
}

int b2dq_wgrad_reduce(const float* partial, float* dw, int splits, int taps, int cout, int cin,
                      int accumulate, cudaStream_t stream) {
  const long long per = (long long)taps * cout * cin;
  if (per <= 0) return 0;
  const int wblocks = (int)((per + 255) / 256);
  wgrad_reduce_kernel<<<(unsigned)wblocks, 256, 0, stream>>>(partial, dw, splits, taps, cout, cin, accumulate,
                                                            nullptr, nullptr, wblocks);
  return (int)cudaGetLastError();
}

// b2dq_wgrad_reduce + b2dq_colsum_reduce(colsum -> db [cout]) in one launch.
int b2dq_wgrad_reduce_bias(const float* partial, float* dw, int splits, int taps, int cout, int cin,
                           const float* colsum, float* db, cudaStream_t stream) {
  const long long per = (long long)taps * cout * cin;
  if (per <= 0) return 0;
  if (!colsum || !db) return -1;
  const int wblocks = (int)((per + 255) / 256);
  wgrad_reduce_kernel<<<(unsigned)(wblocks + (cout + 255) / 256), 256, 0, stream>>>(partial, dw, splits, taps, cout,
                                                                                   cin, 0, colsum, db, wblocks);
  return (int)cudaGetLastError();
}

}  // extern "C"
