// Vector-quantization bottleneck of DQ-VAE stage 1, fused for sm_100a.
//
// Replaces, in one pass over the latent, what the reference does with ~25 ATen launches
// (modules/vector_quantization/quantize2_mask.py):
//   :29-48  compute_distances   ||e||^2 - 2 x.e^T  (row constant ||x||^2 dropped: argmin-neutral)
//   :50-55  find_nearest_embedding (argmin, lowest index on ties)
//   :66-84  one-hot scatter + onehot@x   -> per-code counts and vector sums (red.add)
//   :123    embed (gather of the PRE-update codebook row)
//   :172-179 masked commitment loss       -> sum_rows m * sum_c (e - x)^2
// The [N,K] distance matrix never leaves the SM: x.e^T tiles are produced by tcgen05.mma
// into TMEM (two 256-column buffers) and reduced to a running (min, argmin) per row by
// the search warps while the tensor core works on the next codebook tile.
//
// CTA = 576 threads, four roles that only meet through mbarriers:
//   warp 0        TMA producer: x tile in four 16 KB channel chunks (each chunk is re-loaded for the
//                 next tile as soon as the last MMA that reads it has retired; the next tile's rows
//                 are prefetched into L2 a tile ahead) + the codebook through a 4-stage 32 KB ring
//   warp 1        MMA issuer (+ TMEM owner)
//   warps 2..9    search: TMEM -> registers, d = ||e||^2 - 2 x.e, running (min, argmin); ||e||^2 of
//                 the NEXT codebook tile is fetched while the current one is reduced and read back
//                 from a per-warp shared-memory slot (no global-load latency inside the loop)
//   warps 10..17  gather: codes -> fp32 codebook rows, loss, EMA sums; works on tile t while the
//                 search warps and the tensor core are already on tile t+1
// Small N (residual quantizer, stage-2 sampling): the codebook is split over `splits` CTAs per row
// tile; the partial minima meet in a 64-bit atomicMin (ordered distance bits | index, so the lowest
// index still wins ties) and the last CTA of a tile to arrive runs its gather.
#include "common.cuh"
#include "tmap.h"

namespace b2 {

constexpr int VQ_BM = 128;        // latent rows per tile
constexpr int VQ_BN = 256;        // codebook entries per MMA tile
constexpr int VQ_STAGES = 4;      // codebook ring depth
constexpr int VQ_SEARCH_WARPS = 8;
constexpr int VQ_GATHER_WARPS = 8;
constexpr int VQ_THREADS = 32 * (2 + VQ_SEARCH_WARPS + VQ_GATHER_WARPS);
constexpr uint32_t VQ_A_CHUNK = VQ_BM * 128;   // 16 KB: 128 rows x 64 bf16
constexpr uint32_t VQ_B_CHUNK = VQ_BN * 128;   // 32 KB
constexpr uint32_t VQ_SCRATCH = 16384;         // barriers, argmin hand-off, ||e||^2 slots
constexpr uint32_t VQ_SMEM = 4 * VQ_A_CHUNK + VQ_STAGES * VQ_B_CHUNK + 1024 /*align*/ + VQ_SCRATCH;

struct VqParams {
  const __nv_bfloat16* x_bf16;   // [N,C] search operand (and loss/EMA operand when x_f32 == null)
  const float* x_f32;            // [N,C] optional fp32 copy of x for loss / EMA sums
  const float* weight_f32;       // [K(+1),C] fp32 codebook (pre-update), gather source
  const float* cb_sqnorm;        // [round_up(K,256)] ||e||^2 of the bf16 codebook, +inf padded
  const float* row_mask;         // [N] or null
  long long* codes;              // [N] int64
  __nv_bfloat16* xq_bf16;        // [N,C] or null
  float* xq_f32;                 // [N,C] or null
  float* loss_acc;               // [1] += sum m*(e-x)^2   (null: skip)
  float* counts;                 // [K] += 1 per assigned row   (null: no EMA accumulation)
  float* sums;                   // [K,C] += x row
  unsigned long long* keys;      // [tiles*128] split mode: running min of (ordered d | index), preset to ~0
  unsigned int* tile_done;       // [tiles]     split mode: arrival counter, preset to 0xFFFFFFFF
  int N, C, K;
  // Work schedule (VqSched): `full_rounds` whole row tiles per CTA (tile = round * gridDim.x + blockIdx.x), then
  // the remaining `tail_tiles` row tiles are cut stream-K style into runs of `tail_q` codebook tiles per CTA
  // (a run may cover the end of one row tile and the start of the next).  tail_q == codebook tiles per row
  // tile means "no split".  keys / tile_done are indexed by the tail-local tile.
  int full_rounds, tail_tiles, tail_q;
};

struct VqItem {
  int tile;      // row tile
  int j0, j1;    // codebook tiles [j0, j1) of 256 entries
  int nsplit;    // CTAs that share this row tile (1: this CTA owns it and gathers directly)
  int tl;        // tail-local tile index (keys / tile_done slot) when nsplit > 1
};
// The same pure function of (blockIdx, ordinal) in every role of the CTA.
struct VqSched {
  int G, b, R, nn, q, tailbase, U, i, u, uend;
  __device__ VqSched(const VqParams& p, int nn_)
      : G(gridDim.x), b(blockIdx.x), R(p.full_rounds), nn(nn_), q(p.tail_q), tailbase(p.full_rounds * gridDim.x),
        U(p.tail_tiles * nn_), i(0), u(0), uend(0) {}
  __device__ bool next(VqItem& it) {
    if (i < R) {
      it.tile = i * G + b; it.j0 = 0; it.j1 = nn; it.nsplit = 1; it.tl = 0;
      ++i;
      return true;
    }
    if (i == R) {                       // enter the tail: this CTA's run of codebook-tile units
      const long long u0 = static_cast<long long>(b) * q;
      u = u0 < U ? static_cast<int>(u0) : U;
      uend = min(U, u + q);
      ++i;
    }
    if (u >= uend) return false;
    const int tl = u / nn, j0 = u - tl * nn;
    const int j1 = min(nn, j0 + (uend - u));
    it.tile = tailbase + tl; it.j0 = j0; it.j1 = j1; it.tl = tl;
    it.nsplit = ((tl + 1) * nn - 1) / q - (tl * nn) / q + 1;
    u += j1 - j0;
    return true;
  }
};

__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* tm, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tm), "r"(c0),
               "r"(c1)
               : "memory");
}
// fp32 -> uint32 whose unsigned order is the float order (for the split-mode atomicMin key)
__device__ __forceinline__ uint32_t ordered_bits(float d) {
  const uint32_t u = __float_as_uint(d);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int VQ_RG = 4;          // rows a gather warp keeps in flight

// Gather / loss / EMA sums for VQ_RG consecutive rows (row0 ..) of the latent; a lane covers 8 channels
// (C <= 256).  Every load of the group is issued before the first use - rows past N re-read row N-1 and
// are masked at the stores only - so a group costs one memory round trip, not one per row.  The loss is
// accumulated per LANE (sum over its 8 channels, weighted by the row mask); the caller reduces it over
// the warp once at the end of the kernel, so there is no shuffle in here.
//   X32   the rows for loss / EMA sums come from x_f32 (else from x_bf16)
//   NEEDX loss or EMA sums requested (a plain eval gather never touches x)
template <bool X32, bool NEEDX>
__device__ __forceinline__ void vq_gather_group(const VqParams& p, const int (&idx)[VQ_RG], int row0,
                                                int lane, float& loss_lane) {
  const int c = lane * 8;
  if (c < p.C) {
    float4 ea[VQ_RG], eb[VQ_RG];
    float4 xa[(X32 && NEEDX) ? VQ_RG : 1], xb[(X32 && NEEDX) ? VQ_RG : 1];
    uint4 raw[(!X32 && NEEDX) ? VQ_RG : 1];
    float mk[VQ_RG];
#pragma unroll
    for (int u = 0; u < VQ_RG; ++u) {
      const int lrow = min(row0 + u, p.N - 1);
      mk[u] = 1.0f;
      if (NEEDX) {
        if (p.row_mask) mk[u] = __ldg(p.row_mask + lrow);
        const long long off = static_cast<long long>(lrow) * p.C + c;
        if (X32) {
          const float4* xp = reinterpret_cast<const float4*>(p.x_f32 + off);
          xa[(X32 && NEEDX) ? u : 0] = xp[0];
          xb[(X32 && NEEDX) ? u : 0] = xp[1];
        } else {
          raw[(!X32 && NEEDX) ? u : 0] = *reinterpret_cast<const uint4*>(p.x_bf16 + off);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < VQ_RG; ++u) {
      const float4* ep = reinterpret_cast<const float4*>(p.weight_f32 + static_cast<long long>(idx[u]) * p.C + c);
      ea[u] = __ldg(ep);
      eb[u] = __ldg(ep + 1);
    }
#pragma unroll
    for (int u = 0; u < VQ_RG; ++u) {
      const bool live = row0 + u < p.N;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (NEEDX) {
        if (X32) {
          va = xa[(X32 && NEEDX) ? u : 0];
          vb = xb[(X32 && NEEDX) ? u : 0];
        } else {
          const uint4 w4 = raw[(!X32 && NEEDX) ? u : 0];
          va = make_float4(bf16_lo(w4.x), bf16_hi(w4.x), bf16_lo(w4.y), bf16_hi(w4.y));
          vb = make_float4(bf16_lo(w4.z), bf16_hi(w4.z), bf16_lo(w4.w), bf16_hi(w4.w));
        }
        if (p.loss_acc) {
          const float d0 = ea[u].x - va.x, d1 = ea[u].y - va.y, d2 = ea[u].z - va.z, d3 = ea[u].w - va.w;
          const float d4 = eb[u].x - vb.x, d5 = eb[u].y - vb.y, d6 = eb[u].z - vb.z, d7 = eb[u].w - vb.w;
          const float s0 = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
          const float s1 = d4 * d4 + d5 * d5 + d6 * d6 + d7 * d7;
          loss_lane += (live ? mk[u] : 0.f) * (s0 + s1);
        }
      }
      if (!live) continue;
      const long long off = static_cast<long long>(row0 + u) * p.C + c;
      if (p.xq_f32) {
        float4* op = reinterpret_cast<float4*>(p.xq_f32 + off);
        op[0] = ea[u];
        op[1] = eb[u];
      }
      if (p.xq_bf16) {
        uint4 o;
        o.x = pack_bf16x2(ea[u].x, ea[u].y);
        o.y = pack_bf16x2(ea[u].z, ea[u].w);
        o.z = pack_bf16x2(eb[u].x, eb[u].y);
        o.w = pack_bf16x2(eb[u].z, eb[u].w);
        *reinterpret_cast<uint4*>(p.xq_bf16 + off) = o;
      }
      if (NEEDX && p.sums) {
        float* dst = p.sums + static_cast<long long>(idx[u]) * p.C + c;
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(va.x), "f"(va.y), "f"(va.z),
                     "f"(va.w)
                     : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + 4), "f"(vb.x), "f"(vb.y),
                     "f"(vb.z), "f"(vb.w)
                     : "memory");
      }
    }
  }
  // lanes 0..3 publish the code (and the per-code count) of one row each
  if (lane < VQ_RG && row0 + lane < p.N) {
    const int mine = lane == 0 ? idx[0] : lane == 1 ? idx[1] : lane == 2 ? idx[2] : idx[3];
    p.codes[row0 + lane] = mine;
    if (p.counts) atomicAdd(p.counts + mine, 1.0f);
  }
}

__global__ void __launch_bounds__(VQ_THREADS, 1)
vq_search_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const VqParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + 4 * VQ_A_CHUNK;
  const uint32_t sX = sB + VQ_STAGES * VQ_B_CHUNK;  // barriers + scratch
  uint8_t* gen = smem_raw + (sX - smem_u32(smem_raw));
  // barrier slots (8 B each)
  const uint32_t bar_full = sX;                          // [VQ_STAGES] codebook stage landed
  const uint32_t bar_empty = bar_full + 8 * VQ_STAGES;   // [VQ_STAGES] codebook stage consumed
  const uint32_t bar_afull = bar_empty + 8 * VQ_STAGES;  // [4] x chunk landed
  const uint32_t bar_aempty = bar_afull + 32;            // [4] x chunk no longer read by any MMA
  const uint32_t bar_tfull = bar_aempty + 32;            // [2] accumulator buffer complete
  const uint32_t bar_tempty = bar_tfull + 16;            // [2] accumulator buffer drained
  const uint32_t bar_cfull = bar_tempty + 16;            // [2] argmin of a tile published
  const uint32_t bar_cempty = bar_cfull + 16;            // [2] argmin slot consumed by the gather warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + 256);
  volatile int* s_flag = reinterpret_cast<volatile int*>(gen + 272);   // [2] split mode: "this CTA gathers"
  float* s_best = reinterpret_cast<float*>(gen + 512);                  // [2 parity][2 half][128]
  int* s_idx = reinterpret_cast<int*>(gen + 512 + 2048);                // [2][2][128]
  float* s_sq = reinterpret_cast<float*>(gen + 512 + 4096);             // [8 warps][2 slots][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KC = p.C >> 6;
  const int num_m_tiles = (p.N + VQ_BM - 1) / VQ_BM;
  const int num_n_tiles = (p.K + VQ_BN - 1) / VQ_BN;

  if (threadIdx.x == 0) {
    for (int i = 0; i < VQ_STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar_afull + 8 * i, 1);
      mbar_init(bar_aempty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, VQ_SEARCH_WARPS);
      mbar_init(bar_cfull + 8 * i, VQ_SEARCH_WARPS);
      mbar_init(bar_cempty + 8 * i, VQ_GATHER_WARPS);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------ TMA producer
      uint32_t stage = 0, phase = 0, it = 0;
      VqSched sched(p, num_n_tiles);
      VqItem wi;
      for (; sched.next(wi); ++it) {
        const int tile = wi.tile, j0 = wi.j0, j1 = wi.j1;
        {  // rows of the next work item: HBM -> L2 now, L2 -> shared memory when the chunk frees up
          VqSched peek = sched;
          VqItem nx;
          if (peek.next(nx) && nx.tile != tile)
            for (int kc = 0; kc < KC; ++kc) tma_prefetch_l2_2d(&tmA, kc * 64, nx.tile * VQ_BM);
        }
        for (int j = j0; j < j1; ++j)
          for (int kc = 0; kc < KC; ++kc) {
            if (j == j0) {
              mbar_wait(bar_aempty + 8 * kc, (it & 1) ^ 1);
              mbar_arrive_expect_tx(bar_afull + 8 * kc, VQ_A_CHUNK);
              tma_load_2d(sA + kc * VQ_A_CHUNK, &tmA, bar_afull + 8 * kc, kc * 64, tile * VQ_BM);
            }
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            mbar_arrive_expect_tx(bar_full + 8 * stage, VQ_B_CHUNK);
            tma_load_2d(sB + stage * VQ_B_CHUNK, &tmB, bar_full + 8 * stage, kc * 64, j * VQ_BN);
            if (++stage == VQ_STAGES) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------ MMA issuer
      constexpr uint32_t idesc = make_idesc_bf16(VQ_BM, VQ_BN, 0, 0);
      uint32_t stage = 0, phase = 0, it = 0, jj = 0;
      VqSched sched(p, num_n_tiles);
      VqItem wi;
      for (; sched.next(wi); ++it) {
        const int j0 = wi.j0, j1 = wi.j1;
        for (int j = j0; j < j1; ++j, ++jj) {
          const uint32_t buf = jj & 1;
          mbar_wait(bar_tempty + 8 * buf, ((jj >> 1) & 1) ^ 1);
          tc_fence_after();
          for (int kc = 0; kc < KC; ++kc) {
            if (j == j0) mbar_wait(bar_afull + 8 * kc, it & 1);
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = make_smem_desc(sA + kc * VQ_A_CHUNK + k * 32, 0, 1024);
              const uint64_t db = make_smem_desc(sB + stage * VQ_B_CHUNK + k * 32, 0, 1024);
              umma_bf16(tmem_base + buf * VQ_BN, da, db, idesc, (kc | k) ? 1u : 0u);
            }
            umma_commit(bar_empty + 8 * stage);
            if (j == j1 - 1) umma_commit(bar_aempty + 8 * kc);   // last reader of this x chunk
            if (++stage == VQ_STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(bar_tfull + 8 * buf);
        }
      }
    }
  } else if (warp < 2 + VQ_SEARCH_WARPS) {
    // ------------------------------------------------ search: running argmin over the codebook
    const int ew = warp - 2;          // 0..7
    const int q = warp & 3;           // TMEM lane quadrant this warp may read
    const int half = ew >> 2;         // which 128 columns of each 256-column tile
    const int row = q * 32 + lane;    // row of the tile owned by this thread
    float* my_sq = s_sq + ew * 256;   // two slots of 128 squared norms, private to this warp
    uint32_t jj = 0, it = 0;
    VqSched sched(p, num_n_tiles);
    VqItem wi;
    {
      VqSched peek = sched;
      VqItem first;
      if (peek.next(first)) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.cb_sqnorm + first.j0 * VQ_BN + half * 128) + lane);
        *reinterpret_cast<float4*>(my_sq + lane * 4) = v;
      }
      __syncwarp();
    }
    for (; sched.next(wi); ++it) {
      const int tile = wi.tile, j0 = wi.j0, j1 = wi.j1;
      int jnext_item = -1;
      {
        VqSched peek = sched;
        VqItem nx;
        if (peek.next(nx)) jnext_item = nx.j0;
      }
      float best = __int_as_float(0x7f800000);
      int bi = 0;
      for (int j = j0; j < j1; ++j, ++jj) {
        const uint32_t buf = jj & 1;
        // squared norms of the codebook tile after this one: in flight while this one is reduced
        const int nj = (j + 1 < j1) ? j + 1 : jnext_item;
        float4 nsq = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nj >= 0) nsq = __ldg(reinterpret_cast<const float4*>(p.cb_sqnorm + nj * VQ_BN + half * 128) + lane);
        mbar_wait(bar_tfull + 8 * buf, (jj >> 1) & 1);
        tc_fence_after();
        const uint32_t tcol = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * VQ_BN +
                              half * 128;
        const int colbase = j * VQ_BN + half * 128;
        const float* sq = my_sq + (jj & 1) * 128;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tcol + c0, r);
          float4 s[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) s[i] = *reinterpret_cast<const float4*>(sq + c0 + 4 * i);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float d0 = fmaf(-2.f, __uint_as_float(r[4 * i + 0]), s[i].x);
            const float d1 = fmaf(-2.f, __uint_as_float(r[4 * i + 1]), s[i].y);
            const float d2 = fmaf(-2.f, __uint_as_float(r[4 * i + 2]), s[i].z);
            const float d3 = fmaf(-2.f, __uint_as_float(r[4 * i + 3]), s[i].w);
            const int c = colbase + c0 + 4 * i;
            if (d0 < best) { best = d0; bi = c; }
            if (d1 < best) { best = d1; bi = c + 1; }
            if (d2 < best) { best = d2; bi = c + 2; }
            if (d3 < best) { best = d3; bi = c + 3; }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        if (nj >= 0) *reinterpret_cast<float4*>(my_sq + ((jj + 1) & 1) * 128 + lane * 4) = nsq;
        __syncwarp();
      }
      // publish (min, argmin) of this work item to the gather warps
      const int par = it & 1;
      mbar_wait(bar_cempty + 8 * par, ((it >> 1) & 1) ^ 1);
      if (wi.nsplit == 1) {
        s_best[(par * 2 + half) * 128 + row] = best;
        s_idx[(par * 2 + half) * 128 + row] = bi;
      } else {
        if (static_cast<long long>(tile) * VQ_BM + row < p.N)
          atomicMin(p.keys + static_cast<long long>(wi.tl) * VQ_BM + row,
                    (static_cast<unsigned long long>(ordered_bits(best)) << 32) | static_cast<unsigned int>(bi));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_cfull + 8 * par);
    }
  } else {
    // ------------------------------------------------ gather / loss / EMA accumulation
    // Each warp takes 16 rows of the tile, four at a time (all loads of a group are issued before
    // their first use); a lane covers 8 consecutive channels.
    const int gw = warp - (2 + VQ_SEARCH_WARPS);   // 0..7
    uint32_t it = 0;
    float loss_local = 0.f;
    VqSched sched(p, num_n_tiles);
    VqItem wi;
    for (; sched.next(wi); ++it) {
      const int tile = wi.tile;
      const int S = wi.nsplit;
      const int par = it & 1;
      mbar_wait(bar_cfull + 8 * par, (it >> 1) & 1);
      bool gather_here = true;
      if (S > 1) {
        // the CTA that completes a tile's set of codebook splits gathers for it
        if (gw == 0 && lane == 0) {
          __threadfence();
          const unsigned int prev = atomicAdd(p.tile_done + wi.tl, 1u);   // preset 0xFFFFFFFF
          s_flag[par] = (prev == static_cast<unsigned int>(S - 2)) ? 1 : 0;
          __threadfence();
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        gather_here = s_flag[par] != 0;
      }
      if (gather_here) {
        for (int rr0 = 0; rr0 < 16; rr0 += VQ_RG) {
          int idx[VQ_RG];
          const int r_in0 = gw * 16 + rr0;
          const int row0 = tile * VQ_BM + r_in0;
#pragma unroll
          for (int u = 0; u < VQ_RG; ++u) {
            const int r_in = r_in0 + u;
            if (S == 1) {
              const float b0 = s_best[(par * 2 + 0) * 128 + r_in];
              const float b1 = s_best[(par * 2 + 1) * 128 + r_in];
              const int i0 = s_idx[(par * 2 + 0) * 128 + r_in];
              const int i1 = s_idx[(par * 2 + 1) * 128 + r_in];
              idx[u] = (b1 < b0) ? i1 : i0;   // lower column half wins ties
            } else {
              idx[u] = (row0 + u < p.N)
                           ? static_cast<int>(__ldcg(p.keys + static_cast<long long>(wi.tl) * VQ_BM + r_in) & 0xFFFFFFFFull)
                           : 0;
            }
            if (idx[u] >= p.K || idx[u] < 0) idx[u] = 0;
          }
          if (!p.loss_acc && !p.sums) vq_gather_group<false, false>(p, idx, row0, lane, loss_local);
          else if (p.x_f32) vq_gather_group<true, true>(p, idx, row0, lane, loss_local);
          else vq_gather_group<false, true>(p, idx, row0, lane, loss_local);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_cempty + 8 * par);
    }
    if (p.loss_acc) {
      loss_local = warp_sum(loss_local);
      if (lane == 0 && loss_local != 0.f) atomicAdd(p.loss_acc, loss_local);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------
// Codebook preparation: fp32 weight[:K] -> bf16 copy + ||bf16(e)||^2 (fp32), +inf padded
// to a multiple of 256 so the search epilogue needs no bounds check.
// (quantize2_mask.py:31,38: codebook = weight[:-1]; norms of the operand actually multiplied.)
__global__ void vq_prepare_codebook_kernel(const float* __restrict__ w, __nv_bfloat16* cb,
                                           float* sqn, int K, int C, int Kpad) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= Kpad) return;
  if (warp >= K) {
    if (lane == 0) sqn[warp] = __int_as_float(0x7f800000);
    return;
  }
  float s = 0.f;
  for (int c = lane; c < C; c += 32) {
    const __nv_bfloat16 b = __float2bfloat16_rn(w[static_cast<long long>(warp) * C + c]);
    cb[static_cast<long long>(warp) * C + c] = b;
    const float f = __bfloat162float(b);
    s = fmaf(f, f, s);
  }
  s = warp_sum(s);
  if (lane == 0) sqn[warp] = s;
}

// ---------------------------------------------------------------------------------
// EMA finalize, step 1 (single CTA): cluster_size_ema update, restart bookkeeping, n = sum.
// quantize2_mask.py:90 (EMA of counts), :102-105 (usage mask, dead codes reset to 1), :110.
__global__ void vq_ema_counts_kernel(const float* __restrict__ counts, float* cluster_size_ema,
                                     unsigned char* dead, float* n_out, int K, float decay,
                                     int restart) {
  __shared__ float red[32];
  float local = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float cs = cluster_size_ema[k] * decay + counts[k] * (1.f - decay);
    unsigned char d = 0;
    if (restart && !(cs >= 1.f)) { d = 1; cs = 1.f; }
    cluster_size_ema[k] = cs;
    dead[k] = d;
    local += cs;
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) *n_out = v;
  }
}
// step 2: embed_ema update (+restart rows), weight = embed_ema / smoothed cluster size.
// quantize2_mask.py:91,103 and :107-115.
__global__ void vq_ema_embed_kernel(const float* __restrict__ sums,
                                    const float* __restrict__ restart_rows,
                                    const unsigned char* __restrict__ dead,
                                    const float* __restrict__ cluster_size_ema,
                                    const float* __restrict__ n_in, float* embed_ema,
                                    float* weight, int K, int C, float decay, float eps) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(K) * C) return;
  const int k = static_cast<int>(i / C);
  float e = embed_ema[i] * decay + sums[i] * (1.f - decay);
  if (dead[k]) e = restart_rows[i];
  embed_ema[i] = e;
  const float n = *n_in;
  const float norm = n * (cluster_size_ema[k] + eps) / (n + K * eps);
  weight[i] = e / norm;
}

// VQ backward (quantize2_mask.py:172-182): g_x = g_xq (straight-through) +
//   g_loss * 2*beta/(N*C) * m * (x - e);   the (xq - sg(x))^2 term has no trainable input.
__global__ void vq_bwd_kernel(const __nv_bfloat16* __restrict__ g_xq,
                              const __nv_bfloat16* __restrict__ x,
                              const __nv_bfloat16* __restrict__ xq,
                              const float* __restrict__ row_mask, const float* __restrict__ g_loss,
                              float coef, __nv_bfloat16* g_x, long long n_rows, int C) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 2;
  if (i >= n_rows * C) return;
  const long long r = i / C;
  const float s = coef * g_loss[0] * (row_mask ? row_mask[r] : 1.f);
  const uint32_t gu = *reinterpret_cast<const uint32_t*>(g_xq + i);
  const uint32_t xu = *reinterpret_cast<const uint32_t*>(x + i);
  const uint32_t qu = *reinterpret_cast<const uint32_t*>(xq + i);
  const float o0 = bf16_lo(gu) + s * (bf16_lo(xu) - bf16_lo(qu));
  const float o1 = bf16_hi(gu) + s * (bf16_hi(xu) - bf16_hi(qu));
  *reinterpret_cast<uint32_t*>(g_x + i) = pack_bf16x2(o0, o1);
}

}  // namespace b2

// =================================================================================== C ABI
using namespace b2;

static int num_sms() { return device_sm_count(); }

extern "C" {

int b2dq_vq_prepare_codebook(const float* weight_f32, void* cb_bf16, float* cb_sqnorm, int K, int C,
                             cudaStream_t stream) {
  if (K <= 0 || C <= 0) return -1;
  const int Kpad = (K + 255) / 256 * 256;
  const int warps_per_block = 8;
  const int blocks = (Kpad + warps_per_block - 1) / warps_per_block;
  vq_prepare_codebook_kernel<<<blocks, warps_per_block * 32, 0, stream>>>(
      weight_f32, reinterpret_cast<__nv_bfloat16*>(cb_bf16), cb_sqnorm, K, C, Kpad);
  return (int)cudaGetLastError();
}

// Work schedule (see VqSched): whole row tiles for `rounds` full waves of the grid, then the leftover row tiles
// stream-K split into runs of `q` codebook tiles so that the last wave keeps every SM busy.  Covers both
// "512 row tiles on 148 SMs" (3 full rounds + 68 tiles cut in halves: 3.5 rounds instead of 4) and the small-N
// calls of the residual quantizer / stage-2 sampling (16 row tiles, 64 codebook tiles each -> 147 CTAs).
struct VqPlan {
  int grid, rounds, tail_tiles, q, split;   // split: some row tile is shared between CTAs (needs the workspace)
};
// codebook tiles (of 256 entries) a row tile needs before its tail may be shared between CTAs
static int vq_split_min_tiles() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B2DQ_VQ_SPLIT_MIN_TILES");
    v = e ? atoi(e) : 8;
    if (v < 2) v = 2;
  }
  return v;
}
static VqPlan vq_plan_for(int N, int K, int G, bool allow_split) {
  VqPlan pl;
  const int tiles = (N + VQ_BM - 1) / VQ_BM;
  const int nn = (K + VQ_BN - 1) / VQ_BN;
  if (G < 1) G = 1;
  pl.rounds = tiles / G;
  pl.tail_tiles = tiles - pl.rounds * G;
  pl.q = nn;
  pl.split = 0;
  // a split tail costs tail_tiles/G of a round (plus the re-load of shared row tiles and the atomics)
  // instead of a whole one: worth it when the tail leaves at least ~15 % of the SMs idle
  // ... and when a row tile has enough codebook tiles to amortise what a shared tile costs (second load of its
  // rows, 128 atomics per sharer, the workspace memset): measured on B200, K = 1024 (4 codebook tiles) loses
  // (N = 32768: 38.8 vs 25.7 us), K >= 8192 gains 8-15 % and small-N calls 3.5-6x
  if (allow_split && nn >= vq_split_min_tiles() && pl.tail_tiles > 0 && pl.tail_tiles * 20 <= G * 17) {
    const long long U = (long long)pl.tail_tiles * nn;
    int q = (int)((U + G - 1) / G);
    if (q < 1) q = 1;
    if (q < nn || q % nn) {
      pl.q = q;
      pl.split = 1;
    }
  }
  if (pl.rounds > 0) {
    pl.grid = G;
  } else {
    const long long U = (long long)pl.tail_tiles * nn;
    pl.grid = (int)((U + pl.q - 1) / pl.q);
  }
  return pl;
}

static VqPlan vq_plan(int N, int K, int max_ctas, bool allow_split) {
  int G = num_sms();
  if (max_ctas > 0 && max_ctas < G) G = max_ctas;
  return vq_plan_for(N, K, G, allow_split);
}

// The schedule a launch on `num_ctas` CTAs would use: out5 = {grid, full_rounds, tail_tiles, tail_q, split}.
// Host-only (no device needed): lets callers and tests inspect the stream-K cut.
int b2dq_vq_search_plan(int N, int K, int num_ctas, int allow_split, int* out5) {
  if (N <= 0 || K <= 0 || num_ctas <= 0 || !out5) return -1;
  const VqPlan pl = vq_plan_for(N, K, num_ctas, allow_split != 0);
  out5[0] = pl.grid; out5[1] = pl.rounds; out5[2] = pl.tail_tiles; out5[3] = pl.q; out5[4] = pl.split;
  return 0;
}

// Scratch for the shared row tiles: an upper bound that holds for any max_ctas (0: never needed).
int b2dq_vq_search_workspace_bytes(int N, int K) {
  if (N <= 0 || K <= (vq_split_min_tiles() - 1) * VQ_BN) return 0;
  const int tiles = (N + VQ_BM - 1) / VQ_BM;
  const int cap = tiles < 1024 ? tiles : 1024;   // the tail is shorter than one wave of the grid
  return cap * VQ_BM * 8 + cap * 4;
}

int b2dq_vq_search_gather(const void* x_bf16, const float* x_f32, const void* cb_bf16,
                          const float* cb_sqnorm, const float* weight_f32, const float* row_mask,
                          long long* codes, void* xq_bf16, float* xq_f32, float* loss_acc,
                          float* counts, float* sums, int N, int C, int K, int max_ctas,
                          void* workspace, int workspace_bytes, cudaStream_t stream) {
  if (N <= 0) return 0;
  if (C % 64 != 0 || C > 256 || C <= 0 || K <= 0) return -1;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)N};
    uint64_t str[2] = {1, (uint64_t)C};
    uint32_t box[2] = {64, VQ_BM};
    int r = make_tmap_bf16(&tmA, x_bf16, 2, dims, str, box);
    if (r) return r;
  }
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)K};
    uint64_t str[2] = {1, (uint64_t)C};
    uint32_t box[2] = {64, VQ_BN};
    int r = make_tmap_bf16(&tmB, cb_bf16, 2, dims, str, box);
    if (r) return r;
  }
  static unsigned long long attr_mask = 0;
  if (int e = set_max_smem_once(vq_search_kernel, VQ_SMEM, attr_mask)) return e;
  VqParams p;
  p.x_bf16 = reinterpret_cast<const __nv_bfloat16*>(x_bf16);
  p.x_f32 = x_f32;
  p.weight_f32 = weight_f32;
  p.cb_sqnorm = cb_sqnorm;
  p.row_mask = row_mask;
  p.codes = codes;
  p.xq_bf16 = reinterpret_cast<__nv_bfloat16*>(xq_bf16);
  p.xq_f32 = xq_f32;
  p.loss_acc = loss_acc;
  p.counts = counts;
  p.sums = sums;
  p.keys = nullptr;
  p.tile_done = nullptr;
  p.N = N; p.C = C; p.K = K;
  VqPlan pl = vq_plan(N, K, max_ctas, workspace != nullptr);
  if (pl.split) {
    const int need = pl.tail_tiles * VQ_BM * 8 + pl.tail_tiles * 4;
    if (workspace_bytes < need) return -3;
    if (reinterpret_cast<uintptr_t>(workspace) & 7) return -4;
    p.keys = reinterpret_cast<unsigned long long*>(workspace);
    p.tile_done = reinterpret_cast<unsigned int*>(p.keys + (size_t)pl.tail_tiles * VQ_BM);
    // keys start at the largest key, arrival counters at 0xFFFFFFFF (first arrival reads it back)
    cudaError_t e = cudaMemsetAsync(workspace, 0xFF, need, stream);
    if (e != cudaSuccess) return (int)e;
  }
  p.full_rounds = pl.rounds;
  p.tail_tiles = pl.tail_tiles;
  p.tail_q = pl.q;
  const int grid = pl.grid;
  vq_search_kernel<<<grid, VQ_THREADS, VQ_SMEM, stream>>>(tmA, tmB, p);
  return (int)cudaGetLastError();
}

int b2dq_vq_ema_finalize(const float* counts, const float* sums, const float* restart_rows,
                         float* cluster_size_ema, float* embed_ema, float* weight_f32,
                         unsigned char* dead_scratch, float* n_scratch, int K, int C, float decay,
                         float eps, int restart, cudaStream_t stream) {
  if (K <= 0 || C <= 0) return -1;
  if (restart && !restart_rows) return -2;
  vq_ema_counts_kernel<<<1, 1024, 0, stream>>>(counts, cluster_size_ema, dead_scratch, n_scratch, K,
                                               decay, restart);
  const long long total = (long long)K * C;
  const int blocks = (int)((total + 255) / 256);
  vq_ema_embed_kernel<<<blocks, 256, 0, stream>>>(sums, restart_rows, dead_scratch,
                                                  cluster_size_ema, n_scratch, embed_ema,
                                                  weight_f32, K, C, decay, eps);
  return (int)cudaGetLastError();
}

int b2dq_vq_bwd(const void* g_xq, const void* x, const void* xq, const float* row_mask,
                const float* g_loss, float coef, void* g_x, long long n_rows, int C,
                cudaStream_t stream) {
  if (n_rows <= 0) return 0;
  if (C % 2) return -1;
  const long long pairs = n_rows * C / 2;
  const int blocks = (int)((pairs + 255) / 256);
  vq_bwd_kernel<<<blocks, 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(g_xq), reinterpret_cast<const __nv_bfloat16*>(x),
      reinterpret_cast<const __nv_bfloat16*>(xq), row_mask, g_loss, coef,
      reinterpret_cast<__nv_bfloat16*>(g_x), n_rows, C);
  return (int)cudaGetLastError();
}

}  // extern "C"
