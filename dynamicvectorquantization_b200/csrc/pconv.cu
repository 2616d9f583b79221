// Persistent 3x3 (stride 1, pad 1) convolution for the wide, 128-output-channel layers of the
// DQ-VAE stacks (the 128->128 convolutions at 256x256 / 128x128 that carry 54 % of the model FLOPs;
// reference: modules/diffusionmodules/model.py:88-102 ResnetBlock conv1/conv2, :43-47 Upsample conv)
// - forward and data gradient.
//
// Why a second kernel next to tapgemm_kernel: at Cout = 128 the one-tile-per-CTA tap GEMM needs
// 128 B of L2->SM operand traffic per tensor-core cycle and saturates the L2 fabric at ~43 % tensor
// utilisation (profiles/r01_ncu_full_summary.md).  This kernel cuts the traffic per FLOP 2.4x:
//   * the three horizontal taps (s = 0,1,2) of a filter row read ONE activation strip of 130
//     pixels x 64 channels (TMA box {64,130}); tap s is the same shared-memory strip addressed from
//     row s (UMMA descriptor start + s*128 B; the 128 B swizzle is a function of absolute smem
//     address bits, so the shifted view needs no descriptor base offset);
//   * two 128-pixel output tiles share every weight tile (6 MMA groups per 3 weight tiles);
//   * the CTA is persistent (one per SM) with two TMEM accumulator sets, so the epilogue of one
//     tile pair overlaps the main loop of the next and the TMA ring never drains;
//   * the epilogue never touches global memory with per-thread accesses: a thread owns one pixel row of
//     the tile, so its 16 B stores / residual loads would hit 32 different 128 B lines per warp
//     instruction (8192 LSU wavefronts per tile pair with a residual - more than the 4608 tensor cycles of
//     the pair: measured 627 us vs 485 us per launch).  Instead the bf16 tile is staged in shared memory in
//     the 128 B-swizzled layout of a TMA box {64 ch, 128 px} and written by ONE bulk tensor store per
//     64-channel half; the residual half-tile arrives the same way (TMA load into the same staging buffer
//     two units ahead, added in place).  Three 16 KB staging buffers rotate.
// Per pipeline stage: 2 strips (2 x 16.6 KB) + 3 weight tiles (3 x 16 KB) feed 24 MMAs
// (128x128x16) = 1536 tensor cycles -> 53 B/cycle/SM instead of 128.
#include <cstring>
#include "common.cuh"
#include "tmap.h"

namespace b2 {

constexpr int PC_BN = 128;
constexpr int PC_MT = 2;
constexpr int PC_STAGES = 2;
constexpr uint32_t PC_STRIP_ROWS = 130;
constexpr uint32_t PC_STRIP_BYTES = PC_STRIP_ROWS * 128;           // 16,640
constexpr uint32_t PC_STRIP_SLOT = 17 * 1024;                      // padded, keeps 1024 B alignment
constexpr uint32_t PC_B_BYTES = PC_BN * 128;                       // 16 KB
constexpr uint32_t PC_STAGE_BYTES = PC_MT * PC_STRIP_SLOT + 3 * PC_B_BYTES;   // 83,968
constexpr uint32_t PC_OUT_BUFS = 3;                                // staging buffers of the epilogue
constexpr uint32_t PC_OUT_BYTES = 128 * 128;                       // 128 pixels x 64 channels bf16
constexpr uint32_t PC_SMEM = PC_STAGES * PC_STAGE_BYTES + PC_OUT_BUFS * PC_OUT_BYTES + 1024 + 256 + 2048;

constexpr int PC_MAX_ROWS = 9;
struct PconvParams {
  int kchunks;                 // channels per tap / 64
  int nr;                      // row taps: strips loaded per 64-channel chunk (3 filter rows for the 3x3 layers)
  int row_c[PC_MAX_ROWS];      // A-map coordinates of row tap r: channel base, parity plane, row offset
  int row_p[PC_MAX_ROWS];
  int row_dh[PC_MAX_ROWS];
  int strip_dw;                // the strip of an output tile starts at pixel ow0 + strip_dw
  int col_off[3];              // strip row (pixel) offset of column tap s: 0..2
  int wcol[PC_MAX_ROWS * 3];   // weight column base of tap (r, s)
  int tiles_w, H, NB, W;       // tiles per image row, image height, images, width
  int num_tiles;               // NB * H * tiles_w
  int Cout;
  const float* bias;
  int has_residual;            // the residual tensor (same shape as the output) is read through tmR
  float* gn_part;              // optional [num_tiles][32 groups][2]: per-tile (sum, sum of squares) of the
                               // OUTPUT per GroupNorm group of 4 channels (statistics for the next GroupNorm)
};

// NS = column taps served by one strip (3 for the 3x3 layers; 2 / 1 for the parity classes of the folded
// up-convolution and of the stride-2 data gradient, b2dq_pconv_taps).
template <int NS>
__global__ void __launch_bounds__(192, 1)
pconv3x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR,
                const __grid_constant__ PconvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sOut = base + PC_STAGES * PC_STAGE_BYTES;          // 1024 B aligned (stage size is a multiple)
  const uint32_t sBar = sOut + PC_OUT_BUFS * PC_OUT_BYTES;
  const uint32_t bar_full = sBar, bar_empty = sBar + 16, bar_tfull = sBar + 32, bar_tempty = sBar + 48;
  const uint32_t bar_res = sBar + 96;                               // [PC_OUT_BUFS] residual half-tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + (sBar + 64 - smem_u32(smem_raw)));
  float* bias_s = reinterpret_cast<float*>(smem_raw + (sBar + 256 - smem_u32(smem_raw)));
  float* gn_red = bias_s + 128;                  // [4 warps][64]
  uint8_t* out_stage = smem_raw + (sOut - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = (p.num_tiles + PC_MT - 1) / PC_MT;
  const int kiters = p.nr * p.kchunks;

  if (threadIdx.x == 0) {
    for (int i = 0; i < PC_STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 4);
    }
    for (int i = 0; i < (int)PC_OUT_BUFS; ++i) mbar_init(bar_res + 8 * i, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    if (p.has_residual) tma_prefetch_desc(&tmR);
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(tmem_slot));
  if (threadIdx.x >= 64) {
    const int t = threadIdx.x - 64;
    bias_s[t] = (p.bias && t < p.Cout) ? p.bias[t] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_coords = [&](int t, int& ow0, int& oh, int& n) {
    ow0 = (t % p.tiles_w) * 128;
    oh = (t / p.tiles_w) % p.H;
    n = t / (p.tiles_w * p.H);          // tiles past the end give n >= NB: loads read zeros, stores masked
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int ow0[PC_MT], oh[PC_MT], n[PC_MT];
#pragma unroll
        for (int j = 0; j < PC_MT; ++j) tile_coords(item * PC_MT + j, ow0[j], oh[j], n[j]);
        for (int r = 0; r < p.nr; ++r)
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            const uint32_t sa = base + stage * PC_STAGE_BYTES;
            const uint32_t fb = bar_full + 8 * stage;
            mbar_arrive_expect_tx(fb, PC_MT * PC_STRIP_BYTES + NS * PC_B_BYTES);
#pragma unroll
            for (int j = 0; j < PC_MT; ++j)
              tma_load_5d(sa + j * PC_STRIP_SLOT, &tmA, fb, p.row_c[r] + kc * 64, ow0[j] + p.strip_dw, p.row_p[r],
                          oh[j] + p.row_dh[r], n[j]);
#pragma unroll
            for (int s = 0; s < NS; ++s)
              tma_load_2d(sa + PC_MT * PC_STRIP_SLOT + s * PC_B_BYTES, &tmB, fb, p.wcol[r * 3 + s] + kc * 64, 0);
            if (++stage == PC_STAGES) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, PC_BN, 0, 0);
      uint32_t stage = 0, phase = 0, it_item = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it_item) {
        const uint32_t buf = it_item & 1;
        mbar_wait(bar_tempty + 8 * buf, ((it_item >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = base + stage * PC_STAGE_BYTES;
#pragma unroll
          for (int j = 0; j < PC_MT; ++j) {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
              const uint32_t a0 = sa + j * PC_STRIP_SLOT + p.col_off[s] * 128;
              const uint32_t b0 = sa + PC_MT * PC_STRIP_SLOT + s * PC_B_BYTES;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base + buf * (PC_MT * PC_BN) + j * PC_BN, make_smem_desc(a0 + k * 32, 0, 1024),
                          make_smem_desc(b0 + k * 32, 0, 1024), idesc, (it | s | k) ? 1u : 0u);
            }
          }
          umma_commit(bar_empty + 8 * stage);
          if (++stage == PC_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(bar_tfull + 8 * buf);
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue (warps 2..5)
    // A "unit" is one 64-channel half of one 128-pixel tile: 16 KB of bf16 staged in out_stage[unit % 3] and
    // written by one TMA store.  Units run in the fixed order item -> tile -> half; the leader thread keeps the
    // residual half-tiles two units ahead of the arithmetic.
    const int q = warp & 3;
    const int m = q * 32 + lane;                 // pixel row of the tile == TMEM lane
    const bool leader = threadIdx.x == 64;
    const uint32_t swz = static_cast<uint32_t>(m & 7);
    const int my_items = (num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                         static_cast<int>(gridDim.x);
    const int total_units = my_items * (PC_MT * 2);
    auto unit_coords = [&](int u, int& c0, int& ow0, int& oh, int& n) {
      const int item = static_cast<int>(blockIdx.x) + (u / (PC_MT * 2)) * static_cast<int>(gridDim.x);
      tile_coords(item * PC_MT + ((u >> 1) & (PC_MT - 1)), ow0, oh, n);
      c0 = (u & 1) * 64;
    };
    auto load_residual = [&](int u) {            // leader only
      if (!p.has_residual || u >= total_units) return;
      int c0, ow0, oh, n;
      unit_coords(u, c0, ow0, oh, n);
      if (n >= p.NB) return;
      const uint32_t rb = bar_res + 8 * (u % PC_OUT_BUFS);
      mbar_arrive_expect_tx(rb, PC_OUT_BYTES);
      tma_load_5d(sOut + (u % PC_OUT_BUFS) * PC_OUT_BYTES, &tmR, rb, c0, ow0, 0, oh, n);
    };
    if (leader) { load_residual(0); load_residual(1); }
    uint32_t it_item = 0;
    int u = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it_item) {
      const uint32_t buf = it_item & 1;
      mbar_wait(bar_tfull + 8 * buf, (it_item >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < PC_MT; ++j) {
        int ow0, oh, n;
        tile_coords(item * PC_MT + j, ow0, oh, n);
        const bool valid = n < p.NB;             // W % 128 == 0: a tile is a full row segment or past the end
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * (PC_MT * PC_BN) + j * PC_BN;
#pragma unroll 1
        for (int half = 0; half < 2; ++half, ++u) {
          const uint32_t ob = u % PC_OUT_BUFS;
          uint8_t* row = out_stage + ob * PC_OUT_BYTES + m * 128;
          if (p.has_residual && valid) mbar_wait(bar_res + 8 * ob, (u / PC_OUT_BUFS) & 1);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c0 = half * 64 + cc * 32;
            uint32_t r[32];
            tmem_ld_32x32(trow + c0, r);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + bias_s[c0 + i];
            if (p.has_residual && valid) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 uu = *reinterpret_cast<const uint4*>(row + (((cc * 4 + i) ^ swz) << 4));
                v[8 * i + 0] += bf16_lo(uu.x); v[8 * i + 1] += bf16_hi(uu.x);
                v[8 * i + 2] += bf16_lo(uu.y); v[8 * i + 3] += bf16_hi(uu.y);
                v[8 * i + 4] += bf16_lo(uu.z); v[8 * i + 5] += bf16_hi(uu.z);
                v[8 * i + 6] += bf16_lo(uu.w); v[8 * i + 7] += bf16_hi(uu.w);
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 uu;
              uu.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
              uu.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
              uu.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
              uu.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
              *reinterpret_cast<uint4*>(row + (((cc * 4 + i) ^ swz) << 4)) = uu;
            }
            if (p.gn_part) {
              // per-group (4 channels) sum / sum of squares of this row, then a fixed-order transposed
              // butterfly over the warp's 32 rows: 16 values -> lane pair (2i, 2i+1) holds total i
              float a8[8], a4[4], a2[2], a1;
              {
                float vals[16];
#pragma unroll
                for (int g8 = 0; g8 < 8; ++g8) {
                  const float x0 = valid ? v[4 * g8] : 0.f, x1 = valid ? v[4 * g8 + 1] : 0.f;
                  const float x2 = valid ? v[4 * g8 + 2] : 0.f, x3 = valid ? v[4 * g8 + 3] : 0.f;
                  vals[g8] = (x0 + x1) + (x2 + x3);
                  vals[8 + g8] = fmaf(x0, x0, x1 * x1) + fmaf(x2, x2, x3 * x3);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float send = (lane & 16) ? vals[i] : vals[i + 8];
                  const float keep = (lane & 16) ? vals[i + 8] : vals[i];
                  a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float send = (lane & 8) ? a8[i] : a8[i + 4];
                const float keep = (lane & 8) ? a8[i + 4] : a8[i];
                a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
              }
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const float send = (lane & 4) ? a4[i] : a4[i + 2];
                const float keep = (lane & 4) ? a4[i + 2] : a4[i];
                a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
              }
              {
                const float send = (lane & 2) ? a2[0] : a2[1];
                const float keep = (lane & 2) ? a2[1] : a2[0];
                a1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
              }
              a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
              if ((lane & 1) == 0) {
                const int idx = lane >> 1;                     // 0..7: group sums, 8..15: group sums of squares
                const int group = (c0 >> 2) + (idx & 7);
                gn_red[q * 64 + group * 2 + (idx >> 3)] = a1;
              }
            }
          }
          // staged half-tile -> global: make the generic-proxy writes visible to the async proxy, then one TMA store
          fence_proxy_async_smem();
          asm volatile("bar.sync 2, 128;" ::: "memory");
          if (leader) {
            if (valid) {
              asm volatile(
                  "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];" ::"l"(&tmO),
                  "r"(half * 64), "r"(ow0), "r"(0), "r"(oh), "r"(n), "r"(sOut + ob * PC_OUT_BYTES)
                  : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // the buffer of unit u+2 was last read by the store of unit u-1: wait for it, then refill / release it
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            load_residual(u + 2);
          }
          if (half == 1 && p.gn_part) {
            const int te = threadIdx.x - 64;
            const int tile = item * PC_MT + j;
            if (te < 64 && tile < p.num_tiles)
              p.gn_part[static_cast<long long>(tile) * 64 + te] =
                  (gn_red[te] + gn_red[64 + te]) + (gn_red[128 + te] + gn_red[192 + te]);
            asm volatile("bar.sync 2, 128;" ::: "memory");     // gn_red is rewritten by the next tile
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
    }
    if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace b2

using namespace b2;

extern "C" {

// 3x3 stride-1 pad-1 convolution (or its data gradient) of an NHWC bf16 tensor with Cout == 128,
// W % 128 == 0, Cin % 64 == 0.  b_ptr: [128, 9*Cin] bf16, column = (r*3+s)*Cin + ci.
// dgrad != 0: tap (r,s) reads the pixel at (+1-r, +1-s) instead of (r-1, s-1).
int b2dq_pconv3x3(const void* a_ptr, const void* b_ptr, void* out, const float* bias,
                  const void* residual, float* gn_part, int NB, int H, int W, int Cin, int dgrad, int max_ctas,
                  cudaStream_t stream) {
  if (NB <= 0 || H <= 0 || W <= 0) return 0;
  if (W % 128 != 0 || Cin % 64 != 0 || Cin <= 0) return -1;
  CUtensorMap tmA, tmB, tmO, tmR;
  {
    uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, 1, (uint64_t)H, (uint64_t)NB};
    uint64_t str[5] = {1, (uint64_t)Cin, (uint64_t)W * Cin, (uint64_t)W * Cin, (uint64_t)H * W * Cin};
    uint32_t box[5] = {64, PC_STRIP_ROWS, 1, 1, 1};
    int r = make_tmap_bf16(&tmA, a_ptr, 5, dims, str, box);
    if (r) return r;
  }
  {
    uint64_t dims[2] = {(uint64_t)9 * Cin, (uint64_t)PC_BN};
    uint64_t str[2] = {1, (uint64_t)9 * Cin};
    uint32_t box[2] = {64, PC_BN};
    int r = make_tmap_bf16(&tmB, b_ptr, 2, dims, str, box);
    if (r) return r - 1000;
  }
  {
    // output (and residual) [NB,H,W,128]: half-tiles of 128 pixels x 64 channels, 128 B-swizzled in shared memory
    uint64_t dims[5] = {(uint64_t)PC_BN, (uint64_t)W, 1, (uint64_t)H, (uint64_t)NB};
    uint64_t str[5] = {1, (uint64_t)PC_BN, (uint64_t)W * PC_BN, (uint64_t)W * PC_BN, (uint64_t)H * W * PC_BN};
    uint32_t box[5] = {64, 128, 1, 1, 1};
    int r = make_tmap_bf16(&tmO, out, 5, dims, str, box);
    if (r) return r - 2000;
    r = make_tmap_bf16(&tmR, residual ? residual : out, 5, dims, str, box);
    if (r) return r - 3000;
  }
  static unsigned long long attr_mask = 0;
  if (int e = set_max_smem_once(pconv3x3_kernel<3>, PC_SMEM, attr_mask)) return e;
  PconvParams p;
  memset(&p, 0, sizeof(p));
  p.kchunks = Cin / 64; p.nr = 3; p.strip_dw = -1;
  for (int i = 0; i < 3; ++i) {
    p.row_dh[i] = dgrad ? 1 - i : i - 1;
    p.col_off[i] = dgrad ? 2 - i : i;
    for (int s = 0; s < 3; ++s) p.wcol[i * 3 + s] = (i * 3 + s) * Cin;
  }
  p.tiles_w = W / 128; p.H = H; p.NB = NB; p.W = W;
  p.num_tiles = NB * H * p.tiles_w;
  p.Cout = PC_BN;
  p.bias = bias;
  p.has_residual = residual != nullptr;
  p.gn_part = gn_part;
  int grid = device_sm_count();
  if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
  const int items = (p.num_tiles + PC_MT - 1) / PC_MT;
  if (items < grid) grid = items;
  pconv3x3_kernel<3><<<grid, 192, PC_SMEM, stream>>>(tmA, tmB, tmO, tmR, p);
  return (int)cudaGetLastError();
}

// The same persistent kernel for any group of taps that can be read from row strips: `nr` row taps (A-map channel
// base / parity plane / row offset each) x `ns` column taps (pixel offsets col_dw[s], at most 2 apart) - the 2x2
// parity classes of the folded up-convolution and of the stride-2 data gradient, whose four-tap K loop is too short
// for the one-shot tap GEMM (36 % tensor pipe: prologue and epilogue are not overlapped there).
struct b2dq_pconv_taps_desc {
  const void* a_ptr;
  long long a_dims[5];
  long long a_strides[5];     // (c, w, p, h, n) view of the input, elements
  const void* b_ptr;          // [128][b_k] bf16
  long long b_k;
  int kchunks;                // channels per tap / 64
  int nr, ns;                 // row taps (<= 9), column taps (1..3)
  int row_c[9], row_p[9], row_dh[9];
  int col_dw[3];              // pixel offset of column tap s (max - min <= 2)
  int wcol[27];               // weight column base of tap (r, s) at [r*3 + s]
  int NB, H, W;               // output tile grid: NB images x H rows x W pixels, W % 128 == 0
  void* out;                  // bf16, pixel (n, h, w) at out + n*oN + h*oH + w*oW, 128 channels
  long long oN, oH, oW;
  const float* bias;
};

int b2dq_pconv_taps(const b2dq_pconv_taps_desc* d, int max_ctas, cudaStream_t stream) {
  if (!d || d->nr < 1 || d->nr > PC_MAX_ROWS || d->ns < 1 || d->ns > 3 || d->kchunks < 1) return -1;
  if (d->NB <= 0 || d->H <= 0 || d->W <= 0) return 0;
  if (d->W % 128 != 0) return -1;
  int dw_min = d->col_dw[0], dw_max = d->col_dw[0];
  for (int s = 1; s < d->ns; ++s) {
    dw_min = d->col_dw[s] < dw_min ? d->col_dw[s] : dw_min;
    dw_max = d->col_dw[s] > dw_max ? d->col_dw[s] : dw_max;
  }
  if (dw_max - dw_min > 2) return -1;
  if (d->oW % 8 || d->oH % 8 || d->oN % 8 || (reinterpret_cast<uintptr_t>(d->out) & 15)) return -1;
  CUtensorMap tmA, tmB, tmO;
  {
    uint64_t dims[5], str[5];
    for (int i = 0; i < 5; ++i) { dims[i] = (uint64_t)d->a_dims[i]; str[i] = (uint64_t)d->a_strides[i]; }
    uint32_t box[5] = {64, PC_STRIP_ROWS, 1, 1, 1};
    int r = make_tmap_bf16(&tmA, d->a_ptr, 5, dims, str, box);
    if (r) return r;
  }
  {
    uint64_t dims[2] = {(uint64_t)d->b_k, (uint64_t)PC_BN};
    uint64_t str[2] = {1, (uint64_t)d->b_k};
    uint32_t box[2] = {64, PC_BN};
    int r = make_tmap_bf16(&tmB, d->b_ptr, 2, dims, str, box);
    if (r) return r - 1000;
  }
  {
    uint64_t dims[5] = {(uint64_t)PC_BN, (uint64_t)d->W, 1, (uint64_t)d->H, (uint64_t)d->NB};
    uint64_t str[5] = {1, (uint64_t)d->oW, (uint64_t)d->oH, (uint64_t)d->oH, (uint64_t)d->oN};
    uint32_t box[5] = {64, 128, 1, 1, 1};
    int r = make_tmap_bf16(&tmO, d->out, 5, dims, str, box);
    if (r) return r - 2000;
  }
  PconvParams p;
  memset(&p, 0, sizeof(p));
  p.kchunks = d->kchunks; p.nr = d->nr; p.strip_dw = dw_min;
  for (int r = 0; r < d->nr; ++r) {
    p.row_c[r] = d->row_c[r]; p.row_p[r] = d->row_p[r]; p.row_dh[r] = d->row_dh[r];
    for (int s = 0; s < d->ns; ++s) p.wcol[r * 3 + s] = d->wcol[r * 3 + s];
  }
  for (int s = 0; s < d->ns; ++s) p.col_off[s] = d->col_dw[s] - dw_min;
  p.tiles_w = d->W / 128; p.H = d->H; p.NB = d->NB; p.W = d->W;
  p.num_tiles = d->NB * d->H * p.tiles_w;
  p.Cout = PC_BN;
  p.bias = d->bias;
  p.has_residual = 0;
  p.gn_part = nullptr;
  int grid = device_sm_count();
  if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
  const int items = (p.num_tiles + PC_MT - 1) / PC_MT;
  if (items < grid) grid = items;
  static unsigned long long m1 = 0, m2 = 0, m3 = 0;
  if (d->ns == 1) {
    if (int e = set_max_smem_once(pconv3x3_kernel<1>, PC_SMEM, m1)) return e;
    pconv3x3_kernel<1><<<grid, 192, PC_SMEM, stream>>>(tmA, tmB, tmO, tmO, p);
  } else if (d->ns == 2) {
    if (int e = set_max_smem_once(pconv3x3_kernel<2>, PC_SMEM, m2)) return e;
    pconv3x3_kernel<2><<<grid, 192, PC_SMEM, stream>>>(tmA, tmB, tmO, tmO, p);
  } else {
    if (int e = set_max_smem_once(pconv3x3_kernel<3>, PC_SMEM, m3)) return e;
    pconv3x3_kernel<3><<<grid, 192, PC_SMEM, stream>>>(tmA, tmB, tmO, tmO, p);
  }
  return (int)cudaGetLastError();
}

// stats[n][g] = (mean, rstd) from the per-tile partials written by b2dq_pconv3x3 (tiles of an image are
// contiguous; added in tile order: deterministic).  count = H*W*4 elements per group.
__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(0xffffffffu, lo, m);
  hi = __shfl_xor_sync(0xffffffffu, hi, m);
  return __hiloint2double(hi, lo);
}
// One warp per (image, group): lane l adds tiles l, l+32, ... in order, then a fixed butterfly.
__global__ void gn_finalize_tiles_kernel(const float* __restrict__ part, float* stats, int N,
                                         int tiles_per_image, double inv_count, float eps) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= N * 32) return;
  const int n = i / 32, g = i % 32;
  double a = 0.0, b = 0.0;
  const float* pp = part + (static_cast<long long>(n) * tiles_per_image) * 64 + g * 2;
  for (int t = lane; t < tiles_per_image; t += 32) {
    const float2 v = *reinterpret_cast<const float2*>(pp + static_cast<long long>(t) * 64);
    a += v.x;
    b += v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += shfl_xor_f64(a, o);
    b += shfl_xor_f64(b, o);
  }
  if (lane == 0) {
    const double mean = a * inv_count;
    double var = b * inv_count - mean * mean;
    if (var < 0) var = 0;
    stats[2 * i] = static_cast<float>(mean);
    stats[2 * i + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
}

int b2dq_gn_finalize_tiles(const float* gn_part, float* stats, int N, int H, int W, float eps,
                           cudaStream_t stream) {
  if (N <= 0) return 0;
  if (W % 128) return -1;
  const int tpi = H * (W / 128);
  gn_finalize_tiles_kernel<<<(N * 32 * 32 + 255) / 256, 256, 0, stream>>>(
      gn_part, stats, N, tpi, 1.0 / (static_cast<double>(H) * W * 4), eps);
  return (int)cudaGetLastError();
}

}  // extern "C"
