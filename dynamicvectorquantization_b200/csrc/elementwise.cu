// Memory-bound helpers of the DQ-VAE hot path (all NHWC bf16 unless noted): layout changes at the
// NCHW-fp32 module boundary, nearest-neighbour 2x up-sampling (model.py:49-50) and its gradient,
// row softmax of the attention logits (model.py:182) and its gradient, bias gradients, the small
// im2col used for the 3-channel edge convolutions, and their weight gradient.
#include "common.cuh"

namespace b2 {

// [N][C][HW] (src) -> [N][HW][C] (dst) tiled transpose, 32x32 tiles; TI/TO = element types.
template <typename TI, typename TO>
__global__ void transpose_chw_hwc_kernel(const TI* __restrict__ src, TO* __restrict__ dst, int C,
                                         int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const TI* s = src + static_cast<long long>(n) * C * HW;
  TO* d = dst + static_cast<long long>(n) * C * HW;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, hw = hw0 + threadIdx.x;
    if (c < C && hw < HW) tile[j][threadIdx.x] = static_cast<float>(s[static_cast<long long>(c) * HW + hw]);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int hw = hw0 + j, c = c0 + threadIdx.x;
    if (c < C && hw < HW) d[static_cast<long long>(hw) * C + c] = static_cast<TO>(tile[threadIdx.x][j]);
  }
}
// [N][HW][C] -> [N][C][HW]
template <typename TI, typename TO>
__global__ void transpose_hwc_chw_kernel(const TI* __restrict__ src, TO* __restrict__ dst, int C,
                                         int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const TI* s = src + static_cast<long long>(n) * C * HW;
  TO* d = dst + static_cast<long long>(n) * C * HW;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int hw = hw0 + j, c = c0 + threadIdx.x;
    if (c < C && hw < HW) tile[j][threadIdx.x] = static_cast<float>(s[static_cast<long long>(hw) * C + c]);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, hw = hw0 + threadIdx.x;
    if (c < C && hw < HW) d[static_cast<long long>(c) * HW + hw] = static_cast<TO>(tile[threadIdx.x][j]);
  }
}

// Few-channel images (C <= 4: the RGB input and reconstruction): one thread per pixel; every channel plane is
// read / written with unit stride across the warp and the interleaved side is a contiguous run of 32*C elements
// per warp (the 32x32 tile kernel above leaves 29 of 32 lanes idle at C = 3: 134 us for a 25 MB image batch).
template <typename TI, typename TO, bool TO_HWC>
__global__ void transpose_fewc_kernel(const TI* __restrict__ src, TO* __restrict__ dst, int C, int HW, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // n*HW + hw
  if (i >= total) return;
  const long long n = i / HW;
  const int hw = static_cast<int>(i - n * HW);
  const long long plane0 = n * C * HW + hw, inter0 = i * C;
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    if (TO_HWC) dst[inter0 + c] = static_cast<TO>(static_cast<float>(src[plane0 + static_cast<long long>(c) * HW]));
    else dst[plane0 + static_cast<long long>(c) * HW] = static_cast<TO>(static_cast<float>(src[inter0 + c]));
  }
}

// nearest 2x: out[n, 2h+a, 2w+b, :] = in[n, h, w, :]
__global__ void upsample2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int H,
                                  int W, int vecs, long long total_out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total_out) return;
  const int v = static_cast<int>(i % vecs);
  long long r = i / vecs;
  const int ow = static_cast<int>(r % (2 * W)); r /= (2 * W);
  const int oh = static_cast<int>(r % (2 * H));
  const long long n = r / (2 * H);
  out[i] = __ldg(in + ((n * H + (oh >> 1)) * W + (ow >> 1)) * vecs + v);
}
// gradient: in-grad[n,h,w,:] = sum of the 4 out-grads
__global__ void upsample2x_bwd_kernel(const uint4* __restrict__ gout, uint4* __restrict__ gin,
                                      int H, int W, int vecs, long long total_in) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total_in) return;
  const int v = static_cast<int>(i % vecs);
  long long r = i / vecs;
  const int w = static_cast<int>(r % W); r /= W;
  const int h = static_cast<int>(r % H);
  const long long n = r / H;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const uint4 u = __ldg(gout + ((n * 2 * H + 2 * h + a) * (2 * W) + 2 * w + b) * vecs + v);
      acc[0] += bf16_lo(u.x); acc[1] += bf16_hi(u.x); acc[2] += bf16_lo(u.y); acc[3] += bf16_hi(u.y);
      acc[4] += bf16_lo(u.z); acc[5] += bf16_hi(u.z); acc[6] += bf16_lo(u.w); acc[7] += bf16_hi(u.w);
    }
  uint4 o;
  o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
  o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
  gin[i] = o;
}

// One warp per row; T even.  p = softmax(s) row-wise; logits fp32 (TS=float) or bf16.
template <typename TS>
__global__ void softmax_rows_kernel(const TS* __restrict__ s, __nv_bfloat16* __restrict__ p,
                                    long long rows, int T) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const TS* sp = s + row * T;
  uint32_t* pp = reinterpret_cast<uint32_t*>(p + row * T);
  float mx = -INFINITY;
  for (int i = lane; i < T / 2; i += 32)
    mx = fmaxf(mx, fmaxf(static_cast<float>(sp[2 * i]), static_cast<float>(sp[2 * i + 1])));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int i = lane; i < T / 2; i += 32)
    sum += __expf(static_cast<float>(sp[2 * i]) - mx) + __expf(static_cast<float>(sp[2 * i + 1]) - mx);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int i = lane; i < T / 2; i += 32)
    pp[i] = pack_bf16x2(__expf(static_cast<float>(sp[2 * i]) - mx) * inv,
                        __expf(static_cast<float>(sp[2 * i + 1]) - mx) * inv);
}
// ds = scale * p * (dp - sum_j p_j dp_j)
__global__ void softmax_bwd_rows_kernel(const __nv_bfloat16* __restrict__ p,
                                        const __nv_bfloat16* __restrict__ dp,
                                        __nv_bfloat16* __restrict__ ds, long long rows, int T,
                                        float scale) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const uint32_t* pp = reinterpret_cast<const uint32_t*>(p + row * T);
  const uint32_t* dpp = reinterpret_cast<const uint32_t*>(dp + row * T);
  uint32_t* dsp = reinterpret_cast<uint32_t*>(ds + row * T);
  float dot = 0.f;
  for (int i = lane; i < T / 2; i += 32) {
    const uint32_t a = pp[i], b = dpp[i];
    dot += bf16_lo(a) * bf16_lo(b) + bf16_hi(a) * bf16_hi(b);
  }
  dot = warp_sum(dot);
  for (int i = lane; i < T / 2; i += 32) {
    const uint32_t a = pp[i], b = dpp[i];
    dsp[i] = pack_bf16x2(scale * bf16_lo(a) * (bf16_lo(b) - dot), scale * bf16_hi(a) * (bf16_hi(b) - dot));
  }
}

__global__ void add_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                uint4* __restrict__ o, long long nvec) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  const uint4 x = __ldg(a + i), y = __ldg(b + i);
  uint4 r;
  r.x = pack_bf16x2(bf16_lo(x.x) + bf16_lo(y.x), bf16_hi(x.x) + bf16_hi(y.x));
  r.y = pack_bf16x2(bf16_lo(x.y) + bf16_lo(y.y), bf16_hi(x.y) + bf16_hi(y.y));
  r.z = pack_bf16x2(bf16_lo(x.z) + bf16_lo(y.z), bf16_hi(x.z) + bf16_hi(y.z));
  r.w = pack_bf16x2(bf16_lo(x.w) + bf16_lo(y.w), bf16_hi(x.w) + bf16_hi(y.w));
  o[i] = r;
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ a, __nv_bfloat16* __restrict__ o,
                                     long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) o[i] = __float2bfloat16_rn(a[i]);
}

// im2col for a 3x3 pad-1 stride-1 window over a few-channel NHWC bf16 image:
// dst[pixel][t*Cs + c] = src[pixel + off_t][c], zero padded to 64 columns.
// flip = 0: off_t = (r-1, s-1)   (forward / weight gradient);  flip = 1: off_t = (1-r, 1-s) (dgrad).
// One thread produces 8 columns (one 16 B store).
// CS > 0: channel count known at compile time (RGB: 3): the column -> (tap, channel) divisions become
// multiply-shifts (the runtime-Cs version spends its time in integer division sequences).
template <int CS>
__global__ void im2col3x3_small_kernel(const __nv_bfloat16* __restrict__ src,
                                       __nv_bfloat16* __restrict__ dst, int H, int W, int Cs_rt,
                                       int flip, long long total_vecs) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total_vecs) return;
  const int Cs = CS > 0 ? CS : Cs_rt;
  const int cv = static_cast<int>(i & 7);
  const long long pix = i >> 3;
  const int w = static_cast<int>(pix % W);
  const int h = static_cast<int>((pix / W) % H);
  const long long n = pix / (static_cast<long long>(W) * H);
  const __nv_bfloat16* img = src + n * H * W * Cs;
  const int sgn = flip ? -1 : 1;
  __align__(16) __nv_bfloat16 vals[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int col = cv * 8 + k;
    __nv_bfloat16 v = __float2bfloat16_rn(0.f);
    if (col < 9 * Cs) {
      const int t = col / Cs, c = col - t * Cs;     // divisions by a literal when CS > 0
      const int r = t / 3, sx = t - 3 * r;
      const int hh = h + sgn * (r - 1), ww = w + sgn * (sx - 1);
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = img[(hh * W + ww) * Cs + c];
    }
    vals[k] = v;
  }
  reinterpret_cast<uint4*>(dst)[i] = *reinterpret_cast<const uint4*>(vals);
}

// Three-channel case (RGB image in, RGB gradient out): ONE thread builds the whole 64-column row of a
// pixel - 27 two-byte loads that hit L1 (neighbouring pixels share them), eight 16 B stores - instead of
// eight threads each resolving (tap, channel) per element.
__global__ void __launch_bounds__(256)
im2col3x3_c3_kernel(const unsigned short* __restrict__ src, uint4* __restrict__ dst, int H,
                    int W, int flip, long long npix) {
  // the 128 B rows of the block's 256 pixels are staged in shared memory (16 B units XOR-swizzled by the row so
  // that both sides are conflict-free) and written out with unit stride: a thread storing its own row would touch
  // 32 different lines per warp instruction
  __shared__ uint4 tile[256 * 8];
  const long long pix0 = static_cast<long long>(blockIdx.x) * 256;
  const long long pix = pix0 + threadIdx.x;
  const uint32_t tsw = threadIdx.x & 7;
  if (pix < npix) {
    const int w = static_cast<int>(pix % W);
    const int h = static_cast<int>((pix / W) % H);
    const long long n = pix / (static_cast<long long>(W) * H);
    const unsigned short* img = src + n * H * W * 3;
    const int sgn = flip ? -1 : 1;
    unsigned short v[32];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int r = t / 3, sx = t - 3 * r;
      const int hh = h + sgn * (r - 1), ww = w + sgn * (sx - 1);
      const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
      const unsigned short* q = img + (hh * W + ww) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[t * 3 + c] = ok ? __ldg(q + c) : static_cast<unsigned short>(0);
    }
#pragma unroll
    for (int k = 27; k < 32; ++k) v[k] = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 o;
      o.x = v[8 * j + 0] | (static_cast<uint32_t>(v[8 * j + 1]) << 16);
      o.y = v[8 * j + 2] | (static_cast<uint32_t>(v[8 * j + 3]) << 16);
      o.z = v[8 * j + 4] | (static_cast<uint32_t>(v[8 * j + 5]) << 16);
      o.w = v[8 * j + 6] | (static_cast<uint32_t>(v[8 * j + 7]) << 16);
      tile[threadIdx.x * 8 + (j ^ tsw)] = o;
    }
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int j = 4; j < 8; ++j) tile[threadIdx.x * 8 + (j ^ tsw)] = z;
  }
  __syncthreads();
  const long long units = (npix - pix0 < 256 ? npix - pix0 : 256) * 8;
  uint4* out = dst + pix0 * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = i * 256 + threadIdx.x;
    if (k < units) {
      const int row = k >> 3, unit = k & 7;
      out[k] = tile[row * 8 + (unit ^ (row & 7))];
    }
  }
}

// part[block][c] = sum over the block's rows of dy[row][c] (bias gradient), dy bf16 [rows][C],
// C % 8 == 0.  Thread = (8-channel vector, row lane); 16 B loads, 4 rows in flight; the per-thread
// sums are combined in a fixed order (deterministic), bias_grad_reduce adds the blocks in order.
__global__ void bias_grad_kernel(const __nv_bfloat16* __restrict__ dy, float* part, long long rows,
                                 int C, int rows_per_block) {
  extern __shared__ float sh[];  // [rstep][C]
  const int vecs = C >> 3;
  const int v = threadIdx.x % vecs;
  const int rlane = threadIdx.x / vecs;
  const int rstep = blockDim.x / vecs;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const __nv_bfloat16* base = dy + v * 8;
  long long r = r0 + rlane;
  for (; r + 3 * rstep < r1; r += 4 * rstep) {
    uint4 u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = __ldg(reinterpret_cast<const uint4*>(base + (r + j * rstep) * C));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s[0] += bf16_lo(u[j].x); s[1] += bf16_hi(u[j].x); s[2] += bf16_lo(u[j].y); s[3] += bf16_hi(u[j].y);
      s[4] += bf16_lo(u[j].z); s[5] += bf16_hi(u[j].z); s[6] += bf16_lo(u[j].w); s[7] += bf16_hi(u[j].w);
    }
  }
  for (; r < r1; r += rstep) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + r * C));
    s[0] += bf16_lo(u.x); s[1] += bf16_hi(u.x); s[2] += bf16_lo(u.y); s[3] += bf16_hi(u.y);
    s[4] += bf16_lo(u.z); s[5] += bf16_hi(u.z); s[6] += bf16_lo(u.w); s[7] += bf16_hi(u.w);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) sh[rlane * C + v * 8 + k] = s[k];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
    for (int rl = 0; rl < rstep; ++rl) a += sh[rl * C + c];
    part[static_cast<long long>(blockIdx.x) * C + c] = a;
  }
}
// generic variant for channel counts that are not a multiple of 8
__global__ void bias_grad_generic_kernel(const __nv_bfloat16* __restrict__ dy, float* part,
                                         long long rows, int C, int rows_per_block) {
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (long long r = r0; r < r1; ++r) s += __bfloat162float(dy[r * C + c]);
    part[static_cast<long long>(blockIdx.x) * C + c] = s;
  }
}
// out[c] = sum_b part[b][c].  256 threads = 32 channels x 8 block lanes: lane l adds blocks l, l+8, ... in
// order, the eight lane sums are then added in lane order - a fixed order for a given `blocks`, so the
// result is reproducible (the old one-thread-per-channel loop was ~1200 dependent loads long).
__global__ void bias_grad_reduce_kernel(const float* __restrict__ part, float* out, int blocks, int C) {
  __shared__ float sh[8][32];
  const int cl = threadIdx.x & 31, seg = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float a0 = 0.f, a1 = 0.f;
  if (c < C) {
    int b = seg;
    for (; b + 8 < blocks; b += 16) {
      a0 += part[static_cast<long long>(b) * C + c];
      a1 += part[static_cast<long long>(b + 8) * C + c];
    }
    if (b < blocks) a0 += part[static_cast<long long>(b) * C + c];
  }
  sh[seg][cl] = a0 + a1;
  __syncthreads();
  if (seg == 0 && c < C) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) a += sh[k][cl];
    out[c] = a;
  }
}

// OIHW fp32 -> the two bf16 GEMM packings of a convolution weight in one pass:
//   fwd  [Cout][R*S*Cin]  (tap-major, Cin contiguous)     dgr  [Cin][R*S*Cout]  (contraction over Cout)
// Either output may be null.  One thread per weight element, read in memory order.
__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* fwd, __nv_bfloat16* dgr,
                                    int Cout, int Cin, int RS, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int t = static_cast<int>(i % RS);
  const int ci = static_cast<int>((i / RS) % Cin);
  const int co = static_cast<int>(i / (static_cast<long long>(RS) * Cin));
  const __nv_bfloat16 v = __float2bfloat16_rn(w[i]);
  if (fwd) fwd[(static_cast<long long>(co) * RS + t) * Cin + ci] = v;
  if (dgr) dgr[(static_cast<long long>(ci) * RS + t) * Cout + co] = v;
}

// The same for MANY weights in one launch (every convolution weight is repacked after each optimizer step: one
// launch per layer costs more in launch latency than in bytes).  A block owns a tile of 32 output x 32 input channels
// (all R*S taps) of one weight: it is read in memory order into shared memory and written out so that both packings
// get 64 B runs (32 consecutive Cin of a tap for fwd, 32 consecutive Cout for dgrad).
struct PackItem {
  const float* w;
  __nv_bfloat16* fwd;
  __nv_bfloat16* dgr;
  int cout, cin, rs, tiles_ci;
  long long tile_start;          // first block of this weight
};
constexpr int PK_T = 32;
constexpr int PK_MAX_RS = 16;
__global__ void __launch_bounds__(256)
pack_weights_multi_kernel(const PackItem* __restrict__ items, int n_items) {
  extern __shared__ float pk_tile[];             // [32 co][32*RS + 1]
  int lo = 0, hi = n_items - 1;                  // last item with tile_start <= blockIdx.x
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (items[mid].tile_start <= static_cast<long long>(blockIdx.x)) lo = mid; else hi = mid - 1;
  }
  const PackItem it = items[lo];
  const int local = static_cast<int>(blockIdx.x - it.tile_start);
  const int co0 = (local / it.tiles_ci) * PK_T, ci0 = (local % it.tiles_ci) * PK_T;
  const int RS = it.rs, row = PK_T * RS, ld = row + 1;
  const int nco = min(PK_T, it.cout - co0), nci = min(PK_T, it.cin - ci0);
  // load: for each co of the tile, the nci*RS floats starting at w[(co*Cin + ci0)*RS] are contiguous
  for (int co = threadIdx.x >> 5; co < nco; co += 8) {
    const float* src = it.w + (static_cast<long long>(co0 + co) * it.cin + ci0) * RS;
    for (int j = threadIdx.x & 31; j < nci * RS; j += 32) pk_tile[co * ld + j] = src[j];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  if (it.fwd) {                                  // fwd[(co*RS + t)*Cin + ci]: lane = ci
    for (int q = wrp; q < nco * RS; q += 8) {
      const int co = q / RS, t = q - co * RS;
      if (lane < nci)
        it.fwd[(static_cast<long long>(co0 + co) * RS + t) * it.cin + ci0 + lane] =
            __float2bfloat16_rn(pk_tile[co * ld + lane * RS + t]);
    }
  }
  if (it.dgr) {                                  // dgr[(ci*RS + t)*Cout + co]: lane = co
    for (int q = wrp; q < nci * RS; q += 8) {
      const int ci = q / RS, t = q - ci * RS;
      if (lane < nco)
        it.dgr[(static_cast<long long>(ci0 + ci) * RS + t) * it.cout + co0 + lane] =
            __float2bfloat16_rn(pk_tile[lane * ld + ci * RS + t]);
    }
  }
}

// ---------------------------------------------------------------- nearest x2 + 3x3 convolution, folded
// Upsample.forward (modules/diffusionmodules/model.py:49-53) as four 2x2 convolutions of the low-resolution input,
// one per output parity class (ph, pw).  Filter row r lands on low-resolution row slot a of parity ph when
// r is in ROWS(ph, a): (0,0) -> {0}, (0,1) -> {1,2}, (1,0) -> {0,1}, (1,1) -> {2}; the same table serves the columns.
__device__ __forceinline__ void up_range(int parity, int slot, int& lo, int& hi) {
  lo = (parity == 0) ? (slot == 0 ? 0 : 1) : (slot == 0 ? 0 : 2);
  hi = (parity == 0) ? (slot == 0 ? 0 : 2) : (slot == 0 ? 1 : 2);
}
// w fp32 [Cout][Cin][3][3] -> fwd bf16 [Cout][16*Cin] and dgr bf16 [Cin][16*Cout], column block ((ph*2+pw)*2+a)*2+b:
// folded weight = sum of the 3x3 taps that land on slot (a, b) of class (ph, pw), summed in (r, s) order in fp32.
__global__ void upconv_pack_kernel(const float* __restrict__ w, __nv_bfloat16* fwd, __nv_bfloat16* dgr, int Cout,
                                   int Cin) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(Cout) * Cin) return;
  const int ci = static_cast<int>(i % Cin), co = static_cast<int>(i / Cin);
  float k[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) k[t] = w[i * 9 + t];
#pragma unroll
  for (int blk = 0; blk < 16; ++blk) {
    const int ph = blk >> 3, pw = (blk >> 2) & 1, a = (blk >> 1) & 1, b = blk & 1;
    int r0, r1, s0, s1;
    up_range(ph, a, r0, r1);
    up_range(pw, b, s0, s1);
    float v = 0.f;
    for (int r = r0; r <= r1; ++r)
      for (int s_ = s0; s_ <= s1; ++s_) v += k[r * 3 + s_];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    if (fwd) fwd[(static_cast<long long>(co) * 16 + blk) * Cin + ci] = h;
    if (dgr) dgr[(static_cast<long long>(ci) * 16 + blk) * Cout + co] = h;
  }
}
// partial fp32 [4 classes][splits][4 slots (a*2+b)][Cout][Cin] (split-K partial weight gradients of the four class
// GEMMs) -> dw fp32 [Cout][Cin][3][3]: splits added in index order, then every class/slot sum is scattered onto the
// 3x3 taps it was folded from (the transpose of the fold above).
__global__ void upconv_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int splits,
                                           int Cout, int Cin) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long per = static_cast<long long>(Cout) * Cin;
  if (i >= per) return;
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
#pragma unroll
  for (int blk = 0; blk < 16; ++blk) {
    const int cls = blk >> 2, slot = blk & 3;
    const float* src = partial + (static_cast<long long>(cls) * splits * 4 + slot) * per + i;
    float v = 0.f;
    for (int sp = 0; sp < splits; ++sp) v += __ldcs(src + static_cast<long long>(sp) * 4 * per);
    const int ph = cls >> 1, pw = cls & 1, a = slot >> 1, b = slot & 1;
    int r0, r1, s0, s1;
    up_range(ph, a, r0, r1);
    up_range(pw, b, s0, s1);
    for (int r = r0; r <= r1; ++r)
      for (int s_ = s0; s_ <= s1; ++s_) acc[r * 3 + s_] += v;
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) dw[i * 9 + t] = acc[t];
}

// 2x2 / stride 2 max pooling, NHWC bf16, one thread per 8 output channels.
__device__ __forceinline__ void max8(float (&m)[8], const uint4& u) {
  float f[8];
  unpack8(u, f);
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], f[k]);
}
__global__ void maxpool2x2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int H, int W, int CV,
                                  long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cv = static_cast<int>(i % CV);
  const long long pix = i / CV;
  const int Wo = W >> 1, Ho = H >> 1;
  const int wo = static_cast<int>(pix % Wo);
  const int ho = static_cast<int>((pix / Wo) % Ho);
  const long long n = pix / (static_cast<long long>(Wo) * Ho);
  const uint4* p00 = x + ((n * H + 2 * ho) * W + 2 * wo) * CV + cv;
  const uint4 a = __ldg(p00), b = __ldg(p00 + CV), c = __ldg(p00 + static_cast<long long>(W) * CV),
              d = __ldg(p00 + static_cast<long long>(W) * CV + CV);
  float m[8];
  unpack8(a, m);
  max8(m, b); max8(m, c); max8(m, d);
  y[i] = pack8(m);
}
// dx[window position] = dy if that position holds the FIRST maximum of its window in scan order
// (0,0),(0,1),(1,0),(1,1) - ATen's max_pool2d keeps the first index on ties (ReLU outputs tie at 0 a lot).
__global__ void maxpool2x2_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x,
                                      uint4* __restrict__ dx, int H, int W, int CV, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cv = static_cast<int>(i % CV);
  const long long pix = i / CV;
  const int Wo = W >> 1, Ho = H >> 1;
  const int wo = static_cast<int>(pix % Wo);
  const int ho = static_cast<int>((pix / Wo) % Ho);
  const long long n = pix / (static_cast<long long>(Wo) * Ho);
  const long long o00 = ((n * H + 2 * ho) * W + 2 * wo) * CV + cv;
  const long long offs[4] = {o00, o00 + CV, o00 + static_cast<long long>(W) * CV,
                             o00 + static_cast<long long>(W) * CV + CV};
  float f[4][8], g[8], o[4][8];
#pragma unroll
  for (int q = 0; q < 4; ++q) unpack8(__ldg(x + offs[q]), f[q]);
  unpack8(__ldg(dy + i), g);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int best = 0;
    float m = f[0][k];
#pragma unroll
    for (int q = 1; q < 4; ++q)
      if (f[q][k] > m) { m = f[q][k]; best = q; }
#pragma unroll
    for (int q = 0; q < 4; ++q) o[q][k] = (q == best) ? g[k] : 0.f;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) dx[offs[q]] = pack8(o[q]);
}
// dx = dy where the ReLU output y is positive, else 0 (8 bf16 per thread).
__global__ void relu_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y, uint4* __restrict__ dx,
                                long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float g[8], v[8];
  unpack8(__ldg(dy + i), g);
  unpack8(__ldg(y + i), v);
#pragma unroll
  for (int k = 0; k < 8; ++k) g[k] = v[k] > 0.f ? g[k] : 0.f;
  dx[i] = pack8(g);
}

// dx = dy where the LeakyReLU output y is positive, else slope * dy (slope > 0 keeps the sign of the input).
__global__ void lrelu_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y, uint4* __restrict__ dx,
                                 float slope, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float g[8], v[8];
  unpack8(__ldg(dy + i), g);
  unpack8(__ldg(y + i), v);
#pragma unroll
  for (int k = 0; k < 8; ++k) g[k] = v[k] > 0.f ? g[k] : slope * g[k];
  dx[i] = pack8(g);
}

// Window gather ("im2col") for the few-channel edge convolutions of the PatchGAN (4x4 windows,
// modules/discriminator/model.py:37,66): dst[n, oh, ow, t*Cs + c] = src[n, oh*stride + sgn*r + off, ow*stride +
// sgn*s + off, c] with t = r*K + s (zero outside the source, zero in the unused columns up to 64).
//   forward window of a KxK stride-`stride` pad-`pad` convolution: sgn = +1, off = -pad
//   window of its data gradient (stride 1):                       sgn = -1, off = +pad
__global__ void im2col_window_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int Hs,
                                     int Ws, int Ho, int Wo, int Cs, int K, int stride, int sgn, int off,
                                     long long total_vecs) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total_vecs) return;
  const int cv = static_cast<int>(i & 7);
  const long long pix = i >> 3;
  const int w = static_cast<int>(pix % Wo);
  const int h = static_cast<int>((pix / Wo) % Ho);
  const long long n = pix / (static_cast<long long>(Wo) * Ho);
  const __nv_bfloat16* img = src + n * Hs * Ws * Cs;
  __align__(16) __nv_bfloat16 vals[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int col = cv * 8 + k;
    __nv_bfloat16 v = __float2bfloat16_rn(0.f);
    if (col < K * K * Cs) {
      const int t = col / Cs, c = col - t * Cs;
      const int r = t / K, sx = t - K * r;
      const int hh = h * stride + sgn * r + off, ww = w * stride + sgn * sx + off;
      if (hh >= 0 && hh < Hs && ww >= 0 && ww < Ws) v = img[(static_cast<long long>(hh) * Ws + ww) * Cs + c];
    }
    vals[k] = v;
  }
  reinterpret_cast<uint4*>(dst)[i] = *reinterpret_cast<const uint4*>(vals);
}

}  // namespace b2

using namespace b2;

template <typename K, typename... Args>
static int launch1d(K kernel, long long total, cudaStream_t stream, Args... args) {
  if (total <= 0) return 0;
  kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(args...);
  return (int)cudaGetLastError();
}

extern "C" {

int b2dq_version() { return 0; }

int b2dq_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int N, int C, int HW, cudaStream_t st) {
  if (C <= 4)
    return launch1d(transpose_fewc_kernel<float, __nv_bfloat16, true>, (long long)N * HW, st, src,
                    reinterpret_cast<__nv_bfloat16*>(dst), C, HW, (long long)N * HW);
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
  transpose_chw_hwc_kernel<float, __nv_bfloat16><<<grid, block, 0, st>>>(
      src, reinterpret_cast<__nv_bfloat16*>(dst), C, HW);
  return (int)cudaGetLastError();
}
int b2dq_nchw_f32_to_nhwc_f32(const float* src, float* dst, int N, int C, int HW, cudaStream_t st) {
  if (C <= 4)
    return launch1d(transpose_fewc_kernel<float, float, true>, (long long)N * HW, st, src, dst, C, HW, (long long)N * HW);
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
  transpose_chw_hwc_kernel<float, float><<<grid, block, 0, st>>>(src, dst, C, HW);
  return (int)cudaGetLastError();
}
int b2dq_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int N, int C, int HW, cudaStream_t st) {
  if (C <= 4)
    return launch1d(transpose_fewc_kernel<__nv_bfloat16, float, false>, (long long)N * HW, st,
                    reinterpret_cast<const __nv_bfloat16*>(src), dst, C, HW, (long long)N * HW);
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
  transpose_hwc_chw_kernel<__nv_bfloat16, float><<<grid, block, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), dst, C, HW);
  return (int)cudaGetLastError();
}
int b2dq_nhwc_f32_to_nchw_f32(const float* src, float* dst, int N, int C, int HW, cudaStream_t st) {
  if (C <= 4)
    return launch1d(transpose_fewc_kernel<float, float, false>, (long long)N * HW, st, src, dst, C, HW, (long long)N * HW);
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
  transpose_hwc_chw_kernel<float, float><<<grid, block, 0, st>>>(src, dst, C, HW);
  return (int)cudaGetLastError();
}

int b2dq_upsample2x(const void* in, void* out, int N, int H, int W, int C, cudaStream_t st) {
  if (C % 8) return -1;
  const int vecs = C / 8;
  const long long total = (long long)N * 4 * H * W * vecs;
  return launch1d(upsample2x_kernel, total, st, reinterpret_cast<const uint4*>(in),
                  reinterpret_cast<uint4*>(out), H, W, vecs, total);
}
int b2dq_upsample2x_bwd(const void* gout, void* gin, int N, int H, int W, int C, cudaStream_t st) {
  if (C % 8) return -1;
  const int vecs = C / 8;
  const long long total = (long long)N * H * W * vecs;
  return launch1d(upsample2x_bwd_kernel, total, st, reinterpret_cast<const uint4*>(gout),
                  reinterpret_cast<uint4*>(gin), H, W, vecs, total);
}

int b2dq_softmax_rows(const void* s, void* p, long long rows, int T, int in_f32, cudaStream_t st) {
  if (T % 2) return -1;
  if (in_f32)
    return launch1d(softmax_rows_kernel<float>, rows * 32, st, reinterpret_cast<const float*>(s),
                    reinterpret_cast<__nv_bfloat16*>(p), rows, T);
  return launch1d(softmax_rows_kernel<__nv_bfloat16>, rows * 32, st,
                  reinterpret_cast<const __nv_bfloat16*>(s), reinterpret_cast<__nv_bfloat16*>(p),
                  rows, T);
}
int b2dq_softmax_bwd_rows(const void* p, const void* dp, void* ds, long long rows, int T,
                          float scale, cudaStream_t st) {
  if (T % 2) return -1;
  return launch1d(softmax_bwd_rows_kernel, rows * 32, st, reinterpret_cast<const __nv_bfloat16*>(p),
                  reinterpret_cast<const __nv_bfloat16*>(dp), reinterpret_cast<__nv_bfloat16*>(ds),
                  rows, T, scale);
}

int b2dq_add_bf16(const void* a, const void* b, void* o, long long n, cudaStream_t st) {
  if (n % 8) return -1;
  return launch1d(add_bf16_kernel, n / 8, st, reinterpret_cast<const uint4*>(a),
                  reinterpret_cast<const uint4*>(b), reinterpret_cast<uint4*>(o), n / 8);
}

int b2dq_cast_f32_to_bf16(const float* a, void* o, long long n, cudaStream_t st) {
  return launch1d(cast_f32_bf16_kernel, n, st, a, reinterpret_cast<__nv_bfloat16*>(o), n);
}

int b2dq_im2col3x3_small(const void* src, void* dst, int N, int H, int W, int Cs, int flip,
                         cudaStream_t st) {
  if (9 * Cs > 64) return -1;
  const long long total = (long long)N * H * W * 8;
  if (Cs == 3)
    return launch1d(im2col3x3_c3_kernel, total / 8, st, reinterpret_cast<const unsigned short*>(src),
                    reinterpret_cast<uint4*>(dst), H, W, flip, total / 8);
  return launch1d(im2col3x3_small_kernel<0>, total, st, reinterpret_cast<const __nv_bfloat16*>(src),
                  reinterpret_cast<__nv_bfloat16*>(dst), H, W, Cs, flip, total);
}

// Number of row blocks b2dq_bias_grad uses (scratch = blocks * C floats).
int b2dq_bias_grad_blocks(long long rows) {
  if (rows <= 0) return 0;
  long long rpb = (rows + 148 * 8 - 1) / (148 * 8);
  if (rpb < 64) rpb = 64;
  return (int)((rows + rpb - 1) / rpb);
}

int b2dq_bias_grad(const void* dy, float* out, float* part, long long rows, int C, cudaStream_t st) {
  if (rows <= 0) return 0;
  long long rpb = (rows + 148 * 8 - 1) / (148 * 8);
  if (rpb < 64) rpb = 64;
  const unsigned blocks = (unsigned)((rows + rpb - 1) / rpb);
  if (C % 8 == 0 && C / 8 <= 256) {
    // threads = the largest multiple of the vectors per row that fits 256 (e.g. the 768 / 1536 channels of the
    // fused q|k|v gradient: 96 / 192 vectors -> 192 threads)
    const int threads = (256 / (C / 8)) * (C / 8);
    bias_grad_kernel<<<blocks, threads, threads * 8 * sizeof(float), st>>>(
        reinterpret_cast<const __nv_bfloat16*>(dy), part, rows, C, (int)rpb);
  } else {
    bias_grad_generic_kernel<<<blocks, 128, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dy), part,
                                                     rows, C, (int)rpb);
  }
  bias_grad_reduce_kernel<<<(C + 31) / 32, 256, 0, st>>>(part, out, (int)blocks, C);
  return (int)cudaGetLastError();
}

int b2dq_maxpool2x2(const void* x, void* y, int N, int H, int W, int C, cudaStream_t st) {
  if (N <= 0) return 0;
  if (H % 2 || W % 2 || C % 8) return -1;
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
  return launch1d(maxpool2x2_kernel, total, st, reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), H, W,
                  C / 8, total);
}
int b2dq_maxpool2x2_bwd(const void* dy, const void* x, void* dx, int N, int H, int W, int C, cudaStream_t st) {
  if (N <= 0) return 0;
  if (H % 2 || W % 2 || C % 8) return -1;
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
  return launch1d(maxpool2x2_bwd_kernel, total, st, reinterpret_cast<const uint4*>(dy),
                  reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(dx), H, W, C / 8, total);
}
int b2dq_relu_bwd(const void* dy, const void* y, void* dx, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  if (n % 8) return -1;
  return launch1d(relu_bwd_kernel, n / 8, st, reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(y),
                  reinterpret_cast<uint4*>(dx), n / 8);
}

int b2dq_lrelu_bwd(const void* dy, const void* y, void* dx, long long n, float slope, cudaStream_t st) {
  if (n <= 0) return 0;
  if (n % 8) return -1;
  return launch1d(lrelu_bwd_kernel, n / 8, st, reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(y),
                  reinterpret_cast<uint4*>(dx), slope, n / 8);
}

// src [N,Hs,Ws,Cs] bf16 -> dst [N,Ho,Wo,64] bf16, K*K*Cs <= 64 (see im2col_window_kernel).
int b2dq_im2col_window(const void* src, void* dst, int N, int Hs, int Ws, int Ho, int Wo, int Cs, int K, int stride,
                       int sgn, int off, cudaStream_t st) {
  if (K <= 0 || Cs <= 0 || K * K * Cs > 64 || (sgn != 1 && sgn != -1) || stride <= 0) return -1;
  const long long total = (long long)N * Ho * Wo * 8;
  if (total <= 0) return 0;
  return launch1d(im2col_window_kernel, total, st, reinterpret_cast<const __nv_bfloat16*>(src),
                  reinterpret_cast<__nv_bfloat16*>(dst), Hs, Ws, Ho, Wo, Cs, K, stride, sgn, off, total);
}

// weight [Cout,Cin,R,S] fp32 -> fwd [Cout, R*S*Cin] and/or dgrad [Cin, R*S*Cout] bf16 (null = skip).
int b2dq_pack_weights(const float* weight, void* fwd, void* dgrad, int Cout, int Cin, int R, int S,
                      cudaStream_t st) {
  if (Cout <= 0 || Cin <= 0 || R <= 0 || S <= 0) return -1;
  const long long total = (long long)Cout * Cin * R * S;
  return launch1d(pack_weights_kernel, total, st, weight, reinterpret_cast<__nv_bfloat16*>(fwd),
                  reinterpret_cast<__nv_bfloat16*>(dgrad), Cout, Cin, R * S, total);
}

int b2dq_upconv_pack(const float* weight, void* fwd, void* dgrad, int Cout, int Cin, cudaStream_t st) {
  if (Cout <= 0 || Cin <= 0) return -1;
  return launch1d(upconv_pack_kernel, (long long)Cout * Cin, st, weight, reinterpret_cast<__nv_bfloat16*>(fwd),
                  reinterpret_cast<__nv_bfloat16*>(dgrad), Cout, Cin);
}
int b2dq_upconv_wgrad_reduce(const float* partial, float* dw, int splits, int Cout, int Cin, cudaStream_t st) {
  if (Cout <= 0 || Cin <= 0 || splits <= 0) return -1;
  return launch1d(upconv_wgrad_reduce_kernel, (long long)Cout * Cin, st, partial, dw, splits, Cout, Cin);
}

// items_dev: device array of n_items records of 6 x int64 {weight ptr, fwd ptr (or 0), dgrad ptr (or 0),
// Cout | Cin << 32, R*S | tiles_ci << 32, tile_start} with tiles_ci = ceil(Cin / 32), tile_start = running sum of
// ceil(Cout/32) * tiles_ci; total_tiles = that sum over all items; max_rs = the largest R*S (<= 16).
int b2dq_pack_weights_multi(const void* items_dev, int n_items, long long total_tiles, int max_rs, cudaStream_t st) {
  if (n_items <= 0 || total_tiles <= 0) return 0;
  if (max_rs < 1 || max_rs > PK_MAX_RS || total_tiles > 0x7fffffffll) return -1;
  static_assert(sizeof(PackItem) == 48, "PackItem must be 6 x 8 bytes");
  const size_t smem = static_cast<size_t>(PK_T) * (PK_T * max_rs + 1) * sizeof(float);
  static unsigned long long attr_mask = 0;
  if (smem > 48 * 1024)
    if (int e = set_max_smem_once(pack_weights_multi_kernel, 72 * 1024, attr_mask)) return e;
  pack_weights_multi_kernel<<<(unsigned)total_tiles, 256, smem, st>>>(reinterpret_cast<const PackItem*>(items_dev),
                                                                      n_items);
  return (int)cudaGetLastError();
}

}  // extern "C"
