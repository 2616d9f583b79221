// Operator-level entry points of libb200dq.so: one call per nn.Conv2d / GroupNorm(+swish) / attention core
// forward or backward.  Everything below is host code: it derives the tile shapes, tap tables, split-K factors
// and kernel choice from the operator geometry and enqueues the descriptor-level kernels of the same library
// (b2dq_tapgemm, b2dq_pconv3x3, b2dq_mmgemm, b2dq_gn_*), so a binding in another host language needs no copy of
// the heuristics that live in dynamicvectorquantization_b200/kernels.py.  The same inputs give bit-identical
// results on both routes (tests/test_gpu_oplevel.py).
//
// Reference operators replaced: nn.Conv2d as used by modules/diffusionmodules/model.py:43-47,62-72,88-115,146-165;
// Normalize + nonlinearity (model.py:29-35); AttnBlock.forward's softmax(q k^T / sqrt(C)) v (model.py:176-188).
#include <cstdint>
#include <cstring>
#include "../../include/b200dq.h"

namespace {

constexpr int kSMs = 148;

int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

void tile_shape(int wout, int hout, int pixels, int* tw, int* th, int* tn) {
  *tw = pow2ceil(wout) < pixels ? pow2ceil(wout) : pixels;
  const int rest = pixels / *tw;
  *th = pow2ceil(hout) < rest ? pow2ceil(hout) : rest;
  *tn = pixels / (*tw * *th);
}

int ceil_div(int a, int b) { return (a + b - 1) / b; }

int pick_block_n(long long m_tiles, int cout) {
  if (cout % 256 == 0 && m_tiles * (cout / 256) >= 4 * kSMs) return 256;
  if (cout <= 16) return 16;
  if (cout <= 64) return 64;
  return 128;
}

bool geom_ok(const b2dq_conv2d_geom* g) {
  if (!g || g->N <= 0 || g->H <= 0 || g->W <= 0 || g->Cin <= 0 || g->Cout <= 0) return false;
  if (g->ksize != 1 && g->ksize != 3) return false;
  if (g->stride == 1) return true;
  return g->stride == 2 && g->ksize == 3 && g->H % 2 == 0 && g->W % 2 == 0;
}

// persistent strip kernel: 3x3 stride 1 to 128 channels over rows of a multiple of 128 pixels, >= 2 tiles per SM
bool pconv_ok(int ksize, int stride, int w, int cin, int cout, int nb, int h) {
  return ksize == 3 && stride == 1 && cout == 128 && w % 128 == 0 && cin % 64 == 0 &&
         (long long)nb * h * (w / 128) >= 2 * kSMs;
}

void out_hw(const b2dq_conv2d_geom* g, int* ho, int* wo) {
  *ho = g->H / g->stride;
  *wo = g->W / g->stride;
}

// (c, w, p, h, n) view of a contiguous NHWC tensor
void nhwc_view(int nb, int h, int w, int c, long long* dims, long long* strs) {
  dims[0] = c; dims[1] = w; dims[2] = 1; dims[3] = h; dims[4] = nb;
  strs[0] = 1; strs[1] = c; strs[2] = (long long)w * c; strs[3] = (long long)w * c; strs[4] = (long long)h * w * c;
}

// stride-2 view [N, H/2, 2, W/2, 2C]
void parity_view(int nb, int h, int w, int c, long long* dims, long long* strs) {
  dims[0] = 2 * c; dims[1] = w / 2; dims[2] = 2; dims[3] = h / 2; dims[4] = nb;
  strs[0] = 1; strs[1] = 2 * c; strs[2] = (long long)w * c; strs[3] = 2ll * w * c; strs[4] = (long long)h * w * c;
}

int wgrad_splits(int kblocks, int ctas_per_split) {
  int target = kSMs / (ctas_per_split > 0 ? ctas_per_split : 1);
  if (target < 1) target = 1;
  int cap = kblocks >= 4 ? kblocks / 4 : 1;
  int s = target < cap ? target : cap;
  return s < 1 ? 1 : s;
}

struct WgradPlan {
  int ntaps, kw, kh, kn, ktw, kth, kblocks, splits, fuse_bias, strip;
  long long partial_floats, colsum_floats, bias_part_floats;
};

WgradPlan wgrad_plan(const b2dq_conv2d_geom* g, int want_bias) {
  WgradPlan p;
  int ho, wo;
  out_hw(g, &ho, &wo);
  p.ntaps = g->ksize * g->ksize;
  tile_shape(wo, ho, 64, &p.kw, &p.kh, &p.kn);
  p.ktw = ceil_div(wo, p.kw);
  p.kth = ceil_div(ho, p.kh);
  p.kblocks = p.ktw * p.kth * ceil_div(g->N, p.kn);
  const int mt = ceil_div(g->Cout, 128), nt = ceil_div(g->Cin, 128), ngroups = (p.ntaps + 2) / 3;
  p.splits = wgrad_splits(p.kblocks, mt * nt * ngroups);
  p.fuse_bias = want_bias && p.ntaps >= 3;
  p.strip = g->ksize == 3 && g->stride == 1 && p.kw == 64 && p.kh == 1 && p.kn == 1;
  p.partial_floats = (long long)p.splits * p.ntaps * g->Cout * g->Cin;
  p.colsum_floats = p.fuse_bias ? (long long)p.splits * g->Cout : 0;
  p.bias_part_floats = (want_bias && !p.fuse_bias)
                           ? (long long)b2dq_bias_grad_blocks((long long)g->N * ho * wo) * g->Cout : 0;
  return p;
}

long long align256(long long v) { return (v + 255) & ~255ll; }

// batched out[z] = alpha * A B^T, A [M,K] K-major (or stored [K,M] when a_mn), rows `lda` elements apart, batches
// `sa` apart; same for B; out rows ldo apart, batches so apart
int batched_mm(const void* a, int a_mn, long long lda, long long sa, const void* b, int b_mn, long long ldb,
               long long sb, int M, int N, int K, void* out, long long ldo, long long so, int batches, float alpha,
               int out_f32, cudaStream_t stream) {
  b2dq_mm_desc d;
  std::memset(&d, 0, sizeof(d));
  d.a_ptr = a; d.b_ptr = b;
  const long long ad0 = a_mn ? M : K, ad1 = a_mn ? K : M, bd0 = b_mn ? N : K, bd1 = b_mn ? K : N;
  d.a_dims[0] = ad0; d.a_dims[1] = ad1; d.a_dims[2] = 1; d.a_dims[3] = 1; d.a_dims[4] = batches;
  d.a_strides[0] = 1; d.a_strides[1] = lda; d.a_strides[2] = sa; d.a_strides[3] = sa; d.a_strides[4] = sa;
  d.b_dims[0] = bd0; d.b_dims[1] = bd1; d.b_dims[2] = 1; d.b_dims[3] = 1; d.b_dims[4] = batches;
  d.b_strides[0] = 1; d.b_strides[1] = ldb; d.b_strides[2] = sb; d.b_strides[3] = sb; d.b_strides[4] = sb;
  d.a_mn = a_mn; d.b_mn = b_mn; d.ntaps = 1; d.taps_per_cta = 0;
  d.KW = 64; d.KH = 1; d.KN = 1;
  const int kb = (K + 63) / 64;
  d.ktiles_w = kb; d.ktiles_h = 1; d.kblocks = kb;
  d.splits = 1; d.batches = batches; d.M = M; d.N = N;
  d.out = out; d.oZ = so; d.oT = 0; d.oM = ldo;
  d.alpha = alpha; d.out_f32 = out_f32; d.block_n = 0; d.b_strip = 0; d.colsum = nullptr;
  return b2dq_mmgemm(&d, stream);
}

}  // namespace

extern "C" {

int b2dq_conv2d_out_hw(const b2dq_conv2d_geom* g, int* out2) {
  if (!geom_ok(g) || !out2) return -1;
  out_hw(g, &out2[0], &out2[1]);
  return 0;
}

int b2dq_conv2d_fwd_workspace_bytes(const b2dq_conv2d_geom* g) {
  if (!geom_ok(g)) return -1;
  if (!pconv_ok(g->ksize, g->stride, g->W, g->Cin, g->Cout, g->N, g->H)) return 0;
  return (int)((long long)g->N * g->H * (g->W / 128) * 64 * sizeof(float));
}

int b2dq_conv2d_fwd(const b2dq_conv2d_geom* g, const void* x, const void* wpack, const float* bias,
                    const void* residual, void* y, int act, float* gn_stats, void* ws, long long ws_bytes,
                    cudaStream_t stream) {
  if (!geom_ok(g) || g->Cin % 64) return -1;
  if (act == 0 && pconv_ok(g->ksize, g->stride, g->W, g->Cin, g->Cout, g->N, g->H)) {
    float* part = nullptr;
    if (gn_stats) {
      if (!ws || ws_bytes < b2dq_conv2d_fwd_workspace_bytes(g)) return -2;
      part = static_cast<float*>(ws);
    }
    int rc = b2dq_pconv3x3(x, wpack, y, bias, residual, part, g->N, g->H, g->W, g->Cin, 0, 0, stream);
    if (rc) return rc;
    if (gn_stats) rc = b2dq_gn_finalize_tiles(part, gn_stats, g->N, g->H, g->W, 1e-6f, stream);
    return rc;
  }
  if (gn_stats) return -3;     // only the strip kernel emits statistics: ask b2dq_conv2d_fwd_workspace_bytes first
  b2dq_tapgemm_desc d;
  std::memset(&d, 0, sizeof(d));
  int ho, wo;
  out_hw(g, &ho, &wo);
  const int cin = g->Cin;
  d.a_ptr = x;
  if (g->stride == 1) {
    nhwc_view(g->N, g->H, g->W, cin, d.a_dims, d.a_strides);
    if (g->ksize == 3) {
      for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) {
          const int t = r * 3 + s;
          d.tap_w[t] = s - 1; d.tap_h[t] = r - 1; d.tap_bk[t] = t * cin;
        }
      d.num_taps = 9;
    } else {
      d.num_taps = 1;
    }
  } else {
    parity_view(g->N, g->H, g->W, cin, d.a_dims, d.a_strides);
    for (int r = 0; r < 3; ++r)
      for (int s = 0; s < 3; ++s) {
        const int t = r * 3 + s;
        d.tap_c[t] = (s % 2) * cin; d.tap_w[t] = s / 2; d.tap_p[t] = r % 2; d.tap_h[t] = r / 2; d.tap_bk[t] = t * cin;
      }
    d.num_taps = 9;
  }
  d.kchunks = cin / 64;
  d.b_ptr = wpack; d.b_rows = g->Cout; d.b_k = (long long)g->ksize * g->ksize * cin; d.b_batch = 1;
  tile_shape(wo, ho, 128, &d.TW, &d.TH, &d.TN);
  d.Wout = wo; d.Hout = ho; d.NB = g->N; d.Cout = g->Cout;
  d.out = y;
  d.oN = (long long)ho * wo * g->Cout; d.oH = (long long)wo * g->Cout; d.oW = g->Cout;
  d.bias = bias; d.residual = residual; d.rN = d.oN; d.rH = d.oH; d.rW = d.oW;
  d.alpha = 1.f; d.out_f32 = 0;
  d.block_n = pick_block_n((long long)ceil_div(wo, d.TW) * ceil_div(ho, d.TH) * ceil_div(g->N, d.TN), g->Cout);
  d.m_tiles_per_cta = 0; d.relu = act;
  return b2dq_tapgemm(&d, stream);
}

int b2dq_conv2d_dgrad(const b2dq_conv2d_geom* g, const void* dy, const void* wpack_dgrad, void* dx,
                      cudaStream_t stream) {
  if (!geom_ok(g) || g->Cout % 64) return -1;
  int ho, wo;
  out_hw(g, &ho, &wo);
  const int cout = g->Cout, cin = g->Cin;
  if (pconv_ok(g->ksize, g->stride, wo, cout, cin, g->N, ho))
    return b2dq_pconv3x3(dy, wpack_dgrad, dx, nullptr, nullptr, nullptr, g->N, ho, wo, cout, 1, 0, stream);
  b2dq_tapgemm_desc d;
  std::memset(&d, 0, sizeof(d));
  d.a_ptr = dy;
  nhwc_view(g->N, ho, wo, cout, d.a_dims, d.a_strides);
  d.kchunks = cout / 64;
  d.b_ptr = wpack_dgrad; d.b_rows = cin; d.b_k = (long long)g->ksize * g->ksize * cout; d.b_batch = 1;
  d.NB = g->N; d.Cout = cin;
  d.alpha = 1.f;
  if (g->stride == 1) {
    if (g->ksize == 3) {
      for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) {
          const int t = r * 3 + s;
          d.tap_w[t] = 1 - s; d.tap_h[t] = 1 - r; d.tap_bk[t] = t * cout;
        }
      d.num_taps = 9;
    } else {
      d.num_taps = 1;
    }
    tile_shape(g->W, g->H, 128, &d.TW, &d.TH, &d.TN);
    d.Wout = g->W; d.Hout = g->H;
    d.out = dx;
    d.oN = (long long)g->H * g->W * cin; d.oH = (long long)g->W * cin; d.oW = cin;
    d.block_n = pick_block_n((long long)ceil_div(g->W, d.TW) * ceil_div(g->H, d.TH) * ceil_div(g->N, d.TN), cin);
    return b2dq_tapgemm(&d, stream);
  }
  // y[oh,ow] = sum W[r,s] x[2oh+r, 2ow+s]: input pixel (2i+ph, 2j+pw) gathers the taps with r = ph (mod 2)
  static const int sel[2][2][2] = {{{0, 0}, {2, -1}}, {{1, 0}, {0, 0}}};   // [parity][slot] = (filter index, offset)
  static const int nsel[2] = {2, 1};
  tile_shape(wo, ho, 128, &d.TW, &d.TH, &d.TN);
  d.Wout = wo; d.Hout = ho;
  d.oN = (long long)g->H * g->W * cin; d.oH = 2ll * g->W * cin; d.oW = 2 * cin;
  d.block_n = pick_block_n((long long)ceil_div(wo, d.TW) * ceil_div(ho, d.TH) * ceil_div(g->N, d.TN), cin);
  // wide 128-channel outputs: the persistent strip kernel (one strip serves a class's column taps)
  const bool strips = cin == 128 && wo % 128 == 0 && cout % 64 == 0 && (long long)g->N * ho * (wo / 128) >= 2 * kSMs;
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      if (strips) {
        b2dq_pconv_taps_desc q;
        std::memset(&q, 0, sizeof(q));
        q.a_ptr = dy;
        nhwc_view(g->N, ho, wo, cout, q.a_dims, q.a_strides);
        q.b_ptr = wpack_dgrad; q.b_k = 9ll * cout; q.kchunks = cout / 64;
        q.nr = nsel[ph]; q.ns = nsel[pw];
        for (int a = 0; a < q.nr; ++a) {
          q.row_dh[a] = sel[ph][a][1];
          for (int b = 0; b < q.ns; ++b) q.wcol[a * 3 + b] = (sel[ph][a][0] * 3 + sel[pw][b][0]) * cout;
        }
        for (int b = 0; b < q.ns; ++b) q.col_dw[b] = sel[pw][b][1];
        q.NB = g->N; q.H = ho; q.W = wo;
        q.out = static_cast<char*>(dx) + 2ll * ((long long)(ph * g->W + pw) * cin);
        q.oN = d.oN; q.oH = d.oH; q.oW = d.oW;
        const int rc = b2dq_pconv_taps(&q, 0, stream);
        if (rc) return rc;
        continue;
      }
      int t = 0;
      for (int a = 0; a < nsel[ph]; ++a)
        for (int b = 0; b < nsel[pw]; ++b, ++t) {
          const int r = sel[ph][a][0], dh = sel[ph][a][1], s = sel[pw][b][0], dw = sel[pw][b][1];
          d.tap_c[t] = 0; d.tap_p[t] = 0; d.tap_w[t] = dw; d.tap_h[t] = dh; d.tap_bk[t] = (r * 3 + s) * cout;
        }
      d.num_taps = t;
      d.out = static_cast<char*>(dx) + 2ll * ((long long)(ph * g->W + pw) * cin);
      const int rc = b2dq_tapgemm(&d, stream);
      if (rc) return rc;
    }
  return 0;
}

int b2dq_conv2d_wgrad_workspace_bytes(const b2dq_conv2d_geom* g, int want_bias) {
  if (!geom_ok(g)) return -1;
  const WgradPlan p = wgrad_plan(g, want_bias);
  const long long bytes = align256(p.partial_floats * 4) + align256(p.colsum_floats * 4) + align256(p.bias_part_floats * 4);
  return bytes > 0x7fffffffll ? -1 : (int)bytes;
}

int b2dq_conv2d_wgrad(const b2dq_conv2d_geom* g, const void* x, const void* dy, float* dw, float* db, void* ws,
                      long long ws_bytes, cudaStream_t stream) {
  if (!geom_ok(g) || !dw) return -1;
  const int want_bias = db != nullptr;
  const WgradPlan p = wgrad_plan(g, want_bias);
  if (!ws || ws_bytes < b2dq_conv2d_wgrad_workspace_bytes(g, want_bias)) return -2;
  float* partial = static_cast<float*>(ws);
  float* colsum = reinterpret_cast<float*>(static_cast<char*>(ws) + align256(p.partial_floats * 4));
  float* bias_part = reinterpret_cast<float*>(reinterpret_cast<char*>(colsum) + align256(p.colsum_floats * 4));
  int ho, wo;
  out_hw(g, &ho, &wo);
  b2dq_mm_desc d;
  std::memset(&d, 0, sizeof(d));
  d.a_ptr = dy;
  nhwc_view(g->N, ho, wo, g->Cout, d.a_dims, d.a_strides);
  d.b_ptr = x;
  if (g->stride == 1) {
    nhwc_view(g->N, g->H, g->W, g->Cin, d.b_dims, d.b_strides);
    if (g->ksize == 3)
      for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) { d.tap_w[r * 3 + s] = s - 1; d.tap_h[r * 3 + s] = r - 1; }
  } else {
    parity_view(g->N, g->H, g->W, g->Cin, d.b_dims, d.b_strides);
    for (int r = 0; r < 3; ++r)
      for (int s = 0; s < 3; ++s) {
        const int t = r * 3 + s;
        d.tap_c[t] = (s % 2) * g->Cin; d.tap_w[t] = s / 2; d.tap_p[t] = r % 2; d.tap_h[t] = r / 2;
      }
  }
  d.a_mn = 1; d.b_mn = 1; d.ntaps = p.ntaps; d.taps_per_cta = p.ntaps < 3 ? p.ntaps : 3;
  d.KW = p.kw; d.KH = p.kh; d.KN = p.kn;
  d.ktiles_w = p.ktw; d.ktiles_h = p.kth; d.kblocks = p.kblocks;
  d.splits = p.splits; d.batches = 1; d.M = g->Cout; d.N = g->Cin;
  d.out = partial;
  d.oZ = (long long)p.ntaps * g->Cout * g->Cin; d.oT = (long long)g->Cout * g->Cin; d.oM = g->Cin;
  d.alpha = 1.f; d.out_f32 = 1; d.block_n = 128; d.b_strip = p.strip;
  d.colsum = p.fuse_bias ? colsum : nullptr;
  int rc = b2dq_mmgemm(&d, stream);
  if (rc) return rc;
  if (p.fuse_bias) return b2dq_wgrad_reduce_bias(partial, dw, p.splits, p.ntaps, g->Cout, g->Cin, colsum, db, stream);
  rc = b2dq_wgrad_reduce(partial, dw, p.splits, p.ntaps, g->Cout, g->Cin, 0, stream);
  if (rc || !want_bias) return rc;
  return b2dq_bias_grad(dy, db, bias_part, (long long)g->N * ho * wo, g->Cout, stream);
}

// ------------------------------------------------------------------------------------------ GroupNorm (+ swish)
int b2dq_groupnorm_workspace_bytes(int N, int HW, int C, int G, int backward) {
  if (N <= 0 || HW <= 0 || C <= 0 || G <= 0 || C % G) return -1;
  const int fused = backward ? b2dq_gn_bwd_fused_workspace_bytes(N, HW, C, G) : b2dq_gn_fwd_fused_workspace_bytes(N, HW, C, G);
  const long long chunks = b2dq_gn_chunks(N, HW);
  const long long split = backward ? align256((long long)N * chunks * C * 2 * 4) + align256((long long)N * C * 2 * 4)
                                   : align256((long long)N * chunks * G * 2 * 4);
  const long long bytes = fused > split ? fused : split;
  return bytes > 0x7fffffffll ? -1 : (int)bytes;
}

int b2dq_groupnorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stats, void* ws,
                       long long ws_bytes, int N, int HW, int C, int G, float eps, int act, cudaStream_t stream) {
  if (!ws || ws_bytes < b2dq_groupnorm_workspace_bytes(N, HW, C, G, 0)) return -2;
  if ((act == 0 || act == 1) && b2dq_gn_fwd_fused_workspace_bytes(N, HW, C, G) > 0)
    return b2dq_gn_fwd_fused(x, gamma, beta, y, stats, ws, ws_bytes, N, HW, C, G, eps, act, stream);
  const int rc = b2dq_gn_stats(x, stats, static_cast<float*>(ws), N, HW, C, G, eps, stream);
  if (rc) return rc;
  return b2dq_gn_apply(x, stats, gamma, beta, y, N, HW, C, G, act, stream);
}

int b2dq_groupnorm_bwd(const void* dy, const void* x, const float* stats, const float* gamma, const float* beta,
                       void* dx, float* dgb, const void* add, void* ws, long long ws_bytes, int N, int HW, int C,
                       int G, int act, cudaStream_t stream) {
  if (!ws || ws_bytes < b2dq_groupnorm_workspace_bytes(N, HW, C, G, 1)) return -2;
  if ((act == 0 || act == 1) && b2dq_gn_bwd_fused_workspace_bytes(N, HW, C, G) > 0)
    return b2dq_gn_bwd_fused(dy, x, stats, gamma, beta, dx, dgb, add, ws, ws_bytes, N, HW, C, G, act, stream);
  const long long chunks = b2dq_gn_chunks(N, HW);
  float* part = static_cast<float*>(ws);
  float* ws_nc = reinterpret_cast<float*>(static_cast<char*>(ws) + align256((long long)N * chunks * C * 2 * 4));
  const int rc = b2dq_gn_bwd_stats(dy, x, stats, gamma, beta, part, ws_nc, N, HW, C, G, act, stream);
  if (rc) return rc;
  return b2dq_gn_bwd_apply(dy, x, stats, gamma, beta, ws_nc, dx, dgb, add, N, HW, C, G, act, stream);
}

// Names of SURVEY 8b: GroupNorm(32, 1e-6) + swish, the pair every ResnetBlock / AttnBlock / output head uses.
int b2dq_groupnorm_swish_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stats, void* ws,
                             long long ws_bytes, int N, int HW, int C, cudaStream_t stream) {
  return b2dq_groupnorm_fwd(x, gamma, beta, y, stats, ws, ws_bytes, N, HW, C, 32, 1e-6f, 1, stream);
}
int b2dq_groupnorm_swish_bwd(const void* dy, const void* x, const float* stats, const float* gamma, const float* beta,
                             void* dx, float* dgb, const void* add, void* ws, long long ws_bytes, int N, int HW, int C,
                             cudaStream_t stream) {
  return b2dq_groupnorm_bwd(dy, x, stats, gamma, beta, dx, dgb, add, ws, ws_bytes, N, HW, C, 32, 1, stream);
}

// ------------------------------------------------------------------------------------------ attention core
// qkv [N][T][3C] bf16: q | k | v side by side (the output of ONE [C -> 3C] 1x1 convolution).
int b2dq_attention_workspace_bytes(int N, int T, int C, int backward) {
  if (N <= 0 || T <= 0 || C <= 0) return -1;
  const long long tt = (long long)N * T * T;
  const long long bytes = backward ? 2 * align256(tt * 2) : align256(tt * 4);
  return bytes > 0x7fffffffll ? -1 : (int)bytes;
}

int b2dq_attention_fwd(const void* qkv, void* out, void* probs, void* ws, long long ws_bytes, int N, int T, int C,
                       float scale, cudaStream_t stream) {
  if (C % 8 || !qkv || !out || !probs) return -1;
  if (!ws || ws_bytes < b2dq_attention_workspace_bytes(N, T, C, 0)) return -2;
  const char* q = static_cast<const char*>(qkv);
  const char* k = q + 2ll * C;
  const char* v = q + 4ll * C;
  const long long ld = 3ll * C, sb = (long long)T * 3 * C;
  float* s = static_cast<float*>(ws);
  int rc = batched_mm(q, 0, ld, sb, k, 0, ld, sb, T, T, C, s, T, (long long)T * T, N, scale, 1, stream);
  if (rc) return rc;
  rc = b2dq_softmax_rows(s, probs, (long long)N * T, T, 1, stream);
  if (rc) return rc;
  return batched_mm(probs, 0, T, (long long)T * T, v, 1, ld, sb, T, C, T, out, C, (long long)T * C, N, 1.f, 0, stream);
}

int b2dq_attention_bwd(const void* qkv, const void* probs, const void* dout, void* dqkv, void* ws, long long ws_bytes,
                       int N, int T, int C, float scale, cudaStream_t stream) {
  if (C % 8 || !qkv || !probs || !dout || !dqkv) return -1;
  if (!ws || ws_bytes < b2dq_attention_workspace_bytes(N, T, C, 1)) return -2;
  const char* q = static_cast<const char*>(qkv);
  const char* k = q + 2ll * C;
  const char* v = q + 4ll * C;
  char* dq = static_cast<char*>(dqkv);
  char* dk = dq + 2ll * C;
  char* dv = dq + 4ll * C;
  const long long ld = 3ll * C, sb = (long long)T * 3 * C, tt = (long long)T * T;
  void* dp = ws;
  void* ds = static_cast<char*>(ws) + align256((long long)N * tt * 2);
  int rc = batched_mm(probs, 1, T, tt, dout, 1, C, (long long)T * C, T, C, T, dv, ld, sb, N, 1.f, 0, stream);  // dV = P^T dO
  if (rc) return rc;
  rc = batched_mm(dout, 0, C, (long long)T * C, v, 0, ld, sb, T, T, C, dp, T, tt, N, 1.f, 0, stream);            // dP = dO V^T
  if (rc) return rc;
  rc = b2dq_softmax_bwd_rows(probs, dp, ds, (long long)N * T, T, scale, stream);
  if (rc) return rc;
  rc = batched_mm(ds, 0, T, tt, k, 1, ld, sb, T, C, T, dq, ld, sb, N, 1.f, 0, stream);                           // dQ = dS K
  if (rc) return rc;
  return batched_mm(ds, 1, T, tt, q, 1, ld, sb, T, C, T, dk, ld, sb, N, 1.f, 0, stream);                         // dK = dS^T Q
}

}  // extern "C"
