// GroupNorm(32, eps 1e-6, affine)(+swish) BACKWARD as ONE persistent kernel: 2 reads + 1 write of HBM.
// Reference: the autograd of modules/diffusionmodules/model.py:29-35 (Normalize + nonlinearity) as used by
// ResnetBlock (:119-127), AttnBlock (:170) and the output heads (EncoderDual.py:116-117, DecoderPositional.py:142-143).
//
// The backward of GroupNorm needs two per-(image, group) sums over the whole image (S1 = sum dz*gamma,
// S2 = sum dz*gamma*xhat) before any dx can be written, so the separate-kernel version (norm.cu) reads dy and x
// twice from HBM: 5 passes over the tensor instead of the algorithmic 3.  Here a TEAM of CTAs owns one image at a
// time: every CTA streams its row slice of (dy, x) once for the sums (phase 1), the team meets at a per-image
// barrier in global memory (arrival counter + fixed-order sum of the per-CTA partials: deterministic), and each CTA
// then streams the SAME slice again for dx (phase 2) - the second read hits the 126 MB L2, because the number of
// images in flight (teams) is chosen so that their dy + x stay below an L2 budget.  HBM sees 2 reads + 1 write.
//
// Mechanics: 2 CTAs per SM (persistent, all co-resident: grid <= 2 x SM count), each 6 compute warps + 1 producer
// warp (+1 idle warp: 8 warps keep the register allocation at 128 per thread).  Two CTAs per SM belong to different
// teams whenever there are at least two teams, so one streams while the other waits at its team barrier.  The
// producer lane issues 1-D bulk async copies (cp.async.bulk, 6 KB per tensor and ring stage: a row slice of an
// NHWC image is contiguous, no tensor map needed) into a 5-stage shared-memory ring signalled by mbarriers and runs
// ahead of the consumers across the phase boundary and the team barrier.  Consumers keep every per-channel
// constant in registers, use packed fp32x2 arithmetic (fma.rn.f32x2: half the issue slots) and ONE MUFU op per
// sigmoid (sigmoid(z) = 0.5 + 0.5 tanh(z/2)) - the two-kernel version was instruction-bound at ~0.55 of HBM speed.
// dgamma / dbeta are accumulated per CTA over its images and combined by the last CTA to finish, in CTA order.
#include "common.cuh"

namespace b2 {

constexpr int GF_CONSUMERS = 256;                 // 8 compute warps
constexpr int GF_THREADS = GF_CONSUMERS + 32;     // + producer warp (warp 8)
constexpr int GF_VPT = 2;                         // 16 B vectors per consumer thread and stage
constexpr int GF_STAGES = 3;
constexpr int GF_CHUNK = GF_CONSUMERS * 16 * GF_VPT;   // bytes per tensor per stage (6 KB)
constexpr int GF_MAXC = 512;
constexpr int GF_MAXG = 64;
constexpr int GF_MAXS = 148;                      // CTAs per team

struct GnFusedParams {
  const __nv_bfloat16* dy;
  const __nv_bfloat16* x;
  const __nv_bfloat16* add;      // optional: summed into dx (gradient arriving through a skip connection)
  const float* stats;            // [N][G][2] mean, rstd
  const float* gamma;
  const float* beta;
  __nv_bfloat16* dx;
  float* dgb;                    // [2][C] dgamma, dbeta (overwritten)
  float* part;                   // [N][S][G][2] per-CTA group partials
  float* dgb_part;               // [grid][2][C] per-CTA (dgamma, dbeta) over its images
  float* team_part;              // [T][2][C] per-team sums of the above
  unsigned* flags;               // [N] arrival counters | [N] finished teams | [N+1] error flag | [N+2 .. N+2+T) finished
                                 // CTAs per team (all zeroed before launch)
  int N, HW, C, G, S, T, rows_per_cta;
};

// L2 eviction priorities: phase-1 reads are marked evict_last (the same bytes are read again after the team
// barrier), phase-2 reads and the residual-gradient read evict_first (never needed again by this kernel)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(pol)
      : "memory");
}
__device__ __forceinline__ float tanh_approx(float v) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(GF_CONSUMERS) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float2 unpack2(uint32_t u) { return make_float2(bf16_lo(u), bf16_hi(u)); }

// dz = dy * swish'(z) for a channel pair; zh = z / 2 = xhat * (gamma/2) + beta/2
template <bool SW>
__device__ __forceinline__ float2 dz_pair(float2 d, float2 xh, float2 Gh, float2 Bh) {
  if (!SW) return d;
  const float2 zh = __ffma2_rn(xh, Gh, Bh);
  const float2 t = make_float2(tanh_approx(zh.x), tanh_approx(zh.y));
  const float2 sg = __ffma2_rn(t, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
  const float2 om2 = __ffma2_rn(sg, make_float2(-2.f, -2.f), make_float2(2.f, 2.f));   // 2 (1 - sg)
  const float2 w = __ffma2_rn(zh, om2, make_float2(1.f, 1.f));                         // 1 + z (1 - sg)
  return __fmul2_rn(d, __fmul2_rn(sg, w));
}

// shared memory: ring [STAGES][3][CHUNK] | red [2][CONSUMERS][8] | chan [2][MAXC] | dg [2][MAXC] | sk [2][MAXG] | bars
constexpr int GF_SMEM = GF_STAGES * 3 * GF_CHUNK + (2 * GF_CONSUMERS * 8 + 4 * GF_MAXC + 2 * GF_MAXG) * 4 + 2 * GF_STAGES * 8;

template <bool SW>
__global__ void __launch_bounds__(GF_THREADS, 2) gn_bwd_fused_kernel(const GnFusedParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* ring = smem;
  float* red = reinterpret_cast<float*>(smem + GF_STAGES * 3 * GF_CHUNK);
  float* chan = red + 2 * GF_CONSUMERS * 8;
  float* dg = chan + 2 * GF_MAXC;
  float* sk = dg + 2 * GF_MAXC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sk + 2 * GF_MAXG);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + GF_STAGES);
  const int tid = threadIdx.x;
  const int C = p.C, G = p.G, HW = p.HW;
  const int vecs = C >> 3;
  const int rstep = GF_CONSUMERS / vecs;          // rows covered by one vector step of the compute warps
  const int srows = rstep * GF_VPT;               // rows per ring stage
  const int team = blockIdx.x / p.S, s_idx = blockIdx.x % p.S;
  const int r0 = s_idx * p.rows_per_cta;
  const int r1 = min(HW, r0 + p.rows_per_cta);
  const int nstages = (r1 - r0 + srows - 1) / srows;
  const bool has_add = p.add != nullptr;

  if (tid == 0) {
    for (int i = 0; i < GF_STAGES; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, GF_CONSUMERS / 32);
    }
    fence_barrier_init();
  }
  for (int i = tid; i < 2 * C; i += GF_THREADS) dg[i] = 0.f;
  __syncthreads();

  if (tid >= GF_CONSUMERS) {
    // ------------------------------------------------------------------ producer (one lane of warp 6)
    if (tid == GF_CONSUMERS) {
      uint32_t stage = 0, phase = 0;
      const uint64_t keep = l2_policy_evict_last(), drop = l2_policy_evict_first();
      for (int n = team; n < p.N; n += p.T) {
        const long long img = static_cast<long long>(n) * HW * C;
        for (int ph = 0; ph < 2; ++ph) {
          const bool with_add = ph == 1 && has_add;
          const uint64_t pol = ph == 0 ? keep : drop;
          for (int c = 0; c < nstages; ++c) {
            const int row = r0 + c * srows;
            const uint32_t bytes = static_cast<uint32_t>(min(srows, r1 - row)) * C * 2;
            const long long off = img + static_cast<long long>(row) * C;
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            const uint32_t fb = full0 + 8 * stage;
            const uint32_t dst = smem_u32(ring + stage * 3 * GF_CHUNK);
            mbar_arrive_expect_tx(fb, bytes * (with_add ? 3 : 2));
            bulk_load(dst, p.dy + off, bytes, fb, pol);
            bulk_load(dst + GF_CHUNK, p.x + off, bytes, fb, pol);
            if (with_add) bulk_load(dst + 2 * GF_CHUNK, p.add + off, bytes, fb, drop);
            if (++stage == GF_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ consumers (6 warps)
    const int v = tid % vecs, rlane = tid / vecs;
    const int cg = C / G;
    const int lane = tid & 31;
    uint32_t stage = 0, phase = 0;
    const float inv_cnt = 1.f / (static_cast<float>(HW) * cg);
    // image-independent per-channel constants: z/2 = xhat * (gamma/2) + beta/2
    float2 Gh[4], Bh[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      Gh[k] = make_float2(0.5f * p.gamma[v * 8 + 2 * k], 0.5f * p.gamma[v * 8 + 2 * k + 1]);
      Bh[k] = make_float2(0.5f * p.beta[v * 8 + 2 * k], 0.5f * p.beta[v * 8 + 2 * k + 1]);
    }
    for (int n = team; n < p.N; n += p.T) {
      // per-(image, group) constants: xhat = x*R + M.  C/G is even, so a channel pair lies in one group and the
      // constants are scalars (broadcast operands of the packed instructions)
      float R[4], M[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int g0 = (v * 8 + 2 * k) / cg;
        const float m0 = p.stats[(n * G + g0) * 2], s0 = p.stats[(n * G + g0) * 2 + 1];
        R[k] = s0;
        M[k] = -m0 * s0;
      }
      // ---------------- phase 1: sum_rows dz, sum_rows dz * xhat per channel
      float2 sa[4], sb[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { sa[k] = make_float2(0.f, 0.f); sb[k] = make_float2(0.f, 0.f); }
      for (int c = 0; c < nstages; ++c) {
        mbar_wait(full0 + 8 * stage, phase);
        const uint8_t* st = ring + stage * 3 * GF_CHUNK + tid * 16;
#pragma unroll
        for (int u = 0; u < GF_VPT; ++u) {
          if (r0 + c * srows + u * rstep + rlane < r1) {
            const uint4 ud = *reinterpret_cast<const uint4*>(st + u * (GF_CONSUMERS * 16));
            const uint4 ux = *reinterpret_cast<const uint4*>(st + u * (GF_CONSUMERS * 16) + GF_CHUNK);
            const uint32_t wd[4] = {ud.x, ud.y, ud.z, ud.w}, wx[4] = {ux.x, ux.y, ux.z, ux.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 xh = __ffma2_rn(unpack2(wx[k]), make_float2(R[k], R[k]), make_float2(M[k], M[k]));
              const float2 dz = dz_pair<SW>(unpack2(wd[k]), xh, Gh[k], Bh[k]);
              sa[k] = __fadd2_rn(sa[k], dz);
              sb[k] = __ffma2_rn(dz, xh, sb[k]);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * stage);
        if (++stage == GF_STAGES) { stage = 0; phase ^= 1; }
      }
      // ---------------- CTA reduction (fixed order) -> per-channel sums, per-group partials
      float* red_a = red;
      float* red_b = red + GF_CONSUMERS * 8;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        *reinterpret_cast<float2*>(red_a + rlane * C + v * 8 + 2 * k) = sa[k];
        *reinterpret_cast<float2*>(red_b + rlane * C + v * 8 + 2 * k) = sb[k];
      }
      consumer_sync();
      for (int o = tid; o < 2 * C; o += GF_CONSUMERS) {       // o < C: sum dz (-> dbeta), else sum dz*xhat (-> dgamma)
        const float* src = (o < C) ? red_a + o : red_b + (o - C);
        float t = 0.f;
#pragma unroll 4
        for (int rl = 0; rl < rstep; ++rl) t += src[rl * C];
        chan[(o < C) ? o : GF_MAXC + (o - C)] = t;
        dg[(o < C) ? C + o : (o - C)] += t;                    // dg[0..C) = dgamma, dg[C..2C) = dbeta
      }
      consumer_sync();
      if (tid < G) {
        float S1 = 0.f, S2 = 0.f;
        for (int cc = tid * cg; cc < (tid + 1) * cg; ++cc) {
          const float g_ = p.gamma[cc];
          S1 = fmaf(g_, chan[cc], S1);
          S2 = fmaf(g_, chan[GF_MAXC + cc], S2);
        }
        float* o = p.part + ((static_cast<long long>(n) * p.S + s_idx) * G + tid) * 2;
        __stcg(reinterpret_cast<float2*>(o), make_float2(S1, S2));
      }
      consumer_sync();
      // ---------------- team barrier on image n
      if (tid == 0) {
        __threadfence();
        red_release_add(p.flags + n, 1u);
        const long long t0 = clock64();
        while (ld_acquire(p.flags + n) < static_cast<unsigned>(p.S)) {
          if (clock64() - t0 > 4000000000LL) { atomicExch(p.flags + p.N + 1, 1u); break; }   // ~2 s: never hang the GPU
        }
      }
      consumer_sync();
      {
        // sum the S partials of every (group, 2) in a fixed order: thread (j, e4) takes CTAs j, j+16, ... of float4 e4
        const int e4n = (2 * G) / 4;                          // float4 per CTA row (16 for G = 32)
        const int lanes_s = GF_CONSUMERS / e4n;               // 16
        const int e4 = tid % e4n, j = tid / e4n;
        const float4* src = reinterpret_cast<const float4*>(p.part + static_cast<long long>(n) * p.S * G * 2) + e4;
        constexpr int MAXL = (GF_MAXS + 15) / 16;             // loads per thread, all issued before the first add
        float4 t[MAXL];
#pragma unroll
        for (int i = 0; i < MAXL; ++i) {
          const int s = j + i * lanes_s;
          t[i] = (s < p.S) ? __ldcg(src + static_cast<long long>(s) * e4n) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float4 acc = t[0];
#pragma unroll
        for (int i = 1; i < MAXL; ++i) { acc.x += t[i].x; acc.y += t[i].y; acc.z += t[i].z; acc.w += t[i].w; }
        reinterpret_cast<float4*>(red)[j * e4n + e4] = acc;
        consumer_sync();
        if (tid < 2 * G) {
          float tsum = 0.f;
#pragma unroll 4
          for (int jj = 0; jj < lanes_s; ++jj) tsum += red[jj * 2 * G + tid];
          sk[tid] = tsum * inv_cnt;                           // sk[g*2 + 0] = S1/cnt, sk[g*2 + 1] = S2/cnt
        }
        consumer_sync();
      }
      // ---------------- phase 2: dx = rstd * (dz*gamma - S1/cnt - xhat * S2/cnt) (+ add)
      // dx = R * (dz*gamma - k1 - xhat*k2) = 2R * (dz*(gamma/2) - k1/2 - xhat*k2/2): reuses Gh, three small constants
      float R2[4], K1[4], K2[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int g0 = (v * 8 + 2 * k) / cg;
        R2[k] = 2.f * R[k];
        K1[k] = -0.5f * sk[g0 * 2];
        K2[k] = -0.5f * sk[g0 * 2 + 1];
      }
      __nv_bfloat16* out = p.dx + static_cast<long long>(n) * HW * C + v * 8;
      for (int c = 0; c < nstages; ++c) {
        mbar_wait(full0 + 8 * stage, phase);
        const uint8_t* st = ring + stage * 3 * GF_CHUNK + tid * 16;
#pragma unroll
        for (int u = 0; u < GF_VPT; ++u) {
          const int row = r0 + c * srows + u * rstep + rlane;
          if (row < r1) {
            const uint4 ud = *reinterpret_cast<const uint4*>(st + u * (GF_CONSUMERS * 16));
            const uint4 ux = *reinterpret_cast<const uint4*>(st + u * (GF_CONSUMERS * 16) + GF_CHUNK);
            uint4 ua = make_uint4(0u, 0u, 0u, 0u);
            if (has_add) ua = *reinterpret_cast<const uint4*>(st + u * (GF_CONSUMERS * 16) + 2 * GF_CHUNK);
            const uint32_t wd[4] = {ud.x, ud.y, ud.z, ud.w}, wx[4] = {ux.x, ux.y, ux.z, ux.w};
            const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w};
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 xh = __ffma2_rn(unpack2(wx[k]), make_float2(R[k], R[k]), make_float2(M[k], M[k]));
              const float2 dz = dz_pair<SW>(unpack2(wd[k]), xh, Gh[k], Bh[k]);
              float2 r = __ffma2_rn(dz, Gh[k], make_float2(K1[k], K1[k]));
              r = __ffma2_rn(xh, make_float2(K2[k], K2[k]), r);
              const float2 r2 = make_float2(R2[k], R2[k]);
              r = has_add ? __ffma2_rn(r, r2, unpack2(wa[k])) : __fmul2_rn(r, r2);
              o[k] = pack_bf16x2(r.x, r.y);
            }
            __stcs(reinterpret_cast<uint4*>(out + static_cast<long long>(row) * C), make_uint4(o[0], o[1], o[2], o[3]));
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * stage);
        if (++stage == GF_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  }
  // ------------------------------------------------------------------ dgamma / dbeta, two fixed-order levels:
  // the last CTA of a team to finish sums the team's S per-CTA partials, the last team to finish sums the T team sums
  __syncthreads();
  float* mine = p.dgb_part + static_cast<long long>(blockIdx.x) * 2 * C;
  for (int i = tid; i < 2 * C; i += GF_THREADS) __stcg(mine + i, dg[i]);
  __shared__ unsigned s_last;
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(p.flags + p.N + 2 + team, 1u) == static_cast<unsigned>(p.S) - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  {
    const float* src = p.dgb_part + static_cast<long long>(team) * p.S * 2 * C;
    for (int i = tid; i < 2 * C; i += GF_THREADS) {
      float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;           // four fixed interleaved chains (more loads in flight)
      int b = 0;
      for (; b + 3 < p.S; b += 4) {
        t0 += __ldcg(src + static_cast<long long>(b) * 2 * C + i);
        t1 += __ldcg(src + static_cast<long long>(b + 1) * 2 * C + i);
        t2 += __ldcg(src + static_cast<long long>(b + 2) * 2 * C + i);
        t3 += __ldcg(src + static_cast<long long>(b + 3) * 2 * C + i);
      }
      for (; b < p.S; ++b) t0 += __ldcg(src + static_cast<long long>(b) * 2 * C + i);
      __stcg(p.team_part + static_cast<long long>(team) * 2 * C + i, (t0 + t1) + (t2 + t3));
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(p.flags + p.N, 1u) == static_cast<unsigned>(p.T) - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const unsigned poisoned = __ldcg(p.flags + p.N + 1);        // a team barrier timed out: never pass silently
  for (int i = tid; i < 2 * C; i += GF_THREADS) {
    float t = 0.f;
    for (int b = 0; b < p.T; ++b) t += __ldcg(p.team_part + static_cast<long long>(b) * 2 * C + i);
    p.dgb[i] = poisoned ? __int_as_float(0x7fc00000) : t;
  }
}


// =============================================================================================================
// GroupNorm(+swish) FORWARD with the same team scheme: phase 1 streams the team's image once for the per-group sums
// (sum, sum of squares; fp32 per CTA, fp64 across the team's CTAs), team barrier, phase 2 streams it again - from
// L2 - and writes y = act(x * rstd*gamma + beta - mean*rstd*gamma).  HBM sees 1 read + 1 write instead of the
// 2 reads + 1 write of gn_partial_stats + gn_finalize + gn_apply.  stats [N][G][2] = (mean, rstd) is written for
// the backward.  One tensor per ring stage; exact sigmoid (ex2 + rcp) like the separate apply kernel.
constexpr int GW_STAGES = 5;
constexpr int GW_SMEM = GW_STAGES * GF_CHUNK + (2 * GF_CONSUMERS * 8 + 2 * GF_MAXC + 2 * GF_MAXG) * 4 + 2 * GW_STAGES * 8;

struct GnFwdParams {
  const __nv_bfloat16* x;
  const float* gamma;
  const float* beta;
  __nv_bfloat16* y;
  float* stats;                  // [N][G][2] mean, rstd (output)
  float* part;                   // [N][S][G][2] per-CTA group partials
  unsigned* flags;               // [N] arrival counters | [N] error flag (zeroed before launch)
  int N, HW, C, G, S, T, rows_per_cta;
  float eps;
};

template <bool SW>
__global__ void __launch_bounds__(GF_THREADS, 2) gn_fwd_fused_kernel(const GnFwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* ring = smem;
  float* red = reinterpret_cast<float*>(smem + GW_STAGES * GF_CHUNK);
  float* chan = red + 2 * GF_CONSUMERS * 8;
  float* sk = chan + 2 * GF_MAXC;               // [G][2] mean, rstd of the current image
  uint64_t* bars = reinterpret_cast<uint64_t*>(sk + 2 * GF_MAXG);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + GW_STAGES);
  const int tid = threadIdx.x;
  const int C = p.C, G = p.G, HW = p.HW;
  const int vecs = C >> 3;
  const int rstep = GF_CONSUMERS / vecs;
  const int srows = rstep * GF_VPT;
  const int team = blockIdx.x / p.S, s_idx = blockIdx.x % p.S;
  const int r0 = s_idx * p.rows_per_cta;
  const int r1 = min(HW, r0 + p.rows_per_cta);
  const int nstages = (r1 - r0 + srows - 1) / srows;

  if (tid == 0) {
    for (int i = 0; i < GW_STAGES; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, GF_CONSUMERS / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (tid >= GF_CONSUMERS) {
    if (tid == GF_CONSUMERS) {                    // producer lane
      uint32_t stage = 0, phase = 0;
      const uint64_t keep = l2_policy_evict_last(), drop = l2_policy_evict_first();
      for (int n = team; n < p.N; n += p.T) {
        const long long img = static_cast<long long>(n) * HW * C;
        for (int ph = 0; ph < 2; ++ph) {
          for (int c = 0; c < nstages; ++c) {
            const int row = r0 + c * srows;
            const uint32_t bytes = static_cast<uint32_t>(min(srows, r1 - row)) * C * 2;
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            const uint32_t fb = full0 + 8 * stage;
            mbar_arrive_expect_tx(fb, bytes);
            bulk_load(smem_u32(ring + stage * GF_CHUNK), p.x + img + static_cast<long long>(row) * C, bytes, fb,
                      ph == 0 ? keep : drop);
            if (++stage == GW_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    return;
  }
  // ------------------------------------------------------------------ consumers
  const int v = tid % vecs, rlane = tid / vecs;
  const int cg = C / G;
  const int lane = tid & 31;
  uint32_t stage = 0, phase = 0;
  const double inv_cnt = 1.0 / (static_cast<double>(HW) * cg);
  float gm[8], bt[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { gm[k] = p.gamma[v * 8 + k]; bt[k] = p.beta[v * 8 + k]; }
  for (int n = team; n < p.N; n += p.T) {
    // ---------------- phase 1: per-channel sum and sum of squares over this CTA's rows
    float2 sa[4], sb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { sa[k] = make_float2(0.f, 0.f); sb[k] = make_float2(0.f, 0.f); }
    for (int c = 0; c < nstages; ++c) {
      mbar_wait(full0 + 8 * stage, phase);
      const uint8_t* st = ring + stage * GF_CHUNK + tid * 16;
#pragma unroll
      for (int u = 0; u < GF_VPT; ++u) {
        if (r0 + c * srows + u * rstep + rlane < r1) {
          const uint4 ux = *reinterpret_cast<const uint4*>(st + u * (GF_CONSUMERS * 16));
          const uint32_t wx[4] = {ux.x, ux.y, ux.z, ux.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 xv = unpack2(wx[k]);
            sa[k] = __fadd2_rn(sa[k], xv);
            sb[k] = __ffma2_rn(xv, xv, sb[k]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * stage);
      if (++stage == GW_STAGES) { stage = 0; phase ^= 1; }
    }
    float* red_a = red;
    float* red_b = red + GF_CONSUMERS * 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      *reinterpret_cast<float2*>(red_a + rlane * C + v * 8 + 2 * k) = sa[k];
      *reinterpret_cast<float2*>(red_b + rlane * C + v * 8 + 2 * k) = sb[k];
    }
    consumer_sync();
    for (int o = tid; o < 2 * C; o += GF_CONSUMERS) {
      const float* src = (o < C) ? red_a + o : red_b + (o - C);
      float t = 0.f;
#pragma unroll 4
      for (int rl = 0; rl < rstep; ++rl) t += src[rl * C];
      chan[(o < C) ? o : GF_MAXC + (o - C)] = t;
    }
    consumer_sync();
    if (tid < G) {
      float S1 = 0.f, S2 = 0.f;
      for (int cc = tid * cg; cc < (tid + 1) * cg; ++cc) { S1 += chan[cc]; S2 += chan[GF_MAXC + cc]; }
      float* o = p.part + ((static_cast<long long>(n) * p.S + s_idx) * G + tid) * 2;
      __stcg(reinterpret_cast<float2*>(o), make_float2(S1, S2));
    }
    consumer_sync();
    // ---------------- team barrier on image n
    if (tid == 0) {
      __threadfence();
      red_release_add(p.flags + n, 1u);
      const long long t0 = clock64();
      while (ld_acquire(p.flags + n) < static_cast<unsigned>(p.S)) {
        if (clock64() - t0 > 4000000000LL) { atomicExch(p.flags + p.N, 1u); break; }
      }
    }
    consumer_sync();
    {
      // fixed-order sum of the S partials per (group, 2): lanes over CTAs in fp64, then over the 16 lanes in order
      const int e4n = (2 * G) / 4;
      const int lanes_s = GF_CONSUMERS / e4n;
      const int e4 = tid % e4n, j = tid / e4n;
      const float4* src = reinterpret_cast<const float4*>(p.part + static_cast<long long>(n) * p.S * G * 2) + e4;
      constexpr int MAXL = (GF_MAXS + 15) / 16;
      float4 t[MAXL];
#pragma unroll
      for (int i = 0; i < MAXL; ++i) {
        const int s = j + i * lanes_s;
        t[i] = (s < p.S) ? __ldcg(src + static_cast<long long>(s) * e4n) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
      for (int i = 0; i < MAXL; ++i) { a0 += t[i].x; a1 += t[i].y; a2 += t[i].z; a3 += t[i].w; }
      double* rd = reinterpret_cast<double*>(red);           // [lanes_s][2G] doubles (16 x 64 x 8 B = 8 KB)
      rd[j * 2 * G + e4 * 4 + 0] = a0; rd[j * 2 * G + e4 * 4 + 1] = a1;
      rd[j * 2 * G + e4 * 4 + 2] = a2; rd[j * 2 * G + e4 * 4 + 3] = a3;
      consumer_sync();
      if (tid < G) {
        double S1 = 0.0, S2 = 0.0;
        for (int jj = 0; jj < lanes_s; ++jj) { S1 += rd[jj * 2 * G + 2 * tid]; S2 += rd[jj * 2 * G + 2 * tid + 1]; }
        const double mean = S1 * inv_cnt;
        double var = S2 * inv_cnt - mean * mean;
        if (var < 0) var = 0;
        const float m = static_cast<float>(mean), r = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps)));
        sk[2 * tid] = m;
        sk[2 * tid + 1] = r;
        if (s_idx == 0) {
          p.stats[(static_cast<long long>(n) * G + tid) * 2] = m;
          p.stats[(static_cast<long long>(n) * G + tid) * 2 + 1] = r;
        }
      }
      consumer_sync();
    }
    // ---------------- phase 2: y = act(x * a + b)
    float2 A[4], B[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int g0 = (v * 8 + 2 * k) / cg;                   // C/G even: the pair lies in one group
      const float m = sk[2 * g0], r = sk[2 * g0 + 1];
      A[k] = make_float2(r * gm[2 * k], r * gm[2 * k + 1]);
      B[k] = make_float2(bt[2 * k] - m * A[k].x, bt[2 * k + 1] - m * A[k].y);
    }
    __nv_bfloat16* out = p.y + static_cast<long long>(n) * HW * C + v * 8;
    for (int c = 0; c < nstages; ++c) {
      mbar_wait(full0 + 8 * stage, phase);
      const uint8_t* st = ring + stage * GF_CHUNK + tid * 16;
#pragma unroll
      for (int u = 0; u < GF_VPT; ++u) {
        const int row = r0 + c * srows + u * rstep + rlane;
        if (row < r1) {
          const uint4 ux = *reinterpret_cast<const uint4*>(st + u * (GF_CONSUMERS * 16));
          const uint32_t wx[4] = {ux.x, ux.y, ux.z, ux.w};
          uint32_t o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float2 z = __ffma2_rn(unpack2(wx[k]), A[k], B[k]);
            if (SW) {
              z.x *= __fdividef(1.f, 1.f + __expf(-z.x));
              z.y *= __fdividef(1.f, 1.f + __expf(-z.y));
            }
            o[k] = pack_bf16x2(z.x, z.y);
          }
          *reinterpret_cast<uint4*>(out + static_cast<long long>(row) * C) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * stage);
      if (++stage == GW_STAGES) { stage = 0; phase ^= 1; }
    }
  }
}

}  // namespace b2

using namespace b2;

// Images whose dy + x may be in flight at once (teams): their bytes must stay L2-resident between the two phases.
static long long gf_l2_budget() {
  static long long v = -1;
  if (v < 0) {
    const char* e = getenv("B2DQ_GN_L2_BUDGET_MB");
    v = (e ? atoll(e) : 70) << 20;
  }
  return v;
}

struct GfPlan { int T, S, rows_per_cta, grid; };
static GfPlan gf_plan(int N, int HW, int C, int tensors = 2) {
  const int sms = 2 * device_sm_count();                        // 2 CTAs per SM
  const int rstep = GF_VPT * (GF_CONSUMERS / (C / 8));          // rows per ring stage
  const int chunks_img = (HW + rstep - 1) / rstep;
  const long long per_img = 2LL * tensors * HW * C;              // bf16 tensors re-read from L2 (dy + x | x)
  long long T = gf_l2_budget() / (per_img > 0 ? per_img : 1);
  if (T < 1) T = 1;
  if (T > N) T = N;
  if (T > sms) T = sms;
  for (long long t = T; 2 * t > T; --t)                         // same number of images per team when a nearby T allows it
    if (N % t == 0) { T = t; break; }
  int S = sms / static_cast<int>(T);
  if (S > GF_MAXS) S = GF_MAXS;
  if (S > chunks_img) S = chunks_img;
  if (S < 1) S = 1;
  const int chunks_cta = (chunks_img + S - 1) / S;
  GfPlan pl;
  pl.rows_per_cta = chunks_cta * rstep;
  pl.S = (HW + pl.rows_per_cta - 1) / pl.rows_per_cta;          // CTAs with a non-empty slice
  pl.T = static_cast<int>(T);
  pl.grid = pl.T * pl.S;
  return pl;
}

extern "C" {

// Scratch the fused backward needs for an [N, HW, C] tensor with G groups, in bytes (0: shape not supported,
// use b2dq_gn_bwd_stats + b2dq_gn_bwd_apply).
int b2dq_gn_bwd_fused_workspace_bytes(int N, int HW, int C, int G) {
  if (N <= 0 || HW <= 0) return 0;
  if (C % 8 || C > GF_MAXC || G > GF_MAXG || G <= 0 || C % G || (C / G) % 2 || GF_CONSUMERS % (C / 8) || (2 * G) % 4 ||
      GF_CONSUMERS % ((2 * G) / 4) || (GF_CONSUMERS / ((2 * G) / 4)) * ((GF_MAXS + 15) / 16) < GF_MAXS)
    return 0;
  const GfPlan pl = gf_plan(N, HW, C);
  const long long part = 1LL * N * pl.S * G * 2 * 4;
  const long long dgbp = 1LL * (pl.grid + pl.T) * 2 * C * 4;   // per-CTA partials, then per-team sums
  const long long flags = (1LL * N + 2 + pl.T) * 4;
  return static_cast<int>(((part + 15) / 16 + (dgbp + 15) / 16 + (flags + 15) / 16) * 16);
}

// dx = d/dx [ swish?(GroupNorm(x)) ] . dy (+ add);  dgb [2*C] = (dgamma, dbeta).  ws: b2dq_gn_bwd_fused_workspace_bytes.
// Returns 0, a cudaError_t, -1 (unsupported shape) or -2 (workspace too small).
int b2dq_gn_bwd_fused(const void* dy, const void* x, const float* stats, const float* gamma, const float* beta,
                      void* dx, float* dgb, const void* add, void* ws, long long ws_bytes, int N, int HW, int C, int G,
                      int swish, cudaStream_t stream) {
  if (N <= 0 || HW <= 0) return 0;
  const long long need = b2dq_gn_bwd_fused_workspace_bytes(N, HW, C, G);
  if (need == 0 || (swish != 0 && swish != 1)) return -1;
  if (ws_bytes < need || ws == nullptr) return -2;
  const GfPlan pl = gf_plan(N, HW, C);
  GnFusedParams p;
  p.dy = reinterpret_cast<const __nv_bfloat16*>(dy);
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.add = reinterpret_cast<const __nv_bfloat16*>(add);
  p.stats = stats; p.gamma = gamma; p.beta = beta;
  p.dx = reinterpret_cast<__nv_bfloat16*>(dx);
  p.dgb = dgb;
  uint8_t* w = reinterpret_cast<uint8_t*>(ws);
  const long long part = ((1LL * N * pl.S * G * 2 * 4 + 15) / 16) * 16;
  const long long dgbp = ((1LL * (pl.grid + pl.T) * 2 * C * 4 + 15) / 16) * 16;
  p.part = reinterpret_cast<float*>(w);
  p.dgb_part = reinterpret_cast<float*>(w + part);
  p.team_part = p.dgb_part + 1LL * pl.grid * 2 * C;
  p.flags = reinterpret_cast<unsigned*>(w + part + dgbp);
  p.N = N; p.HW = HW; p.C = C; p.G = G; p.S = pl.S; p.T = pl.T; p.rows_per_cta = pl.rows_per_cta;
  cudaError_t e = cudaMemsetAsync(p.flags, 0, (N + 2 + pl.T) * sizeof(unsigned), stream);
  if (e != cudaSuccess) return (int)e;
  static unsigned long long m0 = 0, m1 = 0;
  if (swish) {
    if (int r = set_max_smem_once(gn_bwd_fused_kernel<true>, GF_SMEM, m1)) return r;
    gn_bwd_fused_kernel<true><<<pl.grid, GF_THREADS, GF_SMEM, stream>>>(p);
  } else {
    if (int r = set_max_smem_once(gn_bwd_fused_kernel<false>, GF_SMEM, m0)) return r;
    gn_bwd_fused_kernel<false><<<pl.grid, GF_THREADS, GF_SMEM, stream>>>(p);
  }
  return (int)cudaGetLastError();
}

// ---- forward: y = act(GroupNorm(x)), stats [N][G][2] = (mean, rstd) written for the backward
int b2dq_gn_fwd_fused_workspace_bytes(int N, int HW, int C, int G) {
  if (b2dq_gn_bwd_fused_workspace_bytes(N, HW, C, G) == 0) return 0;     // same shape constraints
  const GfPlan pl = gf_plan(N, HW, C, 1);
  const long long part = ((1LL * N * pl.S * G * 2 * 4 + 15) / 16) * 16;
  return static_cast<int>(part + ((1LL * N + 1) * 4 + 15) / 16 * 16);
}

int b2dq_gn_fwd_fused(const void* x, const float* gamma, const float* beta, void* y, float* stats, void* ws,
                      long long ws_bytes, int N, int HW, int C, int G, float eps, int swish, cudaStream_t stream) {
  if (N <= 0 || HW <= 0) return 0;
  const long long need = b2dq_gn_fwd_fused_workspace_bytes(N, HW, C, G);
  if (need == 0 || (swish != 0 && swish != 1)) return -1;
  if (ws_bytes < need || ws == nullptr) return -2;
  const GfPlan pl = gf_plan(N, HW, C, 1);
  GnFwdParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.gamma = gamma; p.beta = beta;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.stats = stats;
  uint8_t* w = reinterpret_cast<uint8_t*>(ws);
  const long long part = ((1LL * N * pl.S * G * 2 * 4 + 15) / 16) * 16;
  p.part = reinterpret_cast<float*>(w);
  p.flags = reinterpret_cast<unsigned*>(w + part);
  p.N = N; p.HW = HW; p.C = C; p.G = G; p.S = pl.S; p.T = pl.T; p.rows_per_cta = pl.rows_per_cta; p.eps = eps;
  cudaError_t e = cudaMemsetAsync(p.flags, 0, (N + 1) * sizeof(unsigned), stream);
  if (e != cudaSuccess) return (int)e;
  static unsigned long long m0 = 0, m1 = 0;
  if (swish) {
    if (int r = set_max_smem_once(gn_fwd_fused_kernel<true>, GW_SMEM, m1)) return r;
    gn_fwd_fused_kernel<true><<<pl.grid, GF_THREADS, GW_SMEM, stream>>>(p);
  } else {
    if (int r = set_max_smem_once(gn_fwd_fused_kernel<false>, GW_SMEM, m0)) return r;
    gn_fwd_fused_kernel<false><<<pl.grid, GF_THREADS, GW_SMEM, stream>>>(p);
  }
  return (int)cudaGetLastError();
}

// The plan of the call above (tests / bench): teams, CTAs per team, rows per CTA, grid.
int b2dq_gn_bwd_fused_plan(int N, int HW, int C, int* out4) {
  if (N <= 0 || HW <= 0 || C < 8 || !out4) return -1;
  const GfPlan pl = gf_plan(N, HW, C);
  out4[0] = pl.T; out4[1] = pl.S; out4[2] = pl.rows_per_cta; out4[3] = pl.grid;
  return 0;
}

}  // extern "C"
