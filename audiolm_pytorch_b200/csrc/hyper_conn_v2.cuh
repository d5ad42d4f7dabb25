// Hyper-Connections kernels, second generation: 2 warps per token, 4 tokens per CTA, no CTA-wide barrier.
// (The first generation in hyper_conn.cu used one CTA per token with 5-6 __syncthreads per token and ran
// 8-14x off the HBM roofline: profiles/r01_launches_bench_step_v1.csv.)
//
//  - a token's 64 threads each own NCH chunks of 8 channels (16-B vector loads, fully coalesced);
//  - reductions: warp shuffle + one 64-thread named barrier (bar.sync id, 64) through a tiny smem mailbox;
//  - per-channel parameters live in shared memory (fp32); parameter gradients are accumulated in shared
//    memory with red.shared and flushed once per CTA with global atomics.
#pragma once
#include "alm_common.cuh"

namespace alm {
namespace hc2 {

constexpr int S = 4, T = 5;
constexpr int THREADS = 256;      // a CTA holds THREADS / TPT token slots; TPT = threads per token (64 or 128)
constexpr int MAILW = 32;        // floats per warp row of the reduction mailbox (largest reduction: 28 values)
constexpr int AUX = S * T + S + S + (S * T + S) + 2;  // ta[20] tb[4] inv[4] z[24] (pre-tanh) mean rstd

template <int TPT>
__device__ __forceinline__ void bar_slot(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(TPT) : "memory");
}

// sum N values over the TPT threads of a token slot; all of them get the result.
template <int N, int TPT>
__device__ __forceinline__ void slot_sum(float (&v)[N], float* mail /*[2][TPT/32][MAILW]*/, int& which, int w2, int lane,
                                         int bar_id) {
  constexpr int WPT = TPT / 32;
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
  float* b = mail + which * (WPT * MAILW);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) b[w2 * MAILW + i] = v[i];
  }
  bar_slot<TPT>(bar_id);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float a = b[i];
#pragma unroll
    for (int w = 1; w < WPT; ++w) a += b[w * MAILW + i];
    v[i] = a;
  }
  which ^= 1;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint32_t pk(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pk(f[0], f[1]), pk(f[2], f[3]), pk(f[4], f[5]), pk(f[6], f[7]));
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ void lds4(const float* p, float* f) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));  // MUFU.TANH, rel. err ~2^-11: far below the bf16 noise floor
  return y;
}

struct Params {
  const float* gamma_hc; const float* dyn_alpha; const float* dyn_beta; const float* static_alpha;
  const float* static_beta; const float* alpha_scale; const float* beta_scale; const float* ln_gamma;
};
struct Grads {
  float* gamma_hc; float* dyn_alpha; float* dyn_beta; float* static_alpha; float* static_beta;
  float* alpha_scale; float* beta_scale; float* ln_gamma;
};

// smem: [0,d) g1 = (gamma+1)*sqrt(d); [d,2d) dyn_beta; [2d,3d) ln_gamma; [3d, 8d) dyn_alpha transposed [T][d]
__device__ __forceinline__ void stage_params(float* sm, const Params& p, int d) {
  const float sqrt_d = sqrtf((float)d);
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    sm[i] = (p.gamma_hc[i] + 1.f) * sqrt_d;
    sm[d + i] = p.dyn_beta[i];
    sm[2 * d + i] = p.ln_gamma[i];
#pragma unroll
    for (int t = 0; t < T; ++t) sm[(3 + t) * d + i] = p.dyn_alpha[(size_t)i * T + t];
  }
}

// ------------------------------------------------------------------------------------------------
template <int NCH, int TPT>
__global__ void __launch_bounds__(THREADS, 2)
pre_fwd_kernel(const __nv_bfloat16* __restrict__ R_in, const __nv_bfloat16* __restrict__ Y,
               const float* __restrict__ beta_prev, const float* __restrict__ x_expand, Params prm,
               __nv_bfloat16* __restrict__ R_out, __nv_bfloat16* __restrict__ bin, __nv_bfloat16* __restrict__ xn,
               float* __restrict__ beta_out, float* __restrict__ aux, int M, int d) {
  extern __shared__ float sm[];
  constexpr int TOK = THREADS / TPT, WPT = TPT / 32;
  float* mailbox = sm + 8 * d;  // [TOK][2][WPT][MAILW]
  stage_params(sm, prm, d);
  __syncthreads();
  const float* sG1 = sm;
  const float* sBf = sm + d;
  const float* sLn = sm + 2 * d;
  const float* sA = sm + 3 * d;
  const int slot = threadIdx.x / TPT, lt = threadIdx.x % TPT, w2 = lt >> 5, lane = lt & 31;
  float* mail = mailbox + slot * (2 * WPT * MAILW);
  int which = 0;
  const int bar_id = 1 + slot;
  const float a_scale = *prm.alpha_scale, b_scale = *prm.beta_scale;
  float Astat[S][T], Bstat[S];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    Bstat[s] = prm.static_beta[s];
#pragma unroll
    for (int t = 0; t < T; ++t) Astat[s][t] = prm.static_alpha[s * T + t];
  }
  int ch[NCH];
  bool act[NCH];
#pragma unroll
  for (int k = 0; k < NCH; ++k) { ch[k] = (lt + TPT * k) * 8; act[k] = ch[k] < d; }

  for (int m = blockIdx.x * TOK + slot; m < M; m += gridDim.x * TOK) {
    {  // pull the next token of this slot towards L2 while this one is processed (one lane per 128-B line)
      const int mn = m + gridDim.x * TOK;
      if (mn < M && (lt & 7) == 0) {
#pragma unroll
        for (int k = 0; k < NCH; ++k)
          if (act[k]) {
            if (x_expand != nullptr) {
              prefetch_l2(x_expand + (size_t)mn * d + ch[k]);
              prefetch_l2(x_expand + (size_t)mn * d + ch[k] + 32);
            } else {
              prefetch_l2(Y + (size_t)mn * d + ch[k]);
#pragma unroll
              for (int s = 0; s < S; ++s) prefetch_l2(R_in + ((size_t)mn * S + s) * d + ch[k]);
            }
          }
      }
    }
    float R[S][NCH][8];
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      if (act[k]) {
        if (x_expand != nullptr) {
          const float4 a = *reinterpret_cast<const float4*>(x_expand + (size_t)m * d + ch[k]);
          const float4 b = *reinterpret_cast<const float4*>(x_expand + (size_t)m * d + ch[k] + 4);
          const float xv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
          for (int s = 0; s < S; ++s)
#pragma unroll
            for (int e = 0; e < 8; ++e) R[s][k][e] = xv[e];
        } else {
          float yv[8];
          unpack8(*reinterpret_cast<const uint4*>(Y + (size_t)m * d + ch[k]), yv);
#pragma unroll
          for (int s = 0; s < S; ++s) {
            float rv[8];
            unpack8(*reinterpret_cast<const uint4*>(R_in + ((size_t)m * S + s) * d + ch[k]), rv);
            const float bp = beta_prev[(size_t)m * S + s];
#pragma unroll
            for (int e = 0; e < 8; ++e) R[s][k][e] = fmaf(bp, yv[e], rv[e]);
          }
        }
      } else {
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
          for (int e = 0; e < 8; ++e) R[s][k][e] = 0.f;
      }
    }
    // ONE reduction for the stream norms and every dynamic-map dot product: the dots are taken on the raw residual
    // (z[s][c] = inv_s * sum_d R_s[d] g1[d] P_c[d]) so they do not have to wait for inv_s
    float w[S * T + S + S];  // [0, S*T+S): raw dots, then S sums of squares
#pragma unroll
    for (int i = 0; i < S * T + S + S; ++i) w[i] = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      if (act[k]) {
#pragma unroll
        for (int h4 = 0; h4 < 2; ++h4) {  // 4 channels at a time: 16-B shared loads, no bank-conflict replays
          const int c = ch[k] + h4 * 4;
          float g1[4], bf[4], av[T][4];
          lds4(sG1 + c, g1);
          lds4(sBf + c, bf);
#pragma unroll
          for (int t = 0; t < T; ++t) lds4(sA + t * d + c, av[t]);
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const float rv = R[s][k][h4 * 4 + e];
              const float nv = rv * g1[e];
              w[S * T + S + s] = fmaf(rv, rv, w[S * T + S + s]);
#pragma unroll
              for (int t = 0; t < T; ++t) w[s * T + t] = fmaf(nv, av[t][e], w[s * T + t]);
              w[S * T + s] = fmaf(nv, bf[e], w[S * T + s]);
            }
        }
      }
    }
    slot_sum<S * T + S + S, TPT>(w, mail, which, w2, lane, bar_id);
    float inv[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      inv[s] = 1.f / fmaxf(sqrtf(w[S * T + S + s]), 1e-12f);
#pragma unroll
      for (int t = 0; t < T; ++t) w[s * T + t] *= inv[s];
      w[S * T + s] *= inv[s];
    }
    if (lt == 0) {  // pre-activations: the backward's RMS-norm term needs them (hyper_conn_v3.cuh)
      float* az = aux + (size_t)m * AUX + S * T + S + S;
#pragma unroll
      for (int i = 0; i < S * T + S; ++i) az[i] = w[i];
    }
    float alpha[S][T], beta[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
#pragma unroll
      for (int t = 0; t < T; ++t) {
        w[s * T + t] = tanh_fast(w[s * T + t]);
        alpha[s][t] = fmaf(w[s * T + t], a_scale, Astat[s][t]);
      }
      w[S * T + s] = tanh_fast(w[S * T + s]);
      beta[s] = fmaf(w[S * T + s], b_scale, Bstat[s]);
    }
    // mixed residual streams out; branch input kept for the LayerNorm
    float bi[NCH][8];
    float st[2] = {0.f, 0.f};  // sum, sum of squares of the branch input (LayerNorm statistics in one reduction)
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float a = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) a = fmaf(alpha[s][0], R[s][k][e], a);
        bi[k][e] = a;
        st[0] += a;
        st[1] = fmaf(a, a, st[1]);
      }
      if (act[k]) {
#pragma unroll
        for (int t = 1; t < T; ++t) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float a = 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) a = fmaf(alpha[s][t], R[s][k][e], a);
            o[e] = a;
          }
          *reinterpret_cast<uint4*>(R_out + ((size_t)m * S + (t - 1)) * d + ch[k]) = pack8(o);
        }
        *reinterpret_cast<uint4*>(bin + (size_t)m * d + ch[k]) = pack8(bi[k]);
      }
    }
    slot_sum<2, TPT>(st, mail, which, w2, lane, bar_id);
    const float mean = st[0] / d;
    const float rstd = rsqrtf(fmaxf(st[1] / d - mean * mean, 0.f) + 1e-5f);
#pragma unroll
    for (int k = 0; k < NCH; ++k)
      if (act[k]) {
        float o[8], lg[8];
        lds4(sLn + ch[k], lg);
        lds4(sLn + ch[k] + 4, lg + 4);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (bi[k][e] - mean) * rstd * lg[e];
        *reinterpret_cast<uint4*>(xn + (size_t)m * d + ch[k]) = pack8(o);
      }
    if (lt == 0) {
      float* a = aux + (size_t)m * AUX;
#pragma unroll
      for (int i = 0; i < S * T + S; ++i) a[i] = w[i];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        a[S * T + S + s] = inv[s];
        beta_out[(size_t)m * S + s] = beta[s];
      }
      a[AUX - 2] = mean;
      a[AUX - 1] = rstd;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// smem for the backward adds PRIVATE per-slot gradient accumulators [TOK][8][d]: gG, gBf, gLn, gA[T]
template <int NCH, int TPT>
__global__ void __launch_bounds__(THREADS, TPT == 128 ? 2 : 1)
pre_bwd_kernel(const __nv_bfloat16* __restrict__ R_in, const __nv_bfloat16* __restrict__ Y,
               const float* __restrict__ beta_prev, const float* __restrict__ x_expand, Params prm,
               const float* __restrict__ aux, const __nv_bfloat16* __restrict__ dR_out,
               const __nv_bfloat16* __restrict__ dxn, const __nv_bfloat16* __restrict__ dbin_extra,
               const float* __restrict__ dbeta, __nv_bfloat16* __restrict__ dR_in, __nv_bfloat16* __restrict__ dY,
               float* __restrict__ dbeta_prev, float* __restrict__ dx_expand, float dx_scale, Grads gr, int M,
               int d) {
  extern __shared__ float sm[];
  constexpr int TOK = THREADS / TPT, WPT = TPT / 32;
  float* sGradAll = sm + 8 * d;              // [TOK][8][d]: private per token slot -> plain RMW, no atomics
  float* mailbox = sm + (8 + 8 * TOK) * d;   // [TOK][2][WPT][MAILW]
  stage_params(sm, prm, d);
  for (int i = threadIdx.x; i < 8 * TOK * d; i += blockDim.x) sGradAll[i] = 0.f;
  __syncthreads();
  const float* sG1 = sm;
  const float* sBf = sm + d;
  const float* sLn = sm + 2 * d;
  const float* sA = sm + 3 * d;
  const int slot = threadIdx.x / TPT, lt = threadIdx.x % TPT, w2 = lt >> 5, lane = lt & 31;
  float* sGrad = sGradAll + (size_t)slot * 8 * d;
  float* gG = sGrad;
  float* gBf = sGrad + d;
  float* gLn = sGrad + 2 * d;
  float* gA = sGrad + 3 * d;
  float* mail = mailbox + slot * (2 * WPT * MAILW);
  int which = 0;
  const int bar_id = 1 + slot;
  const float sqrt_d = sqrtf((float)d);
  const float a_scale = *prm.alpha_scale, b_scale = *prm.beta_scale;
  float Astat[S][T];
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int t = 0; t < T; ++t) Astat[s][t] = prm.static_alpha[s * T + t];
  float acc_small[S * T + S + 2];
#pragma unroll
  for (int i = 0; i < S * T + S + 2; ++i) acc_small[i] = 0.f;
  int ch[NCH];
  bool act[NCH];
#pragma unroll
  for (int k = 0; k < NCH; ++k) { ch[k] = (lt + TPT * k) * 8; act[k] = ch[k] < d; }

  for (int m = blockIdx.x * TOK + slot; m < M; m += gridDim.x * TOK) {
    {  // L2 prefetch of the next token's rows (the kernel is latency-bound at 8 warps / SM)
      const int mn = m + gridDim.x * TOK;
      if (mn < M && (lt & 7) == 0) {
#pragma unroll
        for (int k = 0; k < NCH; ++k)
          if (act[k]) {
            if (x_expand != nullptr) {
              prefetch_l2(x_expand + (size_t)mn * d + ch[k]);
              prefetch_l2(x_expand + (size_t)mn * d + ch[k] + 32);
            } else {
              prefetch_l2(Y + (size_t)mn * d + ch[k]);
#pragma unroll
              for (int s = 0; s < S; ++s) prefetch_l2(R_in + ((size_t)mn * S + s) * d + ch[k]);
            }
#pragma unroll
            for (int s = 0; s < S; ++s) prefetch_l2(dR_out + ((size_t)mn * S + s) * d + ch[k]);
            prefetch_l2(dxn + (size_t)mn * d + ch[k]);
            if (dbin_extra != nullptr) prefetch_l2(dbin_extra + (size_t)mn * d + ch[k]);
          }
      }
    }
    const float* a = aux + (size_t)m * AUX;
    float ta[S][T], tb[S], inv[S], alpha[S][T], bp[S], dbe[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
#pragma unroll
      for (int t = 0; t < T; ++t) {
        ta[s][t] = a[s * T + t];
        alpha[s][t] = fmaf(ta[s][t], a_scale, Astat[s][t]);
      }
      tb[s] = a[S * T + s];
      inv[s] = a[S * T + S + s];
      dbe[s] = dbeta[(size_t)m * S + s];
      bp[s] = (x_expand == nullptr) ? beta_prev[(size_t)m * S + s] : 0.f;
    }
    const float mean = a[AUX - 2], rstd = a[AUX - 1];

    float R[S][NCH][8], yv[NCH][8], dmix[T][NCH][8];
    float lnred[2] = {0.f, 0.f};
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
#pragma unroll
      for (int e = 0; e < 8; ++e) yv[k][e] = 0.f;
      if (act[k]) {
        if (x_expand != nullptr) {
          const float4 xa = *reinterpret_cast<const float4*>(x_expand + (size_t)m * d + ch[k]);
          const float4 xb = *reinterpret_cast<const float4*>(x_expand + (size_t)m * d + ch[k] + 4);
          const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
          for (int s = 0; s < S; ++s)
#pragma unroll
            for (int e = 0; e < 8; ++e) R[s][k][e] = xv[e];
        } else {
          unpack8(*reinterpret_cast<const uint4*>(Y + (size_t)m * d + ch[k]), yv[k]);
#pragma unroll
          for (int s = 0; s < S; ++s) {
            float rv[8];
            unpack8(*reinterpret_cast<const uint4*>(R_in + ((size_t)m * S + s) * d + ch[k]), rv);
#pragma unroll
            for (int e = 0; e < 8; ++e) R[s][k][e] = fmaf(bp[s], yv[k][e], rv[e]);
          }
        }
#pragma unroll
        for (int s = 0; s < S; ++s)
          unpack8(*reinterpret_cast<const uint4*>(dR_out + ((size_t)m * S + s) * d + ch[k]), dmix[s + 1][k]);
        // LayerNorm backward, part 1 (dmix[0] temporarily holds gl = dxn * ln_gamma)
        float dx8[8], lg[8], gl8[8];
        unpack8(*reinterpret_cast<const uint4*>(dxn + (size_t)m * d + ch[k]), dx8);
        lds4(sLn + ch[k], lg);
        lds4(sLn + ch[k] + 4, lg + 4);
        lds4(gLn + ch[k], gl8);
        lds4(gLn + ch[k] + 4, gl8 + 4);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float b = 0.f;
#pragma unroll
          for (int s = 0; s < S; ++s) b = fmaf(alpha[s][0], R[s][k][e], b);
          const float xhat = (b - mean) * rstd;
          const float gl = dx8[e] * lg[e];
          gl8[e] = fmaf(dx8[e], xhat, gl8[e]);
          lnred[0] += gl;
          lnred[1] = fmaf(gl, xhat, lnred[1]);
          dmix[0][k][e] = gl;
        }
        *reinterpret_cast<float4*>(gLn + ch[k]) = make_float4(gl8[0], gl8[1], gl8[2], gl8[3]);
        *reinterpret_cast<float4*>(gLn + ch[k] + 4) = make_float4(gl8[4], gl8[5], gl8[6], gl8[7]);
      } else {
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
          for (int e = 0; e < 8; ++e) { R[s][k][e] = 0.f; dmix[s + 1][k][e] = 0.f; }
#pragma unroll
        for (int e = 0; e < 8; ++e) dmix[0][k][e] = 0.f;
      }
    }
    slot_sum<2, TPT>(lnred, mail, which, w2, lane, bar_id);
    const float m1 = lnred[0] / d, m2 = lnred[1] / d;
#pragma unroll
    for (int k = 0; k < NCH; ++k)
      if (act[k]) {
        float ex[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) ex[e] = 0.f;
        if (dbin_extra != nullptr) unpack8(*reinterpret_cast<const uint4*>(dbin_extra + (size_t)m * d + ch[k]), ex);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float b = 0.f;
#pragma unroll
          for (int s = 0; s < S; ++s) b = fmaf(alpha[s][0], R[s][k][e], b);
          const float xhat = (b - mean) * rstd;
          dmix[0][k][e] = rstd * (dmix[0][k][e] - m1 - xhat * m2) + ex[e];
        }
      }
    // d alpha
    float dal[S * T];
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
      for (int t = 0; t < T; ++t) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < NCH; ++k)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc = fmaf(dmix[t][k][e], R[s][k][e], acc);
        dal[s * T + t] = acc;
      }
    slot_sum<S * T, TPT>(dal, mail, which, w2, lane, bar_id);
    float dwa[S][T], dwb[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const float g = dal[s * T + t];
        dwa[s][t] = g * a_scale * (1.f - ta[s][t] * ta[s][t]);
        acc_small[s * T + t] += g;
        acc_small[S * T + S] = fmaf(g, ta[s][t], acc_small[S * T + S]);
      }
      dwb[s] = dbe[s] * b_scale * (1.f - tb[s] * tb[s]);
      acc_small[S * T + s] += dbe[s];
      acc_small[S * T + S + 1] = fmaf(dbe[s], tb[s], acc_small[S * T + S + 1]);
    }
    // dR (written in place over dmix[0..3]) + parameter-gradient contributions
    float udot[S] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      if (act[k]) {
#pragma unroll
        for (int h4 = 0; h4 < 2; ++h4) {
          const int c = ch[k] + h4 * 4;
          float g1[4], bf[4], av[T][4], pG[4], pBf[4], pA[T][4];
          lds4(sG1 + c, g1);
          lds4(sBf + c, bf);
          lds4(gG + c, pG);
          lds4(gBf + c, pBf);
#pragma unroll
          for (int t = 0; t < T; ++t) { lds4(sA + t * d + c, av[t]); lds4(gA + t * d + c, pA[t]); }
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const int e = h4 * 4 + e4;
            float dm[T];
#pragma unroll
            for (int t = 0; t < T; ++t) dm[t] = dmix[t][k][e];
#pragma unroll
            for (int s = 0; s < S; ++s) {
              float acc = 0.f;
#pragma unroll
              for (int t = 0; t < T; ++t) acc = fmaf(alpha[s][t], dm[t], acc);
              float dn = dwb[s] * bf[e4];
#pragma unroll
              for (int t = 0; t < T; ++t) dn = fmaf(dwa[s][t], av[t][e4], dn);
              const float rn = R[s][k][e] * inv[s];
              const float nv = rn * g1[e4];
              pG[e4] = fmaf(dn * rn, sqrt_d, pG[e4]);
              pBf[e4] = fmaf(nv, dwb[s], pBf[e4]);
#pragma unroll
              for (int t = 0; t < T; ++t) pA[t][e4] = fmaf(nv, dwa[s][t], pA[t][e4]);
              const float u = dn * g1[e4];
              udot[s] = fmaf(u, R[s][k][e], udot[s]);
              dmix[s][k][e] = fmaf(u, inv[s], acc);  // dR[s] (dm[] was read above)
            }
          }
          *reinterpret_cast<float4*>(gG + c) = make_float4(pG[0], pG[1], pG[2], pG[3]);
          *reinterpret_cast<float4*>(gBf + c) = make_float4(pBf[0], pBf[1], pBf[2], pBf[3]);
#pragma unroll
          for (int t = 0; t < T; ++t)
            *reinterpret_cast<float4*>(gA + t * d + c) = make_float4(pA[t][0], pA[t][1], pA[t][2], pA[t][3]);
        }
      }
    }
    slot_sum<S, TPT>(udot, mail, which, w2, lane, bar_id);
    float dbp[S] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const float kk = udot[s] * inv[s] * inv[s] * inv[s];
#pragma unroll
      for (int k = 0; k < NCH; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          dmix[s][k][e] = fmaf(-R[s][k][e], kk, dmix[s][k][e]);
          dbp[s] = fmaf(dmix[s][k][e], yv[k][e], dbp[s]);
        }
    }
    if (x_expand != nullptr) {
#pragma unroll
      for (int k = 0; k < NCH; ++k)
        if (act[k]) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            o[e] = (dmix[0][k][e] + dmix[1][k][e] + dmix[2][k][e] + dmix[3][k][e]) * dx_scale;
          float* dst = dx_expand + (size_t)m * d + ch[k];
          *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
    } else {
      slot_sum<S, TPT>(dbp, mail, which, w2, lane, bar_id);
#pragma unroll
      for (int k = 0; k < NCH; ++k)
        if (act[k]) {
          float dy[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float acc = 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) acc = fmaf(bp[s], dmix[s][k][e], acc);
            dy[e] = acc;
          }
          *reinterpret_cast<uint4*>(dY + (size_t)m * d + ch[k]) = pack8(dy);
#pragma unroll
          for (int s = 0; s < S; ++s)
            *reinterpret_cast<uint4*>(dR_in + ((size_t)m * S + s) * d + ch[k]) = pack8(dmix[s][k]);
        }
      if (lt == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) dbeta_prev[(size_t)m * S + s] = dbp[s];
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    float acc[8];
#pragma unroll
    for (int a8 = 0; a8 < 8; ++a8) {
      acc[a8] = 0.f;
#pragma unroll
      for (int sl = 0; sl < TOK; ++sl) acc[a8] += sGradAll[((size_t)sl * 8 + a8) * d + i];
    }
    atomicAdd(gr.gamma_hc + i, acc[0]);
    atomicAdd(gr.dyn_beta + i, acc[1]);
    atomicAdd(gr.ln_gamma + i, acc[2]);
#pragma unroll
    for (int t = 0; t < T; ++t) atomicAdd(gr.dyn_alpha + (size_t)i * T + t, acc[3 + t]);
  }
  if (lt == 0) {  // one thread per token slot holds that slot's scalar-parameter partial sums
#pragma unroll
    for (int i = 0; i < S * T; ++i) atomicAdd(gr.static_alpha + i, acc_small[i]);
#pragma unroll
    for (int s = 0; s < S; ++s) atomicAdd(gr.static_beta + s, acc_small[S * T + s]);
    atomicAdd(gr.alpha_scale, acc_small[S * T + S]);
    atomicAdd(gr.beta_scale, acc_small[S * T + S + 1]);
  }
}

inline size_t fwd_smem(int d, int tpt) { return (size_t)(8 * d + (THREADS / tpt) * 2 * (tpt / 32) * MAILW) * sizeof(float); }
inline size_t bwd_smem(int d, int tpt) {
  return (size_t)((8 + 8 * (THREADS / tpt)) * d + (THREADS / tpt) * 2 * (tpt / 32) * MAILW) * sizeof(float);
}

}  // namespace hc2
}  // namespace alm
