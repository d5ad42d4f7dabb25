// Register-tiled causal Conv1d for the SoundStream stacks (soundstream.py:332-345, 362-395), fp32 CUDA cores.
//
// The first-generation kernel (causal_conv1d_kernel in codec.cu) gave each thread a 4 x 4 output tile and spent
// 5 shared-memory loads per 16 FMAs with run-time kernel size / stride / dilation: 10.8 TFLOP/s (14 % of the
// fp32 FMA peak, profiles/r01_bench_n1_v8.json).  Here
//   * (K, stride, dilation) are template parameters (the 10 shapes SoundStream uses; anything else falls back),
//     so the tap loop is fully unrolled and all smem offsets are immediates;
//   * a thread owns COT output channels x TQ output samples (8 x 8 = 64 accumulators): per (channel, tap) it
//     issues COT/4 broadcast LDS.128 for the weights + TQ conflict-free LDS.32 for the samples and COT*TQ FMAs;
//   * the staged input is split by phase (sample p -> xs[c][p % S][p / S]) so strided convs read consecutive
//     words across a warp as well;
//   * the accumulation order per output is unchanged (input channels ascending, taps ascending, one fp32 FMA
//     chain), so results are bit-identical to the first-generation kernel and the RVQ indices stay bit-exact.
// fp32 on CUDA cores is deliberate: the RVQ code search downstream is compared bit-exactly against the fp32
// oracle; a tf32 / bf16 tensor-core conv would flip near-tie codes.
#pragma once
#include "alm_common.cuh"

namespace alm {
namespace cvt {

constexpr int THREADS = 256, CI = 8;

template <int K, int S, int D, int TQ>
struct Geo {
  static constexpr int T_TILE = 32 * TQ;
  static constexpr int SPAN = (T_TILE - 1) * S + (K - 1) * D + 1;  // padded-signal samples per tile and channel
  static constexpr int LI = (SPAN + S - 1) / S + 1;                 // row length of one phase (+1: bank skew)
  static constexpr int XS = CI * S * LI;                            // floats
};

template <int K, int S, int D, int COT, int TQ>
constexpr size_t smem_bytes() {
  return (size_t)(Geo<K, S, D, TQ>::XS + CI * K * 8 * COT) * sizeof(float);
}

template <int K, int S, int D, int COT, int TQ>
__global__ void __launch_bounds__(THREADS, 2)
conv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
            const float* __restrict__ residual, float* __restrict__ y, int Cin, int Cout, int T, int Tout, int pad,
            int pad_mode, int act, int w_packed) {
  using G = Geo<K, S, D, TQ>;
  constexpr int CO_TILE = 8 * COT;
  extern __shared__ float smem[];
  float* xs = smem;            // [CI][S][LI]
  float* ws = smem + G::XS;    // [CI][K][CO_TILE]
  const int t0 = blockIdx.x * G::T_TILE, o0 = blockIdx.y * CO_TILE, b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int in0 = t0 * S;
  float acc[COT][TQ];
#pragma unroll
  for (int i = 0; i < COT; ++i)
#pragma unroll
    for (int q = 0; q < TQ; ++q) acc[i][q] = 0.f;

  for (int c0 = 0; c0 < Cin; c0 += CI) {
    __syncthreads();
    {  // stage the samples: warp <-> input channel, lanes along time (coalesced), phase-split on the way in
      const int c = warp;
      const bool cok = c0 + c < Cin;
      const float* xrow = x + ((size_t)b * Cin + c0 + c) * T;
      float* xrow_s = xs + c * (S * G::LI);
      for (int p = lane; p < G::SPAN; p += 32) {
        float v = 0.f;
        if (cok) {
          const int i = in0 + p;
          int src = i - pad;
          if (i < pad) src = pad_mode == 0 ? pad - i : (pad_mode == 1 ? -1 : 0);
          if (src >= 0 && src < T) v = __ldg(xrow + src);
        }
        xrow_s[(p % S) * G::LI + p / S] = v;
      }
    }
    // stage the weights -> ws[c][j][o]
    if (w_packed) {
      // pre-transposed copy [Cin][K][Cout]: consecutive threads read consecutive output channels (coalesced) and
      // write consecutive words (conflict-free)
      for (int i = threadIdx.x; i < CO_TILE * CI * K; i += THREADS) {
        const int o = i % CO_TILE, r = i / CO_TILE;  // r = c * K + j
        float v = 0.f;
        if (o0 + o < Cout && c0 + r / K < Cin) v = __ldg(w + ((size_t)c0 * K + r) * Cout + o0 + o);
        ws[r * CO_TILE + o] = v;
      }
    } else {
      // torch layout w[o][c][j]: runs of CI*K contiguous floats per output channel (slow path: strided smem writes)
      for (int i = threadIdx.x; i < CO_TILE * CI * K; i += THREADS) {
        const int o = i / (CI * K), r = i - o * (CI * K);
        float v = 0.f;
        if (o0 + o < Cout && c0 + r / K < Cin) v = __ldg(w + ((size_t)(o0 + o) * Cin + c0) * K + r);
        ws[r * CO_TILE + o] = v;
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CI; ++c) {
      const float* xc = xs + c * (S * G::LI) + lane;
      const float* wc = ws + c * (K * CO_TILE) + warp * COT;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        float wv[COT], xv[TQ];
#pragma unroll
        for (int i4 = 0; i4 < COT / 4; ++i4) {
          const float4 t4 = *reinterpret_cast<const float4*>(wc + j * CO_TILE + i4 * 4);
          wv[i4 * 4 + 0] = t4.x; wv[i4 * 4 + 1] = t4.y; wv[i4 * 4 + 2] = t4.z; wv[i4 * 4 + 3] = t4.w;
        }
        const int off = ((j * D) % S) * G::LI + (j * D) / S;  // compile-time after unrolling
#pragma unroll
        for (int q = 0; q < TQ; ++q) xv[q] = xc[off + 32 * q];
#pragma unroll
        for (int i = 0; i < COT; ++i)
#pragma unroll
          for (int q = 0; q < TQ; ++q) acc[i][q] = fmaf(wv[i], xv[q], acc[i][q]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < COT; ++i) {
    const int o = o0 + warp * COT + i;
    if (o >= Cout) continue;
    const float bv = bias ? __ldg(bias + o) : 0.f;
#pragma unroll
    for (int q = 0; q < TQ; ++q) {
      const int t = t0 + lane + 32 * q;
      if (t >= Tout) continue;
      float v = acc[i][q] + bv;
      if (act) v = v > 0.f ? v : expm1f(v);
      const size_t idx = ((size_t)b * Cout + o) * Tout + t;
      if (residual) v += residual[idx];
      y[idx] = v;
    }
  }
}

template <int K, int S, int D, int COT, int TQ>
inline int launch(const float* x, const float* w, const float* bias, const float* residual, float* y, int B, int Cin,
                  int Cout, int T, int Tout, int pad, int pad_mode, int act, int w_packed, cudaStream_t stream) {
  constexpr size_t smem = smem_bytes<K, S, D, COT, TQ>();
  static_assert(smem <= 100 * 1024, "conv tile does not fit two CTAs per SM");
  static bool attr = false;
  if (!attr) {
    ALM_CUDA_OK(cudaFuncSetAttribute(conv_kernel<K, S, D, COT, TQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    attr = true;
  }
  dim3 grid(ceil_div(Tout, Geo<K, S, D, TQ>::T_TILE), ceil_div(Cout, 8 * COT), B);
  conv_kernel<K, S, D, COT, TQ><<<grid, THREADS, smem, stream>>>(x, w, bias, residual, y, Cin, Cout, T, Tout, pad,
                                                                  pad_mode, act, w_packed);
  ALM_CHECK_LAUNCH();
  ALM_LAUNCHED(1);
  return ALM_OK;
}

// returns -1 when (K, stride, dilation) is not one of the specialised shapes
inline int dispatch(const float* x, const float* w, const float* bias, const float* residual, float* y, int B, int Cin,
                    int Cout, int T, int Tout, int K, int stride, int dil, int pad, int pad_mode, int act,
                    int w_packed, cudaStream_t stream) {
#define CVT_CASE(KK, SS, DD, TQQ)                                                                                  \
  if (K == KK && stride == SS && dil == DD) {                                                                      \
    if (Cout >= 64)                                                                                                \
      return launch<KK, SS, DD, 8, TQQ>(x, w, bias, residual, y, B, Cin, Cout, T, Tout, pad, pad_mode, act, w_packed, stream); \
    return launch<KK, SS, DD, 4, TQQ>(x, w, bias, residual, y, B, Cin, Cout, T, Tout, pad, pad_mode, act, w_packed, stream);  \
  }
  CVT_CASE(7, 1, 1, 8)
  CVT_CASE(7, 1, 3, 8)
  CVT_CASE(7, 1, 9, 8)
  CVT_CASE(1, 1, 1, 8)
  CVT_CASE(3, 1, 1, 8)
  CVT_CASE(4, 2, 1, 8)
  CVT_CASE(6, 3, 1, 4)
  CVT_CASE(8, 4, 1, 4)
  CVT_CASE(10, 5, 1, 4)
  CVT_CASE(16, 8, 1, 4)
#undef CVT_CASE
  return -1;
}

}  // namespace cvt
}  // namespace alm
