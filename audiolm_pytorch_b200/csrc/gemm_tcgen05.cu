// Persistent, warp-specialised bf16 GEMM for sm_100a.
//   warp 0 : TMA producer (one lane)      — cp.async.bulk.tensor into a STAGES-deep SW128 smem ring
//   warp 1 : UMMA issuer  (one lane)      — tcgen05.mma kind::f16, fp32 accumulators in TMEM
//   warp 2 : TMEM allocator / deallocator
//   warps 4-7 : epilogue                  — tcgen05.ld -> registers -> alpha/bias -> global
// Two TMEM accumulator stages let the epilogue of tile i overlap the mainloop of tile i+1.
// Operands may be K-major or MN-major (UMMA descriptor major bits), which covers forward
// (x W^T), dgrad (dy W) and wgrad (dy^T x) without materialising any transpose.
#include "alm_common.cuh"
#include "ptx_sm100.cuh"

namespace alm {

struct GemmParams {
  void* C;
  const float* bias;
  long long ldc, strideC;
  int M, N, K, batch;
  int m_blocks, n_blocks, k_blocks, split_k;
  int c_fp32, acc_mode;
  int tma_store;  // bf16 overwrite outputs: stage through smem and store with TMA (full 128-B lines)
  float alpha;
  // fused logit head + cross entropy (CE kernel variant only; alm_gemm_head_ce):
  //   ce_mode 1: nothing is stored; every (row, n tile) emits its soft-max partial {max, sum 2^(t - max)} of
  //              t = logit * log2(e) into ce_part [M][n_blocks][2], and the tile that holds the label its logit into ce_lab
  //   ce_mode 2: C (bf16) = (softmax - onehot) * (*ce_num / *ce_den), zero rows where label == ce_ignore
  int ce_mode;
  const long long* ce_labels;
  long long ce_ignore;
  float* ce_part;
  float* ce_lab;
  const float* ce_lse;   // natural-log LSE per row (mode 2)
  const float* ce_num;
  const float* ce_den;
};

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_THREADS = 256;

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int STAGES = BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static constexpr int STAGING_BYTES = 2 * GEMM_BLOCK_M * 128;  // two [128 rows x 128 B] SW128 output tiles
  static constexpr int SMEM_BYTES =
      STAGES * (A_BYTES + B_BYTES) + STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BLOCK_N, bool A_MN, bool B_MN, bool CE = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint8_t* smem_c = smem + STAGES * (Cfg::A_BYTES + Cfg::B_BYTES);  // epilogue staging (TMA store)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;                  // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;        // [STAGES]  MMA -> TMA
  uint64_t* acc_full_bar = bars + 2 * STAGES; // [2]       MMA -> epilogue
  uint64_t* acc_empty_bar = acc_full_bar + 2; // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full_bar[i], 1);
      mbar_init(&acc_empty_bar[i], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_batch = p.m_blocks * p.n_blocks * p.split_k;
  const int total_tiles = tiles_per_batch * p.batch;
  const int kb_per_split = (p.k_blocks + p.split_k - 1) / p.split_k;

  // tile -> (batch, split, m block, n block); n fastest so consecutive CTAs share the A panel
  auto decode = [&](int t, int& b, int& s, int& mb, int& nb) {
    b = t / tiles_per_batch;
    int r = t - b * tiles_per_batch;
    s = r / (p.m_blocks * p.n_blocks);
    r -= s * (p.m_blocks * p.n_blocks);
    mb = r / p.n_blocks;
    nb = r - mb * p.n_blocks;
  };
  auto k_range = [&](int s, int& kb0, int& kb1) {
    kb0 = s * kb_per_split;
    kb1 = min(p.k_blocks, kb0 + kb_per_split);
  };

  if (warp == 0) {
    {
      // ===================== TMA producer =====================
      // the whole warp runs the (uniform) loop; one elected lane issues.  Issuing from a divergent `if (lane == 0)` region
      // makes the compiler wrap every UTMALDG / UTCHMMA / commit in an ELECT + BRA.U.ANY loop (see ptx_sm100.cuh)
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int b, s, mb, nb, kb0, kb1;
        decode(t, b, s, mb, nb);
        k_range(s, kb0, kb1);
        const int m0 = mb * GEMM_BLOCK_M, n0 = nb * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
          uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
          const int k0 = kb * GEMM_BLOCK_K;
          if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::A_BYTES + Cfg::B_BYTES);
          if constexpr (!A_MN) {
            tma_load_3d(sa, &tmA, &full_bar[stage], k0, m0, b);
          } else {
#pragma unroll
            for (int i = 0; i < GEMM_BLOCK_M / 64; ++i)
              tma_load_3d(sa + i * (GEMM_BLOCK_K * 128), &tmA, &full_bar[stage], m0 + i * 64, k0, b);
          }
          if constexpr (!B_MN) {
            tma_load_3d(sb, &tmB, &full_bar[stage], k0, n0, b);
          } else {
#pragma unroll
            for (int i = 0; i < BLOCK_N / 64; ++i)
              tma_load_3d(sb + i * (GEMM_BLOCK_K * 128), &tmB, &full_bar[stage], n0 + i * 64, k0, b);
          }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    {
      // ===================== UMMA issuer (whole warp, one elected lane issues) =====================
      constexpr uint32_t idesc = umma_idesc_bf16_f32(GEMM_BLOCK_M, BLOCK_N, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++iter) {
        int b, s, mb, nb, kb0, kb1;
        decode(t, b, s, mb, nb);
        k_range(s, kb0, kb1);
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1u;
        mbar_wait(&acc_empty_bar[acc], acc_phase ^ 1u);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::B_BYTES);
          if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            const uint64_t da = A_MN ? umma_smem_desc_sw128(a_addr + k * 2048, 1024, GEMM_BLOCK_K * 128)
                                     : umma_smem_desc_sw128(a_addr + k * 32, 1024, 0);
            const uint64_t db = B_MN ? umma_smem_desc_sw128(b_addr + k * 2048, 1024, GEMM_BLOCK_K * 128)
                                     : umma_smem_desc_sw128(b_addr + k * 32, 1024, 0);
            umma_bf16_ss(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (kb == kb1 - 1) umma_commit(&acc_full_bar[acc]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    int iter = 0;
    uint32_t store_seq = 0;  // running count of staged chunks: picks the staging buffer (also across tiles)
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++iter) {
      int b, s, mb, nb;
      decode(t, b, s, mb, nb);
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1u;
      mbar_wait(&acc_full_bar[acc], acc_phase);
      tc_fence_after_sync();
      const int gm = mb * GEMM_BLOCK_M + row;
      const int n0 = nb * BLOCK_N;
      const bool row_ok = gm < p.M;
      const long long row_off = (long long)b * p.strideC + (long long)gm * p.ldc;
      const uint32_t taddr = tmem_base + acc * BLOCK_N + (uint32_t(q * 32) << 16);
      // fused head + cross entropy: this thread's row state for the tile
      float ce_m = -INFINITY, ce_s = 0.f, ce_lse2 = 0.f, ce_scale = 0.f;
      long long ce_label = -1;
      if constexpr (CE) {
        if (row_ok) {
          ce_label = p.ce_labels[gm];
          if (p.ce_mode == 2) {
            ce_lse2 = p.ce_lse[gm] * 1.4426950408889634f;
            ce_scale = ce_label == p.ce_ignore ? 0.f : __ldg(p.ce_num) / __ldg(p.ce_den);
          }
        }
      }
      if (p.tma_store) {
        // TMEM -> registers -> bf16 -> SW128 smem tile -> cp.async.bulk.tensor store (tails clipped by TMA)
        const bool leader = (warp == 4 && lane == 0);
#pragma unroll 1
        for (int cc = 0; cc < BLOCK_N / 64; ++cc) {
          if (n0 + cc * 64 >= p.N) break;  // CTA-uniform: the remaining chunks lie beyond N (nothing staged/stored)
          // staging buffers alternate per ISSUED store (also across tiles), so "at most one bulk group still
          // reading" below always means the other buffer
          uint8_t* stage_c = smem_c + (store_seq & 1u) * (GEMM_BLOCK_M * 128);
          ++store_seq;
          if (leader) tma_store_wait_read<1>();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          uint32_t r0[32], r1[32];
          tmem_ld_32x32b_x32(taddr + cc * 64, r0);
          tmem_ld_32x32b_x32(taddr + cc * 64 + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(r0[g * 8 + 0]) * p.alpha, __uint_as_float(r0[g * 8 + 1]) * p.alpha);
            o.y = pack_bf16x2(__uint_as_float(r0[g * 8 + 2]) * p.alpha, __uint_as_float(r0[g * 8 + 3]) * p.alpha);
            o.z = pack_bf16x2(__uint_as_float(r0[g * 8 + 4]) * p.alpha, __uint_as_float(r0[g * 8 + 5]) * p.alpha);
            o.w = pack_bf16x2(__uint_as_float(r0[g * 8 + 6]) * p.alpha, __uint_as_float(r0[g * 8 + 7]) * p.alpha);
            *reinterpret_cast<uint4*>(stage_c + sw128_offset(row, g)) = o;
            o.x = pack_bf16x2(__uint_as_float(r1[g * 8 + 0]) * p.alpha, __uint_as_float(r1[g * 8 + 1]) * p.alpha);
            o.y = pack_bf16x2(__uint_as_float(r1[g * 8 + 2]) * p.alpha, __uint_as_float(r1[g * 8 + 3]) * p.alpha);
            o.z = pack_bf16x2(__uint_as_float(r1[g * 8 + 4]) * p.alpha, __uint_as_float(r1[g * 8 + 5]) * p.alpha);
            o.w = pack_bf16x2(__uint_as_float(r1[g * 8 + 6]) * p.alpha, __uint_as_float(r1[g * 8 + 7]) * p.alpha);
            *reinterpret_cast<uint4*>(stage_c + sw128_offset(row, 4 + g)) = o;
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (leader) {
            tma_store_3d(&tmC, stage_c, n0 + cc * 64, mb * GEMM_BLOCK_M, b);
            tma_store_commit();
          }
        }
        // every TMEM read of this accumulator has completed (tcgen05.wait::ld above): hand it back to the MMA warp
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty_bar[acc]);
        continue;
      }
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t r[32];
        __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the predicated stores below
        tmem_ld_32x32b_x32(taddr + c * 32, r);
        tmem_ld_wait();
        if (c == BLOCK_N / 32 - 1) {
          // all TMEM reads of this accumulator are done: hand it back to the MMA warp
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty_bar[acc]);
        }
        const int nbase = n0 + c * 32;
        if (row_ok && nbase < p.N) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nbase + j < p.N) v[j] += __ldg(p.bias + nbase + j);
        }
        if constexpr (CE) {
          if (p.ce_mode == 1) {
            float cm = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (nbase + j == ce_label) p.ce_lab[gm] = v[j];
              v[j] = nbase + j < p.N ? v[j] * 1.4426950408889634f : -INFINITY;
              cm = fmaxf(cm, v[j]);
            }
            const float m_new = fmaxf(ce_m, cm);   // finite: the chunk has at least one valid column
            float add = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) add += exp2f(v[j] - m_new);
            ce_s = ce_s * exp2f(ce_m - m_new) + add;
            ce_m = m_new;
            continue;   // nothing is stored in this mode
          }
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = (exp2f(v[j] * 1.4426950408889634f - ce_lse2) - (nbase + j == ce_label ? 1.f : 0.f)) * ce_scale;
        }
        const bool full = nbase + 32 <= p.N;
        if (p.c_fp32) {
          float* dst = reinterpret_cast<float*>(p.C) + row_off + nbase;
          if (p.acc_mode == 2) {
            if (full && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
              // 16-byte vector reductions: a quarter of the L2 atomic operations of scalar red.add (the split-K
              // weight gradients were bound by them: profiles/r01_ncu_gemm_wgrad_w1.txt, 447 MB of red traffic)
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]),
                             "f"(v[j + 2]), "f"(v[j + 3])
                             : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nbase + j < p.N) atomicAdd(dst + j, v[j]);
            }
          } else if (full && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              if (p.acc_mode == 1) {
                float4 old = *reinterpret_cast<float4*>(dst + j);
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
              }
              *reinterpret_cast<float4*>(dst + j) = o;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nbase + j < p.N) dst[j] = (p.acc_mode == 1 ? dst[j] : 0.f) + v[j];
          }
        } else {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + row_off + nbase;
          if (p.acc_mode == 0 && full && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 o;
              o.x = pack_bf16x2(v[j], v[j + 1]);
              o.y = pack_bf16x2(v[j + 2], v[j + 3]);
              o.z = pack_bf16x2(v[j + 4], v[j + 5]);
              o.w = pack_bf16x2(v[j + 6], v[j + 7]);
              *reinterpret_cast<uint4*>(dst + j) = o;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nbase + j < p.N) {
                float o = v[j];
                if (p.acc_mode != 0) o += __bfloat162float(dst[j]);
                dst[j] = __float2bfloat16_rn(o);
              }
          }
        }
        }  // row_ok
      }
      if constexpr (CE) {
        if (p.ce_mode == 1 && row_ok) {
          float* pp = p.ce_part + ((size_t)gm * p.n_blocks + nb) * 2;
          pp[0] = ce_m;
          pp[1] = ce_s;
        }
      }
    }
  }

  if (p.tma_store && warp == 4 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int BLOCK_N, bool A_MN, bool B_MN, bool CE = false>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  auto kfn = gemm_bf16_tcgen05_kernel<BLOCK_N, A_MN, B_MN, CE>;
  static bool attr_set = false;
  if (!attr_set) {
    ALM_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int total = p.m_blocks * p.n_blocks * p.split_k * p.batch;
  const int grid = total < num_sms() ? total : num_sms();
  kfn<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmC, p);
  ALM_CHECK_LAUNCH();
  ALM_LAUNCHED(1);
  return ALM_OK;
}

static int pick_block_n(int N) {
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  // 128x256 tiles have 33 % more FLOP per operand byte than 128x128 (85 vs 64 FLOP/B of smem fill), which is what
  // decides throughput for K ~ 1024; accept up to ~10 % padded columns before falling back to 128-wide tiles
  const int pad256 = ceil_div(N, 256) * 256, pad128 = ceil_div(N, 128) * 128;
  return (pad128 * 10 < pad256 * 9) ? 128 : 256;
}

}  // namespace alm

namespace alm {
struct CeEpilogue {
  int mode;
  const long long* labels;
  long long ignore;
  float* part;
  float* lab;
  const float* lse;
  const float* num;
  const float* den;
};
}  // namespace alm

static int gemm_common(const void* A, int a_mn, int64_t lda, int64_t strideA, const void* B, int b_mn, int64_t ldb,
                       int64_t strideB, void* C, int c_fp32, int64_t ldc, int64_t strideC, int M, int N, int K, int batch,
                       float alpha, const float* bias, int acc_mode, int split_k, cudaStream_t stream,
                       const alm::CeEpilogue* ce) {
  using namespace alm;
  ALM_REQUIRE(A && B && (C || (ce && ce->mode == 1)), ALM_ERR_ARG);
  ALM_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0, ALM_ERR_ARG);
  ALM_REQUIRE(acc_mode >= 0 && acc_mode <= 2 && split_k >= 1, ALM_ERR_ARG);
  ALM_REQUIRE(split_k == 1 || (acc_mode == 2 && c_fp32), ALM_ERR_ARG);
  ALM_REQUIRE(!(a_mn && !b_mn), ALM_ERR_UNSUPPORTED);  // (MN,K) is never needed on this path
  ALM_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && strideA % 8 == 0 && strideB % 8 == 0, ALM_ERR_ALIGN);

  const int BN = pick_block_n(N);
  GemmParams p;
  p.C = C;
  p.bias = bias;
  p.ldc = ldc;
  p.strideC = strideC;
  p.M = M; p.N = N; p.K = K; p.batch = batch;
  p.m_blocks = ceil_div(M, GEMM_BLOCK_M);
  p.n_blocks = ceil_div(N, BN);
  p.k_blocks = ceil_div(K, GEMM_BLOCK_K);
  if (split_k > p.k_blocks) split_k = p.k_blocks;
  // every split must own at least one k block
  while (split_k > 1 && (split_k - 1) * ceil_div(p.k_blocks, split_k) >= p.k_blocks) --split_k;
  p.split_k = split_k;
  p.c_fp32 = c_fp32;
  p.acc_mode = acc_mode;
  p.alpha = alpha;
  p.ce_mode = 0;
  if (ce != nullptr) {
    p.ce_mode = ce->mode;
    p.ce_labels = ce->labels;
    p.ce_ignore = ce->ignore;
    p.ce_part = ce->part;
    p.ce_lab = ce->lab;
    p.ce_lse = ce->lse;
    p.ce_num = ce->num;
    p.ce_den = ce->den;
  }

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[3], strides[3];
    uint32_t box[3];
    if (!a_mn) {
      dims[0] = (uint64_t)K; dims[1] = (uint64_t)M;
      box[0] = GEMM_BLOCK_K; box[1] = GEMM_BLOCK_M;
    } else {
      dims[0] = (uint64_t)M; dims[1] = (uint64_t)K;
      box[0] = 64; box[1] = GEMM_BLOCK_K;
    }
    dims[2] = (uint64_t)batch; box[2] = 1;
    strides[0] = 2; strides[1] = (uint64_t)lda * 2;
    strides[2] = batch > 1 ? (uint64_t)strideA * 2 : dims[1] * strides[1];
    int rc = make_tensor_map(&tmA, A, 2, 3, dims, strides, box, true);
    if (rc != ALM_OK) return rc;
  }
  {
    uint64_t dims[3], strides[3];
    uint32_t box[3];
    if (!b_mn) {
      dims[0] = (uint64_t)K; dims[1] = (uint64_t)N;
      box[0] = GEMM_BLOCK_K; box[1] = (uint32_t)BN;
    } else {
      dims[0] = (uint64_t)N; dims[1] = (uint64_t)K;
      box[0] = 64; box[1] = GEMM_BLOCK_K;
    }
    dims[2] = (uint64_t)batch; box[2] = 1;
    strides[0] = 2; strides[1] = (uint64_t)ldb * 2;
    strides[2] = batch > 1 ? (uint64_t)strideB * 2 : dims[1] * strides[1];
    int rc = make_tensor_map(&tmB, B, 2, 3, dims, strides, box, true);
    if (rc != ALM_OK) return rc;
  }

  // bf16 overwrite outputs with 16-B aligned rows take the smem-staged TMA-store epilogue
  CUtensorMap tmC = tmA;
  p.tma_store = 0;
  if (ce == nullptr && !c_fp32 && acc_mode == 0 && bias == nullptr && ldc % 8 == 0 && (batch == 1 || strideC % 8 == 0) &&
      (reinterpret_cast<uintptr_t>(C) & 15u) == 0) {
    uint64_t dims[3] = {(uint64_t)N, (uint64_t)M, (uint64_t)batch};
    uint64_t strides[3] = {2, (uint64_t)ldc * 2, batch > 1 ? (uint64_t)strideC * 2 : (uint64_t)M * ldc * 2};
    uint32_t box[3] = {64, GEMM_BLOCK_M, 1};
    int rc = make_tensor_map(&tmC, C, 2, 3, dims, strides, box, true);
    if (rc != ALM_OK) return rc;
    p.tma_store = 1;
  }

  if (ce != nullptr) {   // (row-major x, row-major head weight: the only layout the fused head needs)
    if (BN == 256) return launch_gemm<256, false, false, true>(tmA, tmB, tmC, p, stream);
    if (BN == 128) return launch_gemm<128, false, false, true>(tmA, tmB, tmC, p, stream);
    return launch_gemm<64, false, false, true>(tmA, tmB, tmC, p, stream);
  }
#define ALM_GEMM_DISPATCH(BN_)                                                             \
  if (!a_mn && !b_mn) return launch_gemm<BN_, false, false>(tmA, tmB, tmC, p, stream);     \
  if (!a_mn && b_mn) return launch_gemm<BN_, false, true>(tmA, tmB, tmC, p, stream);       \
  return launch_gemm<BN_, true, true>(tmA, tmB, tmC, p, stream);
  if (BN == 256) { ALM_GEMM_DISPATCH(256) }
  if (BN == 128) { ALM_GEMM_DISPATCH(128) }
  { ALM_GEMM_DISPATCH(64) }
#undef ALM_GEMM_DISPATCH
}

extern "C" int alm_gemm_bf16(const void* A, int a_mn, int64_t lda, int64_t strideA, const void* B, int b_mn,
                             int64_t ldb, int64_t strideB, void* C, int c_fp32, int64_t ldc, int64_t strideC, int M,
                             int N, int K, int batch, float alpha, const float* bias, int acc_mode, int split_k,
                             alm_stream_t stream_) {
  return gemm_common(A, a_mn, lda, strideA, B, b_mn, ldb, strideB, C, c_fp32, ldc, strideC, M, N, K, batch, alpha, bias,
                     acc_mode, split_k, reinterpret_cast<cudaStream_t>(stream_), nullptr);
}

// number of n tiles of the head GEMM for a vocabulary of V (= the second dimension of `part` below)
extern "C" int alm_gemm_head_ce_tiles(int V) { return alm::ceil_div(V, alm::pick_block_n(V)); }

// Fused logit head + cross entropy (audiolm_pytorch.py:621, 798, 965-983, 1325-1361 heads; :1561-1565, 1836-1854,
// 2119-2137 F.cross_entropy): the [M, V] fp32 logits never reach HBM.
//   mode 1: logits = X W^T (+ bias) are reduced in the GEMM epilogue to per-(row, n tile) soft-max partials
//           part [M, tiles, 2] = {max, sum 2^(t - max)} of t = logit * log2(e), and lab_logit [M] = logit[label];
//           alm_ce_finish turns them into the row LSE and loss
//   mode 2: the GEMM is recomputed and its epilogue writes d(loss)/d(logits) = (softmax - onehot) * (*scale_num /
//           *scale_den) as bf16 [M, ldd] (rows with label == ignore_index are zero; columns >= V are not written)
extern "C" int alm_gemm_head_ce(const void* X, int64_t ldx, const void* W, int64_t ldw, const float* bias,
                                const int64_t* labels, int64_t ignore_index, int mode, float* part, float* lab_logit,
                                const float* lse, const float* scale_num, const float* scale_den, void* dlogits,
                                int64_t ldd, int M, int V, int K, alm_stream_t stream_) {
  ALM_REQUIRE(mode == 1 || mode == 2, ALM_ERR_ARG);
  ALM_REQUIRE(labels != nullptr, ALM_ERR_ARG);
  if (mode == 1) ALM_REQUIRE(part && lab_logit, ALM_ERR_ARG);
  else ALM_REQUIRE(lse && scale_num && scale_den && dlogits && ldd >= V, ALM_ERR_ARG);
  alm::CeEpilogue ce{mode, reinterpret_cast<const long long*>(labels), (long long)ignore_index, part, lab_logit, lse,
                     scale_num, scale_den};
  return gemm_common(X, 0, ldx, 0, W, 0, ldw, 0, dlogits, 0, ldd, 0, M, V, K, 1, 1.f, bias, 0, 1,
                     reinterpret_cast<cudaStream_t>(stream_), &ce);
}
