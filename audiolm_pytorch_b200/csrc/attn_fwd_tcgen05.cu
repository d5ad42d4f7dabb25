// Multi-query causal attention forward for sm_100a (replaces attend.py:69-146 as called from
// audiolm_pytorch.py:390): softmax(q k^T * d^-1/2, masked by key-padding mask and right-aligned causal
// mask) v, with ONE shared k/v head of width 64 for all query heads.
//
// One CTA = (batch b, head h, 128 queries).  Q K^T and P V run on tcgen05 with fp32 accumulators in
// TMEM; Q/K/V tiles arrive by TMA into 128-B-swizzled smem; the softmax is an online (flash) softmax
// held in registers by 256 threads: two warps per TMEM lane quadrant, each thread owns one query row and
// half of the key columns of a tile (the row maximum is taken over all 128 columns by both).
//   warps 0-7 : softmax / rescale / output          warp 8 : TMA producer      warp 9 : UMMA issuer
// An optional additive score bias [h, n_q, n_k] (flash_attn=False path, attend.py:122-124) is added in the
// softmax warps; tiles that lie fully below the causal diagonal take a predicate-free path.
// TMEM: S[128x128] at columns 0..127, (P V)[128x64] at columns 128..191 (256 columns allocated so two
// CTAs can be resident per SM and overlap each other's softmax and MMA phases).
#include "alm_common.cuh"
#include "ptx_sm100.cuh"

namespace alm {

constexpr int ATT_BM = 128;     // queries per CTA
constexpr int ATT_BN = 128;     // keys per tile
constexpr int ATT_D = 64;       // head width (dim_head)
constexpr int ATT_KV_STAGES = 2;
constexpr int ATT_THREADS = 320;   // warps 0-7 softmax (2 per TMEM lane quadrant), warp 8 TMA, warp 9 MMA
constexpr int ATT_TMA_WARP = 8, ATT_MMA_WARP = 9;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB: one [128 x 64] bf16 SW128 tile
constexpr int ATT_SMEM_BYTES = ATT_TILE_BYTES * (1 + 2 + 2 * ATT_KV_STAGES) + 256;  // 2 CTAs/SM: must stay <= 113 KB
constexpr int ATT_TMEM_COLS = 256;

struct AttnFwdParams {
  __nv_bfloat16* o;       // [b, n_q, h*64] row stride ldo
  float* lse;             // [b, h, lse_stride] log2-domain LSE of the scaled scores (for backward); may be null
  const uint32_t* kmask;  // packed key mask (alm_pack_key_mask): bit i of word w of row b = key 32 w + i may be attended; may be null
  int kb_stride;          // words per batch row: 4 * ceil(n_k / 128)
  const float* bias;      // [h, n_q, bias_rs] additive score bias (natural-log domain, added after the scale); may be null
  long long bias_hs, bias_rs;  // element strides between heads / query rows (bias_rs % 4 == 0, >= n_k)
  long long ldo, lse_stride;
  int b, h, n_q, n_k;
  int causal;
  float scale_log2;       // d^-1/2 * log2(e)
};

__device__ __forceinline__ float att_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool HAS_BIAS>
__global__ void __launch_bounds__(ATT_THREADS, 2)
mqa_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) {  // SW128 tiles need 1024-B alignment
    if (threadIdx.x == 0) printf("[alm] attn fwd: dynamic smem base not 1024-B aligned\n");
    __trap();
  }
  uint8_t* sQ = smem;
  uint8_t* sP = sQ + ATT_TILE_BYTES;                   // two [128 x 64-key] tiles
  uint8_t* sK = sP + 2 * ATT_TILE_BYTES;               // [stages]
  uint8_t* sV = sK + ATT_KV_STAGES * ATT_TILE_BYTES;   // [stages]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_KV_STAGES * ATT_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                        // [stages]
  uint64_t* kv_empty = kv_full + ATT_KV_STAGES;        // [stages]
  uint64_t* s_full = kv_empty + ATT_KV_STAGES;
  uint64_t* p_full = s_full + 1;
  uint64_t* pv_full = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_qblocks = (p.n_q + ATT_BM - 1) / ATT_BM;
  const int qb = n_qblocks - 1 - (int)blockIdx.x;  // heavy (late) query blocks first
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int q0 = qb * ATT_BM;
  const int off = p.n_k - p.n_q;  // right alignment of queries against keys (KV cache)
  int kv_end = p.n_k;
  if (p.causal) kv_end = min(p.n_k, q0 + ATT_BM + off);
  const int n_tiles = kv_end > 0 ? (kv_end + ATT_BN - 1) / ATT_BN : 0;

  if (warp == ATT_TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < ATT_KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);  // one arrive per softmax warp
    mbar_init(pv_full, 1);
    fence_mbar_init();
  }
  if (warp == ATT_MMA_WARP) tmem_alloc(tmem_slot, ATT_TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_PV = tmem_base + ATT_BN;

  if (warp == ATT_TMA_WARP) {
    if (n_tiles > 0) {
      // ---------------- TMA producer (whole warp runs the loop, one elected lane issues: see elect_one_sync) -------
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
        tma_load_3d(sQ, &tmQ, q_full, head * ATT_D, q0, batch);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&kv_full[stage], 2 * ATT_TILE_BYTES);
          tma_load_3d(sK + stage * ATT_TILE_BYTES, &tmK, &kv_full[stage], 0, j * ATT_BN, batch);
          tma_load_3d(sV + stage * ATT_TILE_BYTES, &tmV, &kv_full[stage], 0, j * ATT_BN, batch);
        }
        if (++stage == ATT_KV_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == ATT_MMA_WARP) {
    if (n_tiles > 0) {
      // ---------------- UMMA issuer (whole warp, one elected lane issues) ----------------
      constexpr uint32_t idesc_s = umma_idesc_bf16_f32(ATT_BM, ATT_BN, false, false);   // Q K^T
      constexpr uint32_t idesc_pv = umma_idesc_bf16_f32(ATT_BM, ATT_D, false, true);    // P V (V is MN-major)
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t p_addr = smem_u32(sP);
      mbar_wait(q_full, 0);
      auto issue_s = [&](int stage) {
        const uint32_t k_addr = smem_u32(sK + stage * ATT_TILE_BYTES);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k)
            umma_bf16_ss(tmem_S, umma_smem_desc_sw128(q_addr + k * 32, 1024, 0),
                         umma_smem_desc_sw128(k_addr + k * 32, 1024, 0), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(s_full);
        }
      };
      int stage = 0;
      uint32_t phase = 0;
      mbar_wait(&kv_full[0], 0);
      tc_fence_after_sync();
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(p_full, j & 1);
        tc_fence_after_sync();
        const uint32_t v_addr = smem_u32(sV + stage * ATT_TILE_BYTES);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < ATT_BN / 16; ++k)
            umma_bf16_ss(tmem_PV, umma_smem_desc_sw128(p_addr + (k >> 2) * ATT_TILE_BYTES + (k & 3) * 32, 1024, 0),
                         umma_smem_desc_sw128(v_addr + k * 2048, 1024, 0), idesc_pv, k > 0 ? 1u : 0u);
          umma_commit(pv_full);
          umma_commit(&kv_empty[stage]);
        }
        if (++stage == ATT_KV_STAGES) { stage = 0; phase ^= 1u; }
        if (j + 1 < n_tiles) {
          mbar_wait(&kv_full[stage], phase);
          tc_fence_after_sync();
          issue_s(stage);  // S buffer is free: the softmax warps finished reading S_j before p_full
        }
      }
    }
  } else {
    // ---------------- softmax warps ----------------
    // thread == query row; warps w and w+4 share a TMEM lane quadrant: `half` 0/1 owns key columns
    // [0,64) / [64,128) of every S tile (== P k-tile 0 / 1) and output channels [0,32) / [32,64).
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;
    const int qi = q0 + row;
    const uint32_t lane_sel = uint32_t(quad * 32) << 16;
    const int pair_bar = 1 + quad;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;
    float o_acc[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) o_acc[d] = 0.f;
    const int q_limit = p.causal ? qi + off : p.n_k - 1;  // last key index this query may see
    const uint32_t* mrow = p.kmask ? p.kmask + (long long)batch * p.kb_stride : nullptr;
    uint8_t* ptile = sP + half * ATT_TILE_BYTES;

    auto add_pv = [&](float a) {
      uint32_t r[32];
      __syncwarp();
      tmem_ld_32x32b_x32(tmem_PV + lane_sel + half * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) o_acc[e] = fmaf(o_acc[e], a, __uint_as_float(r[e]));
    };

    constexpr float kLog2e = 1.4426950408889634f;
    const float* brow = nullptr;  // this query row of the bias (rows past n_q are clamped: their output is dropped)
    if constexpr (HAS_BIAS)
      brow = p.bias + (long long)head * p.bias_hs + (long long)min(qi, p.n_q - 1) * p.bias_rs;
    // bias values of 4 consecutive keys, pre-multiplied into the log2 domain; columns past the padded row read 0
    auto bias4 = [&](int col) -> float4 {
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (HAS_BIAS) {
        if (col + 3 < p.bias_rs) bv = __ldg(reinterpret_cast<const float4*>(brow + col));
        bv.x *= kLog2e; bv.y *= kLog2e; bv.z *= kLog2e; bv.w *= kLog2e;
      }
      return bv;
    };

    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after_sync();
      const int kbase = j * ATT_BN;
      // CTA-uniform: every row of the block sees every key of this tile (no key mask, fully below the diagonal)
      const bool tile_full = mrow == nullptr && kbase + ATT_BN <= p.n_k && (!p.causal || kbase + ATT_BN - 1 <= q0 + off);
      // key validity bits of the whole 128-key tile (the row max needs all of it)
      uint32_t valid[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
      if (!tile_full) {
        if (mrow != nullptr) {  // one 16-B load: the 128 key bits of this tile (shared by every row of the CTA)
          const uint4 mv = __ldg(reinterpret_cast<const uint4*>(mrow + j * 4));
          valid[0] = mv.x; valid[1] = mv.y; valid[2] = mv.z; valid[3] = mv.w;
        }
        const int lim = min(q_limit, p.n_k - 1) - kbase;  // keys 0..lim of this tile are in range
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const int hi = lim - w * 32;
          const uint32_t range = hi >= 31 ? 0xFFFFFFFFu : (hi < 0 ? 0u : ((2u << hi) - 1u));
          valid[w] &= range;
        }
      }
      // pass 1: row max over all 128 keys (both warps of a quadrant compute it redundantly: TMEM reads are
      // cheap, and it avoids a cross-warp exchange); pass 2 below only touches this warp's 64 columns.
      // m_tile is in the log2 domain when a bias is present, else raw (scaled once after the loop).
      float m_tile = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld_32x32b_x32(tmem_S + lane_sel + c * 32, r);
        tmem_ld_wait();
        if constexpr (HAS_BIAS) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 bv = bias4(kbase + c * 32 + g * 4);
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float sv = fmaf(__uint_as_float(r[g * 4 + e]), p.scale_log2, bb[e]);
              if (tile_full || ((valid[c] >> (g * 4 + e)) & 1u)) m_tile = fmaxf(m_tile, sv);
            }
          }
        } else if (tile_full) {
#pragma unroll
          for (int e = 0; e < 32; ++e) m_tile = fmaxf(m_tile, __uint_as_float(r[e]));
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if ((valid[c] >> e) & 1u) m_tile = fmaxf(m_tile, __uint_as_float(r[e]));
        }
      }
      const float m_new = fmaxf(m_run, HAS_BIAS ? m_tile : m_tile * p.scale_log2);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = att_ex2(m_run - m_use);  // m_run == -inf -> 0
      // previous tile's P V must be consumed before sP is overwritten
      if (j > 0) {
        mbar_wait(pv_full, (j - 1) & 1);
        tc_fence_after_sync();
        add_pv(alpha_prev);
      }
      // pass 2: p = exp2(s*scale + bias - m) -> bf16 P in the SW128 K-major A-operand layout
      float l_tile = 0.f;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld_32x32b_x32(tmem_S + lane_sel + half * 64 + cc * 32, r);
        tmem_ld_wait();
        float pe[32];
        if constexpr (HAS_BIAS) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 bv = bias4(kbase + half * 64 + cc * 32 + g * 4);
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const bool ok = tile_full || ((valid[half * 2 + cc] >> (g * 4 + e)) & 1u);
              const float pv = ok ? att_ex2(fmaf(__uint_as_float(r[g * 4 + e]), p.scale_log2, bb[e]) - m_use) : 0.f;
              pe[g * 4 + e] = pv;
              l_tile += pv;
            }
          }
        } else if (tile_full) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float pv = att_ex2(fmaf(__uint_as_float(r[e]), p.scale_log2, -m_use));
            pe[e] = pv;
            l_tile += pv;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float pv = ((valid[half * 2 + cc] >> e) & 1u)
                                 ? att_ex2(fmaf(__uint_as_float(r[e]), p.scale_log2, -m_use)) : 0.f;
            pe[e] = pv;
            l_tile += pv;
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 pk;
          pk.x = pack_bf16x2(pe[g * 8 + 0], pe[g * 8 + 1]);
          pk.y = pack_bf16x2(pe[g * 8 + 2], pe[g * 8 + 3]);
          pk.z = pack_bf16x2(pe[g * 8 + 4], pe[g * 8 + 5]);
          pk.w = pack_bf16x2(pe[g * 8 + 6], pe[g * 8 + 7]);
          *reinterpret_cast<uint4*>(ptile + sw128_offset(row, cc * 4 + g)) = pk;
        }
      }
      l_run = fmaf(l_run, alpha, l_tile);  // partial sum over my 64 columns; halves are added at the end
      m_run = m_new;
      alpha_prev = alpha;
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }

    if (n_tiles > 0) {
      mbar_wait(pv_full, (n_tiles - 1) & 1);
      tc_fence_after_sync();
      add_pv(alpha_prev);
    }
    // total row sum = both halves (the P staging tile is free now: every P V has completed)
    float* lx = reinterpret_cast<float*>(sP);
    lx[half * 128 + row] = l_run;
    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    const float l_tot = lx[row] + lx[128 + row];
    if (qi < p.n_q) {
      const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
      __nv_bfloat16* dst = p.o + ((long long)batch * p.n_q + qi) * p.ldo + head * ATT_D + half * 32;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 pk;
        pk.x = pack_bf16x2(o_acc[g * 8 + 0] * inv, o_acc[g * 8 + 1] * inv);
        pk.y = pack_bf16x2(o_acc[g * 8 + 2] * inv, o_acc[g * 8 + 3] * inv);
        pk.z = pack_bf16x2(o_acc[g * 8 + 4] * inv, o_acc[g * 8 + 5] * inv);
        pk.w = pack_bf16x2(o_acc[g * 8 + 6] * inv, o_acc[g * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(dst + g * 8) = pk;
      }
      if (p.lse != nullptr && half == 0) {
        const float lse = l_tot > 0.f ? (m_run + log2f(l_tot)) : INFINITY;  // log2 domain (x ln2 = natural)
        p.lse[((long long)batch * p.h + head) * p.lse_stride + qi] = lse;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == ATT_MMA_WARP) tmem_dealloc(tmem_base, ATT_TMEM_COLS);
}

}  // namespace alm

namespace alm {
// key mask bytes [b, n_k] (non-zero = attend) -> bits [b, 4 * ceil(n_k / 128)] (keys past n_k: 0)
__global__ void pack_key_mask_kernel(const uint8_t* __restrict__ mask, uint32_t* __restrict__ bits, int n_k, int words) {
  const int b = blockIdx.y;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= words) return;
  const int k = w * 32 + (threadIdx.x & 31);
  const bool on = k < n_k && mask[(long long)b * n_k + k] != 0;
  const uint32_t v = __ballot_sync(0xffffffffu, on);
  if ((threadIdx.x & 31) == 0) bits[(long long)b * words + w] = v;
}
}  // namespace alm

extern "C" int alm_pack_key_mask(const void* key_mask, void* bits, int b, int n_k, alm_stream_t stream_) {
  using namespace alm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALM_REQUIRE(key_mask && bits && b > 0 && n_k > 0, ALM_ERR_ARG);
  const int words = (n_k + 127) / 128 * 4;
  dim3 grid(ceil_div(words, 8), b);
  pack_key_mask_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint8_t*>(key_mask),
                                                 reinterpret_cast<uint32_t*>(bits), n_k, words);
  ALM_CHECK_LAUNCH();
  ALM_LAUNCHED(1);
  return ALM_OK;
}

extern "C" int alm_mqa_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, int64_t k_bstride,
                                const void* v, int64_t ldv, int64_t v_bstride, const void* key_mask, void* o,
                                int64_t ldo, float* lse, int64_t lse_stride, const float* bias, int64_t bias_hstride,
                                int64_t bias_rstride, int b, int h, int n_q, int n_k, int causal, float scale,
                                alm_stream_t stream_) {
  using namespace alm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALM_REQUIRE(q && k && v && o, ALM_ERR_ARG);
  ALM_REQUIRE(b > 0 && h > 0 && n_q > 0 && n_k > 0 && n_k >= n_q, ALM_ERR_ARG);
  ALM_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, ALM_ERR_ALIGN);
  ALM_REQUIRE(k_bstride % 8 == 0 && v_bstride % 8 == 0, ALM_ERR_ALIGN);
  ALM_REQUIRE((reinterpret_cast<uintptr_t>(o) & 15u) == 0, ALM_ERR_ALIGN);
  if (bias != nullptr) {
    ALM_REQUIRE(bias_rstride >= n_k && bias_rstride % 4 == 0 && bias_hstride % 4 == 0, ALM_ERR_ALIGN);
    ALM_REQUIRE((reinterpret_cast<uintptr_t>(bias) & 15u) == 0, ALM_ERR_ALIGN);
  }

  CUtensorMap tmQ, tmK, tmV;
  {
    uint64_t dims[3] = {(uint64_t)h * ATT_D, (uint64_t)n_q, (uint64_t)b};
    uint64_t strides[3] = {2, (uint64_t)ldq * 2, (uint64_t)n_q * ldq * 2};
    uint32_t box[3] = {ATT_D, ATT_BM, 1};
    int rc = make_tensor_map(&tmQ, q, 2, 3, dims, strides, box, true);
    if (rc != ALM_OK) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)ATT_D, (uint64_t)n_k, (uint64_t)b};
    uint64_t strides[3] = {2, (uint64_t)ldk * 2, (uint64_t)k_bstride * 2};
    uint32_t box[3] = {ATT_D, ATT_BN, 1};
    int rc = make_tensor_map(&tmK, k, 2, 3, dims, strides, box, true);
    if (rc != ALM_OK) return rc;
    strides[1] = (uint64_t)ldv * 2;
    strides[2] = (uint64_t)v_bstride * 2;
    rc = make_tensor_map(&tmV, v, 2, 3, dims, strides, box, true);
    if (rc != ALM_OK) return rc;
  }
  AttnFwdParams p;
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.lse = lse;
  p.kmask = reinterpret_cast<const uint32_t*>(key_mask);
  p.kb_stride = (n_k + 127) / 128 * 4;
  p.bias = bias;
  p.bias_hs = bias_hstride;
  p.bias_rs = bias_rstride;
  p.ldo = ldo;
  p.lse_stride = lse_stride;
  p.b = b; p.h = h; p.n_q = n_q; p.n_k = n_k;
  p.causal = causal;
  p.scale_log2 = scale * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    ALM_CUDA_OK(cudaFuncSetAttribute(mqa_attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     ATT_SMEM_BYTES));
    ALM_CUDA_OK(cudaFuncSetAttribute(mqa_attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     ATT_SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((n_q + ATT_BM - 1) / ATT_BM, h, b);
  if (bias != nullptr)
    mqa_attn_fwd_kernel<true><<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  else
    mqa_attn_fwd_kernel<false><<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  ALM_CHECK_LAUNCH();
  ALM_LAUNCHED(1);
  return ALM_OK;
}
