// Multi-query causal attention backward for sm_100a (autograd of attend.py:69-146), two kernels:
//
//  dKV kernel: one CTA = (batch, 128 keys).  It walks every (head, query block) that can see these keys
//      S^T = K Q^T,  dP^T = V dO^T                       (tcgen05, accumulators in TMEM)
//      P^T = exp(S^T*scale - lse),  dS^T = scale * P^T (dP^T - delta)   (128 threads, thread == key row)
//      dV += P^T dO,  dK += dS^T Q                       (accumulated in TMEM over all heads: MQA shares k/v,
//                                                          so no atomics are needed)
//  dQ kernel: one CTA = (batch, head, 128 queries), walks the key tiles:
//      S = Q K^T, dP = dO V^T, dS = scale * P (dP - delta), dQ += dS K.
//
// Q/K/V/dO tiles arrive by TMA (SWIZZLE_128B); the SAME smem tile serves as a K-major operand
// (contraction over the 64-wide head dim) and as an MN-major operand (contraction over its 128 rows).
#include "alm_common.cuh"
#include "ptx_sm100.cuh"

namespace alm {

constexpr int AB_T = 128;                     // tile edge (queries or keys)
constexpr int AB_D = 64;
constexpr int AB_TILE = AB_T * AB_D * 2;      // 16 KB
constexpr int AB_THREADS = 576;          // warps 0-15 compute (4 per TMEM lane quadrant), 16 TMA, 17 MMA
constexpr int AB_TMA_WARP = 16, AB_MMA_WARP = 17;
constexpr int AB_STAGES = 2;
constexpr float LOG2E = 1.4426950408889634f;

struct AttnBwdParams {
  const float* lse;      // [b, h, n_q_pad]  log2-domain LSE (m + log2 l) as written by the forward
  const float* delta;    // [b, h, n_q_pad]
  const uint32_t* kmask;  // packed key mask bits (alm_pack_key_mask) or null
  int kb_stride;          // words per batch row
  const float* bias;     // [h, n_q, bias_rs] additive score bias (as given to the forward) or null
  float* dbias;          // same layout, fp32: d(bias) is ACCUMULATED (red.add) over batches / calls; or null
  long long bias_hs, bias_rs;
  __nv_bfloat16* dq;     // [b, n_q, h*64], row stride lddq
  __nv_bfloat16* dk;     // [b, n_k, 64], row stride lddk
  __nv_bfloat16* dv;
  long long lddq, lddk, lddv;
  int b, h, n_q, n_k, n_q_pad;
  int causal;
  float scale, scale_log2;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// write 8 consecutive bf16 (columns col8*8 .. +7 of `row`) into a [128 x 128] bf16 operand stored as two
// [128 x 64] SW128 tiles (K-major A operand)
__device__ __forceinline__ void store_a_chunk(uint8_t* tiles, int row, int col8, const float* v) {
  uint4 pk;
  pk.x = pack_bf16x2(v[0], v[1]);
  pk.y = pack_bf16x2(v[2], v[3]);
  pk.z = pack_bf16x2(v[4], v[5]);
  pk.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(tiles + (col8 >> 3) * AB_TILE + sw128_offset(row, col8 & 7)) = pk;
}

// ================================================================================================
// dK / dV
// ================================================================================================
constexpr int DKV_SMEM = AB_TILE * (2 + 2 * AB_STAGES + 2 + 2) + AB_STAGES * 2 * 512 + 256;

template <bool HAS_BIAS>
__global__ void __launch_bounds__(AB_THREADS, 1)
mqa_attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                        const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sK = smem;
  uint8_t* sV = sK + AB_TILE;
  uint8_t* sQ = sV + AB_TILE;                    // [stages]
  uint8_t* sdO = sQ + AB_STAGES * AB_TILE;       // [stages]
  uint8_t* sPT = sdO + AB_STAGES * AB_TILE;      // 2 tiles
  uint8_t* sdST = sPT + 2 * AB_TILE;             // 2 tiles
  float* sLse = reinterpret_cast<float*>(sdST + 2 * AB_TILE);  // [stages][128]
  float* sDelta = sLse + AB_STAGES * AB_T;                     // [stages][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + AB_STAGES * AB_T);
  uint64_t* kv_full = bars;
  uint64_t* qdo_full = bars + 1;              // [stages]
  uint64_t* qdo_empty = qdo_full + AB_STAGES; // [stages]
  uint64_t* s_full = qdo_empty + AB_STAGES;
  uint64_t* p_full = s_full + 1;
  uint64_t* acc_full = p_full + 1;
  uint64_t* tmem_free = acc_full + 1;   // compute -> MMA: S^T / dP^T of this iteration are in registers
  uint64_t* pds_free = tmem_free + 1;   // MMA -> compute: dV / dK MMAs of this iteration consumed P^T / dS^T
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pds_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // heaviest key blocks first across ALL batches (LPT order): CTA x -> (kb = x / b, batch = x % b).  With the
  // old (kb fastest) order a late-starting kb=0 CTA stretched the makespan to 200 iterations vs 118 ideal.
  const int kb = blockIdx.x / p.b, batch = blockIdx.x % p.b;
  const int k0 = kb * AB_T;
  const int off = p.n_k - p.n_q;
  const int n_qblocks = (p.n_q + AB_T - 1) / AB_T;
  int qb_min = 0;
  if (p.causal && k0 - off > 0) qb_min = (k0 - off) / AB_T;
  const int q_per_head = n_qblocks - qb_min;
  const int n_iter = q_per_head > 0 ? q_per_head * p.h : 0;

  if (warp == AB_TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    mbar_init(kv_full, 1);
    for (int i = 0; i < AB_STAGES; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, 16);
    mbar_init(acc_full, 1);
    mbar_init(tmem_free, 16);
    mbar_init(pds_free, 1);
    fence_mbar_init();
  }
  if (warp == AB_MMA_WARP) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_ST = tmem_base, tmem_dPT = tmem_base + 128, tmem_dV = tmem_base + 256,
                 tmem_dK = tmem_base + 320;

  if (warp == AB_TMA_WARP) {
    if (n_iter > 0) {  // whole warp runs the loop; one elected lane issues (see elect_one_sync)
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(kv_full, 2 * AB_TILE);
        tma_load_3d(sK, &tmK, kv_full, 0, k0, batch);
        tma_load_3d(sV, &tmV, kv_full, 0, k0, batch);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_iter; ++it) {
        const int head = it / q_per_head, qb = qb_min + it % q_per_head;
        mbar_wait(&qdo_empty[stage], phase ^ 1u);
        const size_t roff = ((size_t)batch * p.h + head) * p.n_q_pad + (size_t)qb * AB_T;
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&qdo_full[stage], 2 * AB_TILE + 2 * AB_T * 4);
          tma_load_3d(sQ + stage * AB_TILE, &tmQ, &qdo_full[stage], head * AB_D, qb * AB_T, batch);
          tma_load_3d(sdO + stage * AB_TILE, &tmdO, &qdo_full[stage], head * AB_D, qb * AB_T, batch);
          bulk_load_1d(sLse + stage * AB_T, p.lse + roff, AB_T * 4, &qdo_full[stage]);
          bulk_load_1d(sDelta + stage * AB_T, p.delta + roff, AB_T * 4, &qdo_full[stage]);
        }
        if (++stage == AB_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == AB_MMA_WARP) {
    if (n_iter > 0) {  // whole warp runs the loop; one elected lane issues (see elect_one_sync)
      constexpr uint32_t idesc_s = umma_idesc_bf16_f32(AB_T, AB_T, false, false);
      constexpr uint32_t idesc_acc = umma_idesc_bf16_f32(AB_T, AB_D, false, true);
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), pt_addr = smem_u32(sPT),
                     dst_addr = smem_u32(sdST);
      mbar_wait(kv_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      auto issue_s = [&](int st) {
        const uint32_t q_addr = smem_u32(sQ + st * AB_TILE), do_addr = smem_u32(sdO + st * AB_TILE);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < AB_D / 16; ++k)
            umma_bf16_ss(tmem_ST, umma_smem_desc_sw128(k_addr + k * 32, 1024, 0),
                         umma_smem_desc_sw128(q_addr + k * 32, 1024, 0), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < AB_D / 16; ++k)
            umma_bf16_ss(tmem_dPT, umma_smem_desc_sw128(v_addr + k * 32, 1024, 0),
                         umma_smem_desc_sw128(do_addr + k * 32, 1024, 0), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(s_full);
        }
      };
      mbar_wait(&qdo_full[0], 0);
      tc_fence_after_sync();
      issue_s(0);
      for (int it = 0; it < n_iter; ++it) {
        // software pipeline: as soon as the compute warps hold S^T / dP^T of iteration `it` in registers, the next
        // iteration's S^T / dP^T MMAs are issued, so they run under this iteration's exp / dS arithmetic
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == AB_STAGES) { nstage = 0; nphase ^= 1u; }
        if (it + 1 < n_iter) {
          mbar_wait(tmem_free, it & 1);
          mbar_wait(&qdo_full[nstage], nphase);
          tc_fence_after_sync();
          issue_s(nstage);
        }
        mbar_wait(p_full, it & 1);
        tc_fence_after_sync();
        const uint32_t q_addr = smem_u32(sQ + stage * AB_TILE), do_addr = smem_u32(sdO + stage * AB_TILE);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < AB_T / 16; ++k)
            umma_bf16_ss(tmem_dV, umma_smem_desc_sw128(pt_addr + (k >> 2) * AB_TILE + (k & 3) * 32, 1024, 0),
                         umma_smem_desc_sw128(do_addr + k * 2048, 1024, 0), idesc_acc, (it > 0 || k > 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < AB_T / 16; ++k)
            umma_bf16_ss(tmem_dK, umma_smem_desc_sw128(dst_addr + (k >> 2) * AB_TILE + (k & 3) * 32, 1024, 0),
                         umma_smem_desc_sw128(q_addr + k * 2048, 1024, 0), idesc_acc, (it > 0 || k > 0) ? 1u : 0u);
          umma_commit(&qdo_empty[stage]);
          umma_commit(pds_free);
          if (it == n_iter - 1) umma_commit(acc_full);
        }
        stage = nstage;
        phase = nphase;
      }
    }
  } else {
    // compute warps: thread == key row; the 4 warps {q, q+4, q+8, q+12} share TMEM lane quadrant q and each
    // owns one 32-query column chunk (16 warps in flight per SM hide the ALU / TMEM-load latencies)
    const int quad = warp & 3, part = warp >> 2;
    const int row = quad * 32 + lane;
    const int kj = k0 + row;
    const uint32_t lane_sel = uint32_t(quad * 32) << 16;
    bool key_ok = kj < p.n_k;
    if (key_ok && p.kmask != nullptr) key_ok = (p.kmask[(size_t)batch * p.kb_stride + (kj >> 5)] >> (kj & 31)) & 1u;
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int qb = qb_min + it % q_per_head;
      const int q0 = qb * AB_T;
      [[maybe_unused]] const long long bias_base =
          HAS_BIAS ? (long long)(it / q_per_head) * p.bias_hs + min(kj, p.n_k - 1) : 0;
      mbar_wait(&qdo_full[stage], phase);  // lse / delta staged in smem
      mbar_wait(s_full, it & 1);           // S^T, dP^T ready
      tc_fence_after_sync();
      const float* lse_s = sLse + stage * AB_T;
      const float* del_s = sDelta + stage * AB_T;
      // whole tile below the causal diagonal and inside n_q: only the per-row key flag matters
      const bool tile_full = (q0 + AB_T <= p.n_q) && (!p.causal || k0 + AB_T - 1 <= q0 + off);
      {
        const int c = part;
        uint32_t rs[32], rp[32];
        __syncwarp();
        tmem_ld_32x32b_x32(tmem_ST + lane_sel + c * 32, rs);
        tmem_ld_32x32b_x32(tmem_dPT + lane_sel + c * 32, rp);
        tmem_ld_wait();
        // S^T / dP^T are in registers: let the MMA warp overwrite them with the next iteration's products
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_free);
        // P^T / dS^T staging tiles are still read by the previous iteration's dV / dK MMAs
        if (it > 0) mbar_wait(pds_free, (it - 1) & 1);
        if (!HAS_BIAS && tile_full) {
          // fast path (88 % of the tiles at C3): no per-element predicates; a masked / out-of-range key row is
          // zeroed through its scale factors.  lse / delta come as 16-byte shared loads.
          const float keyf = key_ok ? 1.f : 0.f;
          const float sc = key_ok ? p.scale : 0.f;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float pv[8], dsv[8];
            const float4 l0 = *reinterpret_cast<const float4*>(lse_s + c * 32 + g * 8);
            const float4 l1 = *reinterpret_cast<const float4*>(lse_s + c * 32 + g * 8 + 4);
            const float4 d0 = *reinterpret_cast<const float4*>(del_s + c * 32 + g * 8);
            const float4 d1 = *reinterpret_cast<const float4*>(del_s + c * 32 + g * 8 + 4);
            const float lv[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
            const float dv8[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float pe = ex2_approx(fmaf(__uint_as_float(rs[g * 8 + e]), p.scale_log2, -lv[e]));
              pv[e] = pe * keyf;
              dsv[e] = (pe * sc) * (__uint_as_float(rp[g * 8 + e]) - dv8[e]);
            }
            store_a_chunk(sPT, row, c * 4 + g, pv);
            store_a_chunk(sdST, row, c * 4 + g, dsv);
          }
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float pv[8], dsv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int col = c * 32 + g * 8 + e;
              const int qi = q0 + col;
              const bool ok = key_ok && (tile_full || (qi < p.n_q && (!p.causal || kj <= qi + off)));
              const float s = __uint_as_float(rs[g * 8 + e]);
              float shift = -lse_s[col];
              [[maybe_unused]] long long bidx = 0;
              if constexpr (HAS_BIAS) {  // lanes hold consecutive keys of one query row: coalesced
                bidx = bias_base + (long long)min(qi, p.n_q - 1) * p.bias_rs;
                shift = fmaf(__ldg(p.bias + bidx), LOG2E, shift);
              }
              const float pe = ok ? ex2_approx(fmaf(s, p.scale_log2, shift)) : 0.f;
              pv[e] = pe;
              const float ds = ok ? pe * (__uint_as_float(rp[g * 8 + e]) - del_s[col]) : 0.f;
              dsv[e] = ds * p.scale;
              if constexpr (HAS_BIAS) {
                if (ok && p.dbias != nullptr) atomicAdd(p.dbias + bidx, ds);
              }
            }
            store_a_chunk(sPT, row, c * 4 + g, pv);
            store_a_chunk(sdST, row, c * 4 + g, dsv);
          }
        }
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (++stage == AB_STAGES) { stage = 0; phase ^= 1u; }
    }
    // epilogue: dV, dK from TMEM
    if (n_iter > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after_sync();
    }
    {
      const int which = part >> 1;  // parts 0,1 write dV columns [0,32),[32,64); parts 2,3 write dK
      __nv_bfloat16* dst = (which == 0 ? p.dv : p.dk) +
                           ((size_t)batch * p.n_k + (kj < p.n_k ? kj : 0)) * (which == 0 ? p.lddv : p.lddk);
      {
        const int c = part & 1;
        uint32_t r[32];
        if (n_iter > 0) {
          __syncwarp();
          tmem_ld_32x32b_x32((which == 0 ? tmem_dV : tmem_dK) + lane_sel + c * 32, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) r[e] = 0u;
        }
        if (kj < p.n_k) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 pk;
            pk.x = pack_bf16x2(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1]));
            pk.y = pack_bf16x2(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3]));
            pk.z = pack_bf16x2(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5]));
            pk.w = pack_bf16x2(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7]));
            *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = pk;
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == AB_MMA_WARP) tmem_dealloc(tmem_base, 512);
}

// ================================================================================================
// dQ
// ================================================================================================
constexpr int DQ_SMEM = AB_TILE * (2 + 2 * AB_STAGES + 2) + 256;

template <bool HAS_BIAS>
__global__ void __launch_bounds__(AB_THREADS, 1)
mqa_attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                       const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + AB_TILE;
  uint8_t* sK = sdO + AB_TILE;                 // [stages]
  uint8_t* sV = sK + AB_STAGES * AB_TILE;      // [stages]
  uint8_t* sdS = sV + AB_STAGES * AB_TILE;     // 2 tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + 2 * AB_TILE);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + AB_STAGES;
  uint64_t* s_full = kv_empty + AB_STAGES;
  uint64_t* p_full = s_full + 1;
  uint64_t* acc_full = p_full + 1;
  uint64_t* tmem_free = acc_full + 1;  // compute -> MMA: S / dP of this tile are in registers
  uint64_t* ds_free = tmem_free + 1;   // MMA -> compute: the dQ MMAs of this tile consumed the dS staging tiles
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ds_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qblocks = (p.n_q + AB_T - 1) / AB_T;
  const int qb = n_qblocks - 1 - (int)blockIdx.x;
  const int head = blockIdx.y, batch = blockIdx.z;
  const int q0 = qb * AB_T;
  const int off = p.n_k - p.n_q;
  int kv_end = p.n_k;
  if (p.causal) kv_end = min(p.n_k, q0 + AB_T + off);
  const int n_tiles = kv_end > 0 ? (kv_end + AB_T - 1) / AB_T : 0;

  if (warp == AB_TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    mbar_init(q_full, 1);
    for (int i = 0; i < AB_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, 16);
    mbar_init(acc_full, 1);
    mbar_init(tmem_free, 16);
    mbar_init(ds_free, 1);
    fence_mbar_init();
  }
  if (warp == AB_MMA_WARP) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 128, tmem_dQ = tmem_base + 256;

  if (warp == AB_TMA_WARP) {
    if (n_tiles > 0) {  // whole warp runs the loop; one elected lane issues (see elect_one_sync)
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(q_full, 2 * AB_TILE);
        tma_load_3d(sQ, &tmQ, q_full, head * AB_D, q0, batch);
        tma_load_3d(sdO, &tmdO, q_full, head * AB_D, q0, batch);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&kv_full[stage], 2 * AB_TILE);
          tma_load_3d(sK + stage * AB_TILE, &tmK, &kv_full[stage], 0, j * AB_T, batch);
          tma_load_3d(sV + stage * AB_TILE, &tmV, &kv_full[stage], 0, j * AB_T, batch);
        }
        if (++stage == AB_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == AB_MMA_WARP) {
    if (n_tiles > 0) {  // whole warp runs the loop; one elected lane issues (see elect_one_sync)
      constexpr uint32_t idesc_s = umma_idesc_bf16_f32(AB_T, AB_T, false, false);
      constexpr uint32_t idesc_acc = umma_idesc_bf16_f32(AB_T, AB_D, false, true);
      const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sdO), ds_addr = smem_u32(sdS);
      mbar_wait(q_full, 0);
      auto issue_s = [&](int st) {
        const uint32_t k_addr = smem_u32(sK + st * AB_TILE), v_addr = smem_u32(sV + st * AB_TILE);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < AB_D / 16; ++k)
            umma_bf16_ss(tmem_S, umma_smem_desc_sw128(q_addr + k * 32, 1024, 0),
                         umma_smem_desc_sw128(k_addr + k * 32, 1024, 0), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < AB_D / 16; ++k)
            umma_bf16_ss(tmem_dP, umma_smem_desc_sw128(do_addr + k * 32, 1024, 0),
                         umma_smem_desc_sw128(v_addr + k * 32, 1024, 0), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(s_full);
        }
      };
      int stage = 0;
      uint32_t phase = 0;
      mbar_wait(&kv_full[0], 0);
      tc_fence_after_sync();
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        // software pipeline (see the dK/dV kernel): the next tile's S / dP MMAs are issued as soon as the compute
        // warps hold this tile's S / dP in registers
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == AB_STAGES) { nstage = 0; nphase ^= 1u; }
        if (j + 1 < n_tiles) {
          mbar_wait(tmem_free, j & 1);
          mbar_wait(&kv_full[nstage], nphase);
          tc_fence_after_sync();
          issue_s(nstage);
        }
        mbar_wait(p_full, j & 1);
        tc_fence_after_sync();
        const uint32_t k_addr = smem_u32(sK + stage * AB_TILE);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < AB_T / 16; ++k)
            umma_bf16_ss(tmem_dQ, umma_smem_desc_sw128(ds_addr + (k >> 2) * AB_TILE + (k & 3) * 32, 1024, 0),
                         umma_smem_desc_sw128(k_addr + k * 2048, 1024, 0), idesc_acc, (j > 0 || k > 0) ? 1u : 0u);
          umma_commit(&kv_empty[stage]);
          umma_commit(ds_free);
          if (j == n_tiles - 1) umma_commit(acc_full);
        }
        stage = nstage;
        phase = nphase;
      }
    }
  } else {
    const int quad = warp & 3, part = warp >> 2;
    const int row = quad * 32 + lane;
    const int qi = q0 + row;
    const uint32_t lane_sel = uint32_t(quad * 32) << 16;
    const size_t roff = ((size_t)batch * p.h + head) * p.n_q_pad + qi;
    const float lse = p.lse[roff];          // log2-domain; n_q_pad >= n_qblocks*128: always in bounds
    const float delta = p.delta[roff];
    const int q_limit = p.causal ? qi + off : p.n_k - 1;
    const uint32_t* mrow = p.kmask ? p.kmask + (size_t)batch * p.kb_stride : nullptr;
    [[maybe_unused]] const float* brow =
        HAS_BIAS ? p.bias + (long long)head * p.bias_hs + (long long)min(qi, p.n_q - 1) * p.bias_rs : nullptr;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after_sync();
      const int kbase = j * AB_T;
      const bool tile_full = mrow == nullptr && (q0 + AB_T <= p.n_q) && (kbase + AB_T <= p.n_k) &&
                             (!p.causal || kbase + AB_T - 1 <= q0 + off);
      {
        const int c = part;
        const uint32_t mbits = mrow != nullptr ? __ldg(mrow + j * 4 + c) : 0xFFFFFFFFu;  // this warp's 32 keys
        uint32_t rs[32], rp[32];
        __syncwarp();
        tmem_ld_32x32b_x32(tmem_S + lane_sel + c * 32, rs);
        tmem_ld_32x32b_x32(tmem_dP + lane_sel + c * 32, rp);
        tmem_ld_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_free);
        if (j > 0) mbar_wait(ds_free, (j - 1) & 1);  // previous tile's dQ MMAs still read the dS staging tiles
        if (!HAS_BIAS && tile_full) {
          const float nlse = -lse;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float dsv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float pe = ex2_approx(fmaf(__uint_as_float(rs[g * 8 + e]), p.scale_log2, nlse));
              dsv[e] = (pe * p.scale) * (__uint_as_float(rp[g * 8 + e]) - delta);
            }
            store_a_chunk(sdS, row, c * 4 + g, dsv);
          }
        } else
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float dsv[8];
          [[maybe_unused]] float bb[8];
          if constexpr (HAS_BIAS) {
            const int col = kbase + c * 32 + g * 8;
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
            // bias_rs is a multiple of 4: each float4 is either fully inside the padded row or skipped
            if (col + 3 < p.bias_rs) b0 = __ldg(reinterpret_cast<const float4*>(brow + col));
            if (col + 7 < p.bias_rs) b1 = __ldg(reinterpret_cast<const float4*>(brow + col + 4));
            bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w;
            bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int kj = kbase + c * 32 + g * 8 + e;
            bool ok = tile_full;
            if (!tile_full) {
              ok = qi < p.n_q && kj < p.n_k && kj <= q_limit;
              ok = ok && ((mbits >> (g * 8 + e)) & 1u);
            }
            const float s = __uint_as_float(rs[g * 8 + e]);
            float shift = -lse;
            if constexpr (HAS_BIAS) shift = fmaf(bb[e], LOG2E, shift);
            const float pe = ok ? ex2_approx(fmaf(s, p.scale_log2, shift)) : 0.f;
            dsv[e] = ok ? pe * (__uint_as_float(rp[g * 8 + e]) - delta) * p.scale : 0.f;
          }
          store_a_chunk(sdS, row, c * 4 + g, dsv);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    if (n_tiles > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after_sync();
    }
    if (part < 2) {
      const int c = part;
      uint32_t r[32];
      if (n_tiles > 0) {
        __syncwarp();
        tmem_ld_32x32b_x32(tmem_dQ + lane_sel + c * 32, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = 0u;
      }
      if (qi < p.n_q) {
        __nv_bfloat16* dst = p.dq + ((size_t)batch * p.n_q + qi) * p.lddq + head * AB_D + c * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1]));
          pk.y = pack_bf16x2(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3]));
          pk.z = pack_bf16x2(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5]));
          pk.w = pack_bf16x2(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7]));
          *reinterpret_cast<uint4*>(dst + g * 8) = pk;
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == AB_MMA_WARP) tmem_dealloc(tmem_base, 512);
}

}  // namespace alm

extern "C" int alm_mqa_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, int64_t k_bstride,
                                const void* v, int64_t ldv, int64_t v_bstride, const void* d_o, int64_t lddo,
                                const void* key_mask, const float* lse, const float* delta, int n_q_pad, void* dq,
                                int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, const float* bias,
                                float* dbias, int64_t bias_hstride, int64_t bias_rstride, int b, int h, int n_q,
                                int n_k, int causal, float scale, alm_stream_t stream_) {
  using namespace alm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALM_REQUIRE(q && k && v && d_o && lse && delta && dq && dk && dv, ALM_ERR_ARG);
  ALM_REQUIRE(b > 0 && h > 0 && n_q > 0 && n_k >= n_q, ALM_ERR_ARG);
  ALM_REQUIRE(n_q_pad % AB_T == 0 && n_q_pad >= n_q, ALM_ERR_ARG);
  if (bias != nullptr) {
    ALM_REQUIRE(bias_rstride >= n_k && bias_rstride % 4 == 0 && bias_hstride % 4 == 0, ALM_ERR_ALIGN);
    ALM_REQUIRE((reinterpret_cast<uintptr_t>(bias) & 15u) == 0, ALM_ERR_ALIGN);
  } else {
    ALM_REQUIRE(dbias == nullptr, ALM_ERR_ARG);
  }
  ALM_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && lddo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 &&
                  lddv % 8 == 0 && k_bstride % 8 == 0 && v_bstride % 8 == 0,
              ALM_ERR_ALIGN);
  CUtensorMap tmQ, tmK, tmV, tmdO;
  {
    uint64_t dims[3] = {(uint64_t)h * AB_D, (uint64_t)n_q, (uint64_t)b};
    uint64_t strides[3] = {2, (uint64_t)ldq * 2, (uint64_t)n_q * ldq * 2};
    uint32_t box[3] = {AB_D, AB_T, 1};
    int rc = make_tensor_map(&tmQ, q, 2, 3, dims, strides, box, true);
    if (rc != ALM_OK) return rc;
    strides[1] = (uint64_t)lddo * 2;
    strides[2] = (uint64_t)n_q * lddo * 2;
    rc = make_tensor_map(&tmdO, d_o, 2, 3, dims, strides, box, true);
    if (rc != ALM_OK) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)AB_D, (uint64_t)n_k, (uint64_t)b};
    uint64_t strides[3] = {2, (uint64_t)ldk * 2, (uint64_t)k_bstride * 2};
    uint32_t box[3] = {AB_D, AB_T, 1};
    int rc = make_tensor_map(&tmK, k, 2, 3, dims, strides, box, true);
    if (rc != ALM_OK) return rc;
    strides[1] = (uint64_t)ldv * 2;
    strides[2] = (uint64_t)v_bstride * 2;
    rc = make_tensor_map(&tmV, v, 2, 3, dims, strides, box, true);
    if (rc != ALM_OK) return rc;
  }
  AttnBwdParams p;
  p.lse = lse; p.delta = delta;
  p.kmask = reinterpret_cast<const uint32_t*>(key_mask);
  p.kb_stride = (n_k + 127) / 128 * 4;
  p.dq = (__nv_bfloat16*)dq; p.dk = (__nv_bfloat16*)dk; p.dv = (__nv_bfloat16*)dv;
  p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  p.bias = bias; p.dbias = dbias; p.bias_hs = bias_hstride; p.bias_rs = bias_rstride;
  p.b = b; p.h = h; p.n_q = n_q; p.n_k = n_k; p.n_q_pad = n_q_pad;
  p.causal = causal;
  p.scale = scale;
  p.scale_log2 = scale * LOG2E;
  static bool attr_set = false;
  if (!attr_set) {
    ALM_CUDA_OK(cudaFuncSetAttribute(mqa_attn_bwd_dkv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV_SMEM));
    ALM_CUDA_OK(cudaFuncSetAttribute(mqa_attn_bwd_dq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM));
    ALM_CUDA_OK(cudaFuncSetAttribute(mqa_attn_bwd_dkv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV_SMEM));
    ALM_CUDA_OK(cudaFuncSetAttribute(mqa_attn_bwd_dq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM));
    attr_set = true;
  }
  dim3 grid_kv(((n_k + AB_T - 1) / AB_T) * b);
  if (bias != nullptr)
    mqa_attn_bwd_dkv_kernel<true><<<grid_kv, AB_THREADS, DKV_SMEM, stream>>>(tmQ, tmK, tmV, tmdO, p);
  else
    mqa_attn_bwd_dkv_kernel<false><<<grid_kv, AB_THREADS, DKV_SMEM, stream>>>(tmQ, tmK, tmV, tmdO, p);
  ALM_CHECK_LAUNCH();
  dim3 grid_q((n_q + AB_T - 1) / AB_T, h, b);
  if (bias != nullptr)
    mqa_attn_bwd_dq_kernel<true><<<grid_q, AB_THREADS, DQ_SMEM, stream>>>(tmQ, tmK, tmV, tmdO, p);
  else
    mqa_attn_bwd_dq_kernel<false><<<grid_q, AB_THREADS, DQ_SMEM, stream>>>(tmQ, tmK, tmV, tmdO, p);
  ALM_CHECK_LAUNCH();
  ALM_LAUNCHED(2);
  return ALM_OK;
}
