/*
 * libalm_b200 — C ABI of the B200-native AudioLM hot path.
 *
 * The reference (lucidrains/audiolm-pytorch) has no FFI / plugin registry: its hot path is reached
 * through Python classes that call torch library kernels.  This header is the boundary a maintainer
 * would bind instead (ctypes stub shown in INTEGRATION.md).  Each entry point names the reference
 * call it replaces (file:line under /root/reference/audiolm_pytorch/).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch allocates everything);
 *     the library allocates nothing, never synchronises, and launches on the given stream;
 *   - return value: ALM_OK (0) or a negative alm_status; details go to stderr;
 *   - bf16 tensors are raw uint16 storage (torch.bfloat16), "f32" is IEEE float, ids are int64;
 *   - all row-major; leading dimensions in ELEMENTS.
 */
#ifndef ALM_B200_H_
#define ALM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* alm_stream_t; /* == cudaStream_t */

typedef enum alm_status {
  ALM_OK = 0,
  ALM_ERR_ARG = -1,
  ALM_ERR_ALIGN = -2,
  ALM_ERR_CUDA = -3,
  ALM_ERR_UNSUPPORTED = -4
} alm_status;

/* ---- runtime ------------------------------------------------------------------------------ */
int alm_version(void);
const char* alm_status_string(int code);
unsigned long long alm_launch_count(void); /* kernels launched by this library since the last reset */
void alm_reset_launch_count(void);

/* ---- dense contractions (tcgen05 + TMA) ---------------------------------------------------- */
/*
 * C[b,m,n] (op)= alpha * sum_k A(b,m,k) * B(b,n,k)  [+ bias[n]]        bf16 x bf16 -> fp32 accumulate in TMEM
 *   a_mn = 0: A(m,k) = A[b*strideA + m*lda + k]   ("K-major", e.g. activations x[M,K])
 *   a_mn = 1: A(m,k) = A[b*strideA + k*lda + m]   ("MN-major", e.g. dy^T for weight gradients)
 *   b_mn likewise for B(n,k).   c_fp32: 0 -> bf16 output, 1 -> fp32 output.
 *   acc_mode: 0 overwrite, 1 C += (read-modify-write), 2 C += with fp32 atomics (required when split_k > 1).
 * Replaces every nn.Linear / einsum on the transformer path: audiolm_pytorch.py:255,259 (FFN),
 * :293-294,303 (q/kv/out projections), :621,798 (logit Linear), :972,979,1335,1350,1357 (grouped
 * logit einsums), and their autograd backward (dgrad: b_mn=1, wgrad: a_mn=b_mn=1).
 * Alignment: A/B base 16 B, lda/ldb/strides multiples of 8 elements.
 */
int alm_gemm_bf16(const void* A, int a_mn, int64_t lda, int64_t strideA, const void* B, int b_mn, int64_t ldb,
                  int64_t strideB, void* C, int c_fp32, int64_t ldc, int64_t strideC, int M, int N, int K, int batch,
                  float alpha, const float* bias, int acc_mode, int split_k, alm_stream_t stream);

/* ---- multi-query attention (tcgen05 + TMA, flash-style online softmax) ----------------------- */
/*
 * o[b,i,h*64:(h+1)*64] = softmax_j( q[b,i,h,:]·k[b,j,:] * scale, masked ) · v[b,j,:]
 *   one shared k/v head of width 64 (MQA); key_mask[b,j] (uint8, 1 = attend) optional;
 *   causal: query i sees keys j <= i + (n_k - n_q) (right-aligned, as needed by the KV cache).
 *   lse[b,h,i] (optional) = log-sum-exp of the scaled, masked scores (natural log) for the backward.
 * Replaces Attend.forward / flash_attn (attend.py:69-146) as called by Attention.forward
 * (audiolm_pytorch.py:390).  Fully masked rows produce zeros (the reference's flash path yields NaN).
 * q rows stride ldq (q may be a column slice of a fused qkv buffer); k/v rows stride ldk/ldv, batch
 * strides k_bstride/v_bstride (elements).
 */
int alm_mqa_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, int64_t k_bstride, const void* v,
                     int64_t ldv, int64_t v_bstride, const void* key_mask, void* o, int64_t ldo, float* lse, int b,
                     int h, int n_q, int n_k, int causal, float scale, alm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ALM_B200_H_ */
