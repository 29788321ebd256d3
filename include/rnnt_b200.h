/*
 * rnnt_b200.h -- C ABI of librnnt_b200.so: the B200-native (sm_100a) fused
 * JointNet + RNN-T loss path.  Plain pointers and sizes only; no torch types.
 *
 * What it replaces in the reference (YooSungHyun/RNNTransducer):
 *   - networks/transducer.py:54-71   JointNet.joint (repeat/cat/GELU/Linear)
 *   - model.py:39,57,74              Warp_RNNTLoss(blank, reduction)(acts, labels, act_lens, label_lens)
 *   - model.py:31,57                 Torch_RNNTLoss(...) (same call, fp16 path)
 * The reference binds those through warp-transducer's C API (rnnt.h:
 * get_workspace_size / compute_rnnt_loss, rnntStatus_t) and torchaudio's
 * torch.ops.torchaudio.rnnt_loss_forward; the entry points below are what a
 * ctypes/cffi binding of this path binds instead (INTEGRATION.md shows it).
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer on the caller's current CUDA device
 *     unless named host_*; the library never allocates device memory and holds
 *     no global state (re-entrant: autograd calls *_bwd from another thread);
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work
 *     and returns (no host synchronisation, CUDA-graph capturable);
 *   - lengths stay on the device: act_lens[B], label_lens[B] are int32 device
 *     arrays and are never read by the host (reference README.md:65 complaint);
 *   - lattice planes are [B, T, U1] row-major, 4 bytes per cell, U1 = max_label_len + 1:
 *       lp2   float2  (log p(blank|t,u), log p(y_{u+1}|t,u)), natural log
 *       lse   float   log-sum-exp of the cell's logits, natural log
 *       alpha, beta   rnntb200_e16m16_t: a 32-bit wide-exponent float ("e16m16"), value =
 *                     (1 + (q & 0xFFFF) / 65536) * 2^(q >> 16)  (arithmetic shift), i.e.
 *                     ln(alpha) = ((q >> 16) + log2(1 + (q & 0xFFFF) / 65536)) * ln(2).
 *                     2^-17 relative precision at any magnitude (|log2| < 32768), so the
 *                     occupancy alpha * beta / P(y|x) is formed from exact integer exponents
 *                     by the gradient kernels; beta[b,0,0] is P(y|x) of utterance b.
 *     entries outside an utterance's (T_b, U_b + 1) box are left untouched by
 *     forward calls and are written as exact zeros in gradient outputs;
 *   - return value: rnntb200_status_t, modelled on warp-transducer's
 *     rnntStatus_t; 0 = success.  rnntb200_status_string() names it.
 *   - There is NO CPU fallback anywhere in this library.
 */
#ifndef RNNT_B200_H_
#define RNNT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RNNTB200_VERSION 100 /* major*100 + minor */

#if defined(__GNUC__)
#define RNNTB200_API __attribute__((visibility("default")))
#else
#define RNNTB200_API
#endif

typedef enum {
    RNNTB200_STATUS_SUCCESS = 0,
    RNNTB200_STATUS_MEMOPS_FAILED = 1,
    RNNTB200_STATUS_INVALID_VALUE = 2,
    RNNTB200_STATUS_EXECUTION_FAILED = 3,
    RNNTB200_STATUS_UNKNOWN_ERROR = 4
} rnntb200_status_t;

/* element type of logits / activations handed in by the caller */
typedef enum { RNNTB200_F32 = 0, RNNTB200_F16 = 1, RNNTB200_BF16 = 2 } rnntb200_dtype_t;

/* alpha / beta plane element: e16m16 wide-exponent float (see conventions above) */
typedef int32_t rnntb200_e16m16_t;

/* joint function.  CONCAT_GELU is the reference's joint (transducer.py:64-69):
 *   logits = fc(gelu_tanh([enc_t ; dec_u])), fc.weight [V, He+Hd].
 * ADD_TANH is the north_star's alternative: logits = fc(tanh(enc_t + dec_u)), fc.weight [V, H]
 * (semantics of torchaudio.models.rnnt._Joiner(activation="tanh")). */
typedef enum { RNNTB200_JOINT_CONCAT_GELU = 0, RNNTB200_JOINT_ADD_TANH = 1 } rnntb200_joint_mode_t;

/* arithmetic of the H x V contraction in ADD_TANH mode */
typedef enum {
    RNNTB200_GEMM_FP32 = 0, /* CUDA-core FFMA, fp32 parity tolerance                  */
    RNNTB200_GEMM_BF16 = 1, /* tcgen05 kind::f16 (bf16 in, fp32 accumulate in TMEM)    */
    RNNTB200_GEMM_TF32X3 = 2 /* tcgen05 kind::tf32, 3-pass split, fp32-class accuracy   */
} rnntb200_gemm_t;

RNNTB200_API int rnntb200_version(void);
RNNTB200_API const char* rnntb200_status_string(int status);

/* ------------------------------------------------------------------------------------------
 * Lattice sweeps (alpha and beta in ONE launch, one CTA per utterance and direction).
 *   lp2   [B,T,U1] float2 = (log p(blank | t,u), log p(y_{u+1} | t,u)), natural log
 *   alpha, beta [B,T,U1] e16m16 out;  costs[B] = -ln beta(0,0) = -log P(y|x), fp32 natural log
 *   ll_alpha[B] optional (may be NULL): ln(alpha(T-1,U) p_blank(T-1,U)), a cross-check of -costs.
 * Replaces warp-transducer compute_alphas/compute_betas and torchaudio's
 * ComputeAlphasBetasCosts.  Requires U1 <= 1024. */
RNNTB200_API int rnntb200_lattice_sweep(const void* lp2, const int32_t* act_lens, const int32_t* label_lens,
                           int B, int T, int U1, rnntb200_e16m16_t* alpha, rnntb200_e16m16_t* beta,
                           float* costs, float* ll_alpha, void* stream);

/* ------------------------------------------------------------------------------------------
 * Dense-logits RNNTLoss (compat path: `loss(logits, labels, act_lens, label_lens)`, model.py:57).
 *   logits [B,T,U1,V] contiguous, dtype in {F32,F16,BF16}; labels [B,U1-1] int32 zero padded.
 * fwd: per-cell log-sum-exp -> lp2, lse; sweeps -> alpha, beta, costs.
 * bwd: grad_logits[b,t,u,v] = grad_costs[b] * d cost_b / d logits (SURVEY.md 8(a) closed form),
 *      same dtype as logits, padded cells written as 0.  Softmax is recomputed from logits + lse. */
RNNTB200_API int rnntb200_loss_dense_fwd(const void* logits, int dtype, const int32_t* labels,
                            const int32_t* act_lens, const int32_t* label_lens, int B, int T,
                            int U1, int V, int blank, float* costs, void* lp2, float* lse,
                            rnntb200_e16m16_t* alpha, rnntb200_e16m16_t* beta, void* stream);

RNNTB200_API int rnntb200_loss_dense_bwd(const void* logits, int dtype, const int32_t* labels,
                            const int32_t* act_lens, const int32_t* label_lens, int B, int T,
                            int U1, int V, int blank, const float* lse, const rnntb200_e16m16_t* alpha,
                            const rnntb200_e16m16_t* beta, const float* grad_costs,
                            void* grad_logits, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused joint + loss, CONCAT_GELU (reference-exact) mode, factorised:
 *   fc(gelu([e_t;d_u])) = P_enc[t,:] + P_dec[u,:],
 *   P_enc = gelu(enc) W[:, :He]^T + bias  [B,T,V],   P_dec = gelu(dec) W[:, He:]^T  [B,U1,V]
 * (two small projections done by the caller or by rnntb200_project_*).  The [B,T,U1,V] logits
 * are never formed.
 * fwd: per-cell V-wide add + log-sum-exp -> lp2, lse; sweeps -> alpha, beta, costs.
 * bwd: d_penc[B,T,V] = sum_u g, d_pdec[B,U1,V] = sum_t g with g = grad_costs[b] * dcost/dlogits;
 *      both fully written (zeros outside the valid box).  `deterministic` != 0 selects the
 *      two-pass reduction through `workspace` (rnntb200_joint_cg_bwd_workspace_bytes) instead of
 *      fp32 atomics for d_pdec. */
/* The two projections themselves, on the tensor cores at fp32 accuracy (bf16 hi/lo split, three
 * tcgen05 MMAs per product, fp32 accumulation):
 *   penc[r,:] = gelu_tanh(enc[r,:]) weight[:, :He]^T + bias,   pdec[r,:] = gelu_tanh(dec[r,:]) weight[:, He:]^T
 * enc [rows_enc, He], dec [rows_dec, Hd], weight [V, He+Hd] (the reference's fc.weight), bias [V].
 * x_dtype (rnntb200_dtype_t) is the element type of enc / dec -- and of d_enc / d_dec in the backward: the
 * AMP mode of the reference's shipped script (scripts/run_train.sh:32 --precision=16) hands the joint fp16
 * activations; they are converted exactly on load and the gradients rounded to nearest on store, the
 * arithmetic in between and every other tensor (weight, bias, penc, pdec, d_weight, d_bias) stay fp32.
 * rnntb200_joint_cg_project_workspace_bytes returns 0 when the shape is not supported (V > 80 or
 * He/Hd not multiples of 64): the caller then uses a library GEMM for this step. */
RNNTB200_API size_t rnntb200_joint_cg_project_workspace_bytes(int V, int He, int Hd);

RNNTB200_API int rnntb200_joint_cg_project(const void* enc, const void* dec, int x_dtype, const float* weight,
                              const float* bias, int rows_enc, int rows_dec, int He, int Hd, int V,
                              float* penc, float* pdec, void* workspace, size_t workspace_bytes,
                              void* stream);

/* Backward of rnntb200_joint_cg_project, same arithmetic: d_enc [rows_enc, He], d_dec [rows_dec, Hd],
 * d_weight [V, He+Hd] and d_bias [V] are fully overwritten.  Supported when the workspace query
 * returns > 0 (V <= 80; He, Hd multiples of 128, <= 512).  workspace_holds_split != 0: `workspace`
 * is the very buffer rnntb200_joint_cg_project filled for the same weight (same layout and size),
 * so the bf16 split of the weight is not redone. */
RNNTB200_API size_t rnntb200_joint_cg_project_bwd_workspace_bytes(int V, int He, int Hd);

RNNTB200_API int rnntb200_joint_cg_project_bwd(const void* enc, const void* dec, int x_dtype, const float* weight,
                                  const float* d_penc, const float* d_pdec, int rows_enc, int rows_dec,
                                  int He, int Hd, int V, void* d_enc, void* d_dec, float* d_weight,
                                  float* d_bias, void* workspace, size_t workspace_bytes,
                                  int workspace_holds_split, void* stream);

/* `factors` (rnntb200_joint_cg_factors_bytes, 16-byte aligned; the query returns 0 and NULL is
 * accepted only when RNNTB200_CG_GENERIC selects the generic per-cell kernels) receives the factor planes of this step -- 2^((P - rowmax) log2 e) of both
 * projections plus per-row scalars, computed once instead of per frame tile.  The forward (or
 * rnntb200_joint_cg_logprobs) writes them; rnntb200_joint_cg_bwd for the same penc / pdec reads
 * them: like lse / alpha / beta they are state saved between the two passes, owned by the caller. */
RNNTB200_API size_t rnntb200_joint_cg_factors_bytes(int B, int T, int U1, int V);

RNNTB200_API int rnntb200_joint_cg_fwd(const float* penc, const float* pdec, const int32_t* labels,
                          const int32_t* act_lens, const int32_t* label_lens, int B, int T, int U1,
                          int V, int blank, float* costs, void* lp2, float* lse,
                          rnntb200_e16m16_t* alpha, rnntb200_e16m16_t* beta, void* factors,
                          size_t factors_bytes, void* stream);

RNNTB200_API size_t rnntb200_joint_cg_bwd_workspace_bytes(int B, int T, int U1, int V, int deterministic);

RNNTB200_API int rnntb200_joint_cg_bwd(const float* penc, const float* pdec, const int32_t* labels,
                          const int32_t* act_lens, const int32_t* label_lens, int B, int T, int U1,
                          int V, int blank, const float* lse, const rnntb200_e16m16_t* alpha,
                          const rnntb200_e16m16_t* beta, const float* grad_costs, float* d_penc,
                          float* d_pdec, int deterministic, void* workspace, size_t workspace_bytes,
                          const void* factors, size_t factors_bytes, void* stream);

/* The same two calls with the reduction over the utterances folded in (what `RNNTLoss(reduction="mean" |
 * "sum")` computes, model.py:39,57): the sweep's last-arriving utterance adds the B costs in index order
 * (bit-reproducible) and writes loss[0] = loss_scale * sum_b costs[b] (loss_scale = 1/B for "mean"); the
 * backward takes the ONE upstream value grad_loss[0] and uses grad_scale * grad_loss[0] for every
 * utterance.  Saves the reduction kernel and the broadcast multiply of a step.  `ticket`: one int32 in
 * device memory, zero before the first call; every call leaves it zero (calls sharing a ticket must be
 * stream-ordered).  costs[B] is still written. */
RNNTB200_API int rnntb200_joint_cg_fwd_loss(const float* penc, const float* pdec, const int32_t* labels,
                               const int32_t* act_lens, const int32_t* label_lens, int B, int T, int U1,
                               int V, int blank, float* costs, void* lp2, float* lse,
                               rnntb200_e16m16_t* alpha, rnntb200_e16m16_t* beta, void* factors,
                               size_t factors_bytes, float* loss, int32_t* ticket, float loss_scale,
                               void* stream);

RNNTB200_API int rnntb200_joint_cg_bwd_loss(const float* penc, const float* pdec, const int32_t* labels,
                               const int32_t* act_lens, const int32_t* label_lens, int B, int T, int U1,
                               int V, int blank, const float* lse, const rnntb200_e16m16_t* alpha,
                               const rnntb200_e16m16_t* beta, const float* grad_loss, float grad_scale,
                               float* d_penc, float* d_pdec, int deterministic, void* workspace,
                               size_t workspace_bytes, const void* factors, size_t factors_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused joint + loss, ADD_TANH mode: logits(t,u,:) = tanh(enc_t + dec_u) W^T + bias, the one
 * dense H x V contraction per lattice cell.  enc [B,T,H], dec [B,U1,H], weight [V,H], bias [V].
 * fwd emits lp2 / lse only; bwd recomputes the logits tile, forms g and feeds it straight into
 * dgrad (g W (1 - z^2)) and wgrad (g^T z): d_enc [B,T,H], d_dec [B,U1,H], d_weight [V,H],
 * d_bias [V]; all four are fully overwritten.  lp2 is the plane the forward wrote.
 * `workspace` (rnntb200_joint_at_workspace_bytes, 16-byte aligned, may be NULL when that is 0)
 * holds the bf16 copy of the weight the TMA streams in RNNTB200_GEMM_BF16 mode. */
RNNTB200_API size_t rnntb200_joint_at_workspace_bytes(int V, int H, int gemm);

RNNTB200_API int rnntb200_joint_at_fwd(const float* enc, const float* dec, const float* weight,
                          const float* bias, int gemm, const int32_t* labels,
                          const int32_t* act_lens, const int32_t* label_lens, int B, int T, int U1,
                          int V, int H, int blank, float* costs, void* lp2, float* lse,
                          rnntb200_e16m16_t* alpha, rnntb200_e16m16_t* beta, void* workspace,
                          size_t workspace_bytes, void* stream);

RNNTB200_API int rnntb200_joint_at_bwd(const float* enc, const float* dec, const float* weight,
                          const float* bias, int gemm, const int32_t* labels,
                          const int32_t* act_lens, const int32_t* label_lens, int B, int T, int U1,
                          int V, int H, int blank, const void* lp2, const float* lse,
                          const rnntb200_e16m16_t* alpha, const rnntb200_e16m16_t* beta,
                          const float* grad_costs, float* d_enc, float* d_dec, float* d_weight,
                          float* d_bias, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stage entry points: the per-cell front-ends alone (no sweep), so that each kernel can be
 * timed and profiled by itself (bench.py's per-kernel roofline) or composed by a caller that
 * batches several front-ends before one sweep.  Same arguments as the *_fwd calls above;
 * outputs: lp2 [B,T,U1] float2 and lse [B,T,U1]. */
RNNTB200_API int rnntb200_dense_logprobs(const void* logits, int dtype, const int32_t* labels,
                            const int32_t* act_lens, const int32_t* label_lens, int B, int T,
                            int U1, int V, int blank, void* lp2, float* lse, void* stream);

RNNTB200_API int rnntb200_joint_cg_logprobs(const float* penc, const float* pdec, const int32_t* labels,
                               const int32_t* act_lens, const int32_t* label_lens, int B, int T,
                               int U1, int V, int blank, void* lp2, float* lse, void* factors,
                               size_t factors_bytes, void* stream);

RNNTB200_API int rnntb200_joint_at_logprobs(const float* enc, const float* dec, const float* weight,
                               const float* bias, int gemm, const int32_t* labels,
                               const int32_t* act_lens, const int32_t* label_lens, int B, int T,
                               int U1, int V, int H, int blank, void* lp2, float* lse, void* workspace,
                               size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * One-shot gradient all-reduce over NVLink peer memory (one process per GPU, one node): what replaces the
 * NCCL all-reduce DDP issues for the fc gradients this path produces (reference train.py:45-48,
 * model.py:59).  Unlike every other entry point these own memory: a rank's buffer must be a cudaMalloc
 * allocation of its own so that its cudaIpc handle names exactly it.
 *   rnntb200_comm_buffer_bytes(max_floats, world)   size of one rank's buffer (one slot per source rank
 *                                            and step parity) for vectors of <= max_floats
 *   rnntb200_comm_alloc / _free              the buffer (zeroed; synchronises the device once)
 *   rnntb200_comm_export(ptr, handle[64])    cudaIpcMemHandle_t of the buffer, to be sent to the peers
 *   rnntb200_comm_import(handle, &ptr) / rnntb200_comm_release(ptr)   map / unmap a peer's buffer
 *   rnntb200_comm_allreduce(peer_ptrs[world] (host array; [rank] = own buffer), rank, world,
 *        segments[n] (host array of device pointers), segment_floats[n], n <= 4, max_floats, scale, stream)
 *     enqueues ONE kernel: segments (in place) <- scale * sum over ranks, summed in rank order (bit-identical
 *     on every rank).  No host synchronisation, no per-step argument changes: CUDA-graph capturable.  Every
 *     rank must call it the same number of times with the same segment sizes; world <= 8. */
RNNTB200_API size_t rnntb200_comm_buffer_bytes(size_t max_floats, int world);
RNNTB200_API int rnntb200_comm_alloc(size_t bytes, void** dev_ptr);
RNNTB200_API int rnntb200_comm_free(void* dev_ptr);
RNNTB200_API int rnntb200_comm_export(void* dev_ptr, unsigned char* handle64);
RNNTB200_API int rnntb200_comm_import(const unsigned char* handle64, void** peer_ptr);
RNNTB200_API int rnntb200_comm_release(void* peer_ptr);
RNNTB200_API int rnntb200_comm_allreduce(void* const* peer_ptrs, int rank, int world, float* const* segments,
                                         const int* segment_floats, int n_segments, size_t max_floats,
                                         float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RNNT_B200_H_ */
