// common.cuh -- shared device helpers for librnnt_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rnnt_b200.h"

namespace rnntb200 {

// Finite stand-in for log(0): keeps every logaddexp branch-free and NaN-free
// (NEG + NEG stays finite, exp2(NEG - x) == 0).
constexpr float kNegInf = -1.0e30f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float fast_rcp(float x) {  // MUFU.RCP, <= 1 ulp; x normal and finite
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// log2(2^a + 2^b), base-2 log domain: 1 MUFU.EX2 + 1 MUFU.LG2 on the dependent chain.
__device__ __forceinline__ float logaddexp2(float a, float b) {
    const float m = fmaxf(a, b);
    const float d = -fabsf(a - b);
    return m + fast_lg2(1.0f + fast_ex2(d));
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// alpha / beta planes hold "e16m16" wide-exponent floats (lattice.cu): value = (1 + lo16 / 65536) *
// 2^(hi16 as signed).  log2 of the occupancy ratio alpha * beta / P(y|x): exponents combine exactly
// in integers, mantissas as one small float; pass bq = 0 for a factor of 1.
__device__ __forceinline__ float e16m16_mant(int q) { return __int_as_float(0x3f800000 | ((q & 0xFFFF) << 7)); }
__device__ __forceinline__ float e16m16_log2_ratio(int aq, int bq, int llq) {
    const int e = (aq >> 16) + (bq >> 16) - (llq >> 16);
    const float m = e16m16_mant(aq) * e16m16_mant(bq) * fast_rcp(e16m16_mant(llq));
    return (float)e + fast_lg2(m);
}

// Lengths and labels are device data nobody validates on the hot path (that would cost a D2H sync):
// every kernel reads them through these, so out-of-range values are clamped IDENTICALLY everywhere
// (act_lens into [1, T], label_lens into [0, U1-1], labels into [0, V-1]) -- never an out-of-bounds
// access, never one kernel disagreeing with another about an utterance's box.
// RNNTLoss(check_inputs=True) raises on such inputs instead (one device sync).
__device__ __forceinline__ int len_T(const int32_t* act_lens, int b, int T) { return min(max(__ldg(act_lens + b), 1), T); }
__device__ __forceinline__ int len_U(const int32_t* label_lens, int b, int U1) { return min(max(__ldg(label_lens + b), 0), U1 - 1); }
__device__ __forceinline__ int label_at(const int32_t* labels, int b, int U1, int u, int V) {
    return min(max(__ldg(labels + (size_t)b * (U1 - 1) + u), 0), V - 1);
}

// Activations (enc / dec and their gradients) may be fp32, fp16 or bf16 (rnntb200_dtype_t): the projection
// kernels read / write them through these -- four consecutive elements at element offset e (e % 4 == 0, rows
// 16-byte aligned for fp32 / 8-byte for the half types), converted exactly on the way in and rounded to
// nearest on the way out.  The arithmetic in between is the same fp32-class arithmetic for every dtype.
__device__ __forceinline__ float4 ldx4(const void* x, int dt, size_t e) {
    if (dt == RNNTB200_F32) return __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(x) + e));
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(x) + e));
    if (dt == RNNTB200_F16) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u),
                       __uint_as_float(r.y << 16), __uint_as_float(r.y & 0xffff0000u));
}
__device__ __forceinline__ void stx4(void* x, int dt, size_t e, float4 v) {
    if (dt == RNNTB200_F32) {
        *reinterpret_cast<float4*>(static_cast<float*>(x) + e) = v;
        return;
    }
    uint2 r;
    if (dt == RNNTB200_F16) {
        const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        r.x = *reinterpret_cast<const uint32_t*>(&a), r.y = *reinterpret_cast<const uint32_t*>(&b);
    } else {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        r.x = *reinterpret_cast<const uint32_t*>(&a), r.y = *reinterpret_cast<const uint32_t*>(&b);
    }
    *reinterpret_cast<uint2*>(static_cast<uint16_t*>(x) + e) = r;
}

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------
// The cfg-2 step is a chain of dependent one-wave kernels, 25-55 us each, and the gap between two of them is a
// microsecond or two per boundary.  The kernels of the chain are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and call pdl_wait() (griddepcontrol.wait) after their own
// prologue, before they touch anything in global memory: the next kernel is then PRE-LAUNCHED and starts the
// moment its predecessor's last CTA exits (-4 us per cfg-2 step, measured).  griddepcontrol.wait returns once
// the preceding grid has completed and its writes are visible; it is a no-op when the kernel was launched
// without the attribute or behind a non-PDL predecessor.  An explicit early trigger
// (griddepcontrol.launch_dependents at kernel entry, -DRNNTB200_PDL_EARLY) was measured 30 us SLOWER per step --
// the next kernel's early-resident CTAs disturb the running one -- and is compiled out.  The attribute is only
// set for small problems (pdl_ok); RNNTB200_PDL=0 switches it off (A/B timing).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
#ifdef RNNTB200_PDL_EARLY  // (the measured-slower variant: dependents become resident while this kernel still runs)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
bool pdl_ok(long long work_rows);  // api.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int status_from_cuda(cudaError_t e) {
    if (e == cudaSuccess) return RNNTB200_STATUS_SUCCESS;
    if (e == cudaErrorInvalidValue || e == cudaErrorInvalidConfiguration) return RNNTB200_STATUS_INVALID_VALUE;
    if (e == cudaErrorMemoryAllocation) return RNNTB200_STATUS_MEMOPS_FAILED;
    return RNNTB200_STATUS_EXECUTION_FAILED;
}

// After a launch: report launch-configuration errors without synchronising.
inline int launch_status() { return status_from_cuda(cudaGetLastError()); }

// ---- internal launchers (defined in the .cu files, called from api.cu) ------------------------
struct LossReduce;
int launch_lattice_sweep(const float2* lp2, const int32_t* act_lens, const int32_t* label_lens, int B,
                         int T, int U1, int32_t* alpha, int32_t* beta, float* costs, float* ll_alpha,
                         cudaStream_t stream, const LossReduce* reduce = nullptr);

int launch_dense_lse(const void* logits, int dtype, const int32_t* labels, const int32_t* act_lens,
                     const int32_t* label_lens, int B, int T, int U1, int V, int blank, float2* lp2,
                     float* lse, cudaStream_t stream);
int launch_dense_grad(const void* logits, int dtype, const int32_t* labels, const int32_t* act_lens,
                      const int32_t* label_lens, int B, int T, int U1, int V, int blank,
                      const float* lse, const int32_t* alpha, const int32_t* beta,
                      const float* grad_costs, void* grad_logits, cudaStream_t stream);

// Factor planes of the factorised concat-GELU joint (joint_cg_mm.cu): written once per
// step by the forward into caller memory (cg_factors_bytes), read by the cell kernels of both passes.
struct CgFactors {
    float* Ea;   // [B*T][Vk]   2^((P_enc - rowmax) log2e), pad columns zero
    float* Eb;   // [B*U1][Vk]  same for P_dec
    uint32_t* Ea2;  // the same two planes as packed (bf16 hi | bf16 lo << 16) pairs, hi + lo = value to
    uint32_t* Eb2;  // 2^-18: the gradient kernel's tensor-core operands (no splitting in its loops)
    float* mA;   // [B*T]   row maxima, base 2
    float* lAb;  // [B*T]   log2 Ea[.][blank] (exact: not taken from the possibly underflowed Ea)
    float* mB;   // [B*U1]
    float* lBb;  // [B*U1]
    float* lBy;  // [B*U1]  log2 Eb[u][y_u], 0 for u >= U_b
    int Vk;      // V rounded up to a multiple of 8
};
size_t cg_factors_bytes(int B, int T, int U1, int V);  // 0: RNNTB200_CG_GENERIC, the generic kernels need no factors

// factors == nullptr is allowed only when cg_factors_bytes(...) == 0
int launch_cg_lse(const float* penc, const float* pdec, const int32_t* labels, const int32_t* act_lens,
                  const int32_t* label_lens, int B, int T, int U1, int V, int blank, float2* lp2,
                  float* lse, void* factors, size_t factors_bytes, cudaStream_t stream);
// Upstream gradient of the per-utterance costs: a [B] vector (stride 1) or ONE value broadcast to every
// utterance (stride 0: the backward of a fused mean / sum, scale = 1/B or 1) -- the latter saves the
// broadcast-multiply launch between the loss reduction's backward and the gradient kernel.
struct GradCosts {
    const float* p;
    int stride;
    float scale;
    __device__ __forceinline__ float at(int b) const { return __ldg(p + (size_t)b * stride) * scale; }
};

// Fused reduction of the costs inside the sweep (last-arriving utterance sums all B costs in index order:
// bit-reproducible): out[0] = scale * sum_b costs[b].  ticket: one int32 in device memory, zero before the
// first launch and left zero by every launch.  out == nullptr: off.
struct LossReduce {
    float* out = nullptr;
    int* ticket = nullptr;
    float scale = 0.f;
};

int launch_cg_grad(const float* penc, const float* pdec, const int32_t* labels, const int32_t* act_lens,
                   const int32_t* label_lens, int B, int T, int U1, int V, int blank, const float* lse,
                   const int32_t* alpha, const int32_t* beta, GradCosts grad_costs,
                   float* d_penc, float* d_pdec, int deterministic, void* workspace,
                   size_t workspace_bytes, const void* factors, size_t factors_bytes, cudaStream_t stream);
size_t cg_grad_workspace_bytes(int B, int T, int U1, int V, int deterministic);

int launch_at_lse(const float* enc, const float* dec, const float* weight, const float* bias, int gemm,
                  const int32_t* labels, const int32_t* act_lens, const int32_t* label_lens, int B,
                  int T, int U1, int V, int H, int blank, float2* lp2, float* lse, void* workspace,
                  size_t workspace_bytes, cudaStream_t stream);
size_t at_workspace_bytes(int V, int H, int gemm);

bool proj_tc_supported(int V, int He, int Hd);
size_t proj_tc_workspace_bytes(int V, int He, int Hd);
bool proj_tc_bwd_supported(int V, int He, int Hd);
size_t proj_tc_bwd_workspace_bytes(int V, int He, int Hd);
int launch_proj_tc_bwd(const void* enc, const void* dec, int x_dtype, const float* weight, const float* d_penc,
                       const float* d_pdec, int rows_enc, int rows_dec, int He, int Hd, int V, void* d_enc,
                       void* d_dec, float* d_weight, float* d_bias, void* workspace, size_t workspace_bytes,
                       int workspace_holds_split, cudaStream_t stream);
int launch_proj_tc(const void* enc, const void* dec, int x_dtype, const float* weight, const float* bias, int rows_enc,
                   int rows_dec, int He, int Hd, int V, float* penc, float* pdec, void* workspace,
                   size_t workspace_bytes, cudaStream_t stream);
int launch_at_grad(const float* enc, const float* dec, const float* weight, const float* bias, int gemm,
                   const int32_t* labels, const int32_t* act_lens, const int32_t* label_lens, int B,
                   int T, int U1, int V, int H, int blank, const float2* lp2, const float* lse,
                   const int32_t* alpha, const int32_t* beta, const float* grad_costs, float* d_enc,
                   float* d_dec, float* d_weight, float* d_bias, void* workspace, size_t workspace_bytes,
                   cudaStream_t stream);

}  // namespace rnntb200
