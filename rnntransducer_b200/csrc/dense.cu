// dense.cu -- dense-logits front-end and gradient of the RNN-T loss (the plain RNNTLoss drop-in:
// `loss(logits[B,T,U1,V], labels, act_lens, label_lens)`, reference model.py:57,74).
//
// Both kernels are HBM-bound streaming passes over the logits (12*V bytes per cell for fwd+bwd,
// SURVEY.md 8(d)): one warp owns kRows consecutive lattice cells, issues the loads of all of them
// before any reduction (bytes in flight), reduces over V with warp shuffles, and writes only
// (lp_blank, lp_label) as one float2 and the log-sum-exp per cell.  The gradient kernel recomputes
// the softmax from logits + lse, so no [B,T,U1,V] probability tensor is ever stored.
// Replaces torchaudio's ReduceMax2D / ReduceLogSumExpGivenMax2D / ComputeLogProbs / ComputeGradients
// and warp-transducer's reduce_max / reduce_exp / compute_grad_kernel.
#include "common.cuh"

namespace rnntb200 {

namespace {

constexpr int kRows = 4;        // cells per warp (independent loads in flight)
constexpr int kWarpsPerCta = 8;

struct CellCoord {
    int b, t, u;
    bool valid;
};

__device__ __forceinline__ CellCoord decode_cell(long long c, long long cells, int T, int U1,
                                                 const int32_t* act_lens, const int32_t* label_lens) {
    CellCoord k;
    k.valid = false;
    k.b = k.t = k.u = 0;
    if (c >= cells) return k;
    const int tu = T * U1;
    k.b = (int)(c / tu);
    const int r = (int)(c - (long long)k.b * tu);
    k.t = r / U1;
    k.u = r - k.t * U1;
    k.valid = k.t < len_T(act_lens, k.b, T) && k.u <= len_U(label_lens, k.b, U1);
    return k;
}

// NV > 0: row kept in NV registers per lane (V <= 32*NV).  NV == 0: generic V, online softmax.
template <typename T, int NV>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
dense_lse_kernel(const T* __restrict__ logits, const int32_t* __restrict__ labels,
                 const int32_t* __restrict__ act_lens, const int32_t* __restrict__ label_lens,
                 long long cells, int T_, int U1, int V, int blank, float2* __restrict__ lp2,
                 float* __restrict__ lse_out) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const long long c0 = warp * kRows;
    if (c0 >= cells) return;

    CellCoord cc[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) cc[r] = decode_cell(c0 + r, cells, T_, U1, act_lens, label_lens);

    float lse[kRows];
    if (NV > 0) {
        float x[kRows][NV > 0 ? NV : 1];
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const T* row = logits + (c0 + r) * V;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const int v = lane + 32 * k;
                x[r][k] = (cc[r].valid && v < V) ? to_f32<T>(row[v]) : -INFINITY;
            }
        }
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            float m = x[r][0];
#pragma unroll
            for (int k = 1; k < NV; ++k) m = fmaxf(m, x[r][k]);
            m = warp_max(m);
            const float m2 = (m == -INFINITY ? 0.f : m) * kLog2e;
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < NV; ++k) s += fast_ex2(fmaf(x[r][k], kLog2e, -m2));
            s = warp_sum(s);
            lse[r] = (m2 + fast_lg2(s)) * kLn2;
        }
    } else {
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            float m = -INFINITY, s = 0.f;
            if (cc[r].valid) {
                const T* row = logits + (c0 + r) * V;
                for (int v = lane; v < V; v += 32) {
                    const float xv = to_f32<T>(row[v]) * kLog2e;
                    const float mn = fmaxf(m, xv);
                    const float ms = mn == -INFINITY ? 0.f : mn;  // all -inf so far: keep s = 0
                    s = s * fast_ex2(m - ms) + fast_ex2(xv - ms);
                    m = mn;
                }
            }
            const float mw = warp_max(m);
            const float mw0 = mw == -INFINITY ? 0.f : mw;
            s = warp_sum(m == -INFINITY ? 0.f : s * fast_ex2(m - mw0));
            lse[r] = (mw0 + fast_lg2(s)) * kLn2;
        }
    }

    if (lane < kRows) {
        // lane r finishes cell r: pick the blank / label logits (L1-resident re-read)
        float my_lse = lse[0];
        CellCoord my = cc[0];
#pragma unroll
        for (int r = 1; r < kRows; ++r)
            if (lane == r) { my_lse = lse[r]; my = cc[r]; }
        if (my.valid) {
            const T* row = logits + (c0 + lane) * V;
            const float lb = fmaxf(to_f32<T>(row[blank]) - my_lse, kNegInf);
            float ll = 0.f;
            if (my.u < len_U(label_lens, my.b, U1)) {
                const int y = label_at(labels, my.b, U1, my.u, V);
                ll = fmaxf(to_f32<T>(row[y]) - my_lse, kNegInf);
            }
            lp2[c0 + lane] = make_float2(lb, ll);
            lse_out[c0 + lane] = my_lse;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
dense_grad_kernel(const T* __restrict__ logits, const int32_t* __restrict__ labels,
                  const int32_t* __restrict__ act_lens, const int32_t* __restrict__ label_lens,
                  long long cells, int T_, int U1, int V, int blank, const float* __restrict__ lse,
                  const int32_t* __restrict__ alpha, const int32_t* __restrict__ beta,
                  const float* __restrict__ grad_costs, T* __restrict__ grad) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const long long c0 = warp * kRows;
    if (c0 >= cells) return;

#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const long long c = c0 + r;
        if (c >= cells) break;
        const CellCoord k = decode_cell(c, cells, T_, U1, act_lens, label_lens);
        T* g = grad + c * V;
        if (!k.valid) {
            for (int v = lane; v < V; v += 32) g[v] = from_f32<T>(0.f);
            continue;
        }
        const T* row = logits + c * V;
        const int Tb = len_T(act_lens, k.b, T_), Ub = len_U(label_lens, k.b, U1);
        const int aq = alpha[c], llq = beta[(long long)k.b * T_ * U1];  // beta(0,0) = log2 P(y|x)
        const float z2 = lse[c] * kLog2e, gc = grad_costs[k.b];
        const float c_all = e16m16_log2_ratio(aq, beta[c], llq) - z2;  // log2(occupancy / partition)
        // corrections at the blank and label columns
        float corr_b = 0.f, corr_l = 0.f;
        int y = -1;
        const float lb2 = to_f32<T>(row[blank]) * kLog2e - z2;
        if (k.t < Tb - 1) corr_b = fast_ex2(e16m16_log2_ratio(aq, beta[c + U1], llq) + lb2);
        else if (k.u == Ub) corr_b = fast_ex2(e16m16_log2_ratio(aq, 0, llq) + lb2);
        if (k.u < Ub) {
            y = label_at(labels, k.b, U1, k.u, V);
            const float ll2 = to_f32<T>(row[y]) * kLog2e - z2;
            corr_l = fast_ex2(e16m16_log2_ratio(aq, beta[c + 1], llq) + ll2);
        }
        for (int v = lane; v < V; v += 32) {
            float gv = fast_ex2(fmaf(to_f32<T>(row[v]), kLog2e, c_all));
            if (v == blank) gv -= corr_b;
            if (v == y) gv -= corr_l;
            g[v] = from_f32<T>(gv * gc);
        }
    }
}

template <typename T>
int launch_lse_t(const T* logits, const int32_t* labels, const int32_t* act_lens,
                 const int32_t* label_lens, int B, int T_, int U1, int V, int blank, float2* lp2,
                 float* lse, cudaStream_t stream) {
    const long long cells = (long long)B * T_ * U1;
    const long long warps = (cells + kRows - 1) / kRows;
    const unsigned grid = (unsigned)((warps + kWarpsPerCta - 1) / kWarpsPerCta);
    const int threads = kWarpsPerCta * 32;
#define RNNT_LSE(NV)                                                                              \
    dense_lse_kernel<T, NV><<<grid, threads, 0, stream>>>(logits, labels, act_lens, label_lens,  \
                                                          cells, T_, U1, V, blank, lp2, lse)
    if (V <= 32) RNNT_LSE(1);
    else if (V <= 64) RNNT_LSE(2);
    else if (V <= 96) RNNT_LSE(3);
    else if (V <= 128) RNNT_LSE(4);
    else RNNT_LSE(0);
#undef RNNT_LSE
    return launch_status();
}

template <typename T>
int launch_grad_t(const T* logits, const int32_t* labels, const int32_t* act_lens,
                  const int32_t* label_lens, int B, int T_, int U1, int V, int blank, const float* lse,
                  const int32_t* alpha, const int32_t* beta, const float* grad_costs,
                  T* grad, cudaStream_t stream) {
    const long long cells = (long long)B * T_ * U1;
    const long long warps = (cells + kRows - 1) / kRows;
    const unsigned grid = (unsigned)((warps + kWarpsPerCta - 1) / kWarpsPerCta);
    dense_grad_kernel<T><<<grid, kWarpsPerCta * 32, 0, stream>>>(
        logits, labels, act_lens, label_lens, cells, T_, U1, V, blank, lse, alpha, beta,
        grad_costs, grad);
    return launch_status();
}

}  // namespace

int launch_dense_lse(const void* logits, int dtype, const int32_t* labels, const int32_t* act_lens,
                     const int32_t* label_lens, int B, int T, int U1, int V, int blank, float2* lp2,
                     float* lse, cudaStream_t stream) {
    if ((long long)B * T * U1 == 0) return RNNTB200_STATUS_SUCCESS;
    switch (dtype) {
        case RNNTB200_F32:
            return launch_lse_t((const float*)logits, labels, act_lens, label_lens, B, T, U1, V, blank, lp2, lse, stream);
        case RNNTB200_F16:
            return launch_lse_t((const __half*)logits, labels, act_lens, label_lens, B, T, U1, V, blank, lp2, lse, stream);
        case RNNTB200_BF16:
            return launch_lse_t((const __nv_bfloat16*)logits, labels, act_lens, label_lens, B, T, U1, V, blank, lp2, lse, stream);
    }
    return RNNTB200_STATUS_INVALID_VALUE;
}

int launch_dense_grad(const void* logits, int dtype, const int32_t* labels, const int32_t* act_lens,
                      const int32_t* label_lens, int B, int T, int U1, int V, int blank,
                      const float* lse, const int32_t* alpha, const int32_t* beta,
                      const float* grad_costs, void* grad_logits, cudaStream_t stream) {
    if ((long long)B * T * U1 == 0) return RNNTB200_STATUS_SUCCESS;
    switch (dtype) {
        case RNNTB200_F32:
            return launch_grad_t((const float*)logits, labels, act_lens, label_lens, B, T, U1, V, blank, lse, alpha, beta, grad_costs, (float*)grad_logits, stream);
        case RNNTB200_F16:
            return launch_grad_t((const __half*)logits, labels, act_lens, label_lens, B, T, U1, V, blank, lse, alpha, beta, grad_costs, (__half*)grad_logits, stream);
        case RNNTB200_BF16:
            return launch_grad_t((const __nv_bfloat16*)logits, labels, act_lens, label_lens, B, T, U1, V, blank, lse, alpha, beta, grad_costs, (__nv_bfloat16*)grad_logits, stream);
    }
    return RNNTB200_STATUS_INVALID_VALUE;
}

}  // namespace rnntb200
