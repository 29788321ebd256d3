// joint_cg.cu -- fused joint + log-softmax + gradient for the reference's joint
// (networks/transducer.py:54-71: repeat/cat -> GELU(tanh) -> Linear(2H -> V)).
//
// The activation is elementwise on a concat and the layer after it is linear, so
//     fc(gelu([e_t ; d_u])) = P_enc[t,:] + P_dec[u,:]
// with P_enc = gelu(enc) W[:, :He]^T + bias and P_dec = gelu(dec) W[:, He:]^T.
// These kernels take the two small projections and never form the [B,T,U1,V] logits or any of
// the reference's [B,T,U1,2H] intermediates:
//   cg_lse_kernel   per lattice cell: V-wide add + log-sum-exp -> (lp_blank, lp_label), lse
//   cg_grad_kernel  recomputes the softmax per cell and reduces g = dcost/dlogits straight into
//                   d_penc[t,:] = sum_u g and d_pdec[u,:] = sum_t g.
#include "common.cuh"

namespace rnntb200 {

// joint_cg_mm.cu: the factorised (exp(a+b) = exp(a) exp(b)) kernels used by default
bool cg_mm_supported(int V);
int cg_mm_tile_rows();
CgFactors cg_factors_layout(void* mem, int B, int T, int U1, int V);
int launch_cg_factor_rows(const float* penc, const float* pdec, const int32_t* labels, const int32_t* label_lens,
                          int B, int T, int U1, int V, int blank, const CgFactors& F, cudaStream_t stream);
int launch_cg_lse_mm(const float* penc, const float* pdec, const CgFactors& F, const int32_t* labels,
                     const int32_t* act_lens, const int32_t* label_lens, int B, int T, int U1, int V, int blank,
                     float2* lp2, float* lse, cudaStream_t stream);
int launch_cg_grad_mm(const float* penc, const float* pdec, const CgFactors& F, const int32_t* labels,
                      const int32_t* act_lens, const int32_t* label_lens, int B, int T, int U1, int V, int blank,
                      const float* lse, const int32_t* alpha, const int32_t* beta, GradCosts grad_costs,
                      float* d_penc, float* d_pdec, float* partial, cudaStream_t stream);

namespace {

// ---------------------------------------------------------------------------------------------
// forward: one thread per cell; lanes span u (P_dec rows, odd smem stride -> conflict-free), warps
// span t (P_enc row broadcast inside the warp).
constexpr int kTT = 8;   // t rows per CTA (= warps)
constexpr int kUU = 32;  // u columns per chunk (= lanes)

__global__ void __launch_bounds__(kTT * kUU)
cg_lse_kernel(const float* __restrict__ penc, const float* __restrict__ pdec,
              const int32_t* __restrict__ labels, const int32_t* __restrict__ act_lens,
              const int32_t* __restrict__ label_lens, int T, int U1, int V, int Vs, int blank,
              float2* __restrict__ lp2, float* __restrict__ lse_out) {
    extern __shared__ float smem[];
    float* pe_s = smem;             // [kTT][Vs]
    float* pd_s = smem + kTT * Vs;  // [kUU][Vs]
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * kTT;
    const int Tb = len_T(act_lens, b, T), Ub = len_U(label_lens, b, U1);
    if (t0 >= Tb) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = t0 + warp;

    // stage the P_enc rows of this tile (pre-scaled to the base-2 domain)
    for (int i = threadIdx.x; i < kTT * V; i += kTT * kUU) {
        const int r = i / V, v = i - r * V;
        pe_s[r * Vs + v] = (t0 + r < Tb) ? penc[((size_t)b * T + t0 + r) * V + v] * kLog2e : 0.f;
    }
    for (int u0 = 0; u0 <= Ub; u0 += kUU) {
        __syncthreads();
        for (int i = threadIdx.x; i < kUU * V; i += kTT * kUU) {
            const int r = i / V, v = i - r * V;
            pd_s[r * Vs + v] = (u0 + r <= Ub) ? pdec[((size_t)b * U1 + u0 + r) * V + v] * kLog2e : 0.f;
        }
        __syncthreads();
        const int u = u0 + lane;
        if (t < Tb && u <= Ub) {
            const float* pe = pe_s + warp * Vs;
            const float* pd = pd_s + lane * Vs;
            float m = -INFINITY;
            for (int v = 0; v < V; ++v) m = fmaxf(m, pe[v] + pd[v]);
            const float ms = m == -INFINITY ? 0.f : m;
            float s = 0.f;
            for (int v = 0; v < V; ++v) s += fast_ex2(pe[v] + pd[v] - ms);
            const float lse2 = ms + fast_lg2(s);
            const float lb = fmaxf((pe[blank] + pd[blank] - lse2) * kLn2, kNegInf);
            float ll = 0.f;
            if (u < Ub) {
                const int y = label_at(labels, b, U1, u, V);
                ll = fmaxf((pe[y] + pd[y] - lse2) * kLn2, kNegInf);
            }
            const size_t c = ((size_t)b * T + t) * U1 + u;
            lp2[c] = make_float2(lb, ll);
            lse_out[c] = lse2 * kLn2;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward: one thread per vocabulary column v.  A CTA owns (b, t-tile): d_penc rows of the tile
// are reduced over u in registers and stored once; d_pdec partial sums over the tile's t go to
// global memory with fp32 atomics (or to a per-tile workspace slab in deterministic mode).
constexpr int kGT = 8;    // t rows per CTA
constexpr int kGUC = 64;  // u chunk whose per-cell scalars are staged in smem

struct CellScalars {
    float c_all;   // log2( occupancy / partition ) = (alpha + beta + cost - lse) * log2e
    float corr_b;  // blank-column correction
    float corr_l;  // label-column correction
};

__global__ void __launch_bounds__(256)
cg_grad_kernel(const float* __restrict__ penc, const float* __restrict__ pdec,
               const int32_t* __restrict__ labels, const int32_t* __restrict__ act_lens,
               const int32_t* __restrict__ label_lens, int T, int U1, int V, int blank,
               const float* __restrict__ lse, const int32_t* __restrict__ alpha,
               const int32_t* __restrict__ beta,
               GradCosts grad_costs, float* __restrict__ d_penc,
               float* __restrict__ d_pdec, float* __restrict__ partial /* deterministic slabs or null */) {
    __shared__ CellScalars sc[kGT][kGUC];
    __shared__ int ys[kGUC];
    const int b = blockIdx.y;
    const int tile = blockIdx.x;
    const int t0 = tile * kGT;
    const int Tb = len_T(act_lens, b, T), Ub = len_U(label_lens, b, U1);
    const float gc = grad_costs.at(b);
    const int llq = beta[(size_t)b * T * U1];  // beta(0,0) = log2 P(y|x), e16m16
    const int n_tiles = gridDim.x;
    // deterministic mode: slab [b][tile][U1][V]
    float* slab = partial ? partial + ((size_t)b * n_tiles + tile) * U1 * V : nullptr;

    if (t0 >= Tb) {
        // tile entirely in the padding: its d_penc rows are exact zeros
        for (int i = threadIdx.x; i < kGT * V; i += blockDim.x) {
            const int r = i / V, v = i - r * V;
            if (t0 + r < T) d_penc[((size_t)b * T + t0 + r) * V + v] = 0.f;
        }
        if (slab)
            for (int i = threadIdx.x; i < U1 * V; i += blockDim.x) slab[i] = 0.f;
        return;
    }

    for (int v0 = 0; v0 < V; v0 += blockDim.x) {
        const int v = v0 + threadIdx.x;
        const bool v_on = v < V;
        float pe[kGT], acc_e[kGT];
#pragma unroll
        for (int r = 0; r < kGT; ++r) {
            acc_e[r] = 0.f;
            pe[r] = (v_on && t0 + r < Tb) ? penc[((size_t)b * T + t0 + r) * V + v] * kLog2e : 0.f;
        }
        for (int u0 = 0; u0 < U1; u0 += kGUC) {
            __syncthreads();
            // per-cell scalars of the (tile x chunk) block
            for (int i = threadIdx.x; i < kGT * kGUC; i += blockDim.x) {
                const int r = i / kGUC, uu = i - r * kGUC;
                const int t = t0 + r, u = u0 + uu;
                CellScalars s;
                s.c_all = -INFINITY;
                s.corr_b = 0.f;
                s.corr_l = 0.f;
                if (t < Tb && u <= Ub) {
                    const size_t c = ((size_t)b * T + t) * U1 + u;
                    const int aq = alpha[c];
                    const float z2 = lse[c] * kLog2e;
                    s.c_all = e16m16_log2_ratio(aq, beta[c], llq) - z2;
                    const float* per = penc + ((size_t)b * T + t) * V;
                    const float* pdr = pdec + ((size_t)b * U1 + u) * V;
                    const float lb2 = (per[blank] + pdr[blank]) * kLog2e - z2;
                    if (t < Tb - 1) s.corr_b = fast_ex2(e16m16_log2_ratio(aq, beta[c + U1], llq) + lb2);
                    else if (u == Ub) s.corr_b = fast_ex2(e16m16_log2_ratio(aq, 0, llq) + lb2);
                    if (u < Ub) {
                        const int y = label_at(labels, b, U1, u, V);
                        const float ll2 = (per[y] + pdr[y]) * kLog2e - z2;
                        s.corr_l = fast_ex2(e16m16_log2_ratio(aq, beta[c + 1], llq) + ll2);
                    }
                }
                sc[r][uu] = s;
            }
            for (int i = threadIdx.x; i < kGUC; i += blockDim.x) {
                const int u = u0 + i;
                ys[i] = (u < Ub) ? label_at(labels, b, U1, u, V) : -1;
            }
            __syncthreads();
            if (v_on) {
                const int u_end = min(kGUC, U1 - u0);
                for (int uu = 0; uu < u_end; ++uu) {
                    const int u = u0 + uu;
                    float acc_d = 0.f;
                    if (u <= Ub) {
                        const float pd = pdec[((size_t)b * U1 + u) * V + v] * kLog2e;
                        const bool is_b = v == blank, is_y = v == ys[uu];
#pragma unroll
                        for (int r = 0; r < kGT; ++r) {
                            const CellScalars s = sc[r][uu];
                            float g = fast_ex2(pe[r] + pd + s.c_all);
                            if (is_b) g -= s.corr_b;
                            if (is_y) g -= s.corr_l;
                            acc_e[r] += g;
                            acc_d += g;
                        }
                    }
                    if (slab) slab[(size_t)u * V + v] = acc_d * gc;
                    else if (u <= Ub) atomicAdd(d_pdec + ((size_t)b * U1 + u) * V + v, acc_d * gc);
                }
            }
        }
        if (v_on) {
#pragma unroll
            for (int r = 0; r < kGT; ++r)
                if (t0 + r < T) d_penc[((size_t)b * T + t0 + r) * V + v] = acc_e[r] * gc;
        }
    }
}

// deterministic mode: d_pdec[b,u,v] = sum over tiles of slab[b][tile][u][v], fixed order
__global__ void cg_reduce_slabs_kernel(const float* __restrict__ partial, int n_tiles, size_t uv,
                                       float* __restrict__ d_pdec) {
    const int b = blockIdx.y;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= uv) return;
    const float* p = partial + (size_t)b * n_tiles * uv + i;
    float s = 0.f;
    for (int k = 0; k < n_tiles; ++k) s += p[(size_t)k * uv];
    d_pdec[(size_t)b * uv + i] = s;
}

}  // namespace

int launch_cg_lse(const float* penc, const float* pdec, const int32_t* labels, const int32_t* act_lens,
                  const int32_t* label_lens, int B, int T, int U1, int V, int blank, float2* lp2,
                  float* lse, void* factors, size_t factors_bytes, cudaStream_t stream) {
    if ((long long)B * T * U1 == 0) return RNNTB200_STATUS_SUCCESS;
    if (cg_mm_supported(V)) {
        if (!factors || factors_bytes < cg_factors_bytes(B, T, U1, V) || ((uintptr_t)factors & 15))
            return RNNTB200_STATUS_INVALID_VALUE;
        const CgFactors F = cg_factors_layout(factors, B, T, U1, V);
        const int st = launch_cg_factor_rows(penc, pdec, labels, label_lens, B, T, U1, V, blank, F, stream);
        if (st != RNNTB200_STATUS_SUCCESS) return st;
        return launch_cg_lse_mm(penc, pdec, F, labels, act_lens, label_lens, B, T, U1, V, blank, lp2, lse, stream);
    }
    const int Vs = V | 1;  // odd row stride: lanes reading different rows hit different banks
    const size_t smem = (size_t)(kTT + kUU) * Vs * sizeof(float);
    if (smem > 227 * 1024) return RNNTB200_STATUS_INVALID_VALUE;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(cg_lse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return status_from_cuda(e);
    }
    dim3 grid((T + kTT - 1) / kTT, B);
    cg_lse_kernel<<<grid, kTT * kUU, smem, stream>>>(penc, pdec, labels, act_lens, label_lens, T, U1, V,
                                                     Vs, blank, lp2, lse);
    return launch_status();
}

size_t cg_grad_workspace_bytes(int B, int T, int U1, int V, int deterministic) {
    if (!deterministic) return 0;
    const int rows = cg_mm_supported(V) ? cg_mm_tile_rows() : kGT;
    const size_t n_tiles = (T + rows - 1) / rows;
    return (size_t)B * n_tiles * U1 * V * sizeof(float);
}

int launch_cg_grad(const float* penc, const float* pdec, const int32_t* labels, const int32_t* act_lens,
                   const int32_t* label_lens, int B, int T, int U1, int V, int blank, const float* lse,
                   const int32_t* alpha, const int32_t* beta, GradCosts grad_costs,
                   float* d_penc, float* d_pdec, int deterministic, void* workspace,
                   size_t workspace_bytes, const void* factors, size_t factors_bytes, cudaStream_t stream) {
    if ((long long)B * T * U1 == 0) return RNNTB200_STATUS_SUCCESS;
    const bool mm = cg_mm_supported(V);
    if (mm && (!factors || factors_bytes < cg_factors_bytes(B, T, U1, V) || ((uintptr_t)factors & 15)))
        return RNNTB200_STATUS_INVALID_VALUE;
    const int rows = mm ? cg_mm_tile_rows() : kGT;
    const int n_tiles = (T + rows - 1) / rows;
    float* partial = nullptr;
    if (deterministic) {
        if (!workspace || workspace_bytes < cg_grad_workspace_bytes(B, T, U1, V, 1))
            return RNNTB200_STATUS_INVALID_VALUE;
        partial = (float*)workspace;
    } else {
        cudaError_t e = cudaMemsetAsync(d_pdec, 0, (size_t)B * U1 * V * sizeof(float), stream);
        if (e != cudaSuccess) return RNNTB200_STATUS_MEMOPS_FAILED;
    }
    int st;
    if (mm) {
        // the forward of this step filled the factor planes for the same penc / pdec
        const CgFactors F = cg_factors_layout(const_cast<void*>(factors), B, T, U1, V);
        st = launch_cg_grad_mm(penc, pdec, F, labels, act_lens, label_lens, B, T, U1, V, blank, lse, alpha, beta,
                               grad_costs, d_penc, d_pdec, partial, stream);
    } else {
        const int threads = min(256, ((V + 31) / 32) * 32);
        dim3 grid(n_tiles, B);
        cg_grad_kernel<<<grid, threads, 0, stream>>>(penc, pdec, labels, act_lens, label_lens, T, U1, V,
                                                     blank, lse, alpha, beta, grad_costs, d_penc,
                                                     d_pdec, partial);
        st = launch_status();
    }
    if (st != RNNTB200_STATUS_SUCCESS || !deterministic) return st;
    const size_t uv = (size_t)U1 * V;
    dim3 rgrid((unsigned)((uv + 255) / 256), B);
    cg_reduce_slabs_kernel<<<rgrid, 256, 0, stream>>>(partial, n_tiles, uv, d_pdec);
    return launch_status();
}

}  // namespace rnntb200
