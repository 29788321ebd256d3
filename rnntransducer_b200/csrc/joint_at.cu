// joint_at.cu -- fused joint + log-softmax + gradient, ADD_TANH mode, CUDA-core (FFMA) kernels.
//
//     logits(t,u,:) = tanh(enc_t + dec_u) W^T + bias          (north_star's joint; semantics of
//                                                              torchaudio.models.rnnt._Joiner("tanh"))
// The [B,T,U1,V] logits and the [B,T,U1,H] activations never reach HBM: a CTA owns a tile of
// lattice cells, builds z = tanh(e_t + d_u) for the tile in shared memory, contracts it with W in
// column tiles of 64, and reduces each tile on the spot
//   forward : online log-sum-exp over the column tiles -> (lp_blank, lp_label), lse
//   backward: g = grad_cost * (softmax * occupancy - corrections) for the column tile, then
//             dz += g W (dgrad), dW += g^T z (wgrad), dbias += sum g, and after the last column
//             tile dpre = dz * (1 - z^2) reduced over u into d_enc and over t into d_dec.
// These are the fp32-exact kernels (RNNTB200_GEMM_FP32) and, with kBf16 = true, an emulation of
// the tensor-core numerics (operands rounded to bf16, fp32 accumulation) that serves as the
// backward of RNNTB200_GEMM_BF16 until the tcgen05 gradient kernel lands; the tcgen05 forward
// lives in joint_at_tc.cu.
#include "common.cuh"

namespace rnntb200 {

int launch_at_lse_tc(const float* enc, const float* dec, const float* weight, const float* bias,
                     const int32_t* labels, const int32_t* act_lens, const int32_t* label_lens, int B,
                     int T, int U1, int V, int H, int blank, float2* lp2, float* lse, void* workspace,
                     size_t workspace_bytes, cudaStream_t stream);  // joint_at_tc.cu
bool at_tc_supported(int V, int H);
size_t at_tc_workspace_bytes(int V, int H);
bool at_tc_bwd_supported(int V, int H);  // joint_at_tc_bwd.cu
int launch_at_grad_tc(const float* enc, const float* dec, const float* weight, const float* bias,
                      const int32_t* labels, const int32_t* act_lens, const int32_t* label_lens, int B, int T,
                      int U1, int V, int H, int blank, const float2* lp2, const float* lse, const int32_t* alpha,
                      const int32_t* beta, const float* grad_costs, float* d_enc, float* d_dec, float* d_weight,
                      float* d_bias, void* workspace, size_t workspace_bytes, cudaStream_t stream);

namespace {

constexpr int kNT = 64;   // vocabulary columns per tile
constexpr int kKT = 32;   // K slab of W staged per step
constexpr int kWs = kKT + 1;

template <bool kBf16>
__device__ __forceinline__ float rnd(float x) {
    return kBf16 ? __bfloat162float(__float2bfloat16_rn(x)) : x;
}

// z tile [kCells][Hs] = tanh(enc[t] + dec[u]), cell r = tt * kUU + uu; rows past the lattice are
// clamped (their results are never written).  Columns H..Hs-1 are zero.
template <bool kBf16, int kTT, int kUU>
__device__ __forceinline__ void build_z(float* Z, int Hs, const float* __restrict__ enc,
                                        const float* __restrict__ dec, int b, int t0, int u0, int T,
                                        int U1, int H) {
    constexpr int kCells = kTT * kUU;
    for (int i = threadIdx.x; i < kCells * Hs; i += blockDim.x) {
        const int r = i / Hs, k = i - r * Hs;
        float z = 0.f;
        if (k < H) {
            const int t = min(t0 + r / kUU, T - 1), u = min(u0 + r % kUU, U1 - 1);
            z = rnd<kBf16>(tanhf(__ldg(enc + ((size_t)b * T + t) * H + k) +
                                 __ldg(dec + ((size_t)b * U1 + u) * H + k)));
        }
        Z[i] = z;
    }
}

// stage W[n0 .. n0+63][k0 .. k0+31] as Ws[col][kk] (zero padded)
template <bool kBf16>
__device__ __forceinline__ void stage_w(float* Ws, const float* __restrict__ W, int n0, int k0, int V,
                                        int H) {
    for (int i = threadIdx.x; i < kNT * kKT; i += blockDim.x) {
        const int c = i / kKT, kk = i - c * kKT;
        const int n = n0 + c, k = k0 + kk;
        Ws[c * kWs + kk] = (n < V && k < H) ? rnd<kBf16>(__ldg(W + (size_t)n * H + k)) : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
// forward: CTA = 8 t x 8 u cells, 256 threads as 16 (rows: 4 cells each) x 16 (cols: tx + 16 j)
template <bool kBf16>
__global__ void __launch_bounds__(256)
at_lse_simt_kernel(const float* __restrict__ enc, const float* __restrict__ dec,
                   const float* __restrict__ W, const float* __restrict__ bias,
                   const int32_t* __restrict__ labels, const int32_t* __restrict__ act_lens,
                   const int32_t* __restrict__ label_lens, int T, int U1, int V, int H, int Hs,
                   int blank, float2* __restrict__ lp2, float* __restrict__ lse_out) {
    constexpr int kTT = 8, kUU = 8, kCells = 64;
    extern __shared__ float smem[];
    float* Z = smem;               // [64][Hs]
    float* Ws = Z + kCells * Hs;   // [64][33]
    float* pick = Ws + kNT * kWs;  // [64][2] logits (base 2) at the blank / label columns
    __shared__ int ylab[kUU];

    const int b = blockIdx.z, t0 = blockIdx.y * kTT, u0 = blockIdx.x * kUU;
    const int Tb = len_T(act_lens, b, T), Ub = len_U(label_lens, b, U1);
    if (t0 >= Tb || u0 > Ub) return;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

    if (threadIdx.x < kUU) {
        const int u = u0 + threadIdx.x;
        ylab[threadIdx.x] = u < Ub ? label_at(labels, b, U1, u, V) : -1;
    }
    if (threadIdx.x < 2 * kCells) pick[threadIdx.x] = 0.f;
    build_z<kBf16, kTT, kUU>(Z, Hs, enc, dec, b, t0, u0, T, U1, H);

    float m_run[4], s_run[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { m_run[i] = -INFINITY; s_run[i] = 0.f; }

    for (int n0 = 0; n0 < V; n0 += kNT) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int k0 = 0; k0 < H; k0 += kKT) {
            __syncthreads();
            stage_w<kBf16>(Ws, W, n0, k0, V, H);
            __syncthreads();
#pragma unroll 8
            for (int kk = 0; kk < kKT; ++kk) {
                float a[4], w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = Z[(ty * 4 + i) * Hs + k0 + kk];
#pragma unroll
                for (int j = 0; j < 4; ++j) w[j] = Ws[(tx + 16 * j) * kWs + kk];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
            }
        }
        // epilogue of this column tile: online log-sum-exp (base 2) per cell row
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = ty * 4 + i;
            float x[4], cmax = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = n0 + tx + 16 * j;
                x[j] = col < V ? (acc[i][j] + __ldg(bias + col)) * kLog2e : -INFINITY;
                cmax = fmaxf(cmax, x[j]);
                if (col == blank) pick[2 * r] = x[j];
                if (col < V && col == ylab[r % kUU]) pick[2 * r + 1] = x[j];
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
            const float m_new = fmaxf(m_run[i], cmax);
            float csum = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) csum += fast_ex2(x[j] - m_new);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
            s_run[i] = s_run[i] * fast_ex2(m_run[i] - m_new) + csum;
            m_run[i] = m_new;
        }
    }
    __syncthreads();
    if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = ty * 4 + i;
            const int t = t0 + r / kUU, u = u0 + r % kUU;
            if (t < Tb && u <= Ub) {
                const float lse2 = m_run[i] + fast_lg2(s_run[i]);
                const size_t c = ((size_t)b * T + t) * U1 + u;
                const float lb = fmaxf((pick[2 * r] - lse2) * kLn2, kNegInf);
                const float ll = u < Ub ? fmaxf((pick[2 * r + 1] - lse2) * kLn2, kNegInf) : 0.f;
                lp2[c] = make_float2(lb, ll);
                lse_out[c] = lse2 * kLn2;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward: CTA = 4 t x 8 u cells, 256 threads.  Requires H <= 512 (dz lives in registers:
// thread owns columns h = tid and tid + 256 for all 32 cells).
struct GradCell {
    float c_all;   // log2(occupancy / partition)
    float corr_b;  // blank-column correction
    float corr_l;  // label-column correction
    int y;         // label column or -1
};

template <bool kBf16>
__global__ void __launch_bounds__(256)
at_grad_simt_kernel(const float* __restrict__ enc, const float* __restrict__ dec,
                    const float* __restrict__ W, const float* __restrict__ bias,
                    const int32_t* __restrict__ labels, const int32_t* __restrict__ act_lens,
                    const int32_t* __restrict__ label_lens, int T, int U1, int V, int H, int Hs,
                    int blank, const float2* __restrict__ lp2, const float* __restrict__ lse,
                    const int32_t* __restrict__ alpha, const int32_t* __restrict__ beta,
                    const float* __restrict__ grad_costs, float* __restrict__ d_enc,
                    float* __restrict__ d_dec, float* __restrict__ d_w, float* __restrict__ d_b) {
    constexpr int kTT = 4, kUU = 8, kCells = 32, kGs = kNT + 1;
    extern __shared__ float smem[];
    float* Z = smem;              // [32][Hs]
    float* Ws = Z + kCells * Hs;  // [64][33]
    float* G = Ws + kNT * kWs;    // [32][65]
    __shared__ GradCell gc_s[kCells];

    const int b = blockIdx.z, t0 = blockIdx.y * kTT, u0 = blockIdx.x * kUU;
    const int Tb = len_T(act_lens, b, T), Ub = len_U(label_lens, b, U1);
    if (t0 >= Tb || u0 > Ub) return;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;  // rows ty*2 .. +1, cols tx + 16 j
    const float gscale = grad_costs[b];

    if (tid < kCells) {
        const int t = t0 + tid / kUU, u = u0 + tid % kUU;
        GradCell g{-INFINITY, 0.f, 0.f, -1};
        if (t < Tb && u <= Ub) {
            const size_t c = ((size_t)b * T + t) * U1 + u;
            const int aq = alpha[c], llq = beta[(size_t)b * T * U1];
            const float2 lp = lp2[c];
            g.c_all = e16m16_log2_ratio(aq, beta[c], llq) - lse[c] * kLog2e;
            if (t < Tb - 1) g.corr_b = fast_ex2(e16m16_log2_ratio(aq, beta[c + U1], llq) + lp.x * kLog2e);
            else if (u == Ub) g.corr_b = fast_ex2(e16m16_log2_ratio(aq, 0, llq) + lp.x * kLog2e);
            if (u < Ub) {
                g.y = label_at(labels, b, U1, u, V);
                g.corr_l = fast_ex2(e16m16_log2_ratio(aq, beta[c + 1], llq) + lp.y * kLog2e);
            }
        }
        gc_s[tid] = g;
    }
    build_z<kBf16, kTT, kUU>(Z, Hs, enc, dec, b, t0, u0, T, U1, H);

    float dz[kCells][2];
#pragma unroll
    for (int r = 0; r < kCells; ++r) dz[r][0] = dz[r][1] = 0.f;
    const int h0 = tid, h1 = tid + 256;

    for (int n0 = 0; n0 < V; n0 += kNT) {
        // (1) logits of the column tile, recomputed exactly as the forward did
        float acc[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int k0 = 0; k0 < H; k0 += kKT) {
            __syncthreads();
            stage_w<kBf16>(Ws, W, n0, k0, V, H);
            __syncthreads();
#pragma unroll 8
            for (int kk = 0; kk < kKT; ++kk) {
                float a[2], w[4];
#pragma unroll
                for (int i = 0; i < 2; ++i) a[i] = Z[(ty * 2 + i) * Hs + k0 + kk];
#pragma unroll
                for (int j = 0; j < 4; ++j) w[j] = Ws[(tx + 16 * j) * kWs + kk];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
            }
        }
        // (2) g = grad_cost * d cost / d logit for the tile
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = ty * 2 + i;
            const GradCell g = gc_s[r];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int cl = tx + 16 * j, col = n0 + cl;
                float gv = 0.f;
                if (col < V) {
                    gv = fast_ex2((acc[i][j] + __ldg(bias + col)) * kLog2e + g.c_all);
                    if (col == blank) gv -= g.corr_b;
                    if (col == g.y) gv -= g.corr_l;
                    gv *= gscale;
                }
                G[r * kGs + cl] = gv;
            }
        }
        __syncthreads();
        // (3) dbias, dgrad (dz += g W) and wgrad (dW += g^T z) for this column tile
        if (tid < kNT && n0 + tid < V) {
            float s = 0.f;
#pragma unroll 8
            for (int r = 0; r < kCells; ++r) s += G[r * kGs + tid];
            atomicAdd(d_b + n0 + tid, s);
        }
        const int nv = min(kNT, V - n0);
        for (int v = 0; v < nv; ++v) {
            const float* wrow = W + (size_t)(n0 + v) * H;
            const float w0 = h0 < H ? rnd<kBf16>(__ldg(wrow + h0)) : 0.f;
            const float w1 = h1 < H ? rnd<kBf16>(__ldg(wrow + h1)) : 0.f;
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int r = 0; r < kCells; ++r) {
                const float gv = G[r * kGs + v];
                dz[r][0] = fmaf(gv, w0, dz[r][0]);
                dz[r][1] = fmaf(gv, w1, dz[r][1]);
                s0 = fmaf(gv, Z[r * Hs + h0], s0);
                if (h1 < Hs) s1 = fmaf(gv, Z[r * Hs + h1], s1);
            }
            if (h0 < H) atomicAdd(d_w + (size_t)(n0 + v) * H + h0, s0);
            if (h1 < H) atomicAdd(d_w + (size_t)(n0 + v) * H + h1, s1);
        }
    }
    // (4) dpre = dz * (1 - z^2); d_enc[t] += sum_u dpre, d_dec[u] += sum_t dpre
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int h = q == 0 ? h0 : h1;
        if (h >= H) continue;
        float du[kUU];
#pragma unroll
        for (int uu = 0; uu < kUU; ++uu) du[uu] = 0.f;
#pragma unroll
        for (int tt = 0; tt < kTT; ++tt) {
            float de = 0.f;
#pragma unroll
            for (int uu = 0; uu < kUU; ++uu) {
                const int r = tt * kUU + uu;
                const float z = Z[r * Hs + h];
                const float dp = dz[r][q] * (1.f - z * z);
                de += dp;
                du[uu] += dp;
            }
            if (t0 + tt < Tb) atomicAdd(d_enc + ((size_t)b * T + t0 + tt) * H + h, de);
        }
#pragma unroll
        for (int uu = 0; uu < kUU; ++uu)
            if (u0 + uu <= Ub) atomicAdd(d_dec + ((size_t)b * U1 + u0 + uu) * H + h, du[uu]);
    }
}

template <typename K>
int set_smem(K kernel, size_t smem) {
    if (smem > 227 * 1024) return RNNTB200_STATUS_INVALID_VALUE;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return status_from_cuda(e);
    }
    return RNNTB200_STATUS_SUCCESS;
}

}  // namespace

size_t at_workspace_bytes(int V, int H, int gemm) {
    return (gemm == RNNTB200_GEMM_BF16 && at_tc_supported(V, H)) ? at_tc_workspace_bytes(V, H) : 0;
}

int launch_at_lse(const float* enc, const float* dec, const float* weight, const float* bias, int gemm,
                  const int32_t* labels, const int32_t* act_lens, const int32_t* label_lens, int B,
                  int T, int U1, int V, int H, int blank, float2* lp2, float* lse, void* workspace,
                  size_t workspace_bytes, cudaStream_t stream) {
    if ((long long)B * T * U1 == 0) return RNNTB200_STATUS_SUCCESS;
    if (gemm == RNNTB200_GEMM_TF32X3) return RNNTB200_STATUS_INVALID_VALUE;  // reserved
    if (gemm == RNNTB200_GEMM_BF16 && at_tc_supported(V, H))
        return launch_at_lse_tc(enc, dec, weight, bias, labels, act_lens, label_lens, B, T, U1, V, H,
                                blank, lp2, lse, workspace, workspace_bytes, stream);
    const int Hs = (H + kKT - 1) / kKT * kKT + 1;  // odd stride; columns H..Hs-1 are zero
    const size_t smem = ((size_t)64 * Hs + kNT * kWs + 128) * sizeof(float);
    dim3 grid((U1 + 7) / 8, (T + 7) / 8, B);
    int st;
    if (gemm == RNNTB200_GEMM_BF16) {
        if ((st = set_smem(at_lse_simt_kernel<true>, smem)) != 0) return st;
        at_lse_simt_kernel<true><<<grid, 256, smem, stream>>>(enc, dec, weight, bias, labels, act_lens,
                                                              label_lens, T, U1, V, H, Hs, blank, lp2, lse);
    } else {
        if ((st = set_smem(at_lse_simt_kernel<false>, smem)) != 0) return st;
        at_lse_simt_kernel<false><<<grid, 256, smem, stream>>>(enc, dec, weight, bias, labels, act_lens,
                                                               label_lens, T, U1, V, H, Hs, blank, lp2, lse);
    }
    return launch_status();
}

int launch_at_grad(const float* enc, const float* dec, const float* weight, const float* bias, int gemm,
                   const int32_t* labels, const int32_t* act_lens, const int32_t* label_lens, int B,
                   int T, int U1, int V, int H, int blank, const float2* lp2, const float* lse,
                   const int32_t* alpha, const int32_t* beta, const float* grad_costs, float* d_enc,
                   float* d_dec, float* d_weight, float* d_bias, void* workspace, size_t workspace_bytes,
                   cudaStream_t stream) {
    if (gemm == RNNTB200_GEMM_TF32X3) return RNNTB200_STATUS_INVALID_VALUE;  // reserved
    const bool tc = gemm == RNNTB200_GEMM_BF16 && at_tc_bwd_supported(V, H);
    if (!tc && H > 512) return RNNTB200_STATUS_INVALID_VALUE;  // dz register tile (see kernel)
    // all four outputs are accumulated with fp32 atomics: clear them first (also the B == 0 case)
    if (cudaMemsetAsync(d_enc, 0, (size_t)B * T * H * sizeof(float), stream) != cudaSuccess ||
        cudaMemsetAsync(d_dec, 0, (size_t)B * U1 * H * sizeof(float), stream) != cudaSuccess ||
        cudaMemsetAsync(d_weight, 0, (size_t)V * H * sizeof(float), stream) != cudaSuccess ||
        cudaMemsetAsync(d_bias, 0, (size_t)V * sizeof(float), stream) != cudaSuccess)
        return RNNTB200_STATUS_MEMOPS_FAILED;
    if ((long long)B * T * U1 == 0) return RNNTB200_STATUS_SUCCESS;
    if (tc)
        return launch_at_grad_tc(enc, dec, weight, bias, labels, act_lens, label_lens, B, T, U1, V, H, blank, lp2,
                                 lse, alpha, beta, grad_costs, d_enc, d_dec, d_weight, d_bias, workspace,
                                 workspace_bytes, stream);
    const int Hs = (H + kKT - 1) / kKT * kKT + 1;
    const size_t smem = ((size_t)32 * Hs + kNT * kWs + 32 * (kNT + 1)) * sizeof(float);
    dim3 grid((U1 + 7) / 8, (T + 3) / 4, B);
    int st;
    if (gemm == RNNTB200_GEMM_BF16) {
        if ((st = set_smem(at_grad_simt_kernel<true>, smem)) != 0) return st;
        at_grad_simt_kernel<true><<<grid, 256, smem, stream>>>(
            enc, dec, weight, bias, labels, act_lens, label_lens, T, U1, V, H, Hs, blank, lp2, lse, alpha,
            beta, grad_costs, d_enc, d_dec, d_weight, d_bias);
    } else {
        if ((st = set_smem(at_grad_simt_kernel<false>, smem)) != 0) return st;
        at_grad_simt_kernel<false><<<grid, 256, smem, stream>>>(
            enc, dec, weight, bias, labels, act_lens, label_lens, T, U1, V, H, Hs, blank, lp2, lse, alpha,
            beta, grad_costs, d_enc, d_dec, d_weight, d_bias);
    }
    return launch_status();
}

}  // namespace rnntb200
