// joint_cg_mm.cu -- the reference-exact joint (concat -> GELU -> Linear) without ANY per-cell
// transcendental work.
//
// logits(t,u,v) = P_enc[t,v] + P_dec[u,v] is a SUM of a t-term and a u-term, so its exponential
// factorises:  exp(logits) = A[t,v] * B[u,v],  A = exp(P_enc - rowmax), B = exp(P_dec - rowmax).
// Hence, per utterance,
//   partition   S[t,u]   = sum_v A[t,v] B[u,v]                       = (A B^T)[t,u]      (K = V)
//   d P_enc[t,v]         = A[t,v] * sum_u C[t,u] B[u,v] - corrections = A .* (C B)        (K = U1)
//   d P_dec[u,v]         = B[u,v] * sum_t C[t,u] A[t,v] - corrections = B .* (C^T A)      (K = T)
// with C[t,u] = grad_cost * occupancy(t,u) / S[t,u] (one exp per CELL, not per cell x vocabulary)
// and the corrections touching only the blank column and the label column of each cell.
// The C*V exponentials of the straightforward evaluation (76 M at B=32,T=400,U=80,V=73 -- the MUFU
// floor of cg_lse_kernel / cg_grad_kernel) become FFMAs of three small batched GEMMs.
//
// Range: A, B are in (0,1].  If the peaks of the two rows do not line up, S can be tiny; cells
// with S < 2^-66 (logit ranges beyond ~45 nats in BOTH rows, never seen with real activations) take
// an exact per-cell path instead (log-sum-exp with the true maximum / explicit V-wide gradient).
//
// Used for V <= 128 (thread tiles keep 16*NC vocabulary columns in registers); larger vocabularies
// run the generic kernels of joint_cg.cu.
#include "common.cuh"

namespace rnntb200 {

namespace {

constexpr float kTinyLog2 = -66.f;  // log2 of the partition threshold below which a cell goes exact

// =================================================================================================
// forward: CTA = (utterance, 32 frames); 4 warps; warp item = 16 t x 16 u, thread = 4 t x 2 u cells
constexpr int kFT = 32;    // frames per CTA
constexpr int kFUC = 64;   // label positions per staged chunk

__device__ __forceinline__ void stage_rows_exp(float* dst, float* rowmax, const float* __restrict__ src,
                                               int n_rows, int n_valid, int V, int Vs) {
    // one warp per row: row max (base 2), then A = 2^(x - max); rows >= n_valid become zeros
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int r = warp; r < n_rows; r += n_warps) {
        float* d = dst + r * Vs;
        if (r >= n_valid) {
            for (int v = lane; v < V; v += 32) d[v] = 0.f;
            if (lane == 0) rowmax[r] = 0.f;
            continue;
        }
        const float* s = src + (size_t)r * V;
        float m = -INFINITY;
        for (int v = lane; v < V; v += 32) m = fmaxf(m, __ldg(s + v));
        m = warp_max(m) * kLog2e;
        for (int v = lane; v < V; v += 32) d[v] = fast_ex2(fmaf(__ldg(s + v), kLog2e, -m));
        if (lane == 0) rowmax[r] = m;
    }
}

__global__ void __launch_bounds__(128)
cg_lse_mm_kernel(const float* __restrict__ penc, const float* __restrict__ pdec,
                 const int32_t* __restrict__ labels, const int32_t* __restrict__ act_lens,
                 const int32_t* __restrict__ label_lens, int T, int U1, int V, int Vs, int blank,
                 float2* __restrict__ lp2, float* __restrict__ lse_out) {
    extern __shared__ float smem[];
    float* As = smem;               // [kFT][Vs]
    float* Bs = As + kFT * Vs;      // [kFUC][Vs]
    float* mA = Bs + kFUC * Vs;     // [kFT]   row maxima, base 2
    float* mB = mA + kFT;           // [kFUC]
    const int b = blockIdx.y, t0 = blockIdx.x * kFT;
    const int Tb = min(__ldg(act_lens + b), T), Ub = min(__ldg(label_lens + b), U1 - 1);
    if (t0 >= Tb) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ty = lane >> 3, tx = lane & 7;

    stage_rows_exp(As, mA, penc + ((size_t)b * T + t0) * V, kFT, min(kFT, Tb - t0), V, Vs);
    for (int u0 = 0; u0 <= Ub; u0 += kFUC) {
        __syncthreads();
        stage_rows_exp(Bs, mB, pdec + ((size_t)b * U1 + u0) * V, kFUC, min(kFUC, Ub + 1 - u0), V, Vs);
        __syncthreads();
        for (int item = warp; item < (kFT / 16) * (kFUC / 16); item += 4) {
            const int rt = (item / (kFUC / 16)) * 16 + 4 * ty;  // first of this thread's 4 frames
            const int ru = (item % (kFUC / 16)) * 16 + 2 * tx;  // first of its 2 label positions
            if (t0 + (item / (kFUC / 16)) * 16 >= Tb || u0 + (item % (kFUC / 16)) * 16 > Ub) continue;
            const float* a = As + rt * Vs;
            const float* bb = Bs + ru * Vs;
            float s[4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i) s[i][0] = s[i][1] = 0.f;
#pragma unroll 4
            for (int v = 0; v < V; ++v) {
                const float b0 = bb[v], b1 = bb[Vs + v];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float av = a[i * Vs + v];
                    s[i][0] = fmaf(av, b0, s[i][0]);
                    s[i][1] = fmaf(av, b1, s[i][1]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int t = t0 + rt + i, u = u0 + ru + k;
                    if (t >= Tb || u > Ub) continue;
                    const float* ar = a + i * Vs;
                    const float* br = bb + k * Vs;
                    const float mm = mA[rt + i] + mB[ru + k];
                    float lgs = fast_lg2(s[i][k]);
                    const int y = u < Ub ? __ldg(labels + (size_t)b * (U1 - 1) + u) : blank;
                    float lb2, ll2;
                    if (lgs < kTinyLog2) {
                        // exact path: the row peaks do not line up, redo this cell in the log domain
                        const float* pe = penc + ((size_t)b * T + t) * V;
                        const float* pd = pdec + ((size_t)b * U1 + u) * V;
                        float mx = -INFINITY;
                        for (int v = 0; v < V; ++v) mx = fmaxf(mx, (pe[v] + pd[v]) * kLog2e);
                        float se = 0.f;
                        for (int v = 0; v < V; ++v) se += fast_ex2((pe[v] + pd[v]) * kLog2e - mx);
                        lgs = mx + fast_lg2(se) - mm;
                        lb2 = (pe[blank] + pd[blank]) * kLog2e - mm - lgs;
                        ll2 = (pe[y] + pd[y]) * kLog2e - mm - lgs;
                    } else {
                        lb2 = fast_lg2(ar[blank] * br[blank]) - lgs;
                        ll2 = fast_lg2(ar[y] * br[y]) - lgs;
                    }
                    const size_t c = ((size_t)b * T + t) * U1 + u;
                    lp2[c] = make_float2(fmaxf(lb2 * kLn2, kNegInf), u < Ub ? fmaxf(ll2 * kLn2, kNegInf) : 0.f);
                    lse_out[c] = (mm + lgs) * kLn2;
                }
        }
    }
}

// =================================================================================================
// backward: CTA = (utterance, 32 frames), 128 threads as 8 (ty) x 16 (tx); thread tile for
// E = C B: frames 4ty..4ty+3 x columns tx + 16c; for D = C^T A: positions ty + 8i x the same columns.
constexpr int kGT2 = 32;   // frames per CTA
constexpr int kGUC2 = 48;  // label positions per chunk (6 per thread row group)
constexpr int kCs = kGUC2 + 1;

template <int NC>
__global__ void __launch_bounds__(128)
cg_grad_mm_kernel(const float* __restrict__ penc, const float* __restrict__ pdec,
                  const int32_t* __restrict__ labels, const int32_t* __restrict__ act_lens,
                  const int32_t* __restrict__ label_lens, int T, int U1, int V, int Vs, int blank,
                  const float* __restrict__ lse, const int32_t* __restrict__ alpha,
                  const int32_t* __restrict__ beta, const float* __restrict__ grad_costs,
                  float* __restrict__ d_penc, float* __restrict__ d_pdec,
                  float* __restrict__ partial /* deterministic slabs or null */) {
    extern __shared__ float smem[];
    float* As = smem;                  // [32][Vs]  A = 2^(P_enc - max)
    float* Bs = As + kGT2 * Vs;        // [48][Vs]  B chunk
    float* Oe = Bs + kGUC2 * Vs;       // [32][Vs]  corrections of d_penc (negative), all chunks
    float* Cs = Oe + kGT2 * Vs;        // [32][49]  C chunk
    float* mA = Cs + kGT2 * kCs;       // [32]
    float* mB = mA + kGT2;             // [48]
    float* ub = mB + kGUC2;            // [48] sum_t corr_blank
    float* ul = ub + kGUC2;            // [48] sum_t corr_label
    __shared__ int ys[kGUC2];
    __shared__ int n_exact;

    const int b = blockIdx.y, tile = blockIdx.x, t0 = tile * kGT2;
    const int Tb = min(__ldg(act_lens + b), T), Ub = min(__ldg(label_lens + b), U1 - 1);
    const int n_tiles = gridDim.x;
    float* slab = partial ? partial + ((size_t)b * n_tiles + tile) * U1 * V : nullptr;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

    if (t0 >= Tb) {  // tile entirely in the padding: exact zeros
        for (int i = tid; i < kGT2 * V; i += 128) {
            const int r = i / V, v = i - r * V;
            if (t0 + r < T) d_penc[((size_t)b * T + t0 + r) * V + v] = 0.f;
        }
        if (slab)
            for (int i = tid; i < U1 * V; i += 128) slab[i] = 0.f;
        return;
    }
    const float gc = grad_costs[b];
    const int llq = beta[(size_t)b * T * U1];  // beta(0,0) = P(y|x), e16m16
    const int rows_t = min(kGT2, Tb - t0);

    stage_rows_exp(As, mA, penc + ((size_t)b * T + t0) * V, kGT2, rows_t, V, Vs);
    for (int i = tid; i < kGT2 * Vs; i += 128) Oe[i] = 0.f;
    if (tid == 0) n_exact = 0;

    float E[4][NC];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < NC; ++c) E[i][c] = 0.f;

    for (int u0 = 0; u0 < U1; u0 += kGUC2) {
        if (!slab && u0 > Ub) break;  // nothing left to add (slabs must be written in full)
        const int rows_u = max(0, min(kGUC2, Ub + 1 - u0));
        __syncthreads();  // previous chunk fully consumed
        stage_rows_exp(Bs, mB, pdec + ((size_t)b * U1 + u0) * V, kGUC2, rows_u, V, Vs);
        if (tid < kGUC2) {
            const int u = u0 + tid;
            ys[tid] = u < Ub ? __ldg(labels + (size_t)b * (U1 - 1) + u) : -1;
            ub[tid] = 0.f;
            ul[tid] = 0.f;
        }
        __syncthreads();
        // per-cell scalars of the (32 x 48) block: C and the two corrections
        for (int i = tid; i < kGT2 * kGUC2; i += 128) {
            const int r = i / kGUC2, uu = i - r * kGUC2;
            const int t = t0 + r, u = u0 + uu;
            float cval = 0.f;
            if (t < Tb && u <= Ub) {
                const size_t c = ((size_t)b * T + t) * U1 + u;
                const int aq = alpha[c];
                const float shift = mA[r] + mB[uu] - lse[c] * kLog2e;  // -log2 S(t,u)
                const float* ar = As + r * Vs;
                const float* br = Bs + uu * Vs;
                const int y = ys[uu];
                const bool exact = shift > -kTinyLog2;
                float pb, pl = 0.f;  // p(blank), p(label) of the cell
                if (!exact) {
                    cval = gc * fast_ex2(e16m16_log2_ratio(aq, beta[c], llq) + shift);
                    pb = ar[blank] * br[blank] * fast_ex2(shift);
                    if (y >= 0) pl = ar[y] * br[y] * fast_ex2(shift);
                } else {
                    atomicAdd(&n_exact, 1);
                    const float* pe = penc + ((size_t)b * T + t) * V;
                    const float* pd = pdec + ((size_t)b * U1 + u) * V;
                    const float z2 = lse[c] * kLog2e;
                    pb = fast_ex2((pe[blank] + pd[blank]) * kLog2e - z2);
                    if (y >= 0) pl = fast_ex2((pe[y] + pd[y]) * kLog2e - z2);
                }
                float cb = 0.f, cl = 0.f;
                if (t < Tb - 1) cb = gc * pb * fast_ex2(e16m16_log2_ratio(aq, beta[c + U1], llq));
                else if (u == Ub) cb = gc * pb * fast_ex2(e16m16_log2_ratio(aq, 0, llq));
                if (y >= 0) cl = gc * pl * fast_ex2(e16m16_log2_ratio(aq, beta[c + 1], llq));
                if (cb != 0.f) { atomicAdd(Oe + r * Vs + blank, -cb); atomicAdd(ub + uu, cb); }
                if (cl != 0.f) { atomicAdd(Oe + r * Vs + y, -cl); atomicAdd(ul + uu, cl); }
            }
            Cs[r * kCs + uu] = cval;
        }
        __syncthreads();
        // E += C B   (K = label positions of the chunk)
        for (int uu = 0; uu < rows_u; ++uu) {
            float bv[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) bv[c] = Bs[uu * Vs + tx + 16 * c];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float cv = Cs[(4 * ty + i) * kCs + uu];
#pragma unroll
                for (int c = 0; c < NC; ++c) E[i][c] = fmaf(cv, bv[c], E[i][c]);
            }
        }
        // D = C^T A   (K = frames of the tile), then d_pdec partial = B .* D - corrections
        float D[6][NC];
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int c = 0; c < NC; ++c) D[i][c] = 0.f;
        for (int r = 0; r < rows_t; ++r) {
            float av[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) av[c] = As[r * Vs + tx + 16 * c];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const float cv = Cs[r * kCs + ty + 8 * i];
#pragma unroll
                for (int c = 0; c < NC; ++c) D[i][c] = fmaf(cv, av[c], D[i][c]);
            }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int uu = ty + 8 * i, u = u0 + uu;
            if (u >= U1) continue;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int v = tx + 16 * c;
                if (v >= V) continue;
                float g = 0.f;
                if (uu < rows_u) {
                    g = Bs[uu * Vs + v] * D[i][c];
                    if (v == blank) g -= ub[uu];
                    if (v == ys[uu]) g -= ul[uu];
                }
                if (slab) slab[(size_t)u * V + v] = g;  // exact-path cells are added below
                else if (uu < rows_u) atomicAdd(d_pdec + ((size_t)b * U1 + u) * V + v, g);
            }
        }
        // exact path (cold): cells whose partition underflows the factorised form
        if (n_exact > 0) {  // uniform: n_exact was final at the barrier above
            __syncthreads();  // slab writes of this chunk are complete
            for (int i = tid; i < kGT2 * kGUC2; i += 128) {
                const int r = i / kGUC2, uu = i - r * kGUC2;
                const int t = t0 + r, u = u0 + uu;
                if (t >= Tb || u > Ub) continue;
                const size_t c = ((size_t)b * T + t) * U1 + u;
                const float z2 = lse[c] * kLog2e;
                if (!(mA[r] + mB[uu] - z2 > -kTinyLog2)) continue;
                const float occ = e16m16_log2_ratio(alpha[c], beta[c], llq) - z2;
                const float* pe = penc + ((size_t)b * T + t) * V;
                const float* pd = pdec + ((size_t)b * U1 + u) * V;
                for (int v = 0; v < V; ++v) {
                    const float g = gc * fast_ex2((pe[v] + pd[v]) * kLog2e + occ);
                    atomicAdd(Oe + r * Vs + v, g);
                    if (slab) atomicAdd(slab + (size_t)u * V + v, g);
                    else atomicAdd(d_pdec + ((size_t)b * U1 + u) * V + v, g);
                }
            }
        }
    }
    __syncthreads();
    // d_penc = A .* E + corrections (blank / label columns, exact-path cells); padded rows are zero
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = 4 * ty + i;
        if (t0 + r >= T) continue;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int v = tx + 16 * c;
            if (v < V) d_penc[((size_t)b * T + t0 + r) * V + v] = fmaf(As[r * Vs + v], E[i][c], Oe[r * Vs + v]);
        }
    }
}

template <int NC>
int launch_grad_mm(const float* penc, const float* pdec, const int32_t* labels, const int32_t* act_lens,
                   const int32_t* label_lens, int B, int T, int U1, int V, int blank, const float* lse,
                   const int32_t* alpha, const int32_t* beta, const float* grad_costs, float* d_penc,
                   float* d_pdec, float* partial, cudaStream_t stream) {
    const int Vs = V | 1;
    const size_t smem = ((size_t)(2 * kGT2 + kGUC2) * Vs + kGT2 * kCs + kGT2 + 3 * kGUC2) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(cg_grad_mm_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return status_from_cuda(e);
    dim3 grid((T + kGT2 - 1) / kGT2, B);
    cg_grad_mm_kernel<NC><<<grid, 128, smem, stream>>>(penc, pdec, labels, act_lens, label_lens, T, U1, V, Vs,
                                                      blank, lse, alpha, beta, grad_costs, d_penc, d_pdec, partial);
    return launch_status();
}

}  // namespace

bool cg_mm_supported(int V) { return V <= 128; }
int cg_mm_tile_rows() { return kGT2; }

int launch_cg_lse_mm(const float* penc, const float* pdec, const int32_t* labels, const int32_t* act_lens,
                     const int32_t* label_lens, int B, int T, int U1, int V, int blank, float2* lp2,
                     float* lse, cudaStream_t stream) {
    const int Vs = V | 1;
    const size_t smem = ((size_t)(kFT + kFUC) * Vs + kFT + kFUC) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(cg_lse_mm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return status_from_cuda(e);
    dim3 grid((T + kFT - 1) / kFT, B);
    cg_lse_mm_kernel<<<grid, 128, smem, stream>>>(penc, pdec, labels, act_lens, label_lens, T, U1, V, Vs, blank,
                                                  lp2, lse);
    return launch_status();
}

int launch_cg_grad_mm(const float* penc, const float* pdec, const int32_t* labels, const int32_t* act_lens,
                      const int32_t* label_lens, int B, int T, int U1, int V, int blank, const float* lse,
                      const int32_t* alpha, const int32_t* beta, const float* grad_costs, float* d_penc,
                      float* d_pdec, float* partial, cudaStream_t stream) {
#define RNNT_MM(NC)                                                                                      \
    return launch_grad_mm<NC>(penc, pdec, labels, act_lens, label_lens, B, T, U1, V, blank, lse, alpha,  \
                              beta, grad_costs, d_penc, d_pdec, partial, stream)
    if (V <= 16) RNNT_MM(1);
    if (V <= 32) RNNT_MM(2);
    if (V <= 48) RNNT_MM(3);
    if (V <= 80) RNNT_MM(5);
    RNNT_MM(8);
#undef RNNT_MM
}

}  // namespace rnntb200
