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

__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gmem_src) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(gmem_src) : "memory");
}

// Stage n_rows rows of src ([.., V] row-major) as E = 2^((x - rowmax) log2e).  The raw rows go to
// shared memory with cp.async (fire-and-forget: every load of the tile is in flight at once, no
// register dependency), then one warp per row takes the row maximum and exponentiates in place.
// Rows >= n_valid become zeros.  Optionally records the normalised base-2 log at column `col` of
// every row (the blank column) and at a per-row column cols[r] (the row's label): those can
// underflow in E although they are representable -- and may lie on the best path.
// Contains two block barriers.
__device__ __forceinline__ void stage_rows_exp(float* dst, float* rowmax, float* log_at_col, int col,
                                               const float* __restrict__ src, int n_rows, int n_valid,
                                               int V, int Vs, float* log_at_cols = nullptr,
                                               const int* cols = nullptr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int r = warp; r < n_valid; r += n_warps)
        for (int v = lane; v < V; v += 32) cp_async_4(dst + r * Vs + v, src + (size_t)r * V + v);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int r = n_valid + warp; r < n_rows; r += n_warps) {
        for (int v = lane; v < V; v += 32) dst[r * Vs + v] = 0.f;
        if (lane == 0) {
            rowmax[r] = 0.f;
            if (log_at_col) log_at_col[r] = 0.f;
            if (log_at_cols) log_at_cols[r] = 0.f;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int r = warp; r < n_valid; r += n_warps) {
        float* d = dst + r * Vs;
        float m = -INFINITY;
        for (int v = lane; v < V; v += 32) m = fmaxf(m, d[v]);
        m = warp_max(m);
        if (lane == 0) {
            rowmax[r] = m * kLog2e;
            if (log_at_col) log_at_col[r] = (d[col] - m) * kLog2e;
            if (log_at_cols) log_at_cols[r] = cols[r] >= 0 ? (d[cols[r]] - m) * kLog2e : 0.f;
        }
        __syncwarp();
        for (int v = lane; v < V; v += 32) d[v] = fast_ex2((d[v] - m) * kLog2e);
    }
    __syncthreads();
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
    float* lAb = mB + kFUC;         // [kFT]   log2 A[t][blank]
    float* lBb = lAb + kFT;         // [kFUC]  log2 B[u][blank]
    float* lBy = lBb + kFUC;        // [kFUC]  log2 B[u][y_u]
    __shared__ int ys[kFUC];
    const int b = blockIdx.y, t0 = blockIdx.x * kFT;
    const int Tb = min(__ldg(act_lens + b), T), Ub = min(__ldg(label_lens + b), U1 - 1);
    if (t0 >= Tb) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ty = lane >> 3, tx = lane & 7;

    stage_rows_exp(As, mA, lAb, blank, penc + ((size_t)b * T + t0) * V, kFT, min(kFT, Tb - t0), V, Vs);
    for (int u0 = 0; u0 <= Ub; u0 += kFUC) {
        if (u0 > 0) __syncthreads();  // every warp is done with the previous chunk
        if (threadIdx.x < kFUC) {
            const int u = u0 + threadIdx.x;
            ys[threadIdx.x] = u < Ub ? __ldg(labels + (size_t)b * (U1 - 1) + u) : -1;
        }
        __syncthreads();
        stage_rows_exp(Bs, mB, lBb, blank, pdec + ((size_t)b * U1 + u0) * V, kFUC, min(kFUC, Ub + 1 - u0), V, Vs,
                       lBy, ys);
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
                    const float mm = mA[rt + i] + mB[ru + k];
                    const float* pe = penc + ((size_t)b * T + t) * V;
                    const float* pd = pdec + ((size_t)b * U1 + u) * V;
                    float lgs = fast_lg2(s[i][k]);
                    const int y = u < Ub ? ys[ru + k] : blank;
                    float lb2, ll2;
                    if (lgs < kTinyLog2) {
                        // exact path: the row peaks do not line up, redo this cell in the log domain
                        float mx = -INFINITY;
                        for (int v = 0; v < V; ++v) mx = fmaxf(mx, (pe[v] + pd[v]) * kLog2e);
                        float se = 0.f;
                        for (int v = 0; v < V; ++v) se += fast_ex2((pe[v] + pd[v]) * kLog2e - mx);
                        lgs = mx + fast_lg2(se) - mm;
                    }
                    // blank / label log-probs straight in the log domain (A * B would underflow
                    // below 2^-126 although such steps are representable -- and may be on the path)
                    lb2 = lAb[rt + i] + lBb[ru + k] - lgs;
                    const float ay = a[i * Vs + y];  // A[t][y]: its log unless it underflowed
                    const float lay = ay > 1e-30f ? fast_lg2(ay) : __ldg(pe + y) * kLog2e - mA[rt + i];
                    ll2 = lay + lBy[ru + k] - lgs;
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
// The blank / label corrections ride along as rank-1 style terms with a fixed summation order, so
// the only order-dependent arithmetic is the cross-tile accumulation of d_pdec (fp32 atomics, or
// per-tile slabs + a fixed-order reduction in deterministic mode) and the cold exact path.
constexpr int kGT2 = 32;   // frames per CTA
constexpr int kGUC2 = 48;  // label positions per chunk (6 per thread row group)
constexpr int kCs = kGUC2 + 1;
constexpr int kCellsPerThread = kGT2 * kGUC2 / 128;  // 12
constexpr int kBatch = 4;  // cells whose global loads are in flight together

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
    float* Xs = Bs + kGUC2 * Vs;       // [32][Vs]  corrections of d_penc: -(blank, label terms) (+ exact path)
    float* Cs = Xs + kGT2 * Vs;        // [32][49]  C chunk
    float* CBs = Cs + kGT2 * kCs;      // [32][49]  blank corrections
    float* CLs = CBs + kGT2 * kCs;     // [32][49]  label corrections
    float* mA = CLs + kGT2 * kCs;      // [32] row maxima (base 2)
    float* mB = mA + kGT2;             // [48]
    float* lAb = mB + kGUC2;           // [32] log2 A[t][blank]
    float* lBb = lAb + kGT2;           // [48] log2 B[u][blank]
    float* ub = lBb + kGUC2;           // [48] sum_t corr_blank
    float* ul = ub + kGUC2;            // [48] sum_t corr_label
    float* rb = ul + kGUC2;            // [32] sum_u corr_blank (all chunks)
    float* lBy = rb + kGT2;            // [48] log2 B[u][y_u]
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

    for (int i = tid; i < kGT2 * Vs; i += 128) Xs[i] = 0.f;
    if (tid == 0) n_exact = 0;
    stage_rows_exp(As, mA, lAb, blank, penc + ((size_t)b * T + t0) * V, kGT2, rows_t, V, Vs);

    float E[4][NC];  // (C B)[t][v]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < NC; ++c) E[i][c] = 0.f;

    for (int u0 = 0; u0 < U1; u0 += kGUC2) {
        if (!slab && u0 > Ub) break;  // nothing left to add (slabs must be written in full)
        const int rows_u = max(0, min(kGUC2, Ub + 1 - u0));
        if (u0 > 0) __syncthreads();  // previous chunk fully consumed
        if (tid < kGUC2) {
            const int u = u0 + tid;
            ys[tid] = u < Ub ? __ldg(labels + (size_t)b * (U1 - 1) + u) : -1;
        }
        __syncthreads();  // ys visible to the staging warps
        stage_rows_exp(Bs, mB, lBb, blank, pdec + ((size_t)b * U1 + u0) * V, kGUC2, rows_u, V, Vs, lBy, ys);

        // per-cell scalars of the (32 x 48) block: C and the two corrections.  The global loads of
        // kBatch cells are issued together before any of them is used.
#pragma unroll 1
        for (int q0 = 0; q0 < kCellsPerThread; q0 += kBatch) {
            int aq[kBatch], bq[kBatch], bdn[kBatch], brt[kBatch];
            float z2[kBatch], pey[kBatch];
#pragma unroll
            for (int q = 0; q < kBatch; ++q) {
                const int i = tid + 128 * (q0 + q), r = i / kGUC2, uu = i - r * kGUC2;
                const int t = t0 + r, u = u0 + uu;
                aq[q] = bq[q] = bdn[q] = brt[q] = 0;
                z2[q] = pey[q] = 0.f;
                if (t < Tb && u <= Ub) {
                    const size_t c = ((size_t)b * T + t) * U1 + u;
                    aq[q] = alpha[c];
                    bq[q] = beta[c];
                    if (t < Tb - 1) bdn[q] = beta[c + U1];
                    if (u < Ub) {
                        brt[q] = beta[c + 1];
                        const float ay = As[r * Vs + ys[uu]];  // log2 A[t][y_u]; gather if it underflowed
                        pey[q] = ay > 1e-30f ? fast_lg2(ay)
                                             : __ldg(penc + ((size_t)b * T + t) * V + ys[uu]) * kLog2e - mA[r];
                    }
                    z2[q] = lse[c] * kLog2e;
                }
            }
#pragma unroll
            for (int q = 0; q < kBatch; ++q) {
                const int i = tid + 128 * (q0 + q), r = i / kGUC2, uu = i - r * kGUC2;
                const int t = t0 + r, u = u0 + uu;
                float cval = 0.f, cb = 0.f, cl = 0.f;
                if (t < Tb && u <= Ub) {
                    const float mm = mA[r] + mB[uu];
                    const float shift = mm - z2[q];  // -log2 S(t,u)
                    if (shift > -kTinyLog2) atomicAdd(&n_exact, 1);  // C = 0: handled by the exact path
                    else cval = gc * fast_ex2(e16m16_log2_ratio(aq[q], bq[q], llq) + shift);
                    // log-domain p(blank), p(label): representable far below 2^-126
                    const float lb2 = lAb[r] + lBb[uu] + shift;
                    if (t < Tb - 1) cb = gc * fast_ex2(e16m16_log2_ratio(aq[q], bdn[q], llq) + lb2);
                    else if (u == Ub) cb = gc * fast_ex2(e16m16_log2_ratio(aq[q], 0, llq) + lb2);
                    if (u < Ub) {
                        const float ll2 = pey[q] + lBy[uu] + shift;
                        cl = gc * fast_ex2(e16m16_log2_ratio(aq[q], brt[q], llq) + ll2);
                    }
                }
                Cs[r * kCs + uu] = cval;
                CBs[r * kCs + uu] = cb;
                CLs[r * kCs + uu] = cl;
            }
        }
        __syncthreads();
        // corrections, in a fixed order.  Row r of d_penc loses cb at the blank column and cl at the
        // label column of every position: one thread per frame walks the positions sequentially
        // (two positions may share a label, so this must not be a parallel scatter).
        if (tid < kGT2) {
            float* xr = Xs + tid * Vs;
            float sb = 0.f;
            for (int uu = 0; uu < rows_u; ++uu) {
                sb += CBs[tid * kCs + uu];
                const int y = ys[uu];
                if (y >= 0) xr[y] -= CLs[tid * kCs + uu];
            }
            xr[blank] -= sb;
        } else if (tid >= 64 && tid < 64 + kGUC2) {
            const int uu = tid - 64;
            float sb = 0.f, sl = 0.f;
            for (int r = 0; r < kGT2; ++r) { sb += CBs[r * kCs + uu]; sl += CLs[r * kCs + uu]; }
            ub[uu] = sb;
            ul[uu] = sl;
        }
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
        __syncthreads();  // ub / ul are complete
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
                const float zz = lse[c] * kLog2e;
                if (!(mA[r] + mB[uu] - zz > -kTinyLog2)) continue;
                const float occ = e16m16_log2_ratio(alpha[c], beta[c], llq) - zz;
                const float* pe = penc + ((size_t)b * T + t) * V;
                const float* pd = pdec + ((size_t)b * U1 + u) * V;
                for (int v = 0; v < V; ++v) {
                    const float g = gc * fast_ex2((pe[v] + pd[v]) * kLog2e + occ);
                    atomicAdd(Xs + r * Vs + v, g);
                    if (slab) atomicAdd(slab + (size_t)u * V + v, g);
                    else atomicAdd(d_pdec + ((size_t)b * U1 + u) * V + v, g);
                }
            }
        }
    }
    __syncthreads();
    // d_penc = A .* E + corrections (+ exact-path cells); padded rows come out as zeros
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = 4 * ty + i;
        if (t0 + r >= T) continue;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int v = tx + 16 * c;
            if (v >= V) continue;
            d_penc[((size_t)b * T + t0 + r) * V + v] = fmaf(As[r * Vs + v], E[i][c], Xs[r * Vs + v]);
        }
    }
}

template <int NC>
int launch_grad_mm(const float* penc, const float* pdec, const int32_t* labels, const int32_t* act_lens,
                   const int32_t* label_lens, int B, int T, int U1, int V, int blank, const float* lse,
                   const int32_t* alpha, const int32_t* beta, const float* grad_costs, float* d_penc,
                   float* d_pdec, float* partial, cudaStream_t stream) {
    const int Vs = V | 1;
    const size_t smem = ((size_t)(2 * kGT2 + kGUC2) * Vs + 3 * kGT2 * kCs + 3 * kGT2 + 5 * kGUC2) * sizeof(float);
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
    const size_t smem = ((size_t)(kFT + kFUC) * Vs + 2 * kFT + 3 * kFUC) * sizeof(float);
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
