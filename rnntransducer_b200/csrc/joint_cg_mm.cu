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
// floor of cg_lse_kernel / cg_grad_kernel) become three small batched GEMMs.
//
// The GEMMs are per-CTA products of shared-memory tiles (32 x 64 x V, 32 x V x 48, 48 x V x 32) and
// run as warp-level tensor-core MMAs on split operands, hi*hi + hi*lo + lo*hi:
//   forward (the partition feeds every log-probability of the lattice): mma.sync m16n8k8, TF32
//     halves split in the loop, dropped term 2^-22 relative, i.e. fp32 accuracy;
//   backward: mma.sync m16n8k16 on operands stored as packed (bf16 hi | bf16 lo) pairs -- the pair
//     rides in the instruction's k dimension, so the loops contain no conversions at all; 2^-17
//     relative, far inside the 1e-4 gradient tolerance.
// One warp instruction replaces 1024-2048 FFMA lanes; tcgen05 would add a TMEM round trip per
// 32-frame tile for products this small (K = V <= 128).
//
// Range: A, B are in (0,1].  If the peaks of the two rows do not line up, S can be tiny; cells
// with S < 2^-66 (logit ranges beyond ~45 nats in BOTH rows, never seen with real activations) take
// an exact per-cell path instead (log-sum-exp with the true maximum / explicit V-wide gradient).
//
// V <= 128: a CTA holds whole rows of the factor planes.  V > 128 ("wide"): the same products in
// 128-column chunks -- the partition accumulates over the chunks in the MMA accumulators (the vocabulary is
// its K dimension), the gradient takes one CTA per (utterance, 32 frames, 128 columns): the vocabulary is its
// N dimension and the per-cell scalars are cheap enough to recompute per chunk.  With RNNTB200_CG_GENERIC
// set: the generic per-cell kernels of joint_cg.cu (one exponential per cell and column; kept as an
// independent evaluation for the tests).
#include <cstdlib>

#include "common.cuh"

namespace rnntb200 {

namespace {

constexpr float kTinyLog2 = -66.f;  // log2 of the partition threshold below which a cell goes exact

// x = hi + lo with hi, lo representable in TF32 (round-to-nearest both times: no one-sided bias)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float r = x - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}

// x -> (bf16 hi | bf16 lo << 16) with hi + lo = x to 2^-18 relative; and back
__device__ __forceinline__ uint32_t pack_hilo(float x) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const float hf = __bfloat162float(h);
    const __nv_bfloat162 p = __floats2bfloat162_rn(hf, x - hf);  // .x (low half) = hi exactly, .y = lo
    return *reinterpret_cast<const uint32_t*>(&p);
}
__device__ __forceinline__ float unpack_hilo(uint32_t p) {
    return __uint_as_float(p << 16) + __uint_as_float(p & 0xffff0000u);
}

// D += A B with bf16 operands, m16n8k16, fp32 accumulate.  Register r of a fragment holds the two
// CONSECUTIVE k indices (2q', 2q'+1).  The kernels below feed it packed (hi, lo) pairs, i.e. the
// instruction's k index runs over (element, half): lane (g, q) then holds for A the elements
// [g][q], [g+8][q], [g][q+4], [g+8][q+4] and for B [q][g], [q+4][g] -- the same addressing as the
// TF32 fragments -- and one instruction covers 8 elements of the real K dimension.
// RNNTB200_EXP_NO_MMA (scripts/build_exp_nomma.sh, never the product build): the products and with them their
// operand loads / conversions disappear -- what remains is staging, per-cell scalars and epilogues, i.e. what
// ANY tensor-core formulation (mma.sync or tcgen05) of these kernels would still have to do.
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
#ifdef RNNTB200_EXP_NO_MMA
    d[0] += 1e-3f;
    return;
#endif
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// (a_hi + a_lo)(b_hi + b_lo) ~ a_hi b_hi + a_lo b_hi + a_hi b_lo from packed operands: A as stored;
// B once as (hi, hi) and once as (lo, 0) -- a byte permute and a shift instead of any conversion
__device__ __forceinline__ void mma_hilo(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    const uint32_t hh[2] = {__byte_perm(b0, b0, 0x1010), __byte_perm(b1, b1, 0x1010)};
    const uint32_t l0[2] = {b0 >> 16, b1 >> 16};
    mma_bf16(d, a, hh);
    mma_bf16(d, a, l0);
}

// D += A B, A 16x8 (row), B 8x8 (col), fp32 accumulate.  Lane (g = lane/4, q = lane%4) holds
//   a0 = A[g][q]  a1 = A[g+8][q]  a2 = A[g][q+4]  a3 = A[g+8][q+4];  b0 = B[q][g]  b1 = B[q+4][g];
//   d0 = D[g][2q]  d1 = D[g][2q+1]  d2 = D[g+8][2q]  d3 = D[g+8][2q+1]
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
#ifdef RNNTB200_EXP_NO_MMA
    d[0] += 1e-3f, d[1] += 1e-3f, d[2] += 1e-3f, d[3] += 1e-3f;
    return;
#endif
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Factor planes (CgFactors, common.cuh): cg_factor_rows_kernel turns every row of P_enc / P_dec into
// E = 2^((x - rowmax) log2e) (row stride Vk = V rounded up to 8, pad columns zero) ONCE per step,
// together with the per-row scalars the cell kernels need: the row maximum (base 2), the
// normalised base-2 log at the blank column and, for predictor rows, at the row's label (those can
// underflow in E although they are representable -- and may lie on the best path).  The cell
// kernels then stage tiles of E with 16-byte cp.async; no CTA repeats the exponentials (every
// utterance's P_dec used to be re-exponentiated by each of its T/32 frame tiles).
// One warp per row.
__global__ void __launch_bounds__(256)
cg_factor_rows_kernel(const float* __restrict__ penc, const float* __restrict__ pdec,
                      const int32_t* __restrict__ labels, const int32_t* __restrict__ label_lens, int rows_enc,
                      int rows_dec, int U1, int V, int Vk, int blank, float* __restrict__ Ea,
                      float* __restrict__ mA, float* __restrict__ lAb, float* __restrict__ Eb,
                      float* __restrict__ mB, float* __restrict__ lBb, float* __restrict__ lBy,
                      uint32_t* __restrict__ Ea2, uint32_t* __restrict__ Eb2) {
    pdl_launch_dependents();
    pdl_wait();  // penc / pdec come from the projection kernel
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows_enc + rows_dec) return;
    const bool is_dec = row >= rows_enc;
    const int r = is_dec ? row - rows_enc : row;
    const float* x = (is_dec ? pdec : penc) + (size_t)r * V;
    float* e = (is_dec ? Eb : Ea) + (size_t)r * Vk;
    uint32_t* e2 = (is_dec ? Eb2 : Ea2) + (size_t)r * Vk;
    float xv[4];  // V <= 128
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int v = lane + 32 * i;
        xv[i] = v < V ? __ldg(x + v) : -INFINITY;
        m = fmaxf(m, xv[i]);
    }
    m = warp_max(m);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int v = lane + 32 * i;
        if (v < Vk) {
            const float ev = v < V ? fast_ex2((xv[i] - m) * kLog2e) : 0.f;
            e[v] = ev;
            e2[v] = pack_hilo(ev);
        }
    }
    if (lane == 0) {
        const float lb = (__ldg(x + blank) - m) * kLog2e;
        if (!is_dec) {
            mA[r] = m * kLog2e;
            lAb[r] = lb;
        } else {
            mB[r] = m * kLog2e;
            lBb[r] = lb;
            const int b = r / U1, u = r - b * U1;
            const int Ub = len_U(label_lens, b, U1);
            lBy[r] = u < Ub ? (__ldg(x + label_at(labels, b, U1, u, V)) - m) * kLog2e : 0.f;
        }
    }
}

// The same for rows wider than 128 columns: two passes over the (L2-resident) row instead of registers.
__global__ void __launch_bounds__(256)
cg_factor_rows_wide_kernel(const float* __restrict__ penc, const float* __restrict__ pdec,
                           const int32_t* __restrict__ labels, const int32_t* __restrict__ label_lens, int rows_enc,
                           int rows_dec, int U1, int V, int Vk, int blank, float* __restrict__ Ea,
                           float* __restrict__ mA, float* __restrict__ lAb, float* __restrict__ Eb,
                           float* __restrict__ mB, float* __restrict__ lBb, float* __restrict__ lBy,
                           uint32_t* __restrict__ Ea2, uint32_t* __restrict__ Eb2) {
    pdl_launch_dependents();
    pdl_wait();  // penc / pdec come from the projections
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows_enc + rows_dec) return;
    const bool is_dec = row >= rows_enc;
    const int r = is_dec ? row - rows_enc : row;
    const float* x = (is_dec ? pdec : penc) + (size_t)r * V;
    float* e = (is_dec ? Eb : Ea) + (size_t)r * Vk;
    uint32_t* e2 = (is_dec ? Eb2 : Ea2) + (size_t)r * Vk;
    float m = -INFINITY;
    for (int v = lane; v < V; v += 32) m = fmaxf(m, __ldg(x + v));
    m = warp_max(m);
    for (int v = lane; v < Vk; v += 32) {
        const float ev = v < V ? fast_ex2((__ldg(x + v) - m) * kLog2e) : 0.f;
        e[v] = ev;
        e2[v] = pack_hilo(ev);
    }
    if (lane == 0) {
        const float lb = (__ldg(x + blank) - m) * kLog2e;
        if (!is_dec) {
            mA[r] = m * kLog2e;
            lAb[r] = lb;
        } else {
            mB[r] = m * kLog2e;
            lBb[r] = lb;
            const int b = r / U1, u = r - b * U1;
            const int Ub = len_U(label_lens, b, U1);
            lBy[r] = u < Ub ? (__ldg(x + label_at(labels, b, U1, u, V)) - m) * kLog2e : 0.f;
        }
    }
}

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}
// rows [0, n_copy) of a factor plane (row stride Vk floats) -> shared memory (row stride Vs), rows
// [n_copy, n_rows) -> zeros.  Fire-and-forget; stage_wait + a block barrier make the tile visible.
__device__ __forceinline__ void stage_tile(float* dst, const float* __restrict__ src, int n_copy, int n_rows, int Vk,
                                           int Vs) {
    const int cpr = Vk >> 2;  // 16-byte chunks per row (<= 32): one warp per row, one lane per chunk
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    if (lane < cpr) {
        for (int r = warp; r < n_rows; r += n_warps) {
            if (r < n_copy) cp_async_16(dst + r * Vs + 4 * lane, src + (size_t)r * Vk + 4 * lane);
            else *reinterpret_cast<float4*>(dst + r * Vs + 4 * lane) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
// columns [0, width) (width a multiple of 8, <= 128) of rows [0, n_copy) of a plane with row stride
// src_stride -> shared memory (row stride Vs), rows [n_copy, n_rows) -> zeros.  Does NOT commit: the caller
// groups several tiles into one cp.async group.
__device__ __forceinline__ void stage_cols(float* dst, const float* __restrict__ src, int n_copy, int n_rows,
                                           int src_stride, int width, int Vs) {
    const int cpr = width >> 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    if (lane < cpr) {
        for (int r = warp; r < n_rows; r += n_warps) {
            if (r < n_copy) cp_async_16(dst + r * Vs + 4 * lane, src + (size_t)r * src_stride + 4 * lane);
            else *reinterpret_cast<float4*>(dst + r * Vs + 4 * lane) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}
__device__ __forceinline__ void stage_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// per-row scalars of the same rows (0 beyond n_copy); joins the cp.async group committed next
__device__ __forceinline__ void stage_scalars(float* dst, const float* __restrict__ src, int n_copy, int n_rows) {
    for (int i = threadIdx.x; i < n_rows; i += blockDim.x) {
        if (i < n_copy) {
            const unsigned d = (unsigned)__cvta_generic_to_shared(dst + i);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src + i) : "memory");
        } else {
            dst[i] = 0.f;
        }
    }
}

// =================================================================================================
// forward: CTA = (utterance, 32 frames), 8 warps; per 64-position chunk warp w owns the 16-frame
// row tile (w & 1) x two 8-position column tiles (w >> 1): S = A B^T on tensor cores.
constexpr int kFThreads = 256;
constexpr int kFT = 32;    // frames per CTA
constexpr int kFUC = 64;   // label positions per staged chunk

__global__ void __launch_bounds__(kFThreads)
cg_lse_mm_kernel(const float* __restrict__ penc, const float* __restrict__ pdec, const CgFactors F,
                 const int32_t* __restrict__ labels, const int32_t* __restrict__ act_lens,
                 const int32_t* __restrict__ label_lens, int T, int U1, int V, int Vk, int Vs, int blank,
                 float2* __restrict__ lp2, float* __restrict__ lse_out) {
    extern __shared__ float smem[];
    float* As = smem;                   // [kFT][Vs]   (Vs = Vk + 4: conflict-free fragment loads)
    float* Bs0 = As + kFT * Vs;         // [2][kFUC][Vs]: the next chunk's rows land while this one is used
    float* mA = Bs0 + 2 * kFUC * Vs;    // [kFT]      row maxima, base 2
    float* lAb = mA + kFT;              // [kFT]      log2 A[t][blank]
    float* sc0 = lAb + kFT;             // [2][3][kFUC]  per chunk: row maxima, log2 B[u][blank], log2 B[u][y_u]
    __shared__ int ys0[2][kFUC];
    pdl_launch_dependents();
    const int b = blockIdx.y, t0 = blockIdx.x * kFT;
    const int Tb = len_T(act_lens, b, T), Ub = len_U(label_lens, b, U1);
    if (t0 >= Tb) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int mt = warp & 1, nq = warp >> 1;
    pdl_wait();  // the factor planes come from cg_factor_rows_kernel

    // everything chunk c needs travels as one cp.async group (+ plain stores of the labels) into
    // buffer c & 1, issued one chunk ahead
    auto issue_chunk = [&](int c) {
        const int u0 = c * kFUC, n = min(kFUC, Ub + 1 - u0);
        const size_t row = (size_t)b * U1 + u0;
        float* sc = sc0 + (c & 1) * 3 * kFUC;
        stage_scalars(sc, F.mB + row, n, kFUC);
        stage_scalars(sc + kFUC, F.lBb + row, n, kFUC);
        stage_scalars(sc + 2 * kFUC, F.lBy + row, n, kFUC);
        if (threadIdx.x < kFUC) {
            const int u = u0 + threadIdx.x;
            ys0[c & 1][threadIdx.x] = u < Ub ? label_at(labels, b, U1, u, V) : -1;
        }
        stage_tile(Bs0 + (c & 1) * kFUC * Vs, F.Eb + row * Vk, n, kFUC, Vk, Vs);
    };
    {
        const int n = min(kFT, Tb - t0);
        const size_t row = (size_t)b * T + t0;
        stage_scalars(mA, F.mA + row, n, kFT);
        stage_scalars(lAb, F.lAb + row, n, kFT);
        stage_tile(As, F.Ea + row * Vk, n, kFT, Vk, Vs);
    }
    issue_chunk(0);
    int chunk = 0;
    for (int u0 = 0; u0 <= Ub; u0 += kFUC, ++chunk) {
        const float* Bs = Bs0 + (chunk & 1) * kFUC * Vs;
        const float* mB = sc0 + (chunk & 1) * 3 * kFUC;
        const float* lBb = mB + kFUC;
        const float* lBy = lBb + kFUC;
        const int* ys = ys0[chunk & 1];
        stage_wait();
        __syncthreads();  // this chunk has landed; every warp is done with the previous one
        if (u0 + kFUC <= Ub) issue_chunk(chunk + 1);

        const int un0 = nq * 16;  // first position (within the chunk) of this warp's column tiles
        if (t0 + mt * 16 >= Tb || u0 + un0 > Ub) continue;  // warp-uniform
        const bool n_on1 = u0 + un0 + 8 <= Ub;
        float acc[2][4], acs[2][4];  // hi*hi / the two cross terms
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][k] = acs[i][k] = 0.f;
        const float* arow = As + (mt * 16 + g) * Vs + q;
        const float* brow = Bs + (un0 + g) * Vs + q;
#pragma unroll 2
        for (int k0 = 0; k0 < Vk; k0 += 8) {
            uint32_t ah[4], al[4];
            split_tf32(arow[k0], ah[0], al[0]);
            split_tf32(arow[k0 + 8 * Vs], ah[1], al[1]);
            split_tf32(arow[k0 + 4], ah[2], al[2]);
            split_tf32(arow[k0 + 8 * Vs + 4], ah[3], al[3]);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (i == 1 && !n_on1) continue;
                uint32_t bh[2], bl[2];
                split_tf32(brow[i * 8 * Vs + k0], bh[0], bl[0]);
                split_tf32(brow[i * 8 * Vs + k0 + 4], bh[1], bl[1]);
                mma_tf32(acc[i], ah, bh);
                mma_tf32(acs[i], ah, bl);
                mma_tf32(acs[i], al, bh);
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (i == 1 && !n_on1) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int r = mt * 16 + g + 8 * h, uu = un0 + 8 * i + 2 * q + k;
                    const int t = t0 + r, u = u0 + uu;
                    if (t >= Tb || u > Ub) continue;
                    const float mm = mA[r] + mB[uu];
                    const float* pe = penc + ((size_t)b * T + t) * V;
                    float lgs = fast_lg2(acc[i][2 * h + k] + acs[i][2 * h + k]);
                    const int y = u < Ub ? ys[uu] : blank;
                    float lb2, ll2;
                    if (lgs < kTinyLog2) {
                        // exact path: the row peaks do not line up, redo this cell in the log domain
                        const float* pd = pdec + ((size_t)b * U1 + u) * V;
                        float mx = -INFINITY;
                        for (int v = 0; v < V; ++v) mx = fmaxf(mx, (pe[v] + pd[v]) * kLog2e);
                        float se = 0.f;
                        for (int v = 0; v < V; ++v) se += fast_ex2((pe[v] + pd[v]) * kLog2e - mx);
                        lgs = mx + fast_lg2(se) - mm;
                    }
                    // blank / label log-probs straight in the log domain (A * B would underflow
                    // below 2^-126 although such steps are representable -- and may be on the path)
                    lb2 = lAb[r] + lBb[uu] - lgs;
                    const float ay = As[r * Vs + y];  // A[t][y]: its log unless it underflowed
                    const float lay = ay > 1e-30f ? fast_lg2(ay) : __ldg(pe + y) * kLog2e - mA[r];
                    ll2 = lay + lBy[uu] - lgs;
                    const size_t c = ((size_t)b * T + t) * U1 + u;
                    lp2[c] = make_float2(fmaxf(lb2 * kLn2, kNegInf), u < Ub ? fmaxf(ll2 * kLn2, kNegInf) : 0.f);
                    lse_out[c] = (mm + lgs) * kLn2;
                }
        }
    }
}

// =================================================================================================
// forward, wide vocabulary (128 < V): the same CTA / warp tiling; the factor rows no longer fit, so the
// partition's K dimension (the vocabulary) is walked in 128-column chunks -- per (64-position chunk,
// 128-column chunk) step one A tile [32][128] and one B tile [64][128] land (double-buffered, one step
// ahead) and the accumulators carry over the column chunks of a position chunk.  The per-row scalars and
// A[t][y] of the epilogue come straight from the (L2-resident) factor planes.
constexpr int kWC = 128;       // vocabulary columns per staged chunk
constexpr int kWCs = kWC + 4;  // shared-memory row stride (4 mod 32: conflict-free fragment loads)

__global__ void __launch_bounds__(kFThreads)
cg_lse_mmw_kernel(const float* __restrict__ penc, const float* __restrict__ pdec, const CgFactors F,
                  const int32_t* __restrict__ labels, const int32_t* __restrict__ act_lens,
                  const int32_t* __restrict__ label_lens, int T, int U1, int V, int Vk, int blank,
                  float2* __restrict__ lp2, float* __restrict__ lse_out) {
    extern __shared__ float smem[];  // [2] x { A [kFT][kWCs], B [kFUC][kWCs] }
    constexpr int kBuf = (kFT + kFUC) * kWCs;
    pdl_launch_dependents();
    const int b = blockIdx.y, t0 = blockIdx.x * kFT;
    const int Tb = len_T(act_lens, b, T), Ub = len_U(label_lens, b, U1);
    if (t0 >= Tb) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int mt = warp & 1, nq = warp >> 1;
    pdl_wait();  // the factor planes come from cg_factor_rows_wide_kernel

    const int n_vc = (Vk + kWC - 1) / kWC;  // column chunks
    const int n_uc = Ub / kFUC + 1;         // position chunks covering u = 0 .. Ub
    const int n_steps = n_uc * n_vc;
    auto issue = [&](int s) {  // both tiles of step s, one cp.async group, into buffer s & 1
        const int uc = s / n_vc, vc = s - uc * n_vc;
        const int v0 = vc * kWC, w = min(kWC, Vk - v0);
        float* Ad = smem + (s & 1) * kBuf;
        float* Bd = Ad + kFT * kWCs;
        stage_cols(Ad, F.Ea + ((size_t)b * T + t0) * Vk + v0, min(kFT, Tb - t0), kFT, Vk, w, kWCs);
        stage_cols(Bd, F.Eb + ((size_t)b * U1 + uc * kFUC) * Vk + v0, min(kFUC, Ub + 1 - uc * kFUC), kFUC, Vk, w,
                   kWCs);
        stage_commit();
    };
    issue(0);
    float acc[2][4], acs[2][4];  // hi*hi / the two cross terms
    int uc = 0, vc = 0;
    for (int s = 0; s < n_steps; ++s) {
        const int u0 = uc * kFUC;
        const bool last_vc = vc == n_vc - 1;
        const int w = min(kWC, Vk - vc * kWC);
        const float* As = smem + (s & 1) * kBuf;
        const float* Bs = As + kFT * kWCs;
        stage_wait();
        __syncthreads();  // this step's tiles have landed; every warp is done with the previous step's
        if (s + 1 < n_steps) issue(s + 1);

        const int un0 = nq * 16;  // first position (within the chunk) of this warp's column tiles
        const bool warp_on = t0 + mt * 16 < Tb && u0 + un0 <= Ub;  // warp-uniform, the same for every vc of this uc
        const bool n_on1 = u0 + un0 + 8 <= Ub;
        if (warp_on) {
            if (vc == 0) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[i][k] = acs[i][k] = 0.f;
            }
            const float* arow = As + (mt * 16 + g) * kWCs + q;
            const float* brow = Bs + (un0 + g) * kWCs + q;
#pragma unroll 2
            for (int k0 = 0; k0 < w; k0 += 8) {
                uint32_t ah[4], al[4];
                split_tf32(arow[k0], ah[0], al[0]);
                split_tf32(arow[k0 + 8 * kWCs], ah[1], al[1]);
                split_tf32(arow[k0 + 4], ah[2], al[2]);
                split_tf32(arow[k0 + 8 * kWCs + 4], ah[3], al[3]);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (i == 1 && !n_on1) continue;
                    uint32_t bh[2], bl[2];
                    split_tf32(brow[i * 8 * kWCs + k0], bh[0], bl[0]);
                    split_tf32(brow[i * 8 * kWCs + k0 + 4], bh[1], bl[1]);
                    mma_tf32(acc[i], ah, bh);
                    mma_tf32(acs[i], ah, bl);
                    mma_tf32(acs[i], al, bh);
                }
            }
            if (last_vc) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (i == 1 && !n_on1) continue;
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const int r = mt * 16 + g + 8 * h, uu = un0 + 8 * i + 2 * q + k;
                            const int t = t0 + r, u = u0 + uu;
                            if (t >= Tb || u > Ub) continue;
                            const size_t ra = (size_t)b * T + t, rb = (size_t)b * U1 + u;
                            const float mAr = __ldg(F.mA + ra), mm = mAr + __ldg(F.mB + rb);
                            const float* pe = penc + ra * V;
                            float lgs = fast_lg2(acc[i][2 * h + k] + acs[i][2 * h + k]);
                            const int y = u < Ub ? label_at(labels, b, U1, u, V) : blank;
                            if (lgs < kTinyLog2) {
                                // exact path: the row peaks do not line up, redo this cell in the log domain
                                const float* pd = pdec + rb * V;
                                float mx = -INFINITY;
                                for (int v = 0; v < V; ++v) mx = fmaxf(mx, (pe[v] + pd[v]) * kLog2e);
                                float se = 0.f;
                                for (int v = 0; v < V; ++v) se += fast_ex2((pe[v] + pd[v]) * kLog2e - mx);
                                lgs = mx + fast_lg2(se) - mm;
                            }
                            const float lb2 = __ldg(F.lAb + ra) + __ldg(F.lBb + rb) - lgs;
                            const float ay = __ldg(F.Ea + ra * Vk + y);  // A[t][y]: its log unless it underflowed
                            const float lay = ay > 1e-30f ? fast_lg2(ay) : __ldg(pe + y) * kLog2e - mAr;
                            const float ll2 = lay + __ldg(F.lBy + rb) - lgs;
                            const size_t c = ra * U1 + u;
                            lp2[c] = make_float2(fmaxf(lb2 * kLn2, kNegInf), u < Ub ? fmaxf(ll2 * kLn2, kNegInf) : 0.f);
                            lse_out[c] = (mm + lgs) * kLn2;
                        }
                }
            }
        }
        if (last_vc) {
            vc = 0;
            ++uc;
        } else {
            ++vc;
        }
    }
}

// =================================================================================================
// backward: CTA = (utterance, 32 frames), 8 warps.  Per 48-position chunk: the per-cell scalars
// (C and the blank / label corrections) go to shared memory, then on tensor cores
//   E += C B     (32 x Vk, K = 48 positions; warp w: row tile w & 1, column quarter w >> 1; the
//                 accumulators stay in registers across the chunks)
//   D  = C^T A   (48 x Vk, K = 32 frames; warp w: its column quarter, row tiles 0, 2 or 1)
// and d_pdec += B .* D - corrections.  The d_penc corrections (row t loses cb at the blank column and
// cl at the label column of every position) are a fourth product, CL Y + CB Y_blank with one-hot
// columns generated in registers; the d_pdec ones are column sums.  All of it has a fixed summation
// order, so the only order-dependent arithmetic is the cross-tile accumulation of d_pdec (fp32
// atomics, or per-tile slabs + a fixed-order reduction in deterministic mode) and the cold exact path.
constexpr int kGT2 = 32;   // frames per CTA
constexpr int kGUC2 = 48;  // label positions per chunk
constexpr int kGThreads = 256;
constexpr int kCs = 52;    // row stride of the C planes (4 mod 8: conflict-free A fragments of C B)
static_assert(kGT2 * kGUC2 == 6 * kGThreads, "thread (rr, cc) owns 2 rows x 3 columns of the cell block");

// kWide (128 < V): blockIdx.z selects a 128-column chunk of the vocabulary; the CTA stages only those columns
// of A and B, recomputes the per-cell scalars (cheap next to the products) and owns those columns of d_penc /
// d_pdec / the slab.  Everything indexed by a column below is local to the chunk unless it says v_off.
template <int NTW, bool kWide>  // NTW: 8-column tiles per warp, ceil(columns of the CTA / 32)
__global__ void __launch_bounds__(kGThreads, 3)
cg_grad_mm_kernel(const float* __restrict__ penc, const float* __restrict__ pdec, const CgFactors F,
                  const int32_t* __restrict__ labels, const int32_t* __restrict__ act_lens,
                  const int32_t* __restrict__ label_lens, int T, int U1, int V, int Vk, int Vs, int blank,
                  const float* __restrict__ lse, const int32_t* __restrict__ alpha,
                  const int32_t* __restrict__ beta, GradCosts grad_costs,
                  float* __restrict__ d_penc, float* __restrict__ d_pdec,
                  float* __restrict__ partial /* deterministic slabs or null */) {
    // operand planes hold packed (bf16 hi | bf16 lo << 16) pairs (pack_hilo): A / B arrive that way
    // from the factor planes, the per-cell scalars are packed when they are stored
    extern __shared__ float smem[];
    uint32_t* As = reinterpret_cast<uint32_t*>(smem);  // [32][Vs]  A = 2^(P_enc - max)   (Vs = 8 mod 16: conflict-free B fragments)
    uint32_t* Bs0 = As + kGT2 * Vs;       // [2][48][Vs]  B chunk, double-buffered (next chunk lands during this one)
    uint32_t* Cs = Bs0 + 2 * kGUC2 * Vs;  // [32][52]  C chunk
    uint32_t* CBs = Cs + kGT2 * kCs;      // [32][52]  blank corrections
    uint32_t* CLs = CBs + kGT2 * kCs;     // [32][52]  label corrections
    float* mA = reinterpret_cast<float*>(CLs + kGT2 * kCs);  // [32] row maxima (base 2)
    float* lAb = mA + kGT2;            // [32] log2 A[t][blank]
    float* sc0 = lAb + kGT2;           // [2][3][48] per chunk: row maxima, log2 B[u][blank], log2 B[u][y_u]
    float* ub = sc0 + 6 * kGUC2;       // [48] sum_t corr_blank
    float* ul = ub + kGUC2;            // [48] sum_t corr_label
    __shared__ int ys0[2][kGUC2];
    __shared__ int n_exact;

    const int b = blockIdx.y, tile = blockIdx.x, t0 = tile * kGT2;
    const int Tb = len_T(act_lens, b, T), Ub = len_U(label_lens, b, U1);
    const int n_tiles = gridDim.x;
    const int v_off = kWide ? (int)blockIdx.z * kWC : 0;  // first vocabulary column of this CTA
    const int Vc = kWide ? min(kWC, Vk - v_off) : Vk;      // its column count (a multiple of 8)
    float* slab = partial ? partial + ((size_t)b * n_tiles + tile) * U1 * V : nullptr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int mt = warp & 1;                 // 16-frame row tile of E
    const int nt0 = (warp >> 1) * NTW;       // first 8-column tile of this warp (column quarter)
    const int n_nt = Vc >> 3;                // column tiles in use
    constexpr uint32_t kOnes = 0x3f803f80u;  // (1.0, 1.0) in bf16: selects hi + lo of a packed operand
    pdl_launch_dependents();
    pdl_wait();  // nothing in global memory is read OR written before the sweep (the predecessor) has finished

    if (t0 >= Tb) {  // tile entirely in the padding: exact zeros
        if constexpr (kWide) {
            const int nv = min(Vc, V - v_off);  // real columns of this chunk
            for (int i = tid; i < kGT2 * nv; i += kGThreads) {
                const int r = i / nv, v = v_off + i - r * nv;
                if (t0 + r < T) d_penc[((size_t)b * T + t0 + r) * V + v] = 0.f;
            }
            if (slab)
                for (int i = tid; i < U1 * nv; i += kGThreads) {
                    const int u = i / nv, v = v_off + i - u * nv;
                    slab[(size_t)u * V + v] = 0.f;
                }
        } else {
            for (int i = tid; i < kGT2 * V; i += kGThreads) {
                const int r = i / V, v = i - r * V;
                if (t0 + r < T) d_penc[((size_t)b * T + t0 + r) * V + v] = 0.f;
            }
            if (slab)
                for (int i = tid; i < U1 * V; i += kGThreads) slab[i] = 0.f;
        }
        return;
    }
    const float gc = grad_costs.at(b);
    const int llq = beta[(size_t)b * T * U1];  // beta(0,0) = P(y|x), e16m16
    const int rows_t = min(kGT2, Tb - t0);

    if (tid == 0) n_exact = 0;
    auto issue_chunk = [&](int c) {  // everything chunk c needs, into buffer c & 1, one chunk ahead
        const int u0 = c * kGUC2, n = max(0, min(kGUC2, Ub + 1 - u0));
        const size_t row = (size_t)b * U1 + u0;
        float* sc = sc0 + (c & 1) * 3 * kGUC2;
        stage_scalars(sc, F.mB + row, n, kGUC2);
        stage_scalars(sc + kGUC2, F.lBb + row, n, kGUC2);
        stage_scalars(sc + 2 * kGUC2, F.lBy + row, n, kGUC2);
        if (tid < kGUC2) {
            const int u = u0 + tid;
            ys0[c & 1][tid] = u < Ub ? label_at(labels, b, U1, u, V) : -1;
        }
        if constexpr (kWide) {
            stage_cols(reinterpret_cast<float*>(Bs0 + (c & 1) * kGUC2 * Vs),
                       reinterpret_cast<const float*>(F.Eb2 + row * Vk + v_off), n, kGUC2, Vk, Vc, Vs);
            stage_commit();
        } else {
            stage_tile(reinterpret_cast<float*>(Bs0 + (c & 1) * kGUC2 * Vs), reinterpret_cast<const float*>(F.Eb2 + row * Vk),
                       n, kGUC2, Vk, Vs);
        }
    };
    {
        const size_t row = (size_t)b * T + t0;
        stage_scalars(mA, F.mA + row, rows_t, kGT2);
        stage_scalars(lAb, F.lAb + row, rows_t, kGT2);
        if constexpr (kWide) {
            stage_cols(reinterpret_cast<float*>(As), reinterpret_cast<const float*>(F.Ea2 + row * Vk + v_off), rows_t, kGT2,
                       Vk, Vc, Vs);
            stage_commit();
        } else {
            stage_tile(reinterpret_cast<float*>(As), reinterpret_cast<const float*>(F.Ea2 + row * Vk), rows_t, kGT2, Vk, Vs);
        }
    }
    issue_chunk(0);
    int chunk = 0;

    // fragments of (C B)[t][v] and of the d_penc corrections X[t][v] = sum_u cl(t,u) [v = y_u] + cb(t,u) [v = blank]
    // (a product with one-hot columns generated in registers: exact, fixed order, no scatter)
    float E[NTW][4], X[NTW][4];
#pragma unroll
    for (int i = 0; i < NTW; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) E[i][k] = X[i][k] = 0.f;

    for (int u0 = 0; u0 < U1; u0 += kGUC2, ++chunk) {
        if (!slab && u0 > Ub) break;  // nothing left to add (slabs must be written in full)
        const int rows_u = max(0, min(kGUC2, Ub + 1 - u0));
        const uint32_t* Bs = Bs0 + (chunk & 1) * kGUC2 * Vs;
        const float* mB = sc0 + (chunk & 1) * 3 * kGUC2;
        const float* lBb = mB + kGUC2;
        const float* lBy = lBb + kGUC2;
        const int* ys = ys0[chunk & 1];
        stage_wait();
        __syncthreads();  // this chunk has landed; the previous one is fully consumed
        if (u0 + kGUC2 < U1 && (slab || u0 + kGUC2 <= Ub)) issue_chunk(chunk + 1);

        // per-cell scalars of the (32 x 48) block: the two corrections
        //   cb = grad_cost alpha(t,u) beta(t+1,u) p(blank) / P,   cl = grad_cost alpha(t,u) beta(t,u+1) p(label) / P
        // and C = grad_cost occupancy / S.  The beta recursion itself says occupancy = (cb + cl) /
        // grad_cost, so beta(t,u) is not read.  Ratios are formed from the e16m16 planes with integer
        // exponents: value = mantissa product * 2^(exponent sum + log2 p).  Thread (rr, cc) owns rows
        // 2 rr + {0,1} x columns cc + 16 {0,1,2}; the global loads of a row's 3 cells go out together.
        {
            const int rr = tid >> 4, cc = tid & 15;
            const int e_ll = llq >> 16;
            const float k_ll = gc * fast_rcp(e16m16_mant(llq));
#pragma unroll 1
            for (int i = 0; i < 2; ++i) {
                const int r = 2 * rr + i, t = t0 + r;
                const bool t_ok = t < Tb;
                const size_t cbase = ((size_t)b * T + min(t, T - 1)) * U1 + u0;
                int aq[3], bdn[3], brt[3];
                float z2[3], pey[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int uu = cc + 16 * k, u = u0 + uu;
                    aq[k] = bdn[k] = brt[k] = 0;
                    z2[k] = pey[k] = 0.f;
                    if (t_ok && u <= Ub) {
                        const size_t c = cbase + uu;
                        aq[k] = alpha[c];
                        if (t < Tb - 1) bdn[k] = beta[c + U1];
                        if (u < Ub) {
                            brt[k] = beta[c + 1];
                            // log2 A[t][y_u]; gather if it underflowed (wide: y_u may lie outside this CTA's columns)
                            const float ay = kWide ? __ldg(F.Ea + ((size_t)b * T + t) * Vk + ys[uu])
                                                   : unpack_hilo(As[r * Vs + ys[uu]]);
                            pey[k] = ay > 1e-30f ? fast_lg2(ay)
                                                 : __ldg(penc + ((size_t)b * T + t) * V + ys[uu]) * kLog2e - mA[r];
                        }
                        z2[k] = lse[c] * kLog2e;
                    }
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int uu = cc + 16 * k, u = u0 + uu;
                    float cval = 0.f, cb = 0.f, cl = 0.f;
                    if (t_ok && u <= Ub) {
                        const float shift = mA[r] + mB[uu] - z2[k];  // -log2 S(t,u)
                        const int e_a = (aq[k] >> 16) - e_ll;
                        const float m_a = k_ll * e16m16_mant(aq[k]);
                        // log-domain p(blank), p(label): representable far below 2^-126
                        const float lb2 = lAb[r] + lBb[uu] + shift;
                        if (t < Tb - 1) cb = m_a * e16m16_mant(bdn[k]) * fast_ex2((float)(e_a + (bdn[k] >> 16)) + lb2);
                        else if (u == Ub) cb = m_a * fast_ex2((float)e_a + lb2);
                        if (u < Ub) {
                            const float ll2 = pey[k] + lBy[uu] + shift;
                            cl = m_a * e16m16_mant(brt[k]) * fast_ex2((float)(e_a + (brt[k] >> 16)) + ll2);
                        }
                        if (shift > -kTinyLog2) atomicAdd(&n_exact, 1);  // C = 0: handled by the exact path
                        else cval = (cb + cl) * fast_ex2(shift);
                    }
                    Cs[r * kCs + uu] = pack_hilo(cval);
                    CBs[r * kCs + uu] = pack_hilo(cb);
                    CLs[r * kCs + uu] = pack_hilo(cl);
                }
            }
        }
        __syncthreads();
        // column sums of the two correction planes (what d_pdec loses at the blank / label column)
        if (tid >= 64 && tid < 64 + kGUC2) {
            const int uu = tid - 64;
            float sb = 0.f, sl = 0.f;
            for (int r = 0; r < kGT2; ++r) {
                sb += unpack_hilo(CBs[r * kCs + uu]);
                sl += unpack_hilo(CLs[r * kCs + uu]);
            }
            ub[uu] = sb;
            ul[uu] = sl;
        }
        // E += C B, X += CL Y + CB Y_blank  (K = label positions of the chunk; rows beyond rows_u are zeros)
        {
            const uint32_t* crow = Cs + (mt * 16 + g) * kCs + q;
            const int off_b = CBs - Cs, off_l = CLs - Cs;
#pragma unroll 2
            for (int k0 = 0; k0 < kGUC2; k0 += 8) {
                if (k0 >= rows_u) break;
                const uint32_t ca[4] = {crow[k0], crow[k0 + 8 * kCs], crow[k0 + 4], crow[k0 + 8 * kCs + 4]};
                const uint32_t la[4] = {crow[off_l + k0], crow[off_l + k0 + 8 * kCs], crow[off_l + k0 + 4],
                                        crow[off_l + k0 + 8 * kCs + 4]};
                const uint32_t ba[4] = {crow[off_b + k0], crow[off_b + k0 + 8 * kCs], crow[off_b + k0 + 4],
                                        crow[off_b + k0 + 8 * kCs + 4]};
                const int y0 = ys[k0 + q], y1 = ys[k0 + q + 4];  // -1: no label, matches no column
                const uint32_t* bcol = Bs + (k0 + q) * Vs + nt0 * 8 + g;
#pragma unroll
                for (int i = 0; i < NTW; ++i) {
                    if (nt0 + i >= n_nt) continue;
                    mma_hilo(E[i], ca, bcol[8 * i], bcol[8 * i + 4 * Vs]);
                    const int col = v_off + (nt0 + i) * 8 + g;
                    const uint32_t yh[2] = {y0 == col ? kOnes : 0u, y1 == col ? kOnes : 0u};
                    mma_bf16(X[i], la, yh);
                    if ((blank >> 3) == (v_off >> 3) + nt0 + i) {  // warp-uniform
                        const uint32_t yb[2] = {col == blank ? kOnes : 0u, col == blank ? kOnes : 0u};
                        mma_bf16(X[i], ba, yb);
                    }
                }
            }
        }
        // D = C^T A   (K = frames of the tile), then d_pdec partial = B .* D - corrections
        bool synced = false;
#pragma unroll 1
        for (int m0 = mt * 16; m0 < kGUC2; m0 += 32) {  // row tiles 0, 2 (even warps) / 1 (odd warps)
            if (!slab && m0 >= rows_u) break;
            float D[NTW][4];
#pragma unroll
            for (int i = 0; i < NTW; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k) D[i][k] = 0.f;
            if (m0 < rows_u) {
#pragma unroll 2
                for (int k0 = 0; k0 < kGT2; k0 += 8) {
                    if (k0 >= rows_t) break;
                    const uint32_t* ccol = Cs + (k0 + q) * kCs + m0 + g;  // (C^T)[u][t] = C[t][u]
                    const uint32_t ca[4] = {ccol[0], ccol[8], ccol[4 * kCs], ccol[4 * kCs + 8]};
                    const uint32_t* acol = As + (k0 + q) * Vs + nt0 * 8 + g;
#pragma unroll
                    for (int i = 0; i < NTW; ++i) {
                        if (nt0 + i >= n_nt) continue;
                        mma_hilo(D[i], ca, acol[8 * i], acol[8 * i + 4 * Vs]);
                    }
                }
            }
            if (!synced) {  // ub / ul are complete (the barrier is reached by every warp exactly once per chunk:
                synced = true;  // see below for the warps that skip this loop)
                __syncthreads();
            }
#pragma unroll
            for (int i = 0; i < NTW; ++i) {
                if (nt0 + i >= n_nt) continue;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int uu = m0 + g + 8 * h, u = u0 + uu;
                    if (u >= U1) continue;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int v = v_off + (nt0 + i) * 8 + 2 * q + k;
                        if (v >= V) continue;
                        float gv = 0.f;
                        if (uu < rows_u) {
                            gv = unpack_hilo(Bs[uu * Vs + v - v_off]) * D[i][2 * h + k];
                            if (v == blank) gv -= ub[uu];
                            if (v == ys[uu]) gv -= ul[uu];
                        }
                        if (slab) slab[(size_t)u * V + v] = gv;  // exact-path cells are added at the end
                        else if (uu < rows_u) atomicAdd(d_pdec + ((size_t)b * U1 + u) * V + v, gv);
                    }
                }
            }
        }
        if (!synced) __syncthreads();  // warps without a row tile in this chunk
    }
    // d_penc = A .* E - corrections; padded rows come out as zeros
#pragma unroll
    for (int i = 0; i < NTW; ++i) {
        if (nt0 + i >= n_nt) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = mt * 16 + g + 8 * h;
            if (t0 + r >= T) continue;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int v = v_off + (nt0 + i) * 8 + 2 * q + k;
                if (v >= V) continue;
                d_penc[((size_t)b * T + t0 + r) * V + v] =
                    fmaf(unpack_hilo(As[r * Vs + v - v_off]), E[i][2 * h + k], -X[i][2 * h + k]);
            }
        }
    }
    // exact path (cold): cells whose partition underflows the factorised form contributed C = 0
    // above; their full V-wide gradient is added here (this CTA owns these d_penc rows / this slab)
    __syncthreads();
    if (n_exact > 0) {
        for (int i = tid; i < kGT2 * (Ub + 1); i += kGThreads) {
            const int r = i / (Ub + 1), u = i - r * (Ub + 1);
            const int t = t0 + r;
            if (t >= Tb) continue;
            const size_t c = ((size_t)b * T + t) * U1 + u;
            const float zz = lse[c] * kLog2e;
            if (!(mA[r] + __ldg(F.mB + (size_t)b * U1 + u) - zz > -kTinyLog2)) continue;
            const float occ = e16m16_log2_ratio(alpha[c], beta[c], llq) - zz;
            const float* pe = penc + ((size_t)b * T + t) * V;
            const float* pd = pdec + ((size_t)b * U1 + u) * V;
            for (int v = v_off; v < min(V, v_off + Vc); ++v) {  // this CTA's columns
                const float gg = gc * fast_ex2((pe[v] + pd[v]) * kLog2e + occ);
                atomicAdd(d_penc + ((size_t)b * T + t) * V + v, gg);
                if (slab) atomicAdd(slab + (size_t)u * V + v, gg);
                else atomicAdd(d_pdec + ((size_t)b * U1 + u) * V + v, gg);
            }
        }
    }
}

inline int grad_row_stride(int Vk) { return (Vk & 15) == 8 ? Vk : Vk + 8; }  // 8 mod 16

template <int NTW, bool kWide>
int launch_grad_mm(const float* penc, const float* pdec, const CgFactors& F, const int32_t* labels,
                   const int32_t* act_lens, const int32_t* label_lens, int B, int T, int U1, int V, int blank,
                   const float* lse, const int32_t* alpha, const int32_t* beta, GradCosts grad_costs,
                   float* d_penc, float* d_pdec, float* partial, cudaStream_t stream) {
    const int Vk = F.Vk, Vs = grad_row_stride(kWide ? kWC : Vk);
    const size_t smem = ((size_t)(kGT2 + 2 * kGUC2) * Vs + 3 * kGT2 * kCs + 2 * kGT2 + 8 * kGUC2) *
                        sizeof(float);  // 65 KiB at V = 73: three CTAs per SM, the whole cfg-2 grid in one wave
    cudaError_t e = cudaFuncSetAttribute(cg_grad_mm_kernel<NTW, kWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return status_from_cuda(e);
    dim3 grid((T + kGT2 - 1) / kGT2, B, kWide ? (Vk + kWC - 1) / kWC : 1);
    e = launch_pdl(pdl_ok((long long)B * T), cg_grad_mm_kernel<NTW, kWide>, grid, dim3(kGThreads), smem, stream, penc, pdec, F, labels,
                   act_lens, label_lens, T, U1, V, Vk, Vs, blank, lse, alpha, beta, grad_costs, d_penc, d_pdec, partial);
    return e == cudaSuccess ? launch_status() : status_from_cuda(e);
}

}  // namespace

// every V; RNNTB200_CG_GENERIC selects the per-cell kernels of joint_cg.cu (an independent evaluation, kept for the tests; V <~ 1400)
bool cg_mm_supported(int V) { return V >= 1 && !std::getenv("RNNTB200_CG_GENERIC"); }
int cg_mm_tile_rows() { return kGT2; }

// factor planes: E_enc [B*T][Vk], E_dec [B*U1][Vk], then the per-row scalars
size_t cg_factors_bytes(int B, int T, int U1, int V) {
    if (!cg_mm_supported(V)) return 0;  // the generic kernels need no factors
    const size_t Vk = (V + 7) & ~7, re = (size_t)B * T, rd = (size_t)B * U1;
    return (2 * (re + rd) * Vk + 2 * re + 3 * rd) * sizeof(float);
}

CgFactors cg_factors_layout(void* mem, int B, int T, int U1, int V) {
    const size_t Vk = (V + 7) & ~7, re = (size_t)B * T, rd = (size_t)B * U1;
    float* p = static_cast<float*>(mem);
    CgFactors F;
    F.Vk = (int)Vk;
    F.Ea = p;  // the four planes first: each starts 32-byte aligned (rows * Vk floats, Vk a multiple of 8)
    F.Eb = p + re * Vk;
    F.Ea2 = reinterpret_cast<uint32_t*>(F.Eb + rd * Vk);
    F.Eb2 = F.Ea2 + re * Vk;
    F.mA = reinterpret_cast<float*>(F.Eb2 + rd * Vk);
    F.lAb = F.mA + re;
    F.mB = F.lAb + re;
    F.lBb = F.mB + rd;
    F.lBy = F.lBb + rd;
    return F;
}

int launch_cg_factor_rows(const float* penc, const float* pdec, const int32_t* labels, const int32_t* label_lens,
                          int B, int T, int U1, int V, int blank, const CgFactors& F, cudaStream_t stream) {
    const int rows = B * (T + U1);
    if (rows == 0) return RNNTB200_STATUS_SUCCESS;
    const cudaError_t e =
        V > 128 ? launch_pdl(pdl_ok((long long)B * T), cg_factor_rows_wide_kernel, dim3((rows + 7) / 8), dim3(256), (size_t)0,
                             stream, penc, pdec, labels, label_lens, B * T, B * U1, U1, V, F.Vk, blank, F.Ea, F.mA, F.lAb,
                             F.Eb, F.mB, F.lBb, F.lBy, F.Ea2, F.Eb2)
                : launch_pdl(pdl_ok((long long)B * T), cg_factor_rows_kernel, dim3((rows + 7) / 8), dim3(256), (size_t)0,
                             stream, penc, pdec, labels, label_lens, B * T, B * U1, U1, V, F.Vk, blank, F.Ea, F.mA, F.lAb,
                             F.Eb, F.mB, F.lBb, F.lBy, F.Ea2, F.Eb2);
    return e == cudaSuccess ? launch_status() : status_from_cuda(e);
}

int launch_cg_lse_mm(const float* penc, const float* pdec, const CgFactors& F, const int32_t* labels,
                     const int32_t* act_lens, const int32_t* label_lens, int B, int T, int U1, int V, int blank,
                     float2* lp2, float* lse, cudaStream_t stream) {
    const int Vk = F.Vk, Vs = Vk + 4;
    if (V > 128) {
        const size_t smem = (size_t)2 * (kFT + kFUC) * kWCs * sizeof(float);  // 99 KiB: two CTAs per SM
        cudaError_t e = cudaFuncSetAttribute(cg_lse_mmw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return status_from_cuda(e);
        dim3 grid((T + kFT - 1) / kFT, B);
        e = launch_pdl(pdl_ok((long long)B * T), cg_lse_mmw_kernel, grid, dim3(kFThreads), smem, stream, penc, pdec, F, labels,
                       act_lens, label_lens, T, U1, V, Vk, blank, lp2, lse);
        return e == cudaSuccess ? launch_status() : status_from_cuda(e);
    }
    const size_t smem = ((size_t)(kFT + 2 * kFUC) * Vs + 2 * kFT + 6 * kFUC) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(cg_lse_mm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return status_from_cuda(e);
    dim3 grid((T + kFT - 1) / kFT, B);
    e = launch_pdl(pdl_ok((long long)B * T), cg_lse_mm_kernel, grid, dim3(kFThreads), smem, stream, penc, pdec, F, labels, act_lens,
                   label_lens, T, U1, V, Vk, Vs, blank, lp2, lse);
    return e == cudaSuccess ? launch_status() : status_from_cuda(e);
}

int launch_cg_grad_mm(const float* penc, const float* pdec, const CgFactors& F, const int32_t* labels,
                      const int32_t* act_lens, const int32_t* label_lens, int B, int T, int U1, int V, int blank,
                      const float* lse, const int32_t* alpha, const int32_t* beta, GradCosts grad_costs,
                      float* d_penc, float* d_pdec, float* partial, cudaStream_t stream) {
#define RNNT_MM(NTW, WIDE)                                                                                         \
    return launch_grad_mm<NTW, WIDE>(penc, pdec, F, labels, act_lens, label_lens, B, T, U1, V, blank, lse, alpha, \
                                     beta, grad_costs, d_penc, d_pdec, partial, stream)
    if (V <= 32) RNNT_MM(1, false);
    if (V <= 96) RNNT_MM(3, false);
    if (V <= 128) RNNT_MM(4, false);
    RNNT_MM(4, true);
#undef RNNT_MM
}

}  // namespace rnntb200
