// lattice.cu -- alpha / beta recursions over the T x U1 alignment lattice (SURVEY.md 8(a) maths).
//
// Replaces warp-transducer's compute_alphas_kernel / compute_betas_kernel and torchaudio's
// ComputeAlphasBetasCosts -- written from the recursion, not from
// either implementation.
//
// Mapping.  ONE launch, grid (B, 2): blockIdx.y = 0 sweeps alpha, 1 sweeps beta, so both
// directions of every utterance are in flight together.  One thread per label position walks the
// anti-diagonals d = t + u; the working diagonal lives in registers and is handed to the u+1
// neighbour with warp shuffles.  A CTA has W = ceil(U1/32) warps; warp w runs kLag diagonals
// behind warp w-1, so the value crossing a warp boundary is produced kLag+1 steps before it is
// consumed and travels through a small shared-memory ring that needs a block barrier only every
// kLag steps (not one per diagonal).
//
// Arithmetic.  The recursion is linear in the probability domain,
//     alpha(t,u) = alpha(t-1,u) * P_blank(t-1,u) + alpha(t,u-1) * P_label(t,u-1),
// and is evaluated there with every quantity held as (mantissa in [1,2), integer exponent): two
// FMULs, one FFMA and a handful of integer ops on the dependent chain, no MUFU, no overflow or
// underflow for any lattice size, ~1e-7 relative error per step (the log-domain form costs two
// MUFUs per step on the chain and loses ulp(|alpha|) ~ 1e-4 per step once |alpha| reaches 10^3).
//
// Storage.  alpha / beta planes are written as Q16 fixed-point base-2 logs (int32, value =
// log2(alpha) * 65536): 4 bytes per cell like fp32 but with 1.5e-5 resolution over +-32767, so the
// occupancy alpha + beta - log P(y|x) the gradient needs is formed exactly in integers instead of
// cancelling three fp32 numbers of magnitude 10^3.  beta[b,0,0] is log2 P(y|x) in the same format.
//
// Loads.  The (lp_blank, lp_label) pair of a cell is one 8-byte cp.async into a per-thread slot
// of a shared-memory ring, issued kDepth-1 diagonals ahead (the loads do not depend on the
// recursion), so HBM latency is off the dependent chain.
#include "common.cuh"

namespace rnntb200 {

namespace {

constexpr int kDepth = 16;      // cp.async ring depth (diagonals in flight per thread)
constexpr int kLag = 8;         // diagonals warp w trails warp w-1
constexpr int kEdgeRing = 32;   // >= 2*kLag + 1 slots for warp-boundary values
constexpr int kZeroExp = -(1 << 30);  // exponent of the (mantissa, exponent) encoding of 0

struct ME {  // value = m * 2^e, m in [1,2) after normalize(), or m == 0
    float m;
    int e;
};

__device__ __forceinline__ ME me_zero() { return ME{0.f, kZeroExp}; }

// log-probability (natural log, <= 0) -> (mantissa in [1,2), exponent)
__device__ __forceinline__ ME me_from_log(float lp) {
    float x = fmaxf(lp * kLog2e, -16384.f);
    const float fl = floorf(x);
    return ME{fast_ex2(x - fl), (int)fl};
}

__device__ __forceinline__ ME me_mul(ME a, ME b) { return ME{a.m * b.m, a.e + b.e}; }

// a + b for mantissas that are 0 or in [1,4); result mantissa in [1,8) (or 0), not normalised
__device__ __forceinline__ ME me_add(ME a, ME b) {
    const int dd = b.e - a.e;
    const bool b_big = dd > 0;
    const int k = min(abs(dd), 127);
    const float s = __int_as_float((127 - k) << 23);  // 2^-k, and +0.0 when k == 127
    const float big = b_big ? b.m : a.m, small = b_big ? a.m : b.m;
    return ME{fmaf(small, s, big), max(a.e, b.e)};
}

__device__ __forceinline__ ME me_normalize(ME a) {
    const int bits = __float_as_int(a.m);
    const bool z = a.m == 0.f;
    return ME{z ? 0.f : __int_as_float((bits & 0x007fffff) | 0x3f800000),
              z ? kZeroExp : a.e + (bits >> 23) - 127};
}

// Q16 fixed-point log2 of a normalised value
__device__ __forceinline__ int me_to_q16(ME a) {
    const int e = max(min(a.e, 32766), -32767);
    return e * 65536 + __float2int_rn(fast_lg2(a.m) * 65536.f);
}

__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gmem_src) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <bool kMultiWarp>
__global__ void __launch_bounds__(1024, 1)
lattice_sweep_kernel(const float2* __restrict__ lp2, const int32_t* __restrict__ act_lens,
                     const int32_t* __restrict__ label_lens, int T, int U1,
                     int32_t* __restrict__ alpha, int32_t* __restrict__ beta,
                     float* __restrict__ costs, float* __restrict__ ll_alpha) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* ring = reinterpret_cast<float2*>(smem_raw);  // [kDepth][blockDim.x]
    __shared__ float edge_m[kEdgeRing][32];
    __shared__ int edge_e[kEdgeRing][32];

    const int b = blockIdx.x;
    const int dir = blockIdx.y;  // 0: alpha (forward), 1: beta (backward)
    const int j = threadIdx.x;   // position along the sweep: u = j (alpha) or U_b - j (beta)
    const int lane = j & 31, warp = j >> 5;
    const int nthreads = blockDim.x;

    int Tb = act_lens[b], Ub = label_lens[b];
    Tb = min(max(Tb, 1), T);
    Ub = min(max(Ub, 0), U1 - 1);
    const int U1b = Ub + 1;
    const int D = Tb + Ub;  // anti-diagonals of this utterance
    const bool lane_on = j < U1b;
    const int u = dir == 0 ? j : Ub - j;
    const size_t base = (size_t)b * T * U1;
    const float2* src = lp2 + base;
    int32_t* dst = (dir == 0 ? alpha : beta) + base;
    const int lag = kMultiWarp ? warp * kLag : 0;
    const int n_warps_on = (U1b + 31) >> 5;
    // every warp runs the same number of steps (uniform barriers), rounded to the barrier period
    int S = D + (kMultiWarp ? (n_warps_on - 1) * kLag : 0);
    if (kMultiWarp) S = (S + kLag - 1) / kLag * kLag;

    // cell index at local progress tau (tau = d - j): row t = tau (alpha) or T_b-1-tau (beta)
    auto cell = [&](int tau) -> int { return (dir == 0 ? tau : Tb - 1 - tau) * U1 + u; };
    auto prefetch = [&](int s) {  // cell this thread consumes at local step s
        const int tau = s - lag - j;
        if (lane_on && tau >= 0 && tau < Tb) cp_async_8(&ring[(s % kDepth) * nthreads + j], src + cell(tau));
        cp_async_commit();
    };

#pragma unroll 1
    for (int s = 0; s < kDepth - 1; ++s) prefetch(s);

    ME own = me_zero();    // alpha: alpha(t-1,u) * P_blank(t-1,u);   beta: beta(t+1,u)
    ME share = me_zero();  // alpha: alpha(t,u)   * P_label(t,u);     beta: beta(t,u)

#pragma unroll 1
    for (int s = 0; s < S; ++s) {
        if (kMultiWarp && (s % kLag) == 0) __syncthreads();
        const int d = s - lag;
        const int tau = d - j;
        const bool on = lane_on && tau >= 0 && tau < Tb;

        // hand-off from the u-1 neighbour (its value on diagonal d-1)
        ME in;
        in.m = __shfl_up_sync(0xffffffffu, share.m, 1);
        in.e = __shfl_up_sync(0xffffffffu, share.e, 1);
        if (lane == 0) {
            if (kMultiWarp && warp > 0) {
                const int slot = (d - 1) & (kEdgeRing - 1);
                in.m = edge_m[slot][warp - 1];
                in.e = edge_e[slot][warp - 1];
            } else {
                in = me_zero();
            }
        }

        cp_async_wait<kDepth - 2>();
        const float2 lp = on ? ring[(s % kDepth) * nthreads + j] : make_float2(0.f, 0.f);
        prefetch(s + kDepth - 1);  // refills the slot consumed one step ago

        const ME pb = me_from_log(lp.x), pl = me_from_log(lp.y);
        ME val;
        if (dir == 0) {
            val = me_normalize(me_add(own, in));
            if (tau == 0 && j == 0) val = ME{1.f, 0};
            own = me_mul(val, pb);
            share = me_mul(val, pl);
        } else {
            val = me_normalize(me_add(me_mul(own, pb), me_mul(in, pl)));
            if (tau == 0 && j == 0) val = pb;
            own = val;
            share = val;
        }
        if (!on) {
            own = me_zero();
            share = me_zero();
        } else {
            dst[cell(tau)] = me_to_q16(val);
            if (j == Ub && tau == Tb - 1) {
                if (dir == 0) {  // alpha(T-1,U) * P_blank(T-1,U)
                    const ME f = me_normalize(own);
                    if (ll_alpha) ll_alpha[b] = (float)(((double)f.e + (double)fast_lg2(f.m)) * 0.6931471805599453);
                } else {         // beta(0,0) = P(y|x)
                    costs[b] = (float)(-((double)val.e + (double)fast_lg2(val.m)) * 0.6931471805599453);
                }
            }
        }
        if (kMultiWarp && lane == 31) {
            const int slot = d & (kEdgeRing - 1);
            edge_m[slot][warp] = share.m;
            edge_e[slot][warp] = share.e;
        }
    }
    cp_async_wait<0>();
}

}  // namespace

int launch_lattice_sweep(const float2* lp2, const int32_t* act_lens, const int32_t* label_lens, int B,
                         int T, int U1, int32_t* alpha, int32_t* beta, float* costs, float* ll_alpha,
                         cudaStream_t stream) {
    if (B == 0) return RNNTB200_STATUS_SUCCESS;
    if (U1 > 1024) return RNNTB200_STATUS_INVALID_VALUE;
    const int threads = ((U1 + 31) / 32) * 32;
    const size_t smem = (size_t)kDepth * threads * sizeof(float2);  // <= 128 KiB at U1 = 1024
    dim3 grid(B, 2);
    if (threads <= 32) {
        lattice_sweep_kernel<false><<<grid, threads, smem, stream>>>(lp2, act_lens, label_lens, T, U1,
                                                                      alpha, beta, costs, ll_alpha);
    } else {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(lattice_sweep_kernel<true>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return status_from_cuda(e);
        }
        lattice_sweep_kernel<true><<<grid, threads, smem, stream>>>(lp2, act_lens, label_lens, T, U1,
                                                                     alpha, beta, costs, ll_alpha);
    }
    return launch_status();
}

}  // namespace rnntb200
