// lattice.cu -- alpha / beta recursions over the T x U1 alignment lattice (SURVEY.md 8(a) maths).
//
// Replaces warp-transducer's compute_alphas_kernel / compute_betas_kernel and torchaudio's
// ComputeAlphasBetasCosts -- written from the recursion, not from
// either implementation.
//
// Mapping: ONE launch, grid (B, 2): blockIdx.y = 0 sweeps alpha, 1 sweeps beta, so both directions
// of every utterance are in flight together.  One thread per label position u walks the
// anti-diagonals d = t + u; the working diagonal lives in registers, the hand-off to the u+1
// neighbour is a warp shuffle, and only warp-boundary values cross shared memory (double-buffered,
// one barrier per diagonal).  State is kept in the base-2 log domain so the dependent chain per
// diagonal is FADD -> SHFL -> FMNMX/FADD -> MUFU.EX2 -> FADD -> MUFU.LG2 -> FADD.
// The (lp_blank, lp_label) pair of a cell is one 8-byte load, prefetched 4 diagonals ahead into a
// register ring (the loads do not depend on the recursion).
#include "common.cuh"

namespace rnntb200 {

namespace {

constexpr int kPrefetch = 4;

template <bool kMultiWarp>
__global__ void __launch_bounds__(1024, 1)
lattice_sweep_kernel(const float2* __restrict__ lp2, const int32_t* __restrict__ act_lens,
                     const int32_t* __restrict__ label_lens, int T, int U1,
                     float* __restrict__ alpha, float* __restrict__ beta,
                     float* __restrict__ costs, float* __restrict__ ll_alpha) {
    __shared__ float edge[2][32];
    const int b = blockIdx.x;
    const int dir = blockIdx.y;  // 0: alpha (forward), 1: beta (backward)
    const int j = threadIdx.x;   // position along the sweep: u = j (alpha) or U_b - j (beta)
    const int lane = j & 31, warp = j >> 5;

    int Tb = act_lens[b], Ub = label_lens[b];
    Tb = min(max(Tb, 1), T);
    Ub = min(max(Ub, 0), U1 - 1);
    const int U1b = Ub + 1;
    const int D = Tb + U1b - 1;  // number of anti-diagonals
    const bool lane_on = j < U1b;
    const int u = dir == 0 ? j : Ub - j;
    const size_t base = (size_t)b * T * U1;
    const float2* src = lp2 + base;
    float* dst = (dir == 0 ? alpha : beta) + base;

    // cell index at local progress tau (tau = d - j): row t = tau (alpha) or T_b-1-tau (beta)
    auto cell = [&](int tau) -> int { return (dir == 0 ? tau : Tb - 1 - tau) * U1 + u; };

    float2 ring[kPrefetch];
#pragma unroll
    for (int k = 0; k < kPrefetch; ++k) {
        const int tau = k - j;
        ring[k] = make_float2(0.f, 0.f);
        if (lane_on && tau >= 0 && tau < Tb) ring[k] = __ldg(src + cell(tau));
    }

    float own = kNegInf;  // alpha: alpha(t-1,u)+lp_blank(t-1,u);  beta: beta(t+1,u)
    float in = kNegInf;   // alpha: alpha(t,u-1)+lp_label(t,u-1);  beta: beta(t,u+1)

    auto step = [&](int d, float2& slot) {
        const int tau = d - j;
        const bool on = lane_on && tau >= 0 && tau < Tb;
        const float lb = slot.x * kLog2e, ll = slot.y * kLog2e;
        // prefetch the cell this thread needs kPrefetch diagonals from now
        const int tau_pf = tau + kPrefetch;
        if (lane_on && tau_pf >= 0 && tau_pf < Tb) slot = __ldg(src + cell(tau_pf));

        float val, share;
        if (dir == 0) {
            const float up = tau > 0 ? own : kNegInf;
            const float left = j > 0 ? in : kNegInf;
            val = (tau == 0 && j == 0) ? 0.f : logaddexp2(up, left);
            own = val + lb;
            share = val + ll;
        } else {
            const float down = tau > 0 ? own + lb : kNegInf;
            const float right = j > 0 ? in + ll : kNegInf;
            val = (tau == 0 && j == 0) ? lb : logaddexp2(down, right);
            own = val;
            share = val;
        }
        if (on) {
            dst[cell(tau)] = val * kLn2;
            if (j == Ub && tau == Tb - 1) {
                if (dir == 0) {
                    if (ll_alpha) ll_alpha[b] = own * kLn2;  // alpha(T-1,U) + lp_blank(T-1,U)
                } else {
                    costs[b] = -val * kLn2;  // -beta(0,0)
                }
            }
        }
        // hand the value to the u+1 neighbour for the next diagonal
        if (kMultiWarp) {
            if (lane == 31) edge[d & 1][warp] = share;
            __syncthreads();
        }
        in = __shfl_up_sync(0xffffffffu, share, 1);
        if (kMultiWarp && lane == 0 && warp > 0) in = edge[d & 1][warp - 1];
    };

    for (int d0 = 0; d0 < D; d0 += kPrefetch) {
#pragma unroll
        for (int k = 0; k < kPrefetch; ++k) step(d0 + k, ring[k]);
    }
}

}  // namespace

int launch_lattice_sweep(const float2* lp2, const int32_t* act_lens, const int32_t* label_lens, int B,
                         int T, int U1, float* alpha, float* beta, float* costs, float* ll_alpha,
                         cudaStream_t stream) {
    if (B == 0) return RNNTB200_STATUS_SUCCESS;
    if (U1 > 1024) return RNNTB200_STATUS_INVALID_VALUE;
    const int threads = ((U1 + 31) / 32) * 32;
    dim3 grid(B, 2);
    if (threads <= 32)
        lattice_sweep_kernel<false><<<grid, threads, 0, stream>>>(lp2, act_lens, label_lens, T, U1,
                                                                   alpha, beta, costs, ll_alpha);
    else
        lattice_sweep_kernel<true><<<grid, threads, 0, stream>>>(lp2, act_lens, label_lens, T, U1,
                                                                  alpha, beta, costs, ll_alpha);
    return launch_status();
}

}  // namespace rnntb200
