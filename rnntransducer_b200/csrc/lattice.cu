// lattice.cu -- alpha / beta recursions over the T x U1 alignment lattice (SURVEY.md 8(a) maths).
//
// Replaces warp-transducer's compute_alphas_kernel / compute_betas_kernel and torchaudio's
// ComputeAlphasBetasCosts -- written from the recursion, not from
// either implementation.
//
// Common to every kernel of this file.  ONE launch, grid (B, 2): blockIdx.y = 0 sweeps alpha, 1 sweeps beta,
// so both directions of every utterance are in flight together.  One thread per label position walks the
// anti-diagonals d = t + u; the working diagonal lives in registers and is handed to the u+1 neighbour with
// warp shuffles; warp w runs a few diagonals behind warp w-1 and the value crossing a warp boundary travels
// through a small shared-memory ring.  The recursion is linear in the probability domain,
//     alpha(t,u) = alpha(t-1,u) * P_blank(t-1,u) + alpha(t,u-1) * P_label(t,u-1),
// and is evaluated there with every quantity held as (fp32 mantissa, integer exponent): no MUFU on the chain,
// no overflow or underflow in the recursion for any lattice size, ~1e-7 relative error per step (the
// log-domain form costs two MUFUs per step on the chain and loses ulp(|alpha|) ~ 1e-4 per step once |alpha|
// reaches 10^3).  alpha / beta planes are written in a 32-bit wide-exponent float ("e16m16", common.cuh):
// signed 16-bit binary exponent in the high half, 16 mantissa bits in the low half -- 4 bytes per cell like
// fp32, 2^-17 relative precision at ANY magnitude (|log2| < 32768; beyond, the cost is reported as +inf, see
// cost_of), so the occupancy alpha * beta / P(y|x) the gradient needs is formed from exact integer exponents
// instead of cancelling three fp32 logs of magnitude 10^3.  beta[b,0,0] is P(y|x) in that format.
//
// The file, top to bottom:
//   1. the single-role sweep of round 1 (`lattice_sweep_kernel`: one warp per 32 positions does everything,
//      every lane touches its own lattice row) -- last fall-back, RNNTB200_SWEEP_LEGACY;
//   2. the warp-specialised sweep of round 1 (`lattice_sweep_ws_kernel`: a chain warp + loader + two converters
//      + consumer per 32 positions, mbarrier rings, cluster bands) -- fall-back, RNNTB200_SWEEP=ws;
//   3. the decoupled (mantissa | exponent) recursion both newer kernels use (`ws_chain`, `tp_sweep`): the
//      exponents follow an integer max-plus recurrence that runs one step ahead of the mantissas;
//   4. the fused reduction of the costs (last-arriving utterance sums them in index order);
//   5. THE DEFAULT: the self-contained sweep (`lattice_sweep_tp_kernel`): one warp per 32 positions again, but
//      with row-wise global accesses through two lane-private FIFOs, 24 KB per warp, thread-block-cluster
//      bands for long label sequences;
//   6. the dispatch (`launch_lattice_sweep`) with the measurements that decide it.
#include <cstdlib>

#include "tc_common.cuh"

namespace rnntb200 {

namespace {

constexpr int kDepth = 16;        // cp.async ring depth (diagonals in flight per thread)
constexpr int kUnroll = 8;        // steps per loop body = half a ring = kLag
constexpr int kRingStride = 17;   // float2 slots per thread (+1 pad: conflict-free 8-byte accesses)
constexpr int kLag = 8;           // diagonals warp w trails warp w-1
constexpr int kEdgeRing = 32;     // >= 2*kLag + 1 slots for warp-boundary values
constexpr int kLagX = 32;         // diagonals the first warp of a CTA trails the last warp of the previous CTA
constexpr int kEdgeRingX = 128;   // >= 2*kLagX + 1 slots for the value crossing CTAs (cluster mode)
constexpr int kZeroExp = -(1 << 29);  // (1, kZeroExp) stands for 0: it never wins an addition

struct ME {  // value = m * 2^e; m in [1,2) after normalize(), in [0.7, 5.7) before
    float m;
    int e;
};

// log-probability (natural log, <= 0) -> (mantissa in [0.707, 1.414], exponent): round-to-nearest
// split with the 1.5 * 2^23 trick (no F2I / FRND on the path), one MUFU.EX2
__device__ __forceinline__ ME me_from_log(float lp) {
    const float x = fmaxf(lp * kLog2e, -16000.f);
    const float t = x + 12582912.f;
    return ME{fast_ex2(x - (t - 12582912.f)), __float_as_int(t) - 0x4B400000};
}

__device__ __forceinline__ ME me_mul(ME a, ME b) { return ME{a.m * b.m, a.e + b.e}; }

// a + b: the term with the smaller exponent is scaled by 2^-(exponent gap) (0 once the gap >= 127)
__device__ __forceinline__ ME me_add(ME a, ME b) {
    const int dd = b.e - a.e;
    const bool b_big = dd > 0;
    const int k = min(abs(dd), 127);
    const float s = __int_as_float((127 - k) << 23);
    const float big = b_big ? b.m : a.m, small = b_big ? a.m : b.m;
    return ME{fmaf(small, s, big), max(a.e, b.e)};
}

__device__ __forceinline__ ME me_normalize(ME a) {
    const int bits = __float_as_int(a.m);
    return ME{__int_as_float((bits & 0x007fffff) | 0x3f800000), a.e + (bits >> 23) - 127};
}

__device__ __forceinline__ int me_pack(ME a) {  // a normalised -> e16m16, mantissa rounded to nearest
    const int rb = __float_as_int(a.m) + 0x40;
    const int e = max(min(a.e + (rb >> 23) - 127, 32767), -32767);
    return (e << 16) | ((rb >> 7) & 0xFFFF);
}

__device__ __forceinline__ double me_ln(ME a) {
    return ((double)a.e + (double)log2f(a.m)) * 0.6931471805599453;
}

// cost = -ln P(y|x).  The e16m16 planes hold |log2| < 32768 (utterances costing < 22 700 nats); beyond
// that me_pack saturates and the gradients formed from the planes would be silently wrong, so the
// cost is reported as +inf instead (the loss and every gradient of the step then show it).
__device__ __forceinline__ float cost_of(ME p) {
    return p.e < -32766 ? __int_as_float(0x7f800000) : (float)(-me_ln(p));
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

// DIR 0: alpha (u = j, rows walked upwards); DIR 1: beta (u = U_b - j, rows walked downwards).
// The loop is software-pipelined: the (lp_blank, lp_label) pair of step s+1 is read from the ring
// and split into (mantissa, exponent) form during step s, so the only work between receiving the
// neighbour's value and handing the new value on is add -> normalise -> multiply.
// There is no special-casing of the lattice borders inside the loop: the ring starts zeroed
// (log-prob 0 = factor 1) and every absent term is the "zero" (1, kZeroExp), which loses every
// addition, so a thread that has not reached its first cell yet just carries zeros along.
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_nctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_barrier() {  // release / acquire at cluster scope
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store {x, y} at the same shared-memory offset in CTA `rank` of the cluster (DSMEM)
__device__ __forceinline__ void st_cluster_v2(const void* local_smem, unsigned rank, int x, int y) {
    unsigned laddr = (unsigned)__cvta_generic_to_shared(local_smem), raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(rank));
    asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(raddr), "r"(x), "r"(y) : "memory");
}

// kMode 0: one warp.  1: several warps of one CTA (block barrier every kLag diagonals).
// 2: the warps of one sweep are spread over a thread-block cluster (<= 4 warps per CTA, <= 8 CTAs):
//    a long label sequence (U1 = 301 is ten warps) then issues from several SMs instead of
//    queueing on the four schedulers of one.  The warp-boundary value crosses CTAs through
//    distributed shared memory (written into the consumer's ring) and the barrier is the cluster's;
//    the kLag-diagonal skew between consecutive warps hides the DSMEM latency.
template <int DIR, int kMode>
__device__ __forceinline__ void sweep(const float2* __restrict__ lp2, int Tb, int Ub, int T, int U1, int b,
                                      int32_t* __restrict__ out, float* __restrict__ costs,
                                      float* __restrict__ ll_alpha, float2* ring, int2 (*edge)[33], int2* xedge) {
    constexpr bool kMultiWarp = kMode != 0;
    const unsigned rank = kMode == 2 ? cluster_ctarank() : 0, n_rank = kMode == 2 ? cluster_nctarank() : 1;
    const int wl = threadIdx.x >> 5, wpc = blockDim.x >> 5;  // warp within the CTA, warps per CTA
    const int j = rank * blockDim.x + threadIdx.x;            // position along the sweep
    const int lane = j & 31, warp = j >> 5;
    const int U1b = Ub + 1;
    const bool lane_on = j < U1b;
    const unsigned Tb_eff = lane_on ? Tb : 0;  // (unsigned)tau < Tb_eff  <=>  this thread has a cell
    const int u = DIR == 0 ? j : Ub - j;
    // consecutive warps of a CTA run kLag diagonals apart (block barrier every kLag steps); across a
    // CTA boundary the skew is kLagX and the (expensive) cluster barrier comes every kLagX steps
    const int lag = !kMultiWarp ? 0 : warp * kLag + (kMode == 2 ? (int)rank * (kLagX - kLag) : 0);
    const int n_warps_on = (U1b + 31) >> 5;
    const int max_lag = !kMultiWarp ? 0
                        : (n_warps_on - 1) * kLag + (kMode == 2 ? ((n_warps_on - 1) / wpc) * (kLagX - kLag) : 0);
    // every warp runs the same number of steps (uniform barriers), rounded up to the barrier period
    const int round_to = kMode == 2 ? kLagX : kUnroll;
    const int S = (Tb + Ub + max_lag + round_to - 1) / round_to * round_to;

    // this thread's cell at progress tau: row t = tau (alpha) or T_b-1-tau (beta)
    const int stride = DIR == 0 ? U1 : -U1;
    const size_t first = (size_t)b * T * U1 + (size_t)(DIR == 0 ? 0 : Tb - 1) * U1 + u;
    int tau = -lag - j;                 // progress at step 0
    int off = tau * stride;             // cell offset (elements) of the current step from `first`
    const float2* src = lp2 + first;    // + off : cell consumed at the current step
    const float2* src_pf = src + (long long)(kDepth - 1) * stride;  // + off : cell being prefetched
    int32_t* dst = out + first;
    float2* cur = ring + (size_t)threadIdx.x * kRingStride;  // ring half of steps s0 .. s0+7
    float2* oth = cur + kUnroll;                             // ring half of the next 8 steps
    // where lane 0 finds its neighbour's value: column wl-1 (written by the previous warp of this
    // CTA), the xedge ring (written through DSMEM by the last warp of the previous CTA), or column 32 (zero)
    const int edge_col = !kMultiWarp ? 32 : wl > 0 ? wl - 1 : 32;
    const bool from_x = kMode == 2 && wl == 0 && rank > 0;

    // zero ring (factor 1 for steps without a cell), "zero" edge values
#pragma unroll
    for (int k = 0; k < kDepth; ++k) cur[k] = make_float2(0.f, 0.f);
    if (kMultiWarp)
        for (int i = threadIdx.x; i < kEdgeRing * 33; i += blockDim.x) edge[i / 33][i % 33] = make_int2(0x3f800000, kZeroExp);
    if (kMode == 2)
        for (int i = threadIdx.x; i < kEdgeRingX; i += blockDim.x) xedge[i] = make_int2(0x3f800000, kZeroExp);

    // prologue: cells of steps 0 .. kDepth-2 (slot of step s = s mod 16: cur[0..7], oth[0..6])
#pragma unroll
    for (int k = 0; k < kDepth - 1; ++k) {
        if ((unsigned)(tau + k) < Tb_eff) cp_async_8(k < kUnroll ? cur + k : oth + (k - kUnroll), src + (off + k * stride));
        cp_async_commit();
    }
    int es = (-lag - 1) & (kEdgeRing - 1);  // edge slot of diagonal d-1
    int ex = (-lag - 1) & (kEdgeRingX - 1);  // same in the cross-CTA ring

    // alpha: own = alpha(t-1,u) P_blank(t-1,u), share = alpha(t,u) P_label(t,u); the first cell gets
    // alpha(0,0) = 1 as own.  beta: own = share = beta(t+1,u) / beta(t,u+1); the first cell gets 1 so
    // that beta(T-1,U) = 1 * P_blank.
    ME own{1.f, j == 0 ? 0 : kZeroExp};
    ME share{1.f, kZeroExp};
    ME last{1.f, kZeroExp};                               // alpha(T-1,U) P_blank(T-1,U) / beta(0,0)
    const int t_last = (lane_on && j == Ub) ? Tb - 1 : -1;  // progress at which this thread is there

    cp_async_wait<kDepth - 2>();
    ME pb = me_from_log(cur[0].x), pl = me_from_log(cur[0].y);  // factors of step 0

#pragma unroll 1
    for (int s0 = 0; s0 < S; s0 += kUnroll) {
        // kLag == kUnroll: one barrier per 8 diagonals
        if (kMode == 1) __syncthreads();
        if (kMode == 2) {
            if ((s0 & (kLagX - 1)) == 0) cluster_barrier();
            else __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) {
            // hand-off from the u-1 neighbour (its value on diagonal d-1)
            ME in;
            in.m = __shfl_up_sync(0xffffffffu, share.m, 1);
            in.e = __shfl_up_sync(0xffffffffu, share.e, 1);
            int2 ev = kMultiWarp ? edge[es][edge_col] : make_int2(0x3f800000, kZeroExp);
            if (kMode == 2 && from_x) ev = xedge[ex];

            // off the dependent chain: refill the ring, fetch and split the factors of step s+1
            if ((unsigned)(tau + kDepth - 1) < Tb_eff) cp_async_8(k == 0 ? oth + kUnroll - 1 : cur + k - 1, src_pf + off);
            cp_async_commit();
            cp_async_wait<kDepth - 2>();
            const float2 lpn = k + 1 < kUnroll ? cur[k + 1] : oth[0];
            const ME pbn = me_from_log(lpn.x), pln = me_from_log(lpn.y);

            if (lane == 0) in = ME{__int_as_float(ev.x), ev.y};
            ME val;
            if (DIR == 0) {
                val = me_normalize(me_add(own, in));
                own = me_mul(val, pb);
                share = me_mul(val, pl);
            } else {
                val = me_normalize(me_add(me_mul(own, pb), me_mul(in, pl)));
                own = val;
                share = val;
            }
            const int packed = me_pack(val);
            if ((unsigned)tau < Tb_eff) dst[off] = packed;
            if (tau == t_last) last = DIR == 0 ? own : val;  // terminal cell (t_last = -1 elsewhere)
            if (kMultiWarp) {
                es = (es + 1) & (kEdgeRing - 1);  // now the slot of diagonal d
                if (kMode == 2) ex = (ex + 1) & (kEdgeRingX - 1);
                if (lane == 31) {
                    if (kMode == 1 || wl < wpc - 1) edge[es][wl] = make_int2(__float_as_int(share.m), share.e);
                    else if (rank + 1 < n_rank) st_cluster_v2(&xedge[ex], rank + 1, __float_as_int(share.m), share.e);
                }
            }
            pb = pbn;
            pl = pln;
            ++tau;
            off += stride;
        }
        float2* tmp = cur;
        cur = oth;
        oth = tmp;
    }
    cp_async_wait<0>();
    if (kMode == 2) cluster_barrier();  // no CTA may exit while a neighbour can still write into its ring
    if (t_last >= 0) {
        if (DIR == 0) {
            if (ll_alpha) ll_alpha[b] = (float)me_ln(me_normalize(last));
        } else {
            costs[b] = cost_of(last);
        }
    }
}

template <int kMode>
__global__ void __launch_bounds__(kMode == 2 ? 128 : 1024, 1)
lattice_sweep_kernel(const float2* __restrict__ lp2, const int32_t* __restrict__ act_lens,
                     const int32_t* __restrict__ label_lens, int T, int U1,
                     int32_t* __restrict__ alpha, int32_t* __restrict__ beta,
                     float* __restrict__ costs, float* __restrict__ ll_alpha) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* ring = reinterpret_cast<float2*>(smem_raw);  // [blockDim.x][kRingStride]
    __shared__ int2 edge[kEdgeRing][33];
    __shared__ int2 xedge[kMode == 2 ? kEdgeRingX : 1];
    const int b = kMode == 2 ? blockIdx.x / cluster_nctarank() : blockIdx.x;
    const int Tb = len_T(act_lens, b, T);
    const int Ub = len_U(label_lens, b, U1);
    if (blockIdx.y == 0)
        sweep<0, kMode>(lp2, Tb, Ub, T, U1, b, alpha, costs, ll_alpha, ring, edge, xedge);
    else
        sweep<1, kMode>(lp2, Tb, Ub, T, U1, b, beta, costs, ll_alpha, ring, edge, xedge);
}


// =================================================================================================
// Warp-specialised sweep (label sequences of up to 128 positions, i.e. up to four chain warps).
//
// At small batch sizes the sweep is bound by the latency of ONE warp walking T + U dependent steps,
// and two thirds of that warp's instructions are not the recursion at all: fetching and splitting
// the log-probabilities, packing and storing the result, address and predicate arithmetic.  Here
// every chain warp has four helper warps (see the roles below):
//   chain   factors (one 16-byte load) -> hand-off -> add / normalise / multiply -> value (one 8-byte
//           store); nothing else: 24 instructions, ~94 cycles per step when run alone
// Rings are handed over in blocks of 8 steps through mbarriers (one elected arrival per warp after
// a __syncwarp: 32 arrivals on one address serialise); the chain warps of one CTA keep the 8-step
// skew and meet at a named barrier that the helpers never join.
// Measured at cfg 2 (B=32, T=400, U=80): chain alone 29 us; + hand-shakes 35 us; + helper work 53 us.
constexpr int kWsStages = 3;
constexpr int kHalf = 8;        // steps whose loads / arithmetic / stores are batched for instruction-level parallelism
constexpr int kWsEdgeRing = 64; // >= 2 * KB + 1 slots for warp-boundary values
constexpr int kWin = 64;        // rows of the consumer's shared-memory window (power of two)
constexpr int kRowBars = 16;    // ring of row-block barriers (> blocks the loader may run ahead)
constexpr int kBandSkew = 32;   // steps the first chain warp of a CTA trails the last one of the previous CTA (cluster)

// Label sequences beyond four chain warps: the sweep is cut into BANDS of three chain warps, one CTA of a
// thread-block cluster each.  The value crossing a band boundary is forwarded by the consumer helper
// of the band's last chain warp straight into the next CTA's shared memory (st.shared::cluster into
// `xedge`, one slot per step of the whole sweep: no ring, hence no back-pressure), followed by a
// release store of the finished block count (`xdone`) that the receiving chain warp polls locally.
struct WsBand {
    int band, n_bands;   // this CTA's rank in the cluster
    int xlag;            // extra skew of this band: band * (kBandSkew - KB)
    int2* xedge;         // [n_blocks * KB] boundary values from the previous band (local shared memory)
    int* xdone;          // blocks the previous band's last warp has forwarded (local shared memory)
};

// KB = steps per hand-over block = skew between consecutive chain warps; RW = lattice rows in the
// loader's window.
template <int KB, int RW>
struct WsWarp {  // shared memory of one chain warp and its helpers
    uint4 fac[kWsStages][KB][32];             // (m_blank, e_blank, m_label, e_label) per step and lane
    int2 val[kWsStages][KB][32];              // normalised lattice value per step and lane
    float2 raw[RW * 32];                      // producer: window of RW lp2 rows (this warp's 32 columns), filled by cp.async
    int32_t out[kWin * 32];                   // consumer: window of packed output rows
    unsigned long long bars[4 * kWsStages];   // full, empty (factors); vfull, vempty (values)
    unsigned long long rows[kRowBars];        // "the lp2 rows of block b have landed" (loader -> converters)
    int chain_done;                           // blocks the chain has finished (the loader's window-reuse throttle: a
                                              // plain counter, because an mbarrier parity cannot name a phase that
                                              // lies more than one completion back)
};
// ---- decoupled (mantissa | exponent) recursion -------------------------------------------------------
// A lattice value is m * 2^E, E an int and m an fp32 that is NOT kept normalised: between two
// renormalisations (one per block of KB steps) it drifts by at most 2^(1.5 KB), far inside fp32 range.
// The exponents then obey a pure integer max-plus recurrence,
//     E(s) = max(E_own(s-1) + e_blank, E_in(s-1) + e_label),
// that does not depend on the mantissas at all, so it runs ONE STEP AHEAD of the mantissa recurrence
// and the alignment scales 2^(E_term - E) are in registers before the neighbour's mantissa arrives.
// Dependent chain of a step: FMUL -> SHFL -> FFMA (alpha) / SHFL -> FFMA (beta), with IADD -> SHFL ->
// IMNMX running beside it -- instead of SHFL -> exponent compare -> align -> FFMA -> normalise -> FMUL.
// No select on the chain either: lane 0 (no neighbour inside the warp) gets a bias on the shuffled
// exponent so that term's scale is exactly 0, and the value crossing a warp boundary enters as a third
// term whose exponent is "minus infinity" on every other lane.
constexpr int kNoTerm = kZeroExp;  // exponent (or exponent bias) of an absent term; sums of two of these and a
                                   // sweep's worth of factor exponents stay inside int32

__device__ __forceinline__ float pow2_neg(int d) {  // 2^-d for d >= 0, exactly 0 once d >= 127
    return __int_as_float((127 - min(d, 127)) << 23);
}

template <int DIR, bool kMulti, int KB, int RW>
__device__ __forceinline__ void ws_chain(WsWarp<KB, RW>& W, int2 (*edge)[33], int w, int nw, int lane, int n_blocks,
                                         const WsBand& X) {
    const uint32_t bars = tc::smem_u32(W.bars);
    const int wg = X.band * nw + w;  // position of this warp in the whole sweep
    const int j = wg * 32 + lane;
    const int lag = kMulti ? wg * KB + X.xlag : 0;
    const int edge_col = (kMulti && w > 0) ? w - 1 : 32;
    const bool from_band = kMulti && w == 0 && X.band > 0;  // lane 0's neighbour lives in the previous CTA
    const bool has_edge = kMulti && lane == 0;
    const uint32_t xdone = tc::smem_u32(X.xdone);
    const int in_bias = lane == 0 ? kNoTerm : 0;
    int es = (-lag - 1) & (kWsEdgeRing - 1);  // edge slot of diagonal d-1
    // value of the previous step (m, E) and the factors it was / will be multiplied with:
    //   alpha: terms of step k are  val(k-1) p_blank(k-1)  and  shfl(val(k-1) p_label(k-1))
    //   beta:  terms of step k are  val(k-1) p_blank(k)    and  shfl(val(k-1)) p_label(k)
    // alpha's start alpha(0,0) = 1 enters as the own term of step 0 of lane j = 0 (factors 1 "before" it);
    // beta's start is val(-1) = 1 on lane j = 0, so that beta(T-1,U) = 1 * p_blank.
    float m = 1.f;
    int E = j == 0 ? 0 : kZeroExp;
    float pbm_prev = 1.f, plm_prev = 1.f;  // alpha only
    int pbe_prev = 0, ple_prev = DIR == 0 ? kNoTerm : 0;  // alpha: nothing to hand on before the first step
#pragma unroll 1
    for (int blk = 0; blk < n_blocks; ++blk) {
        const int st = blk % kWsStages;
        const uint32_t ph = (blk / kWsStages) & 1;
        if (kMulti) asm volatile("bar.sync 1, %0;" ::"r"(nw * 32) : "memory");  // chain warps only
        tc::mbar_wait(bars + 8 * st, ph);                        // factors of this block are there
        tc::mbar_wait(bars + 8 * (3 * kWsStages + st), ph ^ 1);  // the value slot has been drained
        if (from_band && blk >= kBandSkew / KB) {  // the previous band has forwarded what this block reads
            int done;
            for (;;) {
                asm volatile("ld.acquire.cluster.shared::cta.s32 %0, [%1];" : "=r"(done) : "r"(xdone) : "memory");
                if (done > blk - kBandSkew / KB) break;
                __nanosleep(32);
            }
        }
        // the previous warp is a whole block ahead: the boundary values of this block are already in the
        // ring, fetch them off the dependent chain (lanes other than 0 carry an absent term)
        float evm[KB];
        int evE[KB];
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            int2 ev = make_int2(0x3f800000, kNoTerm);
            if (has_edge) {
                ev = edge[(es + k) & (kWsEdgeRing - 1)][edge_col];
                if (from_band) {
                    const int q = blk * KB + k - lag - 1;  // the previous band's step index of diagonal d-1
                    ev = q >= 0 ? X.xedge[q] : make_int2(0x3f800000, kZeroExp);
                }
            }
            evm[k] = __int_as_float(ev.x);
            evE[k] = ev.y;
        }
        uint4 f = W.fac[st][0][lane];
        // exponent recurrence of step 0 of this block (it cannot run ahead across the renormalisation)
        int En;
        float c_own, c_in, c_edge;
        {
            const int pbe = DIR == 0 ? pbe_prev : (int)f.y, ple = DIR == 0 ? ple_prev : (int)f.w;
            const int oE = E + pbe;
            // (beta's seed val(-1) = 1 on lane j = 0 must not reach lane 1: no shuffled term at the very first step)
            const int iE = __shfl_up_sync(0xffffffffu, DIR == 0 ? E + ple : E, 1) +
                           (DIR == 0 ? in_bias : (blk == 0 ? kNoTerm : in_bias) + ple);
            const int eE = DIR == 0 ? evE[0] : evE[0] + ple;
            En = max(max(oE, iE), eE);
            c_own = pow2_neg(En - oE) * (DIR == 0 ? pbm_prev : __uint_as_float(f.x));
            c_in = pow2_neg(En - iE) * (DIR == 0 ? 1.f : __uint_as_float(f.z));
            c_edge = pow2_neg(En - eE) * (DIR == 0 ? 1.f : __uint_as_float(f.z));
        }
        float shm = DIR == 0 ? m * plm_prev : m;  // what the neighbour receives
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            const uint4 fn = W.fac[st][k + 1 < KB ? k + 1 : k][lane];  // next step's factors, off the chain
            const float in_m = __shfl_up_sync(0xffffffffu, shm, 1);     // mantissa chain: the long-latency hop first
            const float own_term = fmaf(evm[k], c_edge, m * c_own);
            const int Ek = En;
            const float ci = c_in;
            const float pbm = __uint_as_float(f.x), plm = __uint_as_float(f.z);
            const int pbe = (int)f.y, ple = (int)f.w;
            if (k + 1 < KB) {  // exponent recurrence of step k+1, in the shadow of the shuffle above
                const int pbe_n = DIR == 0 ? pbe : (int)fn.y, ple_n = DIR == 0 ? ple : (int)fn.w;
                const int oE = Ek + pbe_n;
                const int iE = __shfl_up_sync(0xffffffffu, DIR == 0 ? Ek + ple_n : Ek, 1) +
                               (DIR == 0 ? in_bias : in_bias + ple_n);
                const int eE = DIR == 0 ? evE[k + 1] : evE[k + 1] + ple_n;
                En = max(max(oE, iE), eE);
                c_own = pow2_neg(En - oE) * (DIR == 0 ? pbm : __uint_as_float(fn.x));
                c_in = pow2_neg(En - iE) * (DIR == 0 ? 1.f : __uint_as_float(fn.z));
                c_edge = pow2_neg(En - eE) * (DIR == 0 ? 1.f : __uint_as_float(fn.z));
            }
            m = fmaf(in_m, ci, own_term);
            E = Ek;
            if (k + 1 == KB) {  // renormalise once per block: the exponent of m moves into E
                const int bits = __float_as_int(m);
                m = __int_as_float((bits & 0x007fffff) | 0x3f800000);
                E += (bits >> 23) - 127;
            }
            W.val[st][k][lane] = make_int2(__float_as_int(m), E);
            shm = DIR == 0 ? m * plm : m;
            if (kMulti) {
                es = (es + 1) & (kWsEdgeRing - 1);  // now the slot of diagonal d
                if (lane == 31) edge[es][w] = make_int2(__float_as_int(shm), DIR == 0 ? E + ple : E);
            }
            pbm_prev = pbm, plm_prev = plm, pbe_prev = pbe, ple_prev = ple;
            f = fn;
        }
        __syncwarp();  // orders every lane's shared-memory traffic before the one arrival below
        if (lane == 0) {
            tc::mbar_arrive(bars + 8 * (kWsStages + st));      // factor slot free
            tc::mbar_arrive(bars + 8 * (2 * kWsStages + st));  // values of this block are there
            asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(tc::smem_u32(&W.chain_done)), "r"(blk + 1) : "memory");
        }
    }
}

// Helper roles of one chain warp.  A single warp retires about one instruction every four cycles
// whatever its instruction-level parallelism, so the ~95-cycle chain step tolerates ~20 helper
// instructions per step and warp: the helper work is cut into four warps.
//   loader     lp2 rows -> cp.async -> shared window; publishes "rows of block b landed"
//   convert x2 one for p(blank), one for p(label): window (diagonal read) -> (mantissa, exponent) -> factor ring
//   consumer   value ring -> e16m16 -> output window (diagonal write) -> global rows
// Global memory is touched ROW-wise (all lanes of one instruction on one lattice row: 256 contiguous
// bytes of lp2 / 128 of the output plane) although the recursion needs the cells DIAGONAL-wise (lane j
// is 1 row behind lane j-1): both directions go through shared-memory windows in which lane j only
// ever touches column j.  (One lattice row per lane and instruction, as the single-warp sweep does,
// costs 32 L1 tag look-ups per instruction.)
template <int DIR, bool kMulti, int KB>
struct WsGeom {  // what every helper derives from (warp, lane, utterance)
    int j, lag, base, tau0, stride;
    unsigned Tb_eff;
    size_t first;
    __device__ WsGeom(int Tb, int Ub, int T, int U1, int b, int wg, int lane, int xlag) {  // wg: warp in the whole sweep
        j = wg * 32 + lane;
        Tb_eff = j < Ub + 1 ? Tb : 0;  // (unsigned)tau < Tb_eff  <=>  this thread has a cell
        const int u = DIR == 0 ? j : Ub - j;
        lag = kMulti ? wg * KB + xlag : 0;
        base = lag + 32 * wg;          // lane 0 of this warp reaches row s - base at step s
        tau0 = -lag - j;               // this lane's progress at step 0
        stride = DIR == 0 ? U1 : -U1;
        first = (size_t)b * T * U1 + (size_t)(DIR == 0 ? 0 : Tb - 1) * U1 + u;
    }
};

template <int DIR, bool kMulti, int KB, int RW>
__device__ __forceinline__ void ws_loader(WsWarp<KB, RW>& W, const float2* __restrict__ lp2, int Tb, int Ub, int T,
                                          int U1, int b, int wg, int lane, int n_blocks, int xlag) {
    constexpr int kRun = RW / KB - 5;             // blocks the loader may run ahead of the chain (window reuse)
    constexpr int kFly = kRun > 5 ? 4 : kRun - 1; // cp.async groups in flight (~750 cycles each)
    static_assert(kRun >= 2 && kRun + 1 < kRowBars, "window too small for this block size");
    const WsGeom<DIR, kMulti, KB> G(Tb, Ub, T, U1, b, wg, lane, xlag);
    const uint32_t rows = tc::smem_u32(W.rows), progress = tc::smem_u32(&W.chain_done);
    float2* rawc = W.raw + lane;                  // column `lane` of the window, row stride 32
    const float2* p = lp2 + G.first + (long long)(-G.base) * G.stride;  // row -base: the first one block 0 needs
    int r = -G.base;
#pragma unroll 1
    for (int pb = 0; pb < n_blocks + kFly; ++pb) {
        if (pb < n_blocks) {
            // the chain (hence the converters) must have finished block pb - kRun before its rows are overwritten
            if (pb >= kRun) {
                int done;
                for (;;) {
                    asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(done) : "r"(progress) : "memory");
                    if (done > pb - kRun) break;
                    __nanosleep(64);
                }
            }
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                if ((unsigned)r < G.Tb_eff) cp_async_8(rawc + (r & (RW - 1)) * 32, p);
                ++r;
                p += G.stride;
            }
        }
        cp_async_commit();
        cp_async_wait<kFly>();  // the rows of block pb - kFly (and everything older) have landed
        const int done = pb - kFly;
        if (done >= 0) {
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(rows + 8 * (done % kRowBars));
        }
    }
}

// COMP 0: p(blank) -> fac[..].xy; COMP 1: p(label) -> fac[..].zw
template <int DIR, bool kMulti, int COMP, int KB, int RW>
__device__ __forceinline__ void ws_convert(WsWarp<KB, RW>& W, int Tb, int Ub, int T, int U1, int b, int wg, int lane,
                                           int n_blocks, int xlag) {
    const WsGeom<DIR, kMulti, KB> G(Tb, Ub, T, U1, b, wg, lane, xlag);
    const uint32_t bars = tc::smem_u32(W.bars), rows = tc::smem_u32(W.rows);
    const float* rawc = reinterpret_cast<const float*>(W.raw + lane) + COMP;
#pragma unroll 1
    for (int blk = 0; blk < n_blocks; ++blk) {
        const int st = blk % kWsStages;
        tc::mbar_wait(rows + 8 * (blk % kRowBars), (blk / kRowBars) & 1);                   // rows landed
        tc::mbar_wait(bars + 8 * (kWsStages + st), ((blk / kWsStages) & 1) ^ 1);            // ring slot free
#pragma unroll 1
        for (int h = 0; h < KB; h += kHalf) {
            const int tau = G.tau0 + blk * KB + h;
            float lp[kHalf];  // all loads, then all arithmetic, then all stores
#pragma unroll
            for (int k = 0; k < kHalf; ++k) lp[k] = rawc[((tau + k) & (RW - 1)) * 64];
            uint2 fv[kHalf];
#pragma unroll
            for (int k = 0; k < kHalf; ++k) {
                const ME f = me_from_log((unsigned)(tau + k) < G.Tb_eff ? lp[k] : 0.f);  // no cell: factor 1
                fv[k] = make_uint2(__float_as_uint(f.m), (unsigned)f.e);
            }
#pragma unroll
            for (int k = 0; k < kHalf; ++k) reinterpret_cast<uint2*>(&W.fac[st][h + k][lane])[COMP] = fv[k];
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bars + 8 * st);  // one of the two arrivals that complete "full"
    }
}

template <int DIR, bool kMulti, int KB, int RW>
__device__ __forceinline__ void ws_consumer(WsWarp<KB, RW>& W, const float2* __restrict__ lp2, int Tb, int Ub, int T,
                                            int U1, int b, int32_t* __restrict__ out, float* __restrict__ costs,
                                            float* __restrict__ ll_alpha, int w, int nw, int lane, int n_blocks,
                                            const WsBand& X, const int2 (*edge)[33]) {
    const WsGeom<DIR, kMulti, KB> G(Tb, Ub, T, U1, b, X.band * nw + w, lane, X.xlag);
    const uint32_t bars = tc::smem_u32(W.bars);
    // the last chain warp of a band forwards its boundary values to the next CTA of the cluster
    const bool to_band = kMulti && w == nw - 1 && X.band + 1 < X.n_bands;
    uint32_t r_xedge = 0, r_xdone = 0;
    if (to_band) {
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r_xedge) : "r"(tc::smem_u32(X.xedge)), "r"(X.band + 1));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r_xdone) : "r"(tc::smem_u32(X.xdone)), "r"(X.band + 1));
    }
    int32_t* outc = W.out + lane;              // column `lane` of the output window
    int tau_v = G.tau0;                        // progress at the step whose value is packed next
    int row_st = -G.base - 31;                 // next row to be stored (complete once lane 31 has passed it)
    int32_t* pst = out + G.first + (long long)row_st * G.stride;
    const int t_last = (G.Tb_eff != 0 && G.j == Ub) ? Tb - 1 : -1;
    ME last{1.f, kZeroExp};
#pragma unroll 1
    for (int vb = 0; vb < n_blocks; ++vb) {
        const int st = vb % kWsStages;
        tc::mbar_wait(bars + 8 * (2 * kWsStages + st), (vb / kWsStages) & 1);
#pragma unroll 1
        for (int h = 0; h < KB; h += kHalf) {
            int2 v[kHalf];
#pragma unroll
            for (int k = 0; k < kHalf; ++k) v[k] = W.val[st][h + k][lane];
            int pk[kHalf];
#pragma unroll
            for (int k = 0; k < kHalf; ++k) pk[k] = me_pack(ME{__int_as_float(v[k].x), v[k].y});
#pragma unroll
            for (int k = 0; k < kHalf; ++k) outc[((tau_v + k) & (kWin - 1)) * 32] = pk[k];
            if ((unsigned)(t_last - tau_v) < (unsigned)kHalf) {  // the terminal cell is here (one lane, once)
#pragma unroll
                for (int k = 0; k < kHalf; ++k)
                    if (tau_v + k == t_last) last = ME{__int_as_float(v[k].x), v[k].y};
            }
            tau_v += kHalf;
        }
        if (to_band && lane == 0) {  // slot q of the receiver = this warp's step index minus its lag
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                const int q = vb * KB + k - G.lag;
                if (q >= 0) {
                    const int2 ev = edge[q & (kWsEdgeRing - 1)][w];
                    asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(r_xedge + 8u * (unsigned)q), "r"(ev.x), "r"(ev.y) : "memory");
                }
            }
            asm volatile("st.release.cluster.shared::cluster.s32 [%0], %1;" ::"r"(r_xdone), "r"(vb + 1) : "memory");
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bars + 8 * (3 * kWsStages + st));
        // lane 31 has now passed rows row_st .. row_st + KB - 1
#pragma unroll 1
        for (int h = 0; h < KB; h += kHalf) {
            int ov[kHalf];
#pragma unroll
            for (int k = 0; k < kHalf; ++k) ov[k] = outc[((row_st + k) & (kWin - 1)) * 32];
#pragma unroll
            for (int k = 0; k < kHalf; ++k) {
                if ((unsigned)(row_st + k) < G.Tb_eff) *pst = ov[k];
                pst += G.stride;
            }
            row_st += kHalf;
        }
    }
    for (; row_st < Tb; ++row_st, pst += G.stride)  // rows the last lanes finished in the final blocks
        if ((unsigned)row_st < G.Tb_eff) *pst = outc[(row_st & (kWin - 1)) * 32];
    if (t_last >= 0) {
        if (DIR == 0) {
            if (ll_alpha) {
                const float2* src = lp2 + G.first;
                ll_alpha[b] = (float)me_ln(me_normalize(me_mul(last, me_from_log(src[(long long)t_last * G.stride].x))));
            }
        } else {
            costs[b] = cost_of(last);
        }
    }
}

template <bool kMulti, int KB, int RW, bool kCluster>
__global__ void __launch_bounds__(640, 1)
lattice_sweep_ws_kernel(const float2* __restrict__ lp2, const int32_t* __restrict__ act_lens,
                        const int32_t* __restrict__ label_lens, int T, int U1, int32_t* __restrict__ alpha,
                        int32_t* __restrict__ beta, float* __restrict__ costs, float* __restrict__ ll_alpha,
                        int xedge_slots, int nw, int spread) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WsWarp<KB, RW>* ws = reinterpret_cast<WsWarp<KB, RW>*>(smem_raw);
    __shared__ int2 edge[kWsEdgeRing][33];
    __shared__ int xdone_slot;
    // warps [0, nw): chains, then nw loaders, 2 nw converters, nw consumers -- or, `spread` (two chain
    // warps): chains on warps 0, 1 and the helpers on warps 2, 3 mod 4 only, so that (warp id mod 4
    // being the scheduler) no helper shares an issue port with a chain warp; warps 0, 1 mod 4 beyond the
    // chains stay idle
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WsBand X;
    X.band = kCluster ? (int)cluster_ctarank() : 0;
    X.n_bands = kCluster ? (int)cluster_nctarank() : 1;
    X.xlag = X.band * (kBandSkew - KB);
    X.xedge = reinterpret_cast<int2*>(ws + nw);  // [xedge_slots], cluster launches only
    X.xdone = &xdone_slot;
    const int b = kCluster ? blockIdx.x / X.n_bands : blockIdx.x;
    const int Tb = len_T(act_lens, b, T);
    const int Ub = len_U(label_lens, b, U1);
    if ((int)threadIdx.x < nw) {
        WsWarp<KB, RW>& W = ws[threadIdx.x];
        for (int i = 0; i < 4 * kWsStages; ++i) tc::mbar_init(tc::smem_u32(W.bars + i), i < kWsStages ? 2 : 1);
        for (int i = 0; i < kRowBars; ++i) tc::mbar_init(tc::smem_u32(W.rows + i), 1);
        W.chain_done = 0;
    }
    if (threadIdx.x == 0) xdone_slot = 0;
    tc::fence_barrier_init();
    if (kMulti)
        for (int i = threadIdx.x; i < kWsEdgeRing * 33; i += blockDim.x) edge[i / 33][i % 33] = make_int2(0x3f800000, kZeroExp);
    __syncthreads();
    if (kCluster) cluster_barrier();  // no CTA may be written to before it has initialised its shared memory
    // every warp of every band runs the same number of steps (uniform barriers), rounded up to whole blocks
    const int n_on = (Ub + 32) / 32;  // chain warps with cells
    const int max_lag = kMulti ? (n_on - 1) * KB + ((n_on - 1) / nw) * (kBandSkew - KB) : 0;
    int n_blocks = (Tb + Ub + max_lag + KB - 1) / KB;
    if (kCluster) n_blocks = min(n_blocks, xedge_slots / KB);  // (the host sized xedge for the longest utterance)
    int w = warp % nw, role = warp / nw;
    if (spread) {
        w = warp & 1;
        role = warp < 2 ? 0 : (warp & 2) ? 1 + (warp >> 2) : 5;  // 2,3 loaders; 6,7 / 10,11 converters; 14,15 consumers
    }
    const int wg = X.band * nw + w;
    const bool fwd = blockIdx.y == 0;
    int32_t* plane = fwd ? alpha : beta;
    if (role == 0) {
        if (fwd) ws_chain<0, kMulti, KB, RW>(ws[w], edge, w, nw, lane, n_blocks, X);
        else ws_chain<1, kMulti, KB, RW>(ws[w], edge, w, nw, lane, n_blocks, X);
    } else if (role == 1) {
        if (fwd) ws_loader<0, kMulti, KB, RW>(ws[w], lp2, Tb, Ub, T, U1, b, wg, lane, n_blocks, X.xlag);
        else ws_loader<1, kMulti, KB, RW>(ws[w], lp2, Tb, Ub, T, U1, b, wg, lane, n_blocks, X.xlag);
    } else if (role == 2) {
        if (fwd) ws_convert<0, kMulti, 0, KB, RW>(ws[w], Tb, Ub, T, U1, b, wg, lane, n_blocks, X.xlag);
        else ws_convert<1, kMulti, 0, KB, RW>(ws[w], Tb, Ub, T, U1, b, wg, lane, n_blocks, X.xlag);
    } else if (role == 3) {
        if (fwd) ws_convert<0, kMulti, 1, KB, RW>(ws[w], Tb, Ub, T, U1, b, wg, lane, n_blocks, X.xlag);
        else ws_convert<1, kMulti, 1, KB, RW>(ws[w], Tb, Ub, T, U1, b, wg, lane, n_blocks, X.xlag);
    } else if (role == 4) {
        if (fwd)
            ws_consumer<0, kMulti, KB, RW>(ws[w], lp2, Tb, Ub, T, U1, b, plane, costs, ll_alpha, w, nw, lane, n_blocks, X, edge);
        else
            ws_consumer<1, kMulti, KB, RW>(ws[w], lp2, Tb, Ub, T, U1, b, plane, costs, ll_alpha, w, nw, lane, n_blocks, X, edge);
    }
    if (kCluster) cluster_barrier();  // no CTA may exit while a neighbour can still write into its shared memory
}

template <bool kMulti, int KB, int RW>
int launch_ws(const float2* lp2, const int32_t* act_lens, const int32_t* label_lens, int B, int T, int U1,
              int32_t* alpha, int32_t* beta, float* costs, float* ll_alpha, int warps, cudaStream_t stream) {
    const size_t smem = (size_t)warps * sizeof(WsWarp<KB, RW>);
    cudaError_t e = cudaFuncSetAttribute(lattice_sweep_ws_kernel<kMulti, KB, RW, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return status_from_cuda(e);
    lattice_sweep_ws_kernel<kMulti, KB, RW, false><<<dim3(B, 2), warps * 160, smem, stream>>>(
        lp2, act_lens, label_lens, T, U1, alpha, beta, costs, ll_alpha, 0, warps, 0);
    return launch_status();
}

// More than four chain warps: bands of four in a thread-block cluster.  Returns -1 when the boundary
// buffer of the longest possible sweep does not fit in shared memory (the caller falls back).
int launch_ws_cluster(const float2* lp2, const int32_t* act_lens, const int32_t* label_lens, int B, int T, int U1,
                      int32_t* alpha, int32_t* beta, float* costs, float* ll_alpha, int warps, int NW,
                      cudaStream_t stream) {
    constexpr int KB = 8, RW = 128;
    const int n_bands = (warps + NW - 1) / NW;
    if (n_bands > 8) return -1;
    const int max_lag = (warps - 1) * KB + (n_bands - 1) * (kBandSkew - KB);
    const int slots = (T + U1 + max_lag + KB - 1) / KB * KB;
    const size_t smem = (size_t)NW * sizeof(WsWarp<KB, RW>) + (size_t)slots * sizeof(int2);
    if (NW < 1 || NW > 3) return -1;
    if (smem + sizeof(int2) * kWsEdgeRing * 33 + 64 > 227 * 1024) return -1;
    auto kern = lattice_sweep_ws_kernel<true, KB, RW, true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return status_from_cuda(e);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * n_bands), 2);
    const int spread = NW == 2;  // helpers on schedulers 2, 3 only (16 warps, six of them idle): 246 -> 236 us at cfg 3
    cfg.blockDim = dim3(spread ? 512 : NW * 160);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)n_bands;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int fits = 0;  // a cluster of n_bands CTAs with this much shared memory must be co-schedulable on one GPC
    e = cudaOccupancyMaxActiveClusters(&fits, kern, &cfg);
    if (e != cudaSuccess || fits < 1) {
        (void)cudaGetLastError();
        return -1;  // the caller falls back to the single-role sweep
    }
    e = cudaLaunchKernelEx(&cfg, kern, lp2, act_lens, label_lens, T, U1, alpha, beta, costs, ll_alpha, slots, NW, spread);
    return e == cudaSuccess ? launch_status() : status_from_cuda(e);
}


// ---- fused reduction of the costs (LossReduce, common.cuh) ---------------------------------------------
// Called by the one thread per utterance that has just written costs[b]: the last utterance to arrive sums
// all B costs in index order (the same bits whatever the arrival order) and re-arms the ticket.
struct ReduceArgs {
    float* out;
    int* ticket;
    float scale;
    int B;
};
__device__ __forceinline__ void reduce_costs_last_arriver(const ReduceArgs& R, const float* costs) {
    if (R.out == nullptr) return;
    __threadfence();  // costs[b] before the ticket
    if (atomicAdd(R.ticket, 1) != R.B - 1) return;
    __threadfence();
    float s = 0.f;
    for (int i = 0; i < R.B; ++i) s += __ldcg(costs + i);
    R.out[0] = s * R.scale;
    *R.ticket = 0;
}
__global__ void reduce_costs_kernel(const float* __restrict__ costs, int B, float scale, float* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {  // fallback sweeps only (same order, same bits)
        float s = 0.f;
        for (int i = 0; i < B; ++i) s += costs[i];
        out[0] = s * scale;
    }
}

// =================================================================================================
// Self-contained sweep ("tp"): one warp does everything for its 32 label positions.
//
// The warp-specialised kernel above buys chain latency with shared memory (179 KB per CTA at U1 = 81: one CTA
// per SM, so a large batch runs in waves) and with helper warps that compete with the chain for the SM.
// With the decoupled recursion the chain is short enough that ONE warp can carry all roles again, provided
// its global accesses stay row-wise.  They do, through two lane-private FIFOs in shared memory:
//   raw  lane l copies ITS column of every lattice row with cp.async (all lanes of one instruction on one
//        row: 256 contiguous bytes), kTpAhead + 1 blocks ahead, and reads it back l steps later -- lane l of
//        the wavefront is l rows behind lane 0.  A lane only ever reads its own copies: cp.async.wait_group
//        is all the synchronisation there is.
//   out  lane l writes its packed value and reads it back 31 - l steps later, when lane 31 has passed the
//        row and the whole row is stored with one 128-byte instruction.
// 24 KB per warp: three CTAs of three warps per SM at U1 = 81, a B = 512 batch is resident in 2.3 waves
// of 9 warps per SM instead of 7 waves of one CTA.  Warps of one sweep run KB steps apart and meet at a
// block barrier per KB steps (edge ring as in the chain warps above); longer label sequences are cut into
// bands of kTpBandWarps warps, one CTA of a thread-block cluster each, the boundary value written by lane
// 31 straight into the next CTA's shared memory (one slot per step of the sweep) followed by a release
// store of the finished block count.
constexpr int kTpKB = 8;        // steps per block = skew between consecutive warps
constexpr int kTpAhead = 2;     // blocks of lp2 rows in flight beyond the one being converted
constexpr int kTpRaw = 64;      // FIFO depth (rows): 31 (lane skew) + (kTpAhead + 2) * kTpKB <= 64
constexpr int kTpOut = 64;      // FIFO depth (rows): 31 + kTpKB + 1 <= 64
constexpr int kTpEdge = 32;     // >= 2 * kTpKB + 1 slots for warp-boundary values
constexpr int kTpBandSkew = 16; // steps the first warp of a band trails the last warp of the previous band
constexpr int kTpNoValue = (int)0x80000000;  // both words of a boundary slot nobody has written yet
static_assert(31 + (kTpAhead + 2) * kTpKB <= kTpRaw && 32 + kTpKB <= kTpOut, "FIFO too shallow");

struct TpWarp {
    float2 raw[kTpRaw][32];
    int32_t out[kTpOut][32];
};

template <int DIR, bool kMulti, bool kCluster>
__device__ __forceinline__ void tp_sweep(TpWarp& W, int2 (*edge)[8], const float2* __restrict__ lp2, int Tb, int Ub,
                                         int T, int U1, int b, int32_t* __restrict__ out, float* __restrict__ costs,
                                         float* __restrict__ ll_alpha, int w, int nw, int lane, const WsBand& X,
                                         const ReduceArgs& R) {
    constexpr int KB = kTpKB;
    const WsGeom<DIR, kMulti, KB> G(Tb, Ub, T, U1, b, X.band * nw + w, lane, X.xlag);
    const int n_on = (Ub + 32) / 32;  // warps with cells
    const int max_lag = kMulti ? (n_on - 1) * KB + ((n_on - 1) / nw) * (kTpBandSkew - KB) : 0;
    const int n_blocks = (Tb + Ub + max_lag + KB - 1) / KB;
    const bool from_band = kCluster && w == 0 && X.band > 0;
    const int edge_col = from_band ? 5 : (kMulti && w > 0) ? w - 1 : 4;  // 5: parked band-boundary values, 4: "zero"
    const bool to_band = kCluster && w == nw - 1 && X.band + 1 < X.n_bands;  // this warp's lane 31 feeds the next band
    uint32_t r_xedge = 0;
    if (to_band)
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r_xedge) : "r"(tc::smem_u32(X.xedge)), "r"(X.band + 1));
    const int in_bias = lane == 0 ? kNoTerm : 0;
    // FIFO slot of (row r, lane l) = (r + l) mod depth: lane l meets row r at step r + base + l, so the slot
    // index is the STEP index (minus base) -- uniform across the lanes and, base being a multiple of KB,
    // the KB slots of a block are consecutive: the chain side addresses both FIFOs with one register and
    // immediates.  (The row-wise sides -- cp.async in, row stores out -- pay the index arithmetic.)
    float2* rawc = &W.raw[0][lane];   // this lane's column of the FIFOs (slot stride 32 elements)
    int32_t* outc = &W.out[0][lane];

    // ---- loader: row r_ld (the same row for every lane of an instruction), this lane's cell of it.  Rows
    // outside [0, Tb) are CLAMPED, not skipped: a lane without a cell at some step then multiplies by a real
    // factor instead of 1, which is harmless -- before its first cell it carries "zero" (and zero times
    // anything finite stays zero), after its last one nothing reads it -- and it frees the conversions from
    // any predicate.  Lanes beyond the label sequence never load: their FIFO column is zeroed once
    // (log-probability 0 = factor 1).
    const bool lane_on = G.Tb_eff != 0;
    if (!lane_on)
        for (int i = 0; i < kTpRaw; ++i) rawc[i * 32] = make_float2(0.f, 0.f);
    int r_ld = -G.base;
    const float2* src0 = lp2 + G.first;
    auto load_block = [&]() {
        if (lane_on) {
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                const int rc = min(max(r_ld + k, 0), Tb - 1);
                cp_async_8(rawc + ((r_ld + k + lane) & (kTpRaw - 1)) * 32, src0 + rc * G.stride);
            }
        }
        r_ld += KB;
        cp_async_commit();
    };

    // prologue: rows of blocks 0 .. kTpAhead requested, block 0 converted
#pragma unroll
    for (int i = 0; i < kTpAhead + 1; ++i) load_block();
    cp_async_wait<kTpAhead>();
    int slot0 = (-G.base) & (kTpRaw - 1);  // slot of the first step of the current block (both FIFOs, same depth)
    static_assert(kTpRaw == kTpOut, "one slot register serves both FIFOs");
    float fm[KB][2];
    int fe[KB][2];
#pragma unroll
    for (int k = 0; k < KB; ++k) {
        const float2 lp = rawc[(slot0 + k) * 32];
        const ME fb = me_from_log(lp.x), fl = me_from_log(lp.y);
        fm[k][0] = fb.m, fe[k][0] = fb.e, fm[k][1] = fl.m, fe[k][1] = fl.e;
    }

    float m = 1.f;
    int E = G.j == 0 ? 0 : kZeroExp;
    float pbm_prev = 1.f, plm_prev = 1.f;
    int pbe_prev = 0, ple_prev = DIR == 0 ? kNoTerm : 0;
    const int edge_bias = lane == 0 ? 0 : kNoTerm;
    int tau = G.tau0;                 // progress at the first step of the current block
    int row_st = -G.base - 31;        // next row to be stored (complete once lane 31 has passed it)
    int32_t* pst = out + G.first + (long long)row_st * G.stride;
    const int t_last = (G.Tb_eff != 0 && G.j == Ub) ? Tb - 1 : -1;
    float last_m = 1.f;
    int last_E = kZeroExp;

#pragma unroll 1
    for (int blk = 0; blk < n_blocks; ++blk) {
        if (kMulti) __syncthreads();  // the previous warp has finished the block this one reads boundary values of
        load_block();                 // rows of block blk + kTpAhead + 1
        // boundary values of this block: every lane reads them (a broadcast), lanes other than 0 push the term
        // out of reach with an exponent bias.  Reader slots are block-aligned: the writer stores its local step
        // p at slot p + 1, the reader of local step q wants p = q - 1, i.e. slot q = er + k.
        float evm[KB];
        int evE[KB];
        if (kMulti) {
            const int er = (blk * KB - G.lag) & (kTpEdge - 1);
            if (kCluster && from_band) {
                // The value crossing a BAND boundary sits in this CTA's xedge array, one slot per step of the
                // sweep, and the slot IS the flag: it is written exactly once, by one 8-byte remote store of
                // the previous band's lane 31; until then both words hold kTpNoValue.  No release / acquire
                // pair, no progress counter: the sender never waits for anything.  Lanes 0..KB-1 poll the
                // block's KB slots in parallel and park them in this warp's private column of the edge ring,
                // from where the common path below picks them up.
                if (lane < KB) {
                    const int q = blk * KB + lane - G.lag - 1;
                    int2 ev = make_int2(0x3f800000, kZeroExp);
                    if (q >= 0) {
                        const volatile int2* slot = X.xedge + q;
                        long long t0 = 0;
                        for (unsigned spins = 0;; ++spins) {
                            ev.x = slot->x, ev.y = slot->y;
                            if (ev.x != kTpNoValue && ev.y != kTpNoValue) break;
                            if ((spins & 4095) == 4095) {  // a protocol bug must fail loudly, not hang the GPU
                                const long long now = clock64();
                                if (t0 == 0) t0 = now;
                                else if (now - t0 > 4000000000LL) __trap();
                            }
                        }
                    }
                    edge[er + lane][edge_col] = ev;
                }
                __syncwarp();
            }
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                const int2 ev = edge[er + k][edge_col];
                evm[k] = __int_as_float(ev.x);
                evE[k] = ev.y + edge_bias;
            }
        } else {
#pragma unroll
            for (int k = 0; k < KB; ++k) evm[k] = 1.f, evE[k] = kNoTerm;
        }
        cp_async_wait<kTpAhead>();  // rows of blocks <= blk + 1 have landed (this lane's own copies)
        const int slot1 = (slot0 + KB) & (kTpRaw - 1);
        float2 lpn[KB];
#pragma unroll
        for (int k = 0; k < KB; ++k) lpn[k] = rawc[(slot1 + k) * 32];

        // exponent recurrence of step 0 of this block (it cannot run ahead across the renormalisation)
        int En;
        float c_own, c_in, c_edge;
        {
            const int pbe = DIR == 0 ? pbe_prev : fe[0][0], ple = DIR == 0 ? ple_prev : fe[0][1];
            const int oE = E + pbe;
            const int iE = __shfl_up_sync(0xffffffffu, DIR == 0 ? E + ple : E, 1) +
                           (DIR == 0 ? in_bias : (blk == 0 ? kNoTerm : in_bias) + ple);
            const int eE = DIR == 0 ? evE[0] : evE[0] + ple;
            En = max(max(oE, iE), eE);
            c_own = pow2_neg(En - oE) * (DIR == 0 ? pbm_prev : fm[0][0]);
            c_in = pow2_neg(En - iE) * (DIR == 0 ? 1.f : fm[0][1]);
            c_edge = pow2_neg(En - eE) * (DIR == 0 ? 1.f : fm[0][1]);
        }
        float shm = DIR == 0 ? m * plm_prev : m;
        const int ew = (blk * KB - G.lag) & (kTpEdge - 1);  // slot base of this warp's own boundary values
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            const float in_m = __shfl_up_sync(0xffffffffu, shm, 1);  // mantissa chain: the long-latency hop first
            const float own_term = kMulti ? fmaf(evm[k], c_edge, m * c_own) : m * c_own;
            const int Ek = En;
            const float ci = c_in;
            const float plm = fm[k][1];
            const int ple = fe[k][1];
            if (k + 1 < KB) {  // exponent recurrence of step k+1, in the shadow of the shuffle above
                const int pbe_n = fe[DIR == 0 ? k : k + 1][0], ple_n = fe[DIR == 0 ? k : k + 1][1];
                const int oE = Ek + pbe_n;
                const int iE = __shfl_up_sync(0xffffffffu, DIR == 0 ? Ek + ple_n : Ek, 1) +
                               (DIR == 0 ? in_bias : in_bias + ple_n);
                const int eE = DIR == 0 ? evE[k + 1] : evE[k + 1] + ple_n;
                En = max(max(oE, iE), eE);
                c_own = pow2_neg(En - oE) * fm[DIR == 0 ? k : k + 1][0];
                c_in = pow2_neg(En - iE) * (DIR == 0 ? 1.f : fm[k + 1][1]);
                c_edge = pow2_neg(En - eE) * (DIR == 0 ? 1.f : fm[k + 1][1]);
            } else {
                pbm_prev = fm[k][0], plm_prev = plm, pbe_prev = fe[k][0], ple_prev = ple;
            }
            m = fmaf(in_m, ci, own_term);
            E = Ek;
            if (k + 1 == KB) {  // renormalise once per block
                const int bits = __float_as_int(m);
                m = __int_as_float((bits & 0x007fffff) | 0x3f800000);
                E += (bits >> 23) - 127;
            }
            shm = DIR == 0 ? m * plm : m;
            const int shE = DIR == 0 ? E + ple : E;
            if (kMulti) {
                if (lane == 31) edge[k + 1 < KB ? ew + k + 1 : (ew + KB) & (kTpEdge - 1)][w] = make_int2(__float_as_int(shm), shE);
            }
            // off the chain: pack and park the value; the factor registers of this step are free now and
            // take the next block's factors of the same slot
            outc[(slot0 + k) * 32] = me_pack(ME{m, E});
            if (tau + k == t_last) last_m = m, last_E = E;
            {
                const ME fb = me_from_log(lpn[k].x), fl = me_from_log(lpn[k].y);
                fm[k][0] = fb.m, fe[k][0] = fb.e, fm[k][1] = fl.m, fe[k][1] = fl.e;
            }
        }
        if (to_band) {  // forward this block's KB boundary values (they sit in the edge ring) to the next band:
            __syncwarp();  // lanes 0..KB-1 each copy one slot into the receiving CTA's xedge[q], q = step - lag
            if (lane < KB) {
                const int q = blk * KB + lane - G.lag;
                if (q >= 0) {
                    const int2 ev = edge[lane + 1 < KB ? ew + lane + 1 : (ew + KB) & (kTpEdge - 1)][w];
                    asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(r_xedge + 8u * (unsigned)q), "r"(ev.x), "r"(ev.y) : "memory");
                }
            }
        }
        // lane 31 has now passed rows row_st .. row_st + KB - 1
        {
            int ov[KB];
#pragma unroll
            for (int k = 0; k < KB; ++k) ov[k] = outc[((row_st + k + lane) & (kTpOut - 1)) * 32];
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                if ((unsigned)(row_st + k) < G.Tb_eff) *pst = ov[k];
                pst += G.stride;
            }
            row_st += KB;
        }
        slot0 = slot1;
        tau += KB;
    }
    cp_async_wait<0>();
    for (; row_st < Tb; ++row_st, pst += G.stride)  // rows the last lanes finished in the final blocks
        if ((unsigned)row_st < G.Tb_eff) *pst = outc[((row_st + lane) & (kTpOut - 1)) * 32];
    if (t_last >= 0) {
        const ME last{last_m, last_E};
        if (DIR == 0) {
            if (ll_alpha) {
                const float2* src = lp2 + G.first;
                ll_alpha[b] = (float)me_ln(me_normalize(me_mul(last, me_from_log(src[(long long)t_last * G.stride].x))));
            }
        } else {
            costs[b] = cost_of(me_normalize(last));
            reduce_costs_last_arriver(R, costs);
        }
    }
}

template <bool kMulti, bool kCluster>
__global__ void __launch_bounds__(128)
lattice_sweep_tp_kernel(const float2* __restrict__ lp2, const int32_t* __restrict__ act_lens,
                        const int32_t* __restrict__ label_lens, int T, int U1, int32_t* __restrict__ alpha,
                        int32_t* __restrict__ beta, float* __restrict__ costs, float* __restrict__ ll_alpha,
                        int xedge_slots, ReduceArgs R) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_launch_dependents();
    TpWarp* tw = reinterpret_cast<TpWarp*>(smem_raw);
    __shared__ int2 edge[kTpEdge][8];  // column w: written by warp w; column 4 stays "zero" (warp 0 of band 0)
    __shared__ int xdone_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    WsBand X;
    X.band = kCluster ? (int)cluster_ctarank() : 0;
    X.n_bands = kCluster ? (int)cluster_nctarank() : 1;
    X.xlag = X.band * (kTpBandSkew - kTpKB);
    X.xedge = reinterpret_cast<int2*>(tw + nw);  // [xedge_slots], cluster launches only
    X.xdone = &xdone_slot;
    const int b = kCluster ? blockIdx.x / X.n_bands : blockIdx.x;
    const int Tb = len_T(act_lens, b, T);
    const int Ub = len_U(label_lens, b, U1);
    if (kMulti) {
        for (int i = threadIdx.x; i < kTpEdge * 8; i += blockDim.x) edge[i / 8][i % 8] = make_int2(0x3f800000, kZeroExp);
        if (threadIdx.x == 0) xdone_slot = 0;
        __syncthreads();
    }
    if (kCluster) {
        for (int i = threadIdx.x; i < xedge_slots; i += blockDim.x) X.xedge[i] = make_int2(kTpNoValue, kTpNoValue);
        cluster_barrier();  // no CTA may be written to before it has initialised its shared memory
    }
    pdl_wait();  // lp2 comes from the front-end kernel; nothing in global memory is touched before this point
    if (blockIdx.y == 0)
        tp_sweep<0, kMulti, kCluster>(tw[warp], edge, lp2, Tb, Ub, T, U1, b, alpha, costs, ll_alpha, warp, nw, lane, X, R);
    else
        tp_sweep<1, kMulti, kCluster>(tw[warp], edge, lp2, Tb, Ub, T, U1, b, beta, costs, ll_alpha, warp, nw, lane, X, R);
    if (kCluster) cluster_barrier();  // no CTA may exit while a neighbour can still write into its shared memory
}

// warps <= 4: one CTA per sweep.  More: bands of `bw` warps in a thread-block cluster.  Returns -1 when the
// configuration cannot be launched (the caller falls back to the other kernels).
int launch_tp(const float2* lp2, const int32_t* act_lens, const int32_t* label_lens, int B, int T, int U1,
              int32_t* alpha, int32_t* beta, float* costs, float* ll_alpha, int warps, int bw, cudaStream_t stream,
              ReduceArgs R) {
    if (warps <= 4) {
        const size_t smem = (size_t)warps * sizeof(TpWarp);
        if (warps == 1) {
            auto kern = lattice_sweep_tp_kernel<false, false>;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
            if (launch_pdl(pdl_ok((long long)B * T), kern, dim3(B, 2), dim3(32), smem, stream, lp2, act_lens, label_lens, T, U1, alpha,
                           beta, costs, ll_alpha, 0, R) != cudaSuccess)
                return status_from_cuda(cudaGetLastError());
        } else {
            auto kern = lattice_sweep_tp_kernel<true, false>;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
            if (launch_pdl(pdl_ok((long long)B * T), kern, dim3(B, 2), dim3(warps * 32), smem, stream, lp2, act_lens, label_lens, T, U1,
                           alpha, beta, costs, ll_alpha, 0, R) != cudaSuccess)
                return status_from_cuda(cudaGetLastError());
        }
        return launch_status();
    }
    if (bw < 1 || bw > 4) return -1;
    const int n_bands = (warps + bw - 1) / bw;
    if (n_bands > 8) return -1;
    const int max_lag = (warps - 1) * kTpKB + (n_bands - 1) * (kTpBandSkew - kTpKB);
    const int slots = (T + U1 + max_lag + kTpKB - 1) / kTpKB * kTpKB;
    const size_t smem = (size_t)bw * sizeof(TpWarp) + (size_t)slots * sizeof(int2);
    if (smem + 4096 > 227 * 1024) return -1;
    auto kern = lattice_sweep_tp_kernel<true, true>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * n_bands), 2);
    cfg.blockDim = dim3((unsigned)(bw * 32));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)n_bands;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int fits = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&fits, kern, &cfg);
    if (e != cudaSuccess || fits < 1) {
        (void)cudaGetLastError();
        return -1;
    }
    e = cudaLaunchKernelEx(&cfg, kern, lp2, act_lens, label_lens, T, U1, alpha, beta, costs, ll_alpha, slots, R);
    return e == cudaSuccess ? launch_status() : status_from_cuda(e);
}

}  // namespace

namespace {
int sweep_dispatch(const float2* lp2, const int32_t* act_lens, const int32_t* label_lens, int B, int T, int U1,
                   int32_t* alpha, int32_t* beta, float* costs, float* ll_alpha, cudaStream_t stream,
                   const ReduceArgs& R, bool* reduced);
}

int launch_lattice_sweep(const float2* lp2, const int32_t* act_lens, const int32_t* label_lens, int B,
                         int T, int U1, int32_t* alpha, int32_t* beta, float* costs, float* ll_alpha,
                         cudaStream_t stream, const LossReduce* reduce) {
    ReduceArgs R{nullptr, nullptr, 0.f, B};
    if (reduce && reduce->out) {
        if (!reduce->ticket) return RNNTB200_STATUS_INVALID_VALUE;
        R = ReduceArgs{reduce->out, reduce->ticket, reduce->scale, B};
    }
    bool reduced = false;
    const int st = sweep_dispatch(lp2, act_lens, label_lens, B, T, U1, alpha, beta, costs, ll_alpha, stream, R, &reduced);
    if (st != RNNTB200_STATUS_SUCCESS || R.out == nullptr || reduced || B == 0) return st;
    reduce_costs_kernel<<<1, 32, 0, stream>>>(costs, B, R.scale, R.out);  // the fallback sweeps do not carry the ticket
    return launch_status();
}

namespace {
int sweep_dispatch(const float2* lp2, const int32_t* act_lens, const int32_t* label_lens, int B, int T, int U1,
                   int32_t* alpha, int32_t* beta, float* costs, float* ll_alpha, cudaStream_t stream,
                   const ReduceArgs& R, bool* reduced) {
    if (B == 0) return RNNTB200_STATUS_SUCCESS;
    if (U1 > 1024) return RNNTB200_STATUS_INVALID_VALUE;
    const int warps = (U1 + 31) / 32;
    static const bool legacy = getenv("RNNTB200_SWEEP_LEGACY") != nullptr;  // A/B timing of the one-warp-does-all sweep
    // A/B switch (timing experiments only): RNNTB200_SWEEP=tp | ws selects the self-contained or the
    // warp-specialised kernel for every shape it can run; RNNTB200_SWEEP_BW = warps per band (tp, long sequences)
    static const char* which = getenv("RNNTB200_SWEEP");
    static const int band_warps_env = getenv("RNNTB200_SWEEP_BW") ? atoi(getenv("RNNTB200_SWEEP_BW")) : 0;
    // bands of two warps (five SMs for U1 = 301) up to 16 warps; wider bands beyond, so that 8 bands cover U1 = 1024
    const int band_warps = band_warps_env ? band_warps_env : (warps <= 16 ? 2 : (warps <= 24 ? 3 : 4));
    const bool force_ws = which && which[0] == 'w';
    // Measured (round 2, sweep alone, us; tp / ws with the decoupled chain / round 1's ws):
    //   U1 = 81  (3 warps) B = 32: 45 / 51 / 54;   B = 512: 188 / 334 / 350
    //   U1 = 101 (4 warps) B = 16: 49 / 67 / 68
    //   U1 = 301 (10 warps, tp as 5 bands of 2 in a cluster) B = 8: 221 / 237 / 236;   B = 296: 1801 / 5300 / 5300
    // -> the self-contained sweep serves every shape it can launch; the others remain as fall-backs.
    if (!legacy && !force_ws) {
        const int st = launch_tp(lp2, act_lens, label_lens, B, T, U1, alpha, beta, costs, ll_alpha, warps, band_warps, stream, R);
        if (st >= 0) {
            *reduced = true;
            return st;
        }
    }
    // Measured (sweep alone, us):  U1 = 81 (3 warps): 53 warp-specialised in one CTA / 59 single-role;
    // U1 = 101 (4 warps): 81 in one CTA (20 warps crowd the SM) / 74 as two bands / 68 single-role;
    // U1 = 301 (10 warps): 245 as five bands of two / 274 as four bands of three / 360 single-role cluster.
    if (!legacy) {
        if (warps == 1) return launch_ws<false, 8, 128>(lp2, act_lens, label_lens, B, T, U1, alpha, beta, costs, ll_alpha, 1, stream);
        if (warps <= 3) return launch_ws<true, 8, 128>(lp2, act_lens, label_lens, B, T, U1, alpha, beta, costs, ll_alpha, warps, stream);
        if (warps >= 5) {  // bands of two (up to 16 warps) or three chain warps in a thread-block cluster
            const int st = launch_ws_cluster(lp2, act_lens, label_lens, B, T, U1, alpha, beta, costs, ll_alpha, warps,
                                             warps <= 16 ? 2 : 3, stream);
            if (st >= 0) return st;
        }
    }
    if (warps == 1) {
        const size_t smem = (size_t)kRingStride * 32 * sizeof(float2);
        lattice_sweep_kernel<0><<<dim3(B, 2), 32, smem, stream>>>(lp2, act_lens, label_lens, T, U1, alpha, beta,
                                                                  costs, ll_alpha);
        return launch_status();
    }
    if (warps <= 4) {  // one CTA: its warps sit on different schedulers of one SM
        const size_t smem = (size_t)kRingStride * warps * 32 * sizeof(float2);
        lattice_sweep_kernel<1><<<dim3(B, 2), warps * 32, smem, stream>>>(lp2, act_lens, label_lens, T, U1, alpha,
                                                                          beta, costs, ll_alpha);
        return launch_status();
    }
    // longer label sequences: spread the sweep over a cluster, <= 4 warps per CTA, <= 8 CTAs
    const int wpc = (warps + 7) / 8, cs = (warps + wpc - 1) / wpc;
    const size_t smem = (size_t)kRingStride * wpc * 32 * sizeof(float2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * cs), 2);
    cfg.blockDim = dim3((unsigned)(wpc * 32));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, lattice_sweep_kernel<2>, lp2, act_lens, label_lens, T, U1, alpha, beta,
                                       costs, ll_alpha);
    return e == cudaSuccess ? launch_status() : status_from_cuda(e);
}
}  // namespace

}  // namespace rnntb200
