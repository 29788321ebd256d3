// joint_at_tc_bwd.cu -- tcgen05 gradient of the ADD_TANH joint + RNN-T loss (row K4 of SURVEY 8):
// the softmax is recomputed from the saved log-sum-exp and d(logits) goes straight into the GEMM
// backward, all on the tensor cores; neither logits, d(logits) nor tanh activations touch HBM.
//
// Per tile of 128 lattice cells (16 t x 8 u), with z = tanh(e_t + d_u) [128 x H] in shared memory:
//   S   = z W^T                      [128 x V]   recompute of the logits            (K = H)
//   G   = grad_cost * (softmax(S) * occupancy - blank/label corrections), bf16 -> shared memory
//   dW^T += z^T G                    [H x V]     accumulated in TMEM across ALL tiles (K = cells)
//   db  += G^T 1                     [V]         same, through a 0/1 selector matrix
//   dZ  = G W                        [128 x H]   in 64-column pieces               (K = V)
//   dP  = dZ * (1 - z^2), bf16, written over z in shared memory
//   d_enc^T, d_dec^T = dP^T R        [H x 16], [H x 8]   R = 0/1 selectors of the cell's t / u
// Every product is a tcgen05.mma (kind::f16, bf16 operands, fp32 accumulators in TMEM).  ONE copy
// of z, ONE copy of G and the TMA-streamed W tile serve all GEMMs: the same bytes are read as a
// K-major operand by one product and as an MN-major operand by another (core matrices are 8 x 16 B
// either way; only the descriptor strides and the major bits of the instruction descriptor change).
//
// Warp roles (one persistent CTA per SM): warps 0-15 build z; warp 16 streams W by TMA (twice per
// tile: for S and for dZ); warp 17 issues every MMA; warps 18-25 (two groups) own one lattice cell (TMEM lane)
// each: they turn S into G, dZ into dP, and move the reduced d_enc / d_dec tiles to HBM (fp32
// atomics).  mbarrier-only synchronisation.  TMEM map (512 columns): dW^T [0,320) | S [320,400) |
// dZ pieces [400,464) and [320,384) (the second aliases S, dead by then) | d_enc^T/d_dec^T [320,448) (aliases S and dZ, both dead by then) | db [464,480).
//
// Supported: H a multiple of 128, H <= 512; any V (one launch per chunk of 80 vocabulary columns;
// KsponSpeech's 73 is a single pass).  Other shapes use the CUDA-core kernel in joint_at.cu.
#include <algorithm>

#include "tc_common.cuh"

namespace rnntb200 {

using namespace tc;

namespace {

constexpr int kTT = 16, kUU = 8;
constexpr int kKB = 64;
constexpr int kProducerWarps = 16;
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kTmaWarp = 16, kMmaWarp = 17;  // warps 18-21 / 22-25: epilogue groups 0 / 1
constexpr int kThreads = 26 * 32;
constexpr int kMaxWStages = 8;
constexpr int kASlotBytes = 128 * kKB * 2;  // one 64-wide K block of z: 16 KiB
constexpr int kGroupBytes = 2048;           // 128 rows x 16 B: one 8-element column group of z / G
constexpr int kRBytes = 16 * 512;           // selector matrix R: [32 rows][128 cells] bf16, K-major
constexpr int kColDW = 0, kColS = 320, kColDZ = 400, kColRed = 320, kColDB = 464;
constexpr int kTmemCols = 512;

struct SmemB {
    int z, g, w, r, dd, bars, total;
    int w_stage_bytes, dd_stride, w_stages;
};

__host__ __device__ inline SmemB smem_layout_b(int H, int NB) {
    SmemB s;
    s.z = 0;
    s.g = (H / kKB) * kASlotBytes;
    s.w = s.g + (NB / 8) * kGroupBytes;  // W stages directly after G (see the db product)
    s.w_stage_bytes = NB * kKB * 2;
    s.dd_stride = H + 4;  // 8 predictor rows 4 banks apart: conflict-free float4 reads
    const int fixed = s.w + kRBytes + kUU * s.dd_stride * 4 + 40 * 8 + 32;
    s.w_stages = (227 * 1024 - fixed) / s.w_stage_bytes;
    s.w_stages = s.w_stages > kMaxWStages ? kMaxWStages : s.w_stages;
    s.r = s.w + s.w_stages * s.w_stage_bytes;
    s.dd = s.r + kRBytes;
    s.bars = (s.dd + kUU * s.dd_stride * 4 + 15) & ~15;
    s.total = s.bars + 40 * 8 + 16;
    return s;
}

__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}

__global__ void __launch_bounds__(kThreads, 1)
at_grad_tc_kernel(const __grid_constant__ CUtensorMap w_map, const float* __restrict__ enc,
                  const float* __restrict__ dec, const float* __restrict__ bias,
                  const int32_t* __restrict__ labels, const int32_t* __restrict__ act_lens,
                  const int32_t* __restrict__ label_lens, int B, int T, int U1, int V, int H, int NB, int v0,
                  int blank, const float2* __restrict__ lp2, const float* __restrict__ lse,
                  const int32_t* __restrict__ alpha, const int32_t* __restrict__ beta,
                  const float* __restrict__ grad_costs, float* __restrict__ d_enc,
                  float* __restrict__ d_dec, float* __restrict__ d_w, float* __restrict__ d_b) {
    extern __shared__ __align__(128) unsigned char smem[];
    const SmemB L = smem_layout_b(H, NB);
    const int n_slots = H / kKB, n_mt = H / 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t sbase = smem_u32(smem);
    const uint32_t z_base = sbase + L.z, g_base = sbase + L.g, w_base = sbase + L.w, r_base = sbase + L.r;
    float* dd = reinterpret_cast<float*>(smem + L.dd);  // [8 predictor rows][dd_stride]
    const int kWStages = L.w_stages;
    const uint32_t bars = sbase + L.bars;
    auto z_full = [&](int i) { return bars + 8 * i; };
    const uint32_t z_empty = bars + 8 * 8;
    auto w_full = [&](int i) { return bars + 8 * (9 + i); };
    auto w_empty = [&](int i) { return bars + 8 * (17 + i); };
    const uint32_t s_full = bars + 8 * 25, g_full = bars + 8 * 26, p_full = bars + 8 * 29,
                   r_full = bars + 8 * 30, r_empty = bars + 8 * 31, done = bars + 8 * 32;
    auto dz_full = [&](int i) { return bars + 8 * (33 + i); };   // one dZ piece buffer per epilogue group
    auto dz_empty = [&](int i) { return bars + 8 * (35 + i); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bars + 40 * 8);

    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(z_full(i), kProducerWarps);
        mbar_init(z_empty, 1);
        for (int i = 0; i < kMaxWStages; ++i) { mbar_init(w_full(i), 1); mbar_init(w_empty(i), 1); }
        mbar_init(s_full, 1);
        mbar_init(g_full, 8);
        for (int i = 0; i < 2; ++i) { mbar_init(dz_full(i), 1); mbar_init(dz_empty(i), 4); }
        mbar_init(p_full, 8);
        mbar_init(r_full, 1);
        mbar_init(r_empty, 8);
        mbar_init(done, 1);
        fence_barrier_init();
    }
    // selector matrix R (K-major B operand, 32 rows x 128 cells): row n < 16 selects the cells of
    // frame n of the tile, 16 <= n < 24 the cells of label position n-16, row 24 is all ones
    for (int i = threadIdx.x; i < 32 * 128; i += kThreads) {
        const int n = i >> 7, k = i & 127;
        const float v = n < 16 ? (k / kUU == n) : n < 24 ? (k % kUU == n - 16) : n == 24 ? 1.f : 0.f;
        *reinterpret_cast<__nv_bfloat16*>(smem + L.r + (k >> 3) * 512 + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) =
            __float2bfloat16_rn(v);
    }
    fence_async_smem();
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int nT = (T + kTT - 1) / kTT, nU = (U1 + kUU - 1) / kUU;
    const int n_tiles = B * nT * nU;
    auto decode = [&](int tile, int& b, int& t0, int& u0) -> bool {
        b = tile / (nT * nU);
        const int r = tile - b * nT * nU;
        t0 = (r / nU) * kTT;
        u0 = (r % nU) * kUU;
        return t0 < len_T(act_lens, b, T) && u0 <= len_U(label_lens, b, U1);
    };

    if (warp < kProducerWarps) {
        // ===== producers: z = tanh(e_t + d_u) -> bf16, core-matrix layout, one K block per slot =====
        const int p = threadIdx.x;
        const int r = p & 127, kc0 = p >> 7;
        const int tt = r / kUU, uu = r % kUU;
        uint32_t n = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            int b, t0, u0;
            if (!decode(tile, b, t0, u0)) continue;
            mbar_wait(z_empty, (n & 1) ^ 1);  // every reader of the previous tile's z / dP is done
            // stage the tile's 8 predictor rows (each is reused by all 16 frames); the encoder rows
            // are read straight from global memory: every element is needed by exactly one warp
            asm volatile("bar.sync 1, 512;" ::: "memory");  // previous tile's readers are done
            const int H4 = H / 4;
            for (int i = p; i < kUU * H4; i += kProducerThreads) {
                const int row = i / H4, c4 = i - row * H4;
                const float* src = dec + ((size_t)b * U1 + min(u0 + row, U1 - 1)) * H;
                *reinterpret_cast<float4*>(dd + row * L.dd_stride + 4 * c4) =
                    __ldg(reinterpret_cast<const float4*>(src) + c4);
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            const float4* erow = reinterpret_cast<const float4*>(enc + ((size_t)b * T + min(t0 + tt, T - 1)) * H);
            const float* drow = dd + uu * L.dd_stride;
            // three register buffers take turns (K loop unrolled by three, roles are compile-time):
            // block kb+2 is requested while block kb is computed
            float4 eb0[4], eb1[4], eb2[4];
            auto load_e = [&](float4(&dst)[4], int kb) {
                if (kb < n_slots) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        dst[2 * i] = __ldg(erow + kb * (kKB / 4) + (kc0 + 4 * i) * 2);
                        dst[2 * i + 1] = __ldg(erow + kb * (kKB / 4) + (kc0 + 4 * i) * 2 + 1);
                    }
                }
            };
            auto block = [&](const float4(&e)[4], int kb) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int kc = kc0 + 4 * i;
                    const int k = kb * kKB + kc * 8;
                    const float4 e0 = e[2 * i], e1 = e[2 * i + 1];
                    const float4 d0 = *reinterpret_cast<const float4*>(drow + k);
                    const float4 d1 = *reinterpret_cast<const float4*>(drow + k + 4);
                    uint4 out;
                    out.x = pack_bf16(tanh_fast(e0.x + d0.x), tanh_fast(e0.y + d0.y));
                    out.y = pack_bf16(tanh_fast(e0.z + d0.z), tanh_fast(e0.w + d0.w));
                    out.z = pack_bf16(tanh_fast(e1.x + d1.x), tanh_fast(e1.y + d1.y));
                    out.w = pack_bf16(tanh_fast(e1.z + d1.z), tanh_fast(e1.w + d1.w));
                    *reinterpret_cast<uint4*>(smem + L.z + kb * kASlotBytes + kc * kGroupBytes + r * 16) = out;
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(z_full(kb));
            };
            load_e(eb0, 0);
            load_e(eb1, 1);
            for (int kb = 0; kb < n_slots; kb += 3) {
                load_e(eb2, kb + 2);
                block(eb0, kb);
                if (kb + 1 < n_slots) {
                    load_e(eb0, kb + 3);
                    block(eb1, kb + 1);
                }
                if (kb + 2 < n_slots) {
                    load_e(eb1, kb + 4);
                    block(eb2, kb + 2);
                }
            }
            ++n;
        }
    } else if (warp == kTmaWarp) {
        // ===== TMA producer: W K-blocks, once for S and once for dZ per tile =====
        if (lane == 0) {
            uint32_t wi = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                int b, t0, u0;
                if (!decode(tile, b, t0, u0)) continue;
                for (int pass = 0; pass < 2; ++pass)
                    for (int kb = 0; kb < n_slots; ++kb, ++wi) {
                        const int st = wi % kWStages;
                        mbar_wait(w_empty(st), ((wi / kWStages) & 1) ^ 1);
                        mbar_arrive_expect_tx(w_full(st), (uint32_t)L.w_stage_bytes);
                        tma_load_3d(w_base + st * L.w_stage_bytes, &w_map, 0, v0, kb * (kKB / 8), w_full(st));
                    }
            }
        }
    } else if (warp == kMmaWarp) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t id_s = umma_idesc_bf16(NB, false, false);   // z (K-major) x W (K-major)
            const uint32_t id_dw = umma_idesc_bf16(NB, true, true);    // z^T (MN) x G (MN)
            const uint32_t id_db = umma_idesc_bf16(16, true, false);   // G^T (MN) x R (K-major)
            const uint32_t id_dz = umma_idesc_bf16(kKB, false, true);  // G (K-major) x W (MN)
            const uint32_t id_red = umma_idesc_bf16(32, true, false);  // dP^T (MN) x R (K-major)
            const uint32_t w_sbo = NB * 16;
            uint32_t n = 0, wi = 0, pi = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                int b, t0, u0;
                if (!decode(tile, b, t0, u0)) continue;
                const uint32_t ph = n & 1;
                // (1) S = z W^T
                mbar_wait(r_empty, ph ^ 1);  // previous tile's reduced tiles (aliasing S / dZ) are read out
                tc_fence_after();
                for (int kb = 0; kb < n_slots; ++kb, ++wi) {
                    const int st = wi % kWStages;
                    mbar_wait(z_full(kb), ph);
                    mbar_wait(w_full(st), (wi / kWStages) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int j = 0; j < kKB / 16; ++j)
                        umma_bf16(tmem + kColS,
                                  umma_desc(z_base + kb * kASlotBytes + j * 2 * kGroupBytes, kGroupBytes, 128),
                                  umma_desc(w_base + st * L.w_stage_bytes + j * 2 * w_sbo, w_sbo, 128), id_s,
                                  (kb | j) != 0);
                    umma_commit(w_empty(st));
                }
                umma_commit(s_full);
                // (2) dW^T += z^T G, db += G^T 1   (K = the tile's 128 cells)
                mbar_wait(g_full, ph);
                tc_fence_after();
                for (int mt = 0; mt < n_mt; ++mt)
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_bf16(tmem + kColDW + mt * NB,
                                  umma_desc(z_base + mt * 16 * kGroupBytes + ks * 256, 128, kGroupBytes),
                                  umma_desc(g_base + ks * 256, 128, kGroupBytes), id_dw, (n | ks) != 0);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_bf16(tmem + kColDB, umma_desc(g_base + ks * 256, 128, kGroupBytes),
                              umma_desc(r_base + 2 * 128 + ks * 2 * 512, 512, 128), id_db, (n | ks) != 0);
                // (3) dZ = G W, one 64-column piece per W K-block
                for (int kb = 0; kb < n_slots; ++kb, ++wi, ++pi) {
                    const int st = wi % kWStages;
                    const int buf = kb & 1;  // n_slots is even: buffer b sees pieces b, b+2, ...
                    mbar_wait(dz_empty(buf), ((pi >> 1) & 1) ^ 1);
                    mbar_wait(w_full(st), (wi / kWStages) & 1);
                    tc_fence_after();
                    for (int j = 0; j < NB / 16; ++j)
                        umma_bf16(tmem + (buf ? kColS : kColDZ),
                                  umma_desc(g_base + j * 2 * kGroupBytes, kGroupBytes, 128),
                                  umma_desc(w_base + st * L.w_stage_bytes + j * 256, 128, w_sbo), id_dz, j != 0);
                    umma_commit(w_empty(st));
                    umma_commit(dz_full(buf));
                }
                // (4) d_enc^T | d_dec^T = dP^T R
                mbar_wait(p_full, ph);
                tc_fence_after();
                for (int mt = 0; mt < n_mt; ++mt)
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_bf16(tmem + kColRed + mt * 32,
                                  umma_desc(z_base + mt * 16 * kGroupBytes + ks * 256, 128, kGroupBytes),
                                  umma_desc(r_base + ks * 2 * 512, 512, 128), id_red, ks != 0);
                umma_commit(r_full);
                umma_commit(z_empty);
                ++n;
            }
            umma_commit(done);
        }
    } else {
        // ===== epilogue: one lattice cell (TMEM lane) per thread, two groups of four warps.  Both
        // groups see all 128 cells; they split the work by columns: group g turns the vocabulary
        // pieces pc = g, g+2, .. of S into G, owns dZ buffer g (pieces kb = g, g+2, ..), and moves the
        // reduced tiles / accumulators of the row blocks mt = g, g+2, .. =====
        const int grp = (warp - 18) >> 2;
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int tt = r / kUU, uu = r % kUU;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        uint32_t n = 0, pi = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            int b, t0, u0;
            if (!decode(tile, b, t0, u0)) continue;
            const uint32_t ph = n & 1;
            const int Tb = len_T(act_lens, b, T), Ub = len_U(label_lens, b, U1);
            const int t = t0 + tt, u = u0 + uu;
            // per-cell scalars (same closed form as every other gradient kernel of the library)
            float c_all = -INFINITY, corr_b = 0.f, corr_l = 0.f;
            int y = -1;
            const float gscale = grad_costs[b];
            if (t < Tb && u <= Ub) {
                const size_t c = ((size_t)b * T + t) * U1 + u;
                const int aq = alpha[c], llq = beta[(size_t)b * T * U1];
                const float2 lp = lp2[c];
                c_all = e16m16_log2_ratio(aq, beta[c], llq) - lse[c] * kLog2e;
                if (t < Tb - 1) corr_b = fast_ex2(e16m16_log2_ratio(aq, beta[c + U1], llq) + lp.x * kLog2e);
                else if (u == Ub) corr_b = fast_ex2(e16m16_log2_ratio(aq, 0, llq) + lp.x * kLog2e);
                if (u < Ub) {
                    y = label_at(labels, b, U1, u, V);
                    corr_l = fast_ex2(e16m16_log2_ratio(aq, beta[c + 1], llq) + lp.y * kLog2e);
                }
            }
            // (a) S -> G (bf16, [v-group][cell][8 v])
            mbar_wait(s_full, ph);
            tc_fence_after();
            for (int pc = grp; pc < NB / 16; pc += 2) {
                float v[16];
                tmem_ld16(tmem + kColS + pc * 16 + lane_sel, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int col = v0 + pc * 16 + i;  // vocabulary column of this launch's chunk
                    float g = 0.f;
                    if (col < V) {
                        g = fast_ex2(fmaf(v[i], kLog2e, __ldg(bias + col) * kLog2e) + c_all);
                        if (col == blank) g -= corr_b;
                        if (col == y) g -= corr_l;
                        g *= gscale;
                    }
                    v[i] = g;
                }
                uint4 lo, hi;
                lo.x = pack_bf16(v[0], v[1]);   lo.y = pack_bf16(v[2], v[3]);
                lo.z = pack_bf16(v[4], v[5]);   lo.w = pack_bf16(v[6], v[7]);
                hi.x = pack_bf16(v[8], v[9]);   hi.y = pack_bf16(v[10], v[11]);
                hi.z = pack_bf16(v[12], v[13]); hi.w = pack_bf16(v[14], v[15]);
                *reinterpret_cast<uint4*>(smem + L.g + (2 * pc) * kGroupBytes + r * 16) = lo;
                *reinterpret_cast<uint4*>(smem + L.g + (2 * pc + 1) * kGroupBytes + r * 16) = hi;
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(g_full);
            // (c) dZ pieces -> dP = dZ * (1 - z^2), bf16, over z
            for (int kb = grp; kb < n_slots; kb += 2, ++pi) {
                mbar_wait(dz_full(grp), pi & 1);
                tc_fence_after();
#pragma unroll
                for (int pc = 0; pc < 4; ++pc) {
                    float v[16];
                    tmem_ld16(tmem + (grp ? kColS : kColDZ) + pc * 16 + lane_sel, v);
#pragma unroll
                    for (int hgrp = 0; hgrp < 2; ++hgrp) {
                        uint4* zp = reinterpret_cast<uint4*>(smem + L.z + kb * kASlotBytes +
                                                             (2 * pc + hgrp) * kGroupBytes + r * 16);
                        const uint4 zz = *zp;
                        const uint32_t zw[4] = {zz.x, zz.y, zz.z, zz.w};
                        uint32_t ow[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float2 z2 = unpack_bf16(zw[k]);
                            ow[k] = pack_bf16(v[hgrp * 8 + 2 * k] * (1.f - z2.x * z2.x),
                                              v[hgrp * 8 + 2 * k + 1] * (1.f - z2.y * z2.y));
                        }
                        *zp = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(dz_empty(grp));
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
            // (d) reduced tiles: TMEM lane = h within the 128-row block, columns = frame / position
            mbar_wait(r_full, ph);
            tc_fence_after();
            for (int mt = grp; mt < n_mt; mt += 2) {
                float ve[16], vd[16];
                tmem_ld16(tmem + kColRed + mt * 32 + lane_sel, ve);
                tmem_ld16(tmem + kColRed + mt * 32 + 16 + lane_sel, vd);
                const int h = mt * 128 + r;
#pragma unroll
                for (int k = 0; k < kTT; ++k)
                    if (t0 + k < Tb) atomicAdd(d_enc + ((size_t)b * T + t0 + k) * H + h, ve[k]);
#pragma unroll
                for (int k = 0; k < kUU; ++k)
                    if (u0 + k <= Ub) atomicAdd(d_dec + ((size_t)b * U1 + u0 + k) * H + h, vd[k]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(r_empty);
            ++n;
        }
        // flush the accumulators that lived in TMEM for the whole kernel
        if (n > 0) {
            mbar_wait(done, 0);
            tc_fence_after();
            for (int mt = grp; mt < n_mt; mt += 2)
                for (int pc = 0; pc < NB / 16; ++pc) {
                    float v[16];
                    tmem_ld16(tmem + kColDW + mt * NB + pc * 16 + lane_sel, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int col = v0 + pc * 16 + i;
                        if (col < V) atomicAdd(d_w + (size_t)col * H + mt * 128 + r, v[i]);
                    }
                }
            if (grp == 0) {
                float v[16];
                tmem_ld16(tmem + kColDB + lane_sel, v);
                if (r < NB && v0 + r < V) atomicAdd(d_b + v0 + r, v[8]);  // column 8 = selector row 24 = all cells
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
    }
}

}  // namespace

// defined in joint_at_tc.cu
int at_tc_prepare_weight(const float* weight, int V, int H, int NB, void* workspace, size_t workspace_bytes,
                         CUtensorMap* map, cudaStream_t stream);

bool at_tc_bwd_supported(int V, int H) { return V >= 1 && H >= 128 && H % 128 == 0 && H <= 512; }

int launch_at_grad_tc(const float* enc, const float* dec, const float* weight, const float* bias,
                      const int32_t* labels, const int32_t* act_lens, const int32_t* label_lens, int B, int T,
                      int U1, int V, int H, int blank, const float2* lp2, const float* lse, const int32_t* alpha,
                      const int32_t* beta, const float* grad_costs, float* d_enc, float* d_dec, float* d_weight,
                      float* d_bias, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!at_tc_bwd_supported(V, H)) return RNNTB200_STATUS_INVALID_VALUE;
    // Vocabularies beyond 80 columns run one pass per chunk of 80: the softmax is normalised by the
    // saved log-sum-exp, so every chunk's G, dW rows and its share of dZ are independent; d_enc and
    // d_dec simply accumulate over the passes (z is recomputed per pass).
    const int NB = std::min(((V + 15) / 16) * 16, 80);
    CUtensorMap map;
    int st = at_tc_prepare_weight(weight, V, H, NB, workspace, workspace_bytes, &map, stream);
    if (st != RNNTB200_STATUS_SUCCESS) return st;
    const SmemB L = smem_layout_b(H, NB);
    cudaError_t e = cudaFuncSetAttribute(at_grad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    if (e != cudaSuccess) return status_from_cuda(e);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_tiles = B * ((T + kTT - 1) / kTT) * ((U1 + kUU - 1) / kUU);
    for (int v0 = 0; v0 < V; v0 += NB) {
        at_grad_tc_kernel<<<std::min(n_tiles, sms), kThreads, L.total, stream>>>(
            map, enc, dec, bias, labels, act_lens, label_lens, B, T, U1, V, H, NB, v0, blank, lp2, lse, alpha, beta,
            grad_costs, d_enc, d_dec, d_weight, d_bias);
        if ((st = launch_status()) != RNNTB200_STATUS_SUCCESS) return st;
    }
    return RNNTB200_STATUS_SUCCESS;
}

}  // namespace rnntb200
