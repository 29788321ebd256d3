// proj_tc_bwd.cu -- backward of the two projections of the reference-exact joint (proj_tc.cu) on
// the tensor cores at fp32 accuracy (bf16 hi/lo split, three tcgen05 MMAs per product):
//     d_x   = (dP W_s) .* gelu_tanh'(x)            [rows, K]     K' = V
//     d_W_s^T += gelu_tanh(x)^T dP                 [K, V]        K' = rows   (TMEM-resident per tile)
//     d_b   += dP^T 1                              [V]           (encoder projection only)
// for (x, W_s) = (enc, W[:, :He]) and (dec, W[:, He:]).  A tile is 128 consecutive rows; one CTA per
// tile.  dP (hi/lo) stays in shared memory for the whole tile and is read K-major by the d_x
// product and MN-major by the d_W product; gelu(x) streams through a 2-stage ring of 128-column
// blocks; W K-blocks arrive by TMA and are read MN-major (K' = V).
#include <algorithm>

#include "tc_common.cuh"

namespace rnntb200 {

using namespace tc;

namespace {

constexpr int kKB = 64;
constexpr int kThreads = 18 * 32;  // 8 producer warps, TMA, MMA, 2 x 4 epilogue warps (one group per dX buffer)
constexpr int kGroup = 2048;              // 128 rows x 16 B
constexpr int kGxHalf = 16 * kGroup;      // 128 columns of gelu(x), hi (lo follows): 32 KiB
constexpr int kRBytes = 16 * 256;         // ones selector, K-major [16 rows][128]
constexpr int kStageBytes = 8 * 2048;     // epilogue transposition buffers: [32 rows][16 floats] per warp
static_assert(kStageBytes >= kRBytes, "the staging area starts on top of the ones selector");
constexpr int kColDX = 320, kColDB = 448;
constexpr int kTmemCols = 512;

struct ProblemB {
    const void* x;    // [rows, K], element type xdt (rnntb200_dtype_t)
    const float* dp;  // [rows, V]
    void* dx;         // [rows, K], same element type
    int rows, K, tiles, w_col0, with_bias, xdt;
};

struct SmemPB {
    int dp, gx, w, r, bars, total, dp_half, w_half;
};
__host__ __device__ inline SmemPB smem_layout_pb(int NB) {
    SmemPB s;
    s.dp = 0;
    s.dp_half = (NB / 8) * kGroup;
    s.gx = 2 * s.dp_half;
    s.w = s.gx + 2 * 2 * kGxHalf;
    s.w_half = NB * kKB * 2;
    s.bars = s.w + 2 * 2 * s.w_half;
    s.r = s.bars + 24 * 8 + 16;  // the ones selector; its 4 KiB are also the head of the epilogue staging
    s.total = s.r + kStageBytes; // area (the d_b products that read it are complete long before)
    return s;
}

__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
    const __nv_bfloat162 h = __halves2bfloat162(ah, bh);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - __bfloat162float(ah), b - __bfloat162float(bh));
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// sigmoid(2y), y = sqrt(2/pi)(x + 0.044715 x^3):  gelu = x s,  gelu' = s + x s (1 - s) d(2y)/dx
__device__ __forceinline__ float gelu_sig(float x) {
    const float a = x * fmaf(-0.10294324f, x * x, -2.3022082f);  // -2y log2(e)
    return fast_rcp(1.f + fast_ex2(a));
}
__device__ __forceinline__ float gelu_val(float x) { return x * gelu_sig(x); }
__device__ __forceinline__ float gelu_grad(float x) {
    const float s = gelu_sig(x);
    const float dy2 = 1.5957691216057308f * fmaf(0.134145f * x, x, 1.f);
    return fmaf(x * s * (1.f - s), dy2, s);
}

__global__ void __launch_bounds__(kThreads, 1)
proj_tc_bwd_kernel(const __grid_constant__ CUtensorMap w0_hi, const __grid_constant__ CUtensorMap w0_lo,
                   const __grid_constant__ CUtensorMap w1_hi, const __grid_constant__ CUtensorMap w1_lo,
                   ProblemB p0, ProblemB p1, int V, int NB, float* __restrict__ d_weight, int ldw,
                   float* __restrict__ d_bias) {
    extern __shared__ __align__(128) unsigned char smem[];
    pdl_launch_dependents();
    const SmemPB L = smem_layout_pb(NB);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool second = (int)blockIdx.x >= p0.tiles;
    const ProblemB P = second ? p1 : p0;
    const int row0 = (second ? blockIdx.x - p0.tiles : blockIdx.x) * 128;
    const int n_blk = P.K / 128, n_kb = P.K / kKB;

    const uint32_t sbase = smem_u32(smem);
    const uint32_t dp_hi = sbase + L.dp, dp_lo = dp_hi + L.dp_half, gx_base = sbase + L.gx,
                   w_base = sbase + L.w, r_base = sbase + L.r, bars = sbase + L.bars;
    const uint32_t dp_full = bars;
    auto gx_full = [&](int i) { return bars + 8 * (1 + i); };
    auto gx_empty = [&](int i) { return bars + 8 * (3 + i); };
    auto w_full = [&](int i) { return bars + 8 * (5 + i); };
    auto w_empty = [&](int i) { return bars + 8 * (7 + i); };
    auto dx_full = [&](int i) { return bars + 8 * (9 + i); };
    auto dx_empty = [&](int i) { return bars + 8 * (11 + i); };
    const uint32_t done = bars + 8 * 13;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bars + 24 * 8);

    if (threadIdx.x == 0) {
        mbar_init(dp_full, 8);
        for (int i = 0; i < 2; ++i) {
            mbar_init(gx_full(i), 8);
            mbar_init(gx_empty(i), 1);
            mbar_init(w_full(i), 1);
            mbar_init(w_empty(i), 1);
            mbar_init(dx_full(i), 1);
            mbar_init(dx_empty(i), 4);
        }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    // ones selector (K-major B operand, 16 rows x 128): row 0 = 1 -> column 0 of the product = sum over rows
    for (int i = threadIdx.x; i < 16 * 128; i += kThreads) {
        const int n = i >> 7, k = i & 127;
        *reinterpret_cast<__nv_bfloat16*>(smem + L.r + (k >> 3) * 256 + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) =
            __float2bfloat16_rn(n == 0 ? 1.f : 0.f);
    }
    fence_async_smem();
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();  // d_penc / d_pdec of the gradient kernel are read from here on

    if (warp < 8) {
        // ===== producers =====
        // A warp covers 16 rows; lane = (row within 8, one of 4 adjacent 16-byte-output groups): every
        // global load instruction touches 8 rows x 128 contiguous bytes (8 cache lines -- one row per
        // lane would touch 32 and is bound by L1 tag look-ups), and a quarter warp stores 8 rows x 16 B
        // = one 128-byte core-matrix column, conflict-free.
        const int r8 = warp * 16 + (lane & 7), c4 = lane >> 3;
        // (1) dP tile -> (hi, lo), [v-group][row][8 v].  All loads of a row go out before the first
        // conversion (the tile is the first thing every MMA of this CTA waits for).
#pragma unroll
        for (int i1 = 0; i1 < 2; ++i1) {
            const int r = r8 + 8 * i1, row = row0 + r;
            const bool row_ok = row < P.rows;
            const float* dprow = P.dp + (size_t)min(row, P.rows - 1) * V;
            float v[3][8];  // v-groups c4, c4 + 4, c4 + 8 (NB <= 80: at most ten groups)
#pragma unroll
            for (int i2 = 0; i2 < 3; ++i2)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int col = (c4 + 4 * i2) * 8 + j;
                    v[i2][j] = (row_ok && col < V) ? __ldg(dprow + col) : 0.f;
                }
#pragma unroll
            for (int i2 = 0; i2 < 3; ++i2) {
                const int gi = c4 + 4 * i2;
                if (gi >= NB / 8) continue;
                uint4 hi, lo;
                split2(v[i2][0], v[i2][1], hi.x, lo.x);
                split2(v[i2][2], v[i2][3], hi.y, lo.y);
                split2(v[i2][4], v[i2][5], hi.z, lo.z);
                split2(v[i2][6], v[i2][7], hi.w, lo.w);
                *reinterpret_cast<uint4*>(smem + L.dp + gi * kGroup + r * 16) = hi;
                *reinterpret_cast<uint4*>(smem + L.dp + L.dp_half + gi * kGroup + r * 16) = lo;
            }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(dp_full);
        // (2) gelu(x) in 128-column blocks -> (hi, lo), [h-group][row][8 h]
        const bool row_ok0 = row0 + r8 < P.rows, row_ok1 = row0 + r8 + 8 < P.rows;
        const size_t xrow0 = (size_t)min(row0 + r8, P.rows - 1) * P.K;  // element offsets of this thread's two rows
        const size_t xrow1 = (size_t)min(row0 + r8 + 8, P.rows - 1) * P.K;
        // each 128-column block is produced as two 64-column halves from two register buffers that
        // alternate: the loads of the next half are in flight while the current one is computed.
        // Item i of a half: row r8 + 8 (i & 1), h-group 8 hf + c4 + 4 (i >> 1).
        float4 xa[8], xb[8];
        auto load_half = [&](float4(&dst)[8], int blk, int hf) {
            if (blk < n_blk) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const size_t xr = (i & 1) ? xrow1 : xrow0;
                    const int kc = 8 * hf + c4 + 4 * (i >> 1);
                    dst[2 * i] = ldx4(P.x, P.xdt, xr + blk * 128 + kc * 8);
                    dst[2 * i + 1] = ldx4(P.x, P.xdt, xr + blk * 128 + kc * 8 + 4);
                }
            }
        };
        auto do_half = [&](const float4(&x)[8], int st, int hf) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int kc = 8 * hf + c4 + 4 * (i >> 1), r = r8 + 8 * (i & 1);
                const float4 x0 = x[2 * i], x1 = x[2 * i + 1];
                uint4 hi, lo;
                split2(gelu_val(x0.x), gelu_val(x0.y), hi.x, lo.x);
                split2(gelu_val(x0.z), gelu_val(x0.w), hi.y, lo.y);
                split2(gelu_val(x1.x), gelu_val(x1.y), hi.z, lo.z);
                split2(gelu_val(x1.z), gelu_val(x1.w), hi.w, lo.w);
                if (!((i & 1) ? row_ok1 : row_ok0)) hi = lo = make_uint4(0, 0, 0, 0);  // rows past the end must not reach d_W
                unsigned char* dst = smem + L.gx + st * 2 * kGxHalf + kc * kGroup + r * 16;
                *reinterpret_cast<uint4*>(dst) = hi;
                *reinterpret_cast<uint4*>(dst + kGxHalf) = lo;
            }
        };
        load_half(xa, 0, 0);
        for (int blk = 0; blk < n_blk; ++blk) {
            const int st = blk & 1;
            load_half(xb, blk, 1);
            mbar_wait(gx_empty(st), ((blk >> 1) & 1) ^ 1);
            do_half(xa, st, 0);
            load_half(xa, blk + 1, 0);
            do_half(xb, st, 1);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(gx_full(st));
        }
    } else if (warp == 8) {
        if (lane == 0) {
            const CUtensorMap* mh = second ? &w1_hi : &w0_hi;
            const CUtensorMap* ml = second ? &w1_lo : &w0_lo;
            for (int kb = 0; kb < n_kb; ++kb) {
                const int st = kb & 1;
                mbar_wait(w_empty(st), ((kb >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(w_full(st), 2u * (uint32_t)L.w_half);
                tma_load_3d(w_base + st * 2 * L.w_half, mh, 0, 0, kb * (kKB / 8), w_full(st));
                tma_load_3d(w_base + st * 2 * L.w_half + L.w_half, ml, 0, 0, kb * (kKB / 8), w_full(st));
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            const uint32_t id_db = umma_idesc_bf16(16, true, false);   // dP^T (MN) x ones (K-major)
            const uint32_t id_dw = umma_idesc_bf16(NB, true, true);    // gelu(x)^T (MN) x dP (MN)
            const uint32_t id_dx = umma_idesc_bf16(kKB, false, true);  // dP (K-major) x W (MN)
            const uint32_t w_sbo = NB * 16;
            mbar_wait(dp_full, 0);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const uint64_t ones = umma_desc(r_base + ks * 2 * 256, 256, 128);
                umma_bf16(tmem + kColDB, umma_desc(dp_lo + ks * 256, 128, kGroup), ones, id_db, ks != 0);
                umma_bf16(tmem + kColDB, umma_desc(dp_hi + ks * 256, 128, kGroup), ones, id_db, 1);
            }
            for (int blk = 0; blk < n_blk; ++blk) {
                const int st = blk & 1;
                mbar_wait(gx_full(st), (blk >> 1) & 1);
                tc_fence_after();
                const uint32_t gh = gx_base + st * 2 * kGxHalf, gl = gh + kGxHalf;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint64_t ah = umma_desc(gh + ks * 256, 128, kGroup), al = umma_desc(gl + ks * 256, 128, kGroup);
                    const uint64_t bh = umma_desc(dp_hi + ks * 256, 128, kGroup), bl = umma_desc(dp_lo + ks * 256, 128, kGroup);
                    umma_bf16(tmem + blk * NB, al, bh, id_dw, ks != 0);
                    umma_bf16(tmem + blk * NB, ah, bl, id_dw, 1);
                    umma_bf16(tmem + blk * NB, ah, bh, id_dw, 1);
                }
                umma_commit(gx_empty(st));
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int kb = blk * 2 + h2, ws = kb & 1, buf = kb & 1;
                    mbar_wait(w_full(ws), (kb >> 1) & 1);
                    mbar_wait(dx_empty(buf), ((kb >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t wh = w_base + ws * 2 * L.w_half, wl = wh + L.w_half;
                    for (int j = 0; j < NB / 16; ++j) {
                        const uint64_t ah = umma_desc(dp_hi + j * 2 * kGroup, kGroup, 128), al = umma_desc(dp_lo + j * 2 * kGroup, kGroup, 128);
                        const uint64_t bh = umma_desc(wh + j * 256, 128, w_sbo), bl = umma_desc(wl + j * 256, 128, w_sbo);
                        umma_bf16(tmem + kColDX + buf * kKB, al, bh, id_dx, j != 0);
                        umma_bf16(tmem + kColDX + buf * kKB, ah, bl, id_dx, 1);
                        umma_bf16(tmem + kColDX + buf * kKB, ah, bh, id_dx, 1);
                    }
                    umma_commit(w_empty(ws));
                    umma_commit(dx_full(buf));
                }
            }
            umma_commit(done);
        }
    } else {
        // ===== epilogue: one row per thread (= TMEM lane); group g (warps 10-13 / 14-17) owns dX buffer
        // g, i.e. the even / odd 64-column pieces, and half of the final d_W flush.
        // Global memory is touched through a per-warp transposition buffer ([32 rows][16 floats], XOR
        // swizzled): 8 rows x 64 contiguous bytes per instruction instead of 32 rows x 16 bytes (a
        // thread-per-row access costs 32 L1 tag look-ups per instruction and starves the producers).
        const int grp = (warp - 10) >> 2;
        const int q = warp & 3, r = q * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        float* stg = reinterpret_cast<float*>(smem + L.r) + (warp - 10) * 512;
        auto sw = [](int rowi, int chunk) { return rowi * 16 + ((chunk ^ ((rowi >> 1) & 3)) << 2); };
        const int lr = lane >> 2, lc = lane & 3;  // transposed ownership: rows lr + 8 j, 16-byte chunk lc
        for (int kb = grp; kb < n_kb; kb += 2) {
            const int buf = grp;
            float4 xg[4][4];  // x of this 64-column piece, transposed ownership, requested before the wait
#pragma unroll
            for (int pc = 0; pc < 4; ++pc)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int grow = min(row0 + q * 32 + lr + 8 * j, P.rows - 1);
                    xg[pc][j] = ldx4(P.x, P.xdt, (size_t)grow * P.K + kb * kKB + pc * 16 + lc * 4);
                }
            mbar_wait(dx_full(buf), (kb >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) {
#pragma unroll
                for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(stg + sw(lr + 8 * j, lc)) = xg[pc][j];
                __syncwarp();
                float4 xr[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) xr[c] = *reinterpret_cast<const float4*>(stg + sw(lane, c));
                float v[16];
                tmem_ld16(tmem + kColDX + buf * kKB + pc * 16 + lane_sel, v);
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<float4*>(stg + sw(lane, c)) =
                        make_float4(v[4 * c] * gelu_grad(xr[c].x), v[4 * c + 1] * gelu_grad(xr[c].y),
                                    v[4 * c + 2] * gelu_grad(xr[c].z), v[4 * c + 3] * gelu_grad(xr[c].w));
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int grow = row0 + q * 32 + lr + 8 * j;
                    if (grow < P.rows)
                        stx4(P.dx, P.xdt, (size_t)grow * P.K + kb * kKB + pc * 16 + lc * 4,
                             *reinterpret_cast<const float4*>(stg + sw(lr + 8 * j, lc)));
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dx_empty(buf));
        }
        mbar_wait(done, 0);
        tc_fence_after();
        for (int blk = grp; blk < n_blk; blk += 2)
            for (int pc = 0; pc < NB / 16; ++pc) {
                float v[16];
                tmem_ld16(tmem + blk * NB + pc * 16 + lane_sel, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int col = pc * 16 + i;
                    if (col < V) atomicAdd(d_weight + (size_t)col * ldw + P.w_col0 + blk * 128 + r, v[i]);
                }
            }
        if (P.with_bias && grp == 0) {
            float v[16];
            tmem_ld16(tmem + kColDB + lane_sel, v);
            if (r < V) atomicAdd(d_bias + r, v[0]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
    }
}

}  // namespace

// proj_tc.cu
size_t proj_tc_workspace_bytes(int V, int He, int Hd);
int proj_tc_prepare(const float* weight, int V, int He, int Hd, int NB, void* workspace, size_t workspace_bytes,
                    bool already_split, CUtensorMap maps[4], cudaStream_t stream);

bool proj_tc_bwd_supported(int V, int He, int Hd) {
    return V >= 1 && V <= 80 && He >= 128 && Hd >= 128 && He % 128 == 0 && Hd % 128 == 0 && He <= 512 && Hd <= 512;
}

size_t proj_tc_bwd_workspace_bytes(int V, int He, int Hd) { return proj_tc_workspace_bytes(V, He, Hd); }

int launch_proj_tc_bwd(const void* enc, const void* dec, int x_dtype, const float* weight, const float* d_penc,
                       const float* d_pdec, int rows_enc, int rows_dec, int He, int Hd, int V, void* d_enc,
                       void* d_dec, float* d_weight, float* d_bias, void* workspace, size_t workspace_bytes,
                       int workspace_holds_split, cudaStream_t stream) {
    if (!proj_tc_bwd_supported(V, He, Hd)) return RNNTB200_STATUS_INVALID_VALUE;
    const int ldw = He + Hd;
    if (d_bias == d_weight + (size_t)V * ldw) {  // one flat gradient buffer (loss.py allocates it so): one memset node
        if (cudaMemsetAsync(d_weight, 0, ((size_t)V * ldw + V) * sizeof(float), stream) != cudaSuccess)
            return RNNTB200_STATUS_MEMOPS_FAILED;
    } else if (cudaMemsetAsync(d_weight, 0, (size_t)V * ldw * sizeof(float), stream) != cudaSuccess ||
               cudaMemsetAsync(d_bias, 0, (size_t)V * sizeof(float), stream) != cudaSuccess) {
        return RNNTB200_STATUS_MEMOPS_FAILED;
    }
    const int NB = ((V + 15) / 16) * 16;
    CUtensorMap m[4];
    int st = proj_tc_prepare(weight, V, He, Hd, NB, workspace, workspace_bytes, workspace_holds_split != 0, m, stream);
    if (st != RNNTB200_STATUS_SUCCESS) return st;
    if (rows_enc + rows_dec == 0) return RNNTB200_STATUS_SUCCESS;
    ProblemB p0{enc, d_penc, d_enc, rows_enc, He, (rows_enc + 127) / 128, 0, 1, x_dtype};
    ProblemB p1{dec, d_pdec, d_dec, rows_dec, Hd, (rows_dec + 127) / 128, He, 0, x_dtype};
    const SmemPB L = smem_layout_pb(NB);
    cudaError_t e = cudaFuncSetAttribute(proj_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    if (e != cudaSuccess) return status_from_cuda(e);
    e = launch_pdl(pdl_ok(rows_enc), proj_tc_bwd_kernel, dim3(p0.tiles + p1.tiles), dim3(kThreads), (size_t)L.total, stream, m[0],
                   m[1], m[2], m[3], p0, p1, V, NB, d_weight, ldw, d_bias);
    return e == cudaSuccess ? launch_status() : status_from_cuda(e);
}

}  // namespace rnntb200
