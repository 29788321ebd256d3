// proj_tc.cu -- the two small projections of the reference-exact joint on the tensor cores:
//     P_enc = gelu_tanh(enc) W[:, :He]^T + bias   [B*T,  V]
//     P_dec = gelu_tanh(dec) W[:, He:]^T          [B*U1, V]
// (networks/transducer.py:64-69 in factorised form), at fp32 accuracy: every
// operand is split into two bf16 numbers (x = hi + lo, |lo| <= 2^-9 |x|) and three tcgen05 MMAs
// accumulate hi*hi + hi*lo + lo*hi in fp32 TMEM; the dropped lo*lo term is 2^-18 relative.
// One launch covers both projections: a tile is 128 consecutive rows of enc or of dec.
//
//   warps 0-15  A producers: x -> gelu_tanh(x) (= x * sigmoid(2y), one EX2 + one RCP) -> (hi, lo)
//               bf16 pair, written in UMMA K-major core-matrix layout, one 64-wide K block per stage
//   warp  16    TMA: the matching K blocks of W_hi and W_lo (bf16 copies made by a prologue kernel)
//   warp  17    MMA issuer: 3 x 4 tcgen05.mma per K block
//   warps 18-21 epilogue: TMEM -> registers -> + bias -> global (one output row per thread)
#include <algorithm>

#include "tc_common.cuh"

namespace rnntb200 {

using namespace tc;

namespace {

constexpr int kKB = 64;
constexpr int kProducerWarps = 16;  // gelu + hi/lo split is the long pole: 4 threads per row
constexpr int kTmaWarp = 16, kMmaWarp = 17;
constexpr int kThreads = 22 * 32;
constexpr int kStages = 4;
constexpr int kAHalf = 128 * kKB * 2;  // 16 KiB: one K block of A_hi (A_lo follows)
constexpr int kTmemCols = 128;

struct Problem {  // one projection
    const void* x;      // [rows, K], element type xdt (rnntb200_dtype_t)
    const float* bias;  // [V] or null
    float* out;         // [rows, V]
    int rows, K, tiles, xdt;
};

struct SmemP {
    int a, w, bars, total, w_half;
};
__host__ __device__ inline SmemP smem_layout_p(int NB) {
    SmemP s;
    s.a = 0;
    s.w = kStages * 2 * kAHalf;
    s.w_half = NB * kKB * 2;
    s.bars = s.w + kStages * 2 * s.w_half;
    s.total = s.bars + 24 * 8 + 16;
    return s;
}

__device__ __forceinline__ float gelu_tanh(float x) {
    // 0.5 x (1 + tanh(y)) = x * sigmoid(2y),  y = sqrt(2/pi) (x + 0.044715 x^3)
    // -2y log2(e) = x (c0 + c1 x^2); ex2 overflows to +inf for very negative x and the reciprocal is then 0
    const float a = x * fmaf(-0.10294324f, x * x, -2.3022082f);
    return x * fast_rcp(1.f + fast_ex2(a));
}

__device__ __forceinline__ void split_store(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
    const __nv_bfloat162 h = __halves2bfloat162(ah, bh);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - __bfloat162float(ah), b - __bfloat162float(bh));
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// fc.weight [V, He+Hd] -> dense bf16 hi / lo copies of both column slices ([V, He] and [V, Hd])
__global__ void split_weight_kernel(const float* __restrict__ w, int V, int He, int Hd,
                                    __nv_bfloat16* __restrict__ e_hi, __nv_bfloat16* __restrict__ e_lo,
                                    __nv_bfloat16* __restrict__ d_hi, __nv_bfloat16* __restrict__ d_lo) {
    const int ldw = He + Hd;
    pdl_launch_dependents();
    pdl_wait();  // the previous step's projection backward may still be reading the split this kernel overwrites
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < V * ldw; i += gridDim.x * blockDim.x) {
        const int v = i / ldw, k = i - v * ldw;
        const float x = w[i];
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
        if (k < He) {
            e_hi[v * He + k] = h;
            e_lo[v * He + k] = l;
        } else {
            d_hi[v * Hd + k - He] = h;
            d_lo[v * Hd + k - He] = l;
        }
    }
}

__global__ void __launch_bounds__(kThreads, 1)
proj_tc_kernel(const __grid_constant__ CUtensorMap w0_hi, const __grid_constant__ CUtensorMap w0_lo,
               const __grid_constant__ CUtensorMap w1_hi, const __grid_constant__ CUtensorMap w1_lo,
               Problem p0, Problem p1, int V, int NB) {
    extern __shared__ __align__(128) unsigned char smem[];
    pdl_launch_dependents();
    const SmemP L = smem_layout_p(NB);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool second = (int)blockIdx.x >= p0.tiles;
    const Problem P = second ? p1 : p0;
    const int row0 = (second ? blockIdx.x - p0.tiles : blockIdx.x) * 128;
    const int n_kb = P.K / kKB;

    const uint32_t sbase = smem_u32(smem);
    const uint32_t a_base = sbase + L.a, w_base = sbase + L.w, bars = sbase + L.bars;
    auto a_full = [&](int i) { return bars + 8 * i; };
    auto a_empty = [&](int i) { return bars + 8 * (4 + i); };
    auto w_full = [&](int i) { return bars + 8 * (8 + i); };
    auto w_empty = [&](int i) { return bars + 8 * (12 + i); };
    const uint32_t acc_full = bars + 8 * 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bars + 24 * 8);

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(a_full(i), kProducerWarps);
            mbar_init(a_empty(i), 1);
            mbar_init(w_full(i), 1);
            mbar_init(w_empty(i), 1);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
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
    pdl_wait();  // the prologue above overlapped the predecessor (the weight split); its output is read from here on

    if (warp < kProducerWarps) {
        // ===== A producers: thread = (row r, K chunks kc0 and kc0+4 of each block) =====
        // a warp covers 8 rows x 4 chunks: one load instruction touches 8 rows x 128 contiguous bytes
        // (8 cache lines; a lane-per-row mapping touches 32 and is bound by L1 tag look-ups), and a
        // quarter warp still stores 8 rows x 16 B = one 128-byte core-matrix column, conflict-free
        const int r = warp * 8 + (lane & 7), kc0 = lane >> 3;
        const int row = min(row0 + r, P.rows - 1);  // rows past the end are computed but never stored
        const size_t xrow = (size_t)row * P.K;  // element offset of this thread's row
        // three register buffers take turns (K loop unrolled by three, compile-time roles, no copy
        // ever waits on a load): block kb+2 is requested while block kb is computed
        float4 xb0[4], xb1[4], xb2[4];
        auto load_x = [&](float4(&dst)[4], int kb) {
            if (kb < n_kb) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    dst[2 * i] = ldx4(P.x, P.xdt, xrow + kb * kKB + (kc0 + 4 * i) * 8);
                    dst[2 * i + 1] = ldx4(P.x, P.xdt, xrow + kb * kKB + (kc0 + 4 * i) * 8 + 4);
                }
            }
        };
        auto block = [&](const float4(&x)[4], int kb) {
            const int st = kb % kStages;
            mbar_wait(a_empty(st), ((kb / kStages) & 1) ^ 1);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int kc = kc0 + 4 * i;
                const float4 x0 = x[2 * i], x1 = x[2 * i + 1];
                uint4 hi, lo;
                split_store(gelu_tanh(x0.x), gelu_tanh(x0.y), hi.x, lo.x);
                split_store(gelu_tanh(x0.z), gelu_tanh(x0.w), hi.y, lo.y);
                split_store(gelu_tanh(x1.x), gelu_tanh(x1.y), hi.z, lo.z);
                split_store(gelu_tanh(x1.z), gelu_tanh(x1.w), hi.w, lo.w);
                unsigned char* dst = smem + L.a + st * 2 * kAHalf + kc * 2048 + r * 16;
                *reinterpret_cast<uint4*>(dst) = hi;
                *reinterpret_cast<uint4*>(dst + kAHalf) = lo;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full(st));
        };
        load_x(xb0, 0);
        load_x(xb1, 1);
        for (int kb = 0; kb < n_kb; kb += 3) {
            load_x(xb2, kb + 2);
            block(xb0, kb);
            if (kb + 1 < n_kb) {
                load_x(xb0, kb + 3);
                block(xb1, kb + 1);
            }
            if (kb + 2 < n_kb) {
                load_x(xb1, kb + 4);
                block(xb2, kb + 2);
            }
        }
    } else if (warp == kTmaWarp) {
        if (lane == 0) {
            const CUtensorMap* mh = second ? &w1_hi : &w0_hi;
            const CUtensorMap* ml = second ? &w1_lo : &w0_lo;
            for (int kb = 0; kb < n_kb; ++kb) {
                const int st = kb % kStages;
                mbar_wait(w_empty(st), ((kb / kStages) & 1) ^ 1);
                mbar_arrive_expect_tx(w_full(st), 2u * (uint32_t)L.w_half);
                tma_load_3d(w_base + st * 2 * L.w_half, mh, 0, 0, kb * (kKB / 8), w_full(st));
                tma_load_3d(w_base + st * 2 * L.w_half + L.w_half, ml, 0, 0, kb * (kKB / 8), w_full(st));
            }
        }
    } else if (warp == kMmaWarp) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(NB, false, false);
            const uint32_t b_lbo = NB * 16;
            for (int kb = 0; kb < n_kb; ++kb) {
                const int st = kb % kStages;
                mbar_wait(a_full(st), (kb / kStages) & 1);
                mbar_wait(w_full(st), (kb / kStages) & 1);
                tc_fence_after();
                const uint32_t ah = a_base + st * 2 * kAHalf, al = ah + kAHalf;
                const uint32_t wh = w_base + st * 2 * L.w_half, wl = wh + L.w_half;
#pragma unroll
                for (int j = 0; j < kKB / 16; ++j) {
                    const uint64_t dah = umma_desc(ah + j * 2 * 2048, 2048, 128), dal = umma_desc(al + j * 2 * 2048, 2048, 128);
                    const uint64_t dwh = umma_desc(wh + j * 2 * b_lbo, b_lbo, 128), dwl = umma_desc(wl + j * 2 * b_lbo, b_lbo, 128);
                    umma_bf16(tmem, dal, dwh, idesc, (kb | j) != 0);  // small terms first
                    umma_bf16(tmem, dah, dwl, idesc, 1);
                    umma_bf16(tmem, dah, dwh, idesc, 1);
                }
                umma_commit(a_empty(st));
                umma_commit(w_empty(st));
            }
            umma_commit(acc_full);
        }
    } else {
        // ===== epilogue: one output row per thread (= TMEM lane) -> shared-memory tile -> coalesced copy
        // The [rows, V] tile is one contiguous span of the output; a thread-per-row store would touch
        // 32 cache lines per instruction.  The A stages are free once the accumulator is complete.
        const int q = warp & 3, r = q * 32 + lane;
        float* stile = reinterpret_cast<float*>(smem + L.a);  // [128][Vp], odd row stride: conflict-free
        const int Vp = V | 1;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        for (int pc = 0; pc < NB / 16; ++pc) {
            float v[16];
            tmem_ld16(tmem + pc * 16 + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int col = pc * 16 + i;
                if (col < V) stile[r * Vp + col] = v[i] + (P.bias ? __ldg(P.bias + col) : 0.f);
            }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps
        const int n_rows = min(128, P.rows - row0);
        const int n = n_rows * V, t = threadIdx.x - (kThreads - 128);
        float* gout = P.out + (size_t)row0 * V;
        if (Vp == V && (n & 3) == 0 && ((uintptr_t)gout & 15) == 0) {  // the tile is flat in both memories
            for (int i = t; i < n / 4; i += 128)
                reinterpret_cast<float4*>(gout)[i] = reinterpret_cast<const float4*>(stile)[i];
        } else {
            for (int i = t; i < n; i += 128) {
                const int rr = i / V;
                gout[i] = stile[rr * Vp + (i - rr * V)];
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// TMA view [K/8][V][8] of a dense bf16 [V, K] matrix, box {8, NB, 8}: lands as UMMA core matrices
bool make_w_map(CUtensorMap* map, void* w, int V, int K, int NB) {
    EncodeTiledFn encode = encode_fn();
    if (!encode) return false;
    const cuuint64_t gdim[3] = {8, (cuuint64_t)V, (cuuint64_t)(K / 8)};
    const cuuint64_t gstride[2] = {(cuuint64_t)K * 2, 16};
    const cuuint32_t box[3] = {8, (cuuint32_t)NB, (cuuint32_t)(kKB / 8)};
    const cuuint32_t estr[3] = {1, 1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, w, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool proj_tc_supported(int V, int He, int Hd) {
    return V >= 1 && V <= 80 && He >= kKB && Hd >= kKB && He % kKB == 0 && Hd % kKB == 0;
}

size_t proj_tc_workspace_bytes(int V, int He, int Hd) {
    return 2 * align256((size_t)V * He * 2) + 2 * align256((size_t)V * Hd * 2);
}

// Splits the weight into the workspace (unless the caller says it already holds this weight's
// split, e.g. the backward reusing the forward's workspace) and builds the four TMA descriptors.
int proj_tc_prepare(const float* weight, int V, int He, int Hd, int NB, void* workspace, size_t workspace_bytes,
                    bool already_split, CUtensorMap maps[4], cudaStream_t stream) {
    if (!workspace || workspace_bytes < proj_tc_workspace_bytes(V, He, Hd) || ((uintptr_t)workspace & 15))
        return RNNTB200_STATUS_INVALID_VALUE;
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    __nv_bfloat16* e_hi = reinterpret_cast<__nv_bfloat16*>(ws);
    __nv_bfloat16* e_lo = reinterpret_cast<__nv_bfloat16*>(ws + align256((size_t)V * He * 2));
    __nv_bfloat16* d_hi = reinterpret_cast<__nv_bfloat16*>(ws + 2 * align256((size_t)V * He * 2));
    __nv_bfloat16* d_lo = reinterpret_cast<__nv_bfloat16*>(ws + 2 * align256((size_t)V * He * 2) + align256((size_t)V * Hd * 2));
    if (!already_split)
        (void)launch_pdl(pdl_ok(1), split_weight_kernel, dim3(std::min((V * (He + Hd) + 255) / 256, 592)), dim3(256), (size_t)0,
                         stream, weight, V, He, Hd, e_hi, e_lo, d_hi, d_lo);
    if (!make_w_map(&maps[0], e_hi, V, He, NB) || !make_w_map(&maps[1], e_lo, V, He, NB) ||
        !make_w_map(&maps[2], d_hi, V, Hd, NB) || !make_w_map(&maps[3], d_lo, V, Hd, NB))
        return RNNTB200_STATUS_EXECUTION_FAILED;
    return launch_status();
}

int launch_proj_tc(const void* enc, const void* dec, int x_dtype, const float* weight, const float* bias, int rows_enc,
                   int rows_dec, int He, int Hd, int V, float* penc, float* pdec, void* workspace,
                   size_t workspace_bytes, cudaStream_t stream) {
    if (!proj_tc_supported(V, He, Hd)) return RNNTB200_STATUS_INVALID_VALUE;
    const int NB = ((V + 15) / 16) * 16;
    CUtensorMap m[4];
    int st = proj_tc_prepare(weight, V, He, Hd, NB, workspace, workspace_bytes, false, m, stream);
    if (st != RNNTB200_STATUS_SUCCESS) return st;
    if (rows_enc + rows_dec == 0) return RNNTB200_STATUS_SUCCESS;
    Problem p0{enc, bias, penc, rows_enc, He, (rows_enc + 127) / 128, x_dtype};
    Problem p1{dec, nullptr, pdec, rows_dec, Hd, (rows_dec + 127) / 128, x_dtype};
    const SmemP L = smem_layout_p(NB);
    cudaError_t e = cudaFuncSetAttribute(proj_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    if (e != cudaSuccess) return status_from_cuda(e);
    e = launch_pdl(pdl_ok(rows_enc), proj_tc_kernel, dim3(p0.tiles + p1.tiles), dim3(kThreads), (size_t)L.total, stream, m[0], m[1],
                   m[2], m[3], p0, p1, V, NB);
    return e == cudaSuccess ? launch_status() : status_from_cuda(e);
}

}  // namespace rnntb200
