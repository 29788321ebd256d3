// joint_at_tc.cu -- tcgen05 forward of the ADD_TANH joint: logits = tanh(enc_t + dec_u) W^T + bias
// on the 5th-gen tensor cores with the log-softmax fused into the accumulator read-out, so that
// only (lp_blank, lp_label) and the log-sum-exp of each lattice cell ever reach HBM.
//
// One persistent CTA per SM walks tiles of 128 lattice cells (16 t x 8 u of one utterance):
//
//   warps 0-15  A producers: z = tanh(e_t + d_u) on the CUDA cores (MUFU.TANH), rounded to bf16 and
//               written straight into the UMMA K-major core-matrix layout, one 64-wide K block per
//               shared-memory slot (the A operand is computed, never loaded);
//   warp  16    TMA producer: streams W (bf16 copy, [V,H]) as [N-chunk x 64] K blocks through a
//               ring of shared-memory stages with cp.async.bulk.tensor (3-D view so that the
//               box lands directly in core-matrix order; rows >= V are zero-filled by the TMA);
//   warp  17    MMA issuer: one elected thread issues tcgen05.mma (M = 128 cells, N = chunk of the
//               vocabulary, K = 16 per instruction), accumulating in TMEM; tcgen05.commit releases
//               W stages / A slots and publishes the accumulator;
//   warps 18-21 epilogue: each thread owns one lattice cell (one TMEM lane), reads its logits with
//               tcgen05.ld, adds the bias and runs an online log-sum-exp across vocabulary chunks
//               (two accumulator stages in TMEM, so chunk c+1 is multiplied while chunk c is
//               reduced), picks the blank / label columns and writes 12 bytes per cell.
//
// Synchronisation is mbarrier-only (full/empty pairs per A slot, W stage and accumulator stage).
// Supported shapes: H a multiple of 64, H <= 512; any V (chunks of <= 128 columns, so V = 1024 is
// 8 chunks with the A tile resident in shared memory and computed once).
#include <algorithm>

#include "tc_common.cuh"

namespace rnntb200 {

using namespace tc;

namespace {

constexpr int kTT = 16, kUU = 8;       // tile = 16 frames x 8 label positions = 128 cells (MMA M)
constexpr int kKB = 64;                // K elements per A slot / W stage
constexpr int kProducerWarps = 16;
constexpr int kProducerThreads = kProducerWarps * 32;  // warps 0-7
constexpr int kTmaWarp = 16, kMmaWarp = 17;  // warps 10-13: epilogue
constexpr int kThreads = 22 * 32;
constexpr int kMaxWStages = 8;  // W K-blocks in flight (TMA latency ~1 us: a 2-deep ring starves the MMA)
constexpr int kAccStride = 128;        // TMEM columns per accumulator stage
constexpr int kTmemCols = 256;
constexpr int kASlotBytes = 128 * kKB * 2;  // 16 KiB
constexpr int kMaxSlots = 8;                // H <= 512

struct Smem {  // offsets into dynamic shared memory
    int a, w, dd, bars, total;
    int w_stage_bytes, dd_stride, w_stages;
};

__host__ __device__ inline Smem smem_layout(int H, int NB) {
    Smem s;
    s.a = 0;
    s.w = (H / kKB) * kASlotBytes;
    s.w_stage_bytes = NB * kKB * 2;
    s.dd_stride = H + 4;  // floats; +4: the 8 predictor rows start 4 banks apart, float4 reads conflict-free
    const int fixed = s.w + kUU * s.dd_stride * 4 + 40 * 8 + 32;
    s.w_stages = (227 * 1024 - fixed) / s.w_stage_bytes;
    s.w_stages = s.w_stages > kMaxWStages ? kMaxWStages : s.w_stages;
    s.dd = s.w + s.w_stages * s.w_stage_bytes;
    s.bars = (s.dd + kUU * s.dd_stride * 4 + 15) & ~15;
    s.total = s.bars + 40 * 8 + 16;
    return s;
}

__global__ void convert_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = __float2bfloat16_rn(src[i]);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
at_lse_tc_kernel(const __grid_constant__ CUtensorMap w_map, const float* __restrict__ enc,
                 const float* __restrict__ dec, const float* __restrict__ bias,
                 const int32_t* __restrict__ labels, const int32_t* __restrict__ act_lens,
                 const int32_t* __restrict__ label_lens, int B, int T, int U1, int V, int H, int NB,
                 int blank, float2* __restrict__ lp2, float* __restrict__ lse_out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Smem L = smem_layout(H, NB);
    const int n_slots = H / kKB;
    const int n_chunks = (V + NB - 1) / NB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t sbase = smem_u32(smem);
    const uint32_t a_base = sbase + L.a, w_base = sbase + L.w;
    float* dd = reinterpret_cast<float*>(smem + L.dd);  // [8 predictor rows][dd_stride]
    const int kWStages = L.w_stages;
    const uint32_t bars = sbase + L.bars;
    auto a_full = [&](int i) { return bars + 8 * i; };
    auto a_empty = [&](int i) { return bars + 8 * (8 + i); };
    auto w_full = [&](int i) { return bars + 8 * (16 + i); };
    auto w_empty = [&](int i) { return bars + 8 * (24 + i); };
    auto acc_full = [&](int i) { return bars + 8 * (32 + i); };
    auto acc_empty = [&](int i) { return bars + 8 * (34 + i); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bars + 40 * 8);

    if (threadIdx.x == 0) {
        for (int i = 0; i < kMaxSlots; ++i) { mbar_init(a_full(i), kProducerWarps); mbar_init(a_empty(i), 1); }
        for (int i = 0; i < kMaxWStages; ++i) { mbar_init(w_full(i), 1); mbar_init(w_empty(i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), 4); }
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
    const uint32_t tmem_base = *tmem_slot;

    const int nT = (T + kTT - 1) / kTT, nU = (U1 + kUU - 1) / kUU;
    const int n_tiles = B * nT * nU;

    // tile -> (b, t0, u0) and whether any of its cells is inside the utterance's lattice
    auto decode = [&](int tile, int& b, int& t0, int& u0) -> bool {
        b = tile / (nT * nU);
        const int r = tile - b * nT * nU;
        t0 = (r / nU) * kTT;
        u0 = (r % nU) * kUU;
        return t0 < len_T(act_lens, b, T) && u0 <= len_U(label_lens, b, U1);
    };

    if (warp < kProducerWarps) {
        // ===== A producers =====
        const int p = threadIdx.x;
        const int r = p & 127, kc0 = p >> 7;  // cell row of the tile, first 8-wide K chunk
        const int tt = r / kUU, uu = r % kUU;
        uint32_t n = 0;  // tiles processed by this CTA so far
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            int b, t0, u0;
            if (!decode(tile, b, t0, u0)) continue;
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
            // This thread's 2 x 8 encoder values per K block live in three register buffers that take
            // turns (the K loop is unrolled by three so the roles are compile-time: no register copy
            // ever has to wait for a load): the values of block kb+2 are requested while block kb is
            // computed, which covers the L2 latency with two blocks of tanh work.
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
                mbar_wait(a_empty(kb), (n & 1) ^ 1);
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
                    // core-matrix layout: K chunk kc at kc * 2048, cell row r at r * 16
                    *reinterpret_cast<uint4*>(smem + L.a + kb * kASlotBytes + kc * 2048 + r * 16) = out;
                }
                fence_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(a_full(kb));
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
        // ===== TMA producer for W =====
        if (lane == 0) {
            uint32_t wi = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                int b, t0, u0;
                if (!decode(tile, b, t0, u0)) continue;
                for (int c = 0; c < n_chunks; ++c)
                    for (int kb = 0; kb < n_slots; ++kb, ++wi) {
                        const int st = wi % kWStages;
                        mbar_wait(w_empty(st), ((wi / kWStages) & 1) ^ 1);
                        mbar_arrive_expect_tx(w_full(st), (uint32_t)L.w_stage_bytes);
                        tma_load_3d(w_base + st * L.w_stage_bytes, &w_map, 0, c * NB, kb * (kKB / 8), w_full(st));
                    }
            }
        }
    } else if (warp == kMmaWarp) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // kind::f16, A/B = bf16 K-major, D = fp32, M = 128, N = NB
            const uint32_t idesc = umma_idesc_bf16(NB, false, false);
            const uint32_t b_lbo = NB * 16;
            uint32_t n = 0, wi = 0, ci = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                int b, t0, u0;
                if (!decode(tile, b, t0, u0)) continue;
                for (int c = 0; c < n_chunks; ++c, ++ci) {
                    const int as = ci & 1;
                    mbar_wait(acc_empty(as), ((ci >> 1) & 1) ^ 1);
                    tc_fence_after();
                    for (int kb = 0; kb < n_slots; ++kb, ++wi) {
                        const int st = wi % kWStages;
                        if (c == 0) mbar_wait(a_full(kb), n & 1);
                        mbar_wait(w_full(st), (wi / kWStages) & 1);
                        tc_fence_after();
#pragma unroll
                        for (int j = 0; j < kKB / 16; ++j) {
                            const uint64_t ad = umma_desc(a_base + kb * kASlotBytes + j * 2 * 2048, 2048, 128);
                            const uint64_t bd = umma_desc(w_base + st * L.w_stage_bytes + j * 2 * b_lbo, b_lbo, 128);
                            umma_bf16(tmem_base + as * kAccStride, ad, bd, idesc, (kb | j) != 0);
                        }
                        umma_commit(w_empty(st));
                        if (c == n_chunks - 1) umma_commit(a_empty(kb));
                    }
                    umma_commit(acc_full(as));
                }
                ++n;
            }
        }
    } else {
        // ===== epilogue: one lattice cell per thread =====
        const int q = warp & 3;               // TMEM lane quarter this warp may read
        const int r = q * 32 + lane;          // cell row of the tile
        const int tt = r / kUU, uu = r % kUU;
        uint32_t ci = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            int b, t0, u0;
            if (!decode(tile, b, t0, u0)) continue;
            const int Tb = len_T(act_lens, b, T), Ub = len_U(label_lens, b, U1);
            const int t = t0 + tt, u = u0 + uu;
            const bool valid = t < Tb && u <= Ub;
            const int y = (valid && u < Ub) ? label_at(labels, b, U1, u, V) : -1;
            float m = -INFINITY, s = 0.f, xb = 0.f, xl = 0.f;
            for (int c = 0; c < n_chunks; ++c, ++ci) {
                const int as = ci & 1;
                mbar_wait(acc_full(as), (ci >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + as * kAccStride + ((uint32_t)(q * 32) << 16);
                // one pass over the chunk: online log-sum-exp with one rescale per 16 columns
                for (int pc = 0; pc < NB / 16; ++pc) {
                    float v[16];
                    tmem_ld16(taddr + pc * 16, v);
                    float pmax = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int col = c * NB + pc * 16 + i;
                        v[i] = col < V ? fmaf(v[i], kLog2e, __ldg(bias + col) * kLog2e) : -INFINITY;
                        pmax = fmaxf(pmax, v[i]);
                        if (col == blank) xb = v[i];
                        if (col == y) xl = v[i];
                    }
                    const float m_new = fmaxf(m, pmax);  // finite: the first piece always has col < V
                    float ps = 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) ps += fast_ex2(v[i] - m_new);
                    s = fmaf(s, fast_ex2(m - m_new), ps);
                    m = m_new;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty(as));
            }
            if (valid) {
                const float lse2 = m + fast_lg2(s);
                const size_t cidx = ((size_t)b * T + t) * U1 + u;
                const float lb = fmaxf((xb - lse2) * kLn2, kNegInf);
                const float ll = u < Ub ? fmaxf((xl - lse2) * kLn2, kNegInf) : 0.f;
                lp2[cidx] = make_float2(lb, ll);
                lse_out[cidx] = lse2 * kLn2;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
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

inline int chunk_cols(int V) { return std::min(((V + 15) / 16) * 16, 128); }

}  // namespace

bool at_tc_supported(int V, int H) { return V >= 1 && H >= kKB && H % kKB == 0 && H <= kKB * kMaxSlots; }

size_t at_tc_workspace_bytes(int V, int H) { return ((size_t)V * H * sizeof(__nv_bfloat16) + 255) & ~(size_t)255; }

// bf16 copy of W into the workspace + the TMA descriptor of its 3-D view [H/8][V][8]: a box
// {8, NB, 8} lands in shared memory as [8-element column group][row][8 elements], i.e. UMMA core
// matrices that can be read K-major (K = H) or MN-major (K = V).
int at_tc_prepare_weight(const float* weight, int V, int H, int NB, void* workspace, size_t workspace_bytes,
                         CUtensorMap* map, cudaStream_t stream) {
    if (!workspace || workspace_bytes < at_tc_workspace_bytes(V, H) || ((uintptr_t)workspace & 15))
        return RNNTB200_STATUS_INVALID_VALUE;
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return RNNTB200_STATUS_EXECUTION_FAILED;
    __nv_bfloat16* wb = static_cast<__nv_bfloat16*>(workspace);
    const size_t nw = (size_t)V * H;
    convert_bf16_kernel<<<(unsigned)std::min<size_t>((nw + 255) / 256, 1184), 256, 0, stream>>>(weight, wb, nw);
    const cuuint64_t gdim[3] = {8, (cuuint64_t)V, (cuuint64_t)(H / 8)};
    const cuuint64_t gstride[2] = {(cuuint64_t)H * 2, 16};
    const cuuint32_t box[3] = {8, (cuuint32_t)NB, (cuuint32_t)(kKB / 8)};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, wb, gdim, gstride, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return RNNTB200_STATUS_INVALID_VALUE;
    return launch_status();
}

int launch_at_lse_tc(const float* enc, const float* dec, const float* weight, const float* bias,
                     const int32_t* labels, const int32_t* act_lens, const int32_t* label_lens, int B,
                     int T, int U1, int V, int H, int blank, float2* lp2, float* lse, void* workspace,
                     size_t workspace_bytes, cudaStream_t stream) {
    if (!at_tc_supported(V, H)) return RNNTB200_STATUS_INVALID_VALUE;
    const int NB = chunk_cols(V);
    CUtensorMap map;
    int st = at_tc_prepare_weight(weight, V, H, NB, workspace, workspace_bytes, &map, stream);
    if (st != RNNTB200_STATUS_SUCCESS) return st;

    const Smem L = smem_layout(H, NB);
    cudaError_t e = cudaFuncSetAttribute(at_lse_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    if (e != cudaSuccess) return status_from_cuda(e);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_tiles = B * ((T + kTT - 1) / kTT) * ((U1 + kUU - 1) / kUU);
    const int grid = std::min(n_tiles, sms);
    at_lse_tc_kernel<<<grid, kThreads, L.total, stream>>>(map, enc, dec, bias, labels, act_lens, label_lens, B,
                                                         T, U1, V, H, NB, blank, lp2, lse);
    return launch_status();
}

}  // namespace rnntb200
