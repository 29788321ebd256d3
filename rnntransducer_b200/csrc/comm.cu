// comm.cu -- one-shot gradient all-reduce over NVLink peer memory (the path's only
// collective is the all-reduce of the fc gradients, 299 KB at cfg 2).
//
// NCCL's all-reduce costs a host-side launch per step plus a multi-kernel protocol tuned for bandwidth; at
// 300 KB the payload is ~1 us of NVLink time and everything else is latency, fully exposed because the fc
// gradients leave the LAST kernel of the step.  Here every rank owns one peer-mapped buffer (cudaMalloc +
// cudaIpc handles exchanged once through torch.distributed), and ONE kernel per step does:
//   1. copy this rank's gradient segments into its own buffer (parity half `step & 1`),
//   2. publish: system-scope fence, then store the step number into slot [rank][cta] of EVERY peer's flag
//      array (remote stores),
//   3. wait until its own flag slots [p][cta] of all peers p show this step (local loads only),
//   4. read the peers' halves through NVLink (cache-volatile loads), add them in rank order -- every rank
//      computes the same sum in the same order, bit-identical results -- scale, write back in place.
// CTA c only ever talks to CTA c of the peers, so there is no grid-wide synchronisation; the step number is
// a per-CTA counter in device memory, so the launch is parameter-free across steps and can sit inside the
// step's CUDA graph: no host work per step at all.  Double buffering by step parity is enough: a peer can
// overwrite the half a slow rank still reads only two steps later, i.e. after that slow rank has published
// the step in between, which it does after its reads (program order).
#include <cstring>

#include "common.cuh"

namespace rnntb200 {
namespace {

constexpr int kMaxRanks = 8;
constexpr int kCommCtas = 64;       // CTA c of every rank handles slice c of the vector
constexpr int kCommThreads = 256;
constexpr int kMaxSegments = 4;

struct CommHeader {                 // at the start of every rank's buffer
    unsigned int flags[kMaxRanks][kCommCtas];   // flags[p][c]: last step rank p's CTA c has published
    unsigned int step[kCommCtas];               // this rank's CTA c: steps done
    unsigned int pad[64];
};
static_assert(sizeof(CommHeader) % 256 == 0, "data halves stay 256-byte aligned");

struct CommArgs {
    unsigned char* peer[kMaxRanks];  // every rank's buffer as mapped into THIS process (peer[rank] = own)
    float* seg[kMaxSegments];        // gradient segments, reduced in place
    int seg_n[kMaxSegments];
    int n_seg, rank, world;
    size_t half_floats;              // floats per parity half
    float scale;
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_cv(const float* p) {  // never served from a stale cache line
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kCommThreads)
peer_allreduce_kernel(CommArgs A) {
    __shared__ unsigned int s_step;
    const int c = blockIdx.x, tid = threadIdx.x;
    CommHeader* own = reinterpret_cast<CommHeader*>(A.peer[A.rank]);
    if (tid == 0) s_step = own->step[c] + 1;
    __syncthreads();
    const unsigned int step = s_step;
    int total = 0;
    for (int s = 0; s < A.n_seg; ++s) total += A.seg_n[s];
    const int per = (total + kCommCtas - 1) / kCommCtas;
    const int lo = min(c * per, total), hi = min(lo + per, total);
    const size_t half = (size_t)(step & 1) * A.half_floats;
    float* mine = reinterpret_cast<float*>(A.peer[A.rank] + sizeof(CommHeader)) + half;

    // 1. this rank's slice -> its own buffer
    for (int i = lo + tid; i < hi; i += kCommThreads) {
        int k = i, s = 0;
        while (k >= A.seg_n[s]) k -= A.seg_n[s++];
        mine[i] = A.seg[s][k];
    }
    __syncthreads();
    // 2. publish to every peer (and to ourselves)
    if (tid < A.world) {
        __threadfence_system();
        CommHeader* ph = reinterpret_cast<CommHeader*>(A.peer[tid]);
        st_release_sys(&ph->flags[A.rank][c], step);
    }
    // 3. wait for every peer's slice c of this step
    if (tid < A.world) {
        const unsigned int* f = &own->flags[tid][c];
        long long t0 = 0;
        for (unsigned spins = 0;; ++spins) {
            if ((int)(ld_acquire_sys(f) - step) >= 0) break;
            if ((spins & 1023) == 1023) {  // a dead peer must fail the launch, not hang the GPU
                const long long now = clock64();
                if (t0 == 0) t0 = now;
                else if (now - t0 > 20000000000LL) __trap();
                __nanosleep(200);
            }
        }
    }
    __syncthreads();
    // 4. sum in rank order, scale, write back in place
    for (int i = lo + tid; i < hi; i += kCommThreads) {
        float acc = 0.f;
        for (int p = 0; p < A.world; ++p)
            acc += ld_cv(reinterpret_cast<const float*>(A.peer[p] + sizeof(CommHeader)) + half + i);
        int k = i, s = 0;
        while (k >= A.seg_n[s]) k -= A.seg_n[s++];
        A.seg[s][k] = acc * A.scale;
    }
    if (tid == 0) own->step[c] = step;
}

}  // namespace
}  // namespace rnntb200

using namespace rnntb200;

extern "C" {

RNNTB200_API size_t rnntb200_comm_buffer_bytes(size_t max_floats) {
    const size_t half = (max_floats + 63) / 64 * 64;
    return sizeof(CommHeader) + 2 * half * sizeof(float);
}

RNNTB200_API int rnntb200_comm_alloc(size_t bytes, void** dev_ptr) {
    if (!dev_ptr || bytes < sizeof(CommHeader)) return RNNTB200_STATUS_INVALID_VALUE;
    cudaError_t e = cudaMalloc(dev_ptr, bytes);
    if (e != cudaSuccess) return status_from_cuda(e);
    e = cudaMemset(*dev_ptr, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    return e == cudaSuccess ? RNNTB200_STATUS_SUCCESS : RNNTB200_STATUS_MEMOPS_FAILED;
}

RNNTB200_API int rnntb200_comm_free(void* dev_ptr) { return status_from_cuda(cudaFree(dev_ptr)); }

RNNTB200_API int rnntb200_comm_export(void* dev_ptr, unsigned char* handle64) {
    if (!dev_ptr || !handle64) return RNNTB200_STATUS_INVALID_VALUE;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, dev_ptr);
    if (e != cudaSuccess) return status_from_cuda(e);
    memcpy(handle64, &h, 64);
    return RNNTB200_STATUS_SUCCESS;
}

RNNTB200_API int rnntb200_comm_import(const unsigned char* handle64, void** peer_ptr) {
    if (!handle64 || !peer_ptr) return RNNTB200_STATUS_INVALID_VALUE;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    return status_from_cuda(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
}

RNNTB200_API int rnntb200_comm_release(void* peer_ptr) { return status_from_cuda(cudaIpcCloseMemHandle(peer_ptr)); }

RNNTB200_API int rnntb200_comm_allreduce(void* const* peer_ptrs, int rank, int world, float* const* segments,
                                         const int* segment_floats, int n_segments, size_t max_floats,
                                         float scale, void* stream) {
    if (!peer_ptrs || !segments || !segment_floats || world < 1 || world > kMaxRanks || rank < 0 || rank >= world ||
        n_segments < 1 || n_segments > kMaxSegments)
        return RNNTB200_STATUS_INVALID_VALUE;
    CommArgs A = {};
    size_t total = 0;
    for (int p = 0; p < world; ++p) {
        if (!peer_ptrs[p]) return RNNTB200_STATUS_INVALID_VALUE;
        A.peer[p] = (unsigned char*)peer_ptrs[p];
    }
    for (int s = 0; s < n_segments; ++s) {
        if (!segments[s] || segment_floats[s] < 0) return RNNTB200_STATUS_INVALID_VALUE;
        A.seg[s] = segments[s];
        A.seg_n[s] = segment_floats[s];
        total += (size_t)segment_floats[s];
    }
    if (total > max_floats || total > 0x7fffffffu) return RNNTB200_STATUS_INVALID_VALUE;
    A.n_seg = n_segments, A.rank = rank, A.world = world, A.scale = scale;
    A.half_floats = (max_floats + 63) / 64 * 64;
    peer_allreduce_kernel<<<kCommCtas, kCommThreads, 0, (cudaStream_t)stream>>>(A);
    return launch_status();
}

}  // extern "C"
