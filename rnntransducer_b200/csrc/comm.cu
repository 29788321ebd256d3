// comm.cu -- one-shot gradient all-reduce over NVLink peer memory (the path's only
// collective is the all-reduce of the fc gradients, 299 KB at cfg 2).
//
// NCCL's all-reduce costs a host-side launch per step plus a protocol tuned for bandwidth; at 300 KB the
// payload is ~1 us of NVLink time and everything else is latency, fully exposed because the fc gradients
// leave the LAST kernel of the step.  Here every rank owns one peer-mapped buffer (cudaMalloc + cudaIpc
// handles exchanged once through torch.distributed) with one slot per source rank, and ONE kernel per step
// does, CTA c of every rank for slice c of the vector:
//   1. PUSH: read this rank's slice of the gradient segments once and store it into slot [rank] of EVERY
//      rank's buffer (posted remote stores over NVLink: nobody waits for a round trip),
//   2. publish: block barrier, system-scope fence, then the step number into flag [rank][c] of every peer,
//   3. wait until its OWN flags [p][c] show this step for every peer p (local loads only),
//   4. add the W slots of its OWN buffer in rank order -- every rank computes the same sum in the same
//      order, bit-identical results -- scale, write back in place (local loads, L1 bypassed).
// CTA c only ever talks to CTA c of the peers, so there is no grid-wide synchronisation; the step number is
// a per-CTA counter in device memory, so the launch is parameter-free across steps and can sit inside the
// step's CUDA graph: no host work per step at all.  Double buffering by step parity is enough: a peer can
// overwrite a slot a slow rank still reads only two steps later, i.e. after that slow rank has published
// the step in between, which it does after its reads (program order).
#include <cstring>

#include "common.cuh"

namespace rnntb200 {
namespace {

constexpr int kMaxRanks = 8;
constexpr int kCommCtas = 74;       // CTA c of every rank handles slice c of the vector (half the SMs)
constexpr int kCommThreads = 256;
constexpr int kMaxSegments = 4;

struct CommHeader {                 // at the start of every rank's buffer
    unsigned int flags[kMaxRanks][128];   // flags[p][c]: last step rank p's CTA c has pushed completely
    unsigned int step[128];               // this rank's CTA c: steps done
};
static_assert(sizeof(CommHeader) % 256 == 0 && kCommCtas <= 128, "slots stay 256-byte aligned");

struct CommArgs {
    unsigned char* peer[kMaxRanks];  // every rank's buffer as mapped into THIS process (peer[rank] = own)
    float* seg[kMaxSegments];        // gradient segments, reduced in place
    int seg_n[kMaxSegments];         // floats
    int seg_q0[kMaxSegments + 1];    // first float4 of each segment in the (4-float padded) concatenation
    int n_seg, rank, world, vec_ok;
    size_t slot_floats;              // floats per (parity, source rank) slot, multiple of 64
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
__device__ __forceinline__ float4 ld_cv4(const float4* p) {  // never served from a stale L1 line
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4* slot_of(const CommArgs& A, int dst_rank, unsigned parity, int src_rank) {
    return reinterpret_cast<float4*>(A.peer[dst_rank] + sizeof(CommHeader)) +
           ((size_t)parity * A.world + src_rank) * (A.slot_floats / 4);
}
// float4 q of the padded concatenation: which segment, which float inside it
__device__ __forceinline__ void locate(const CommArgs& A, int q, int& s, int& k) {
    s = 0;
    while (s + 1 < A.n_seg && q >= A.seg_q0[s + 1]) ++s;
    k = (q - A.seg_q0[s]) * 4;
}
__device__ __forceinline__ float4 seg_load(const CommArgs& A, int q) {
    int s, k;
    locate(A, q, s, k);
    const float* p = A.seg[s] + k;
    if (A.vec_ok && k + 4 <= A.seg_n[s]) return *reinterpret_cast<const float4*>(p);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int n = A.seg_n[s] - k;
    if (n > 0) v.x = p[0];
    if (n > 1) v.y = p[1];
    if (n > 2) v.z = p[2];
    if (n > 3) v.w = p[3];
    return v;
}
__device__ __forceinline__ void seg_store(const CommArgs& A, int q, float4 v) {
    int s, k;
    locate(A, q, s, k);
    float* p = A.seg[s] + k;
    if (A.vec_ok && k + 4 <= A.seg_n[s]) {
        *reinterpret_cast<float4*>(p) = v;
        return;
    }
    const int n = A.seg_n[s] - k;
    if (n > 0) p[0] = v.x;
    if (n > 1) p[1] = v.y;
    if (n > 2) p[2] = v.z;
    if (n > 3) p[3] = v.w;
}

__global__ void __launch_bounds__(kCommThreads)
peer_allreduce_kernel(CommArgs A) {
    __shared__ unsigned int s_step;
    pdl_wait();  // the gradients come from the step's last kernel
    const int c = blockIdx.x, tid = threadIdx.x;
    CommHeader* own = reinterpret_cast<CommHeader*>(A.peer[A.rank]);
    if (tid == 0) s_step = own->step[c] + 1;
    __syncthreads();
    const unsigned int step = s_step, parity = step & 1;
    const int total = A.seg_q0[A.n_seg];  // float4s
    const int per = (total + kCommCtas - 1) / kCommCtas;
    const int lo = min(c * per, total), hi = min(lo + per, total);

    // 1. push this rank's slice into slot [rank] of every rank's buffer
    for (int q = lo + tid; q < hi; q += kCommThreads) {
        const float4 v = seg_load(A, q);
        for (int p = 0; p < A.world; ++p) slot_of(A, p, parity, A.rank)[q] = v;
    }
    __syncthreads();
    // 2. publish to every peer (and to ourselves)
    if (tid < A.world) {
        __threadfence_system();
        CommHeader* ph = reinterpret_cast<CommHeader*>(A.peer[tid]);
        st_release_sys(&ph->flags[A.rank][c], step);
        // 3. wait for every peer's slice c of this step
        const unsigned int* f = &own->flags[tid][c];
        long long t0 = 0;
        for (unsigned spins = 0;; ++spins) {
            if ((int)(ld_acquire_sys(f) - step) >= 0) break;
            if ((spins & 1023) == 1023) {  // a dead peer must fail the launch, not hang the GPU
                const long long now = clock64();
                if (t0 == 0) t0 = now;
                else if (now - t0 > 20000000000LL) __trap();
                __nanosleep(100);
            }
        }
    }
    __syncthreads();
    // 4. sum the slots of OUR buffer in rank order, scale, write back in place
    for (int q = lo + tid; q < hi; q += kCommThreads) {
        float4 v[kMaxRanks];
#pragma unroll
        for (int p = 0; p < kMaxRanks; ++p)
            if (p < A.world) v[p] = ld_cv4(slot_of(A, A.rank, parity, p) + q);
        float4 acc = v[0];
#pragma unroll
        for (int p = 1; p < kMaxRanks; ++p)
            if (p < A.world) acc.x += v[p].x, acc.y += v[p].y, acc.z += v[p].z, acc.w += v[p].w;
        seg_store(A, q, make_float4(acc.x * A.scale, acc.y * A.scale, acc.z * A.scale, acc.w * A.scale));
    }
    if (tid == 0) own->step[c] = step;
}

inline size_t slot_floats_for(size_t max_floats) {  // room for kMaxSegments segments each padded to 4 floats
    return (max_floats + 4 * kMaxSegments + 63) / 64 * 64;
}

}  // namespace
}  // namespace rnntb200

using namespace rnntb200;

extern "C" {

RNNTB200_API size_t rnntb200_comm_buffer_bytes(size_t max_floats, int world) {
    if (world < 1 || world > kMaxRanks) return 0;
    return sizeof(CommHeader) + 2 * (size_t)world * slot_floats_for(max_floats) * sizeof(float);
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
    A.vec_ok = 1;
    int q0 = 0;
    for (int s = 0; s < n_segments; ++s) {
        if (!segments[s] || segment_floats[s] < 0) return RNNTB200_STATUS_INVALID_VALUE;
        A.seg[s] = segments[s];
        A.seg_n[s] = segment_floats[s];
        A.seg_q0[s] = q0;
        q0 += (segment_floats[s] + 3) / 4;
        total += (size_t)segment_floats[s];
        if ((uintptr_t)segments[s] % 16 != 0) A.vec_ok = 0;
    }
    A.seg_q0[n_segments] = q0;
    if (total > max_floats || total > 0x3fffffffu) return RNNTB200_STATUS_INVALID_VALUE;
    A.n_seg = n_segments, A.rank = rank, A.world = world, A.scale = scale;
    A.slot_floats = slot_floats_for(max_floats);
    const cudaError_t e = launch_pdl(pdl_ok(1), peer_allreduce_kernel, dim3(kCommCtas), dim3(kCommThreads), (size_t)0,
                                     (cudaStream_t)stream, A);
    return e == cudaSuccess ? launch_status() : status_from_cuda(e);
}

}  // extern "C"
