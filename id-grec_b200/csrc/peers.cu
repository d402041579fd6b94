// Peer-memory plumbing for the row-partitioned multi-GPU path (one process per GPU, one NVSwitch box).
//
// Every rank allocates one "slab" with cudaMalloc, exports it with CUDA IPC and maps all peers'
// slabs.  Buffers carved at identical offsets are then addressable on every GPU, which lets the
// SpMM epilogue store finished rows straight into the peers' layer buffers (csrc/spmm.cu,
// SpmmArgs::peerY) -- the per-layer all-gather of SURVEY.md section 8 e fused into the kernel that
// produces the rows.  Synchronisation is a device-side flag barrier over the same slabs; NCCL is
// used by the host side only for bootstrap and scalar reductions.
#include <string.h>

#include "idg_common.cuh"

namespace idg {
struct PeerPtrs { char* p[8]; };

__global__ void peer_push_kernel(const float4* __restrict__ src, PeerPtrs dst, int n_dst, int64_t n4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = src[i];
    for (int q = 0; q < n_dst; ++q) reinterpret_cast<float4*>(dst.p[q])[i] = v;
}
__global__ void mc_push_kernel(const float4* __restrict__ src, float* mc_dst, int64_t n4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    multimem_st4(mc_dst + 4 * i, src[i]);
}

// Persistent variants for the chunked exchange (idgrec/dist.py): a FEW CTAs stream a finished block of rows to the peers while
// the next block is still being computed by the propagation kernel on the other SMs.  Measured at the XL shape on 8 GPUs: with
// the peer stores inside the SpMM epilogue, a layer over 1/8 of the rows took 0.93 ms against 0.53 ms for the same layer without
// them -- NVLink back-pressure on the store path stalls the gathers queued behind it on every SM.
__global__ void __launch_bounds__(512) mc_push_persistent_kernel(const float4* __restrict__ src, float* mc_dst, int64_t n4) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (i + u * stride < n4) v[u] = src[i + u * stride];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (i + u * stride < n4) multimem_st4(mc_dst + 4 * (i + u * stride), v[u]);
    }
}
__global__ void __launch_bounds__(512) peer_push_persistent_kernel(const float4* __restrict__ src, PeerPtrs dst, int n_dst, int64_t n4) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = src[i];
        for (int q = 0; q < n_dst; ++q) reinterpret_cast<float4*>(dst.p[q])[i] = v;
    }
}

// state layout (ints, inside the slab): [0] = epoch counter (local), [1] = error word (0 = healthy, else
// 1 + index of the first peer that did not arrive within the time limit), [8 + r] = last epoch announced by rank r.
// The wait is bounded (a dead peer must surface as an error, not as a hung stream): after
// timeout_ns without the peer's flag the thread records the peer in state[1] and returns; every later barrier of a
// failed slab returns immediately, so a captured train step drains instead of spinning, and the host reads the
// error word with idg_peers_status at its next synchronisation point.
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__global__ void peer_barrier_kernel(int* state, PeerPtrs peer_states, int rank, int world, unsigned long long timeout_ns) {
    __shared__ int s_epoch;
    if (threadIdx.x == 0) { s_epoch = state[0] + 1; state[0] = s_epoch; }
    __syncthreads();
    const int e = s_epoch;
    __threadfence_system();
    if (threadIdx.x < world) {
        volatile int* f = reinterpret_cast<volatile int*>(peer_states.p[threadIdx.x]) + 8;
        f[rank] = e;
    }
    __threadfence_system();
    if (threadIdx.x < world) {
        volatile int* st = reinterpret_cast<volatile int*>(state);
        if (st[1] == 0) {
            const unsigned long long t0 = global_ns();
            unsigned spins = 0;
            while (st[8 + threadIdx.x] < e) {
                if ((++spins & 1023u) == 0 && global_ns() - t0 > timeout_ns) {
                    atomicCAS(state + 1, 0, 1 + (int)threadIdx.x);
                    break;
                }
            }
        }
    }
    __threadfence_system();
}
}  // namespace idg

using namespace idg;

extern "C" int idg_device_alloc(int64_t bytes, void** out) {
    if (!out || bytes <= 0) return fail(-1, "idg_device_alloc: bad argument%s");
    IDG_CUDA(cudaMalloc(out, (size_t)bytes));
    IDG_CUDA(cudaMemset(*out, 0, (size_t)bytes));
    return 0;
}
extern "C" int idg_device_free(void* p) {
    IDG_CUDA(cudaFree(p));
    return 0;
}
extern "C" int idg_ipc_get_handle(const void* d_ptr, void* handle64) {
    if (!d_ptr || !handle64) return fail(-1, "idg_ipc_get_handle: null argument%s");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    IDG_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), const_cast<void*>(d_ptr)));
    return 0;
}
extern "C" int idg_ipc_open(const void* handle64, void** out) {
    if (!handle64 || !out) return fail(-1, "idg_ipc_open: null argument%s");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    IDG_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int idg_ipc_close(void* p) {
    IDG_CUDA(cudaIpcCloseMemHandle(p));
    return 0;
}

extern "C" int idg_peers_create(void* local_base, int64_t bytes, int32_t rank, int32_t world, void* const* bases, idg_peers** out) {
    if (!local_base || !bases || !out || world < 1 || world > 8 || rank < 0 || rank >= world) return fail(-1, "idg_peers_create: bad argument%s");
    idg_peers* p = new idg_peers();
    p->local_base = (char*)local_base; p->bytes = bytes; p->rank = rank; p->world = world;
    for (int q = 0; q < world; ++q) p->bases[q] = (q == rank) ? (char*)local_base : (char*)bases[q];
    *out = p;
    return 0;
}
extern "C" void idg_peers_destroy(idg_peers* p) { delete p; }
extern "C" int idg_peers_set_multicast(idg_peers* p, void* mc_base) {
    if (!p) return fail(-1, "idg_peers_set_multicast: null argument%s");
    p->mc_base = (char*)mc_base;
    return 0;
}

static int slab_offset(const idg_peers* p, const void* ptr, int64_t bytes, int64_t* off) {
    const char* c = (const char*)ptr;
    if (c < p->local_base || c + bytes > p->local_base + p->bytes) return fail(-1, "pointer is not inside the peer slab%s");
    *off = c - p->local_base;
    return 0;
}

extern "C" int idg_peers_push(const idg_peers* p, const void* d_src, int64_t bytes, void* stream) {
    if (!p || !d_src || bytes < 0 || (bytes & 15) || ((uintptr_t)d_src & 15)) return fail(-1, "idg_peers_push: bad argument (16-byte granularity)%s");
    if (bytes == 0 || p->world == 1) return 0;
    int64_t off;
    if (int rc = slab_offset(p, d_src, bytes, &off)) return rc;
    const int64_t n4 = bytes / 16;
    if (p->mc_base) {
        mc_push_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)d_src, (float*)(p->mc_base + off), n4);
        IDG_LAUNCH_CHECK("mc_push_kernel");
        return 0;
    }
    PeerPtrs dst;
    int n = 0;
    for (int q = 0; q < p->world; ++q) if (q != p->rank) dst.p[n++] = p->bases[q] + off;
    peer_push_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)d_src, dst, n, n4);
    IDG_LAUNCH_CHECK("peer_push_kernel");
    return 0;
}

extern "C" int idg_peers_push_ctas(const idg_peers* p, const void* d_src, int64_t bytes, int32_t n_ctas, void* stream) {
    if (!p || !d_src || bytes < 0 || (bytes & 15) || ((uintptr_t)d_src & 15) || n_ctas < 1) return fail(-1, "idg_peers_push_ctas: bad argument (16-byte granularity)%s");
    if (bytes == 0 || p->world == 1) return 0;
    int64_t off;
    if (int rc = slab_offset(p, d_src, bytes, &off)) return rc;
    const int64_t n4 = bytes / 16;
    if (p->mc_base) {
        mc_push_persistent_kernel<<<(unsigned)n_ctas, 512, 0, (cudaStream_t)stream>>>((const float4*)d_src, (float*)(p->mc_base + off), n4);
        IDG_LAUNCH_CHECK("mc_push_persistent_kernel");
        return 0;
    }
    PeerPtrs dst;
    int n = 0;
    for (int q = 0; q < p->world; ++q) if (q != p->rank) dst.p[n++] = p->bases[q] + off;
    peer_push_persistent_kernel<<<(unsigned)n_ctas, 512, 0, (cudaStream_t)stream>>>((const float4*)d_src, dst, n, n4);
    IDG_LAUNCH_CHECK("peer_push_persistent_kernel");
    return 0;
}

extern "C" int idg_peers_set_timeout_ms(idg_peers* p, int64_t ms) {
    if (!p || ms <= 0) return fail(-1, "idg_peers_set_timeout_ms: bad argument%s");
    p->timeout_ns = (unsigned long long)ms * 1000000ull;
    return 0;
}

extern "C" int idg_peers_barrier(const idg_peers* p, int32_t* d_state, void* stream) {
    if (!p || !d_state) return fail(-1, "idg_peers_barrier: null argument%s");
    if (p->world == 1) return 0;
    int64_t off;
    if (int rc = slab_offset(p, d_state, 64 * (int64_t)sizeof(int), &off)) return rc;
    PeerPtrs st;
    for (int q = 0; q < 8; ++q) st.p[q] = (q < p->world) ? p->bases[q] + off : nullptr;
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_state, st, p->rank, p->world, p->timeout_ns);
    IDG_LAUNCH_CHECK("peer_barrier_kernel");
    return 0;
}

// Synchronises the stream and reads the barrier's error word: 0 = every barrier so far completed,
// IDG_ERR_PEER_TIMEOUT + r = rank r did not reach a barrier within the time limit (idg_last_error names it).
extern "C" int idg_peers_status(const idg_peers* p, const int32_t* d_state, void* stream) {
    if (!p || !d_state) return fail(-1, "idg_peers_status: null argument%s");
    int word = 0;
    IDG_CUDA(cudaMemcpyAsync(&word, d_state + 1, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    IDG_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (word == 0) return 0;
    return fail(IDG_ERR_PEER_TIMEOUT + (word - 1), "idg_peers_barrier: rank %s%lld did not arrive within %lld ms", "", word - 1,
                (long long)(p->timeout_ns / 1000000ull));
}
