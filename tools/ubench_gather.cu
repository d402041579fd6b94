// Micro-benchmark behind DESIGN section 4 ("where can the gathers of one propagation layer come from?"):
// random 256-byte row gathers (one 128-bit load per lane, 16 lanes per row -- the access pattern of
// spmm_kernel<16>) served from (a) an L2-resident table, (b) the shared memory of the CTAs of a thread-block
// cluster (DSMEM), (c) the CTA's own shared memory, (d) a mix of L2 and DSMEM.  Prints GB/s per variant.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_gather tools/ubench_gather.cu
//   ./gpurun_out/ubench_gather
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ float4 f4fma(float w, float4 x, float4 a) {
    return make_float4(fmaf(w, x.x, a.x), fmaf(w, x.y, a.y), fmaf(w, x.z, a.z), fmaf(w, x.w, a.w));
}

__global__ void fill_idx(int* idx, long long n, unsigned range, unsigned seed) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = (int)(hash32((unsigned)i * 2654435761U + seed) % range);
}

// (a) L2 gather: one half-warp per segment of `per` indices, U independent gathers in flight per lane
template <int U>
__global__ void __launch_bounds__(256) l2_gather(const float* __restrict__ X, const int* __restrict__ idx, int per, int n_seg, float* __restrict__ out) {
    const int lane = threadIdx.x & 31, sub = lane & 15;
    const int seg = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 2 + (lane >> 4);
    if (seg >= n_seg) return;
    const int* ip = idx + (size_t)seg * per;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int k = 0; k + U <= per; k += U) {
        int c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = __ldg(ip + k + u);
        float4 x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = __ldg(reinterpret_cast<const float4*>(X + (size_t)c[u] * 64 + sub * 4));
#pragma unroll
        for (int u = 0; u < U; ++u) acc = f4fma(1.0f, x[u], acc);
    }
    *reinterpret_cast<float4*>(out + (size_t)seg * 64 + sub * 4) = acc;
}

// sequential L2-resident read: every CTA streams the whole table `passes` times (rotated start)
__global__ void __launch_bounds__(256) l2_seq(const float4* __restrict__ X, long long n4, int passes, float4* __restrict__ out) {
    float4 acc = make_float4(0, 0, 0, 0);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; ++p) {
        long long i = ((long long)((blockIdx.x + p * 37) % gridDim.x)) * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n4; i += 4 * stride) {
            float4 a = __ldg(X + i), b = __ldg(X + i + stride), c = __ldg(X + i + 2 * stride), d = __ldg(X + i + 3 * stride);
            acc = f4fma(1.f, a, acc); acc = f4fma(1.f, b, acc); acc = f4fma(1.f, c, acc); acc = f4fma(1.f, d, acc);
        }
    }
    if (acc.x == 123.456f) out[0] = acc;
}

// (b,c,d) cluster kernel: every CTA keeps `rows_per_cta` rows of 256 B in shared memory; each half-warp performs `per`
// gathers.  mode 0: all from the cluster's shared memory (uniform over ranks, local included); mode 1: own shared
// memory only; mode 2: remote ranks only; mode 3+f: f/8 of the gathers from DSMEM (uniform rank), the rest from the L2 table.
template <int U>
__global__ void __launch_bounds__(1024, 1) dsmem_gather(const float* __restrict__ X, unsigned x_rows, int rows_per_cta, int per, int mode, float* __restrict__ out) {
    extern __shared__ __align__(16) float srows[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned C = cluster.num_blocks(), my = cluster.block_rank();
    for (int i = threadIdx.x; i < rows_per_cta * 64; i += blockDim.x) srows[i] = (float)(i & 7);
    cluster.sync();
    const int lane = threadIdx.x & 31, sub = lane & 15;
    const unsigned hw = (blockIdx.x * 32 + (threadIdx.x >> 5)) * 2 + (lane >> 4);
    float4 acc = make_float4(0, 0, 0, 0);
    const int frac = mode >= 3 ? mode - 3 : 8;
    for (int k = 0; k + U <= per; k += U) {
        float4 x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned h = hash32(hw * 65599U + (unsigned)(k + u));
            const unsigned slot = (h >> 8) % (unsigned)rows_per_cta;
            unsigned rank = (h >> 4) % C;
            if (mode == 1) rank = my;
            if (mode == 2 && C > 1) rank = (my + 1 + (h >> 4) % (C - 1)) % C;
            const bool from_l2 = (mode >= 3) && ((h & 7u) >= (unsigned)frac);
            if (from_l2) {
                x[u] = __ldg(reinterpret_cast<const float4*>(X + (size_t)((h >> 3) % x_rows) * 64 + sub * 4));
            } else {
                const float* p = cluster.map_shared_rank(srows + slot * 64 + sub * 4, rank);
                x[u] = *reinterpret_cast<const float4*>(p);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc = f4fma(1.0f, x[u], acc);
    }
    if (acc.x == 123.456f) *reinterpret_cast<float4*>(out + (size_t)hw * 64 + sub * 4) = acc;
    cluster.sync();
}

template <typename F>
static float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

template <int U>
static void run_cluster(const float* X, unsigned x_rows, float* out, int C, int rows_per_cta, int per, int mode, const char* tag) {
    auto kern = dsmem_gather<U>;
    const size_t smem = (size_t)rows_per_cta * 256;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (C > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = smem;
    cfg.gridDim = dim3(C);
    int max_clusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
    if (e != cudaSuccess || max_clusters == 0) { printf("%-34s C=%2d: not launchable (%s)\n", tag, C, cudaGetErrorString(e)); cudaGetLastError(); return; }
    const int grid = max_clusters * C;
    cfg.gridDim = dim3(grid);
    float ms = time_ms([&] { CK(cudaLaunchKernelEx(&cfg, kern, X, x_rows, rows_per_cta, per, mode, out)); });
    const double bytes = (double)grid * 64 * per * 256.0;
    printf("%-34s C=%2d U=%d ctas=%3d (%d clusters) rows/cta=%d: %.3f ms  %.0f GB/s  (%.1f B/clk/SM @1.9GHz)\n", tag, C, U, grid, max_clusters,
           rows_per_cta, ms, bytes / ms / 1e6, bytes / ms / 1e6 / grid / 1.9);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("%s, %d SMs, L2 %d MB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20);
    const unsigned N = 144242;       // amazon-book shape: 36.9 MB table
    const long long M = 4761460;     // gathers per layer
    float *X, *out; int* idx;
    CK(cudaMalloc(&X, (size_t)N * 256)); CK(cudaMemset(X, 0, (size_t)N * 256));
    CK(cudaMalloc(&out, (size_t)1 << 28));
    CK(cudaMalloc(&idx, sizeof(int) * (M + 64)));
    fill_idx<<<(unsigned)((M + 255) / 256), 256>>>(idx, M, N, 1u);
    CK(cudaDeviceSynchronize());

    const int per = 32;
    const int n_seg = (int)(M / per);
    const unsigned grid = (n_seg + 15) / 16;
    const double gbytes = (double)n_seg * per * 256.0;
    { float ms = time_ms([&] { l2_gather<2><<<grid, 256>>>(X, idx, per, n_seg, out); }); printf("L2 gather (37 MB table) U=2: %.3f ms  %.0f GB/s\n", ms, gbytes / ms / 1e6); }
    { float ms = time_ms([&] { l2_gather<4><<<grid, 256>>>(X, idx, per, n_seg, out); }); printf("L2 gather (37 MB table) U=4: %.3f ms  %.0f GB/s\n", ms, gbytes / ms / 1e6); }
    { float ms = time_ms([&] { l2_gather<8><<<grid, 256>>>(X, idx, per, n_seg, out); }); printf("L2 gather (37 MB table) U=8: %.3f ms  %.0f GB/s\n", ms, gbytes / ms / 1e6); }
    {
        const long long n4 = (long long)N * 16;
        const int passes = 8;
        float ms = time_ms([&] { l2_seq<<<148 * 8, 256>>>(reinterpret_cast<const float4*>(X), n4, passes, reinterpret_cast<float4*>(out)); });
        printf("L2 sequential read (37 MB x %d passes): %.3f ms  %.0f GB/s\n", passes, ms, (double)N * 256.0 * passes / ms / 1e6);
    }
    // smaller table (hot set only, 3.3 MB): is the L2 rate set by the fabric or by slice locality?
    {
        fill_idx<<<(unsigned)((M + 255) / 256), 256>>>(idx, M, 12800, 7u);
        float ms = time_ms([&] { l2_gather<2><<<grid, 256>>>(X, idx, per, n_seg, out); });
        printf("L2 gather (3.3 MB hot set) U=2: %.3f ms  %.0f GB/s\n", ms, gbytes / ms / 1e6);
        fill_idx<<<(unsigned)((M + 255) / 256), 256>>>(idx, M, N, 1u);
    }

    const int P = 1024;
    for (int C : {1, 2, 4, 8, 16}) {
        run_cluster<2>(X, N, out, C, 768, P, 0, "cluster smem gather (uniform rank)");
        run_cluster<4>(X, N, out, C, 768, P, 0, "cluster smem gather (uniform rank)");
    }
    run_cluster<4>(X, N, out, 1, 768, P, 1, "own smem only");
    for (int C : {8, 16}) {
        run_cluster<4>(X, N, out, C, 768, P, 2, "remote smem only");
        run_cluster<8>(X, N, out, C, 768, P, 2, "remote smem only");
    }
    for (int C : {8, 16})
        for (int f : {0, 2, 3, 4}) {
            char tag[64]; snprintf(tag, sizeof(tag), "mixed: %d/8 DSMEM, rest L2", f);
            run_cluster<4>(X, N, out, C, 768, P, 3 + f, tag);
        }
    return 0;
}
