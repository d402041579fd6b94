// Normalised bipartite adjacency as canonical CSR, built on the device (sm_100a).
//
// Replaces utility/utility_data/data_graph.py:7-55 (scipy dok/lil slicing: 89 s on the yelp2018
// shape) and the user_item_net construction of data_loader.py:42-45.  Integer/HBM-bound work:
// key generation -> radix sort -> run-length merge of duplicates -> row pointers by binary
// search -> degrees.  The sort and run-length primitives come from CUB (part of the CUDA
// toolkit); they are one-off graph preparation, not the per-step hot loop.
#include <cub/cub.cuh>

#include "idg_common.cuh"

namespace idg {

__global__ void csr_keys_kernel(const int64_t* __restrict__ user, const int64_t* __restrict__ item, int64_t E, int64_t U, int64_t N,
                                int add_self, uint64_t* __restrict__ keys) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < E) {
        const int64_t u = user[t], c = U + item[t];
        keys[2 * t] = (uint64_t)(u * N + c);
        keys[2 * t + 1] = (uint64_t)(c * N + u);
    } else if (add_self && t < E + N) {
        const int64_t r = t - E;
        keys[2 * E + r] = (uint64_t)(r * N + r);
    }
}

__global__ void csr_unpack_kernel(const uint64_t* __restrict__ ukeys, const int* __restrict__ counts, const int* __restrict__ n_runs,
                                  int64_t N, int32_t* __restrict__ indices, float* __restrict__ mult) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *n_runs) return;
    indices[t] = (int32_t)(ukeys[t] % (uint64_t)N);
    mult[t] = (float)counts[t];
}

// indptr[r] = lower_bound(ukeys, r*N); deg[r] = sum of multiplicities in the row
__global__ void csr_rows_kernel(const uint64_t* __restrict__ ukeys, const int* __restrict__ counts, const int* __restrict__ n_runs,
                                int64_t N, int32_t* __restrict__ indptr, double* __restrict__ deg) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > N) return;
    const int n = *n_runs;
    auto lb = [&](uint64_t key) { int lo = 0, hi = n; while (lo < hi) { const int mid = (lo + hi) >> 1; if (ukeys[mid] < key) lo = mid + 1; else hi = mid; } return lo; };
    const int s = lb((uint64_t)r * (uint64_t)N);
    indptr[r] = s;
    if (r < N && deg) {
        const int e = lb((uint64_t)(r + 1) * (uint64_t)N);
        double a = 0.0;
        for (int k = s; k < e; ++k) a += (double)counts[k];
        deg[r] = a;
    }
}

__global__ void csr_norm_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ mult,
                                int n_rows, const float* __restrict__ d32, const double* __restrict__ d64, float* __restrict__ data) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    const int s = indptr[r], e = indptr[r + 1];
    if (d32) {
        const float dr = d32[r];
        // (d[row] * a) * d[col], every product rounded to fp32 (no FMA contraction)
        for (int k = s + lane; k < e; k += 32) data[k] = __fmul_rn(__fmul_rn(dr, mult[k]), d32[indices[k]]);
    } else {
        const double dr = d64[r];
        for (int k = s + lane; k < e; k += 32) data[k] = (float)__dmul_rn(__dmul_rn(dr, (double)mult[k]), d64[indices[k]]);
    }
}

}  // namespace idg

using namespace idg;

extern "C" int idg_csr_structure(const int64_t* d_user, const int64_t* d_item, int64_t E, int32_t U, int32_t I, int add_self,
                                 int32_t* d_indptr, int32_t* d_indices, float* d_mult, double* d_deg, int64_t* h_nnz, void* stream_) {
    if (!d_user || !d_item || !d_indptr || !d_indices || !d_mult || !h_nnz) return fail(-1, "idg_csr_structure: null argument%s");
    if (E < 0 || U <= 0 || I <= 0) return fail(-1, "idg_csr_structure: bad sizes%s");
    const int64_t N = (int64_t)U + I;
    const int64_t M = 2 * E + (add_self ? N : 0);
    if (M >= (1ll << 31)) return fail(-1, "idg_csr_structure: %s%lld entries exceed int32 indexing", "", M);
    cudaStream_t stream = (cudaStream_t)stream_;
    uint64_t *keys = nullptr, *keys2 = nullptr, *ukeys = nullptr;
    int *counts = nullptr, *n_runs = nullptr;
    void* tmp = nullptr;
    int rc = 0;
    auto cleanup = [&]() { cudaFree(keys); cudaFree(keys2); cudaFree(ukeys); cudaFree(counts); cudaFree(n_runs); cudaFree(tmp); };
#define C_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { rc = cuda_fail(_e, #expr); cleanup(); return rc; } } while (0)
    const size_t Mz = (size_t)std::max<int64_t>(M, 1);
    C_CUDA(cudaMalloc(&keys, sizeof(uint64_t) * Mz));
    C_CUDA(cudaMalloc(&keys2, sizeof(uint64_t) * Mz));
    C_CUDA(cudaMalloc(&ukeys, sizeof(uint64_t) * Mz));
    C_CUDA(cudaMalloc(&counts, sizeof(int) * Mz));
    C_CUDA(cudaMalloc(&n_runs, sizeof(int)));
    C_CUDA(cudaMemsetAsync(n_runs, 0, sizeof(int), stream));
    int nrun_h = 0;
    if (M > 0) {
        const int64_t threads = E + (add_self ? N : 0);
        csr_keys_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(d_user, d_item, E, U, N, add_self, keys);
        g_launches.fetch_add(1);
        C_CUDA(cudaGetLastError());
        int end_bit = 1;
        while (end_bit < 64 && ((uint64_t)N * (uint64_t)N >> end_bit)) ++end_bit;
        size_t tb1 = 0, tb2 = 0;
        C_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb1, keys, keys2, (int)M, 0, end_bit, stream));
        C_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, tb2, keys2, ukeys, counts, n_runs, (int)M, stream));
        C_CUDA(cudaMalloc(&tmp, std::max(tb1, tb2)));
        C_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tb1, keys, keys2, (int)M, 0, end_bit, stream));
        C_CUDA(cub::DeviceRunLengthEncode::Encode(tmp, tb2, keys2, ukeys, counts, n_runs, (int)M, stream));
        g_launches.fetch_add(8);
        csr_unpack_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(ukeys, counts, n_runs, N, d_indices, d_mult);
        g_launches.fetch_add(1);
        C_CUDA(cudaGetLastError());
    }
    csr_rows_kernel<<<(unsigned)((N + 1 + 255) / 256), 256, 0, stream>>>(ukeys, counts, n_runs, N, d_indptr, d_deg);
    g_launches.fetch_add(1);
    C_CUDA(cudaGetLastError());
    C_CUDA(cudaMemcpyAsync(&nrun_h, n_runs, sizeof(int), cudaMemcpyDeviceToHost, stream));
    C_CUDA(cudaStreamSynchronize(stream));
#undef C_CUDA
    cleanup();
    *h_nnz = nrun_h;
    return 0;
}

extern "C" int idg_csr_normalise(const int32_t* d_indptr, const int32_t* d_indices, const float* d_mult, int32_t n_rows, int64_t nnz,
                                 const float* d_dinv32, const double* d_dinv64, float* d_data, void* stream) {
    if (!d_indptr || !d_indices || !d_mult || !d_data || n_rows < 0) return fail(-1, "idg_csr_normalise: bad argument%s");
    if ((d_dinv32 != nullptr) == (d_dinv64 != nullptr)) return fail(-1, "idg_csr_normalise: pass exactly one of dinv32 / dinv64%s");
    if (n_rows == 0 || nnz == 0) return 0;
    csr_norm_kernel<<<(n_rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(d_indptr, d_indices, d_mult, n_rows, d_dinv32, d_dinv64, d_data);
    IDG_LAUNCH_CHECK("csr_norm_kernel");
    return 0;
}
