// Propagation layer: Y = A_hat . X with fused epilogues (sm_100a).
//
// Replaces torch.sparse.mm(self.Graph, all_embedding) and the element-wise ops that
// follow it in models/LightGCN.py:43-48, SimGCL.py:47-56, XSimGCL.py:50-62 and their
// autograd transposes (trainer.py:55).  A_hat is symmetric, so backward uses the
// same CSR with the Horner addend epilogue.
//
// Design (HBM/L2-bound gather, no tensor cores):
//   * CSR rows are cut into work items of at most kChunk nonzeros; items are sorted
//     longest-first (degree-sorted schedule) so the power-law tail starts early.
//   * one warp per item.  (col,val) pairs are stored interleaved (int2) and read with
//     one coalesced 8-byte load per lane, then broadcast by shuffle.
//   * an embedding row of d = 4*LPR floats is covered by LPR lanes with one 128-bit
//     load each, so a warp-wide LDG.128 gathers 32/LPR rows (2 rows at d=64); kUnroll
//     independent gathers per lane are in flight before the FMAs.
//   * the reduction order inside a row is a pure function of the row's nonzeros
//     (lane-group partial sums over strided nonzeros, butterfly combine, chunk
//     partials summed in chunk order by the last-arriving warp): deterministic,
//     no float atomics, and independent of how rows are partitioned across GPUs.
#include <algorithm>
#include <vector>

#include "idg_common.cuh"

namespace idg {

constexpr int kChunk = 256;    // max nonzeros per work item
constexpr int kUnroll = 4;     // independent 128-bit gathers in flight per lane
constexpr int kWarpsPerCta = 8;

struct HeavyRow {
    int row;         // local row
    int part_begin;  // first slot in partials
    int n_parts;
    int pad;
};

struct SpmmArgs {
    const int4* items;  // {row | heavy index, start, end, part (-1 = whole row)}
    int n_items;
    const int2* colval;
    const HeavyRow* heavy;
    float* partials;  // [n_parts_total, d]
    int* counters;    // [n_heavy], zero between launches
    int row_offset;
    const float* X;
    float* Y;
    const float* addend;
    const float* addend2;  // y += scale2 * addend2[row]  (XSimGCL contrast-layer gradient)
    float scale2;
    const float* noise;
    float eps;
    const float* acc_in;
    float* acc_out;
    float acc_div;
};

}  // namespace idg

struct idg_graph {
    int32_t n_rows = 0, n_cols = 0, row_offset = 0;
    int64_t nnz = 0;
    int n_items = 0, n_heavy = 0, n_parts = 0;
    int2* colval = nullptr;
    int4* items = nullptr;
    idg::HeavyRow* heavy = nullptr;
    float* partials = nullptr;
    int* counters = nullptr;
};

namespace idg {

template <int LPR>
__device__ __forceinline__ void finish_row(const SpmmArgs& a, int grow, int sub, bool writer, float4 y) {
    constexpr int d = 4 * LPR;
    const size_t off = (size_t)grow * d + sub * 4;
    if (a.addend) y = f4add(y, ldcs4(a.addend + off));
    if (a.addend2) y = f4fma(a.scale2, ldcs4(a.addend2 + off), y);
    if (a.noise) {
        // x += sign(x) * F.normalize(noise, dim=-1) * eps   (SimGCL.py:49-51)
        float4 nz = ldcs4(a.noise + off);
        float ss = nz.x * nz.x + nz.y * nz.y + nz.z * nz.z + nz.w * nz.w;
#pragma unroll
        for (int m = LPR / 2; m >= 1; m >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, m);
        const float nrm = fmaxf(sqrtf(ss), 1e-12f);
        y.x += (sgn(y.x) * (nz.x / nrm)) * a.eps;
        y.y += (sgn(y.y) * (nz.y / nrm)) * a.eps;
        y.z += (sgn(y.z) * (nz.z / nrm)) * a.eps;
        y.w += (sgn(y.w) * (nz.w / nrm)) * a.eps;
    }
    if (!writer) return;
    if (a.Y) st4(a.Y + off, y);
    if (a.acc_out) {
        float4 s = y;
        if (a.acc_in) s = f4add(ldcs4(a.acc_in + off), y);
        s.x /= a.acc_div; s.y /= a.acc_div; s.z /= a.acc_div; s.w /= a.acc_div;
        stcs4(a.acc_out + off, s);
    }
}

template <int LPR>
__global__ void __launch_bounds__(kWarpsPerCta * 32) spmm_kernel(const SpmmArgs a) {
    constexpr int d = 4 * LPR;
    constexpr int G = 32 / LPR;  // rows gathered per warp-wide load
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (item >= a.n_items) return;
    const int4 it = __ldg(a.items + item);
    const int group = lane / LPR, sub = lane % LPR;
    const float* __restrict__ X = a.X;

    float4 acc = f4zero();
    for (int base = it.y; base < it.z; base += 32) {
        const int n = min(32, it.z - base);
        int2 cv = make_int2(0, 0);
        if (lane < n) cv = __ldg(a.colval + base + lane);
        for (int j = 0; j < n; j += G * kUnroll) {
            float4 x[kUnroll];
            float w[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int k = j + u * G + group;
                const int c = __shfl_sync(0xffffffffu, cv.x, k & 31);
                w[u] = __int_as_float(__shfl_sync(0xffffffffu, cv.y, k & 31));
                x[u] = f4zero();
                if (k < n) x[u] = ldg4(X + (size_t)c * d + sub * 4);
                else w[u] = 0.f;
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) acc = f4fma(w[u], x[u], acc);
        }
    }
    // combine the G lane-group partial sums (butterfly: every lane ends with the same bits)
#pragma unroll
    for (int m = LPR; m < 32; m <<= 1) acc = f4add(acc, f4shfl_xor(acc, m));

    if (it.w < 0) {
        finish_row<LPR>(a, a.row_offset + it.x, sub, group == 0, acc);
        return;
    }
    // chunk of a heavy row: publish the partial, the last chunk to arrive reduces in chunk order
    const HeavyRow h = a.heavy[it.x];
    if (group == 0) stcg4(a.partials + (size_t)it.w * d + sub * 4, acc);
    __threadfence();
    __syncwarp();
    int old = 0;
    if (lane == 0) old = atomicAdd(a.counters + it.x, 1);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old != h.n_parts - 1) return;
    __threadfence();
    float4 s = ldcg4(a.partials + (size_t)h.part_begin * d + sub * 4);
    for (int p = 1; p < h.n_parts; ++p) s = f4add(s, ldcg4(a.partials + (size_t)(h.part_begin + p) * d + sub * 4));
    if (lane == 0) a.counters[it.x] = 0;  // ready for the next launch
    finish_row<LPR>(a, a.row_offset + h.row, sub, group == 0, s);
}

__global__ void interleave_kernel(const int32_t* __restrict__ col, const float* __restrict__ val, int2* __restrict__ out, int64_t nnz) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) out[i] = make_int2(col[i], __float_as_int(val[i]));
}

}  // namespace idg

using namespace idg;

extern "C" int idg_graph_create(const int32_t* d_indptr, const int32_t* d_indices, const float* d_data, int32_t n_rows,
                                int32_t n_cols, int64_t nnz, int32_t row_offset, idg_graph** out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!out || !d_indptr || n_rows < 0 || n_cols <= 0 || nnz < 0) return fail(-1, "idg_graph_create: bad argument%s");
    if (nnz > 0 && (!d_indices || !d_data)) return fail(-1, "idg_graph_create: null indices/data%s");
    if (nnz >= (1ll << 31)) return fail(-1, "idg_graph_create: nnz %s exceeds int32 (%lld)", "", nnz);
    std::vector<int32_t> ptr((size_t)n_rows + 1);
    IDG_CUDA(cudaMemcpyAsync(ptr.data(), d_indptr, sizeof(int32_t) * ((size_t)n_rows + 1), cudaMemcpyDeviceToHost, stream));
    IDG_CUDA(cudaStreamSynchronize(stream));
    if (ptr[0] != 0 || ptr[n_rows] != nnz) return fail(-1, "idg_graph_create: indptr[0]!=0 or indptr[n]!=nnz (%s%lld vs %lld)", "", ptr[n_rows], nnz);

    idg_graph* g = new idg_graph();
    g->n_rows = n_rows; g->n_cols = n_cols; g->row_offset = row_offset; g->nnz = nnz;

    std::vector<int4> light, parts;
    std::vector<HeavyRow> heavy;
    light.reserve(n_rows);
    for (int r = 0; r < n_rows; ++r) {
        const int s = ptr[r], e = ptr[r + 1];
        if (e < s) { delete g; return fail(-1, "idg_graph_create: indptr not monotone%s"); }
        if (e - s <= kChunk) {
            light.push_back(make_int4(r, s, e, -1));
        } else {
            HeavyRow h{r, (int)parts.size(), (e - s + kChunk - 1) / kChunk, 0};
            for (int p = 0; p < h.n_parts; ++p)
                parts.push_back(make_int4((int)heavy.size(), s + p * kChunk, std::min(e, s + (p + 1) * kChunk), h.part_begin + p));
            heavy.push_back(h);
        }
    }
    // degree-sorted schedule: full chunks first, then whole rows longest-first (stable => deterministic)
    std::stable_sort(parts.begin(), parts.end(), [](const int4& x, const int4& y) { return (x.z - x.y) > (y.z - y.y); });
    std::stable_sort(light.begin(), light.end(), [](const int4& x, const int4& y) { return (x.z - x.y) > (y.z - y.y); });
    std::vector<int4> items(parts);
    items.insert(items.end(), light.begin(), light.end());
    g->n_items = (int)items.size(); g->n_heavy = (int)heavy.size(); g->n_parts = (int)parts.size();

#define G_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { idg_graph_destroy(g); return cuda_fail(_e, #expr); } } while (0)
    G_CUDA(cudaMalloc(&g->colval, sizeof(int2) * (size_t)std::max<int64_t>(nnz, 1)));
    G_CUDA(cudaMalloc(&g->items, sizeof(int4) * (size_t)std::max(g->n_items, 1)));
    G_CUDA(cudaMalloc(&g->heavy, sizeof(HeavyRow) * (size_t)std::max(g->n_heavy, 1)));
    G_CUDA(cudaMalloc(&g->partials, sizeof(float) * 128 * (size_t)std::max(g->n_parts, 1)));
    G_CUDA(cudaMalloc(&g->counters, sizeof(int) * (size_t)std::max(g->n_heavy, 1)));
    G_CUDA(cudaMemsetAsync(g->counters, 0, sizeof(int) * (size_t)std::max(g->n_heavy, 1), stream));
    if (g->n_items) G_CUDA(cudaMemcpyAsync(g->items, items.data(), sizeof(int4) * items.size(), cudaMemcpyHostToDevice, stream));
    if (g->n_heavy) G_CUDA(cudaMemcpyAsync(g->heavy, heavy.data(), sizeof(HeavyRow) * heavy.size(), cudaMemcpyHostToDevice, stream));
    if (nnz) {
        interleave_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, stream>>>(d_indices, d_data, g->colval, nnz);
        g_launches.fetch_add(1);
        G_CUDA(cudaGetLastError());
    }
    G_CUDA(cudaStreamSynchronize(stream));
#undef G_CUDA
    *out = g;
    return 0;
}

extern "C" void idg_graph_destroy(idg_graph* g) {
    if (!g) return;
    cudaFree(g->colval); cudaFree(g->items); cudaFree(g->heavy); cudaFree(g->partials); cudaFree(g->counters);
    delete g;
}
extern "C" int64_t idg_graph_nnz(const idg_graph* g) { return g ? g->nnz : -1; }
extern "C" int32_t idg_graph_rows(const idg_graph* g) { return g ? g->n_rows : -1; }

static int spmm_launch(const idg_graph* g, const float* d_X, float* d_Y, const float* d_addend, const float* d_addend2,
                       float scale2, const float* d_noise, float eps, const float* d_acc_in, float* d_acc_out, float acc_div,
                       int32_t d, void* stream_) {
    if (!g || !d_X) return fail(-1, "idg_spmm_layer: null graph or X%s");
    if (!d_Y && !d_acc_out) return fail(-1, "idg_spmm_layer: no output requested%s");
    if (d_X == d_Y || d_X == d_acc_out) return fail(-1, "idg_spmm_layer: X must not alias an output%s");
    if (d != 32 && d != 64 && d != 128) return fail(-1, "idg_spmm_layer: d must be 32, 64 or 128 (%s%lld)", "", d);
    if (d_acc_out && !(acc_div > 0.f)) return fail(-1, "idg_spmm_layer: acc_div must be > 0%s");
    if (g->n_items == 0) return 0;
    SpmmArgs a;
    a.items = g->items; a.n_items = g->n_items; a.colval = g->colval; a.heavy = g->heavy;
    a.partials = g->partials; a.counters = g->counters; a.row_offset = g->row_offset;
    a.X = d_X; a.Y = d_Y; a.addend = d_addend; a.addend2 = d_addend2; a.scale2 = scale2; a.noise = d_noise; a.eps = eps;
    a.acc_in = d_acc_in; a.acc_out = d_acc_out; a.acc_div = acc_div;
    const unsigned grid = (unsigned)((g->n_items + kWarpsPerCta - 1) / kWarpsPerCta);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (d == 64) spmm_kernel<16><<<grid, kWarpsPerCta * 32, 0, stream>>>(a);
    else if (d == 32) spmm_kernel<8><<<grid, kWarpsPerCta * 32, 0, stream>>>(a);
    else spmm_kernel<32><<<grid, kWarpsPerCta * 32, 0, stream>>>(a);
    IDG_LAUNCH_CHECK("spmm_kernel");
    return 0;
}

extern "C" int idg_spmm_layer(const idg_graph* g, const float* d_X, float* d_Y, const float* d_addend, const float* d_noise,
                              float eps, const float* d_acc_in, float* d_acc_out, float acc_div, int32_t d, void* stream) {
    return spmm_launch(g, d_X, d_Y, d_addend, nullptr, 0.f, d_noise, eps, d_acc_in, d_acc_out, acc_div, d, stream);
}

// K-layer forward, single GPU (models/LightGCN.py:36-52, SimGCL.py:39-60, XSimGCL.py:40-67).
extern "C" int idg_propagate_fwd(const idg_graph* g, const float* d_X0, int32_t d, int32_t K, int include_layer0,
                                 const float* d_noise, float eps, int32_t cl_layer, float* d_out_mean, float* d_out_cl,
                                 float* d_work, void* stream) {
    if (!g || !d_X0 || !d_out_mean || !d_work) return fail(-1, "idg_propagate_fwd: null argument%s");
    if (K < 1) return fail(-1, "idg_propagate_fwd: K must be >= 1%s");
    if (g->row_offset != 0 || g->n_rows != g->n_cols) return fail(-1, "idg_propagate_fwd: needs the whole square graph (use idg_spmm_layer per rank)%s");
    if (cl_layer > K || (cl_layer > 0 && !d_out_cl)) return fail(-1, "idg_propagate_fwd: bad cl_layer%s");
    const size_t nd = (size_t)g->n_rows * d;
    float* buf[2] = {d_work, d_work + nd};
    const float cnt = (float)(K + (include_layer0 ? 1 : 0));
    const float* x = d_X0;
    for (int l = 1; l <= K; ++l) {
        const bool last = (l == K);
        float* y = (l == cl_layer) ? d_out_cl : buf[l & 1];
        const bool need_y = !last || l == cl_layer;
        // the running layer sum lives in d_out_mean; the last layer divides by the layer count
        const float* acc_in = (l == 1) ? (include_layer0 ? d_X0 : nullptr) : d_out_mean;
        int rc = idg_spmm_layer(g, x, need_y ? y : nullptr, nullptr, d_noise ? d_noise + (size_t)(l - 1) * nd : nullptr, eps,
                                acc_in, d_out_mean, last ? cnt : 1.0f, d, stream);
        if (rc) return rc;
        x = y;
    }
    return 0;
}

// Backward of idg_propagate_fwd w.r.t. X0 (autograd of torch.sparse.mm + stack/mean, trainer.py:55).
// With H_l = cnt * dL/dX_l:  H_K = G (+cnt*Gcl if cl==K);  H_l = G + A H_{l+1} (+cnt*Gcl if cl==l);
// gX0 = (inc0*G + A H_1)/cnt.  The sign-noise perturbation has identity gradient (SimGCL.py:51).
extern "C" int idg_propagate_bwd(const idg_graph* g, const float* d_G, const float* d_Gcl, int32_t d, int32_t K,
                                 int include_layer0, int32_t cl_layer, float* d_gX0, float* d_work, void* stream) {
    if (!g || !d_G || !d_gX0 || !d_work) return fail(-1, "idg_propagate_bwd: null argument%s");
    if (K < 1) return fail(-1, "idg_propagate_bwd: K must be >= 1%s");
    if (g->row_offset != 0 || g->n_rows != g->n_cols) return fail(-1, "idg_propagate_bwd: needs the whole square graph%s");
    if (d_Gcl && (cl_layer < 1 || cl_layer > K)) return fail(-1, "idg_propagate_bwd: bad cl_layer%s");
    const size_t nd = (size_t)g->n_rows * d;
    float* buf[2] = {d_work, d_work + nd};
    const float cnt = (float)(K + (include_layer0 ? 1 : 0));
    const float* h = d_G;
    int pb = 0, rc;
    if (d_Gcl && cl_layer == K) {
        rc = idg_axpby(buf[0], 1.f, d_G, cnt, d_Gcl, (int64_t)nd, stream);
        if (rc) return rc;
        h = buf[0]; pb = 1;
    }
    for (int s = 1; s <= K; ++s) {
        const int layer = K - s;  // index of the H being produced; 0 => gX0
        if (layer > 0) {
            float* y = buf[pb]; pb ^= 1;
            rc = spmm_launch(g, h, y, d_G, (d_Gcl && cl_layer == layer) ? d_Gcl : nullptr, cnt, nullptr, 0.f, nullptr, nullptr, 1.f, d, stream);
            if (rc) return rc;
            h = y;
        } else {
            rc = spmm_launch(g, h, nullptr, include_layer0 ? d_G : nullptr, nullptr, 0.f, nullptr, 0.f, nullptr, d_gX0, cnt, d, stream);
            if (rc) return rc;
        }
    }
    return 0;
}
