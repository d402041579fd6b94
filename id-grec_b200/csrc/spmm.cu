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
//   * one lane group per item: an embedding row of d = 4*LPR floats is covered by LPR
//     lanes with one 128-bit load each, so a warp carries 32/LPR items (2 rows at d=64)
//     of adjacent, hence similar, length.  (col,val) pairs are stored interleaved (int2)
//     and read with one broadcast 8-byte load per nonzero (L1-resident: 16 per line);
//     kUnroll independent gathers per lane are in flight before the FMAs and the next
//     (col,val) quad is fetched while they fly.
//   * the reduction order inside a row is a pure function of the row's nonzeros
//     (ascending column order inside a chunk, chunk partials summed in chunk order by
//     the last-arriving lane group): deterministic, no float atomics, and independent
//     of how rows are partitioned across GPUs.
#include <stdlib.h>
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
    const float* acc_in2;     // optional further layer-sum inputs: s = ((acc_in + acc_in2) + acc_in3) + y
    const float* acc_in3;
    float* peerY[7];          // Y is also stored at the same slab offset of every peer GPU (fused all-gather)
    int n_peers;
    float* mcY;               // ... or once through the NVSwitch multicast mapping (reaches all GPUs incl. this one)
    const int* worklist;      // optional: item ids to run (row-restricted layer), count in *d_wl_count
    const int* d_wl_count;
    // Adam fused into the last backward layer: the row gradient never goes to memory
    float* adam_p; float* adam_m; float* adam_v; const float* adam_regc; const float* adam_scalars;
    float adam_b1, adam_b2, adam_eps;
    float* adam_peer_p[7]; float* adam_mc_p;
    int skip_zero_rows;       // SPARSE kernels: rows that come out exactly zero are not stored (outputs pre-zeroed by the caller)
    const unsigned* rowmask;  // ROWMASK kernels: rows whose bit is clear are skipped (outputs untouched)
    const unsigned* bitmap;   // SPARSE kernels: only columns whose bit is set contribute (rows of X outside are zero)
    const unsigned* addend_mask;  // optional: rows whose bit is clear have an all-zero addend / addend2 (not loaded)
    // SimGCL: the first layer of the clean propagation and of both perturbed views gathers the same A_hat . E0; one launch writes
    // the clean row to Y and row + sign(row) * normalize(noise_v) * eps to view_Y[v] (SimGCL.py:47-51 for each view)
    const float* view_noise[2];
    float* view_Y[2];
};

}  // namespace idg

struct idg_graph {
    const idg_peers* peers = nullptr;
    const unsigned* closure = nullptr;  // optional: batch rows + their neighbours (idg_graph_set_closure)
    int32_t n_rows = 0, n_cols = 0, row_offset = 0;
    int64_t nnz = 0;
    int n_items = 0, n_heavy = 0, n_parts = 0, n_classes = 1;
    int2* colval = nullptr;
    int4* items = nullptr;
    int2* row_items = nullptr;  // per local row: {first item, item count}
    idg::HeavyRow* heavy = nullptr;
    float* partials = nullptr;
    int* counters = nullptr;
};

namespace idg {

__device__ __forceinline__ float4 ldg4_stream(const float* p) {
    // embedding gathers have ~no L1 reuse (ncu: 2.5% hit rate): keep L1 for the (col,val) stream
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

template <bool NA>
__device__ __forceinline__ float4 gat(const float* p) { return NA ? ldg4_stream(p) : ldg4(p); }

template <int LPR, bool ADAM>
__device__ __forceinline__ void finish_row(const SpmmArgs& a, int grow, int sub, unsigned gmask, float4 y) {
    constexpr int d = 4 * LPR;
    const size_t off = (size_t)grow * d + sub * 4;
    // backward chain: the Horner addend (the loss gradient G) is non-zero on the batch rows only -- with the batch-row
    // bitmap at hand the other rows skip the 4 d bytes of zeros (N d 4 bytes per layer at full size)
    const bool has_addend = !a.addend_mask || ((__ldg(a.addend_mask + (grow >> 5)) >> (grow & 31)) & 1u);
    if (a.addend && has_addend) y = f4add(y, ldcs4(a.addend + off));
    if (a.addend2 && has_addend) y = f4fma(a.scale2, ldcs4(a.addend2 + off), y);
    if (a.noise) {
        // x += sign(x) * F.normalize(noise, dim=-1) * eps   (SimGCL.py:49-51)
        float4 nz = ldcs4(a.noise + off);
        float ss = nz.x * nz.x + nz.y * nz.y + nz.z * nz.z + nz.w * nz.w;
#pragma unroll
        for (int m = LPR / 2; m >= 1; m >>= 1) ss += __shfl_xor_sync(gmask, ss, m);
        const float nrm = fmaxf(sqrtf(ss), 1e-12f);
        y.x += (sgn(y.x) * (nz.x / nrm)) * a.eps;
        y.y += (sgn(y.y) * (nz.y / nrm)) * a.eps;
        y.z += (sgn(y.z) * (nz.z / nrm)) * a.eps;
        y.w += (sgn(y.w) * (nz.w / nrm)) * a.eps;
    }
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        if (a.view_noise[v]) {
            const float4 nz = ldcs4(a.view_noise[v] + off);
            float ss = nz.x * nz.x + nz.y * nz.y + nz.z * nz.z + nz.w * nz.w;
#pragma unroll
            for (int m = LPR / 2; m >= 1; m >>= 1) ss += __shfl_xor_sync(gmask, ss, m);
            const float nrm = fmaxf(sqrtf(ss), 1e-12f);
            float4 yv = y;
            yv.x += (sgn(y.x) * (nz.x / nrm)) * a.eps;
            yv.y += (sgn(y.y) * (nz.y / nrm)) * a.eps;
            yv.z += (sgn(y.z) * (nz.z / nrm)) * a.eps;
            yv.w += (sgn(y.w) * (nz.w / nrm)) * a.eps;
            st4(a.view_Y[v] + off, yv);
        }
    }
    if (a.mcY) {
        multimem_st4(a.mcY + off, y);  // one NVLink store, replicated by the switch (NVLS)
    } else if (a.Y) {
        st4(a.Y + off, y);
#pragma unroll 1
        for (int p = 0; p < a.n_peers; ++p) st4(a.peerY[p] + off, y);  // NVLink peer stores, fire-and-forget
    }
    if (ADAM) {
        // g = y/acc_div + regc[row]*p ; torch.optim.Adam update (same formula as adam_dev_kernel); publish p
        float4 g = make_float4(y.x / a.acc_div, y.y / a.acc_div, y.z / a.acc_div, y.w / a.acc_div);
        float4 pv = *reinterpret_cast<const float4*>(a.adam_p + off);
        const float rc = a.adam_regc ? __ldg(a.adam_regc + grow) : 0.f;
        if (rc != 0.f) { g.x = fmaf(rc, pv.x, g.x); g.y = fmaf(rc, pv.y, g.y); g.z = fmaf(rc, pv.z, g.z); g.w = fmaf(rc, pv.w, g.w); }
        float4 mv = *reinterpret_cast<const float4*>(a.adam_m + off), vv = *reinterpret_cast<const float4*>(a.adam_v + off);
        const float step_size = __ldg(a.adam_scalars), bc2_sqrt = __ldg(a.adam_scalars + 1);
        const float b1 = a.adam_b1, b2 = a.adam_b2, eps = a.adam_eps;
        auto upd = [&](float& pp, float gg, float& mm, float& v2) {
            mm = mm + (gg - mm) * (1.f - b1);
            v2 = v2 * b2 + (1.f - b2) * gg * gg;
            const float denom = sqrtf(v2) / bc2_sqrt + eps;
            pp = pp - step_size * (mm / denom);
        };
        upd(pv.x, g.x, mv.x, vv.x); upd(pv.y, g.y, mv.y, vv.y); upd(pv.z, g.z, mv.z, vv.z); upd(pv.w, g.w, mv.w, vv.w);
        st4(a.adam_m + off, mv); st4(a.adam_v + off, vv);
        if (a.adam_mc_p) {
            multimem_st4(a.adam_mc_p + off, pv);
        } else {
            st4(a.adam_p + off, pv);
#pragma unroll 1
            for (int q = 0; q < a.n_peers; ++q) st4(a.adam_peer_p[q] + off, pv);
        }
    }
    if (a.acc_out) {
        float4 s = y;
        if (a.acc_in) {
            float4 t = ldcs4(a.acc_in + off);
            if (a.acc_in2) t = f4add(t, ldcs4(a.acc_in2 + off));
            if (a.acc_in3) t = f4add(t, ldcs4(a.acc_in3 + off));
            s = f4add(t, y);
        }
        s.x /= a.acc_div; s.y /= a.acc_div; s.z /= a.acc_div; s.w /= a.acc_div;
        stcs4(a.acc_out + off, s);
    }
}

// One lane group (LPR lanes = one 4*LPR-float row per 128-bit load) per work item; a warp carries
// 32/LPR items of adjacent (hence similar) length.  Per nonzero: one broadcast 8-byte (col,val)
// load (L1-resident: 16 nonzeros per line), one 128-bit gather per lane, four FFMA.
template <int LPR, int UNROLL = kUnroll, bool NA = false, int MINB = 1, bool SPARSE = false, bool ADAM = false, bool ROWMASK = false>
__global__ void __launch_bounds__(kWarpsPerCta * 32, MINB) spmm_kernel(const SpmmArgs a) {
    constexpr int kU = UNROLL;
    constexpr int d = 4 * LPR;
    constexpr int G = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int group = lane / LPR, sub = lane % LPR;
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (group * LPR));
    const int slot = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * G + group;
    int item = slot;
    if (a.worklist) {
        if (slot >= __ldg(a.d_wl_count)) return;
        item = __ldg(a.worklist + slot);
    } else if (slot >= a.n_items) {
        return;
    }
    const int4 it = __ldg(a.items + item);
    if (ROWMASK) {  // neighbourhood-restricted layer: only rows the next (batch-restricted) layer will read
        const int grow = a.row_offset + ((it.w < 0) ? it.x : a.heavy[it.x].row);
        if (!((__ldg(a.rowmask + (grow >> 5)) >> (grow & 31)) & 1u)) return;
    }
    const float* __restrict__ X = a.X + sub * 4;
    const int2* __restrict__ cvp = a.colval;
    if (ADAM && (sub & 7) == 0) {
        // the epilogue reads this row of p / m / v from HBM: start those lines towards L2 now, under the gather loop
        const size_t poff = (size_t)(a.row_offset + ((it.w < 0) ? it.x : a.heavy[it.x].row)) * d + sub * 4;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.adam_p + poff));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.adam_m + poff));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.adam_v + poff));
    }

    float4 acc = f4zero();
    int k = it.y;
    const int end = it.z;
    if (SPARSE) {
        // X is zero outside the rows flagged in the bitmap (first backward layer: dL/dF touches only
        // the batch rows): stream the (col,val) list, gather only flagged columns, same ascending order.
        // kSU x LPR nonzeros per step: the (col, val) loads and the bitmap tests of a step are independent, so their latencies
        // overlap (hits are rare -- the batch is a few thousand of N rows -- and most steps end after the tests)
        constexpr int kSU = 4;
        for (int base = k; base < end; base += kSU * LPR) {
            int2 c[kSU];
            bool hit[kSU];
#pragma unroll
            for (int u = 0; u < kSU; ++u) {
                const int kk = base + u * LPR + sub;
                c[u] = (kk < end) ? __ldg(cvp + kk) : make_int2(0, 0);
            }
#pragma unroll
            for (int u = 0; u < kSU; ++u) {
                const int kk = base + u * LPR + sub;
                hit[u] = (kk < end) && ((__ldg(a.bitmap + (c[u].x >> 5)) >> (c[u].x & 31)) & 1u);
            }
#pragma unroll
            for (int u = 0; u < kSU; ++u) {
                unsigned m = __ballot_sync(gmask, hit[u]);
                if (LPR < 32) m = (m >> (group * LPR)) & ((1u << (LPR & 31)) - 1u);
                while (m) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    const int col = __shfl_sync(gmask, c[u].x, group * LPR + j);
                    const float w = __int_as_float(__shfl_sync(gmask, c[u].y, group * LPR + j));
                    acc = f4fma(w, ldg4(X + (size_t)col * d), acc);
                }
            }
        }
        k = end;
    }
    if (k + kU <= end) {
        int2 cv[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) cv[u] = __ldg(cvp + k + u);
        for (; k + kU <= end; k += kU) {
            float4 x[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) x[u] = gat<NA>(X + (size_t)cv[u].x * d);
            float w[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) w[u] = __int_as_float(cv[u].y);
            if (k + 2 * kU <= end) {  // software pipeline: next (col,val) quad while the gathers fly
#pragma unroll
                for (int u = 0; u < kU; ++u) cv[u] = __ldg(cvp + k + kU + u);
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) acc = f4fma(w[u], x[u], acc);
        }
    }
    for (; k < end; ++k) {
        const int2 c = __ldg(cvp + k);
        acc = f4fma(__int_as_float(c.y), gat<NA>(X + (size_t)c.x * d), acc);
    }

    if (it.w < 0) {
        if (SPARSE && a.skip_zero_rows) {
            // first backward product: most rows have no batch neighbour and a zero addend -> nothing to publish
            const size_t off = (size_t)(a.row_offset + it.x) * d + sub * 4;
            float4 g = a.addend ? ldcs4(a.addend + off) : f4zero();
            const bool nz = (acc.x != 0.f) | (acc.y != 0.f) | (acc.z != 0.f) | (acc.w != 0.f) | (g.x != 0.f) | (g.y != 0.f) | (g.z != 0.f) | (g.w != 0.f);
            if (!__any_sync(gmask, nz)) return;
        }
        finish_row<LPR, ADAM>(a, a.row_offset + it.x, sub, gmask, acc);
        return;
    }
    // chunk of a heavy row: publish the partial; the last chunk to arrive sums them in chunk order
    const HeavyRow h = a.heavy[it.x];
    stcg4(a.partials + (size_t)it.w * d + sub * 4, acc);
    __threadfence();
    __syncwarp(gmask);
    int old = 0;
    if (sub == 0) old = atomicAdd(a.counters + it.x, 1);
    old = __shfl_sync(gmask, old, group * LPR);
    if (old != h.n_parts - 1) return;
    __threadfence();
    float4 s = ldcg4(a.partials + (size_t)h.part_begin * d + sub * 4);
    for (int p = 1; p < h.n_parts; ++p) s = f4add(s, ldcg4(a.partials + (size_t)(h.part_begin + p) * d + sub * 4));
    if (sub == 0) a.counters[it.x] = 0;  // ready for the next launch
    finish_row<LPR, ADAM>(a, a.row_offset + h.row, sub, gmask, s);
}

__global__ void interleave_kernel(const int32_t* __restrict__ col, const float* __restrict__ val, int2* __restrict__ out, int64_t nnz) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) out[i] = make_int2(col[i], __float_as_int(val[i]));
}

// first and last column of every row (rows hold ascending columns); {-1, -1} for an empty row
__global__ void row_first_last_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int n_rows, int2* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int s = indptr[r], e = indptr[r + 1];
    out[r] = (e > s) ? make_int2(indices[s], indices[e - 1]) : make_int2(-1, -1);
}
}  // namespace idg

using namespace idg;

// Schedule classes for tables that do not fit the L2 (IDG_SPMM_CLASS_SPLIT = 1 forces, 0 disables; default: gather table of
// n_cols 256-byte rows beyond kClassSplitBytes).  In the bipartite adjacency every user row gathers item rows only and every item
// row gathers user rows only, so running all user rows before all item rows halves the table the L2 has to hold at any time
// (XL shape: 512 MB -> 2 x 256 MB; ncu of the mixed schedule: 29.5 GB of DRAM traffic per layer, L2 hit rate 34 %).  A row is
// "upper" when all its columns lie above its own global index and "lower" when all lie below; the split is used only when every
// non-empty row is one or the other.  The order of the work items is scheduling only: every item's sum, the heavy rows' partial
// slots and their chunk-order reduction are unchanged, so results are bit-identical with either schedule.
static const int64_t kClassSplitBytes = 100ll << 20;

static int classify_rows(const int32_t* d_indptr, const int32_t* d_indices, int32_t n_rows, int32_t n_cols, int64_t nnz, int32_t row_offset,
                         cudaStream_t stream, std::vector<unsigned char>& cls, bool* split) {
    *split = false;
    const char* env = getenv("IDG_SPMM_CLASS_SPLIT");
    const int mode = env ? atoi(env) : -1;
    if (nnz == 0 || n_rows == 0 || mode == 0 || (mode != 1 && (int64_t)n_cols * 256 <= kClassSplitBytes)) return 0;
    int2* d_fl = nullptr;
    IDG_CUDA(cudaMalloc(&d_fl, sizeof(int2) * (size_t)n_rows));
    row_first_last_kernel<<<(n_rows + 255) / 256, 256, 0, stream>>>(d_indptr, d_indices, n_rows, d_fl);
    g_launches.fetch_add(1);
    std::vector<int2> fl((size_t)n_rows);
    cudaError_t e = cudaMemcpyAsync(fl.data(), d_fl, sizeof(int2) * (size_t)n_rows, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d_fl);
    if (e != cudaSuccess) return cuda_fail(e, "classify_rows");
    cls.assign((size_t)n_rows, 0);
    size_t n_upper = 0, n_lower = 0;
    unsigned char last = 0;
    for (int r = 0; r < n_rows; ++r) {
        const int grow = row_offset + r;
        if (fl[r].x < 0) { cls[r] = last; continue; }  // empty row: nothing to gather, stays with its neighbours
        if (fl[r].x > grow) { cls[r] = last = 0; ++n_upper; }
        else if (fl[r].y < grow) { cls[r] = last = 1; ++n_lower; }
        else return 0;  // a row reaching across its own index (self loops, general graphs): keep the single schedule
    }
    *split = n_upper > 0 && n_lower > 0;
    return 0;
}

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

    std::vector<unsigned char> cls;
    bool split = false;
    if (int rc = classify_rows(d_indptr, d_indices, n_rows, n_cols, nnz, row_offset, stream, cls, &split)) { delete g; return rc; }
    const int n_cls = split ? 2 : 1;

    // per class: chunks of heavy rows first (contiguous per row), then whole rows longest-first (stable => deterministic);
    // it.w of a chunk is its slot in the partials buffer (numbered in row order, independent of where the item sits)
    std::vector<int4> light[2], parts[2];
    std::vector<HeavyRow> heavy;
    int n_parts_total = 0;
    light[0].reserve(n_rows);
    for (int r = 0; r < n_rows; ++r) {
        const int s = ptr[r], e = ptr[r + 1];
        if (e < s) { delete g; return fail(-1, "idg_graph_create: indptr not monotone%s"); }
        const int c = split ? cls[r] : 0;
        if (e - s <= kChunk) {
            light[c].push_back(make_int4(r, s, e, -1));
        } else {
            HeavyRow h{r, n_parts_total, (e - s + kChunk - 1) / kChunk, 0};
            for (int p = 0; p < h.n_parts; ++p)
                parts[c].push_back(make_int4((int)heavy.size(), s + p * kChunk, std::min(e, s + (p + 1) * kChunk), h.part_begin + p));
            n_parts_total += h.n_parts;
            heavy.push_back(h);
        }
    }
    std::vector<int4> items;
    items.reserve(light[0].size() + light[1].size() + (size_t)n_parts_total);
    std::vector<int2> row_items((size_t)std::max(n_rows, 1));
    for (int c = 0; c < n_cls; ++c) {
        std::stable_sort(light[c].begin(), light[c].end(), [](const int4& x, const int4& y) { return (x.z - x.y) > (y.z - y.y); });
        for (size_t i = 0; i < parts[c].size(); ++i) {
            const HeavyRow& h = heavy[parts[c][i].x];
            if (parts[c][i].w == h.part_begin) row_items[h.row] = make_int2((int)items.size(), h.n_parts);  // first chunk of the row
            items.push_back(parts[c][i]);
        }
        for (size_t i = 0; i < light[c].size(); ++i) {
            row_items[light[c][i].x] = make_int2((int)items.size(), 1);
            items.push_back(light[c][i]);
        }
    }
    g->n_items = (int)items.size(); g->n_heavy = (int)heavy.size(); g->n_parts = n_parts_total; g->n_classes = n_cls;

#define G_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { idg_graph_destroy(g); return cuda_fail(_e, #expr); } } while (0)
    G_CUDA(cudaMalloc(&g->colval, sizeof(int2) * (size_t)std::max<int64_t>(nnz, 1)));
    G_CUDA(cudaMalloc(&g->items, sizeof(int4) * (size_t)std::max(g->n_items, 1)));
    G_CUDA(cudaMalloc(&g->heavy, sizeof(HeavyRow) * (size_t)std::max(g->n_heavy, 1)));
    G_CUDA(cudaMalloc(&g->row_items, sizeof(int2) * row_items.size()));
    G_CUDA(cudaMemcpyAsync(g->row_items, row_items.data(), sizeof(int2) * row_items.size(), cudaMemcpyHostToDevice, stream));
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
    cudaFree(g->colval); cudaFree(g->items); cudaFree(g->row_items); cudaFree(g->heavy); cudaFree(g->partials); cudaFree(g->counters);
    delete g;
}
extern "C" int idg_graph_set_peers(idg_graph* g, const idg_peers* p) {
    if (!g) return fail(-1, "idg_graph_set_peers: null graph%s");
    g->peers = p;
    return 0;
}
extern "C" int64_t idg_graph_nnz(const idg_graph* g) { return g ? g->nnz : -1; }
extern "C" int32_t idg_graph_rows(const idg_graph* g) { return g ? g->n_rows : -1; }
extern "C" int32_t idg_graph_classes(const idg_graph* g) { return g ? g->n_classes : -1; }

struct SpmmExtra {
    const float* acc_in2 = nullptr;
    const float* acc_in3 = nullptr;
    const int* worklist = nullptr;   // row-restricted launch
    const int* d_wl_count = nullptr;
    int max_wl = 0;
    const unsigned* bitmap = nullptr;  // sparse-input launch
    const unsigned* rowmask = nullptr; // row-masked launch
    const unsigned* addend_mask = nullptr;  // rows with a non-zero addend (others skip the addend load)
    const float* view_noise[2] = {nullptr, nullptr};  // extra perturbed copies of the output row (SimGCL's shared first layer)
    float* view_Y[2] = {nullptr, nullptr};
    int skip_zero_rows = 0;
    const idg_adam_args* adam = nullptr;  // Adam-fused epilogue (last backward layer)
};

static int spmm_launch(const idg_graph* g, const float* d_X, float* d_Y, const float* d_addend, const float* d_addend2,
                       float scale2, const float* d_noise, float eps, const float* d_acc_in, float* d_acc_out, float acc_div,
                       int32_t d, void* stream_, const SpmmExtra& ex = SpmmExtra()) {
    if (!g || !d_X) return fail(-1, "idg_spmm_layer: null graph or X%s");
    if (!d_Y && !d_acc_out && !ex.adam) return fail(-1, "idg_spmm_layer: no output requested%s");
    if (d_X == d_Y || d_X == d_acc_out) return fail(-1, "idg_spmm_layer: X must not alias an output%s");
    if (d != 32 && d != 64 && d != 128) return fail(-1, "idg_spmm_layer: d must be 32, 64 or 128 (%s%lld)", "", d);
    if (d_acc_out && !(acc_div > 0.f)) return fail(-1, "idg_spmm_layer: acc_div must be > 0%s");
    if (g->n_items == 0) return 0;
    SpmmArgs a;
    a.items = g->items; a.n_items = g->n_items; a.colval = g->colval; a.heavy = g->heavy;
    a.partials = g->partials; a.counters = g->counters; a.row_offset = g->row_offset;
    a.X = d_X; a.Y = d_Y; a.addend = d_addend; a.addend2 = d_addend2; a.scale2 = scale2; a.noise = d_noise; a.eps = eps;
    a.acc_in = d_acc_in; a.acc_out = d_acc_out; a.acc_div = acc_div;
    a.worklist = ex.worklist; a.d_wl_count = ex.d_wl_count; a.bitmap = ex.bitmap; a.skip_zero_rows = ex.skip_zero_rows;
    a.rowmask = ex.rowmask;
    a.addend_mask = ex.addend_mask;
    for (int v = 0; v < 2; ++v) { a.view_noise[v] = ex.view_noise[v]; a.view_Y[v] = ex.view_Y[v]; }
    a.acc_in2 = ex.acc_in2; a.acc_in3 = ex.acc_in3;
    a.adam_p = nullptr; a.adam_m = a.adam_v = nullptr; a.adam_regc = a.adam_scalars = nullptr; a.adam_mc_p = nullptr;
    a.adam_b1 = a.adam_b2 = a.adam_eps = 0.f;
    for (int p = 0; p < 7; ++p) a.adam_peer_p[p] = nullptr;
    a.n_peers = 0;
    a.mcY = nullptr;
    for (int p = 0; p < 7; ++p) a.peerY[p] = nullptr;
    if (g->peers && d_Y && g->peers->world > 1) {
        const idg_peers* P = g->peers;
        const char* y = (const char*)d_Y;
        if (y >= P->local_base && y < P->local_base + P->bytes) {
            if (P->mc_base) {
                a.mcY = (float*)(P->mc_base + (y - P->local_base));
            } else {
                for (int q = 0; q < P->world; ++q)
                    if (q != P->rank) a.peerY[a.n_peers++] = (float*)(P->bases[q] + (y - P->local_base));
            }
        }
    }
    if (ex.adam) {
        const idg_adam_args* A = ex.adam;
        if (!A->p || !A->m || !A->v || !A->d_scalars || !(acc_div > 0.f)) return fail(-1, "idg_spmm_layer_adam: incomplete Adam arguments%s");
        a.adam_p = A->p; a.adam_m = A->m; a.adam_v = A->v; a.adam_regc = A->regc; a.adam_scalars = A->d_scalars;
        a.adam_b1 = A->beta1; a.adam_b2 = A->beta2; a.adam_eps = A->eps; a.acc_div = acc_div;
        if (g->peers && g->peers->world > 1) {  // parameter rows go straight to the peers as well (replaces the push)
            const idg_peers* P = g->peers;
            const char* pp = (const char*)A->p;
            if (pp >= P->local_base && pp < P->local_base + P->bytes) {
                if (P->mc_base) a.adam_mc_p = (float*)(P->mc_base + (pp - P->local_base));
                else { a.n_peers = 0; for (int q = 0; q < P->world; ++q) if (q != P->rank) a.adam_peer_p[a.n_peers++] = (float*)(P->bases[q] + (pp - P->local_base)); }
            }
        }
    }
    // warps per CTA: 8 by default; IDG_SPMM_WARPS = 2 | 4 | 8 for tuning (smaller CTAs = finer tail, more CTA launches)
    static const int warps_env = getenv("IDG_SPMM_WARPS") ? atoi(getenv("IDG_SPMM_WARPS")) : 0;
    const int warps_per_cta = (warps_env == 2 || warps_env == 4) ? warps_env : kWarpsPerCta;
    const int per_cta = warps_per_cta * (32 / (d / 4));  // items per CTA: one lane group each
    const int n_slots = ex.worklist ? ex.max_wl : g->n_items;
    if (n_slots <= 0) return 0;
    const unsigned grid = (unsigned)((n_slots + per_cta - 1) / per_cta);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int T = warps_per_cta * 32;
    // tuned on B200 (amazon-book shape): 2 gathers in flight per lane at full occupancy (<= 32 registers,
    // 64 warps/SM) beats deeper unrolling at lower occupancy; L1-allocating gathers beat .L1::no_allocate.
    if (ex.adam) {
        if (ex.bitmap) return fail(-1, "idg_spmm_layer_adam: the Adam-fused layer cannot be the sparse-input one (K >= 2)%s");
        // 6 resident CTAs (40 registers): forcing 32 registers for 8 CTAs spills in the epilogue and measured 1.7 % slower per step
        // ... at the amazon-book shape.  With a gather table beyond the L2 (XL shape) the layer is bound by HBM latency and the
        // 64 resident warps of the 32-register build win (IDG_SPMM_ADAM8 = 0 | 1 overrides the size rule)
        const char* a8 = getenv("IDG_SPMM_ADAM8");
        const bool adam8 = d == 64 && (a8 ? atoi(a8) != 0 : (int64_t)g->n_cols * d * 4 > kClassSplitBytes);
        if (adam8) spmm_kernel<16, 2, false, 8, false, true><<<grid, T, 0, stream>>>(a);
        else if (d == 64) spmm_kernel<16, 2, false, 6, false, true><<<grid, T, 0, stream>>>(a);
        else if (d == 32) spmm_kernel<8, 2, false, 6, false, true><<<grid, T, 0, stream>>>(a);
        else spmm_kernel<32, 2, false, 6, false, true><<<grid, T, 0, stream>>>(a);
    } else if (ex.bitmap && ex.rowmask) {
        // sparse-input product restricted to the rows that can come out non-zero (batch rows + their neighbours): the other
        // rows would stream their whole (col, val) list only to find no flagged column
        if (d == 64) spmm_kernel<16, 2, false, 8, true, false, true><<<grid, T, 0, stream>>>(a);
        else if (d == 32) spmm_kernel<8, 2, false, 8, true, false, true><<<grid, T, 0, stream>>>(a);
        else spmm_kernel<32, 2, false, 8, true, false, true><<<grid, T, 0, stream>>>(a);
    } else if (ex.bitmap) {
        if (d == 64) spmm_kernel<16, 2, false, 8, true><<<grid, T, 0, stream>>>(a);
        else if (d == 32) spmm_kernel<8, 2, false, 8, true><<<grid, T, 0, stream>>>(a);
        else spmm_kernel<32, 2, false, 8, true><<<grid, T, 0, stream>>>(a);
    } else if (ex.rowmask) {
        if (d == 64) spmm_kernel<16, 2, false, 8, false, false, true><<<grid, T, 0, stream>>>(a);
        else if (d == 32) spmm_kernel<8, 2, false, 8, false, false, true><<<grid, T, 0, stream>>>(a);
        else spmm_kernel<32, 2, false, 8, false, false, true><<<grid, T, 0, stream>>>(a);
    } else if (ex.worklist) {
        // row-restricted launch: a few thousand rows, latency-bound per row -> deep unrolling instead of occupancy
        if (d == 64) spmm_kernel<16, 8, false, 1><<<grid, T, 0, stream>>>(a);
        else if (d == 32) spmm_kernel<8, 8, false, 1><<<grid, T, 0, stream>>>(a);
        else spmm_kernel<32, 8, false, 1><<<grid, T, 0, stream>>>(a);
    } else {
        if (d == 64) spmm_kernel<16, 2, false, 8><<<grid, T, 0, stream>>>(a);
        else if (d == 32) spmm_kernel<8, 2, false, 8><<<grid, T, 0, stream>>>(a);
        else spmm_kernel<32, 2, false, 8><<<grid, T, 0, stream>>>(a);
    }
    IDG_LAUNCH_CHECK("spmm_kernel");
    return 0;
}

extern "C" int idg_spmm_layer(const idg_graph* g, const float* d_X, float* d_Y, const float* d_addend, const float* d_noise,
                              float eps, const float* d_acc_in, float* d_acc_out, float acc_div, int32_t d, void* stream) {
    return spmm_launch(g, d_X, d_Y, d_addend, nullptr, 0.f, d_noise, eps, d_acc_in, d_acc_out, acc_div, d, stream);
}

// ---- batch row set: unique rows touched by a mini-batch, as a list and as a bitmap ----------------
namespace idg {
// one warp per entry e of the 3B keys (user, U+pos, U+neg); the first occurrence of a row appends it
__global__ void __launch_bounds__(256) batch_rows_kernel(const int64_t* __restrict__ user, const int64_t* __restrict__ pos,
                                                         const int64_t* __restrict__ neg, int B, int U, int* __restrict__ rowlist,
                                                         int* __restrict__ count, unsigned* __restrict__ bitmap,
                                                         unsigned char* __restrict__ lead) {
    extern __shared__ __align__(16) int skeys[];   // [3B] padded with -1 to a multiple of 128 (4 keys per lane and step)
    const int n = 3 * B;
    const int npad = (n + 127) & ~127;
    for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        const int role = i / B, b = i - role * B;
        skeys[i] = (i >= n) ? -1 : (role == 0 ? (int)user[b] : U + (int)(role == 1 ? pos[b] : neg[b]));
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= n) return;
    const int node = skeys[e];
    const int4* skeys4 = reinterpret_cast<const int4*>(skeys);
    for (int base = 0; base < e; base += 128) {
        const int i = base + lane * 4;
        const int4 k = skeys4[i >> 2];
        const bool hit = ((i < e) && (k.x == node)) || ((i + 1 < e) && (k.y == node)) || ((i + 2 < e) && (k.z == node)) || ((i + 3 < e) && (k.w == node));
        if (__any_sync(0xffffffffu, hit)) {
            if (lead && lane == 0) lead[e] = 0;
            return;
        }
    }
    if (lane == 0) {
        if (lead) lead[e] = 1;
        rowlist[atomicAdd(count, 1)] = node;
        atomicOr(bitmap + (node >> 5), 1u << (node & 31));
    }
}

// distinct users / positive items of the batch in order of first appearance (the job of torch.unique at
// SimGCL.py:80-81; InfoNCE is permutation-invariant, the order only has to be deterministic).
// One warp per entry e of the user and positive segments; lead[] comes from batch_rows_kernel.
__global__ void __launch_bounds__(256) batch_unique_kernel(const int64_t* __restrict__ user, const int64_t* __restrict__ pos, int B, int U,
                                                           const unsigned char* __restrict__ lead, int64_t* __restrict__ uidx,
                                                           int* __restrict__ ucnt, int64_t* __restrict__ iidx, int* __restrict__ icnt) {
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= 2 * B) return;
    const int seg = e / B, b = e - seg * B;
    const unsigned char* ld = lead + seg * B;
    int before = 0;
    for (int base = 0; base < b; base += 32) {
        const int i = base + lane;
        before += __popc(__ballot_sync(0xffffffffu, (i < b) && ld[i]));
    }
    if (lane == 0) {
        const bool mine = ld[b] != 0;
        if (mine) {
            if (seg == 0) uidx[before] = user[b]; else iidx[before] = (int64_t)U + pos[b];
        }
        if (b == B - 1) { if (seg == 0) *ucnt = before + (mine ? 1 : 0); else *icnt = before + (mine ? 1 : 0); }
    }
}

// closure[r] = 1 iff row r has a column flagged in `batch` (r is a neighbour of a batch row); one lane group per work
// item of the schedule (heavy rows are already chunked), early exit on the first hit.  The batch rows themselves are
// OR-ed in by closure_or_kernel.
template <int LPR>
__global__ void __launch_bounds__(256) closure_kernel(const int4* __restrict__ items, int n_items, const int2* __restrict__ colval,
                                                      const HeavyRow* __restrict__ heavy, int row_offset, const unsigned* __restrict__ batch,
                                                      unsigned* __restrict__ closure) {
    const int lane = threadIdx.x & 31;
    const int group = lane / LPR, sub = lane % LPR;
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (group * LPR));
    const int item = (blockIdx.x * 8 + (threadIdx.x >> 5)) * (32 / LPR) + group;
    if (item >= n_items) return;
    const int4 it = __ldg(items + item);
    bool any = false;
    for (int base = it.y; base < it.z && !any; base += LPR) {
        const int k = base + sub;
        bool hit = false;
        if (k < it.z) { const int c = __ldg(colval + k).x; hit = (__ldg(batch + (c >> 5)) >> (c & 31)) & 1u; }
        any = __any_sync(gmask, hit);
    }
    if (any && sub == 0) {
        const int grow = row_offset + ((it.w < 0) ? it.x : heavy[it.x].row);
        atomicOr(closure + (grow >> 5), 1u << (grow & 31));
    }
}
// The same closure from the batch side: A_hat has a symmetric structure, so {rows with a neighbour in the batch} is the union of the
// column lists of the batch rows -- ~3 B rows x mean degree entries instead of a pass over every nonzero of the graph (1.6 GB of
// (col, val) pairs at the 1M x 1M size: 0.8 ms per step).  One warp per listed row; needs a handle that holds the listed rows.
__global__ void __launch_bounds__(256) closure_from_rows_kernel(const int* __restrict__ rowlist, const int* __restrict__ count, const int2* __restrict__ row_items,
                                                                const int4* __restrict__ items, const int2* __restrict__ colval, int row_offset, int n_rows,
                                                                unsigned* __restrict__ closure) {
    // one CTA per listed row, its work items (256-nonzero chunks of a popular item's row) spread over the 8 warps
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x;
    if (i >= *count) return;
    const int r = rowlist[i] - row_offset;
    if (r < 0 || r >= n_rows) return;
    const int2 ri = row_items[r];
    for (int q = warp; q < ri.y; q += 8) {
        const int4 it = __ldg(items + ri.x + q);
        for (int k = it.y + lane; k < it.z; k += 32) {
            const int c = __ldg(colval + k).x;
            atomicOr(closure + (c >> 5), 1u << (c & 31));
        }
    }
}
__global__ void closure_or_kernel(const unsigned* __restrict__ batch, unsigned* __restrict__ closure, int w0, int w1) {
    const int i = w0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w1) { const unsigned b = batch[i]; if (b) atomicOr(closure + i, b); }
}

__global__ void batch_rows_clear_kernel(const int* __restrict__ rowlist, const int* __restrict__ count, unsigned* __restrict__ bitmap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *count) bitmap[rowlist[i] >> 5] = 0u;  // every set bit belongs to a listed row
}

// rows -> item ids of the schedule (heavy rows expand to all their chunks); rows outside [row_offset, +n_rows) are skipped
__global__ void expand_rows_kernel(const int* __restrict__ rowlist, const int* __restrict__ count, const int2* __restrict__ row_items,
                                   int row_offset, int n_rows, int* __restrict__ worklist, int* __restrict__ wl_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count) return;
    const int r = rowlist[i] - row_offset;
    if (r < 0 || r >= n_rows) return;
    const int2 ri = row_items[r];
    const int p = atomicAdd(wl_count, ri.y);
    for (int q = 0; q < ri.y; ++q) worklist[p + q] = ri.x + q;
}
}  // namespace idg

static int batch_rows_impl(const int64_t* d_user, const int64_t* d_pos, const int64_t* d_neg, int32_t B, int32_t U,
                           int32_t* d_rowlist, int32_t* d_count, uint32_t* d_bitmap, unsigned char* d_lead, void* stream_) {
    if (!d_user || !d_pos || !d_neg || !d_rowlist || !d_count || !d_bitmap || B <= 0) return fail(-1, "idg_batch_rows: bad argument%s");
    const size_t smem = sizeof(int) * (((3 * (size_t)B) + 127) & ~(size_t)127);
    if (smem > 200 * 1024) return fail(-1, "idg_batch_rows: batch too large (B=%s%lld)", "", B);
    cudaStream_t stream = (cudaStream_t)stream_;
    IDG_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), stream));
    if (smem > 48 * 1024) IDG_CUDA(cudaFuncSetAttribute(batch_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    batch_rows_kernel<<<(3 * B + 7) / 8, 256, smem, stream>>>(d_user, d_pos, d_neg, B, U, d_rowlist, d_count, d_bitmap, d_lead);
    IDG_LAUNCH_CHECK("batch_rows_kernel");
    return 0;
}

extern "C" int idg_batch_rows(const int64_t* d_user, const int64_t* d_pos, const int64_t* d_neg, int32_t B, int32_t U,
                              int32_t* d_rowlist, int32_t* d_count, uint32_t* d_bitmap, void* stream) {
    return batch_rows_impl(d_user, d_pos, d_neg, B, U, d_rowlist, d_count, d_bitmap, nullptr, stream);
}

// idg_batch_rows + the distinct users / (U + positive items) of the batch, in order of first appearance, with their
// counts on the device (d_lead: 3B bytes of scratch).
extern "C" int idg_batch_rows_unique(const int64_t* d_user, const int64_t* d_pos, const int64_t* d_neg, int32_t B, int32_t U,
                                     int32_t* d_rowlist, int32_t* d_count, uint32_t* d_bitmap, unsigned char* d_lead,
                                     int64_t* d_uidx, int32_t* d_ucnt, int64_t* d_iidx, int32_t* d_icnt, void* stream) {
    if (!d_lead || !d_uidx || !d_ucnt || !d_iidx || !d_icnt) return fail(-1, "idg_batch_rows_unique: null argument%s");
    if (int rc = batch_rows_impl(d_user, d_pos, d_neg, B, U, d_rowlist, d_count, d_bitmap, d_lead, stream)) return rc;
    batch_unique_kernel<<<(2 * B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(d_user, d_pos, B, U, d_lead, d_uidx, d_ucnt, d_iidx, d_icnt);
    IDG_LAUNCH_CHECK("batch_unique_kernel");
    return 0;
}

extern "C" int idg_batch_rows_clear(const int32_t* d_rowlist, const int32_t* d_count, int32_t max_rows, uint32_t* d_bitmap, void* stream) {
    if (!d_rowlist || !d_count || !d_bitmap || max_rows <= 0) return fail(-1, "idg_batch_rows_clear: bad argument%s");
    batch_rows_clear_kernel<<<(max_rows + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_rowlist, d_count, d_bitmap);
    IDG_LAUNCH_CHECK("batch_rows_clear_kernel");
    return 0;
}

// closure |= {rows of this handle with a neighbour in the batch} | {batch rows of this handle's row range}
extern "C" int idg_closure_bitmap(const idg_graph* g, const uint32_t* d_batch_bitmap, uint32_t* d_closure, void* stream_) {
    if (!g || !d_batch_bitmap || !d_closure) return fail(-1, "idg_closure_bitmap: null argument%s");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (g->n_items > 0) {
        closure_kernel<16><<<(g->n_items + 15) / 16, 256, 0, stream>>>(g->items, g->n_items, g->colval, g->heavy, g->row_offset, d_batch_bitmap, d_closure);
        IDG_LAUNCH_CHECK("closure_kernel");
    }
    const int w0 = g->row_offset >> 5, w1 = (g->row_offset + g->n_rows + 31) >> 5;
    if (w1 > w0) {
        closure_or_kernel<<<(w1 - w0 + 255) / 256, 256, 0, stream>>>(d_batch_bitmap, d_closure, w0, w1);
        IDG_LAUNCH_CHECK("closure_or_kernel");
    }
    return 0;
}

// closure |= {columns of the listed rows} | {listed rows}: identical to idg_closure_bitmap for a symmetric structure when the handle
// holds every listed row (the whole graph), at a cost proportional to the batch instead of the graph
extern "C" int idg_closure_from_rows(const idg_graph* g, const int32_t* d_rowlist, const int32_t* d_count, int32_t max_rows,
                                     const uint32_t* d_batch_bitmap, uint32_t* d_closure, void* stream_) {
    if (!g || !d_rowlist || !d_count || !d_batch_bitmap || !d_closure || max_rows <= 0) return fail(-1, "idg_closure_from_rows: bad argument%s");
    if (g->row_offset != 0 || g->n_rows != g->n_cols) return fail(-1, "idg_closure_from_rows: needs the whole (square, symmetric) graph%s");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (g->n_items > 0) {
        closure_from_rows_kernel<<<max_rows, 256, 0, stream>>>(d_rowlist, d_count, g->row_items, g->items, g->colval, g->row_offset, g->n_rows, d_closure);
        IDG_LAUNCH_CHECK("closure_from_rows_kernel");
    }
    const int w1 = (g->n_rows + 31) >> 5;
    closure_or_kernel<<<(w1 + 255) / 256, 256, 0, stream>>>(d_batch_bitmap, d_closure, 0, w1);
    IDG_LAUNCH_CHECK("closure_or_kernel");
    return 0;
}

extern "C" int idg_graph_set_closure(idg_graph* g, const uint32_t* d_closure) {
    if (!g) return fail(-1, "idg_graph_set_closure: null graph%s");
    g->closure = d_closure;
    return 0;
}

extern "C" int idg_spmm_layer_masked(const idg_graph* g, const float* d_X, float* d_Y, const float* d_noise, float eps, const float* d_acc_in,
                                     float* d_acc_out, float acc_div, int32_t d, const uint32_t* d_rowmask, void* stream) {
    if (!d_rowmask) return fail(-1, "idg_spmm_layer_masked: null row mask%s");
    SpmmExtra ex;
    ex.rowmask = d_rowmask;
    return spmm_launch(g, d_X, d_Y, nullptr, nullptr, 0.f, d_noise, eps, d_acc_in, d_acc_out, acc_div, d, stream, ex);
}

extern "C" int64_t idg_graph_worklist_ints(const idg_graph* g, int32_t max_rows) {
    return g ? (int64_t)max_rows + g->n_parts + 8 : -1;
}

// one layer restricted to the listed rows (d_worklist: idg_graph_worklist_ints(g, max_rows) ints of scratch)
static int spmm_rows(const idg_graph* g, const float* X, float* Y, const float* noise, float eps, const float* acc_in, float* acc_out,
                     float acc_div, int d, const int* d_rowlist, const int* d_count, int max_rows, int* d_worklist, void* stream_,
                     const float* acc_in2 = nullptr, const float* acc_in3 = nullptr) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int* wl_count = d_worklist;  // first int = number of work items, list follows (aligned to 4 ints)
    IDG_CUDA(cudaMemsetAsync(wl_count, 0, sizeof(int), stream));
    expand_rows_kernel<<<(max_rows + 255) / 256, 256, 0, stream>>>(d_rowlist, d_count, g->row_items, g->row_offset, g->n_rows, d_worklist + 4, wl_count);
    IDG_LAUNCH_CHECK("expand_rows_kernel");
    SpmmExtra ex;
    ex.worklist = d_worklist + 4; ex.d_wl_count = wl_count; ex.max_wl = max_rows + g->n_parts;
    ex.acc_in2 = acc_in2; ex.acc_in3 = acc_in3;
    return spmm_launch(g, X, Y, nullptr, nullptr, 0.f, noise, eps, acc_in, acc_out, acc_div, d, stream_, ex);
}

extern "C" int idg_spmm_layer_rows(const idg_graph* g, const float* d_X, float* d_Y, const float* d_noise, float eps,
                                   const float* d_acc_in, const float* d_acc_in2, const float* d_acc_in3, float* d_acc_out,
                                   float acc_div, int32_t d, const int32_t* d_rowlist, const int32_t* d_count, int32_t max_rows,
                                   int32_t* d_worklist, void* stream) {
    if (!g || !d_rowlist || !d_count || !d_worklist || max_rows <= 0) return fail(-1, "idg_spmm_layer_rows: bad argument%s");
    if ((d_acc_in2 || d_acc_in3) && !d_acc_in) return fail(-1, "idg_spmm_layer_rows: acc_in2/3 need acc_in%s");
    return spmm_rows(g, d_X, d_Y, d_noise, eps, d_acc_in, d_acc_out, acc_div, d, d_rowlist, d_count, max_rows, d_worklist, stream,
                     d_acc_in2, d_acc_in3);
}

extern "C" int idg_spmm_layer_sparse_in(const idg_graph* g, const float* d_X, float* d_Y, const float* d_addend, const float* d_acc_in,
                                        float* d_acc_out, float acc_div, int32_t d, const uint32_t* d_bitmap, int skip_zero_rows,
                                        void* stream) {
    if (!d_bitmap) return fail(-1, "idg_spmm_layer_sparse_in: null bitmap%s");
    if (skip_zero_rows && d_acc_out) return fail(-1, "idg_spmm_layer_sparse_in: skip_zero_rows only with a plain Y output%s");
    SpmmExtra ex;
    ex.bitmap = d_bitmap;
    ex.skip_zero_rows = skip_zero_rows;
    return spmm_launch(g, d_X, d_Y, d_addend, nullptr, 0.f, nullptr, 0.f, d_acc_in, d_acc_out, acc_div, d, stream, ex);
}

// General backward-chain layer on the handle's rows: y = A h + addend + scale2 * addend2 (XSimGCL: the InfoNCE gradient of the
// captured layer enters the Horner chain at that layer), optionally with the sparse-input gather (d_bitmap).
extern "C" int idg_spmm_layer_add2(const idg_graph* g, const float* d_X, float* d_Y, const float* d_addend, const float* d_addend2,
                                   float scale2, int32_t d, const uint32_t* d_bitmap, int skip_zero_rows, void* stream) {
    if (skip_zero_rows && d_addend2) return fail(-1, "idg_spmm_layer_add2: skip_zero_rows cannot be combined with a second addend%s");
    SpmmExtra ex;
    ex.bitmap = d_bitmap;
    ex.skip_zero_rows = d_bitmap ? skip_zero_rows : 0;
    return spmm_launch(g, d_X, d_Y, d_addend, d_addend2, scale2, nullptr, 0.f, nullptr, nullptr, 1.f, d, stream, ex);
}

// idg_spmm_layer_sparse_in evaluated only on the rows flagged in d_rowmask (others untouched): the row-partitioned first backward
// product, whose output is pre-zeroed and can only be non-zero on the batch neighbourhood
extern "C" int idg_spmm_layer_sparse_in_masked(const idg_graph* g, const float* d_X, float* d_Y, const float* d_addend, int32_t d,
                                               const uint32_t* d_bitmap, const uint32_t* d_rowmask, int skip_zero_rows, void* stream) {
    if (!d_bitmap || !d_rowmask) return fail(-1, "idg_spmm_layer_sparse_in_masked: null bitmap%s");
    SpmmExtra ex;
    ex.bitmap = d_bitmap;
    ex.rowmask = d_rowmask;
    ex.skip_zero_rows = skip_zero_rows;
    return spmm_launch(g, d_X, d_Y, d_addend, nullptr, 0.f, nullptr, 0.f, nullptr, nullptr, 1.f, d, stream, ex);
}

// One layer whose (clean) output row also yields up to two sign-noise perturbed copies: SimGCL's three propagations of a step share
// their first product A_hat . E0 (SimGCL.py:62-66 calls aggregate() three times on the same ego table).
extern "C" int idg_spmm_layer_views(const idg_graph* g, const float* d_X, float* d_Y, const float* d_noise_a, float* d_Y_a,
                                    const float* d_noise_b, float* d_Y_b, float eps, int32_t d, void* stream) {
    if (!d_Y) return fail(-1, "idg_spmm_layer_views: the clean output is required%s");
    if ((d_noise_a && !d_Y_a) || (d_noise_b && !d_Y_b)) return fail(-1, "idg_spmm_layer_views: a view needs its output%s");
    if (d_X == d_Y_a || d_X == d_Y_b) return fail(-1, "idg_spmm_layer_views: X must not alias an output%s");
    SpmmExtra ex;
    ex.view_noise[0] = d_noise_a; ex.view_Y[0] = d_Y_a;
    ex.view_noise[1] = d_noise_b; ex.view_Y[1] = d_Y_b;
    return spmm_launch(g, d_X, d_Y, nullptr, nullptr, 0.f, nullptr, eps, nullptr, nullptr, 1.f, d, stream, ex);
}

extern "C" int idg_spmm_layer_adam(const idg_graph* g, const float* d_X, const float* d_addend, float acc_div, int32_t d,
                                   const idg_adam_args* adam, void* stream) {
    if (!adam) return fail(-1, "idg_spmm_layer_adam: null adam%s");
    SpmmExtra ex;
    ex.adam = adam;
    return spmm_launch(g, d_X, nullptr, d_addend, nullptr, 0.f, nullptr, 0.f, nullptr, nullptr, acc_div, d, stream, ex);
}

// K-layer forward, single GPU (models/LightGCN.py:36-52, SimGCL.py:39-60, XSimGCL.py:40-67).  With a row list the
// LAST layer (and with it the layer mean) is evaluated only on those rows: the loss reads nothing else
// (LightGCN.py:57-59), so the result is identical where it is consumed and ~1/K of the gather work disappears.
extern "C" int idg_propagate_fwd_ex(const idg_graph* g, const float* d_X0, int32_t d, int32_t K, int include_layer0,
                                    const float* d_noise, float eps, int32_t cl_layer, float* d_out_mean, float* d_out_cl,
                                    float* d_work, const int32_t* d_rowlist, const int32_t* d_count, int32_t max_rows,
                                    int32_t* d_worklist, void* stream) {
    if (!g || !d_X0 || !d_out_mean || !d_work) return fail(-1, "idg_propagate_fwd: null argument%s");
    if (K < 1) return fail(-1, "idg_propagate_fwd: K must be >= 1%s");
    if (g->row_offset != 0 || g->n_rows != g->n_cols) return fail(-1, "idg_propagate_fwd: needs the whole square graph (use idg_spmm_layer per rank)%s");
    if (cl_layer > K || (cl_layer > 0 && !d_out_cl)) return fail(-1, "idg_propagate_fwd: bad cl_layer%s");
    if (d_rowlist && (!d_count || !d_worklist || max_rows <= 0)) return fail(-1, "idg_propagate_fwd: incomplete row restriction%s");
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
        const float* nz = d_noise ? d_noise + (size_t)(l - 1) * nd : nullptr;
        int rc;
        if (last && d_rowlist)
            rc = spmm_rows(g, x, need_y ? y : nullptr, nz, eps, acc_in, d_out_mean, cnt, d, d_rowlist, d_count, max_rows, d_worklist, stream);
        else if (l == K - 1 && d_rowlist && g->closure) {
            // the batch-restricted last layer only reads this layer at the batch rows and their neighbours
            SpmmExtra ex;
            ex.rowmask = g->closure;
            rc = spmm_launch(g, x, y, nullptr, nullptr, 0.f, nz, eps, acc_in, d_out_mean, 1.0f, d, stream, ex);
        } else
            rc = idg_spmm_layer(g, x, need_y ? y : nullptr, nullptr, nz, eps, acc_in, d_out_mean, last ? cnt : 1.0f, d, stream);
        if (rc) return rc;
        x = y;
    }
    return 0;
}

extern "C" int idg_propagate_fwd(const idg_graph* g, const float* d_X0, int32_t d, int32_t K, int include_layer0,
                                 const float* d_noise, float eps, int32_t cl_layer, float* d_out_mean, float* d_out_cl,
                                 float* d_work, void* stream) {
    return idg_propagate_fwd_ex(g, d_X0, d, K, include_layer0, d_noise, eps, cl_layer, d_out_mean, d_out_cl, d_work, nullptr, nullptr, 0,
                                nullptr, stream);
}

// Backward of idg_propagate_fwd w.r.t. X0 (autograd of torch.sparse.mm + stack/mean, trainer.py:55).
// With H_l = cnt * dL/dX_l:  H_K = G (+cnt*Gcl if cl==K);  H_l = G + A H_{l+1} (+cnt*Gcl if cl==l);
// gX0 = (inc0*G + A H_1)/cnt.  The sign-noise perturbation has identity gradient (SimGCL.py:51).
// With d_bitmap (rows where G / Gcl are non-zero) the first product A*H_K only gathers flagged columns.
static int propagate_bwd_impl(const idg_graph* g, const float* d_G, const float* d_Gcl, int32_t d, int32_t K, int include_layer0,
                              int32_t cl_layer, float* d_gX0, float* d_work, const uint32_t* d_bitmap, const idg_adam_args* adam,
                              void* stream) {
    if (!g || !d_G || (!d_gX0 && !adam) || !d_work) return fail(-1, "idg_propagate_bwd: null argument%s");
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
        SpmmExtra ex;
        ex.addend_mask = d_bitmap;   // G (and Gcl) are zero outside the flagged rows
        if (s == 1) ex.bitmap = d_bitmap;
        // with the closure registered and K >= 3 the next product reads H_{K-1} at closure columns only: rows outside it
        // need not be produced at all (they would be exact zeros)
        if (s == 1 && K >= 3 && d_bitmap && g->closure) ex.rowmask = g->closure;
        // H_{K-1} = G + A G is non-zero only on the batch rows and their neighbours: the next product gathers just those
        if (s == 2 && layer > 0 && d_bitmap && g->closure) ex.bitmap = g->closure;
        if (layer > 0) {
            float* y = buf[pb]; pb ^= 1;
            rc = spmm_launch(g, h, y, d_G, (d_Gcl && cl_layer == layer) ? d_Gcl : nullptr, cnt, nullptr, 0.f, nullptr, nullptr, 1.f, d, stream, ex);
            if (rc) return rc;
            h = y;
        } else {
            ex.adam = adam;
            rc = spmm_launch(g, h, nullptr, include_layer0 ? d_G : nullptr, nullptr, 0.f, nullptr, 0.f, nullptr, adam ? nullptr : d_gX0, cnt, d, stream, ex);
            if (rc) return rc;
        }
    }
    return 0;
}

extern "C" int idg_propagate_bwd_ex(const idg_graph* g, const float* d_G, const float* d_Gcl, int32_t d, int32_t K,
                                    int include_layer0, int32_t cl_layer, float* d_gX0, float* d_work, const uint32_t* d_bitmap,
                                    void* stream) {
    return propagate_bwd_impl(g, d_G, d_Gcl, d, K, include_layer0, cl_layer, d_gX0, d_work, d_bitmap, nullptr, stream);
}

// Same chain, but the last product's epilogue applies the Adam update in place of writing the gradient.
extern "C" int idg_propagate_bwd_adam(const idg_graph* g, const float* d_G, const float* d_Gcl, int32_t d, int32_t K,
                                      int include_layer0, int32_t cl_layer, float* d_work, const uint32_t* d_bitmap,
                                      const idg_adam_args* adam, void* stream) {
    if (!adam) return fail(-1, "idg_propagate_bwd_adam: null adam%s");
    return propagate_bwd_impl(g, d_G, d_Gcl, d, K, include_layer0, cl_layer, nullptr, d_work, d_bitmap, adam, stream);
}

extern "C" int idg_propagate_bwd(const idg_graph* g, const float* d_G, const float* d_Gcl, int32_t d, int32_t K,
                                 int include_layer0, int32_t cl_layer, float* d_gX0, float* d_work, void* stream) {
    return idg_propagate_bwd_ex(g, d_G, d_Gcl, d, K, include_layer0, cl_layer, d_gX0, d_work, nullptr, stream);
}
