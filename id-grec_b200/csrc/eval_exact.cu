// Exact full ranking of selected users (pass C of csrc/eval.cu, and the whole evaluation for shapes outside the tiled
// candidate kernels): score = fp64 dot product of the fp32 embeddings accumulated in index order k = 0..d-1, train
// positives removed, order (score desc, item id asc) -- the T0 definition of SURVEY.md section 7.
//
// Replaces models/LightGCN.py:74-80 + utility_train/batch_test.py:62-68 for the users the candidate filter could not
// settle (candidate overflow from near-ties at Xavier-scale scores, fewer than K unmasked items).  The round-1 version
// gave each such user to ONE CTA that wrote all I scores to global memory and scanned them K times: 70-100 ms for a
// single user at I = 1,000,000, which is what made the 8-GPU XL evaluation 18x slower than the 1-GPU one.
//
//   * work item = (flagged user f, item slice s).  The number of slices per user is chosen ON THE DEVICE from the
//     flagged-user count (few users -> many slices each, so the whole GPU works on them; many users -> one slice each).
//   * a work item streams its slice in chunks of 4,096 items: coalesced row loads staged through shared memory, one
//     fp64 sequential dot product per lane, scores kept in shared memory only.
//   * selection keeps a running sorted top-K list per work item: after the first chunk only entries that beat the
//     current K-th entry are candidates (typically K/j of them in chunk j); one warp inserts them.  No score ever
//     goes to global memory; the K-round block arg-max is only used to seed the list or when a chunk floods it.
//   * a merge kernel (one CTA per user) reduces the per-slice lists to the final K ids.
// Deterministic: every comparison uses the strict total order (score desc, id asc); ids are distinct.
#include <math.h>

#include "idg_common.cuh"

namespace idg {

constexpr int kXChunk = 4096;   // items scored per pass through shared memory
constexpr int kXMaxK = 256;     // list capacity (top_K above this is rejected by the caller)
constexpr int kXMaxD = 1024;
constexpr int kXInsertMax = 128;  // more candidates than this in one chunk -> rebuild the list by arg-max rounds
constexpr int kXThreads = 256;

struct XEntry { double s; int id; int pad; };

__device__ __forceinline__ bool x_better(double s1, int i1, double s2, int i2) { return (s1 > s2) || (s1 == s2 && i1 < i2); }

// arg-best of one round over two (score, id) arrays in shared memory; id < 0 = taken / absent.  All threads call it.
// Returns (array 0/1, index) through shared memory; -1 when both arrays are exhausted.
__device__ void x_argbest(const double* s0, const int* i0, int n0, const double* s1, const int* i1, int n1, double* rb, int* ri, int* rw,
                          int* out_which, int* out_idx) {
    double bs = 0.0; int bi = -1, bw = -1, be = -1;
    for (int e = threadIdx.x; e < n0; e += kXThreads) {
        const int id = i0[e];
        if (id < 0) continue;
        const double v = s0[e];
        if (bi < 0 || x_better(v, id, bs, bi)) { bs = v; bi = id; bw = 0; be = e; }
    }
    for (int e = threadIdx.x; e < n1; e += kXThreads) {
        const int id = i1[e];
        if (id < 0) continue;
        const double v = s1[e];
        if (bi < 0 || x_better(v, id, bs, bi)) { bs = v; bi = id; bw = 1; be = e; }
    }
    int code = (bw << 24) | (be & 0xffffff);   // chunk / list indices are < 2^24
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, bs, m);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
        const int oc = __shfl_xor_sync(0xffffffffu, code, m);
        if (oi >= 0 && (bi < 0 || x_better(os, oi, bs, bi))) { bs = os; bi = oi; code = oc; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { rb[warp] = bs; ri[warp] = bi; rw[warp] = code; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = rb[0]; int ii = ri[0], cc = rw[0];
        for (int q = 1; q < kXThreads / 32; ++q)
            if (ri[q] >= 0 && (ii < 0 || x_better(rb[q], ri[q], b, ii))) { b = rb[q]; ii = ri[q]; cc = rw[q]; }
        *out_which = (ii >= 0) ? (cc >> 24) : -1;
        *out_idx = cc & 0xffffff;
    }
    __syncthreads();
}

// slices per flagged user, the same on every CTA of both kernels
__device__ __forceinline__ int x_slices(int nflag, int I, int target_items) {
    const int max_slices = (I + kXChunk - 1) / kXChunk;
    int ns = (target_items + max(nflag, 1) - 1) / max(nflag, 1);
    return max(1, min(ns, max_slices));
}

__global__ void __launch_bounds__(kXThreads, 2) eval_exact_slices_kernel(const float* __restrict__ Fu, const float* __restrict__ Fi, int I, int d,
                                                                          const int32_t* __restrict__ mptr, const int32_t* __restrict__ mind,
                                                                          const int64_t* __restrict__ users, int K, const int* __restrict__ flag_cnt,
                                                                          const int* __restrict__ flag_list, int target_items, XEntry* __restrict__ part) {
    extern __shared__ __align__(16) unsigned char xsm[];
    double* S = reinterpret_cast<double*>(xsm);                 // [kXChunk] chunk scores, reused as candidate scores
    double* ud = S + kXChunk;                                    // [kXMaxD] user row as doubles
    double* ls = ud + kXMaxD;                                    // [kXMaxK] running list, sorted best first
    double* ns_ = ls + kXMaxK;                                   // [kXMaxK] rebuilt list
    int* CI = reinterpret_cast<int*>(ns_ + kXMaxK);              // [kXChunk] candidate ids (-1 = taken)
    int* li = CI + kXChunk;                                      // [kXMaxK]
    int* ni = li + kXMaxK;                                       // [kXMaxK]
    float* tile = reinterpret_cast<float*>(ni + kXMaxK);         // [8 warps][32][33]
    __shared__ double rb[kXThreads / 32];
    __shared__ int ri[kXThreads / 32], rw[kXThreads / 32];
    __shared__ int s_which, s_idx, s_ncand, s_cnt;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nflag = *flag_cnt;
    const int nsl = x_slices(nflag, I, target_items);
    const int n_chunks_total = (I + kXChunk - 1) / kXChunk;
    const long long n_work = (long long)nflag * nsl;
    float* mytile = tile + warp * 32 * 33;

    for (long long wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
        const int f = (int)(wk / nsl), sl = (int)(wk % nsl);
        const int p = flag_list[f];
        const int u = (int)users[p];
        // chunks [c_begin, c_end) of this slice (balanced split of the chunk range)
        const int c_begin = (int)((long long)n_chunks_total * sl / nsl), c_end = (int)((long long)n_chunks_total * (sl + 1) / nsl);
        __syncthreads();  // previous work item fully done with shared memory
        for (int k = tid; k < d; k += kXThreads) ud[k] = (double)Fu[(size_t)u * d + k];
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        const int mlo = mptr[u], mhi = mptr[u + 1];

        for (int c = c_begin; c < c_end; ++c) {
            const int lo = c * kXChunk, n_c = min(kXChunk, I - lo);
            // ---- exact scores of the chunk: each warp takes 32 rows at a time, k in blocks of 32 ----
            for (int r0 = warp * 32; r0 < n_c; r0 += (kXThreads / 32) * 32) {
                double acc = 0.0;
                for (int k0 = 0; k0 < d; k0 += 32) {
                    __syncwarp();
#pragma unroll 4
                    for (int r = 0; r < 32; ++r) {
                        const int row = lo + r0 + r;
                        mytile[r * 33 + lane] = (r0 + r < n_c && k0 + lane < d) ? __ldg(Fi + (size_t)row * d + k0 + lane) : 0.f;
                    }
                    __syncwarp();
                    const int kn = min(32, d - k0);
                    for (int k = 0; k < kn; ++k) acc = fma(ud[k0 + k], (double)mytile[lane * 33 + k], acc);
                }
                if (r0 + lane < n_c) S[r0 + lane] = acc;
            }
            __syncthreads();
            // ---- train positives inside the chunk score -inf (they stay eligible, in id order, after all real scores) ----
            {
                int a = mlo, b = mhi;
                while (a < b) { const int mid = (a + b) >> 1; if (__ldg(mind + mid) < lo) a = mid + 1; else b = mid; }
                for (int j = a + tid; j < mhi; j += kXThreads) {
                    const int item = __ldg(mind + j);
                    if (item >= lo + n_c) break;
                    S[item - lo] = -INFINITY;
                }
            }
            __syncthreads();
            // ---- candidates: everything while the list is short, else only entries that beat the current K-th ----
            const int cnt = s_cnt;
            const bool full = cnt >= K;
            if (!full) {
                for (int e = tid; e < n_c; e += kXThreads) CI[e] = lo + e;
                if (tid == 0) s_ncand = n_c;
                __syncthreads();
            } else {
                const double ts = ls[K - 1];
                const int ti = li[K - 1];
                if (tid == 0) s_ncand = 0;
                __syncthreads();
                // compact in place: position q <= e for every e, and a slot is only written after ... the scores are
                // first read into registers by all threads, then written: two phases around a barrier
                double v[kXChunk / kXThreads]; int keep = 0;
#pragma unroll
                for (int q = 0; q < kXChunk / kXThreads; ++q) {
                    const int e = tid + q * kXThreads;
                    v[q] = (e < n_c) ? S[e] : 0.0;
                    if (e < n_c && x_better(v[q], lo + e, ts, ti)) keep |= 1 << q;
                }
                __syncthreads();
                if (keep) {
#pragma unroll
                    for (int q = 0; q < kXChunk / kXThreads; ++q)
                        if ((keep >> q) & 1) { const int pos = atomicAdd(&s_ncand, 1); S[pos] = v[q]; CI[pos] = lo + tid + q * kXThreads; }
                }
                __syncthreads();
            }
            const int ncand = s_ncand;
            if (ncand == 0) continue;
            if (full && ncand <= kXInsertMax) {
                // ---- warp 0 inserts the few newcomers into the sorted list (the K-th entry falls off) ----
                if (warp == 0) {
                    for (int j = 0; j < ncand; ++j) {
                        const double cs = S[j]; const int cid = CI[j];
                        int ahead = 0;
                        for (int t = lane; t < K; t += 32) ahead += x_better(ls[t], li[t], cs, cid) ? 1 : 0;
#pragma unroll
                        for (int m = 16; m >= 1; m >>= 1) ahead += __shfl_xor_sync(0xffffffffu, ahead, m);
                        if (ahead >= K) continue;          // a previous insertion of this chunk pushed it out
                        // shift [ahead, K-2] -> [ahead+1, K-1]: read everything, then write
                        double ts_[kXMaxK / 32]; int ti_[kXMaxK / 32];
#pragma unroll
                        for (int q = 0; q < kXMaxK / 32; ++q) {
                            const int t = lane + 32 * q;
                            if (t >= ahead && t < K - 1) { ts_[q] = ls[t]; ti_[q] = li[t]; }
                        }
                        __syncwarp();
#pragma unroll
                        for (int q = 0; q < kXMaxK / 32; ++q) {
                            const int t = lane + 32 * q;
                            if (t >= ahead && t < K - 1) { ls[t + 1] = ts_[q]; li[t + 1] = ti_[q]; }
                        }
                        if (lane == 0) { ls[ahead] = cs; li[ahead] = cid; }
                        __syncwarp();
                    }
                }
                __syncthreads();
            } else {
                // ---- rebuild: K rounds of block arg-best over candidates + current list ----
                int produced = 0;
                for (int r = 0; r < K; ++r) {
                    x_argbest(S, CI, ncand, ls, li, cnt, rb, ri, rw, &s_which, &s_idx);
                    const int which = s_which, idx = s_idx;
                    if (which < 0) break;
                    if (tid == 0) {
                        if (which == 0) { ns_[r] = S[idx]; ni[r] = CI[idx]; CI[idx] = -1; }
                        else { ns_[r] = ls[idx]; ni[r] = li[idx]; li[idx] = -1; }
                    }
                    ++produced;
                    __syncthreads();
                }
                for (int t = tid; t < produced; t += kXThreads) { ls[t] = ns_[t]; li[t] = ni[t]; }
                if (tid == 0) s_cnt = produced;
                __syncthreads();
            }
        }
        __syncthreads();
        // ---- publish this slice's list (absent entries sort after every real one) ----
        XEntry* out = part + (size_t)wk * K;
        const int cnt = s_cnt;
        for (int t = tid; t < K; t += kXThreads) {
            XEntry e;
            e.s = (t < cnt) ? ls[t] : -INFINITY; e.id = (t < cnt) ? li[t] : 0x7fffffff; e.pad = 0;
            out[t] = e;
        }
    }
}

// one CTA per flagged user: K rounds over its n_slices x K partial entries (L2-resident), in place
__global__ void __launch_bounds__(kXThreads) eval_exact_merge_kernel(int I, int K, const int* __restrict__ flag_cnt, const int* __restrict__ flag_list,
                                                                      int target_items, XEntry* __restrict__ part, int64_t* __restrict__ out_ids,
                                                                      float* __restrict__ out_scores) {
    __shared__ double rb[kXThreads / 32];
    __shared__ int ri[kXThreads / 32], re[kXThreads / 32];
    const int nflag = *flag_cnt;
    const int nsl = x_slices(nflag, I, target_items);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int f = blockIdx.x; f < nflag; f += gridDim.x) {
        const int p = flag_list[f];
        XEntry* E = part + (size_t)f * nsl * K;
        const int n = nsl * K;
        for (int r = 0; r < K; ++r) {
            double bs = 0.0; int bi = -1, be = -1;
            for (int e = tid; e < n; e += kXThreads) {
                const int id = E[e].id;
                if (id < 0) continue;
                const double v = E[e].s;
                if (bi < 0 || x_better(v, id, bs, bi)) { bs = v; bi = id; be = e; }
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                const double os = __shfl_xor_sync(0xffffffffu, bs, m);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
                const int oe = __shfl_xor_sync(0xffffffffu, be, m);
                if (oi >= 0 && (bi < 0 || x_better(os, oi, bs, bi))) { bs = os; bi = oi; be = oe; }
            }
            if (lane == 0) { rb[warp] = bs; ri[warp] = bi; re[warp] = be; }
            __syncthreads();
            if (tid == 0) {
                double b = rb[0]; int ii = ri[0], ee = re[0];
                for (int q = 1; q < kXThreads / 32; ++q)
                    if (ri[q] >= 0 && (ii < 0 || x_better(rb[q], ri[q], b, ii))) { b = rb[q]; ii = ri[q]; ee = re[q]; }
                const bool real = ii >= 0 && ii != 0x7fffffff;
                out_ids[(size_t)p * K + r] = real ? ii : 0;
                if (out_scores) out_scores[(size_t)p * K + r] = real ? (float)b : -INFINITY;
                if (ii >= 0) E[ee].id = -1;
            }
            __syncthreads();
        }
    }
}

size_t eval_exact_part_bytes(int nu, int K) { return sizeof(XEntry) * ((size_t)nu + 2 * 2 * kNumSMs + 64) * (size_t)K; }

int launch_eval_exact(const float* Fu, const float* Fi, int I, int d, const int32_t* mptr, const int32_t* mind, const int64_t* users, int K,
                      const int* flag_cnt, const int* flag_list, void* part, int64_t* out_ids, float* out_scores, cudaStream_t stream) {
    if (K > kXMaxK) return fail(-1, "idg_eval_topk: the exact ranking pass keeps at most 256 entries per user (K = %s%lld)", "", K);
    if (d > kXMaxD) return fail(-1, "idg_eval_topk: embedding width above 1024 (%s%lld)", "", d);
    const size_t smem = sizeof(double) * (kXChunk + kXMaxD + 2 * kXMaxK) + sizeof(int) * (kXChunk + 2 * kXMaxK) + sizeof(float) * (kXThreads / 32) * 32 * 33;
    IDG_CUDA(cudaFuncSetAttribute(eval_exact_slices_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = 2 * kNumSMs;
    const int target = 2 * grid;   // work items aimed at when few users are flagged; part holds (nflag + target) * K entries
    eval_exact_slices_kernel<<<grid, kXThreads, smem, stream>>>(Fu, Fi, I, d, mptr, mind, users, K, flag_cnt, flag_list, target, (XEntry*)part);
    IDG_LAUNCH_CHECK("eval_exact_slices_kernel");
    eval_exact_merge_kernel<<<grid, kXThreads, 0, stream>>>(I, K, flag_cnt, flag_list, target, (XEntry*)part, out_ids, out_scores);
    IDG_LAUNCH_CHECK("eval_exact_merge_kernel");
    return 0;
}

}  // namespace idg
