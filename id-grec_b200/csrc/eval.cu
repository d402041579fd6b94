// Full-ranking evaluation: score = <Fu[user], Fi[item]>, train positives removed, top-K by
// (score desc, item id asc), recall/precision/ndcg sums (sm_100a).
//
// Replaces models/LightGCN.py:74-80 (get_rating_for_test: gather + matmul + sigmoid into a
// [1024, I] matrix), utility_train/batch_test.py:54-68 (Python mask lists, rating[eu,ei] = -1,
// torch.topk) and utility_function/metrics.py:4-58.  sigmoid is monotone, so ranking the raw
// dot products ranks the ratings; the [b, I] matrix is never written.
//
// Exactness contract (tier T0): the returned ids equal the order defined
// by the fp64 sequential dot product of the fp32 embeddings, ties by ascending id.
//   pass A  candidate scores s~ with |s~ - s| <= delta_u (fp32 FMA: delta_u = g64*|u|*max|i|).
//           Per user, every unmasked item with s~ >= (running K-th largest s~) - 2*delta_u is
//           kept; that set provably contains the exact top-K including all boundary ties.
//   pass B  survivors are rescored exactly (fp64, k = 0..d-1) and ordered (score desc, id asc).
//   pass C  users whose candidate list overflowed or came up short are ranked exactly over all items by the sliced
//           pass of csrc/eval_exact.cu (the whole GPU cooperates on few users; nothing but top-K lists leaves the SM).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "idg_common.cuh"

namespace idg {

// csrc/eval_tc.cu: the same candidate pass on tcgen05 tensor cores (d = 64, and d = 256 for NGCF's concatenated layers)
int launch_eval_candidates_tc(const float* Fu, const float* Fi, int I, int d, const int32_t* mptr, const int32_t* mind, const int64_t* users,
                              int nu, int K, float* max_norm, int* flag_cnt, int* flag_list, int* cand_cnt, int* cand_ids,
                              float* list_s, int* list_i, const float* aug, float* dbg_tile, cudaStream_t stream);
constexpr int kTcListCap = 80;  // == kTcCap in eval_tc.cu

constexpr int kTU = 64, kTI = 64, kCap = 128, kPruneAt = 64, kCandOut = 64;
// csrc/eval_exact.cu: pass C, exact sliced ranking of the flagged users (and of every user for shapes outside the tiles)
size_t eval_exact_part_bytes(int nu, int K);
int launch_eval_exact(const float* Fu, const float* Fi, int I, int d, const int32_t* mptr, const int32_t* mind, const int64_t* users, int K,
                      const int* flag_cnt, const int* flag_list, void* part, int64_t* out_ids, float* out_scores, cudaStream_t stream);

struct EvalWs {
    float* max_norm;   // [1] max_i |Fi[i]|_2 (as float bits, atomicMax on non-negative floats)
    int* flag_cnt;     // [1]
    int* flag_list;    // [nu] positions (into d_users) that need the exhaustive pass
    int* cand_cnt;     // [nu]
    int* cand_ids;     // [nu, kCandOut]
    float* tc_ls;      // [ceil(nu/128)*128, kTcListCap] tensor-core pass: per-row candidate lists
    int* tc_li;
    float* tc_aug;     // [I, 8] margin operand of the tensor-core pass: column 0 = |i| rounded up to tf32
    float* tc_fir;     // [I, d] item table rounded to nearest tf32 (what the tensor core reads through TMA)
    void* part;        // per-slice top-K lists of the exact pass (eval_exact_part_bytes)
};

__host__ __device__ inline size_t ev_align(size_t x) { return (x + 255) & ~(size_t)255; }

__host__ inline EvalWs eval_carve(void* ws, int nu, int I, int d) {
    char* p = (char*)ws;
    EvalWs w;
    w.max_norm = (float*)p; p += 256;
    w.flag_cnt = (int*)p; p += 256;
    w.flag_list = (int*)p; p += ev_align(sizeof(int) * (size_t)nu);
    w.cand_cnt = (int*)p; p += ev_align(sizeof(int) * (size_t)nu);
    w.cand_ids = (int*)p; p += ev_align(sizeof(int) * (size_t)nu * kCandOut);
    const size_t nup = ((size_t)nu + 127) / 128 * 128;
    w.tc_ls = (float*)p; p += ev_align(sizeof(float) * nup * kTcListCap);
    w.tc_li = (int*)p; p += ev_align(sizeof(int) * nup * kTcListCap);
    w.tc_aug = (float*)p; p += ev_align(sizeof(float) * 8 * (size_t)I);
    w.tc_fir = (float*)p; p += ev_align(sizeof(float) * (size_t)(d == 256 ? 256 : 64) * (size_t)I);
    w.part = (void*)p;
    return w;
}

// max_i |i| (CUDA-core pass: one margin per user) and, for the tensor-core pass, the per-item margin operand
// aug[i] = {|i| rounded UP to a tf32-exact value (plus a relative 2^-20 for the rounding of the norm itself), 0 x 7}
__global__ void item_norm_kernel(const float* __restrict__ Fi, int I, int d, float* __restrict__ max_norm, float* __restrict__ aug,
                                 float* __restrict__ Fi_tf32) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= I) return;
    float ss = 0.f;
    for (int k = lane; k < d; k += 32) {
        const float v = Fi[(size_t)i * d + k];
        ss = fmaf(v, v, ss);
        if (Fi_tf32) {   // the copy the tensor core reads: rounded to nearest tf32, so the unit's own truncation is exact
            uint32_t r;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
            Fi_tf32[(size_t)i * d + k] = __uint_as_float(r);
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, m);
    const float nrm = sqrtf(ss);
    if (lane == 0) atomicMax(reinterpret_cast<int*>(max_norm), __float_as_int(nrm));
    if (aug && lane < 8) {
        const float up = __uint_as_float((__float_as_uint(nrm * 1.000001f) + 0x1fffu) & 0xffffe000u);
        aug[(size_t)i * 8 + lane] = (lane == 0) ? up : 0.f;
    }
}

__device__ __forceinline__ bool masked(const int32_t* __restrict__ ind, int lo, int hi, int item) {
    const int end = hi;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(ind + mid) < item) lo = mid + 1; else hi = mid; }
    return lo < end && __ldg(ind + lo) == item;
}

// Prune one user's smem candidate list (one warp): tau = (K-th largest) - 2*delta; keep >= tau.
__device__ void prune_row(float* __restrict__ ls, int* __restrict__ li, int* __restrict__ cnt_p, float* __restrict__ thr_p,
                          float delta2, int K, int lane) {
    const int m = *cnt_p;
    if (m < K) return;
    float s[4]; int id[4]; int rank[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int idx = lane + 32 * q;
        s[q] = (idx < m) ? ls[idx] : -INFINITY; id[q] = (idx < m) ? li[idx] : 0; rank[q] = 0;
    }
    for (int j = 0; j < m; ++j) {
        const float sj = ls[j];
#pragma unroll
        for (int q = 0; q < 4; ++q) rank[q] += (sj > s[q]) || (sj == s[q] && j < lane + 32 * q);
    }
    float vk = -INFINITY;
#pragma unroll
    for (int q = 0; q < 4; ++q) if (lane + 32 * q < m && rank[q] == K - 1) vk = s[q];
#pragma unroll
    for (int mm = 16; mm >= 1; mm >>= 1) vk = fmaxf(vk, __shfl_xor_sync(0xffffffffu, vk, mm));
    const float tau = vk - delta2;
    __syncwarp();
    int kept = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const bool keep = (lane + 32 * q < m) && (s[q] >= tau);
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (keep) { const int p = kept + __popc(b & ((1u << lane) - 1)); ls[p] = s[q]; li[p] = id[q]; }
        kept += __popc(b);
    }
    __syncwarp();
    if (lane == 0) { *cnt_p = kept; *thr_p = (kept > kPruneAt) ? INFINITY : tau; }
}

// Pass A.  grid = ceil(nu / 64); 256 threads; thread (ty,tx) owns users ty*4.. x items tx*4..
template <int D>
__global__ void __launch_bounds__(256, 2) eval_candidates_kernel(const float* __restrict__ Fu, const float* __restrict__ Fi, int I,
                                                                 const int32_t* __restrict__ mptr, const int32_t* __restrict__ mind,
                                                                 const int64_t* __restrict__ users, int nu, int K, EvalWs w) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Us = reinterpret_cast<float*>(smem_raw);   // [D][64]
    float* Is = Us + D * kTU;                          // [D][64]
    float* ls = Is + D * kTI;                          // [64][kCap]
    int* li = reinterpret_cast<int*>(ls + kTU * kCap); // [64][kCap]
    float* thr = reinterpret_cast<float*>(li + kTU * kCap);  // [64]
    float* del2 = thr + kTU;                           // [64] 2*delta_u
    int* cnt = reinterpret_cast<int*>(del2 + kTU);     // [64]
    int* urow = cnt + kTU;                             // [64] user id or -1
    __shared__ int s_overflow[kTU];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int u0 = blockIdx.x * kTU;
    if (tid < kTU) {
        const int p = u0 + tid;
        urow[tid] = (p < nu) ? (int)users[p] : -1;
        cnt[tid] = 0; s_overflow[tid] = 0;
        thr[tid] = (p < nu) ? -INFINITY : INFINITY;
    }
    __syncthreads();
    // user tile, transposed: lane <-> user row (conflict-free smem stores)
    for (int r = warp; r < 2 * (D / 4); r += 8) {
        const int ul = (r & 1) * 32 + lane, c4 = r >> 1;
        const int u = urow[ul];
        float4 v = (u >= 0) ? ldg4(Fu + (size_t)u * D + c4 * 4) : f4zero();
        Us[(c4 * 4 + 0) * kTU + ul] = v.x; Us[(c4 * 4 + 1) * kTU + ul] = v.y;
        Us[(c4 * 4 + 2) * kTU + ul] = v.z; Us[(c4 * 4 + 3) * kTU + ul] = v.w;
    }
    __syncthreads();
    if (tid < kTU) {
        float ss = 0.f;
        for (int k = 0; k < D; ++k) { const float v = Us[k * kTU + tid]; ss = fmaf(v, v, ss); }
        // fp32 FMA dot of length D: |err| <= gamma_D * |u||i|, gamma_D ~ D*2^-24; 1.25x slack covers
        // the rounding of the norms themselves and the (negligible) fp64 reference error.
        del2[tid] = 2.f * 1.25f * (float)D * 5.9604645e-8f * sqrtf(ss) * (*w.max_norm) + 1e-30f;
    }
    const int ty = tid >> 4, tx = tid & 15;
    int mlo[4], mhi[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int u = urow[ty * 4 + a];
        mlo[a] = (u >= 0) ? mptr[u] : 0; mhi[a] = (u >= 0) ? mptr[u + 1] : 0;
    }

    for (int i0 = 0; i0 < I; i0 += kTI) {
        __syncthreads();  // previous tile fully consumed (and del2/thr visible on the first pass)
        for (int r = warp; r < 2 * (D / 4); r += 8) {
            const int il = (r & 1) * 32 + lane, c4 = r >> 1;
            const int i = i0 + il;
            float4 v = (i < I) ? ldg4(Fi + (size_t)i * D + c4 * 4) : f4zero();
            Is[(c4 * 4 + 0) * kTI + il] = v.x; Is[(c4 * 4 + 1) * kTI + il] = v.y;
            Is[(c4 * 4 + 2) * kTI + il] = v.z; Is[(c4 * 4 + 3) * kTI + il] = v.w;
        }
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 8
        for (int k = 0; k < D; ++k) {
            const float4 uu = *reinterpret_cast<const float4*>(Us + k * kTU + ty * 4);
            const float4 ii = *reinterpret_cast<const float4*>(Is + k * kTI + tx * 4);
            const float ua[4] = {uu.x, uu.y, uu.z, uu.w}, ib[4] = {ii.x, ii.y, ii.z, ii.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(ua[a], ib[b], acc[a][b]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int ul = ty * 4 + a;
            const float t = thr[ul];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int item = i0 + tx * 4 + b;
                if (acc[a][b] >= t && item < I) {
                    if (!masked(mind, mlo[a], mhi[a], item)) {
                        const int p = atomicAdd(&cnt[ul], 1);
                        if (p < kCap) { ls[ul * kCap + p] = acc[a][b]; li[ul * kCap + p] = item; }
                        else s_overflow[ul] = 1;
                    }
                }
            }
        }
        __syncthreads();
        for (int r = warp * 8; r < warp * 8 + 8; ++r) {
            if (cnt[r] > kCap) { if (lane == 0) { cnt[r] = kCap; s_overflow[r] = 1; } __syncwarp(); }
            if (cnt[r] > kPruneAt) prune_row(ls + r * kCap, li + r * kCap, &cnt[r], &thr[r], del2[r], K, lane);
        }
    }
    __syncthreads();
    for (int r = warp * 8; r < warp * 8 + 8; ++r) {
        const int p = u0 + r;
        if (p >= nu) continue;
        prune_row(ls + r * kCap, li + r * kCap, &cnt[r], &thr[r], del2[r], K, lane);
        __syncwarp();
        const int m = cnt[r];
        const bool bad = s_overflow[r] || m > kCandOut || m < K || !(thr[r] < INFINITY);
        if (bad) {
            if (lane == 0) { w.cand_cnt[p] = 0; w.flag_list[atomicAdd(w.flag_cnt, 1)] = p; }
        } else {
            if (lane == 0) w.cand_cnt[p] = m;
            for (int j = lane; j < m; j += 32) w.cand_ids[(size_t)p * kCandOut + j] = li[r * kCap + j];
        }
    }
}

__device__ __forceinline__ double exact_dot(const float* __restrict__ a, const float* __restrict__ b, int d) {
    double acc = 0.0;
    for (int k = 0; k < d; ++k) acc = fma((double)a[k], (double)b[k], acc);  // products exact in fp64; order k = 0..d-1
    return acc;
}

// Pass B: one warp per user; <= 64 candidates, exact rescore, rank by (score desc, id asc).
__global__ void __launch_bounds__(256) eval_rescore_kernel(const float* __restrict__ Fu, const float* __restrict__ Fi, int d,
                                                           const int64_t* __restrict__ users, int nu, int K, EvalWs w,
                                                           int64_t* __restrict__ out_ids, float* __restrict__ out_scores) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (p >= nu) return;
    const int m = w.cand_cnt[p];
    if (m == 0) return;  // flagged
    const float* urow = Fu + (size_t)users[p] * d;
    double s[2]; int id[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int j = lane + 32 * q;
        id[q] = (j < m) ? w.cand_ids[(size_t)p * kCandOut + j] : 0x7fffffff;
        s[q] = (j < m) ? exact_dot(urow, Fi + (size_t)id[q] * d, d) : -INFINITY;
    }
    int rank[2] = {0, 0};
#pragma unroll
    for (int q2 = 0; q2 < 2; ++q2) {
        for (int l = 0; l < 32; ++l) {
            const double sj = __shfl_sync(0xffffffffu, s[q2], l);
            const int idj = __shfl_sync(0xffffffffu, id[q2], l);
#pragma unroll
            for (int q = 0; q < 2; ++q) rank[q] += (sj > s[q]) || (sj == s[q] && idj < id[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (lane + 32 * q < m && rank[q] < K) {
            out_ids[(size_t)p * K + rank[q]] = id[q];
            if (out_scores) out_scores[(size_t)p * K + rank[q]] = (float)s[q];
        }
    }
}

// Shapes outside the fast kernels (d not in {32, 64, 128, 256}, K > 48): every user takes pass C.
__global__ void eval_flag_all_kernel(int nu, EvalWs w) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nu) { w.flag_list[p] = p; w.cand_cnt[p] = 0; }
    if (p == 0) *w.flag_cnt = nu;
}

// metrics.py:4-58 + batch_test.py:80-107: per-user recall/precision/ndcg terms, then an ordered sum.
__global__ void __launch_bounds__(256) eval_metrics_kernel(const int64_t* __restrict__ topk, const int64_t* __restrict__ users, int nu,
                                                           int K, const int32_t* __restrict__ tptr, const int32_t* __restrict__ tind,
                                                           int nk, int k0, int k1, int k2, int k3, int k4, int k5, int k6, int k7,
                                                           double* __restrict__ per_user) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (p >= nu) return;
    const int ks[8] = {k0, k1, k2, k3, k4, k5, k6, k7};
    const int u = (int)users[p];
    const int lo = tptr[u], hi = tptr[u + 1], nt = hi - lo;
    // hit bits of 32 positions at a time (lane <-> position); lane 0 folds them into the running sums in position order
    double nh[8], dcg[8], idcg[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) { nh[t] = 0.0; dcg[t] = 0.0; idcg[t] = 0.0; }
    for (int j0 = 0; j0 < K; j0 += 32) {
        const int j = j0 + lane;
        bool h = false;
        if (j < K) {
            const int item = (int)topk[(size_t)p * K + j];
            int a = lo, b = hi;
            while (a < b) { const int mid = (a + b) >> 1; if (tind[mid] < item) a = mid + 1; else b = mid; }
            h = (a < hi && tind[a] == item);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, h);
        if (lane == 0) {
            const int jn = min(K, j0 + 32);
            for (int jj = j0; jj < jn; ++jj) {
                const double disc = 1.0 / log2((double)(jj + 2));
                const bool hit = (bal >> (jj - j0)) & 1u;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    if (t < nk && jj < ks[t]) {
                        if (hit) { nh[t] += 1.0; dcg[t] += disc; }
                        if (jj < nt) idcg[t] += disc;
                    }
                }
            }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            if (t >= nk) break;
            const double id = (idcg[t] == 0.0) ? 1.0 : idcg[t];
            per_user[(size_t)p * 3 * nk + 3 * t + 0] = (nt > 0) ? nh[t] / (double)nt : 0.0;  // recall term
            per_user[(size_t)p * 3 * nk + 3 * t + 1] = nh[t] / (double)ks[t];                 // precision term
            per_user[(size_t)p * 3 * nk + 3 * t + 2] = dcg[t] / id;                           // ndcg term
        }
    }
}

__global__ void __launch_bounds__(1024) ordered_sum_kernel(const double* __restrict__ per_user, int nu, int ncol, double* __restrict__ sums) {
    __shared__ double sh[1024];
    for (int c = 0; c < ncol; ++c) {
        double a = 0.0;
        for (int i = threadIdx.x; i < nu; i += 1024) a += per_user[(size_t)i * ncol + c];
        sh[threadIdx.x] = a;
        __syncthreads();
        for (int s = 512; s >= 1; s >>= 1) { if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s]; __syncthreads(); }
        if (threadIdx.x == 0) sums[c] = sh[0];
        __syncthreads();
    }
}

// models/LightGCN.py:74-80 as written: rating[b, i] = sigmoid(<Fu[users[b]], Fi[i]>) materialised as a dense matrix.
// Only for API completeness (Test() never calls it); 32x32 tiles, fp32 FMA in the reference's k order.
__global__ void __launch_bounds__(256) rating_matrix_kernel(const float* __restrict__ Fu, const float* __restrict__ Fi, const int64_t* __restrict__ users,
                                                            int nu, int I, int d, float* __restrict__ out) {
    __shared__ float su[32][33], si[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads, 4 rows each
    const int b0 = blockIdx.y * 32, i0 = blockIdx.x * 32;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < d; k0 += 32) {
        for (int r = ty; r < 32; r += 8) {
            const int b = b0 + r, i = i0 + r;
            su[r][tx] = (b < nu && k0 + tx < d) ? Fu[(size_t)users[b] * d + k0 + tx] : 0.f;
            si[r][tx] = (i < I && k0 + tx < d) ? Fi[(size_t)i * d + k0 + tx] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const float iv = si[tx][k];
#pragma unroll
            for (int a = 0; a < 4; ++a) acc[a] = fmaf(su[ty * 4 + a][k], iv, acc[a]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int b = b0 + ty * 4 + a, i = i0 + tx;
        if (b < nu && i < I) out[(size_t)b * I + i] = 1.f / (1.f + expf(-acc[a]));
    }
}

}  // namespace idg

using namespace idg;

extern "C" int idg_rating_matrix(const float* d_Fu, const float* d_Fi, const int64_t* d_users, int32_t nu, int32_t I, int32_t d, float* d_out,
                                 void* stream) {
    if (!d_Fu || !d_Fi || !d_users || !d_out || nu <= 0 || I <= 0 || d <= 0) return fail(-1, "idg_rating_matrix: bad argument%s");
    const dim3 grid((I + 31) / 32, (nu + 31) / 32);
    rating_matrix_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_Fu, d_Fi, d_users, nu, I, d, d_out);
    IDG_LAUNCH_CHECK("rating_matrix_kernel");
    return 0;
}

extern "C" int64_t idg_eval_workspace_bytes(int32_t nu, int32_t I, int32_t d, int32_t K) {
    if (nu <= 0 || I <= 0) return 0;
    const size_t metrics = ev_align(sizeof(double) * (size_t)nu * 3 * 8);
    const size_t nup = ((size_t)nu + 127) / 128 * 128;
    const size_t sel = 512 + 2 * ev_align(sizeof(int) * (size_t)nu) + ev_align(sizeof(int) * (size_t)nu * kCandOut) +
                       ev_align(sizeof(float) * nup * kTcListCap) + ev_align(sizeof(int) * nup * kTcListCap) + ev_align(sizeof(float) * 8 * (size_t)I) + ev_align(sizeof(float) * (size_t)(d == 256 ? 256 : 64) * (size_t)I) +
                       ev_align(eval_exact_part_bytes(nu, K > 0 ? K : 1));
    return (int64_t)(sel > metrics ? sel : metrics);
}

extern "C" int idg_eval_topk(const float* d_Fu, const float* d_Fi, int32_t U, int32_t I, int32_t d, const int32_t* d_mask_indptr,
                             const int32_t* d_mask_indices, const int64_t* d_users, int32_t nu, int32_t K, int64_t* d_out_ids,
                             float* d_out_scores, void* d_ws, void* stream_) {
    if (!d_Fu || !d_Fi || !d_mask_indptr || !d_users || !d_out_ids || !d_ws) return fail(-1, "idg_eval_topk: null argument%s");
    if (nu <= 0 || I <= 0 || U <= 0) return fail(-1, "idg_eval_topk: bad sizes%s");
    if (K < 1 || K > I) return fail(-1, "idg_eval_topk: K must be in [1, I] (%s%lld)", "", K);
    if (d < 1) return fail(-1, "idg_eval_topk: d must be positive (%s%lld)", "", d);
    cudaStream_t stream = (cudaStream_t)stream_;
    EvalWs w = eval_carve(d_ws, nu, I, d);
    IDG_CUDA(cudaMemsetAsync(d_ws, 0, 512, stream));
    if (K > 48 || (d != 32 && d != 64 && d != 128 && d != 256)) {
        // reference-legal but outside the tiled kernels (any embedding_size / top_K parses in the reference): exact
        // exhaustive ranking for every user -- same ids, one fp64 score row per user instead of the candidate filter
        eval_flag_all_kernel<<<(nu + 255) / 256, 256, 0, stream>>>(nu, w);
        IDG_LAUNCH_CHECK("eval_flag_all_kernel");
        return launch_eval_exact(d_Fu, d_Fi, I, d, d_mask_indptr, d_mask_indices, d_users, K, w.flag_cnt, w.flag_list, w.part, d_out_ids, d_out_scores, stream);
    }
    static const bool tc_on = !(getenv("IDG_EVAL_IMPL") && strcmp(getenv("IDG_EVAL_IMPL"), "fma") == 0);
    const bool use_tc = tc_on && (d == 64 || d == 256);   // tcgen05 pass: LightGCN-family width and NGCF's concatenated 256
    item_norm_kernel<<<(I + 7) / 8, 256, 0, stream>>>(d_Fi, I, d, w.max_norm, use_tc ? w.tc_aug : nullptr, use_tc ? w.tc_fir : nullptr);
    IDG_LAUNCH_CHECK("item_norm_kernel");
    const size_t smem = sizeof(float) * ((size_t)d * (kTU + kTI) + (size_t)kTU * kCap) + sizeof(int) * (size_t)kTU * kCap + sizeof(float) * 2 * kTU + sizeof(int) * 2 * kTU;
    const unsigned grid = (unsigned)((nu + kTU - 1) / kTU);
    // IDG_EVAL_IMPL=fma selects the CUDA-core candidate pass (kept as a cross-check of the tensor-core one)
    if (use_tc) {
        if (int rc = launch_eval_candidates_tc(d_Fu, w.tc_fir, I, d, d_mask_indptr, d_mask_indices, d_users, nu, K, w.max_norm, w.flag_cnt,
                                               w.flag_list, w.cand_cnt, w.cand_ids, w.tc_ls, w.tc_li, w.tc_aug, nullptr, stream))
            return rc;
    } else if (d == 64) {
        IDG_CUDA(cudaFuncSetAttribute(eval_candidates_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        eval_candidates_kernel<64><<<grid, 256, smem, stream>>>(d_Fu, d_Fi, I, d_mask_indptr, d_mask_indices, d_users, nu, K, w);
    } else if (d == 32) {
        IDG_CUDA(cudaFuncSetAttribute(eval_candidates_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        eval_candidates_kernel<32><<<grid, 256, smem, stream>>>(d_Fu, d_Fi, I, d_mask_indptr, d_mask_indices, d_users, nu, K, w);
    } else if (d == 128) {
        IDG_CUDA(cudaFuncSetAttribute(eval_candidates_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        eval_candidates_kernel<128><<<grid, 256, smem, stream>>>(d_Fu, d_Fi, I, d_mask_indptr, d_mask_indices, d_users, nu, K, w);
    } else {
        IDG_CUDA(cudaFuncSetAttribute(eval_candidates_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        eval_candidates_kernel<256><<<grid, 256, smem, stream>>>(d_Fu, d_Fi, I, d_mask_indptr, d_mask_indices, d_users, nu, K, w);
    }
    if (!use_tc) IDG_LAUNCH_CHECK("eval_candidates_kernel");
    eval_rescore_kernel<<<(nu + 7) / 8, 256, 0, stream>>>(d_Fu, d_Fi, d, d_users, nu, K, w, d_out_ids, d_out_scores);
    IDG_LAUNCH_CHECK("eval_rescore_kernel");
    return launch_eval_exact(d_Fu, d_Fi, I, d, d_mask_indptr, d_mask_indices, d_users, K, w.flag_cnt, w.flag_list, w.part, d_out_ids, d_out_scores, stream);
}

// Self-check of the tensor-core candidate pass: the raw accumulators (UPPER bounds w = s~ + c|u||i|) of the first 128 users
// against item tile 0 (items 0..127), as the epilogue sees them.  tests/ compare them with the fp64 scores: w >= s must
// hold everywhere and w - s~ must equal the margin product -- the only direct evidence that the ninth k-step is wired.
extern "C" int idg_eval_tc_bounds(const float* d_Fu, const float* d_Fi, int32_t U, int32_t I, int32_t d, const int32_t* d_mask_indptr,
                                  const int32_t* d_mask_indices, const int64_t* d_users, int32_t nu, float* d_out_tile, void* d_ws, void* stream_) {
    if (!d_Fu || !d_Fi || !d_mask_indptr || !d_users || !d_out_tile || !d_ws) return fail(-1, "idg_eval_tc_bounds: null argument%s");
    if ((d != 64 && d != 256) || nu <= 0 || nu > 128 || I <= 0 || U <= 0) return fail(-1, "idg_eval_tc_bounds: needs d = 64 or 256 and 1 <= nu <= 128%s");
    cudaStream_t stream = (cudaStream_t)stream_;
    EvalWs w = eval_carve(d_ws, nu, I, d);
    IDG_CUDA(cudaMemsetAsync(d_ws, 0, 512, stream));
    item_norm_kernel<<<(I + 7) / 8, 256, 0, stream>>>(d_Fi, I, d, w.max_norm, w.tc_aug, w.tc_fir);
    IDG_LAUNCH_CHECK("item_norm_kernel");
    return launch_eval_candidates_tc(d_Fu, w.tc_fir, I, d, d_mask_indptr, d_mask_indices, d_users, nu, 1, w.max_norm, w.flag_cnt, w.flag_list, w.cand_cnt,
                                     w.cand_ids, w.tc_ls, w.tc_li, w.tc_aug, d_out_tile, stream);
}

extern "C" int idg_eval_metrics(const int64_t* d_topk_ids, const int64_t* d_users, int32_t nu, int32_t K, const int32_t* d_test_indptr,
                                const int32_t* d_test_indices, const int32_t* h_ks, int32_t nk, double* d_sums, void* d_ws, void* stream_) {
    if (!d_topk_ids || !d_users || !d_test_indptr || !d_test_indices || !h_ks || !d_sums || !d_ws) return fail(-1, "idg_eval_metrics: null argument%s");
    if (nu <= 0 || K < 1 || nk < 1 || nk > 8) return fail(-1, "idg_eval_metrics: bad sizes (at most 8 cut-offs per call)%s");
    int ks[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < nk; ++t) { if (h_ks[t] < 1 || h_ks[t] > K) return fail(-1, "idg_eval_metrics: k out of range%s"); ks[t] = h_ks[t]; }
    cudaStream_t stream = (cudaStream_t)stream_;
    double* per_user = (double*)d_ws;
    eval_metrics_kernel<<<(nu + 7) / 8, 256, 0, stream>>>(d_topk_ids, d_users, nu, K, d_test_indptr, d_test_indices, nk, ks[0], ks[1], ks[2], ks[3], ks[4], ks[5], ks[6], ks[7], per_user);
    IDG_LAUNCH_CHECK("eval_metrics_kernel");
    ordered_sum_kernel<<<1, 1024, 0, stream>>>(per_user, nu, 3 * nk, d_sums);
    IDG_LAUNCH_CHECK("ordered_sum_kernel");
    return 0;
}
