// Batch x batch losses of the LightGCN-backbone models (SURVEY.md section 8 f, rank 4), forward + backward.
//
// Every one of these reference losses gathers the batch rows of the propagated tables, L2-normalises them
// (F.normalize, eps 1e-12) and reduces an element-wise function of one or two n x n similarity matrices:
//   kind 0  LightCCF  neighbourhood-aggregation loss         models/LightCCF.py:81-94
//   kind 1  LightCSCF margin variant of the same             models/LightCSCF.py:93-104
//   kind 2  SCCF  "down" term  log mean(psi(S) * counts)      models/SCCF.py:72-79
//   kind 3  SCCF  "up"   term  -mean log psi(<a_i,b_i>)       models/SCCF.py:64-70
//   kind 4  DirectAU alignment  mean |a_i - b_i|^2            utility_function/losses.py:61-64
//   kind 5  DirectAU uniformity log mean_{i<j} exp(-2 d_ij^2) utility_function/losses.py:67-69
// a_i = X_i/|X_i|, b_j = Y_j/|Y_j|, S = a b^T, R = a a^T.  With per-row statistics T_i the gradients all have the form
//   dL/dS = H + diag(w),  dL/dR = H   =>   dL/da = H b + (H + H^T) a + w*b,   dL/db = H^T a + w*a
// so the path is: row normalise -> fp32 GEMMs (S, R) -> one row kernel that turns S into H in place (deterministic
// block reductions, fixed order) -> three GEMMs against the [n,d] operands -> normalisation backward.
// n is a mini-batch (<= 4096): the n x n fp32 matrices (<= 64 MB) stay in L2; nothing here touches the [N,d] tables.
// For d = 64 and n >= 256 every contraction runs on the tcgen05 tensor cores as a 3xTF32 split product (the kernels of
// csrc/infonce_tc.cu: TMEM accumulators, error ~2^-21, inside the 1e-5 parity bar): S (+ R) by the scores kernel with a raw
// epilogue, the row kernels write dL/dS directly as (hi, lo) pairs, H (a + b) and H^T a by the [n,n] x [n,64] kernel.
// Other shapes (d != 64, tiny batches) take the fp32 CUDA-core tiles below.
// Duplicate ids in the batch are handled by idg_gather_rows / idg_scatter_add_rows (first occurrence sums all of its
// duplicates in entry order: no float atomics, bit-reproducible).
#include <math.h>

#include "idg_common.cuh"

namespace idg {

// csrc/infonce_tc.cu: 3xTF32 contractions on tcgen05 (n_pad = n rounded up to 128)
int tc_split_rows(const float* A, const float* B, int n, int np, float* AH, float* AL, float* BH, float* BL, cudaStream_t stream);
int tc_scores_raw(const float* XH, const float* XL, const float* YH, const float* YL, int n, int np, float* out, int accumulate, cudaStream_t stream);
int tc_gemm64(const float* EH, const float* EL, const float* TH, const float* TL, int n, int np, float* part, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------------------
// C[M,N] = alpha * op(A) op(B) + beta * C,  row-major, fp32 FMA, k ascending (fixed summation order)
//   op(A)[m,k] = TA ? A[k*lda+m] : A[m*lda+k];   op(B)[k,n] = TB ? B[n*ldb+k] : B[k*ldb+n]
// ------------------------------------------------------------------------------------------------------------
template <bool TA, bool TB, int BM>
__global__ void __launch_bounds__(256) pl_sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                       float* __restrict__ C, int ldc, float alpha, float beta) {
    // BM x 64 output tile per CTA (BM = 64, or 32 for the skinny [n,n] x [n,d] products so that the grid covers the SMs),
    // 16-deep k steps, (BM/16) x 4 micro-tile per thread
    constexpr int RM = BM / 16;
    __shared__ float As[16][BM + 4];
    __shared__ float Bs[16][68];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * 64;
    float acc[RM][4];
#pragma unroll
    for (int a = 0; a < RM; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
        for (int q = 0; q < BM / 16; ++q) {
            const int e = tid + q * 256;
            const int mm = TA ? (e % BM) : (e >> 4), kk = TA ? (e / BM) : (e & 15);
            const int m = m0 + mm, k = k0 + kk;
            float v = 0.f;
            if (m < M && k < K) v = TA ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k];
            As[kk][mm] = v;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = tid + q * 256;
            const int nn = TB ? (e >> 4) : (e & 63), kk = TB ? (e & 15) : (e >> 6);
            const int n = n0 + nn, k = k0 + kk;
            float v = 0.f;
            if (n < N && k < K) v = TB ? B[(size_t)n * ldb + k] : B[(size_t)k * ldb + n];
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float av[RM], bv[4];
#pragma unroll
            for (int a = 0; a < RM; ++a) av[a] = As[kk][ty * RM + a];
#pragma unroll
            for (int b = 0; b < 4; ++b) bv[b] = Bs[kk][tx * 4 + b];
#pragma unroll
            for (int a = 0; a < RM; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < RM; ++a) {
        const int m = m0 + ty * RM + a;
        if (m >= M) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int n = n0 + tx * 4 + b;
            if (n >= N) continue;
            float* c = C + (size_t)m * ldc + n;
            *c = (beta != 0.f) ? fmaf(alpha, acc[a][b], beta * *c) : alpha * acc[a][b];
        }
    }
}

template <int BM>
static void pl_sgemm_launch(bool ta, bool tb, dim3 grid, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                            float alpha, float beta, cudaStream_t st) {
    if (!ta && !tb) pl_sgemm_kernel<false, false, BM><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta);
    else if (!ta && tb) pl_sgemm_kernel<false, true, BM><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta);
    else if (ta && !tb) pl_sgemm_kernel<true, false, BM><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta);
    else pl_sgemm_kernel<true, true, BM><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta);
}

static int pl_sgemm(bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, float alpha,
                    float beta, cudaStream_t st) {
    // 64-row tiles unless that leaves most of the 148 SMs idle (the [n,n] x [n,d] gradient products: N = d or 2d)
    const int tiles64 = ((N + 63) / 64) * ((M + 63) / 64);
    if (tiles64 >= 2 * kNumSMs) pl_sgemm_launch<64>(ta, tb, dim3((N + 63) / 64, (M + 63) / 64), M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, st);
    else pl_sgemm_launch<32>(ta, tb, dim3((N + 63) / 64, (M + 31) / 32), M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, st);
    IDG_LAUNCH_CHECK("pl_sgemm_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// sum over a 256-thread block in a fixed order; every thread gets the result
__device__ __forceinline__ float block_sum256(float v, float* sm) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w];
    return t;
}

// F.normalize(x, dim=-1): x / max(|x|, 1e-12); one warp per row
// (copy, ld_copy): optional second destination with its own leading dimension -- the [Yn | Xn] operand of the merged
// gradient product
__global__ void __launch_bounds__(256) pl_rownorm_kernel(const float* __restrict__ X, int n, int d, float* __restrict__ Xn, float* __restrict__ nrm,
                                                        float* __restrict__ copy, int ld_copy) {
    const int lane = threadIdx.x & 31, i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const float* x = X + (size_t)i * d;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s = fmaf(x[c], x[c], s);
    s = warp_sum(s);
    const float nr = fmaxf(sqrtf(s), 1e-12f);
    for (int c = lane; c < d; c += 32) {
        const float v = x[c] / nr;
        Xn[(size_t)i * d + c] = v;
        if (copy) copy[(size_t)i * ld_copy + c] = v;
    }
    if (lane == 0) nrm[i] = nr;
}

__global__ void pl_diag_kernel(const float* __restrict__ R, int n, float* __restrict__ diag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) diag[i] = R[(size_t)i * n + i];
}

struct PhiVal { float f, df; };
// kinds 0/1: phi(t) = exp(t/tau) [+ exp(relu(t - m)/tau)]
__device__ __forceinline__ PhiVal phi_ccf(float t, float tau, float margin, int use_margin) {
    const float e = expf(t / tau);
    PhiVal r{e, e / tau};
    if (use_margin) {
        const float z = t - margin;
        const float e2 = expf(fmaxf(z, 0.f) / tau);
        r.f += e2;
        if (z > 0.f) r.df += e2 / tau;
    }
    return r;
}
// kinds 2/3: psi(s) = exp(s/tau) + exp(s^2/tau)
__device__ __forceinline__ PhiVal psi_sccf(float s, float tau) {
    const float e1 = expf(s / tau), e2 = expf(s * s / tau);
    return PhiVal{e1 + e2, e1 / tau + 2.f * s * e2 / tau};
}

// kinds 0/1, one CTA per row i:  T_i, loss_i, w_i and S[i,:] <- H[i,:]
__global__ void __launch_bounds__(256) pl_rows_ccf_kernel(float* __restrict__ S, const float* __restrict__ R, int n, float tau, float margin, int use_margin,
                                                         float* __restrict__ w, float* __restrict__ lossi) {
    __shared__ float sm[8];
    const int i = blockIdx.x;
    float* s = S + (size_t)i * n;
    const float* r = R + (size_t)i * n;
    const float sii = s[i];
    float part = 0.f;
    for (int j = threadIdx.x; j < n; j += 256) part += phi_ccf(s[j] + r[j], tau, margin, use_margin).f;
    const float T = block_sum256(part, sm);
    const PhiVal pp = phi_ccf(sii, tau, margin, use_margin);
    const float ratio = pp.f / T;
    const float q = -1.f / ((float)n * (ratio + 1e-5f));          // 10e-6 == 1e-5
    const float hcoef = -q * ratio / T;
    __syncthreads();                                               // every thread holds sii before the row is overwritten
    for (int j = threadIdx.x; j < n; j += 256) s[j] = hcoef * phi_ccf(s[j] + r[j], tau, margin, use_margin).df;
    if (threadIdx.x == 0) {
        w[i] = q * pp.df / T;
        lossi[i] = -logf(ratio + 1e-5f);
    }
}

// kind 2 pass 1: rowT_i = sum_j psi(S_ij)
__global__ void __launch_bounds__(256) pl_rows_sccf_sum_kernel(const float* __restrict__ S, int m, int ld, float tau, float* __restrict__ rowT) {
    __shared__ float sm[8];
    const int i = blockIdx.x;
    const float* s = S + (size_t)i * ld;
    float part = 0.f;
    for (int j = threadIdx.x; j < m; j += 256) part += psi_sccf(s[j], tau).f;
    const float T = block_sum256(part, sm);
    if (threadIdx.x == 0) rowT[i] = T;
}
// kind 2 pass 2: S_ij <- psi'(S_ij) / total
__global__ void __launch_bounds__(256) pl_rows_sccf_grad_kernel(float* __restrict__ S, int64_t total_elems, float tau, const float* __restrict__ scal) {
    const float inv = 1.f / scal[0];
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total_elems; e += (int64_t)gridDim.x * 256) S[e] = psi_sccf(S[e], tau).df * inv;
}

// kind 5: R[i,:] <- E[i,:] = exp(-2 d_ij^2) (0 on the diagonal), rowT_i = sum_j E_ij
__global__ void __launch_bounds__(256) pl_rows_uniform_kernel(float* __restrict__ R, const float* __restrict__ diag, int n, float* __restrict__ rowT) {
    __shared__ float sm[8];
    const int i = blockIdx.x;
    float* r = R + (size_t)i * n;
    const float dii = diag[i];
    float part = 0.f;
    for (int j = threadIdx.x; j < n; j += 256) {
        const float d2 = fmaxf(dii + diag[j] - 2.f * r[j], 0.f);
        const float e = (j == i) ? 0.f : expf(-2.f * d2);
        r[j] = e;
        part += e;
    }
    const float T = block_sum256(part, sm);
    if (threadIdx.x == 0) rowT[i] = T;
}

// kinds 3/4 (element-wise over matching rows), one warp per row
__global__ void __launch_bounds__(256) pl_rows_pairwise_kernel(int kind, const float* __restrict__ Xn, const float* __restrict__ Yn, int n, int d, float tau,
                                                              float* __restrict__ w, float* __restrict__ lossi) {
    const int lane = threadIdx.x & 31, i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const float* a = Xn + (size_t)i * d;
    const float* b = Yn + (size_t)i * d;
    float s = 0.f;
    if (kind == 3) {
        for (int c = lane; c < d; c += 32) s = fmaf(a[c], b[c], s);
        s = warp_sum(s);
        const PhiVal p = psi_sccf(s, tau);
        if (lane == 0) { lossi[i] = logf(p.f); w[i] = -(p.df / p.f) / (float)n; }
    } else {
        for (int c = lane; c < d; c += 32) { const float t = a[c] - b[c]; s = fmaf(t, t, s); }
        s = warp_sum(s);
        if (lane == 0) { lossi[i] = s; w[i] = 0.f; }
    }
}

// ordered sum of v[0..n) (single CTA) and the scalar epilogue of each kind
// *d_loss += scale * loss (callers zero it first); cnt_a/cnt_b (optional, kind 2): the denominator is read from the device
__global__ void __launch_bounds__(1024) pl_reduce_kernel(int kind, const float* __restrict__ v, int n, float denom, float* __restrict__ scal, float* __restrict__ d_loss,
                                                        float scale, const int* __restrict__ cnt_a, const int* __restrict__ cnt_b) {
    __shared__ float sm[32];
    float part = 0.f;
    for (int j = threadIdx.x; j < n; j += 1024) part += v[j];
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int q = 0; q < 32; ++q) t += sm[q];
        scal[0] = t;
        float out;
        if (kind == 0 || kind == 1 || kind == 4) out = t / (float)n;
        else if (kind == 3) out = -t / (float)n;
        else {                            // kind 2: denom = n_unique_users * n_unique_items; kind 5: n (n - 1)
            if (kind == 2 && cnt_a && cnt_b) denom = (float)(*cnt_a) * (float)(*cnt_b);
            out = logf(t / denom);
        }
        *d_loss += scale * out;
    }
}

// gradients w.r.t. the normalised rows assembled per kind, then F.normalize backward; one warp per row
__global__ void __launch_bounds__(256) pl_finish_kernel(int kind, const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ Xn,
                                                       const float* __restrict__ Yn, const float* __restrict__ nx, const float* __restrict__ ny,
                                                       const float* __restrict__ w, const float* __restrict__ rowT, const float* __restrict__ scal,
                                                       const float* __restrict__ tA, int fold_tA, const float* __restrict__ tB, int n, int d,
                                                       float scale, float* __restrict__ gX, float* __restrict__ gY) {
    const int lane = threadIdx.x & 31, i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const size_t o = (size_t)i * d;
    // fold_tA: tA is [n, 2d] = H [Yn | Xn]; the two halves are summed here (H b + H a)
    const size_t oa = fold_tA ? (size_t)i * 2 * d : o;
    const int hi = fold_tA ? d : 0;
    const float wi = w ? w[i] : 0.f;
    const float cu = (kind == 5) ? -8.f / scal[0] : 0.f;
    const float ri = (kind == 5) ? rowT[i] : 0.f;
    float da = 0.f, db = 0.f;
    // pass 1: <a, ga>, <b, gb>
    for (int c = lane; c < d; c += 32) {
        const float a = Xn[o + c], b = Yn ? Yn[o + c] : 0.f;
        float ga, gb = 0.f;
        if (kind <= 1) { ga = (tA[oa + c] + (fold_tA ? tA[oa + hi + c] : 0.f)) + tB[o + c] + wi * b; gb = tB[o + c] + wi * a; }
        else if (kind == 2) { ga = tA[o + c]; gb = tB[o + c]; }
        else if (kind == 3) { ga = wi * b; gb = wi * a; }
        else if (kind == 4) { ga = 2.f * (a - b) / (float)n; gb = -ga; }
        else { ga = cu * (ri * a - tA[o + c]); }
        da = fmaf(a, ga, da);
        db = fmaf(b, gb, db);
    }
    da = warp_sum(da);
    db = warp_sum(db);
    const float na = nx[i], nb = ny ? ny[i] : 1.f;
    for (int c = lane; c < d; c += 32) {
        const float a = Xn[o + c], b = Yn ? Yn[o + c] : 0.f;
        float ga, gb = 0.f;
        if (kind <= 1) { ga = (tA[oa + c] + (fold_tA ? tA[oa + hi + c] : 0.f)) + tB[o + c] + wi * b; gb = tB[o + c] + wi * a; }
        else if (kind == 2) { ga = tA[o + c]; gb = tB[o + c]; }
        else if (kind == 3) { ga = wi * b; gb = wi * a; }
        else if (kind == 4) { ga = 2.f * (a - b) / (float)n; gb = -ga; }
        else { ga = cu * (ri * a - tA[o + c]); }
        // d/dx of x / max(|x|, eps): (g - a <a,g>) / |x| above the clamp, g / eps below it
        gX[o + c] = scale * ((na > 1e-12f) ? (ga - a * da) / na : ga / 1e-12f);
        if (gY) gY[o + c] = scale * ((nb > 1e-12f) ? (gb - b * db) / nb : gb / 1e-12f);
    }
    (void)X; (void)Y;
}

// rows of a table into a dense [n,d] block / deterministic scatter-add of a dense block back into table rows
__global__ void __launch_bounds__(256) pl_gather_kernel(const float* __restrict__ T, const int64_t* __restrict__ idx, int n, int d, float* __restrict__ out) {
    const int lane = threadIdx.x & 31, i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const float* src = T + (size_t)idx[i] * d;
    for (int c = lane; c < d; c += 32) out[(size_t)i * d + c] = src[c];
}

// One warp per entry i; the ids sit in shared memory as int32 (padded with -1 to a multiple of 128) and are scanned 4 per lane
// and step.  An entry with an earlier duplicate returns; the first occurrence sums all of its duplicates in ascending entry
// order (d <= 256: up to 8 floats per lane) and adds the row to the table.
__global__ void __launch_bounds__(256) pl_scatter_add_kernel(const float* __restrict__ G, const int64_t* __restrict__ idx, int n, int d, float* __restrict__ T) {
    extern __shared__ __align__(16) int pl_skeys[];
    const int npad = (n + 127) & ~127;
    for (int i = threadIdx.x; i < npad; i += blockDim.x) pl_skeys[i] = (i < n) ? (int)idx[i] : -1;
    __syncthreads();
    const int lane = threadIdx.x & 31, i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const int me = pl_skeys[i];
    const int4* k4 = reinterpret_cast<const int4*>(pl_skeys);
    for (int base = 0; base < i; base += 128) {
        const int j = base + lane * 4;
        const int4 k = k4[j >> 2];
        const bool hit = ((j < i) && (k.x == me)) || ((j + 1 < i) && (k.y == me)) || ((j + 2 < i) && (k.z == me)) || ((j + 3 < i) && (k.w == me));
        if (__any_sync(0xffffffffu, hit)) return;
    }
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    for (int base = i & ~127; base < n; base += 128) {
        const int j = base + lane * 4;
        const int4 k = k4[j >> 2];
        unsigned m[4];
        m[0] = __ballot_sync(0xffffffffu, (j >= i) && (k.x == me));
        m[1] = __ballot_sync(0xffffffffu, (j + 1 >= i) && (k.y == me));
        m[2] = __ballot_sync(0xffffffffu, (j + 2 >= i) && (k.z == me));
        m[3] = __ballot_sync(0xffffffffu, (j + 3 >= i) && (k.w == me));
        unsigned any = m[0] | m[1] | m[2] | m[3];
        while (any) {
            const int L = __ffs(any) - 1;
            any &= any - 1;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                if (!((m[q4] >> L) & 1u)) continue;
                const float* g = G + (size_t)(base + L * 4 + q4) * d;
#pragma unroll
                for (int q = 0; q < 8; ++q) { const int c = q * 32 + lane; if (c < d) acc[q] += g[c]; }
            }
        }
    }
    float* dst = T + (size_t)(long long)me * d;
#pragma unroll
    for (int q = 0; q < 8; ++q) { const int c = q * 32 + lane; if (c < d) dst[c] += acc[q]; }
}

// ---- tensor-core path helpers (d = 64) -------------------------------------------------------------------------------
__device__ __forceinline__ void pl_split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

// diag[i] = <a_i, b_i> in fp32 (the positive score S_ii / the squared norm R_ii); one warp per row
__global__ void __launch_bounds__(256) pl_diag_dot_kernel(const float* __restrict__ Xn, const float* __restrict__ Yn, int n, int d, float* __restrict__ diag) {
    const int lane = threadIdx.x & 31, i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s = fmaf(Xn[(size_t)i * d + c], Yn[(size_t)i * d + c], s);
    s = warp_sum(s);
    if (lane == 0) diag[i] = s;
}

// kinds 0/1, one CTA per row i < n of T = S + R (ld np): statistics as pl_rows_ccf_kernel, then T[i,:] <- hi(H[i,:]),
// HL[i,:] <- lo(H[i,:]) with zeros in the padded columns [n, np)
__global__ void __launch_bounds__(256) pl_rows_ccf_tc_kernel(float* __restrict__ T, float* __restrict__ HL, const float* __restrict__ sdiag, int n, int np,
                                                            float tau, float margin, int use_margin, float* __restrict__ w, float* __restrict__ lossi) {
    __shared__ float sm[8];
    const int i = blockIdx.x;
    float* t = T + (size_t)i * np;
    float* hl = HL + (size_t)i * np;
    float part = 0.f;
    for (int j = threadIdx.x; j < n; j += 256) part += phi_ccf(t[j], tau, margin, use_margin).f;
    const float Tsum = block_sum256(part, sm);
    const PhiVal pp = phi_ccf(sdiag[i], tau, margin, use_margin);
    const float ratio = pp.f / Tsum;
    const float q = -1.f / ((float)n * (ratio + 1e-5f));
    const float hcoef = -q * ratio / Tsum;
    for (int j = threadIdx.x; j < np; j += 256) {
        float hi = 0.f, lo = 0.f;
        if (j < n) pl_split_tf32(hcoef * phi_ccf(t[j], tau, margin, use_margin).df, hi, lo);
        t[j] = hi; hl[j] = lo;
    }
    if (threadIdx.x == 0) {
        w[i] = q * pp.df / Tsum;
        lossi[i] = -logf(ratio + 1e-5f);
    }
}

// kind 2 pass 2 on the padded matrix: S[i,:] <- hi(psi'(S_ij)/total), HL <- lo, zeros in the padding; one CTA per row
__global__ void __launch_bounds__(256) pl_rows_sccf_grad_tc_kernel(float* __restrict__ S, float* __restrict__ HL, int n, int np, float tau,
                                                                  const float* __restrict__ scal) {
    const int i = blockIdx.x;
    const float inv = 1.f / scal[0];
    float* s = S + (size_t)i * np;
    float* hl = HL + (size_t)i * np;
    for (int j = threadIdx.x; j < np; j += 256) {
        float hi = 0.f, lo = 0.f;
        if (j < n) pl_split_tf32(psi_sccf(s[j], tau).df * inv, hi, lo);
        s[j] = hi; hl[j] = lo;
    }
}

// kind 5 on the padded matrix: R[i,:] <- hi(E[i,:]), HL <- lo, rowT_i = sum_j E_ij (E_ij = exp(-2 d_ij^2), 0 on the diagonal)
__global__ void __launch_bounds__(256) pl_rows_uniform_tc_kernel(float* __restrict__ R, float* __restrict__ HL, const float* __restrict__ diag, int n, int np,
                                                                float* __restrict__ rowT) {
    __shared__ float sm[8];
    const int i = blockIdx.x;
    float* r = R + (size_t)i * np;
    float* hl = HL + (size_t)i * np;
    const float dii = diag[i];
    float part = 0.f;
    for (int j = threadIdx.x; j < np; j += 256) {
        float e = 0.f;
        if (j < n && j != i) e = expf(-2.f * fmaxf(dii + diag[j] - 2.f * r[j], 0.f));
        float hi, lo;
        pl_split_tf32(e, hi, lo);
        r[j] = hi; hl[j] = lo;
        part += e;
    }
    const float Tsum = block_sum256(part, sm);
    if (threadIdx.x == 0) rowT[i] = Tsum;
}

// (OH, OL)[c][r] = (IH, IL)[r][c] on the n x n block, zero elsewhere in [np, np]; 32 x 32 tiles
__global__ void __launch_bounds__(256) pl_transpose_split_kernel(const float* __restrict__ IH, const float* __restrict__ IL, int n, int np,
                                                                float* __restrict__ OH, float* __restrict__ OL) {
    __shared__ float th[32][33], tl[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int q = ty; q < 32; q += 8) {
        const int r = r0 + q, c = c0 + tx;
        const bool ok = r < n && c < n;
        th[q][tx] = ok ? IH[(size_t)r * np + c] : 0.f;
        tl[q][tx] = ok ? IL[(size_t)r * np + c] : 0.f;
    }
    __syncthreads();
    for (int q = ty; q < 32; q += 8) {
        const int c = c0 + q, r = r0 + tx;
        OH[(size_t)c * np + r] = th[tx][q];
        OL[(size_t)c * np + r] = tl[tx][q];
    }
}

// K-major right operand of the [n,n] x [n,64] product: (TH, TL)[k][c] = split(Y[c][k] (+ Y2[c][k])) for c < n, 0 beyond
__global__ void __launch_bounds__(256) pl_operand_t_kernel(const float* __restrict__ Y, const float* __restrict__ Y2, int n, int np,
                                                          float* __restrict__ TH, float* __restrict__ TL) {
    __shared__ float tile[32][65];
    const int c0 = blockIdx.x * 32;
    for (int q = threadIdx.x; q < 32 * 64; q += 256) {
        const int c = q >> 6, k = q & 63;
        float v = 0.f;
        if (c0 + c < n) { v = Y[(size_t)(c0 + c) * 64 + k]; if (Y2) v += Y2[(size_t)(c0 + c) * 64 + k]; }
        tile[c][k] = v;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < 32 * 64; q += 256) {
        const int k = q >> 5, c = q & 31;
        float hi, lo;
        pl_split_tf32(tile[c][k], hi, lo);
        TH[(size_t)k * np + c0 + c] = hi;
        TL[(size_t)k * np + c0 + c] = lo;
    }
}

// out[i][0..63] = sum over the 4 column splits (in split order) of part[s][i][0..63]
__global__ void __launch_bounds__(256) pl_fold_parts_kernel(const float* __restrict__ part, int n, int np, float* __restrict__ out) {
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= n * 64) return;
    const int i = e >> 6, c = e & 63;
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) s += part[((size_t)q * np + i) * 64 + c];
    out[e] = s;
}

constexpr int kTcMinRows = 256;
__host__ inline bool pl_use_tc(int n, int d) { return d == 64 && n >= kTcMinRows; }
__host__ inline int pl_np(int n) { return (n + 127) / 128 * 128; }

struct PlWs {
    float *Xn, *Yn, *YX, *nx, *ny, *S, *R, *w, *rowT, *lossi, *diag, *tA, *tB, *scal;
    // tensor-core path (d = 64, n >= 256): operand splits [np,64] x 4, K-major operands [64,np] x 4, H^T splits [np,np] x 2 (S and R
    // above are [np,np] there and end up holding hi(H) / lo(H)), split-wise partial products [4,np,64] x 2
    float *XH, *XL, *YH, *YL, *T1H, *T1L, *T2H, *T2L, *HtH, *HtL, *partA, *partB;
};
__host__ inline size_t pl_align(size_t x) { return (x + 255) & ~(size_t)255; }
__host__ inline size_t pl_carve(void* ws, int n, int d, PlWs* w) {
    char* p = (char*)ws;
    auto take = [&](size_t bytes) { char* q = p; p += pl_align(bytes); return (float*)q; };
    const size_t nd = sizeof(float) * (size_t)n * d, nn = sizeof(float) * (size_t)n * n, n1 = sizeof(float) * (size_t)n;
    PlWs t;
    t.Xn = take(nd); t.Yn = take(nd); t.YX = take(2 * nd); t.tA = take(2 * nd); t.tB = take(nd);
    t.nx = take(n1); t.ny = take(n1); t.w = take(n1); t.rowT = take(n1); t.lossi = take(n1); t.diag = take(n1);
    t.scal = take(256);
    if (pl_use_tc(n, d)) {
        const size_t np = (size_t)pl_np(n), pd = sizeof(float) * np * 64, pp = sizeof(float) * np * np;
        t.XH = take(pd); t.XL = take(pd); t.YH = take(pd); t.YL = take(pd);
        t.T1H = take(pd); t.T1L = take(pd); t.T2H = take(pd); t.T2L = take(pd);
        t.partA = take(4 * pd); t.partB = take(4 * pd);
        t.S = take(pp + 4096); t.R = take(pp + 4096); t.HtH = take(pp + 4096); t.HtL = take(pp + 4096);
        t.S = (float*)(((uintptr_t)t.S + 1023) & ~(uintptr_t)1023); t.R = (float*)(((uintptr_t)t.R + 1023) & ~(uintptr_t)1023);
        t.HtH = (float*)(((uintptr_t)t.HtH + 1023) & ~(uintptr_t)1023); t.HtL = (float*)(((uintptr_t)t.HtL + 1023) & ~(uintptr_t)1023);
    } else {
        t.XH = t.XL = t.YH = t.YL = t.T1H = t.T1L = t.T2H = t.T2L = t.HtH = t.HtL = t.partA = t.partB = nullptr;
        t.S = take(nn); t.R = take(nn);
    }
    if (w) *w = t;
    return (size_t)(p - (char*)ws);
}

}  // namespace idg

using namespace idg;

extern "C" int64_t idg_pair_loss_workspace_bytes(int32_t n, int32_t d) {
    if (n <= 0 || d <= 0) return 0;
    return (int64_t)pl_carve(nullptr, n, d, nullptr) + 256;
}

static int pair_loss_impl(int32_t kind, const float* d_X, const float* d_Y, int32_t n, int32_t d, float p0, float p1, float scale,
                          const int32_t* d_cnt_a, const int32_t* d_cnt_b, float* d_loss, float* d_gX, float* d_gY, void* d_ws, void* stream_) {
    if (kind < 0 || kind > 5) return fail(-1, "idg_pair_loss: kind must be 0..5%s");
    if (!d_X || !d_loss || !d_ws || n <= 0 || d <= 0 || n > 8192) return fail(-1, "idg_pair_loss: bad argument (n in 1..8192)%s");
    if (kind != 5 && !d_Y) return fail(-1, "idg_pair_loss: kind %s needs both operands", "0..4");
    if (kind == 5 && n < 2) return fail(-1, "idg_pair_loss: uniformity needs n >= 2%s");
    if (kind <= 3 && !(p0 > 0.f)) return fail(-1, "idg_pair_loss: temperature must be > 0%s");
    if (kind == 2 && !(p1 > 0.f) && !(d_cnt_a && d_cnt_b))
        return fail(-1, "idg_pair_loss: SCCF down term needs p1 = n_unique_users * n_unique_items > 0 (or the two device counts)%s");
    cudaStream_t st = (cudaStream_t)stream_;
    PlWs w;
    pl_carve((void*)(((uintptr_t)d_ws + 255) & ~(uintptr_t)255), n, d, &w);
    const bool want_grad = d_gX != nullptr;
    const int rb = (n + 7) / 8;
    const bool merged = want_grad && kind <= 1 && !pl_use_tc(n, d);   // fp32 path: gradient product against [Yn | Xn] in one pass over H
    pl_rownorm_kernel<<<rb, 256, 0, st>>>(d_X, n, d, w.Xn, w.nx, merged ? w.YX + d : nullptr, 2 * d);
    IDG_LAUNCH_CHECK("pl_rownorm_kernel");
    if (kind != 5) {
        pl_rownorm_kernel<<<rb, 256, 0, st>>>(d_Y, n, d, w.Yn, w.ny, merged ? w.YX : nullptr, 2 * d);
        IDG_LAUNCH_CHECK("pl_rownorm_kernel");
    }
    int rc;
    if (pl_use_tc(n, d) && kind != 3 && kind != 4) {
        // ---- tcgen05 path: every n x n x 64 contraction as a 3xTF32 split product --------------------------------------
        const int np = pl_np(n);
        const dim3 tgrid(np / 32, np / 32);
        if ((rc = tc_split_rows(w.Xn, kind == 5 ? w.Xn : w.Yn, n, np, w.XH, w.XL, w.YH, w.YL, st))) return rc;
        if (kind <= 1) {
            if ((rc = tc_scores_raw(w.XH, w.XL, w.YH, w.YL, n, np, w.S, 0, st))) return rc;       // S = a b^T
            if ((rc = tc_scores_raw(w.XH, w.XL, w.XH, w.XL, n, np, w.S, 1, st))) return rc;       // S += a a^T
            pl_diag_dot_kernel<<<rb, 256, 0, st>>>(w.Xn, w.Yn, n, d, w.diag);
            IDG_LAUNCH_CHECK("pl_diag_dot_kernel");
            pl_rows_ccf_tc_kernel<<<n, 256, 0, st>>>(w.S, w.R, w.diag, n, np, p0, p1, kind == 1, w.w, w.lossi);
            IDG_LAUNCH_CHECK("pl_rows_ccf_tc_kernel");
            pl_reduce_kernel<<<1, 1024, 0, st>>>(kind, w.lossi, n, 1.f, w.scal, d_loss, scale, d_cnt_a, d_cnt_b);
            IDG_LAUNCH_CHECK("pl_reduce_kernel");
        } else if (kind == 2) {
            if ((rc = tc_scores_raw(w.XH, w.XL, w.YH, w.YL, n, np, w.S, 0, st))) return rc;
            pl_rows_sccf_sum_kernel<<<n, 256, 0, st>>>(w.S, n, np, p0, w.rowT);   // n columns, row stride np
            IDG_LAUNCH_CHECK("pl_rows_sccf_sum_kernel");
            pl_reduce_kernel<<<1, 1024, 0, st>>>(kind, w.rowT, n, p1, w.scal, d_loss, scale, d_cnt_a, d_cnt_b);
            IDG_LAUNCH_CHECK("pl_reduce_kernel");
            if (want_grad) {
                pl_rows_sccf_grad_tc_kernel<<<n, 256, 0, st>>>(w.S, w.R, n, np, p0, w.scal);
                IDG_LAUNCH_CHECK("pl_rows_sccf_grad_tc_kernel");
            }
        } else {
            if ((rc = tc_scores_raw(w.XH, w.XL, w.XH, w.XL, n, np, w.S, 0, st))) return rc;       // R = a a^T
            pl_diag_dot_kernel<<<rb, 256, 0, st>>>(w.Xn, w.Xn, n, d, w.diag);
            IDG_LAUNCH_CHECK("pl_diag_dot_kernel");
            pl_rows_uniform_tc_kernel<<<n, 256, 0, st>>>(w.S, w.R, w.diag, n, np, w.rowT);
            IDG_LAUNCH_CHECK("pl_rows_uniform_tc_kernel");
            pl_reduce_kernel<<<1, 1024, 0, st>>>(kind, w.rowT, n, (float)n * (float)(n - 1), w.scal, d_loss, scale, d_cnt_a, d_cnt_b);
            IDG_LAUNCH_CHECK("pl_reduce_kernel");
        }
        if (want_grad) {
            if (kind != 5 && !d_gY) return fail(-1, "idg_pair_loss: d_gY is required with d_gX for kind %s", "0..4");
            // (w.S, w.R) now hold (hi, lo) of H.  tA = H (a + b) [kinds 0/1], H b [kind 2], E a [kind 5];  tB = H^T a
            pl_operand_t_kernel<<<np / 32, 256, 0, st>>>(kind == 5 ? w.Xn : w.Yn, kind <= 1 ? w.Xn : nullptr, n, np, w.T1H, w.T1L);
            IDG_LAUNCH_CHECK("pl_operand_t_kernel");
            if ((rc = tc_gemm64(w.S, w.R, w.T1H, w.T1L, n, np, w.partA, st))) return rc;
            pl_fold_parts_kernel<<<(n * 64 + 255) / 256, 256, 0, st>>>(w.partA, n, np, w.tA);
            IDG_LAUNCH_CHECK("pl_fold_parts_kernel");
            if (kind != 5) {
                pl_transpose_split_kernel<<<tgrid, 256, 0, st>>>(w.S, w.R, n, np, w.HtH, w.HtL);
                IDG_LAUNCH_CHECK("pl_transpose_split_kernel");
                pl_operand_t_kernel<<<np / 32, 256, 0, st>>>(w.Xn, nullptr, n, np, w.T2H, w.T2L);
                IDG_LAUNCH_CHECK("pl_operand_t_kernel");
                if ((rc = tc_gemm64(w.HtH, w.HtL, w.T2H, w.T2L, n, np, w.partB, st))) return rc;
                pl_fold_parts_kernel<<<(n * 64 + 255) / 256, 256, 0, st>>>(w.partB, n, np, w.tB);
                IDG_LAUNCH_CHECK("pl_fold_parts_kernel");
            }
            pl_finish_kernel<<<rb, 256, 0, st>>>(kind, d_X, d_Y, w.Xn, kind == 5 ? nullptr : w.Yn, w.nx, kind == 5 ? nullptr : w.ny,
                                                kind <= 1 ? w.w : nullptr, w.rowT, w.scal, w.tA, 0, w.tB, n, d, scale, d_gX, kind == 5 ? nullptr : d_gY);
            IDG_LAUNCH_CHECK("pl_finish_kernel");
        }
        return 0;
    }
    if (kind <= 1) {
        if ((rc = pl_sgemm(false, true, n, n, d, w.Xn, d, w.Yn, d, w.S, n, 1.f, 0.f, st))) return rc;
        if ((rc = pl_sgemm(false, true, n, n, d, w.Xn, d, w.Xn, d, w.R, n, 1.f, 0.f, st))) return rc;
        pl_rows_ccf_kernel<<<n, 256, 0, st>>>(w.S, w.R, n, p0, p1, kind == 1, w.w, w.lossi);
        IDG_LAUNCH_CHECK("pl_rows_ccf_kernel");
        pl_reduce_kernel<<<1, 1024, 0, st>>>(kind, w.lossi, n, 1.f, w.scal, d_loss, scale, d_cnt_a, d_cnt_b);
        IDG_LAUNCH_CHECK("pl_reduce_kernel");
        if (want_grad) {
            // tA [n,2d] = H [b | a] (folded to H b + H a by the finish kernel),  tB = H^T a
            if ((rc = pl_sgemm(false, false, n, 2 * d, n, w.S, n, w.YX, 2 * d, w.tA, 2 * d, 1.f, 0.f, st))) return rc;
            if ((rc = pl_sgemm(true, false, n, d, n, w.S, n, w.Xn, d, w.tB, d, 1.f, 0.f, st))) return rc;
        }
    } else if (kind == 2) {
        if ((rc = pl_sgemm(false, true, n, n, d, w.Xn, d, w.Yn, d, w.S, n, 1.f, 0.f, st))) return rc;
        pl_rows_sccf_sum_kernel<<<n, 256, 0, st>>>(w.S, n, n, p0, w.rowT);
        IDG_LAUNCH_CHECK("pl_rows_sccf_sum_kernel");
        pl_reduce_kernel<<<1, 1024, 0, st>>>(kind, w.rowT, n, p1, w.scal, d_loss, scale, d_cnt_a, d_cnt_b);
        IDG_LAUNCH_CHECK("pl_reduce_kernel");
        if (want_grad) {
            pl_rows_sccf_grad_kernel<<<kNumSMs * 4, 256, 0, st>>>(w.S, (int64_t)n * n, p0, w.scal);
            IDG_LAUNCH_CHECK("pl_rows_sccf_grad_kernel");
            if ((rc = pl_sgemm(false, false, n, d, n, w.S, n, w.Yn, d, w.tA, d, 1.f, 0.f, st))) return rc;
            if ((rc = pl_sgemm(true, false, n, d, n, w.S, n, w.Xn, d, w.tB, d, 1.f, 0.f, st))) return rc;
        }
    } else if (kind == 3 || kind == 4) {
        pl_rows_pairwise_kernel<<<rb, 256, 0, st>>>(kind, w.Xn, w.Yn, n, d, p0, w.w, w.lossi);
        IDG_LAUNCH_CHECK("pl_rows_pairwise_kernel");
        pl_reduce_kernel<<<1, 1024, 0, st>>>(kind, w.lossi, n, 1.f, w.scal, d_loss, scale, d_cnt_a, d_cnt_b);
        IDG_LAUNCH_CHECK("pl_reduce_kernel");
    } else {
        if ((rc = pl_sgemm(false, true, n, n, d, w.Xn, d, w.Xn, d, w.R, n, 1.f, 0.f, st))) return rc;
        pl_diag_kernel<<<(n + 255) / 256, 256, 0, st>>>(w.R, n, w.diag);
        IDG_LAUNCH_CHECK("pl_diag_kernel");
        pl_rows_uniform_kernel<<<n, 256, 0, st>>>(w.R, w.diag, n, w.rowT);
        IDG_LAUNCH_CHECK("pl_rows_uniform_kernel");
        pl_reduce_kernel<<<1, 1024, 0, st>>>(kind, w.rowT, n, (float)n * (float)(n - 1), w.scal, d_loss, scale, d_cnt_a, d_cnt_b);
        IDG_LAUNCH_CHECK("pl_reduce_kernel");
        if (want_grad) {
            if ((rc = pl_sgemm(false, false, n, d, n, w.R, n, w.Xn, d, w.tA, d, 1.f, 0.f, st))) return rc;
        }
    }
    if (want_grad) {
        if (kind != 5 && !d_gY) return fail(-1, "idg_pair_loss: d_gY is required with d_gX for kind %s", "0..4");
        pl_finish_kernel<<<rb, 256, 0, st>>>(kind, d_X, d_Y, w.Xn, kind == 5 ? nullptr : w.Yn, w.nx, kind == 5 ? nullptr : w.ny,
                                            (kind <= 1 || kind == 3) ? w.w : nullptr, w.rowT, w.scal, w.tA, merged ? 1 : 0, w.tB, n, d, scale, d_gX,
                                            kind == 5 ? nullptr : d_gY);
        IDG_LAUNCH_CHECK("pl_finish_kernel");
    }
    return 0;
}

extern "C" int idg_pair_loss(int32_t kind, const float* d_X, const float* d_Y, int32_t n, int32_t d, float p0, float p1, float* d_loss,
                             float* d_gX, float* d_gY, void* d_ws, void* stream) {
    if (!d_loss) return fail(-1, "idg_pair_loss: null d_loss%s");
    IDG_CUDA(cudaMemsetAsync(d_loss, 0, sizeof(float), (cudaStream_t)stream));
    return pair_loss_impl(kind, d_X, d_Y, n, d, p0, p1, 1.f, nullptr, nullptr, d_loss, d_gX, d_gY, d_ws, stream);
}

extern "C" int idg_pair_loss_ex(int32_t kind, const float* d_X, const float* d_Y, int32_t n, int32_t d, float p0, float p1, float scale,
                                const int32_t* d_cnt_a, const int32_t* d_cnt_b, float* d_loss, float* d_gX, float* d_gY, void* d_ws,
                                void* stream) {
    return pair_loss_impl(kind, d_X, d_Y, n, d, p0, p1, scale, d_cnt_a, d_cnt_b, d_loss, d_gX, d_gY, d_ws, stream);
}

extern "C" int idg_gather_rows(const float* d_T, const int64_t* d_idx, int32_t n, int32_t d, float* d_out, void* stream) {
    if (!d_T || !d_idx || !d_out || n <= 0 || d <= 0) return fail(-1, "idg_gather_rows: bad argument%s");
    pl_gather_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)stream>>>(d_T, d_idx, n, d, d_out);
    IDG_LAUNCH_CHECK("pl_gather_kernel");
    return 0;
}

extern "C" int idg_scatter_add_rows(const float* d_G, const int64_t* d_idx, int32_t n, int32_t d, float* d_T, void* stream) {
    if (!d_G || !d_idx || !d_T || n <= 0 || d <= 0 || d > 256 || n > 8192) return fail(-1, "idg_scatter_add_rows: bad argument (d <= 256, n <= 8192)%s");
    const size_t smem = sizeof(int) * (size_t)((n + 127) & ~127);
    pl_scatter_add_kernel<<<(n + 7) / 8, 256, smem, (cudaStream_t)stream>>>(d_G, d_idx, n, d, d_T);
    IDG_LAUNCH_CHECK("pl_scatter_add_kernel");
    return 0;
}
