// In-batch InfoNCE, forward + backward fused (sm_100a).
//
// Replaces utility_function/losses.py:24-35 (normalize x2, matmul n x n, exp, sum, log, mean) and
// its autograd backward for the call sites models/SimGCL.py:83-84 and XSimGCL.py:88-89.
//   a_i = V1[idx_i]/|.|, b_j = V2[idx_j]/|.|, S_ij = <a_i,b_j>/tau
//   loss = mean_i -log(r_i + 1e-5),  r_i = exp(S_ii) / sum_j exp(S_ij)        (10e-6 == 1e-5)
//   w_i  = -r_i / (n (r_i + 1e-5)),  p_ij = exp(S_ij)/ttl_i
//   dL/da_i = (w_i/tau) (b_i - sum_j p_ij b_j);   dL/db_j = (1/tau) (w_j a_j - sum_i w_i p_ij a_i)
//   dL/dv   = (g - a <a,g>) / |v|
// The n x n matrix is never written: each pass recomputes a 64x64 tile of exp(S) in shared memory
// and contracts it immediately.  Column splits write partials that are summed in split order.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "idg_common.cuh"

namespace idg {

// csrc/infonce_tc.cu: the four n x n x 64 contractions on tcgen05 tensor cores (3xTF32)
size_t nce_tc_extra_bytes(int n_max);
int nce_tc_stage(int stage, const float* A, const float* Bm, const float* beta, const int* d_n, int n_max, float inv_tau, float* part_sum,
                 float* part_pb, float* part_qa, void* extra, cudaStream_t stream, int want_grad, int splits);

constexpr int kNT = 64;      // tile edge
constexpr int kNceSplits = 8;
// csrc/infonce_flash.cu: the same contractions without the n x n matrices
int nce_flash(const float* X, const float* Y, const float* V, const float* v_scale, const int* d_n, int n_max, float inv_tau, float* part_sum,
              float* part_out, int splits, cudaStream_t stream);

struct NceWs {
    float* A;      // [n,d] normalised view 1 rows
    float* Bm;     // [n,d] normalised view 2 rows
    float* na;     // [n] |v1|
    float* nb;     // [n] |v2|
    float* pos;    // [n] exp(S_ii)
    float* inv_ttl;  // [n]
    float* wrow;   // [n] w_i
    float* beta;   // [n] w_i / ttl_i
    float* loss_i; // [n]
    float* part_sum;  // [splits, n]
    float* part_pb;   // [splits, n, d]  sum_j exp(S_ij) b_j
    float* part_qa;   // [splits, n, d]  sum_i beta_i exp(S_ij) a_i
    char* extra;      // tensor-core path scratch (splits of the operands, E / E^T)
};

__host__ __device__ inline size_t nce_align(size_t x) { return (x + 255) & ~(size_t)255; }

__host__ inline NceWs nce_carve(void* ws, int n, int d) {
    char* p = (char*)ws;
    NceWs w;
    auto take = [&](size_t bytes) { char* q = p; p += nce_align(bytes); return q; };
    w.A = (float*)take(sizeof(float) * (size_t)n * d);
    w.Bm = (float*)take(sizeof(float) * (size_t)n * d);
    w.na = (float*)take(sizeof(float) * n); w.nb = (float*)take(sizeof(float) * n);
    w.pos = (float*)take(sizeof(float) * n); w.inv_ttl = (float*)take(sizeof(float) * n);
    w.wrow = (float*)take(sizeof(float) * n); w.beta = (float*)take(sizeof(float) * n);
    w.loss_i = (float*)take(sizeof(float) * n);
    w.part_sum = (float*)take(sizeof(float) * (size_t)kNceSplits * n);
    w.part_pb = (float*)take(sizeof(float) * (size_t)kNceSplits * n * d);
    w.part_qa = (float*)take(sizeof(float) * (size_t)kNceSplits * n * d);
    w.extra = p;
    return w;
}

// gather + F.normalize (eps 1e-12) + diagonal term; one warp per row, d = 64
__global__ void __launch_bounds__(256) nce_prep_kernel(const float* __restrict__ V1, const float* __restrict__ V2,
                                                       const int64_t* __restrict__ idx, const int* __restrict__ d_n, int n_in, float inv_tau, NceWs w) {
    const int n = d_n ? *d_n : n_in;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const size_t r = (size_t)idx[i] * 64;
    const float2 a = *reinterpret_cast<const float2*>(V1 + r + lane * 2);
    const float2 b = *reinterpret_cast<const float2*>(V2 + r + lane * 2);
    float sa = a.x * a.x + a.y * a.y, sb = b.x * b.x + b.y * b.y;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, m); sb += __shfl_xor_sync(0xffffffffu, sb, m); }
    const float na = fmaxf(sqrtf(sa), 1e-12f), nb = fmaxf(sqrtf(sb), 1e-12f);
    const float2 an = make_float2(a.x / na, a.y / na), bn = make_float2(b.x / nb, b.y / nb);
    *reinterpret_cast<float2*>(w.A + (size_t)i * 64 + lane * 2) = an;
    *reinterpret_cast<float2*>(w.Bm + (size_t)i * 64 + lane * 2) = bn;
    float dot = an.x * bn.x + an.y * bn.y;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, m);
    if (lane == 0) { w.na[i] = na; w.nb[i] = nb; w.pos[i] = expf(dot * inv_tau); }
}

// One pass over a (row tile, column split):  E_rc = exp(<x_r,y_c>/tau) * (beta ? beta_c : 1)
//   MODE 0: part_sum[split][r]   = sum_c E_rc
//   MODE 1: part_out[split][r][:] = sum_c E_rc * y_c
// TRANS = 0: rows = X, S = <x_r, y_c>;  (S is symmetric in its arguments so TRANS is only naming)
template <int MODE>
__global__ void __launch_bounds__(256) nce_pass_kernel(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ beta,
                                                       const int* __restrict__ d_n, int n_in, int nstride, float inv_tau,
                                                       float* __restrict__ part_sum, float* __restrict__ part_out) {
    const int n = d_n ? *d_n : n_in;
    if ((int)blockIdx.x * kNT >= n) return;
    extern __shared__ __align__(16) float nce_smem[];
    float* Xt = nce_smem;            // [k][r]
    float* Yt = Xt + 64 * kNT;       // [k][c]
    float* Yr = Yt + 64 * kNT;       // [c][k]
    float* Et = Yr + kNT * 64;       // [c][r]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = blockIdx.x * kNT;
    const int split = blockIdx.y;
    const int ntiles = (n + kNT - 1) / kNT;
    const int t_begin = (int)((long long)ntiles * split / kNceSplits), t_end = (int)((long long)ntiles * (split + 1) / kNceSplits);
    for (int q = warp; q < 32; q += 8) {
        const int rl = (q & 1) * 32 + lane, c4 = q >> 1;
        const int r = r0 + rl;
        const float4 v = (r < n) ? ldg4(X + (size_t)r * 64 + c4 * 4) : f4zero();
        Xt[(c4 * 4 + 0) * kNT + rl] = v.x; Xt[(c4 * 4 + 1) * kNT + rl] = v.y; Xt[(c4 * 4 + 2) * kNT + rl] = v.z; Xt[(c4 * 4 + 3) * kNT + rl] = v.w;
    }
    const int ty = tid >> 4, tx = tid & 15;
    float rs[4] = {0.f, 0.f, 0.f, 0.f};
    float out[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) out[a][b] = 0.f;

    for (int t = t_begin; t < t_end; ++t) {
        const int c0 = t * kNT;
        __syncthreads();
        for (int q = warp; q < 32; q += 8) {
            const int cl = (q & 1) * 32 + lane, c4 = q >> 1;
            const int c = c0 + cl;
            const float4 v = (c < n) ? ldg4(Y + (size_t)c * 64 + c4 * 4) : f4zero();
            Yt[(c4 * 4 + 0) * kNT + cl] = v.x; Yt[(c4 * 4 + 1) * kNT + cl] = v.y; Yt[(c4 * 4 + 2) * kNT + cl] = v.z; Yt[(c4 * 4 + 3) * kNT + cl] = v.w;
            if (MODE == 1) *reinterpret_cast<float4*>(Yr + cl * 64 + c4 * 4) = v;
        }
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 8
        for (int k = 0; k < 64; ++k) {
            const float4 xx = *reinterpret_cast<const float4*>(Xt + k * kNT + ty * 4);
            const float4 yy = *reinterpret_cast<const float4*>(Yt + k * kNT + tx * 4);
            const float xa[4] = {xx.x, xx.y, xx.z, xx.w}, yb[4] = {yy.x, yy.y, yy.z, yy.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(xa[a], yb[b], acc[a][b]);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int c = c0 + tx * 4 + b;
            const float bc = (c < n) ? (beta ? __ldg(beta + c) : 1.f) : 0.f;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float e = (c < n) ? expf(acc[a][b] * inv_tau) * bc : 0.f;
                if (MODE == 0) rs[a] += e;
                else Et[(tx * 4 + b) * kNT + ty * 4 + a] = e;
            }
        }
        if (MODE == 1) {
            __syncthreads();
            // out[r][k] += sum_c E[r][c] * Y[c][k];  thread: rows ty*4.., k columns tx*4..
#pragma unroll 8
            for (int c = 0; c < kNT; ++c) {
                const float4 ee = *reinterpret_cast<const float4*>(Et + c * kNT + ty * 4);
                const float4 yy = *reinterpret_cast<const float4*>(Yr + c * 64 + tx * 4);
                const float ea[4] = {ee.x, ee.y, ee.z, ee.w}, yb[4] = {yy.x, yy.y, yy.z, yy.w};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) out[a][b] = fmaf(ea[a], yb[b], out[a][b]);
            }
        }
    }
    if (MODE == 0) {
        // reduce the 16 tx partial row sums in a fixed order through shared memory
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 4; ++a) Et[(ty * 4 + a) * 16 + tx] = rs[a];
        __syncthreads();
        if (tid < kNT && r0 + tid < n) {
            float s = 0.f;
            for (int q = 0; q < 16; ++q) s += Et[tid * 16 + q];
            part_sum[(size_t)split * nstride + r0 + tid] = s;
        }
    } else {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int r = r0 + ty * 4 + a;
            if (r < n) *reinterpret_cast<float4*>(part_out + ((size_t)split * nstride + r) * 64 + tx * 4) = make_float4(out[a][0], out[a][1], out[a][2], out[a][3]);
        }
    }
}

// ttl, loss terms, backward row weights; single CTA, fixed-order loss reduction
__global__ void __launch_bounds__(1024) nce_rows_kernel(NceWs w, const int* __restrict__ d_n, int n_in, int nstride, float loss_scale, float* __restrict__ loss) {
    const int n = d_n ? *d_n : n_in;
    __shared__ float sh[1024];
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += 1024) {
        float ttl = 0.f;
        for (int s = 0; s < kNceSplits; ++s) ttl += w.part_sum[(size_t)s * nstride + i];
        const float r = w.pos[i] / ttl;
        const float li = -logf(r + 1e-5f);
        const float wi = -r / ((float)n * (r + 1e-5f));
        w.inv_ttl[i] = 1.f / ttl; w.wrow[i] = wi; w.beta[i] = wi / ttl; w.loss_i[i] = li;
        a += li;
    }
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int s = 512; s >= 1; s >>= 1) { if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s]; __syncthreads(); }
    // atomic: the user-side and item-side calls of one step run on two streams; two addends onto a zeroed slot commute exactly
    if (threadIdx.x == 0 && n > 0) atomicAdd(loss, loss_scale * (sh[0] / (float)n));
}

// final gradients, through the normalisation, accumulated into rows idx of gV1/gV2; one warp per row
__global__ void __launch_bounds__(256) nce_grad_kernel(NceWs w, const int64_t* __restrict__ idx, const int* __restrict__ d_n, int n_in, int nstride, float inv_tau, float scale,
                                                       float* gV1, float* gV2) {  // may alias (SimGCL accumulates both views into one buffer)
    const int n = d_n ? *d_n : n_in;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const size_t o = (size_t)i * 64 + lane * 2;
    float2 pb = make_float2(0.f, 0.f), qa = make_float2(0.f, 0.f);
    for (int s = 0; s < kNceSplits; ++s) {
        const float2 x = *reinterpret_cast<const float2*>(w.part_pb + (size_t)s * nstride * 64 + o);
        const float2 y = *reinterpret_cast<const float2*>(w.part_qa + (size_t)s * nstride * 64 + o);
        pb.x += x.x; pb.y += x.y; qa.x += y.x; qa.y += y.y;
    }
    const float2 a = *reinterpret_cast<const float2*>(w.A + o), b = *reinterpret_cast<const float2*>(w.Bm + o);
    const float wi = w.wrow[i], it = w.inv_ttl[i];
    float2 ga = make_float2(wi * inv_tau * (b.x - pb.x * it), wi * inv_tau * (b.y - pb.y * it));
    float2 gb = make_float2(inv_tau * (wi * a.x - qa.x), inv_tau * (wi * a.y - qa.y));
    float da = a.x * ga.x + a.y * ga.y, db = b.x * gb.x + b.y * gb.y;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) { da += __shfl_xor_sync(0xffffffffu, da, m); db += __shfl_xor_sync(0xffffffffu, db, m); }
    const float na = w.na[i], nb = w.nb[i];
    const size_t r = (size_t)idx[i] * 64 + lane * 2;
    if (gV1) { float2 g = *reinterpret_cast<float2*>(gV1 + r); g.x += scale * (ga.x - a.x * da) / na; g.y += scale * (ga.y - a.y * da) / na; *reinterpret_cast<float2*>(gV1 + r) = g; }
    if (gV2) { float2 g = *reinterpret_cast<float2*>(gV2 + r); g.x += scale * (gb.x - b.x * db) / nb; g.y += scale * (gb.y - b.y * db) / nb; *reinterpret_cast<float2*>(gV2 + r) = g; }
}

}  // namespace idg

using namespace idg;

static bool nce_use_tc(int n_max) {
    static const bool off = getenv("IDG_NCE_IMPL") && strcmp(getenv("IDG_NCE_IMPL"), "fma") == 0;
    return !off && n_max >= 256;  // below that the tiled CUDA-core passes are launch-bound anyway
}

extern "C" int64_t idg_infonce_workspace_bytes(int32_t n, int32_t d) {
    if (n <= 0 || d <= 0) return 0;
    n = (n + 127) / 128 * 128;
    return (int64_t)nce_tc_extra_bytes(n) + (int64_t)(2 * nce_align(sizeof(float) * (size_t)n * d) + 7 * nce_align(sizeof(float) * (size_t)n) +
                     nce_align(sizeof(float) * (size_t)kNceSplits * n) + 2 * nce_align(sizeof(float) * (size_t)kNceSplits * n * d));
}

static int infonce_impl(const float* d_V1, const float* d_V2, const int64_t* d_idx, const int* d_n, int32_t n, int32_t d, float temperature,
                        float loss_scale, float* d_loss, float* d_gV1, float* d_gV2, void* d_ws, void* stream_) {
    if (!d_V1 || !d_V2 || !d_idx || !d_loss || !d_ws) return fail(-1, "idg_infonce_fwd_bwd: null argument%s");
    if (n <= 0) return fail(-1, "idg_infonce_fwd_bwd: n must be > 0%s");
    if (d != 64) return fail(-1, "idg_infonce_fwd_bwd: d must be 64 (%s%lld)", "", d);
    if (!(temperature > 0.f)) return fail(-1, "idg_infonce_fwd_bwd: temperature must be > 0%s");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int np = (n + 127) / 128 * 128;  // partial buffers use the padded row stride on both paths
    NceWs w = nce_carve(d_ws, np, d);
    const float inv_tau = 1.f / temperature;
    const dim3 grid((n + kNT - 1) / kNT, kNceSplits);
    nce_prep_kernel<<<(n + 7) / 8, 256, 0, stream>>>(d_V1, d_V2, d_idx, d_n, n, inv_tau, w);
    IDG_LAUNCH_CHECK("nce_prep_kernel");
    // IDG_NCE_IMPL=split keeps the round-1 form (E and E^T materialised as (hi, lo) pairs, separate gemm launches) as a cross-check
    static const bool flash = !(getenv("IDG_NCE_IMPL") && strcmp(getenv("IDG_NCE_IMPL"), "split") == 0);
    if (nce_use_tc(n) && flash && (d_gV1 || d_gV2)) {
        // fused tensor-core path (csrc/infonce_flash.cu): E = exp(A B^T / tau) never leaves the SM.  Pass 1: row sums and P B = E B;
        // row kernel: loss, beta; pass 2 (mirrored): Q A = E^T (beta A); gradient kernel.
        if (int rc = nce_flash(w.A, w.Bm, w.Bm, nullptr, d_n, n, inv_tau, w.part_sum, w.part_pb, kNceSplits, stream)) return rc;
        nce_rows_kernel<<<1, 1024, 0, stream>>>(w, d_n, n, np, loss_scale, d_loss);
        IDG_LAUNCH_CHECK("nce_rows_kernel");
        if (int rc = nce_flash(w.Bm, w.A, w.A, w.beta, d_n, n, inv_tau, nullptr, w.part_qa, kNceSplits, stream)) return rc;
        nce_grad_kernel<<<(n + 7) / 8, 256, 0, stream>>>(w, d_idx, d_n, n, np, inv_tau, loss_scale, d_gV1, d_gV2);
        IDG_LAUNCH_CHECK("nce_grad_kernel");
        return 0;
    }
    if (nce_use_tc(n)) {
        // tensor-core path: E = exp(A B^T/tau) and its row sums, then (if gradients are wanted) E^T, PB = E B, QA = E^T (beta A)
        if (int rc = nce_tc_stage(0, w.A, w.Bm, nullptr, d_n, n, inv_tau, w.part_sum, nullptr, nullptr, w.extra, stream, (d_gV1 || d_gV2) ? 1 : 0, kNceSplits)) return rc;
        nce_rows_kernel<<<1, 1024, 0, stream>>>(w, d_n, n, np, loss_scale, d_loss);
        IDG_LAUNCH_CHECK("nce_rows_kernel");
        if (d_gV1 || d_gV2) {
            if (int rc = nce_tc_stage(1, w.A, w.Bm, w.beta, d_n, n, inv_tau, nullptr, w.part_pb, w.part_qa, w.extra, stream, 1, kNceSplits)) return rc;
            nce_grad_kernel<<<(n + 7) / 8, 256, 0, stream>>>(w, d_idx, d_n, n, np, inv_tau, loss_scale, d_gV1, d_gV2);
            IDG_LAUNCH_CHECK("nce_grad_kernel");
        }
        return 0;
    }
    const size_t smem = sizeof(float) * 4 * 64 * kNT;
    IDG_CUDA(cudaFuncSetAttribute(nce_pass_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    IDG_CUDA(cudaFuncSetAttribute(nce_pass_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nce_pass_kernel<0><<<grid, 256, smem, stream>>>(w.A, w.Bm, nullptr, d_n, n, np, inv_tau, w.part_sum, nullptr);
    IDG_LAUNCH_CHECK("nce_pass_kernel<0>");
    nce_rows_kernel<<<1, 1024, 0, stream>>>(w, d_n, n, np, loss_scale, d_loss);
    IDG_LAUNCH_CHECK("nce_rows_kernel");
    if (d_gV1 || d_gV2) {
        nce_pass_kernel<1><<<grid, 256, smem, stream>>>(w.A, w.Bm, nullptr, d_n, n, np, inv_tau, nullptr, w.part_pb);   // sum_j e_ij b_j
        IDG_LAUNCH_CHECK("nce_pass_kernel<1>");
        nce_pass_kernel<1><<<grid, 256, smem, stream>>>(w.Bm, w.A, w.beta, d_n, n, np, inv_tau, nullptr, w.part_qa);    // sum_i beta_i e_ij a_i
        IDG_LAUNCH_CHECK("nce_pass_kernel<1>");
        nce_grad_kernel<<<(n + 7) / 8, 256, 0, stream>>>(w, d_idx, d_n, n, np, inv_tau, loss_scale, d_gV1, d_gV2);
        IDG_LAUNCH_CHECK("nce_grad_kernel");
    }
    return 0;
}

extern "C" int idg_infonce_fwd_bwd(const float* d_V1, const float* d_V2, const int64_t* d_idx, int32_t n, int32_t d, float temperature,
                                   float loss_scale, float* d_loss, float* d_gV1, float* d_gV2, void* d_ws, void* stream) {
    return infonce_impl(d_V1, d_V2, d_idx, nullptr, n, d, temperature, loss_scale, d_loss, d_gV1, d_gV2, d_ws, stream);
}

// Same, with the row count on the device (*d_n <= n_max): no host sync between torch.unique's job and the loss,
// so the whole contrastive step can be captured in a CUDA graph.  Workspace sized for n_max.
extern "C" int idg_infonce_fwd_bwd_dev(const float* d_V1, const float* d_V2, const int64_t* d_idx, const int32_t* d_n, int32_t n_max,
                                       int32_t d, float temperature, float loss_scale, float* d_loss, float* d_gV1, float* d_gV2,
                                       void* d_ws, void* stream) {
    if (!d_n) return fail(-1, "idg_infonce_fwd_bwd_dev: null d_n%s");
    return infonce_impl(d_V1, d_V2, d_idx, d_n, n_max, d, temperature, loss_scale, d_loss, d_gV1, d_gV2, d_ws, stream);
}

// torch.unique(ids) (SimGCL.py:80-81) on the device without a host sync: sorted distinct values (+offset) and
// their count.  One CTA; n <= 4096.  Deterministic: rank = number of distinct smaller values.
namespace idg {
__global__ void __launch_bounds__(1024) unique_rows_kernel(const int64_t* __restrict__ ids, int n, int64_t offset, int64_t* __restrict__ out,
                                                           int* __restrict__ out_count) {
    __shared__ int key[4096];
    __shared__ unsigned char lead[4096];
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    for (int i = threadIdx.x; i < n; i += 1024) key[i] = (int)ids[i];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 1024) {
        const int k = key[i];
        bool first = true;
        for (int j = 0; j < i; ++j) if (key[j] == k) { first = false; break; }
        lead[i] = first;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 1024) {
        if (!lead[i]) continue;
        const int k = key[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += (lead[j] && key[j] < k);
        out[rank] = (int64_t)k + offset;
        atomicAdd(&total, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) *out_count = total;
}
}  // namespace idg

extern "C" int idg_unique_rows(const int64_t* d_ids, int32_t n, int64_t offset, int64_t* d_out, int32_t* d_out_count, void* stream) {
    if (!d_ids || !d_out || !d_out_count || n <= 0 || n > 4096) return fail(-1, "idg_unique_rows: bad argument (n in 1..4096)%s");
    unique_rows_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_ids, n, offset, d_out, d_out_count);
    IDG_LAUNCH_CHECK("unique_rows_kernel");
    return 0;
}
