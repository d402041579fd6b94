// Library-wide state, element-wise utilities, fused Adam, exact negative-sampler replay.
#include <math.h>

#include "idg_common.cuh"

namespace idg {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

__global__ void axpby_kernel(float* __restrict__ out, float a, const float* __restrict__ x, float b,
                             const float* __restrict__ y, int64_t n4, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        float4 xv = reinterpret_cast<const float4*>(x)[i], yv = reinterpret_cast<const float4*>(y)[i];
        reinterpret_cast<float4*>(out)[i] = make_float4(a * xv.x + b * yv.x, a * xv.y + b * yv.y, a * xv.z + b * yv.z, a * xv.w + b * yv.w);
    }
    if (i == 0) for (int64_t k = n4 * 4; k < n; ++k) out[k] = a * x[k] + b * y[k];
}

__global__ void zero_rows_kernel(float* __restrict__ buf, const int64_t* __restrict__ idx, int n, int d4) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int r = (int)(t / d4), c = (int)(t % d4);
    if (r < n) reinterpret_cast<float4*>(buf)[(size_t)idx[r] * d4 + c] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// torch.optim.Adam single-tensor formula (torch/optim/adam.py _single_tensor_adam, default flags):
//   m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, 1-b2);
//   denom = sqrt(v)/sqrt(bc2) + eps;  p.addcdiv_(m, denom, value=-(lr/bc1))
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int64_t n4, int64_t n, float b1, float b2, float eps, float step_size, float bc2_sqrt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        mm = mm + (gg - mm) * (1.f - b1);
        vv = vv * b2 + (1.f - b2) * gg * gg;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        pp = pp - step_size * (mm / denom);
    };
    if (i < n4) {
        float4 pv = reinterpret_cast<float4*>(p)[i], gv = __ldcs(reinterpret_cast<const float4*>(g) + i);
        float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        upd(pv.x, gv.x, mv.x, vv.x); upd(pv.y, gv.y, mv.y, vv.y); upd(pv.z, gv.z, mv.z, vv.z); upd(pv.w, gv.w, mv.w, vv.w);
        reinterpret_cast<float4*>(p)[i] = pv; reinterpret_cast<float4*>(m)[i] = mv; reinterpret_cast<float4*>(v)[i] = vv;
    }
    if (i == 0) for (int64_t k = n4 * 4; k < n; ++k) upd(p[k], g[k], m[k], v[k]);
}

// CUDA-graph friendly variant: the step number lives on the device (d_step holds the number of
// steps already taken) so a captured launch keeps advancing the bias corrections.
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                int64_t n4, int64_t n, float lr, float b1, float b2, float eps, const int* __restrict__ d_step) {
    __shared__ float s_step_size, s_bc2_sqrt;
    if (threadIdx.x == 0) {
        const double t = (double)(*d_step + 1);
        s_step_size = (float)((double)lr / (1.0 - pow((double)b1, t)));
        s_bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, t));
    }
    __syncthreads();
    const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        mm = mm + (gg - mm) * (1.f - b1);
        vv = vv * b2 + (1.f - b2) * gg * gg;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        pp = pp - step_size * (mm / denom);
    };
    if (i < n4) {
        float4 pv = reinterpret_cast<float4*>(p)[i], gv = __ldcs(reinterpret_cast<const float4*>(g) + i);
        float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        upd(pv.x, gv.x, mv.x, vv.x); upd(pv.y, gv.y, mv.y, vv.y); upd(pv.z, gv.z, mv.z, vv.z); upd(pv.w, gv.w, mv.w, vv.w);
        reinterpret_cast<float4*>(p)[i] = pv; reinterpret_cast<float4*>(m)[i] = mv; reinterpret_cast<float4*>(v)[i] = vv;
    }
    if (i == 0) for (int64_t k = n4 * 4; k < n; ++k) upd(p[k], g[k], m[k], v[k]);
}
__global__ void incr_kernel(int* c) { *c += 1; }
// bias-corrected step scalars for the Adam-fused SpMM epilogue: scalars[0] = lr/(1-b1^t), scalars[1] = sqrt(1-b2^t)
__global__ void adam_prepare_kernel(int* d_step, float* scalars, float lr, float b1, float b2) {
    const double t = (double)(*d_step + 1);
    scalars[0] = (float)((double)lr / (1.0 - pow((double)b1, t)));
    scalars[1] = (float)sqrt(1.0 - pow((double)b2, t));
    *d_step += 1;
}
}  // namespace idg

using namespace idg;

extern "C" int idg_version(void) { return 100; }
extern "C" const char* idg_last_error(void) { return g_err; }
extern "C" int64_t idg_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int idg_axpby(float* d_out, float a, const float* d_x, float b, const float* d_y, int64_t n, void* stream) {
    if (!d_out || !d_x || !d_y || n < 0) return fail(-1, "idg_axpby: bad argument%s");
    if (n == 0) return 0;
    if (((uintptr_t)d_out | (uintptr_t)d_x | (uintptr_t)d_y) & 15) return fail(-1, "idg_axpby: pointers must be 16-byte aligned%s");
    const int64_t n4 = n / 4, th = (n4 > 0 ? n4 : 1);
    axpby_kernel<<<(unsigned)((th + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_out, a, d_x, b, d_y, n4, n);
    IDG_LAUNCH_CHECK("axpby_kernel");
    return 0;
}

__global__ void accumulate_f64_kernel(double* __restrict__ acc, const float* __restrict__ x, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acc[i] += (double)x[i];
}
extern "C" int idg_accumulate_f64(double* d_acc, const float* d_x, int32_t n, void* stream) {
    if (!d_acc || !d_x || n < 0) return fail(-1, "idg_accumulate_f64: bad argument%s");
    if (n == 0) return 0;
    accumulate_f64_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_acc, d_x, n);
    IDG_LAUNCH_CHECK("accumulate_f64_kernel");
    return 0;
}

// rows whose bit is set in the bitmap are zeroed: one warp per bitmap word (32 rows), d/4 lanes store one row per set bit,
// so the cost follows the number of flagged rows, not N
__global__ void __launch_bounds__(256) zero_rows_bitmap_kernel(float* __restrict__ buf, const unsigned* __restrict__ bitmap, int n_rows, int d4) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w * 32 >= n_rows) return;
    unsigned m = __ldg(bitmap + w);
    while (m) {
        const int r = w * 32 + __ffs(m) - 1;
        m &= m - 1;
        if (r < n_rows) for (int c = lane; c < d4; c += 32) reinterpret_cast<float4*>(buf)[(size_t)r * d4 + c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
extern "C" int idg_zero_rows_bitmap(float* d_buf, const uint32_t* d_bitmap, int32_t n_rows, int32_t d, void* stream) {
    if (!d_buf || !d_bitmap || n_rows < 0 || d <= 0 || (d & 3)) return fail(-1, "idg_zero_rows_bitmap: bad argument%s");
    if (n_rows == 0) return 0;
    const int words = (n_rows + 31) / 32;
    zero_rows_bitmap_kernel<<<(unsigned)((words + 7) / 8), 256, 0, (cudaStream_t)stream>>>(d_buf, d_bitmap, n_rows, d / 4);
    IDG_LAUNCH_CHECK("zero_rows_bitmap_kernel");
    return 0;
}

extern "C" int idg_zero_rows(float* d_buf, const int64_t* d_idx, int32_t n, int32_t d, void* stream) {
    if (!d_buf || !d_idx || n < 0 || d <= 0 || (d & 3)) return fail(-1, "idg_zero_rows: bad argument%s");
    if (n == 0) return 0;
    const int64_t th = (int64_t)n * (d / 4);
    zero_rows_kernel<<<(unsigned)((th + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_buf, d_idx, n, d / 4);
    IDG_LAUNCH_CHECK("zero_rows_kernel");
    return 0;
}

extern "C" int idg_adam_step(float* d_p, const float* d_g, float* d_m, float* d_v, int64_t n, float lr, float beta1,
                             float beta2, float eps, int32_t step, void* stream) {
    if (!d_p || !d_g || !d_m || !d_v || n < 0 || step < 1) return fail(-1, "idg_adam_step: bad argument%s");
    if (n == 0) return 0;
    if (((uintptr_t)d_p | (uintptr_t)d_g | (uintptr_t)d_m | (uintptr_t)d_v) & 15) return fail(-1, "idg_adam_step: pointers must be 16-byte aligned%s");
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1), bc2_sqrt = (float)sqrt(bc2);
    const int64_t n4 = n / 4, th = (n4 > 0 ? n4 : 1);
    adam_kernel<<<(unsigned)((th + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_p, d_g, d_m, d_v, n4, n, beta1, beta2, eps, step_size, bc2_sqrt);
    IDG_LAUNCH_CHECK("adam_kernel");
    return 0;
}

extern "C" int idg_adam_step_dev(float* d_p, const float* d_g, float* d_m, float* d_v, int64_t n, float lr, float beta1,
                                 float beta2, float eps, int32_t* d_step, void* stream) {
    if (!d_p || !d_g || !d_m || !d_v || !d_step || n < 0) return fail(-1, "idg_adam_step_dev: bad argument%s");
    if (n == 0) return 0;
    if (((uintptr_t)d_p | (uintptr_t)d_g | (uintptr_t)d_m | (uintptr_t)d_v) & 15) return fail(-1, "idg_adam_step_dev: pointers must be 16-byte aligned%s");
    const int64_t n4 = n / 4, th = (n4 > 0 ? n4 : 1);
    adam_dev_kernel<<<(unsigned)((th + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_p, d_g, d_m, d_v, n4, n, lr, beta1, beta2, eps, d_step);
    IDG_LAUNCH_CHECK("adam_dev_kernel");
    incr_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_step);
    IDG_LAUNCH_CHECK("incr_kernel");
    return 0;
}

extern "C" int idg_adam_prepare(int32_t* d_step, float* d_scalars, float lr, float beta1, float beta2, void* stream) {
    if (!d_step || !d_scalars) return fail(-1, "idg_adam_prepare: null argument%s");
    adam_prepare_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_step, d_scalars, lr, beta1, beta2);
    IDG_LAUNCH_CHECK("adam_prepare_kernel");
    return 0;
}

// data_loader.py:108-127 replayed against a bulk candidate stream.  Host code.
extern "C" int idg_neg_sample_replay(const int64_t* h_train_user, int64_t E, const int32_t* h_pos_indptr,
                                     const int32_t* h_pos_indices, const int64_t* h_cand, int64_t n_cand, int64_t* h_neg,
                                     int64_t* h_consumed) {
    if (!h_train_user || !h_pos_indptr || !h_pos_indices || !h_cand || !h_neg || !h_consumed || E < 0)
        return fail(-1, "idg_neg_sample_replay: bad argument%s");
    int64_t j = 0;
    for (int64_t e = 0; e < E; ++e) {
        const int64_t u = h_train_user[e];
        const int32_t* lo = h_pos_indices + h_pos_indptr[u];
        const int32_t* hi = h_pos_indices + h_pos_indptr[u + 1];
        for (;;) {
            if (j >= n_cand) { *h_consumed = j; return fail(-2, "idg_neg_sample_replay: candidate stream exhausted%s"); }
            const int64_t c = h_cand[j++];
            // `neg in all_positive[user]` (data_loader.py:121): binary search in the sorted positives
            const int32_t* a = lo; const int32_t* b = hi;
            while (a < b) { const int32_t* m = a + (b - a) / 2; if (*m < c) a = m + 1; else b = m; }
            if (a < hi && *a == c) continue;
            h_neg[e] = c;
            break;
        }
    }
    *h_consumed = j;
    return 0;
}

// Resumable form of the same walk: edges [e_begin, E) against one chunk of the candidate stream.  Running out of
// candidates is a normal outcome (the caller draws the next chunk from the numpy generator and resumes at
// *h_edges_done), so a first chunk of exactly E candidates plus small follow-up chunks replaces the over-provisioned
// bulk draw and the second full-length draw that re-positions the generator.
extern "C" int idg_neg_sample_walk(const int64_t* h_train_user, int64_t e_begin, int64_t E, const int32_t* h_pos_indptr,
                                   const int32_t* h_pos_indices, const int64_t* h_cand, int64_t n_cand, int64_t* h_neg,
                                   int64_t* h_edges_done, int64_t* h_consumed) {
    if (!h_train_user || !h_pos_indptr || !h_pos_indices || !h_cand || !h_neg || !h_edges_done || !h_consumed || e_begin < 0 || E < e_begin)
        return fail(-1, "idg_neg_sample_walk: bad argument%s");
    int64_t j = 0, e = e_begin;
    for (; e < E; ++e) {
        const int64_t u = h_train_user[e];
        const int32_t* lo = h_pos_indices + h_pos_indptr[u];
        const int32_t* hi = h_pos_indices + h_pos_indptr[u + 1];
        bool placed = false;
        while (j < n_cand) {
            const int64_t c = h_cand[j++];
            const int32_t* a = lo; const int32_t* b = hi;
            while (a < b) { const int32_t* m = a + (b - a) / 2; if (*m < c) a = m + 1; else b = m; }
            if (a < hi && *a == c) continue;      // `neg in all_positive[user]` (data_loader.py:121): draw again
            h_neg[e] = c;
            placed = true;
            break;
        }
        if (!placed) break;                        // chunk exhausted while edge e was still rejecting
    }
    *h_edges_done = e; *h_consumed = j;
    return 0;
}

// tools.shuffle applied on the device (tools.py:35-52): out[0..3)[k] = (user, pos, neg)[perm[k]]
__global__ void permute3_kernel(const int64_t* __restrict__ a, const int64_t* __restrict__ b, const int64_t* __restrict__ c,
                                const int64_t* __restrict__ perm, int64_t n, int64_t* __restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t p = perm[k];
    out[k] = a[p]; out[n + k] = b[p]; out[2 * n + k] = c[p];
}

extern "C" int idg_permute3(const int64_t* d_a, const int64_t* d_b, const int64_t* d_c, const int64_t* d_perm, int64_t n, int64_t* d_out,
                            void* stream) {
    if (!d_a || !d_b || !d_c || !d_perm || !d_out || n < 0) return fail(-1, "idg_permute3: bad argument%s");
    if (n == 0) return 0;
    permute3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_a, d_b, d_c, d_perm, n, d_out);
    IDG_LAUNCH_CHECK("permute3_kernel");
    return 0;
}

// nn.Tanh of EGCF's propagation (models/EGCF.py:42,52-53,71), forward and backward (gx = gy * (1 - y^2))
__global__ void tanh_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = tanhf(x[i]);
}
__global__ void tanh_bwd_kernel(const float* __restrict__ y, const float* __restrict__ gy, float* __restrict__ gx, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float t = y[i]; gx[i] = gy[i] * (1.f - t * t); }
}

// The step's mini-batch out of the epoch's sample arrays, as the FIRST node of the captured step graph: the host then only replays
// (three small device-to-device copies per step were 8 us of the 438 us amazon-book step).  d_ptrs (device, 4 x int64): the three
// array base addresses and the value of the step counter when the epoch's first batch runs; the batch of this replay starts at
// (*d_step - start) * stride.
__global__ void __launch_bounds__(256) batch_fetch_kernel(const long long* __restrict__ ptrs, const int* __restrict__ d_step, int stride, int B,
                                                          int64_t* __restrict__ slab, int slab_stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const int k = blockIdx.y;
    const long long cursor = ((long long)*d_step - ptrs[3]) * stride;
    slab[(size_t)k * slab_stride + i] = reinterpret_cast<const int64_t*>(ptrs[k])[cursor + i];
}

extern "C" int idg_batch_fetch(const void* d_ptrs, const int32_t* d_step, int32_t stride, int32_t B, int64_t* d_slab, int32_t slab_stride, void* stream) {
    if (!d_ptrs || !d_step || !d_slab || stride <= 0 || B <= 0 || B > slab_stride) return fail(-1, "idg_batch_fetch: bad argument%s");
    batch_fetch_kernel<<<dim3((B + 255) / 256, 3), 256, 0, (cudaStream_t)stream>>>((const long long*)d_ptrs, d_step, stride, B, d_slab, slab_stride);
    IDG_LAUNCH_CHECK("batch_fetch_kernel");
    return 0;
}

extern "C" int idg_tanh_fwd(const float* d_x, float* d_y, int64_t n, void* stream) {
    if (!d_x || !d_y || n < 0) return fail(-1, "idg_tanh_fwd: bad argument%s");
    if (n == 0) return 0;
    tanh_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_x, d_y, n);
    IDG_LAUNCH_CHECK("tanh_fwd_kernel");
    return 0;
}

extern "C" int idg_tanh_bwd(const float* d_y, const float* d_gy, float* d_gx, int64_t n, void* stream) {
    if (!d_y || !d_gy || !d_gx || n < 0) return fail(-1, "idg_tanh_bwd: bad argument%s");
    if (n == 0) return 0;
    tanh_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_y, d_gy, d_gx, n);
    IDG_LAUNCH_CHECK("tanh_bwd_kernel");
    return 0;
}
