// Fused BPR + L2-reg loss, forward and backward (sm_100a).
//
// Replaces models/LightGCN.py:57-70 (6 gathers) + utility_function/losses.py:4-21 and the
// autograd index_put scatter (trainer.py:55).  HBM/L2-bound gather work, no tensor cores.
//
//   x_b   = <u,p> - <u,n>;  loss = mean_b -log(sigmoid(x_b) + 1e-7)        (losses.py:6-13, 10e-8 == 1e-7)
//   reg   = lambda * sum_t 0.5*||E0_t||^2 / B                              (losses.py:16-21)
//   c_b   = dloss/dx_b = -s(1-s)/(s+1e-7)/B
//   G[u_b] += c_b (p - n);  G[U+p_b] += c_b u;  G[U+n_b] -= c_b u
//
// Deterministic scatter without float atomics or a sort: the 3B (node) keys are
// scanned by one warp per entry; the first occurrence of a node ("leader") sums the
// contributions of all its occurrences in ascending entry order and writes the row.
#include <math.h>

#include "idg_common.cuh"

namespace idg {

struct BprWs {
    int* keys;       // [3B] global row of entry e = role*B + b
    float* coef;     // [B]  c_b
    float* loss_b;   // [B]
    float* reg_b;    // [B]  0.5*(masked ego norms)
    int* lead_node;  // [3B] row if entry is its node's first occurrence else -1
    int* lead_mult;  // [3B] occurrences of that row among the reg-masked roles
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

__host__ inline BprWs bpr_carve(void* ws, int B) {
    char* p = (char*)ws;
    BprWs w;
    w.keys = (int*)p; p += align256(sizeof(int) * 3 * (size_t)B);
    w.coef = (float*)p; p += align256(sizeof(float) * (size_t)B);
    w.loss_b = (float*)p; p += align256(sizeof(float) * (size_t)B);
    w.reg_b = (float*)p; p += align256(sizeof(float) * (size_t)B);
    w.lead_node = (int*)p; p += align256(sizeof(int) * 3 * (size_t)B);
    w.lead_mult = (int*)p; p += align256(sizeof(int) * 3 * (size_t)B);
    return w;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// one warp per sample; d = 4*LPR*? -> each lane covers d/32 consecutive floats (d in {32,64,128,256})
template <int VPL>  // floats per lane
__global__ void __launch_bounds__(256) bpr_fwd_kernel(const float* __restrict__ F, const float* __restrict__ E0,
                                                      const int64_t* __restrict__ user, const int64_t* __restrict__ pos,
                                                      const int64_t* __restrict__ neg, int B, int U, int reg_mask, BprWs w) {
    constexpr int d = 32 * VPL;
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    const int ru = (int)user[b], rp = U + (int)pos[b], rn = U + (int)neg[b];
    float dp = 0.f, dn = 0.f, rg = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int c = lane * VPL + v;
        const float fu = __ldg(F + (size_t)ru * d + c), fp = __ldg(F + (size_t)rp * d + c), fn = __ldg(F + (size_t)rn * d + c);
        dp = fmaf(fu, fp, dp);
        dn = fmaf(fu, fn, dn);
        if (reg_mask & 1) { const float e = __ldg(E0 + (size_t)ru * d + c); rg = fmaf(e, e, rg); }
        if (reg_mask & 2) { const float e = __ldg(E0 + (size_t)rp * d + c); rg = fmaf(e, e, rg); }
        if (reg_mask & 4) { const float e = __ldg(E0 + (size_t)rn * d + c); rg = fmaf(e, e, rg); }
    }
    dp = warp_sum(dp); dn = warp_sum(dn); rg = warp_sum(rg);
    if (lane == 0) {
        const float x = dp - dn;
        const float s = 1.f / (1.f + expf(-x));
        w.loss_b[b] = -logf(s + 1e-7f);
        w.coef[b] = -(s * (1.f - s)) / (s + 1e-7f) / (float)B;
        w.reg_b[b] = 0.5f * rg;
        w.keys[b] = ru; w.keys[B + b] = rp; w.keys[2 * B + b] = rn;
    }
}

// fixed-order block reduction of the per-sample terms -> loss[0] = bpr, loss[1] = lambda*reg
// ``tail`` (optional): per-step scalar work that would otherwise be two more single-thread launches in the captured
// step -- the epoch loss sums and the bias-corrected Adam scalars of this step (adam_prepare_kernel's job).
__global__ void __launch_bounds__(1024) bpr_reduce_kernel(BprWs w, int B, float reg_lambda, float* __restrict__ loss, idg_step_tail tail) {
    __shared__ float sl[1024], sr[1024];
    float a = 0.f, r = 0.f;
    for (int i = threadIdx.x; i < B; i += 1024) { a += w.loss_b[i]; r += w.reg_b[i]; }
    sl[threadIdx.x] = a; sr[threadIdx.x] = r;
    __syncthreads();
    for (int s = 512; s >= 1; s >>= 1) {
        if (threadIdx.x < s) { sl[threadIdx.x] += sl[threadIdx.x + s]; sr[threadIdx.x] += sr[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float l0 = sl[0] / (float)B, l1 = reg_lambda * (sr[0] / (float)B);
        loss[0] = l0; loss[1] = l1;
        if (tail.d_loss_acc) { tail.d_loss_acc[0] += (double)l0; tail.d_loss_acc[1] += (double)l1; }
        if (tail.d_step) {
            const double t = (double)(*tail.d_step + 1);
            tail.d_scalars[0] = (float)((double)tail.lr / (1.0 - pow((double)tail.beta1, t)));
            tail.d_scalars[1] = (float)sqrt(1.0 - pow((double)tail.beta2, t));
            *tail.d_step += 1;
        }
    }
}

template <int VPL>
__global__ void __launch_bounds__(256) bpr_bwd_kernel(const float* __restrict__ F, int B, int reg_mask, const float* __restrict__ upstream, BprWs w, float* __restrict__ G,
                                                      float* __restrict__ regc, float reg_coef) {
    constexpr int d = 32 * VPL;
    extern __shared__ __align__(16) int skeys[];  // [3B] padded with -1 to a multiple of 128 (4 keys per lane and step)
    const int n = 3 * B;
    const int npad = (n + 127) & ~127;
    for (int i = threadIdx.x; i < npad; i += blockDim.x) skeys[i] = (i < n) ? w.keys[i] : -1;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= n) return;
    const int node = skeys[e];
    const int4* skeys4 = reinterpret_cast<const int4*>(skeys);
    // leader test: any earlier entry with the same row?
    for (int base = 0; base < e; base += 128) {
        const int i = base + lane * 4;
        const int4 k = skeys4[i >> 2];
        const bool hit = ((i < e) && (k.x == node)) || ((i + 1 < e) && (k.y == node)) || ((i + 2 < e) && (k.z == node)) || ((i + 3 < e) && (k.w == node));
        if (__any_sync(0xffffffffu, hit)) {
            if (lane == 0) { w.lead_node[e] = -1; w.lead_mult[e] = 0; }
            return;
        }
    }
    float acc[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v] = 0.f;
    int mult = 0;
    const float up = upstream ? upstream[0] : 1.f;
    for (int base = e & ~127; base < n; base += 128) {
        const int i = base + lane * 4;
        const int4 k = skeys4[i >> 2];
        unsigned m[4];
        m[0] = __ballot_sync(0xffffffffu, (i >= e) && (k.x == node));
        m[1] = __ballot_sync(0xffffffffu, (i + 1 >= e) && (k.y == node));
        m[2] = __ballot_sync(0xffffffffu, (i + 2 >= e) && (k.z == node));
        m[3] = __ballot_sync(0xffffffffu, (i + 3 >= e) && (k.w == node));
        unsigned any = m[0] | m[1] | m[2] | m[3];
        while (any) {                       // ascending entry order: lane-major, then the 4 keys of that lane
            const int L = __ffs(any) - 1;
            any &= any - 1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (!((m[q] >> L) & 1u)) continue;
                const int j = base + L * 4 + q;
                const int role = j / B, b = j - role * B;
                mult += (reg_mask >> role) & 1;
                const float c = w.coef[b] * up;
                if (role == 0) {
                    const float* p = F + (size_t)skeys[B + b] * d;
                    const float* q2 = F + (size_t)skeys[2 * B + b] * d;
#pragma unroll
                    for (int v = 0; v < VPL; ++v) acc[v] += c * (__ldg(p + lane * VPL + v) - __ldg(q2 + lane * VPL + v));
                } else {
                    const float* u = F + (size_t)skeys[b] * d;
                    const float cc = (role == 1) ? c : -c;
#pragma unroll
                    for (int v = 0; v < VPL; ++v) acc[v] += cc * __ldg(u + lane * VPL + v);
                }
            }
        }
    }
#pragma unroll
    for (int v = 0; v < VPL; ++v) G[(size_t)node * d + lane * VPL + v] = acc[v];
    if (lane == 0) {
        w.lead_node[e] = node; w.lead_mult[e] = mult;
        // per-row L2-reg coefficient for the Adam-fused last backward layer: g += regc[row] * E0[row]
        if (regc) regc[node] = reg_coef * (upstream ? upstream[1] : 1.f) * (float)mult;
    }
}

template <int VPL>
__global__ void __launch_bounds__(256) bpr_finish_kernel(const float* __restrict__ E0, float* __restrict__ gE0, float* __restrict__ G,
                                                         int n, float coef, const float* __restrict__ upstream, BprWs w,
                                                         float* __restrict__ regc, unsigned* __restrict__ bitmap) {
    constexpr int d = 32 * VPL;
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= n) return;
    const int node = w.lead_node[e];
    if (node < 0) return;
    const float m = coef * (upstream ? upstream[1] : 1.f) * (float)w.lead_mult[e];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const size_t o = (size_t)node * d + lane * VPL + v;
        if (gE0 && m != 0.f) gE0[o] += m * E0[o];
        if (G) G[o] = 0.f;
    }
    if (regc && lane == 0) regc[node] = 0.f;
    // batch row bitmap of the step (idg_batch_rows): the leaders are exactly its rows; racing leaders of one word all store 0
    if (bitmap && lane == 0) bitmap[node >> 5] = 0u;
}

}  // namespace idg

using namespace idg;

extern "C" int64_t idg_bpr_workspace_bytes(int32_t B) {
    if (B <= 0) return 0;
    return (int64_t)(3 * align256(sizeof(int) * 3 * (size_t)B) + 3 * align256(sizeof(float) * (size_t)B));
}

#define IDG_DISPATCH_D(d, ...)                                                            \
    switch (d) {                                                                            \
        case 32: { constexpr int VPL = 1; __VA_ARGS__; break; }                                    \
        case 64: { constexpr int VPL = 2; __VA_ARGS__; break; }                                    \
        case 128: { constexpr int VPL = 4; __VA_ARGS__; break; }                                   \
        case 192: { constexpr int VPL = 6; __VA_ARGS__; break; } /* NGCF concat, GCN_layer = 2 */   \
        case 256: { constexpr int VPL = 8; __VA_ARGS__; break; }                                   \
        case 320: { constexpr int VPL = 10; __VA_ARGS__; break; } /* NGCF concat, GCN_layer = 4 */  \
        default: return fail(-1, "d must be 32, 64, 128, 192, 256 or 320 (%s%lld)", "", (long long)d); \
    }

static int bpr_forward_impl(const float* d_F, const float* d_E0, const int64_t* d_user, const int64_t* d_pos,
                            const int64_t* d_neg, int32_t B, int32_t U, int32_t N, int32_t d, float reg_lambda, int reg_mask,
                            float* d_loss, const idg_step_tail* tail_in, void* d_ws, void* stream_) {
    if (!d_F || !d_E0 || !d_user || !d_pos || !d_neg || !d_loss || !d_ws) return fail(-1, "idg_bpr_forward: null argument%s");
    if (B <= 0 || U <= 0 || N <= U) return fail(-1, "idg_bpr_forward: bad sizes%s");
    if (3 * (size_t)B * sizeof(int) > 200 * 1024) return fail(-1, "idg_bpr_forward: batch too large for the shared-memory key table (B=%s%lld)", "", B);
    cudaStream_t stream = (cudaStream_t)stream_;
    BprWs w = bpr_carve(d_ws, B);
    IDG_DISPATCH_D(d, (bpr_fwd_kernel<VPL><<<(B + 7) / 8, 256, 0, stream>>>(d_F, d_E0, d_user, d_pos, d_neg, B, U, reg_mask, w)));
    IDG_LAUNCH_CHECK("bpr_fwd_kernel");
    idg_step_tail tail = {nullptr, nullptr, nullptr, 0.f, 0.f, 0.f};
    if (tail_in) tail = *tail_in;
    if (tail.d_step && !tail.d_scalars) return fail(-1, "idg_bpr_forward_tail: d_scalars is required with d_step%s");
    bpr_reduce_kernel<<<1, 1024, 0, stream>>>(w, B, reg_lambda, d_loss, tail);
    IDG_LAUNCH_CHECK("bpr_reduce_kernel");
    return 0;
}

extern "C" int idg_bpr_forward(const float* d_F, const float* d_E0, const int64_t* d_user, const int64_t* d_pos,
                               const int64_t* d_neg, int32_t B, int32_t U, int32_t N, int32_t d, float reg_lambda, int reg_mask,
                               float* d_loss, void* d_ws, void* stream) {
    return bpr_forward_impl(d_F, d_E0, d_user, d_pos, d_neg, B, U, N, d, reg_lambda, reg_mask, d_loss, nullptr, d_ws, stream);
}

extern "C" int idg_bpr_forward_tail(const float* d_F, const float* d_E0, const int64_t* d_user, const int64_t* d_pos,
                                    const int64_t* d_neg, int32_t B, int32_t U, int32_t N, int32_t d, float reg_lambda, int reg_mask,
                                    float* d_loss, const idg_step_tail* tail, void* d_ws, void* stream) {
    return bpr_forward_impl(d_F, d_E0, d_user, d_pos, d_neg, B, U, N, d, reg_lambda, reg_mask, d_loss, tail, d_ws, stream);
}

extern "C" int idg_bpr_backward(const float* d_F, int32_t B, int32_t d, int reg_mask, const float* d_upstream, float* d_G,
                                float reg_lambda, float* d_regc, void* d_ws, void* stream_) {
    if (!d_F || !d_G || !d_ws || B <= 0) return fail(-1, "idg_bpr_backward: bad argument%s");
    cudaStream_t stream = (cudaStream_t)stream_;
    BprWs w = bpr_carve(d_ws, B);
    const size_t smem = sizeof(int) * (((3 * (size_t)B) + 127) & ~(size_t)127);
    IDG_DISPATCH_D(d, {
        if (smem > 48 * 1024) IDG_CUDA(cudaFuncSetAttribute(bpr_bwd_kernel<VPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        bpr_bwd_kernel<VPL><<<(3 * B + 7) / 8, 256, smem, stream>>>(d_F, B, reg_mask, d_upstream, w, d_G, d_regc, reg_lambda / (float)B);
    });
    IDG_LAUNCH_CHECK("bpr_bwd_kernel");
    return 0;
}

static int bpr_finish_impl(const float* d_E0, float* d_gE0, float* d_G, int32_t B, int32_t d, float reg_lambda,
                           const float* d_upstream, float* d_regc, uint32_t* d_bitmap, void* d_ws, void* stream_) {
    if (!d_ws || B <= 0 || (d_gE0 && !d_E0)) return fail(-1, "idg_bpr_finish: bad argument%s");
    if (!d_gE0 && !d_G && !d_regc && !d_bitmap) return 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    BprWs w = bpr_carve(d_ws, B);
    const float coef = reg_lambda / (float)B;
    IDG_DISPATCH_D(d, (bpr_finish_kernel<VPL><<<(3 * B + 7) / 8, 256, 0, stream>>>(d_E0, d_gE0, d_G, 3 * B, coef, d_upstream, w, d_regc, d_bitmap)));
    IDG_LAUNCH_CHECK("bpr_finish_kernel");
    return 0;
}

extern "C" int idg_bpr_finish(const float* d_E0, float* d_gE0, float* d_G, int32_t B, int32_t d, float reg_lambda,
                              const float* d_upstream, float* d_regc, void* d_ws, void* stream) {
    return bpr_finish_impl(d_E0, d_gE0, d_G, B, d, reg_lambda, d_upstream, d_regc, nullptr, d_ws, stream);
}

extern "C" int idg_bpr_finish_clear(const float* d_E0, float* d_gE0, float* d_G, int32_t B, int32_t d, float reg_lambda,
                                    const float* d_upstream, float* d_regc, uint32_t* d_bitmap, void* d_ws, void* stream) {
    return bpr_finish_impl(d_E0, d_gE0, d_G, B, d, reg_lambda, d_upstream, d_regc, d_bitmap, d_ws, stream);
}
