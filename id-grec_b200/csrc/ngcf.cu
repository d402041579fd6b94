// NGCF dense layer epilogue, forward and backward (sm_100a).
//
// Replaces the per-layer element-wise / small-GEMM chain of models/NGCF.py:87-106 (two [N,64]x[64,64] matmuls,
// two bias adds, E*side, LeakyReLU(0.2), Dropout, F.normalize, torch.cat) and its autograd backward:
//   Z    = [side | E (*) side]                       [N,128]          side = A_hat . E from the SpMM kernel
//   S    = Z . [W_gcn ; W_bi] + b_gcn + b_bi         [N,64]
//   D    = LeakyReLU_0.2(S) (*) keep / (1-p)          next layer's E
//   O    = D / max(|D|_2, 1e-12)                      written into the layer's 64-column block of the [N,256] concat
// backward (given dO and the gradient dD_ext arriving from the next layer):
//   dD   = dD_ext + (dO - O <O,dO>)/|D| ;  dS = dD (*) keep/(1-p) (*) (S>0 ? 1 : 0.2)
//   dZ   = dS . [W_gcn ; W_bi]^T ;  dside = dZ1 + dZ2 (*) E ;  dE_direct = dZ2 (*) side
//   dW   = Z^T . dS ,  db = colsum(dS)                per-CTA partials, summed in CTA order (deterministic)
// The forward product runs on tcgen05 (csrc/ngcf_tc.cu, 3xTF32 split, TMEM accumulator); this file keeps the backward on fp32
// CUDA-core tiles (16-18 TFLOP/s) and the entry points.  Measured per layer at the amazon-book shape: fp32 tiles 130 us, a
// 3xTF32 mma.sync.m16n8k8 forward 149 us (the warp-level tf32 MMA of sm_100a issues at ~2x the fp32 FMA rate, so three split
// passes lose), tcgen05 with 2 loader warps 220 us, with 7 operand-builder warps and batched loads ~122 us (now bound by the
// one-tile-per-CTA pipeline and the row-per-thread epilogue stores, not by the MMA).
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <curand_kernel.h>

#include "idg_common.cuh"

namespace idg {

// csrc/ngcf_tc.cu: the forward product on tcgen05 (3xTF32 split, TMEM accumulators)
int ngcf_dense_fwd_tc(const float* E, const float* side, const float* Wg, const float* bg, const float* Wb, const float* bb, const float* keep,
                      const uint32_t* keep_bits, float inv_keep, int N, float* S_pre, float* D, float* out, int out_stride, cudaStream_t stream);

// csrc/ngcf_bwd_tc.cu: the backward products on tcgen05 (both contractions from one shared-memory image of dS)
int ngcf_dense_bwd_tc(const float* E, const float* side, const float* Wg, const float* Wb, const float* keep, const uint32_t* keep_bits, float inv_keep,
                      const float* S_pre,
                      const float* D, const float* dO, int dO_stride, const float* dD_ext, int N, float* dside, float* dE_direct, float* dW_part,
                      float* db_part, int max_parts, int* n_parts, cudaStream_t stream);

constexpr int kNgTile = 64;     // rows per tile
constexpr int kNgCtas = 296;    // persistent grid of the backward kernel (2 per SM)

// ---- backward: persistent CTAs loop over 64-row tiles, accumulating dW/db partials in registers
__global__ void __launch_bounds__(256) ngcf_dense_bwd_kernel(const float* __restrict__ E, const float* __restrict__ side,
                                                             const float* __restrict__ Wg, const float* __restrict__ Wb,
                                                             const float* __restrict__ keep, float inv_keep, const float* __restrict__ S_pre,
                                                             const float* __restrict__ D, const float* __restrict__ dO, int dO_stride,
                                                             const float* __restrict__ dD_ext, int N, float* __restrict__ dside,
                                                             float* __restrict__ dE_direct, float* __restrict__ dW_part,
                                                             float* __restrict__ db_part) {
    extern __shared__ __align__(16) float sm[];
    float* Zt = sm;                         // [64 rows][128 k]  (row-major here: used as Z^T . dS with rows as the contraction)
    float* dSs = Zt + kNgTile * 128;        // [64 rows][64 cols]
    float* Wt = dSs + kNgTile * 64;         // [64 cols][128 k]   Wcat^T: Wt[c][k] = Wcat[k][c]
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    for (int q = tid; q < 64 * 64; q += 256) {
        const int k = q >> 6, c = q & 63;
        Wt[c * 128 + k] = Wg[q];
        Wt[c * 128 + 64 + k] = Wb[q];
    }
    // dW partial owned by this thread: Wcat rows k = ty*8 .. ty*8+7, columns tx*4 .. tx*4+3  (128 x 64 = 256 threads x 32)
    float dw[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dw[a][b] = 0.f;
    float dbv = 0.f;  // threads 0..63: column tid
    const int ntiles = (N + kNgTile - 1) / kNgTile;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int r0 = t * kNgTile;
        __syncthreads();
        // phase 1: per row (16 threads per row, 4 rows per warp-pass): dD, dS, Z
        for (int rr = ty; rr < kNgTile; rr += 16) {
            const int r = r0 + rr;
            float4 d4 = f4zero(), o4 = f4zero(), x4 = f4zero(), s4 = f4zero(), k4 = make_float4(1.f, 1.f, 1.f, 1.f), e4 = f4zero(), sd4 = f4zero();
            if (r < N) {
                d4 = ldg4(D + (size_t)r * 64 + tx * 4);
                o4 = ldg4(dO + (size_t)r * dO_stride + tx * 4);
                if (dD_ext) x4 = ldg4(dD_ext + (size_t)r * 64 + tx * 4);
                s4 = ldg4(S_pre + (size_t)r * 64 + tx * 4);
                if (keep) k4 = ldg4(keep + (size_t)r * 64 + tx * 4);
                e4 = ldg4(E + (size_t)r * 64 + tx * 4);
                sd4 = ldg4(side + (size_t)r * 64 + tx * 4);
            }
            float ss = d4.x * d4.x + d4.y * d4.y + d4.z * d4.z + d4.w * d4.w;
            float dot = d4.x * o4.x + d4.y * o4.y + d4.z * o4.z + d4.w * o4.w;  // <D, dO>
#pragma unroll
            for (int m = 8; m >= 1; m >>= 1) { ss += __shfl_xor_sync(0xffffffffu, ss, m); dot += __shfl_xor_sync(0xffffffffu, dot, m); }
            const float nrm = fmaxf(sqrtf(ss), 1e-12f);
            const float proj = dot / (nrm * nrm);  // <O,dO>/|D| with O = D/|D|
            const float dv[4] = {d4.x, d4.y, d4.z, d4.w}, ov[4] = {o4.x, o4.y, o4.z, o4.w}, xv[4] = {x4.x, x4.y, x4.z, x4.w};
            const float sv[4] = {s4.x, s4.y, s4.z, s4.w}, kv[4] = {k4.x, k4.y, k4.z, k4.w};
            const float ev[4] = {e4.x, e4.y, e4.z, e4.w}, sdv[4] = {sd4.x, sd4.y, sd4.z, sd4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float dD = xv[j] + (ov[j] - dv[j] * proj) / nrm;
                float dS = keep ? dD * kv[j] * inv_keep : dD;
                dS *= (sv[j] > 0.f) ? 1.f : 0.2f;
                if (r >= N) dS = 0.f;
                dSs[rr * 64 + tx * 4 + j] = dS;
                Zt[rr * 128 + tx * 4 + j] = sdv[j];
                Zt[rr * 128 + 64 + tx * 4 + j] = ev[j] * sdv[j];
            }
        }
        __syncthreads();
        // phase 2: dZ = dS . Wcat^T  (64 x 128); thread -> rows ty*4.., k columns tx*8..  then dside / dE_direct
        {
            float dz[4][8];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) dz[a][b] = 0.f;
            for (int c = 0; c < 64; ++c) {
                float sa[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) sa[a] = dSs[(ty * 4 + a) * 64 + c];
                const float4 w0 = *reinterpret_cast<const float4*>(Wt + c * 128 + tx * 8), w1 = *reinterpret_cast<const float4*>(Wt + c * 128 + tx * 8 + 4);
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 8; ++b) dz[a][b] = fmaf(sa[a], wv[b], dz[a][b]);
            }
            // thread tx < 8 holds dZ1[:, tx*8..], its partner lane tx+8 holds dZ2[:, tx*8..]: exchange by shuffle
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int rr = ty * 4 + a, r = r0 + rr;
                float o1[8], o2[8];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const float other = __shfl_xor_sync(0xffffffffu, dz[a][b], 8);
                    const int k = (tx & 7) * 8 + b;
                    const float sdv = Zt[rr * 128 + k];
                    const float ev = (r < N && tx < 8) ? __ldg(E + (size_t)r * 64 + k) : 0.f;
                    o1[b] = dz[a][b] + other * ev;   // dside = dZ1 + dZ2 (*) E        (valid for tx < 8)
                    o2[b] = other * sdv;             // dE_direct = dZ2 (*) side
                }
                if (tx < 8 && r < N) {
                    float* ps = dside + (size_t)r * 64 + tx * 8;
                    float* pe = dE_direct + (size_t)r * 64 + tx * 8;
                    st4(ps, make_float4(o1[0], o1[1], o1[2], o1[3])); st4(ps + 4, make_float4(o1[4], o1[5], o1[6], o1[7]));
                    st4(pe, make_float4(o2[0], o2[1], o2[2], o2[3])); st4(pe + 4, make_float4(o2[4], o2[5], o2[6], o2[7]));
                }
            }
        }
        // phase 3: dWcat[k][c] += sum_rows Z[row][k] * dS[row][c]; thread -> k = ty*8.., c = tx*4..
        for (int rr = 0; rr < kNgTile; ++rr) {
            const float4 s = *reinterpret_cast<const float4*>(dSs + rr * 64 + tx * 4);
            const float sc[4] = {s.x, s.y, s.z, s.w};
            const float4 z0 = *reinterpret_cast<const float4*>(Zt + rr * 128 + ty * 8), z1 = *reinterpret_cast<const float4*>(Zt + rr * 128 + ty * 8 + 4);
            const float zv[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dw[a][b] = fmaf(zv[a], sc[b], dw[a][b]);
        }
        if (tid < 64) for (int rr = 0; rr < kNgTile; ++rr) dbv += dSs[rr * 64 + tid];
    }
    float* wp = dW_part + (size_t)blockIdx.x * 128 * 64;
#pragma unroll
    for (int a = 0; a < 8; ++a) st4(wp + (size_t)(ty * 8 + a) * 64 + tx * 4, make_float4(dw[a][0], dw[a][1], dw[a][2], dw[a][3]));
    if (tid < 64) db_part[(size_t)blockIdx.x * 64 + tid] = dbv;
}

// ordered sum of the per-CTA partials: dWg, dWb [64,64], db [64] (shared by b_gcn and b_bi).  64 consecutive outputs per CTA
// (coalesced rows of the partial tables), 4 interleaved slices of the partials per output, folded in slice order.
__global__ void __launch_bounds__(256) ngcf_reduce_kernel(const float* __restrict__ dW_part, const float* __restrict__ db_part, int n_parts,
                                                         float* __restrict__ dWg, float* __restrict__ dWb, float* __restrict__ db) {
    __shared__ float sm[4][64];
    const int ix = threadIdx.x & 63, sl = threadIdx.x >> 6;
    const int i = blockIdx.x * 64 + ix;
    float a = 0.f;
    if (i < 128 * 64) { for (int p = sl; p < n_parts; p += 4) a += dW_part[(size_t)p * 128 * 64 + i]; }
    else if (i < 128 * 64 + 64) { const int c = i - 128 * 64; for (int p = sl; p < n_parts; p += 4) a += db_part[(size_t)p * 64 + c]; }
    sm[sl][ix] = a;
    __syncthreads();
    if (sl == 0 && i < 128 * 64 + 64) {
        a = (sm[0][ix] + sm[1][ix]) + (sm[2][ix] + sm[3][ix]);
        if (i < 64 * 64) dWg[i] = a;
        else if (i < 128 * 64) dWb[i - 64 * 64] = a;
        else db[i - 128 * 64] = a;
    }
}

// nn.Dropout's Bernoulli(1 - p) draws for all layers of one step in ONE launch (NGCF.py:99-100 constructs nn.Dropout inline, so
// it is always active): keep[l][i] = 1 with probability 1 - p_l.  Counter-based Philox4x32-10 evaluated directly: element quad q of
// step t is the block with counter (t, q) under the key `seed` -- every step reads a fresh block of every quad (t comes from the
// device step counter, so a captured step replays with fresh draws) and the result does not depend on the launch geometry.
// (An earlier version went through curand_init(seed, q, t) + curand_uniform4: offset t in VALUES, i.e. the quad of step t + 1 was
// the quad of step t shifted by one element, and sixteen state initialisations per row made the bit-packed kernel 44 us.)
// torch's own bernoulli_ costs three launches per mask (uniform, compare, cast): 150 us of the 1.76 ms round-1 step.
__device__ __forceinline__ uint32_t keep_quad(unsigned long long seed, unsigned long long step, unsigned long long q, float pk) {
    const uint4 r = curand_Philox4x32_10(make_uint4((unsigned)step, (unsigned)(step >> 32), (unsigned)q, (unsigned)(q >> 32)),
                                         make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    const float4 u = _curand_uniform4(r);   // (0, 1]
    return (u.x <= pk ? 1u : 0u) | (u.y <= pk ? 2u : 0u) | (u.z <= pk ? 4u : 0u) | (u.w <= pk ? 8u : 0u);
}

__global__ void __launch_bounds__(256) ngcf_keep_masks_kernel(float* __restrict__ keep, int64_t per_layer4, int n_layers, float k0, float k1,
                                                              float k2, float k3, unsigned long long seed, const int* __restrict__ d_step) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= per_layer4 * n_layers) return;
    const int layer = (int)(q / per_layer4);
    const float pk = layer == 0 ? k0 : (layer == 1 ? k1 : (layer == 2 ? k2 : k3));
    const uint32_t b = keep_quad(seed, (unsigned long long)(d_step ? *d_step : 0), (unsigned long long)q, pk);
    reinterpret_cast<float4*>(keep)[q] = make_float4((float)(b & 1u), (float)((b >> 1) & 1u), (float)((b >> 2) & 1u), (float)((b >> 3) & 1u));
}

// The same draws packed 64 bits per row (word w, bit b = column 32 w + b): what the tensor-core dense kernels read -- 8 bytes per
// row instead of a 256-byte float mask row in the forward and again in the backward.  Quad for quad the blocks of the kernel above;
// four lanes per row (four quads each), combined by shuffles.
__global__ void __launch_bounds__(256) ngcf_keep_bits_kernel(uint2* __restrict__ bits, int N, int n_layers, float k0, float k1, float k2, float k3,
                                                             unsigned long long seed, const int* __restrict__ d_step) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t t = g >> 2;                 // (layer, row)
    const int part = (int)(g & 3);            // quads 4 part .. 4 part + 3 of the row = 16 mask bits
    const bool in = t < (int64_t)N * n_layers;
    uint32_t h = 0u;
    if (in) {
        const int layer = (int)(t / N);
        const float pk = layer == 0 ? k0 : (layer == 1 ? k1 : (layer == 2 ? k2 : k3));
        const unsigned long long step = (unsigned long long)(d_step ? *d_step : 0);
#pragma unroll
        for (int c = 0; c < 4; ++c) h |= keep_quad(seed, step, (unsigned long long)(t * 16 + part * 4 + c), pk) << (c * 4);
    }
    // lanes 4k .. 4k+3 hold the four 16-bit pieces of one row
    const uint32_t nb = __shfl_down_sync(0xffffffffu, h, 1);
    const uint32_t w = h | (nb << 16);        // valid on even parts: parts (0,1) -> word 0, parts (2,3) -> word 1
    const uint32_t w1 = __shfl_down_sync(0xffffffffu, w, 2);
    if (in && part == 0) bits[t] = make_uint2(w, w1);
}

// dst[row] = src[row] for the rows idx[i] + row_offset (row strides in floats): the 64-column ego block of the [N,256] concat is
// only read at the batch rows by the BPR kernels, so the step copies those instead of the whole table
__global__ void __launch_bounds__(256) copy_rows_strided_kernel(const float* __restrict__ src, int src_stride, const int64_t* __restrict__ idx, int n,
                                                                int row_offset, int d4, float* __restrict__ dst, int dst_stride) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t / d4, c = t % d4;
    if (i >= n) return;
    const size_t row = (size_t)idx[i] + row_offset;
    reinterpret_cast<float4*>(dst + row * dst_stride)[c] = __ldg(reinterpret_cast<const float4*>(src + row * src_stride) + c);
}

}  // namespace idg

using namespace idg;

extern "C" int idg_ngcf_keep_masks(float* d_keep, int64_t per_layer, int32_t n_layers, const float* h_keep_prob, uint64_t seed,
                                   const int32_t* d_step, void* stream) {
    if (!d_keep || !h_keep_prob || per_layer <= 0 || (per_layer & 3) || n_layers < 1 || n_layers > 4) return fail(-1, "idg_ngcf_keep_masks: bad argument%s");
    float k[4] = {1.f, 1.f, 1.f, 1.f};
    for (int l = 0; l < n_layers; ++l) k[l] = h_keep_prob[l];
    const int64_t quads = per_layer / 4 * n_layers;
    ngcf_keep_masks_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_keep, per_layer / 4, n_layers, k[0], k[1], k[2], k[3],
                                                                                        (unsigned long long)seed, d_step);
    IDG_LAUNCH_CHECK("ngcf_keep_masks_kernel");
    return 0;
}

extern "C" int idg_ngcf_keep_bits(uint32_t* d_bits, int32_t N, int32_t n_layers, const float* h_keep_prob, uint64_t seed, const int32_t* d_step,
                                  void* stream) {
    if (!d_bits || !h_keep_prob || N <= 0 || n_layers < 1 || n_layers > 4) return fail(-1, "idg_ngcf_keep_bits: bad argument%s");
    float k[4] = {1.f, 1.f, 1.f, 1.f};
    for (int l = 0; l < n_layers; ++l) k[l] = h_keep_prob[l];
    const int64_t n = (int64_t)N * n_layers * 4;      // four lanes per row
    ngcf_keep_bits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint2*>(d_bits), N, n_layers, k[0], k[1], k[2], k[3],
                                                                                      (unsigned long long)seed, d_step);
    IDG_LAUNCH_CHECK("ngcf_keep_bits_kernel");
    return 0;
}

extern "C" int idg_copy_rows_strided(const float* d_src, int32_t src_stride, const int64_t* d_idx, int32_t n, int32_t row_offset, int32_t d,
                                     float* d_dst, int32_t dst_stride, void* stream) {
    if (!d_src || !d_idx || !d_dst || n < 0 || d <= 0 || (d & 3) || (src_stride & 3) || (dst_stride & 3)) return fail(-1, "idg_copy_rows_strided: bad argument%s");
    if (n == 0) return 0;
    const int th = n * (d / 4);
    copy_rows_strided_kernel<<<(th + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_src, src_stride, d_idx, n, row_offset, d / 4, d_dst, dst_stride);
    IDG_LAUNCH_CHECK("copy_rows_strided_kernel");
    return 0;
}

static int ngcf_fwd_impl(const float* d_E, const float* d_side, const float* d_Wg, const float* d_bg, const float* d_Wb, const float* d_bb,
                         const float* d_keep, const uint32_t* d_keep_bits, float drop_p, int32_t N, float* d_S, float* d_D, float* d_out, int32_t out_stride,
                         void* stream) {
    // d_S (the pre-activation) may be null: only the CUDA-core backward (IDG_NGCF_BWD=fma) reads it, the tensor-core one takes the sign off D
    if (!d_E || !d_side || !d_Wg || !d_bg || !d_Wb || !d_bb || !d_D || !d_out || N <= 0) return fail(-1, "idg_ngcf_dense_fwd: bad argument%s");
    if (drop_p < 0.f || drop_p >= 1.f) return fail(-1, "idg_ngcf_dense_fwd: drop_p must be in [0,1)%s");
    if (out_stride & 3) return fail(-1, "idg_ngcf_dense_fwd: out_stride must be a multiple of 4%s");
    return ngcf_dense_fwd_tc(d_E, d_side, d_Wg, d_bg, d_Wb, d_bb, d_keep, d_keep_bits, 1.f / (1.f - drop_p), N, d_S, d_D, d_out, out_stride, (cudaStream_t)stream);
}

extern "C" int idg_ngcf_dense_fwd(const float* d_E, const float* d_side, const float* d_Wg, const float* d_bg, const float* d_Wb,
                                  const float* d_bb, const float* d_keep, float drop_p, int32_t N, float* d_S, float* d_D, float* d_out,
                                  int32_t out_stride, void* stream) {
    return ngcf_fwd_impl(d_E, d_side, d_Wg, d_bg, d_Wb, d_bb, d_keep, nullptr, drop_p, N, d_S, d_D, d_out, out_stride, stream);
}

extern "C" int idg_ngcf_dense_fwd_bits(const float* d_E, const float* d_side, const float* d_Wg, const float* d_bg, const float* d_Wb,
                                       const float* d_bb, const uint32_t* d_keep_bits, float drop_p, int32_t N, float* d_S, float* d_D, float* d_out,
                                       int32_t out_stride, void* stream) {
    if (!d_keep_bits) return fail(-1, "idg_ngcf_dense_fwd_bits: null mask%s");
    return ngcf_fwd_impl(d_E, d_side, d_Wg, d_bg, d_Wb, d_bb, nullptr, d_keep_bits, drop_p, N, d_S, d_D, d_out, out_stride, stream);
}

extern "C" int64_t idg_ngcf_workspace_bytes(void) { return (int64_t)sizeof(float) * kNgCtas * (128 * 64 + 64); }

static int ngcf_bwd_impl(const float* d_E, const float* d_side, const float* d_Wg, const float* d_Wb, const float* d_keep, const uint32_t* d_keep_bits,
                         float drop_p, const float* d_S, const float* d_D, const float* d_dO, int32_t dO_stride, const float* d_dD_ext, int32_t N,
                         float* d_dside, float* d_dE_direct, float* d_dWg, float* d_dWb, float* d_db, void* d_ws, void* stream_) {
    if (!d_E || !d_side || !d_Wg || !d_Wb || !d_D || !d_dO || !d_dside || !d_dE_direct || !d_dWg || !d_dWb || !d_db || !d_ws || N <= 0)
        return fail(-1, "idg_ngcf_dense_bwd: bad argument%s");
    cudaStream_t stream = (cudaStream_t)stream_;
    float* dW_part = (float*)d_ws;
    float* db_part = dW_part + (size_t)kNgCtas * 128 * 64;
    // IDG_NGCF_BWD=fma selects the CUDA-core tiles (kept as a cross-check of the tensor-core kernel)
    static const bool use_tc = !(getenv("IDG_NGCF_BWD") && strcmp(getenv("IDG_NGCF_BWD"), "fma") == 0);
    if (dO_stride & 3) return fail(-1, "idg_ngcf_dense_bwd: dO_stride must be a multiple of 4%s");
    if (!use_tc && !d_S) return fail(-1, "idg_ngcf_dense_bwd: the CUDA-core kernel needs the pre-activation d_S%s");
    if (!use_tc && d_keep_bits) return fail(-1, "idg_ngcf_dense_bwd_bits: the CUDA-core kernel takes the float mask%s");
    int n_parts = kNgCtas;
    if (use_tc) {
        if (int rc = ngcf_dense_bwd_tc(d_E, d_side, d_Wg, d_Wb, d_keep, d_keep_bits, 1.f / (1.f - drop_p), d_S, d_D, d_dO, dO_stride, d_dD_ext, N, d_dside,
                                       d_dE_direct, dW_part, db_part, kNgCtas, &n_parts, stream))
            return rc;
    } else {
        const size_t smem = sizeof(float) * (kNgTile * 128 + kNgTile * 64 + 64 * 128);
        IDG_CUDA(cudaFuncSetAttribute(ngcf_dense_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ngcf_dense_bwd_kernel<<<kNgCtas, 256, smem, stream>>>(d_E, d_side, d_Wg, d_Wb, d_keep, 1.f / (1.f - drop_p), d_S, d_D, d_dO, dO_stride, d_dD_ext, N,
                                                              d_dside, d_dE_direct, dW_part, db_part);
        IDG_LAUNCH_CHECK("ngcf_dense_bwd_kernel");
    }
    ngcf_reduce_kernel<<<(128 * 64 + 64 + 63) / 64, 256, 0, stream>>>(dW_part, db_part, n_parts, d_dWg, d_dWb, d_db);
    IDG_LAUNCH_CHECK("ngcf_reduce_kernel");
    return 0;
}

extern "C" int idg_ngcf_dense_bwd(const float* d_E, const float* d_side, const float* d_Wg, const float* d_Wb, const float* d_keep, float drop_p,
                                  const float* d_S, const float* d_D, const float* d_dO, int32_t dO_stride, const float* d_dD_ext, int32_t N,
                                  float* d_dside, float* d_dE_direct, float* d_dWg, float* d_dWb, float* d_db, void* d_ws, void* stream) {
    return ngcf_bwd_impl(d_E, d_side, d_Wg, d_Wb, d_keep, nullptr, drop_p, d_S, d_D, d_dO, dO_stride, d_dD_ext, N, d_dside, d_dE_direct, d_dWg, d_dWb, d_db,
                         d_ws, stream);
}

extern "C" int idg_ngcf_dense_bwd_bits(const float* d_E, const float* d_side, const float* d_Wg, const float* d_Wb, const uint32_t* d_keep_bits,
                                       float drop_p, const float* d_D, const float* d_dO, int32_t dO_stride, const float* d_dD_ext, int32_t N,
                                       float* d_dside, float* d_dE_direct, float* d_dWg, float* d_dWb, float* d_db, void* d_ws, void* stream) {
    if (!d_keep_bits) return fail(-1, "idg_ngcf_dense_bwd_bits: null mask%s");
    return ngcf_bwd_impl(d_E, d_side, d_Wg, d_Wb, nullptr, d_keep_bits, drop_p, nullptr, d_D, d_dO, dO_stride, d_dD_ext, N, d_dside, d_dE_direct, d_dWg, d_dWb,
                         d_db, d_ws, stream);
}
