// NGCF dense layer, forward, on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   S = [side | E (*) side] . [W_gcn ; W_bi] + b_gcn + b_bi            (models/NGCF.py:87-97; [N,128] x [128,64])
//   D = LeakyReLU_0.2(S) (*) keep / (1-p);   O = D / max(|D|_2, 1e-12)   (NGCF.py:98-106)
//
// The product is only 64 columns wide: 2.36 GFLOP against 185 MB of [N,64] streams per layer at the amazon-book shape, i.e.
// stream-bound once it leaves the fp32 FMA pipe (the CUDA-core tiles of csrc/ngcf.cu run at 16-18 TFLOP/s = 130 us per layer;
// the warp-level mma.sync tf32 path was measured slower than that).  tcgen05 has no fp32 input kind, so every product is the
// 3xTF32 split  x*y ~= xh*yh + xh*yl + xl*yh  accumulated in one TMEM tile (error ~2^-21, inside the 1e-5 parity bar) -- the
// same scheme and the same pipeline skeleton as nce_tc_gemm_kernel (csrc/infonce_tc.cu): one CTA per 128-row tile, K = 128 in
// four chunks of 32 through a 2-stage shared-memory ring, loader warps that BUILD the operands (Z = [side | E*side] and the
// transposed weights are split and stored in the 128B-swizzled K-major layout with plain stores + fence.proxy.async), one
// MMA-issuing thread, four epilogue warps (thread <-> TMEM lane <-> row: bias, LeakyReLU, dropout, row norm without shuffles).
#include <math.h>

#include "tc_common.cuh"

namespace idg {

constexpr int kNgLoaders = 224;           // warps 1..7 build the operands (warps 4..7 then run the epilogue)
constexpr uint32_t kNgAtomA = 128 * 128;   // bytes: [128 rows x 128 B]  (32 k-values per row)
constexpr uint32_t kNgAtomB = 64 * 128;    // bytes: [ 64 rows x 128 B]
constexpr uint32_t kNgStage = 2 * kNgAtomA + 2 * kNgAtomB;   // A hi/lo + B hi/lo = 48 KB
constexpr int kNgStages = 2;

__device__ __forceinline__ uint32_t ng_sw_atom(int r, int cc) {
    return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((cc ^ (r & 7)) << 4);
}
__device__ __forceinline__ void ng_st_shared4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(256, 2) ngcf_dense_fwd_tc_kernel(const float* __restrict__ E, const float* __restrict__ side,
                                                                   const float* __restrict__ Wg, const float* __restrict__ bg,
                                                                   const float* __restrict__ Wb, const float* __restrict__ bb,
                                                                   const float* __restrict__ keep, float inv_keep, int N,
                                                                   float* __restrict__ S_pre, float* __restrict__ D, float* __restrict__ out,
                                                                   int out_stride) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int S = kNgStages, kChunks = 4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * kNgStage);
    uint64_t *full = bars, *empty = bars + S, *tfull = empty + S;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = blockIdx.x * 128;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full + s, kNgLoaders); mbar_init(empty + s, 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(128, 64);
            for (int t = 0; t < kChunks; ++t) {
                const int s = t % S;
                mbar_wait(full + s, (t / S) & 1);
                tc_fence_after();
                const uint32_t ah = smem_u32(smem + (size_t)s * kNgStage), al = ah + kNgAtomA, bh = al + kNgAtomA, bl = bh + kNgAtomB;
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const uint32_t aa = (p == 2) ? al : ah, bb2 = (p == 1) ? bl : bh;
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_tf32(tmem_base, umma_desc(aa + k * 32u), umma_desc(bb2 + k * 32u), idesc, (t | p | k) != 0);
                }
                umma_commit(empty + s);
            }
            umma_commit(tfull);
        }
    } else {
        // operand builders (warps 1..7): chunk t covers k = 32 t .. 32 t + 31 of Z = [side | E*side] (rows) and of Wcat^T (the 64
        // output columns).  All global loads of a chunk are issued before the first store so that they overlap.
        const int bt = tid - 32;
        for (int t = 0; t < kChunks; ++t) {
            const int s = t % S;
            const uint32_t dst = smem_u32(smem + (size_t)s * kNgStage);
            const int k0 = (t & 1) * 32;          // column offset inside side / E
            const bool second = t >= 2;           // chunks 2, 3: the E (*) side half
            const float* Wsrc = second ? Wb : Wg;   // Wcat rows 0..63 = W_gcn, 64..127 = W_bi; B operand row n holds Wcat[k][n] over k
            float4 zs[5], es[5];
            float wv[3][4];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int c = bt + i * kNgLoaders;
                const int r = c >> 3, cc = c & 7;
                zs[i] = f4zero(); es[i] = make_float4(1.f, 1.f, 1.f, 1.f);
                if (c < 128 * 8 && r0 + r < N) {
                    zs[i] = ldg4(side + (size_t)(r0 + r) * 64 + k0 + cc * 4);
                    if (second) es[i] = ldg4(E + (size_t)(r0 + r) * 64 + k0 + cc * 4);
                }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int c = bt + i * kNgLoaders;
                const int n = c >> 3, cc = c & 7;
#pragma unroll
                for (int j = 0; j < 4; ++j) wv[i][j] = (c < 64 * 8) ? __ldg(Wsrc + (size_t)(k0 + cc * 4 + j) * 64 + n) : 0.f;
            }
            mbar_wait(empty + s, ((t / S) & 1) ^ 1);
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int c = bt + i * kNgLoaders;
                if (c < 128 * 8) {
                    const int r = c >> 3, cc = c & 7;
                    const float z[4] = {zs[i].x * es[i].x, zs[i].y * es[i].y, zs[i].z * es[i].z, zs[i].w * es[i].w};
                    float h[4], l[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) split_tf32(z[j], h[j], l[j]);
                    const uint32_t o = ng_sw_atom(r, cc);
                    ng_st_shared4(dst + o, h[0], h[1], h[2], h[3]);
                    ng_st_shared4(dst + kNgAtomA + o, l[0], l[1], l[2], l[3]);
                }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int c = bt + i * kNgLoaders;
                if (c < 64 * 8) {
                    const int n = c >> 3, cc = c & 7;
                    float h[4], l[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) split_tf32(wv[i][j], h[j], l[j]);
                    const uint32_t o = ng_sw_atom(n, cc);
                    ng_st_shared4(dst + 2 * kNgAtomA + o, h[0], h[1], h[2], h[3]);
                    ng_st_shared4(dst + 2 * kNgAtomA + kNgAtomB + o, l[0], l[1], l[2], l[3]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(full + s);
        }
    }
    if (warp >= 4) {
        const int q = warp & 3;
        const int r = r0 + q * 32 + lane;
        const uint32_t lane_base = ((uint32_t)(q * 32)) << 16;
        mbar_wait(tfull, 0);
        tc_fence_after();
        float sv[64];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t raw[32];
            tmem_ld32(tmem_base + lane_base + (uint32_t)(c * 32), raw);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) sv[c * 32 + j] = __uint_as_float(raw[j]);
        }
        if (r < N) {
            float ss = 0.f;
            float* ps = S_pre + (size_t)r * 64;
            float* pd = D + (size_t)r * 64;
            float* po = out + (size_t)r * out_stride;
#pragma unroll
            for (int j4 = 0; j4 < 16; ++j4) {
                const float4 b1 = ldg4(bg + j4 * 4), b2 = ldg4(bb + j4 * 4);
                float4 kp = make_float4(1.f, 1.f, 1.f, 1.f);
                if (keep) kp = ldg4(keep + (size_t)r * 64 + j4 * 4);
                const float bias[4] = {b1.x + b2.x, b1.y + b2.y, b1.z + b2.z, b1.w + b2.w}, kv[4] = {kp.x, kp.y, kp.z, kp.w};
                float s4[4], d4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s4[j] = sv[j4 * 4 + j] + bias[j];
                    const float act = s4[j] > 0.f ? s4[j] : 0.2f * s4[j];
                    d4[j] = keep ? act * kv[j] * inv_keep : act;
                    ss = fmaf(d4[j], d4[j], ss);
                    sv[j4 * 4 + j] = d4[j];
                }
                st4(ps + j4 * 4, make_float4(s4[0], s4[1], s4[2], s4[3]));
                st4(pd + j4 * 4, make_float4(d4[0], d4[1], d4[2], d4[3]));
            }
            const float nrm = fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
            for (int j4 = 0; j4 < 16; ++j4)
                st4(po + j4 * 4, make_float4(sv[j4 * 4] / nrm, sv[j4 * 4 + 1] / nrm, sv[j4 * 4 + 2] / nrm, sv[j4 * 4 + 3] / nrm));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
    }
}

// launched by idg_ngcf_dense_fwd (csrc/ngcf.cu) for the 64-wide layers of the reference configuration
int ngcf_dense_fwd_tc(const float* E, const float* side, const float* Wg, const float* bg, const float* Wb, const float* bb, const float* keep,
                      float inv_keep, int N, float* S_pre, float* D, float* out, int out_stride, cudaStream_t stream) {
    const size_t smem = (size_t)kNgStages * kNgStage + 128;
    IDG_CUDA(cudaFuncSetAttribute(ngcf_dense_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ngcf_dense_fwd_tc_kernel<<<(N + 127) / 128, 256, smem, stream>>>(E, side, Wg, bg, Wb, bb, keep, inv_keep, N, S_pre, D, out, out_stride);
    IDG_LAUNCH_CHECK("ngcf_dense_fwd_tc_kernel");
    return 0;
}

}  // namespace idg
