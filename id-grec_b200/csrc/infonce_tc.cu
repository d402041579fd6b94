// In-batch InfoNCE contractions on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// The dense work of utility_function/losses.py:24-35 and its backward is four n x n x 64 contractions
// (n = unique users / items of the batch, ~2,000):
//     E  = exp(A B^T / tau)            row sums -> ttl_i          (scores kernel, X = A, Y = B)
//     E' = exp(B A^T / tau) = E^T                                  (scores kernel, X = B, Y = A)
//     PB = E  . B                      sum_j e_ij b_j              (gemm kernel)
//     QA = E' . (beta * A)             sum_i beta_i e_ij a_i       (gemm kernel)
// tcgen05 has no fp32 input kind, and tf32 alone (2^-10) would miss the 1e-5 parity bar, so every product is
// the 3xTF32 split  x*y ~= xh*yh + xh*yl + xl*yh  accumulated in the same TMEM tile (error ~2^-21).  E / E' are
// kept as (hi, lo) fp32 pairs in L2-resident scratch (2 x 16 MB at n = 2048) so the second pair of contractions
// reads exact splits.  Same pipeline as csrc/eval_tc.cu: cp.async loaders writing the 128B-swizzled K-major
// layout, one MMA-issuing thread, four epilogue warps (thread <-> TMEM lane).
#include <math.h>

#include "tc_common.cuh"

namespace idg {

constexpr int kNtLoaders = 64;
constexpr int kNtSplits = 4;           // column splits of the csrc/pairloss.cu entry points (partials are summed in split order)
constexpr uint32_t kAtom128 = 128 * 128;  // bytes: [128 rows x 128 B]
constexpr uint32_t kAtom64 = 64 * 128;    // bytes: [ 64 rows x 128 B]

__device__ __forceinline__ uint32_t sw_atom(int r, int cc) {
    return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((cc ^ (r & 7)) << 4);
}

// ---------------------------------------------------------------------------------------------------
// (hi, lo) splits of the normalised rows; rows >= n are zero.  One warp per row (d = 64).
__global__ void __launch_bounds__(256) nce_tc_split_kernel(const float* __restrict__ A, const float* __restrict__ B, const int* __restrict__ d_n,
                                                           int n_in, int n_pad, float* __restrict__ AH, float* __restrict__ AL,
                                                           float* __restrict__ BH, float* __restrict__ BL) {
    const int n = d_n ? *d_n : n_in;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n_pad) return;
    const size_t o = (size_t)i * 64 + lane * 2;
    float2 a = make_float2(0.f, 0.f), b = a;
    if (i < n) { a = *reinterpret_cast<const float2*>(A + o); b = *reinterpret_cast<const float2*>(B + o); }
    float2 ah, al, bh, bl;
    split_tf32(a.x, ah.x, al.x); split_tf32(a.y, ah.y, al.y); split_tf32(b.x, bh.x, bl.x); split_tf32(b.y, bh.y, bl.y);
    *reinterpret_cast<float2*>(AH + o) = ah; *reinterpret_cast<float2*>(AL + o) = al;
    *reinterpret_cast<float2*>(BH + o) = bh; *reinterpret_cast<float2*>(BL + o) = bl;
}

// Yt(hi,lo)[k][c] = scale_c * Y[c][k] for c < n, 0 beyond: the K-major B operand of the gemm kernel
__global__ void __launch_bounds__(256) nce_tc_transpose_kernel(const float* __restrict__ Y, const float* __restrict__ scale,
                                                               const int* __restrict__ d_n, int n_in, int n_pad, float* __restrict__ TH,
                                                               float* __restrict__ TL) {
    __shared__ float tile[32][65];
    const int n = d_n ? *d_n : n_in;
    const int c0 = blockIdx.x * 32;
    for (int q = threadIdx.x; q < 32 * 64; q += 256) {
        const int c = q >> 6, k = q & 63;
        float v = 0.f;
        if (c0 + c < n) v = Y[(size_t)(c0 + c) * 64 + k] * (scale ? scale[c0 + c] : 1.f);
        tile[c][k] = v;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < 32 * 64; q += 256) {
        const int k = q >> 5, c = q & 31;
        float hi, lo;
        split_tf32(tile[c][k], hi, lo);
        TH[(size_t)k * n_pad + c0 + c] = hi;
        TL[(size_t)k * n_pad + c0 + c] = lo;
    }
}

// ---------------------------------------------------------------------------------------------------
// scores: E[r][c] = exp(<x_r, y_c>/tau) for r, c < n (0 elsewhere) as (hi, lo); optional row-sum partials.
// grid (n_pad/128, kNtSplits); 256 threads.  smem: X hi/lo 64 KB + 2 stages x (Y hi/lo 64 KB) = 192 KB.
// EPI 0: E = exp(S / tau) as (hi, lo) + row-sum partials (InfoNCE).  EPI 1: raw S into EH (one fp32 matrix, ld n_pad, 0 outside
// n x n).  EPI 2: EH += S (the S + R sum of the neighbourhood-aggregation losses, csrc/pairloss.cu).
template <int EPI>
__global__ void __launch_bounds__(256, 1) nce_tc_scores_kernel(const float* __restrict__ XH, const float* __restrict__ XL,
                                                               const float* __restrict__ YH, const float* __restrict__ YL,
                                                               const int* __restrict__ d_n, int n_in, int n_pad, float inv_tau,
                                                               float* __restrict__ EH, float* __restrict__ EL, float* __restrict__ part_sum,
                                                               float* __restrict__ ETH = nullptr, float* __restrict__ ETL = nullptr) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int n = d_n ? *d_n : n_in;
    const int r0 = blockIdx.x * 128;
    if (r0 >= n) return;
    unsigned char* sXH = smem;
    unsigned char* sXL = smem + 2 * kAtom128;
    unsigned char* sY = smem + 4 * kAtom128;   // per stage: YH (32 KB) then YL (32 KB)
    constexpr int S = 2, TB = 4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sY + S * 4 * kAtom128);
    uint64_t *full = bars, *empty = bars + S, *tfull = empty + S, *tempty = tfull + TB, *xfull = tempty + TB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xfull + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_total = (n + 127) / 128;
    const int t_begin = (int)((long long)tiles_total * blockIdx.y / gridDim.y), t_end = (int)((long long)tiles_total * (blockIdx.y + 1) / gridDim.y);
    const int ntiles = t_end - t_begin;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full + s, kNtLoaders); mbar_init(empty + s, 1); }
        for (int b = 0; b < TB; ++b) { mbar_init(tfull + b, 1); mbar_init(tempty + b, 128); }
        mbar_init(xfull, kNtLoaders);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0 && ntiles > 0) {
            const uint32_t idesc = umma_idesc_tf32(128, 128);
            mbar_wait(xfull, 0);
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % S, b = t % TB;
                mbar_wait(full + s, (t / S) & 1);
                mbar_wait(tempty + b, ((t / TB) & 1) ^ 1);
                tc_fence_after();
                const uint32_t xh = smem_u32(sXH), xl = smem_u32(sXL);
                const uint32_t yh = smem_u32(sY + (size_t)s * 4 * kAtom128), yl = yh + 2 * kAtom128;
                const uint32_t acc = tmem_base + (uint32_t)b * 128;
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const uint32_t xa = (p == 2) ? xl : xh, ya = (p == 1) ? yl : yh;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t koff = (uint32_t)(k >> 2) * kAtom128 + (uint32_t)(k & 3) * 32u;
                        umma_tf32(acc, umma_desc(xa + koff), umma_desc(ya + koff), idesc, (p | k) != 0);
                    }
                }
                umma_commit(empty + s);
                umma_commit(tfull + b);
            }
        }
    } else if (warp == 2 || warp == 3) {
        const int lt = tid - 64;
        if (ntiles > 0) {
            for (int c = lt; c < 128 * 16; c += kNtLoaders) {
                const int r = c >> 4, kc = c & 15;
                const uint32_t off = (uint32_t)(kc >> 3) * kAtom128 + sw_atom(r, kc & 7);
                cp_async16(smem_u32(sXH) + off, XH + (size_t)(r0 + r) * 64 + kc * 4, 16);
                cp_async16(smem_u32(sXL) + off, XL + (size_t)(r0 + r) * 64 + kc * 4, 16);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(xfull);
        }
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % S;
            mbar_wait(empty + s, ((t / S) & 1) ^ 1);
            const uint32_t dst = smem_u32(sY + (size_t)s * 4 * kAtom128);
            const int c0 = (t_begin + t) * 128;
#pragma unroll 4
            for (int c = lt; c < 128 * 16; c += kNtLoaders) {
                const int r = c >> 4, kc = c & 15;
                const uint32_t off = (uint32_t)(kc >> 3) * kAtom128 + sw_atom(r, kc & 7);
                cp_async16(dst + off, YH + (size_t)(c0 + r) * 64 + kc * 4, 16);                  // rows < n_pad exist (zero beyond n)
                cp_async16(dst + 2 * kAtom128 + off, YL + (size_t)(c0 + r) * 64 + kc * 4, 16);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (t > 0) {
                asm volatile("cp.async.wait_group 1;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(full + (t - 1) % S);
            }
        }
        if (ntiles > 0) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(full + (ntiles - 1) % S);
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int r = r0 + q * 32 + lane;
        const bool rvalid = r < n;
        const uint32_t lane_base = ((uint32_t)(q * 32)) << 16;
        float rowsum = 0.f;
        for (int t = 0; t < ntiles; ++t) {
            const int b = t % TB;
            mbar_wait(tfull + b, (t / TB) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t raw[32];
                tmem_ld32(tmem_base + lane_base + (uint32_t)(b * 128 + c * 32), raw);
                tmem_ld_wait();
                if (c == 3) { tc_fence_before(); mbar_arrive(tempty + b); }
                const int c0 = (t_begin + t) * 128 + c * 32;
                float4* eh = reinterpret_cast<float4*>(EH + (size_t)r * n_pad + c0);
                float4* el = reinterpret_cast<float4*>(EL + (size_t)r * n_pad + c0);
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    if (EPI == 0) {
                        float h[4], l[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int j = j4 * 4 + jj;
                            const float e = (rvalid && c0 + j < n) ? expf(__uint_as_float(raw[j]) * inv_tau) : 0.f;
                            rowsum += e;
                            split_tf32(e, h[jj], l[jj]);
                        }
                        eh[j4] = make_float4(h[0], h[1], h[2], h[3]);
                        el[j4] = make_float4(l[0], l[1], l[2], l[3]);
                        if (ETH) {
                            // E^T for the second gradient contraction, from the same accumulators (exp(B A^T) = E^T exactly as computed
                            // here): lanes hold consecutive rows r, so each scalar store of a warp covers 128 contiguous bytes of row c
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const size_t o = (size_t)(c0 + j4 * 4 + jj) * n_pad + r;
                                ETH[o] = h[jj]; ETL[o] = l[jj];
                            }
                        }
                    } else {
                        float v[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int j = j4 * 4 + jj;
                            v[jj] = (rvalid && c0 + j < n) ? __uint_as_float(raw[j]) : 0.f;
                        }
                        float4 o = make_float4(v[0], v[1], v[2], v[3]);
                        if (EPI == 2) { const float4 old = eh[j4]; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                        eh[j4] = o;
                    }
                }
            }
        }
        if (part_sum && rvalid) part_sum[(size_t)blockIdx.y * n_pad + r] = rowsum;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// gemm: part[split][r][0..63] = sum over the split's columns c of E[r][c] * Yt[:, c]   (3xTF32)
// grid (n_pad/128, kNtSplits).  Per K chunk of 32 columns: E hi/lo 2 x 16 KB + Yt hi/lo 2 x 8 KB = 48 KB; 4 stages.
__global__ void __launch_bounds__(256, 1) nce_tc_gemm_kernel(const float* __restrict__ EH, const float* __restrict__ EL,
                                                             const float* __restrict__ TH, const float* __restrict__ TL,
                                                             const int* __restrict__ d_n, int n_in, int n_pad, float* __restrict__ part) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int n = d_n ? *d_n : n_in;
    const int r0 = blockIdx.x * 128;
    if (r0 >= n) return;
    constexpr int S = 4;
    constexpr uint32_t kStage = 2 * kAtom128 + 2 * kAtom64;  // 48 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * kStage);
    uint64_t *full = bars, *empty = bars + S, *tfull = empty + S;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunks_total = (n + 31) / 32;
    const int k_begin = (int)((long long)chunks_total * blockIdx.y / gridDim.y), k_end = (int)((long long)chunks_total * (blockIdx.y + 1) / gridDim.y);
    const int nchunks = k_end - k_begin;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full + s, kNtLoaders); mbar_init(empty + s, 1); }
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
        if (lane == 0 && nchunks > 0) {
            const uint32_t idesc = umma_idesc_tf32(128, 64);
            for (int t = 0; t < nchunks; ++t) {
                const int s = t % S;
                mbar_wait(full + s, (t / S) & 1);
                tc_fence_after();
                const uint32_t ah = smem_u32(smem + (size_t)s * kStage), al = ah + kAtom128, bh = al + kAtom128, bl = bh + kAtom64;
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const uint32_t aa = (p == 2) ? al : ah, bb = (p == 1) ? bl : bh;
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_tf32(tmem_base, umma_desc(aa + k * 32u), umma_desc(bb + k * 32u), idesc, (t | p | k) != 0);
                }
                umma_commit(empty + s);
            }
            umma_commit(tfull);
        }
    } else if (warp == 2 || warp == 3) {
        const int lt = tid - 64;
        for (int t = 0; t < nchunks; ++t) {
            const int s = t % S;
            mbar_wait(empty + s, ((t / S) & 1) ^ 1);
            const uint32_t dst = smem_u32(smem + (size_t)s * kStage);
            const int kc0 = (k_begin + t) * 32;  // first column of this K chunk
            for (int c = lt; c < 128 * 8; c += kNtLoaders) {
                const int r = c >> 3, cc = c & 7;
                const size_t src = (size_t)(r0 + r) * n_pad + kc0 + cc * 4;
                cp_async16(dst + sw_atom(r, cc), EH + src, 16);
                cp_async16(dst + kAtom128 + sw_atom(r, cc), EL + src, 16);
            }
            for (int c = lt; c < 64 * 8; c += kNtLoaders) {
                const int r = c >> 3, cc = c & 7;
                const size_t src = (size_t)r * n_pad + kc0 + cc * 4;
                cp_async16(dst + 2 * kAtom128 + sw_atom(r, cc), TH + src, 16);
                cp_async16(dst + 2 * kAtom128 + kAtom64 + sw_atom(r, cc), TL + src, 16);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (t > 1) {  // three chunks in flight
                asm volatile("cp.async.wait_group 2;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(full + (t - 2) % S);
            }
        }
        if (nchunks > 1) {
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(full + (nchunks - 2) % S);
        }
        if (nchunks > 0) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(full + (nchunks - 1) % S);
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int r = r0 + q * 32 + lane;
        const uint32_t lane_base = ((uint32_t)(q * 32)) << 16;
        float* out = part + ((size_t)blockIdx.y * n_pad + r) * 64;
        if (nchunks > 0) {
            mbar_wait(tfull, 0);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t raw[32];
                tmem_ld32(tmem_base + lane_base + (uint32_t)(c * 32), raw);
                tmem_ld_wait();
                if (r < n) {
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4)
                        reinterpret_cast<float4*>(out + c * 32)[j4] = make_float4(__uint_as_float(raw[4 * j4]), __uint_as_float(raw[4 * j4 + 1]),
                                                                                 __uint_as_float(raw[4 * j4 + 2]), __uint_as_float(raw[4 * j4 + 3]));
                }
            }
        } else if (r < n) {
            for (int j4 = 0; j4 < 16; ++j4) reinterpret_cast<float4*>(out)[j4] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
size_t nce_tc_extra_bytes(int n_max) {
    const size_t np = ((size_t)n_max + 127) / 128 * 128;
    return 8 * np * 64 * sizeof(float) + 4 * np * np * sizeof(float) + 4096;
}

// A, Bm: normalised rows [n,64]; beta [n] or NULL until known.  Stage 0: splits + E + row-sum partials.
// Stage 1 (after the row kernel produced beta): E', transposes, the two gemms into part_pb / part_qa.
int nce_tc_stage(int stage, const float* A, const float* Bm, const float* beta, const int* d_n, int n_max, float inv_tau, float* part_sum,
                 float* part_pb, float* part_qa, void* extra, cudaStream_t stream, int want_grad, int splits) {
    const int np = (n_max + 127) / 128 * 128;
    float* p = (float*)(((uintptr_t)extra + 1023) & ~(uintptr_t)1023);
    float *AH = p, *AL = AH + (size_t)np * 64, *BH = AL + (size_t)np * 64, *BL = BH + (size_t)np * 64;
    float *BtH = BL + (size_t)np * 64, *BtL = BtH + (size_t)np * 64, *AtH = BtL + (size_t)np * 64, *AtL = AtH + (size_t)np * 64;
    float *E1H = AtL + (size_t)np * 64, *E1L = E1H + (size_t)np * np, *E2H = E1L + (size_t)np * np, *E2L = E2H + (size_t)np * np;
    const size_t smem_s = 4 * kAtom128 + 2 * 4 * kAtom128 + 128, smem_g = 4 * (2 * kAtom128 + 2 * kAtom64) + 128;
    // `splits` column splits (== kNceSplits of csrc/infonce.cu): 16 row tiles x 8 = 128 CTAs at n ~ 2,000 (4 splits left 84 of
    // the 148 SMs idle and measured 41 us per scores launch)
    const dim3 grid(np / 128, splits);
    if (stage == 0) {
        nce_tc_split_kernel<<<(np + 7) / 8, 256, 0, stream>>>(A, Bm, d_n, n_max, np, AH, AL, BH, BL);
        IDG_LAUNCH_CHECK("nce_tc_split_kernel");
        IDG_CUDA(cudaFuncSetAttribute(nce_tc_scores_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
        // one pass writes E and (when gradients follow) E^T: round 1 recomputed exp(B A^T) with a second launch of this kernel
        nce_tc_scores_kernel<0><<<grid, 256, smem_s, stream>>>(AH, AL, BH, BL, d_n, n_max, np, inv_tau, E1H, E1L, part_sum, want_grad ? E2H : nullptr,
                                                              want_grad ? E2L : nullptr);
        IDG_LAUNCH_CHECK("nce_tc_scores_kernel");
        return 0;
    }
    nce_tc_transpose_kernel<<<np / 32, 256, 0, stream>>>(Bm, nullptr, d_n, n_max, np, BtH, BtL);
    IDG_LAUNCH_CHECK("nce_tc_transpose_kernel");
    nce_tc_transpose_kernel<<<np / 32, 256, 0, stream>>>(A, beta, d_n, n_max, np, AtH, AtL);
    IDG_LAUNCH_CHECK("nce_tc_transpose_kernel");
    IDG_CUDA(cudaFuncSetAttribute(nce_tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
    nce_tc_gemm_kernel<<<grid, 256, smem_g, stream>>>(E1H, E1L, BtH, BtL, d_n, n_max, np, part_pb);
    IDG_LAUNCH_CHECK("nce_tc_gemm_kernel");
    nce_tc_gemm_kernel<<<grid, 256, smem_g, stream>>>(E2H, E2L, AtH, AtL, d_n, n_max, np, part_qa);
    IDG_LAUNCH_CHECK("nce_tc_gemm_kernel");
    return 0;
}

// ---- entry points for csrc/pairloss.cu (same kernels, plain n x n x 64 contractions) ------------------------------------------
// (hi, lo) splits of two [n,64] operands, rows >= n zero (n_pad rows each)
int tc_split_rows(const float* A, const float* B, int n, int np, float* AH, float* AL, float* BH, float* BL, cudaStream_t stream) {
    nce_tc_split_kernel<<<(np + 7) / 8, 256, 0, stream>>>(A, B, nullptr, n, np, AH, AL, BH, BL);
    IDG_LAUNCH_CHECK("nce_tc_split_kernel");
    return 0;
}

// out[n_pad, n_pad] (=|+=) X Y^T on the n x n block, 0 (or unchanged + 0) outside; X, Y given as splits
int tc_scores_raw(const float* XH, const float* XL, const float* YH, const float* YL, int n, int np, float* out, int accumulate, cudaStream_t stream) {
    const size_t smem_s = 4 * kAtom128 + 2 * 4 * kAtom128 + 128;
    const dim3 grid(np / 128, kNtSplits);
    if (accumulate) {
        IDG_CUDA(cudaFuncSetAttribute(nce_tc_scores_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
        nce_tc_scores_kernel<2><<<grid, 256, smem_s, stream>>>(XH, XL, YH, YL, nullptr, n, np, 1.f, out, nullptr, nullptr);
    } else {
        IDG_CUDA(cudaFuncSetAttribute(nce_tc_scores_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
        nce_tc_scores_kernel<1><<<grid, 256, smem_s, stream>>>(XH, XL, YH, YL, nullptr, n, np, 1.f, out, nullptr, nullptr);
    }
    IDG_LAUNCH_CHECK("nce_tc_scores_kernel");
    return 0;
}

// part[kNtSplits][n_pad][64]: split-wise partial sums of E . Yt^T  (E [n_pad,n_pad] and Yt [64,n_pad] as splits, zero outside n)
int tc_gemm64(const float* EH, const float* EL, const float* TH, const float* TL, int n, int np, float* part, cudaStream_t stream) {
    const size_t smem_g = 4 * (2 * kAtom128 + 2 * kAtom64) + 128;
    IDG_CUDA(cudaFuncSetAttribute(nce_tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
    nce_tc_gemm_kernel<<<dim3(np / 128, kNtSplits), 256, smem_g, stream>>>(EH, EL, TH, TL, nullptr, n, np, part);
    IDG_LAUNCH_CHECK("nce_tc_gemm_kernel");
    return 0;
}

}  // namespace idg
