// InfoNCE contractions without the n x n matrices (losses.py:24-35 and its autograd), tcgen05 + TMEM, sm_100a.
//
//   out[r][:] = sum_c exp(<x_r, y_c> / tau) * v_c        (and, optionally, rowsum[r] = sum_c exp(<x_r, y_c> / tau))
//
// over the rows r of a 128-row tile and the columns c of this CTA's column split.  One call with (X, Y, V) = (A, B, B) gives
// the softmax denominators and P B = E B of the forward / dL/dA; the mirrored call (B, A, beta (*) A) gives E^T (beta A) of
// dL/dB.  E = exp(A B^T / tau) never leaves the SM: the score tile is produced in TMEM by one 3xTF32 MMA chain, the epilogue
// threads turn it into E (hi, lo) IN TMEM (tcgen05.ld -> exp -> split -> tcgen05.st), and a second MMA chain takes it from
// there as its A operand (TS form) against the V rows as an MN-major B operand -- the flash-attention structure, minus the
// running maximum (|<x, y>| <= 1 on normalised rows, so exp(s / tau) <= e^{1/tau} needs no rescaling).  Round 1 / early round 2
// wrote E and E^T as (hi, lo) pairs (4 n^2 floats) and read them back in two more launches (csrc/infonce_tc.cu, kept for the
// forward-only call and as the cross-check IDG_NCE_IMPL=split).
//
// grid (n_pad / 128, splits); warp 0 issues the MMAs, warps 1..16 build the operand images (16 lanes per row) and run the
// epilogues (thread <-> TMEM lane = row; four warps per lane quadrant = four 32-column quarters of the score tile).  Measured at
// n = 1,923: 65 us per InfoNCE call with 8 worker warps, 54 us with 16 (93 us with the materialised matrices).
#include <math.h>

#include "tc_common.cuh"

namespace idg {

constexpr int kFlWorkers = 512;               // 16 worker warps (8 left the SM issue-bound in the NGCF kernels of the same build)
constexpr int kFlPasses = 2048 / kFlWorkers;  // a [128 rows x 16 chunks] operand in passes of kFlWorkers / 16 rows
constexpr int kFlPassRows = kFlWorkers / 16;
constexpr uint32_t kFlBlk = 128 * 128;          // [128 rows x 32 fp32]
constexpr uint32_t kFlHalf = 2 * kFlBlk;        // hi (or lo) of a [128 x 64] operand
constexpr uint32_t kFlSmem = 6 * kFlHalf;       // X hi|lo (K-major) | Y hi|lo (K-major) | V hi|lo (MN-major) = 192 KB
// TMEM columns: scores [0,128) | E hi [128,256) | E lo [256,384) | out [384,448)
constexpr uint32_t kFlTmemCols = 512, kFlS = 0, kFlEh = 128, kFlEl = 256, kFlOut = 384;

__device__ __forceinline__ uint32_t fl_sw(int r, int cc) {     // 128 B-row block, 128-byte swizzle
    return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((cc ^ (r & 7)) << 4);
}
__device__ __forceinline__ uint32_t fl_sw32(int r, int cc) {   // 128 B-row block, 128-byte swizzle with 32-byte atoms (MN-major fp32)
    return (uint32_t)r * 128u + (uint32_t)((((cc >> 1) ^ (r & 3)) << 5) | ((cc & 1) << 4));
}
__device__ __forceinline__ void fl_st4(uint32_t addr, const float* v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}
__device__ __forceinline__ void fl_split(float x, float& hi, float& lo) {   // hi = x rounded to nearest tf32, lo = x - hi exactly
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    hi = __uint_as_float(r);
    lo = x - hi;
}

template <bool ROWSUM>
__global__ void __launch_bounds__(kFlWorkers + 32, 1) nce_flash_kernel(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ V,
                                                           const float* __restrict__ v_scale, const int* __restrict__ d_n, int n_in, int n_pad,
                                                           float inv_tau, float* __restrict__ part_sum, float* __restrict__ part_out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int n = d_n ? *d_n : n_in;
    const int r0 = blockIdx.x * 128;
    if (r0 >= n) return;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kFlSmem);
    uint64_t *x_full = bars, *y_full = bars + 1, *y_empty = bars + 2, *s_full = bars + 3, *e_full = bars + 4, *o_full = bars + 5;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
    float* s_row = reinterpret_cast<float*>(bars + 8);     // [3][128] row sums of the column quarters 1..3
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_total = (n + 127) / 128;
    const int t_begin = (int)((long long)tiles_total * blockIdx.y / gridDim.y), t_end = (int)((long long)tiles_total * (blockIdx.y + 1) / gridDim.y);
    const int ntiles = t_end - t_begin;

    if (tid == 0) {
        mbar_init(x_full, kFlWorkers); mbar_init(y_full, kFlWorkers); mbar_init(y_empty, 1); mbar_init(s_full, 1); mbar_init(e_full, kFlWorkers);
        mbar_init(o_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kFlTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t aX = smem_u32(smem), aY = aX + 2 * kFlHalf, aV = aY + 2 * kFlHalf;

    if (warp == 0) {
        if (lane == 0 && ntiles > 0) {
            const uint32_t idesc_s = umma_idesc_tf32(128, 128);                 // scores: X (K-major) . Y^T (K-major)
            const uint32_t idesc_o = umma_idesc_tf32(128, 64) | (1u << 16);     // out += E (TMEM) . V (MN-major)
            mbar_wait(x_full, 0);
            for (int t = 0; t < ntiles; ++t) {
                mbar_wait(y_full, t & 1);
                tc_fence_after();
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const uint32_t xa = aX + (p == 2 ? kFlHalf : 0u), ya = aY + (p == 1 ? kFlHalf : 0u);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t koff = (uint32_t)(k >> 2) * kFlBlk + (uint32_t)(k & 3) * 32u;
                        umma_tf32(tmem_base + kFlS, umma_desc(xa + koff), umma_desc(ya + koff), idesc_s, (p | k) != 0);
                    }
                }
                umma_commit(s_full);
                mbar_wait(e_full, t & 1);
                tc_fence_after();
#pragma unroll
                for (int p = 0; p < 3; ++p) {      // Eh Vh + Eh Vl + El Vh
                    const uint32_t ea = tmem_base + (p == 2 ? kFlEl : kFlEh), va = aV + (p == 1 ? kFlHalf : 0u);
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        umma_tf32_ts(tmem_base + kFlOut, ea + (uint32_t)k * 8u, umma_desc_mn(va + (uint32_t)k * 1024u, kFlBlk), idesc_o, (t | p | k) != 0);
                }
                umma_commit(y_empty);     // Y / V images, the score tile and E are free for the next column tile
            }
            umma_commit(o_full);
        }
    } else {
        const int bt = tid - 32, ty = bt >> 4, tx = bt & 15;
        const int q = warp & 3, cq = (warp - 1) >> 2;      // TMEM lane quadrant, column quarter (32 of the 128 score columns, 16 of the 64 out columns)
        const uint32_t tq = tmem_base + (((uint32_t)(q * 32)) << 16);
        const int row = q * 32 + lane;
        const bool rvalid = r0 + row < n;
        // X tile (hi, lo), once
#pragma unroll
        for (int ps = 0; ps < kFlPasses; ++ps) {
            const int rr = ps * kFlPassRows + ty;
            float4 x = f4zero();
            if (r0 + rr < n) x = ldg4(X + (size_t)(r0 + rr) * 64 + tx * 4);
            const float xv[4] = {x.x, x.y, x.z, x.w};
            float h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) fl_split(xv[j], h[j], l[j]);
            const uint32_t o = (uint32_t)(tx >> 3) * kFlBlk + fl_sw(rr, tx & 7);
            fl_st4(aX + o, h); fl_st4(aX + kFlHalf + o, l);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(x_full);
        float rowsum = 0.f;
        for (int t = 0; t < ntiles; ++t) {
            const int c0 = (t_begin + t) * 128;
            // ---- operand images of the column tile: Y rows K-major (scores), V rows MN-major (the contraction over these rows)
            float4 yv[kFlPasses], vv[kFlPasses];
#pragma unroll
            for (int ps = 0; ps < kFlPasses; ++ps) {
                const int c = c0 + ps * kFlPassRows + ty;
                yv[ps] = f4zero(); vv[ps] = f4zero();
                if (c < n) {
                    yv[ps] = ldg4(Y + (size_t)c * 64 + tx * 4);
                    if (V != Y || v_scale) {
                        vv[ps] = ldg4(V + (size_t)c * 64 + tx * 4);
                        if (v_scale) { const float sc = __ldg(v_scale + c); vv[ps].x *= sc; vv[ps].y *= sc; vv[ps].z *= sc; vv[ps].w *= sc; }
                    } else {
                        vv[ps] = yv[ps];
                    }
                }
            }
            mbar_wait(y_empty, (t & 1) ^ 1);
#pragma unroll
            for (int ps = 0; ps < kFlPasses; ++ps) {
                const int rr = ps * kFlPassRows + ty;
                const float a[4] = {yv[ps].x, yv[ps].y, yv[ps].z, yv[ps].w}, b[4] = {vv[ps].x, vv[ps].y, vv[ps].z, vv[ps].w};
                float h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) fl_split(a[j], h[j], l[j]);
                const uint32_t ok = (uint32_t)(tx >> 3) * kFlBlk + fl_sw(rr, tx & 7);
                fl_st4(aY + ok, h); fl_st4(aY + kFlHalf + ok, l);
#pragma unroll
                for (int j = 0; j < 4; ++j) fl_split(b[j], h[j], l[j]);
                const uint32_t om = (uint32_t)(tx >> 3) * kFlBlk + fl_sw32(rr, tx & 7);
                fl_st4(aV + om, h); fl_st4(aV + kFlHalf + om, l);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(y_full);
            // ---- scores -> E (hi, lo) in TMEM; row sums
            mbar_wait(s_full, t & 1);
            tc_fence_after();
            {
                const uint32_t col = (uint32_t)(cq * 32);
                uint32_t raw[32], lo[32];
                tmem_ld32(tq + kFlS + col, raw);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float e = (rvalid && c0 + (int)col + j < n) ? expf(__uint_as_float(raw[j]) * inv_tau) : 0.f;
                    if (ROWSUM) rowsum += e;
                    float h, l;
                    fl_split(e, h, l);
                    raw[j] = __float_as_uint(h); lo[j] = __float_as_uint(l);
                }
                tmem_st32(tq + kFlEh + col, raw);
                tmem_st32(tq + kFlEl + col, lo);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(e_full);
        }
        // ---- out partial (thread <-> row, this warp's 32 of the 64 columns) and the row sums (two column halves per row)
        if (ntiles > 0) {
            mbar_wait(o_full, 0);
            tc_fence_after();
        }
        if (ROWSUM && cq > 0) s_row[(cq - 1) * 128 + row] = rowsum;
        asm volatile("bar.sync 1, %0;" ::"n"(kFlWorkers) : "memory");
        float* out = part_out + ((size_t)blockIdx.y * n_pad + r0 + row) * 64 + cq * 16;
        if (ntiles > 0) {
            uint32_t raw[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(raw[0]), "=r"(raw[1]), "=r"(raw[2]), "=r"(raw[3]), "=r"(raw[4]), "=r"(raw[5]), "=r"(raw[6]), "=r"(raw[7]), "=r"(raw[8]), "=r"(raw[9]),
                  "=r"(raw[10]), "=r"(raw[11]), "=r"(raw[12]), "=r"(raw[13]), "=r"(raw[14]), "=r"(raw[15])
                : "r"(tq + kFlOut + (uint32_t)(cq * 16))
                : "memory");
            tmem_ld_wait();
            if (rvalid) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4)
                    st4(out + j4 * 4, make_float4(__uint_as_float(raw[4 * j4]), __uint_as_float(raw[4 * j4 + 1]), __uint_as_float(raw[4 * j4 + 2]),
                                                  __uint_as_float(raw[4 * j4 + 3])));
            }
        } else if (rvalid) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) st4(out + j4 * 4, f4zero());
        }
        if (ROWSUM && cq == 0 && rvalid) part_sum[(size_t)blockIdx.y * n_pad + r0 + row] = ((rowsum + s_row[row]) + s_row[128 + row]) + s_row[256 + row];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kFlTmemCols) : "memory");
    }
}

// part_out[split][r][0..63] = sum over the split's columns c < n of exp(<X_r, Y_c> * inv_tau) * v_scale[c] * V_c; with part_sum also
// the row sums of the exponentials.  X, Y, V: [n, 64] rows (rows >= n are never read); partial buffers have row stride n_pad.
int nce_flash(const float* X, const float* Y, const float* V, const float* v_scale, const int* d_n, int n_max, float inv_tau, float* part_sum,
              float* part_out, int splits, cudaStream_t stream) {
    const int np = (n_max + 127) / 128 * 128;
    const dim3 grid(np / 128, splits);
    const size_t smem = (size_t)kFlSmem + 2048;
    if (part_sum) {
        IDG_CUDA(cudaFuncSetAttribute(nce_flash_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        nce_flash_kernel<true><<<grid, kFlWorkers + 32, smem, stream>>>(X, Y, V, v_scale, d_n, n_max, np, inv_tau, part_sum, part_out);
    } else {
        IDG_CUDA(cudaFuncSetAttribute(nce_flash_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        nce_flash_kernel<false><<<grid, kFlWorkers + 32, smem, stream>>>(X, Y, V, v_scale, d_n, n_max, np, inv_tau, nullptr, part_out);
    }
    IDG_LAUNCH_CHECK("nce_flash_kernel");
    return 0;
}

}  // namespace idg
