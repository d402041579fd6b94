// NGCF dense layer, forward, on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   S = [side | E (*) side] . [W_gcn ; W_bi] + b_gcn + b_bi            (models/NGCF.py:87-97; [N,128] x [128,64])
//   D = LeakyReLU_0.2(S) (*) keep / (1-p);   O = D / max(|D|_2, 1e-12)   (NGCF.py:98-106)
//
// The product is only 64 columns wide: 2.36 GFLOP against 185 MB of [N,64] streams per layer at the amazon-book shape, i.e.
// stream-bound once it leaves the fp32 FMA pipe (the CUDA-core tiles ran at 16-18 TFLOP/s = 130 us per layer; the warp-level
// mma.sync tf32 path was measured slower than that).  tcgen05 has no fp32 input kind, so every product is the 3xTF32 split
// x*y ~= xh*yh + xh*yl + xl*yh accumulated in one TMEM tile (error ~2^-21, inside the 1e-5 parity bar).
//
// Round 2: persistent CTAs (one per SM) instead of one tile per CTA (122 us per layer, bound by the un-overlapped
// load -> build -> MMA -> epilogue chain and by row-per-thread epilogue stores):
//   * Wcat^T (hi, lo) is split and laid out once per CTA and stays in shared memory (B operand, 4 k-chunks of 32).
//   * side / E / keep arrive through a four-slot cp.async ring in 16 KB units ([128 rows x 32 columns] of one stream); a thread
//     copies exactly the 16-byte pieces it later consumes, so the ring needs no barrier and keeps three units (48 KB per SM) in
//     flight while earlier ones are processed.
//   * a pair of units (side, E of one 32-column half) becomes two A chunks (k = the side half, k = 64 + the E*side half) in a
//     two-stage ring of 128B-swizzled K-major tiles; one thread issues the MMAs; two TMEM accumulators, so the epilogue of tile
//     t-1 runs after the operands of tile t have been handed over.
//   * epilogue: TMEM (thread <-> row) + bias -> shared-memory tile -> 8 lanes per row: activation, dropout, row norm by shuffles,
//     S / D / O written as full 128-byte lines.
#include <math.h>

#include "tc_common.cuh"

namespace idg {

constexpr int kFwBuilders = 512;               // warps 1..16 (8 warps left the SM issue-bound: `wait` stalls, two warps per scheduler)
constexpr int kFwPasses = 1024 / kFwBuilders;  // a [128 rows x 8 chunks] unit in passes of kFwBuilders / 8 rows
constexpr int kFwPassRows = kFwBuilders / 8;
constexpr uint32_t kFwBlkA = 128 * 128;        // bytes: [128 rows x 32 fp32] A chunk (hi or lo)
constexpr uint32_t kFwStage = 2 * kFwBlkA;     // hi | lo
constexpr uint32_t kFwBlkB = 64 * 128;         // bytes: [64 rows x 32 fp32] B chunk
constexpr uint32_t kFwWHalf = 4 * kFwBlkB;     // hi (or lo) of Wcat^T [64 x 128]
constexpr uint32_t kFwUnit = 128 * 128;        // one staging slot: [128 rows x 32 fp32] of one stream
constexpr int kFwSlots = 4, kFwAhead = 3, kFwUnits = 6;   // per iteration: side c0, E c0, side c1, E c1 (tile t), keep c0, keep c1 (tile t-1)
constexpr uint32_t kFwTile = 128 * 256;        // epilogue tile: S + bias, [128 rows x 64 fp32]
constexpr uint32_t kFwSmem = 2 * kFwWHalf + 2 * kFwStage + kFwSlots * kFwUnit + kFwTile;   // 64 + 64 + 64 + 32 KB
constexpr uint32_t kFwTmemCols = 128;          // two accumulators of 64 columns

__device__ __forceinline__ uint32_t fw_sw(int r, int cc) {   // 128 B-row block, 128-byte swizzle: chunk cc ^ (r & 7)
    return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((cc ^ (r & 7)) << 4);
}
__device__ __forceinline__ void fw_st4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 fw_ld4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void fw_split_store(uint32_t hi_addr, uint32_t lo_addr, float4 x) {
    float h[4], l[4];     // hi rounded to NEAREST tf32: |lo| <= 2^-12 |x|, half of what truncation leaves
    const float v[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t r;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v[j]));
        h[j] = __uint_as_float(r);
        l[j] = v[j] - h[j];
    }
    fw_st4(hi_addr, h[0], h[1], h[2], h[3]);
    fw_st4(lo_addr, l[0], l[1], l[2], l[3]);
}

// KEEP: 0 = no dropout, 1 = float mask rows (keep, through the ring), 2 = 64 mask bits per row (kbits, read in the epilogue)
template <int KEEP>
__global__ void __launch_bounds__(kFwBuilders + 32, 1) ngcf_dense_fwd_tc_kernel(const float* __restrict__ E, const float* __restrict__ side,
                                                                   const float* __restrict__ Wg, const float* __restrict__ bg,
                                                                   const float* __restrict__ Wb, const float* __restrict__ bb,
                                                                   const float* __restrict__ keep, const uint2* __restrict__ kbits, float inv_keep,
                                                                   int N, float* __restrict__ S_pre, float* __restrict__ D,
                                                                   float* __restrict__ out, int out_stride) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sW = smem;                            // Wcat^T hi | lo
    unsigned char* sA = sW + 2 * kFwWHalf;               // A ring: stage = chunk hi | lo
    unsigned char* sG = sA + 2 * kFwStage;               // cp.async staging slots
    unsigned char* sT = sG + kFwSlots * kFwUnit;         // epilogue tile
    uint64_t* bars = reinterpret_cast<uint64_t*>(sT + kFwTile);
    uint64_t *w_full = bars, *full = bars + 1, *empty = bars + 3, *tfull = bars + 5, *tempty = bars + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    float* sBias = reinterpret_cast<float*>(bars + 10);  // [64] b_gcn + b_bi
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ntiles = (N + 127) / 128;
    const int T = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // tiles of this CTA

    if (tid == 0) {
        mbar_init(w_full, kFwBuilders);
        for (int s = 0; s < 2; ++s) { mbar_init(full + s, kFwBuilders); mbar_init(empty + s, 1); mbar_init(tfull + s, 1); mbar_init(tempty + s, kFwBuilders); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 64) sBias[tid] = __ldg(bg + tid) + __ldg(bb + tid);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kFwTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t aW = smem_u32(sW), aA = smem_u32(sA), aG = smem_u32(sG), aT = smem_u32(sT);

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(128, 64);
            mbar_wait(w_full, 0);
            tc_fence_after();
            for (int it = 0; it < T; ++it) {
                const int buf = it & 1;
                mbar_wait(tempty + buf, ((it >> 1) & 1) ^ 1);      // the epilogue of tile it-2 has read this accumulator
                tc_fence_after();
#pragma unroll
                for (int p = 0; p < 2; ++p) {
#pragma unroll
                    for (int st = 0; st < 2; ++st) {
                        mbar_wait(full + st, (it * 2 + p) & 1);
                        tc_fence_after();
                        const int kc = st == 0 ? p : 2 + p;        // k-chunk: side half p, E*side half p
                        const uint32_t ah = aA + (uint32_t)st * kFwStage, bh = aW + (uint32_t)kc * kFwBlkB;
#pragma unroll
                        for (int sp = 0; sp < 3; ++sp) {
                            const uint32_t a = ah + (sp == 2 ? kFwBlkA : 0u), b = bh + (sp == 1 ? kFwWHalf : 0u);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                umma_tf32(tmem_base + (uint32_t)buf * 64u, umma_desc(a + ks * 32u), umma_desc(b + ks * 32u), idesc, (p | st | sp | ks) != 0);
                        }
                        umma_commit(empty + st);
                    }
                }
                umma_commit(tfull + buf);
            }
        }
    } else {
        const int bt = tid - 32, urow = bt >> 3, uj = bt & 7;      // unit mapping: row urow + kFwPassRows*pass, 16-byte chunk uj of the 128 B half-row
        const int q = warp & 3, cq = (warp - 1) >> 2;               // TMEM role: lane quadrant, column quarter (16 columns)
        const uint32_t tq = tmem_base + (((uint32_t)(q * 32)) << 16);
        // Wcat^T (hi, lo), once: B chunk kc, row n holds Wcat[32 kc .. 32 kc + 31][n]
#pragma unroll
        for (int i = 0; i < 2048 / kFwBuilders; ++i) {
            const int idx = bt + i * kFwBuilders;       // 64 columns n x 32 k-quads
            const int n = idx & 63, kq = idx >> 6;
            float wv[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int k = kq * 4 + jj;
                wv[jj] = __ldg((k < 64 ? Wg + (size_t)k * 64 : Wb + (size_t)(k - 64) * 64) + n);
            }
            const uint32_t o = (uint32_t)(kq >> 3) * kFwBlkB + fw_sw(n, kq & 7);
            fw_split_store(aW + o, aW + kFwWHalf + o, make_float4(wv[0], wv[1], wv[2], wv[3]));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(w_full);

        // flat unit index U = 6 * iteration + w; slot U & 3; one commit group per unit, even when empty
        auto issue = [&](int U) {
            const int it_u = U / kFwUnits, w = U - it_u * kFwUnits;
            const float* src = nullptr;
            int tt = 0, col0 = 0;
            if (w < 4) { if (it_u < T) { src = (w & 1) ? E : side; tt = (int)blockIdx.x + it_u * (int)gridDim.x; col0 = (w >> 1) * 32; } }
            else if (KEEP == 1 && it_u >= 1 && it_u <= T) { src = keep; tt = (int)blockIdx.x + (it_u - 1) * (int)gridDim.x; col0 = (w - 4) * 32; }
            if (src) {
                const uint32_t dst = aG + (uint32_t)(U & 3) * kFwUnit + (uint32_t)bt * 16u;
#pragma unroll
                for (int ps = 0; ps < kFwPasses; ++ps) {
                    const int r = tt * 128 + ps * kFwPassRows + urow;
                    cp_async16(dst + (uint32_t)ps * (kFwBuilders * 16u), src + (size_t)(r < N ? r : 0) * 64 + col0 + uj * 4, r < N ? 16 : 0);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // unit U has landed (three younger units stay in flight); returns this thread's first piece in its slot
        auto get = [&](int U) -> uint32_t {
            issue(U + kFwAhead);
            asm volatile("cp.async.wait_group %0;" ::"n"(kFwAhead) : "memory");
            return aG + (uint32_t)(U & 3) * kFwUnit + (uint32_t)bt * 16u;
        };
#pragma unroll
        for (int w = 0; w < kFwAhead; ++w) issue(w);

        for (int it = 0; it <= T; ++it) {
            const int U0 = it * kFwUnits;
            // ---- operands of tile it: two pairs of units -> two A chunks each
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                float4 sd[kFwPasses], e[kFwPasses];
                const uint32_t s0 = get(U0 + 2 * p);
                if (it < T) {
#pragma unroll
                    for (int ps = 0; ps < kFwPasses; ++ps) sd[ps] = fw_ld4(s0 + (uint32_t)ps * (kFwBuilders * 16u));
                }
                const uint32_t s1 = get(U0 + 2 * p + 1);
                if (it < T) {
#pragma unroll
                    for (int ps = 0; ps < kFwPasses; ++ps) e[ps] = fw_ld4(s1 + (uint32_t)ps * (kFwBuilders * 16u));
                    const int n = it * 2 + p;
                    mbar_wait(empty + 0, (n & 1) ^ 1);
                    mbar_wait(empty + 1, (n & 1) ^ 1);
#pragma unroll
                    for (int ps = 0; ps < kFwPasses; ++ps) {
                        const uint32_t o = fw_sw(ps * kFwPassRows + urow, uj);
                        fw_split_store(aA + o, aA + kFwBlkA + o, sd[ps]);
                        fw_split_store(aA + kFwStage + o, aA + kFwStage + kFwBlkA + o,
                                       make_float4(e[ps].x * sd[ps].x, e[ps].y * sd[ps].y, e[ps].z * sd[ps].z, e[ps].w * sd[ps].w));
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(full + 0);
                    mbar_arrive(full + 1);
                }
            }
            // ---- epilogue of tile it-1
            float4 k0[kFwPasses], k1[kFwPasses];
            const uint32_t g0 = get(U0 + 4);
            if (KEEP == 1 && it >= 1) {
#pragma unroll
                for (int ps = 0; ps < kFwPasses; ++ps) k0[ps] = fw_ld4(g0 + (uint32_t)ps * (kFwBuilders * 16u));
            }
            const uint32_t g1 = get(U0 + 5);
            if (KEEP == 1 && it >= 1) {
#pragma unroll
                for (int ps = 0; ps < kFwPasses; ++ps) k1[ps] = fw_ld4(g1 + (uint32_t)ps * (kFwBuilders * 16u));
            }
            if (it >= 1) {
                const int pt = it - 1, buf = pt & 1;
                const int r0 = ((int)blockIdx.x + pt * (int)gridDim.x) * 128;
                uint2 kw[kFwPasses];     // bit-packed dropout draws of this thread's rows (64 bits per row): 8 B per row instead of 256
                if (KEEP == 2) {
#pragma unroll
                    for (int ps = 0; ps < kFwPasses; ++ps) {
                        const int r = r0 + ps * kFwPassRows + urow;
                        kw[ps] = r < N ? __ldg(kbits + r) : make_uint2(0u, 0u);
                    }
                }
                mbar_wait(tfull + buf, (pt >> 1) & 1);
                tc_fence_after();
                {   // thread <-> row: S + bias into the tile, 16-byte chunk index XOR row (conflict-free here and for the reads below)
                    const int row = q * 32 + lane;
                    uint32_t raw[16];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(raw[0]), "=r"(raw[1]), "=r"(raw[2]), "=r"(raw[3]), "=r"(raw[4]), "=r"(raw[5]), "=r"(raw[6]), "=r"(raw[7]), "=r"(raw[8]),
                          "=r"(raw[9]), "=r"(raw[10]), "=r"(raw[11]), "=r"(raw[12]), "=r"(raw[13]), "=r"(raw[14]), "=r"(raw[15])
                        : "r"(tq + (uint32_t)(buf * 64 + cq * 16))
                        : "memory");
                    tmem_ld_wait();
                    tc_fence_before();
                    mbar_arrive(tempty + buf);
                    const uint32_t rb = aT + (uint32_t)row * 256u + (uint32_t)(cq >> 1) * 128u;
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 b = *reinterpret_cast<const float4*>(sBias + cq * 16 + j4 * 4);
                        const int c8 = (cq & 1) * 4 + j4;      // 16-byte chunk inside the 128 B half-row
                        fw_st4(rb + (uint32_t)(((c8 ^ row) & 7) << 4), __uint_as_float(raw[j4 * 4]) + b.x, __uint_as_float(raw[j4 * 4 + 1]) + b.y,
                               __uint_as_float(raw[j4 * 4 + 2]) + b.z, __uint_as_float(raw[j4 * 4 + 3]) + b.w);
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kFwBuilders) : "memory");
#pragma unroll
                for (int ps = 0; ps < kFwPasses; ++ps) {     // 8 lanes per row: columns 4 uj .. +3 and 32 + 4 uj .. +3
                    const int rr = ps * kFwPassRows + urow, r = r0 + rr;
                    const uint32_t o = aT + (uint32_t)rr * 256u + (uint32_t)(((uj ^ rr) & 7) << 4);
                    const float4 sa = fw_ld4(o), sb = fw_ld4(o + 128u);
                    const float s8[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
                    float kv[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
                    if (KEEP == 1) { kv[0] = k0[ps].x; kv[1] = k0[ps].y; kv[2] = k0[ps].z; kv[3] = k0[ps].w; kv[4] = k1[ps].x; kv[5] = k1[ps].y; kv[6] = k1[ps].z; kv[7] = k1[ps].w; }
                    if (KEEP == 2) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            kv[j] = (float)((kw[ps].x >> (uj * 4 + j)) & 1u);
                            kv[4 + j] = (float)((kw[ps].y >> (uj * 4 + j)) & 1u);
                        }
                    }
                    float d8[8], ss = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float act = s8[j] > 0.f ? s8[j] : 0.2f * s8[j];
                        d8[j] = KEEP ? act * kv[j] * inv_keep : act;
                        ss = fmaf(d8[j], d8[j], ss);
                    }
                    ss += __shfl_xor_sync(0xffffffffu, ss, 4); ss += __shfl_xor_sync(0xffffffffu, ss, 2); ss += __shfl_xor_sync(0xffffffffu, ss, 1);
                    const float nrm = fmaxf(sqrtf(ss), 1e-12f);
                    if (r < N) {
                        if (S_pre) { st4(S_pre + (size_t)r * 64 + uj * 4, sa); st4(S_pre + (size_t)r * 64 + 32 + uj * 4, sb); }
                        st4(D + (size_t)r * 64 + uj * 4, make_float4(d8[0], d8[1], d8[2], d8[3]));
                        st4(D + (size_t)r * 64 + 32 + uj * 4, make_float4(d8[4], d8[5], d8[6], d8[7]));
                        st4(out + (size_t)r * out_stride + uj * 4, make_float4(d8[0] / nrm, d8[1] / nrm, d8[2] / nrm, d8[3] / nrm));
                        st4(out + (size_t)r * out_stride + 32 + uj * 4, make_float4(d8[4] / nrm, d8[5] / nrm, d8[6] / nrm, d8[7] / nrm));
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kFwBuilders) : "memory");     // the tile is free for the next epilogue
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kFwTmemCols) : "memory");
    }
}

// launched by idg_ngcf_dense_fwd / _bits (csrc/ngcf.cu) for the 64-wide layers of the reference configuration; S_pre may be null;
// the dropout draws come as a float mask (keep), as 64 bits per row (keep_bits: word w, bit b = column 32 w + b) or not at all
int ngcf_dense_fwd_tc(const float* E, const float* side, const float* Wg, const float* bg, const float* Wb, const float* bb, const float* keep,
                      const uint32_t* keep_bits, float inv_keep, int N, float* S_pre, float* D, float* out, int out_stride, cudaStream_t stream) {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        IDG_CUDA(cudaGetDevice(&dev));
        IDG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int ntiles = (N + 127) / 128;
    const int grid = sms < ntiles ? sms : ntiles;
    const size_t smem = (size_t)kFwSmem + 512;
    auto launch = [&](auto kern) -> int {
        IDG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kFwBuilders + 32, smem, stream>>>(E, side, Wg, bg, Wb, bb, keep, reinterpret_cast<const uint2*>(keep_bits), inv_keep, N, S_pre, D, out, out_stride);
        return 0;
    };
    if (int rc = keep_bits ? launch(ngcf_dense_fwd_tc_kernel<2>) : (keep ? launch(ngcf_dense_fwd_tc_kernel<1>) : launch(ngcf_dense_fwd_tc_kernel<0>))) return rc;
    IDG_LAUNCH_CHECK("ngcf_dense_fwd_tc_kernel");
    return 0;
}

}  // namespace idg
