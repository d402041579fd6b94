// Full-ranking candidate pass on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same contract as eval_candidates_kernel (csrc/eval.cu): per user keep every unmasked item whose
// approximate score can still belong to the exact top-K; the survivors are rescored exactly in fp64
// by eval_rescore_kernel, so the returned ids do not depend on the tensor-core rounding.
//
//   scores:  S[128 users, 128 items] = Fu_tile . Fi_tile^T, d = 64, one tcgen05.mma.kind::tf32 chain
//            (8 instructions of K = 8) per item tile, accumulators in TMEM (4 buffers x 128 columns).
//            operands are rounded to nearest tf32 before the unit sees them: |s~ - s| <= m_ui = c * |u||i|, c ~ 1.01e-3 (tc_margin_coef).
//            The margin is PER ITEM and costs nothing in the epilogue: a ninth k-step multiplies an extra operand column
//            (c|u| per user row, |i| per item row, both rounded up to tf32) so the accumulator holds the UPPER bound
//            w_ui = s~_ui + m_ui directly.  An item is kept iff w_ui >= L_u, L_u = the K-th largest LOWER bound
//            w - 2m seen so far (computed when a row's list is pruned): that set provably contains the exact top-K.
//            (Round 1 used one margin per user from the largest item norm: after a few epochs of training the largest
//            norm is ~8x the median and the lists flooded -- 12 ms instead of 5.5 ms at the amazon-book shape.)
//   roles:   warp 0      MMA issuer (one elected thread) + TMEM alloc/dealloc
//            warps 2-3   operand loaders: the user tile (gathered rows) by cp.async into the 128B-swizzled K-major
//                        UMMA layout; item tiles by TMA (cp.async.bulk.tensor.2d, 128B-swizzle tensor map, one
//                        elected thread, mbarrier expect_tx), 3-stage ring
//            warps 4-7   epilogue: thread <-> user row (TMEM lane), tcgen05.ld 32 columns at a time,
//                        train-mask by a per-row cursor over the sorted positives, threshold filter
//                        in registers, per-row candidate lists in shared memory, warp-cooperative prune
//   bound:   TMEM drain (64 B/clk/SM = 16 scores/clk/SM), not the MMA rate.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <math.h>

#include "tc_common.cuh"

namespace idg {

constexpr int kTcM = 128, kTcN = 128;
constexpr int kTcStages = 2, kTcBufs = 2;   // per CTA; two CTAs share an SM (2 x 97 KB smem, 2 x 256 TMEM columns)
constexpr uint32_t kTcTmemCols = kTcBufs * 128;
// per-row candidate list (in L2): pruned when it exceeds kTcTrig entries.  The filter threshold only moves at a prune, so the trigger
// trades prune count against appends through a stale threshold: 128 / 160 was measured and is 2x SLOWER than 48 / 80 (18 vs 9.6 ms
// after three epochs of training, when ~35 entries survive each prune because of the margin band at the top of the ranking).
constexpr int kTcCap = 80, kTcTrig = 48, kTcCandOut = 64;
constexpr int kTcPruneQ = (kTcCap + 31) / 32;   // list entries per lane in the warp-cooperative prune (80 -> 3)
static_assert(kTcPruneQ * 32 >= kTcCap && kTcCap >= kTcTrig + 32, "the prune must see the whole list; a chunk appends up to 32 entries");
constexpr int kTcLoaders = 64;
constexpr uint32_t kSubTile = 128 * 128;  // bytes of one [128 rows x 128 B] swizzle-atom column
constexpr uint32_t kAugTile = 128 * 32;   // bytes of one [128 rows x 32 B] margin operand (32-byte swizzle atoms)
// Both operands reach the tensor core already ROUNDED TO NEAREST tf32 (cvt.rna: the item table as a rounded copy written by
// item_norm_kernel, the user tile rounded by its loader), so the unit reads them exactly: per element |delta| <= 2^-11, per product
// <= 2^-10 + 2^-22, per score <= (2^-10 + 2^-22) sum_k |u_k i_k| <= (2^-10 + 2^-22) |u||i|, plus the fp32 accumulation of D + 1 terms
// (<= (D + 1) * 2^-24 |u||i|: 3.9e-6 at D = 64, 1.54e-5 at D = 256).  3 % slack on top.  (Round 1 let the hardware truncate: 2^-9, and
// the candidate lists of a trained table carried ~35 entries through every prune.)
template <int D> __host__ __device__ constexpr float tc_margin_coef() { return 1.03f * (0.0009765625f + 2.4e-7f + (D == 64 ? 3.9e-6f : 1.54e-5f)); }
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

struct EvalWsTc {
    float* max_norm;
    int* flag_cnt;
    int* flag_list;
    int* cand_cnt;
    int* cand_ids;
    float* list_s;  // [ceil(nu/128)*128, kTcCap] per-row candidate scores in L2 (rare appends): frees shared memory for 2 CTAs/SM
    int* list_i;
    const float* aug;  // [I, 8]: column 0 = |i| rounded up to tf32, columns 1..7 zero (the item half of the margin k-step)
    float* dbg_tile;   // self-check (idg_eval_tc_bounds): CTA 0 dumps the raw accumulators of item tile 0, [128,128]; else NULL
};

// x >= 0 rounded UP to a value the tensor core reads exactly (tf32: 10 explicit mantissa bits)
__device__ __forceinline__ float tf32_ceil(float x) { return __uint_as_float((__float_as_uint(x) + 0x1fffu) & 0xffffe000u); }

// c|u| of one user row, rounded up to tf32; the loader (operand) and the epilogue (lower bounds) must agree bit for bit
template <int D>
__device__ __forceinline__ float tc_user_coef(const float* __restrict__ urow) {
    float ss = 0.f;
    for (int k = 0; k < D; ++k) { const float v = __ldg(urow + k); ss = fmaf(v, v, ss); }
    return tf32_ceil(tc_margin_coef<D>() * sqrtf(ss) + 1e-30f);
}

// smem byte offset of the 16-byte chunk kc (0..15) of row r inside an operand tile [128 rows x 64 fp32]
__device__ __forceinline__ uint32_t sw128_offset(int r, int kc) {
    return (uint32_t)(kc >> 3) * kSubTile + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)(((kc & 7) ^ (r & 7)) << 4);
}

// Out-of-line append of the survivors of one 4-column group (columns j0, j0+8, j0+16, j0+24 of the chunk).  Kept
// out of the epilogue's hot loop on purpose: ncu showed the loop stalled on instruction fetch (no_instruction) when
// the 32 append sites were inlined.
__device__ __noinline__ int tc_group_append(float v0, float v1, float v2, float v3, float tau, int id0, int I, float* ls, int* li, int cnt) {
    if (v0 >= tau && id0 < I) { ls[cnt] = v0; li[cnt] = id0; ++cnt; }
    if (v1 >= tau && id0 + 8 < I) { ls[cnt] = v1; li[cnt] = id0 + 8; ++cnt; }
    if (v2 >= tau && id0 + 16 < I) { ls[cnt] = v2; li[cnt] = id0 + 16; ++cnt; }
    if (v3 >= tau && id0 + 24 < I) { ls[cnt] = v3; li[cnt] = id0 + 24; ++cnt; }
    return cnt;
}

// warp-cooperative prune of one row's candidate list.  Entries hold UPPER bounds w; lower_j = w_j - c2 * |i_j| (c2 = 2 c|u|).
// L = K-th largest lower bound; keep w >= L.
__device__ __forceinline__ void tc_prune(float* ls, int* li, int m, float c2, const float* __restrict__ aug, int K, int lane, int& new_cnt,
                                         float& new_tau) {
    float s[kTcPruneQ], lo[kTcPruneQ]; int id[kTcPruneQ]; int rank[kTcPruneQ];
#pragma unroll
    for (int q = 0; q < kTcPruneQ; ++q) {
        const int idx = lane + 32 * q;
        s[q] = (idx < m) ? __ldcg(ls + idx) : -INFINITY; id[q] = (idx < m) ? __ldcg(li + idx) : 0; rank[q] = 0;
        lo[q] = (idx < m) ? fmaf(-c2, __ldg(aug + (size_t)id[q] * 8), s[q]) : -INFINITY;
    }
    // rank counting over the lower bounds held in registers (kTcPruneQ entries per lane), broadcast by shuffle
#pragma unroll
    for (int qq = 0; qq < kTcPruneQ; ++qq) {
        for (int l = 0; l < 32; ++l) {
            const int j = l + 32 * qq;
            if (j >= m) break;
            const float sj = __shfl_sync(0xffffffffu, lo[qq], l);
#pragma unroll
            for (int q = 0; q < kTcPruneQ; ++q) rank[q] += (sj > lo[q]) || (sj == lo[q] && j < lane + 32 * q);
        }
    }
    float vk = -INFINITY;
#pragma unroll
    for (int q = 0; q < kTcPruneQ; ++q) if (lane + 32 * q < m && rank[q] == K - 1) vk = lo[q];
#pragma unroll
    for (int mm = 16; mm >= 1; mm >>= 1) vk = fmaxf(vk, __shfl_xor_sync(0xffffffffu, vk, mm));
    const float tau = vk;
    __syncwarp();
    int kept = 0;
#pragma unroll
    for (int q = 0; q < kTcPruneQ; ++q) {
        const bool keep = (lane + 32 * q < m) && (s[q] >= tau);
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (keep) { const int p = kept + __popc(b & ((1u << lane) - 1)); ls[p] = s[q]; li[p] = id[q]; }
        kept += __popc(b);
    }
    __syncwarp();
    new_cnt = kept; new_tau = tau;
}

// D = 64: the whole K extent is one 64-wide chunk per item tile, two CTAs per SM.  D = 256 (NGCF's concatenated layers,
// models/NGCF.py:108,132-138): the user tile (128 KB) stays resident and each item tile streams through the stage ring as four
// 64-wide k-chunks accumulating in the same TMEM buffer; one CTA per SM.  The margin operand rides with the LAST chunk of a tile.
template <int D>
__global__ void __launch_bounds__(256, (D == 64) ? 2 : 1) eval_candidates_tc_kernel(const float* __restrict__ Fu, const float* __restrict__ Fi, int I,
                                                                    const int32_t* __restrict__ mptr, const int32_t* __restrict__ mind,
                                                                    const int64_t* __restrict__ users, int nu, int K, EvalWsTc w,
                                                                    const __grid_constant__ CUtensorMap tmap_items,
                                                                    const __grid_constant__ CUtensorMap tmap_aug) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int NCH = D / 64;                                      // k-chunks of 64 floats per item tile
    unsigned char* sA = smem;                                        // D / 32 sub-tiles of 16 KB (32 KB at D = 64, 128 KB at D = 256)
    unsigned char* sB = smem + (D / 32) * kSubTile;                  // kTcStages x 32 KB: one k-chunk of one item tile per stage
    unsigned char* sAaug = sB + kTcStages * 2 * kSubTile;            // 4 KB: c|u| per user row (32-byte swizzle layout)
    unsigned char* sBaug = sAaug + kAugTile;                         // kTcStages x 4 KB: |i| per item row, by TMA
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBaug + kTcStages * kAugTile);
    uint64_t* full = bars;                   // [kTcStages] loaders -> MMA
    uint64_t* empty = bars + kTcStages;      // [kTcStages] MMA -> loaders
    uint64_t* tfull = empty + kTcStages;     // [kTcBufs]   MMA -> epilogue
    uint64_t* tempty = tfull + kTcBufs;      // [kTcBufs]   epilogue -> MMA
    uint64_t* afull = tempty + kTcBufs;      // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(afull + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int u0 = blockIdx.x * kTcM;
    float* ls = w.list_s + (size_t)u0 * kTcCap;  // this CTA's 128 rows
    int* li = w.list_i + (size_t)u0 * kTcCap;
    const int ntiles = (I + kTcN - 1) / kTcN;

    if (tid == 0) {
        for (int s = 0; s < kTcStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }  // full: one arrive.expect_tx + TMA bytes
        for (int b = 0; b < kTcBufs; ++b) { mbar_init(tfull + b, 1); mbar_init(tempty + b, 128); }
        mbar_init(afull, kTcLoaders);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTcTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(kTcM, kTcN);
            mbar_wait(afull, 0);
            for (int t = 0; t < ntiles; ++t) {
                const int b = t % kTcBufs;
                mbar_wait(tempty + b, ((t / kTcBufs) & 1) ^ 1);
#pragma unroll 1
                for (int ch = 0; ch < NCH; ++ch) {
                    const int it = t * NCH + ch, s = it % kTcStages;
                    mbar_wait(full + s, (it / kTcStages) & 1);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA) + (uint32_t)(2 * ch) * kSubTile, b0 = smem_u32(sB + (size_t)s * 2 * kSubTile);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t koff = (uint32_t)(k >> 2) * kSubTile + (uint32_t)(k & 3) * 32u;
                        umma_tf32(tmem_base + (uint32_t)b * kTcN, umma_desc(a0 + koff), umma_desc(b0 + koff), idesc, (ch | k) != 0);
                    }
                    // margin k-step after the last chunk: + c|u| * |i|  (operands in 32-byte-swizzle K-major atoms)
                    if (ch == NCH - 1)
                        umma_tf32(tmem_base + (uint32_t)b * kTcN, umma_desc_sw32(smem_u32(sAaug)), umma_desc_sw32(smem_u32(sBaug + (size_t)s * kAugTile)), idesc, 1);
                    umma_commit(empty + s);
                }
                umma_commit(tfull + b);
            }
        }
    } else if (warp == 2 || warp == 3) {
        // ===================== operand loaders =====================
        const int lt = tid - 64;  // 0..63
        // A: the CTA's 128 user rows (gathered through users[]), once
        for (int c = lt; c < kTcM * (D / 4); c += kTcLoaders) {
            const int r = c / (D / 4), kc = c % (D / 4);
            const int p = u0 + r;
            const int64_t u = (p < nu) ? users[p] : users[0];
            float4 v = (p < nu) ? __ldg(reinterpret_cast<const float4*>(Fu + (size_t)u * D + kc * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            v = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
            *reinterpret_cast<float4*>(sA + sw128_offset(r, kc)) = v;
        }
        // margin operand of the user tile: row r = [c|u_r|, 0, 0, 0 | c|u_r|, 0, 0, 0].  Both 16-byte chunks carry the value so the
        // product with the item row [|i|, 0, ... 0] is c|u||i| whichever way the 32-byte swizzle orders the two chunks of a row.
        for (int r = lt; r < kTcM; r += kTcLoaders) {
            const int p = u0 + r;
            const float cu = (p < nu) ? tc_user_coef<D>(Fu + (size_t)users[p] * D) : 0.f;
            float4* dst = reinterpret_cast<float4*>(sAaug + (size_t)(r >> 3) * 256 + (size_t)(r & 7) * 32);
            dst[0] = make_float4(cu, 0.f, 0.f, 0.f);
            dst[1] = make_float4(cu, 0.f, 0.f, 0.f);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(afull);
        // B: item tiles by TMA -- the tensor map (box 32 floats x 128 rows, 128B swizzle) lands each [128 x 32] column block
        // directly in the UMMA K-major layout; rows past I are zero-filled by the hardware.  One elected thread.
        if (lt == 0) {
            tma_prefetch_desc(&tmap_items);
            tma_prefetch_desc(&tmap_aug);
            for (int it = 0; it < ntiles * NCH; ++it) {
                const int t = it / NCH, ch = it % NCH, s = it % kTcStages;
                mbar_wait(empty + s, ((it / kTcStages) & 1) ^ 1);
                const uint32_t dst = smem_u32(sB + (size_t)s * 2 * kSubTile);
                const bool last = ch == NCH - 1;   // the margin operand of the tile travels with its last k-chunk (same stage, same barrier)
                mbar_arrive_expect_tx(full + s, 2 * kSubTile + (last ? kAugTile : 0u));
                tma_load_2d(dst, &tmap_items, ch * 64, t * kTcN, full + s);
                tma_load_2d(dst + kSubTile, &tmap_items, ch * 64 + 32, t * kTcN, full + s);
                if (last) tma_load_2d(smem_u32(sBaug + (size_t)s * kAugTile), &tmap_aug, 0, t * kTcN, full + s);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: one thread per user row =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int p = u0 + row;
        const bool valid = p < nu;
        const int u = valid ? (int)users[p] : 0;
        // tf32 operand truncation: |err| <= (2*2^-10 + 2^-20)|u||i| plus fp32 accumulation (tc_margin_coef).
        // The accumulator already holds s~ + c|u||i|; lower bound of an entry = w - 2 c|u||i|.
        const float delta2 = valid ? 2.f * tc_user_coef<D>(Fu + (size_t)u * D) : 0.f;
        float tau = valid ? -3.0e38f : INFINITY;  // finite: masked scores (-inf) never pass
        int cnt = 0, overflow = 0;
        int cur = valid ? mptr[u] : 0;
        const int cend = valid ? mptr[u + 1] : 0;
        int next_mask = (cur < cend) ? __ldg(mind + cur) : 0x7fffffff;
        float* my_ls = ls + row * kTcCap;
        int* my_li = li + row * kTcCap;
        const uint32_t lane_base = ((uint32_t)(q * 32)) << 16;

        int next2 = (cur + 1 < cend) ? __ldg(mind + cur + 1) : 0x7fffffff;  // one positive of look-ahead: no exposed load latency
        for (int t = 0; t < ntiles; ++t) {
            const int b = t % kTcBufs;
            mbar_wait(tfull + b, (t / kTcBufs) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < kTcN / 32; ++c) {  // kept rolled: the epilogue is instruction-cache sensitive (ncu: no_instruction stalls)
                uint32_t raw[32];
                tmem_ld32(tmem_base + lane_base + (uint32_t)(b * kTcN + c * 32), raw);
                tmem_ld_wait();
                if (c == kTcN / 32 - 1) {  // accumulator fully read: hand the buffer back to the MMA warp
                    tc_fence_before();
                    mbar_arrive(tempty + b);
                }
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
                if (w.dbg_tile && t == 0 && blockIdx.x == 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) w.dbg_tile[(size_t)row * kTcN + c * 32 + j] = v[j];
                }
                const int c0 = t * kTcN + c * 32;
                // train positives of this row that fall into [c0, c0+32): remove (batch_test.py:62-65)
                while (next_mask < c0 + 32) {
                    const int j = next_mask - c0;
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) if (jj == j) v[jj] = -INFINITY;
                    ++cur;
                    next_mask = next2;
                    next2 = (cur + 1 < cend) ? __ldg(mind + cur + 1) : 0x7fffffff;
                }
                float m8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) m8[j] = fmaxf(fmaxf(v[j], v[j + 8]), fmaxf(v[j + 16], v[j + 24]));
                const float m = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
                if (m >= tau) {
                    // usually one or two survivors per warp-chunk: descend only into the 4-column groups whose maximum passes
#pragma unroll
                    for (int g8 = 0; g8 < 8; ++g8)
                        if (m8[g8] >= tau) cnt = tc_group_append(v[g8], v[g8 + 8], v[g8 + 16], v[g8 + 24], tau, c0 + g8, I, my_ls, my_li, cnt);
                }
                unsigned need = __ballot_sync(0xffffffffu, cnt > kTcTrig);
                while (need) {
                    const int L = __ffs(need) - 1;
                    need &= need - 1;
                    const int mL = __shfl_sync(0xffffffffu, cnt, L);
                    const float dL = __shfl_sync(0xffffffffu, delta2, L);
                    int nc; float nt;
                    __syncwarp();
                    tc_prune(ls + (q * 32 + L) * kTcCap, li + (q * 32 + L) * kTcCap, mL, dL, w.aug, K, lane, nc, nt);
                    if (lane == L) {
                        cnt = nc;
                        if (nc > kTcTrig) { overflow = 1; tau = INFINITY; } else tau = nt;
                    }
                }
            }
        }
        // final prune of every row, then publish the candidate ids
        for (int L = 0; L < 32; ++L) {
            const int mL = __shfl_sync(0xffffffffu, cnt, L);
            const float dL = __shfl_sync(0xffffffffu, delta2, L);
            if (mL >= K) {
                int nc; float nt;
                __syncwarp();
                tc_prune(ls + (q * 32 + L) * kTcCap, li + (q * 32 + L) * kTcCap, mL, dL, w.aug, K, lane, nc, nt);
                if (lane == L) cnt = nc;
            }
        }
        __syncwarp();
        if (valid) {
            const bool bad = overflow || cnt > kTcCandOut || cnt < K;
            if (bad) {
                w.cand_cnt[p] = 0;
                w.flag_list[atomicAdd(w.flag_cnt, 1)] = p;
            } else {
                w.cand_cnt[p] = cnt;
                for (int j = 0; j < cnt; ++j) w.cand_ids[(size_t)p * kTcCandOut + j] = __ldcg(my_li + j);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTcTmemCols) : "memory");
    }
}

template <int D>
static int launch_eval_candidates_tc_d(const float* Fu, const float* Fi, int I, const int32_t* mptr, const int32_t* mind, const int64_t* users,
                                       int nu, int K, const EvalWsTc& w, cudaStream_t stream) {
    // tensor map of the (rounded) item table [I rows x D fp32]: driver entry point fetched through the runtime (no libcuda link)
    static PFN_cuTensorMapEncodeTiled encode = nullptr;
    if (!encode) {
        cudaDriverEntryPointQueryResult qres;
        void* fn = nullptr;
        IDG_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return fail(-1, "cuTensorMapEncodeTiled is not available%s");
        encode = (PFN_cuTensorMapEncodeTiled)fn;
    }
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)I};
    const cuuint64_t gstride[1] = {(cuuint64_t)D * sizeof(float)};
    const cuuint32_t box[2] = {32, (cuuint32_t)kTcN};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(Fi), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(-1, "cuTensorMapEncodeTiled failed (%s%lld)", "", (long long)cr);
    // margin operand [I rows x 8 fp32] (idg::item_norm_kernel): box 8 floats x 128 rows, 32-byte swizzle = the UMMA SWIZZLE_32B K-major atom
    CUtensorMap tmap_aug;
    const cuuint64_t gdim2[2] = {8, (cuuint64_t)I};
    const cuuint64_t gstride2[1] = {8 * sizeof(float)};
    const cuuint32_t box2[2] = {8, (cuuint32_t)kTcN};
    const CUresult cr2 = encode(&tmap_aug, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(w.aug), gdim2, gstride2, box2, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr2 != CUDA_SUCCESS) return fail(-1, "cuTensorMapEncodeTiled (margin operand) failed (%s%lld)", "", (long long)cr2);
    const size_t smem = (size_t)(D / 32) * kSubTile + (size_t)kTcStages * 2 * kSubTile + (1 + kTcStages) * (size_t)kAugTile +
                        sizeof(uint64_t) * (2 * kTcStages + 2 * kTcBufs + 1) + 16;
    IDG_CUDA(cudaFuncSetAttribute(eval_candidates_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    eval_candidates_tc_kernel<D><<<(unsigned)((nu + kTcM - 1) / kTcM), 256, smem, stream>>>(Fu, Fi, I, mptr, mind, users, nu, K, w, tmap, tmap_aug);
    IDG_LAUNCH_CHECK("eval_candidates_tc_kernel");
    return 0;
}

int launch_eval_candidates_tc(const float* Fu, const float* Fi, int I, int d, const int32_t* mptr, const int32_t* mind, const int64_t* users,
                              int nu, int K, float* max_norm, int* flag_cnt, int* flag_list, int* cand_cnt, int* cand_ids,
                              float* list_s, int* list_i, const float* aug, float* dbg_tile, cudaStream_t stream) {
    EvalWsTc w{max_norm, flag_cnt, flag_list, cand_cnt, cand_ids, list_s, list_i, aug, dbg_tile};
    if (d == 64) return launch_eval_candidates_tc_d<64>(Fu, Fi, I, mptr, mind, users, nu, K, w, stream);
    if (d == 256) return launch_eval_candidates_tc_d<256>(Fu, Fi, I, mptr, mind, users, nu, K, w, stream);
    return fail(-1, "eval_candidates_tc: d must be 64 or 256 (%s%lld)", "", d);
}

}  // namespace idg
