// NGCF dense layer, backward, on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Autograd of models/NGCF.py:87-106 for one layer (same algebra as ngcf_dense_bwd_kernel in csrc/ngcf.cu, which stays as the
// CUDA-core cross-check, IDG_NGCF_BWD=fma):
//   dD = dD_ext + (dO - O <O,dO>)/|D| ;  dS = dD (*) keep/(1-p) (*) (S > 0 ? 1 : 0.2)                 [N,64]   CUDA cores, per row
//   dZ = dS . Wcat^T                      [N,64] x [64,128]   (Wcat = [W_gcn ; W_bi], Z = [side | E (*) side])   tensor cores
//   dWcat = Z^T . dS                      [128,N] x [N,64]                                                        tensor cores
//   dside = dZ1 + dZ2 (*) E ;  dE_direct = dZ2 (*) side ;  db = colsum(dS)                                        CUDA cores
// Every product is the 3xTF32 split (xh yh + xh yl + xl yh, fp32 accumulation in TMEM; error ~2^-21).
//
// Persistent CTAs (one per SM) walk 128-row tiles.  Shared memory holds Wcat (hi, lo) for the whole kernel, the tile's dS (hi, lo)
// in TWO images and a two-stage ring of 16-row Z chunks (hi, lo).  Both products read row-major [rows x 32 fp32] blocks (128 B per
// row) as the builder threads hold them -- no transposes anywhere: for dZ = dS . Wcat^T the rows are the M index and the block is a
// K-major operand (128-byte swizzle); for dWcat = Z^T . dS the rows are the CONTRACTION index and the blocks are MN-major operands
// (8 rows per k-step, the 32-column blocks one LBO apart).  tcgen05 takes 32-bit MN-major operands only in the "128-byte swizzle
// with 32-byte atomicity" layout (cute: Layout_MN_SW128_32B_Atom, descriptor layout type 1), hence the second image of dS.
// dWcat accumulates in one TMEM tile across ALL tiles of the CTA and leaves as a per-CTA partial (summed in CTA order by
// ngcf_reduce_kernel: deterministic).
//
//   warp 0      : TMEM allocation, one thread issues every tcgen05.mma
//   warps 1..8  : build Wcat once; per tile build dS (16 lanes per row: the row norm and <D,dO> by shuffles), stream the eight Z
//                 chunks through the ring, then read dZ from TMEM (thread <-> row, two warps per lane quadrant = two column halves),
//                 stage it through shared memory and write dside / dE_direct as full rows
#include <math.h>

#include "tc_common.cuh"

namespace idg {

constexpr int kBwBuilders = 256;              // warps 1..8
constexpr uint32_t kBwBlk = 128 * 128;        // bytes of one [128 rows x 32 fp32] block
constexpr uint32_t kBwHalf = 2 * kBwBlk;      // hi (or lo) of Wcat [128 x 64] or of dS [128 x 64]: two blocks
constexpr int kBwZRows = 16;                  // rows per Z ring stage (2 k-steps of the dW product)
constexpr int kBwZChunks = 128 / kBwZRows;    // stages per tile
constexpr uint32_t kBwZBlk = kBwZRows * 128;  // [16 rows x 32 fp32]
constexpr uint32_t kBwZHalf = 4 * kBwZBlk;    // hi (or lo) of one stage: the four 32-feature blocks of Z
constexpr uint32_t kBwZStage = 2 * kBwZHalf;
constexpr int kBwZStages = 2;
constexpr uint32_t kBwTmemCols = 256;         // dZ: columns [0,128), dWcat of the current tile: [128,192), running dWcat: [192,256)
constexpr uint32_t kBwSmem = 2 * kBwHalf + 4 * kBwHalf + kBwZStages * kBwZStage;   // Wcat | dS K-major | dS MN-major | Z ring = 224 KB

// MN-major fp32 operand, 128-byte swizzle with 32-byte atomicity: 32 fp32 of the M/N index are contiguous (one 128 B row), 4 rows
// of the K index form one 512 B atom in which the 32-byte chunk index is XORed with the row index; LBO = distance between 32-wide
// M/N blocks, SBO = distance between 4-row k groups (cute/atom/mma_traits_sm100.hpp: Swizzle<2,5,2> o ((T,8,m),(4,k)):((1,T,LBO),(8T,SBO))).
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]),
        "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]),
        "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]),
        "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t bw_sw32(int r, int cc) {   // same block shape, 32-byte-atom swizzle: chunk pair (cc>>1) ^ (r & 3)
    return (uint32_t)r * 128u + (uint32_t)((((cc >> 1) ^ (r & 3)) << 5) | ((cc & 1) << 4));
}
__device__ __forceinline__ uint32_t bw_sw(int r, int cc) {   // byte offset of 16-byte chunk cc of row r inside a 128 B-row block
    return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((cc ^ (r & 7)) << 4);
}
__device__ __forceinline__ void bw_st4(uint32_t addr, const float* v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}
// hi = x rounded to NEAREST tf32 (|lo| <= 2^-12 |x|, half of what truncation leaves), lo = x - hi exactly
__device__ __forceinline__ void bw_split(float x, float& hi, float& lo) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    hi = __uint_as_float(r);
    lo = x - hi;
}
__device__ __forceinline__ void bw_split_store(uint32_t hi_addr, uint32_t lo_addr, const float* x) {
    float h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bw_split(x[j], h[j], l[j]);
    bw_st4(hi_addr, h);
    bw_st4(lo_addr, l);
}

__global__ void __launch_bounds__(288, 1) ngcf_dense_bwd_tc_kernel(const float* __restrict__ E, const float* __restrict__ side,
                                                                   const float* __restrict__ Wg, const float* __restrict__ Wb,
                                                                   const float* __restrict__ keep, float inv_keep, const float* __restrict__ S_pre,
                                                                   const float* __restrict__ D, const float* __restrict__ dO, int dO_stride,
                                                                   const float* __restrict__ dD_ext, int N, float* __restrict__ dside,
                                                                   float* __restrict__ dE_direct, float* __restrict__ dW_part,
                                                                   float* __restrict__ db_part) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sW = smem;                          // Wcat hi | lo
    unsigned char* sS = sW + 2 * kBwHalf;              // dS hi | lo, K-major image (dZ product)
    unsigned char* sM = sS + 2 * kBwHalf;              // dS hi | lo, MN-major image (dWcat product)
    unsigned char* sZ = sM + 2 * kBwHalf;              // ring: stage s = Z chunk hi | lo
    uint64_t* bars = reinterpret_cast<uint64_t*>(sZ + kBwZStages * kBwZStage);
    uint64_t *w_full = bars, *ds_full = bars + 1, *ds_empty = bars + 2, *d1_full = bars + 3, *d2_full = bars + 4;
    uint64_t *z_full = bars + 5, *z_empty = z_full + kBwZStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(z_empty + kBwZStages);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ntiles = (N + 127) / 128;

    if (tid == 0) {
        mbar_init(w_full, kBwBuilders); mbar_init(ds_full, kBwBuilders); mbar_init(ds_empty, 1); mbar_init(d1_full, 1); mbar_init(d2_full, 1);
        for (int s = 0; s < kBwZStages; ++s) { mbar_init(z_full + s, kBwBuilders); mbar_init(z_empty + s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kBwTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t aW = smem_u32(sW), aS = smem_u32(sS), aM = smem_u32(sM), aZ = smem_u32(sZ);
    float4 dbv = f4zero();   // builders: column sums of dS over this thread's rows (columns tx*4 .. tx*4+3)

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t idesc1 = umma_idesc_tf32(128, 128);                             // dZ = dS . Wcat^T: both operands K-major
            const uint32_t idesc2 = umma_idesc_tf32(128, 64) | (1u << 15) | (1u << 16);    // dWcat = Z^T . dS: both operands MN-major
            mbar_wait(w_full, 0);
            tc_fence_after();
            int it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                mbar_wait(ds_full, it & 1);
                tc_fence_after();
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const uint32_t a = aS + (p == 2 ? kBwHalf : 0u), b = aW + (p == 1 ? kBwHalf : 0u);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint32_t off = (uint32_t)(ks >> 2) * kBwBlk + (uint32_t)(ks & 3) * 32u;
                        umma_tf32(tmem_base, umma_desc(a + off), umma_desc(b + off), idesc1, (p | ks) != 0);
                    }
                }
                umma_commit(d1_full);
#pragma unroll 1
                for (int c = 0; c < kBwZChunks; ++c) {
                    const int g = it * kBwZChunks + c, s = g % kBwZStages;
                    mbar_wait(z_full + s, (g / kBwZStages) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
                        const uint32_t a = aZ + (uint32_t)s * kBwZStage + (p == 2 ? kBwZHalf : 0u);
                        const uint32_t b = aM + (p == 1 ? kBwHalf : 0u) + (uint32_t)c * (kBwZRows * 128u);
#pragma unroll
                        for (int ks = 0; ks < kBwZRows / 8; ++ks)
                            umma_tf32(tmem_base + 128u, umma_desc_mn(a + ks * 1024u, kBwZBlk), umma_desc_mn(b + ks * 1024u, kBwBlk), idesc2,
                                      (c | p | ks) != 0);
                    }
                    umma_commit(z_empty + s);
                }
                umma_commit(ds_empty);
                umma_commit(d2_full);
            }
        }
    } else {
        const int bt = tid - 32, ty = bt >> 4, tx = bt & 15;
        // Wcat (hi, lo), once: B operand of the dZ product, row n = Wcat row (0..63 W_gcn, 64..127 W_bi), K = its 64 columns
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = bt + i * kBwBuilders;      // 128 rows x 16 chunks
            const int n = c >> 4, q = c & 15;
            const float4 w = ldg4((n < 64 ? Wg + (size_t)n * 64 : Wb + (size_t)(n - 64) * 64) + q * 4);
            const float wv[4] = {w.x, w.y, w.z, w.w};
            const uint32_t o = (uint32_t)(q >> 3) * kBwBlk + bw_sw(n, q & 7);
            bw_split_store(aW + o, aW + kBwHalf + o, wv);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(w_full);

        const int q = warp & 3, half = (warp - 1) >> 2;     // epilogue role: TMEM lane quadrant, column half
        const uint32_t tq = tmem_base + (((uint32_t)(q * 32)) << 16);
        // dWcat: the tensor core accumulates ONE tile (48 k-steps) at a time; the tiles are added on the CUDA cores in fp32
        // round-to-nearest into a running sum that lives in TMEM columns [192,256) (this thread's share: row q*32+lane, columns
        // half*32 ..).  Accumulating every tile of the CTA in the MMA accumulator lost 8e-6 relative at the amazon-book shape (the
        // accumulator truncates); per tile it stays at the ~1e-6 of the split itself.
        {
            uint32_t zero[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) zero[j] = 0u;
            tmem_st32(tq + 192u + (uint32_t)(half * 32), zero);
            tmem_st_wait();
        }
        // one round = 2 passes of 16 rows (16 lanes per row); S_pre is not read: where keep = 1 the sign of S is the sign of D
        // (D = LeakyReLU(S)/(1-p)), where keep = 0 the gradient is zero either way
        auto load_round = [&](int rbase, float4* d4, float4* o4, float4* x4, float4* k4) {
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
                const int r = rbase + pp * 16 + ty;
                d4[pp] = f4zero(); o4[pp] = f4zero(); x4[pp] = f4zero(); k4[pp] = make_float4(1.f, 1.f, 1.f, 1.f);
                if (r < N) {
                    d4[pp] = ldg4(D + (size_t)r * 64 + tx * 4);
                    o4[pp] = ldg4(dO + (size_t)r * dO_stride + tx * 4);
                    if (dD_ext) x4[pp] = ldg4(dD_ext + (size_t)r * 64 + tx * 4);
                    if (keep) k4[pp] = ldg4(keep + (size_t)r * 64 + tx * 4);
                }
            }
        };
        auto build_round = [&](int r0, int rrbase, const float4* d4, const float4* o4, const float4* x4, const float4* k4) {
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
                const int rr = rrbase + pp * 16 + ty;
                const bool in = r0 + rr < N;
                float ss = d4[pp].x * d4[pp].x + d4[pp].y * d4[pp].y + d4[pp].z * d4[pp].z + d4[pp].w * d4[pp].w;
                float dot = d4[pp].x * o4[pp].x + d4[pp].y * o4[pp].y + d4[pp].z * o4[pp].z + d4[pp].w * o4[pp].w;   // <D, dO>
#pragma unroll
                for (int m = 8; m >= 1; m >>= 1) { ss += __shfl_xor_sync(0xffffffffu, ss, m); dot += __shfl_xor_sync(0xffffffffu, dot, m); }
                const float nrm = fmaxf(sqrtf(ss), 1e-12f);
                const float proj = dot / (nrm * nrm);            // <O,dO>/|D| with O = D/|D|
                const float dv[4] = {d4[pp].x, d4[pp].y, d4[pp].z, d4[pp].w}, ov[4] = {o4[pp].x, o4[pp].y, o4[pp].z, o4[pp].w};
                const float xv[4] = {x4[pp].x, x4[pp].y, x4[pp].z, x4[pp].w}, kv[4] = {k4[pp].x, k4[pp].y, k4[pp].z, k4[pp].w};
                float ds[4], h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float dD = xv[j] + (ov[j] - dv[j] * proj) / nrm;
                    float v = keep ? dD * kv[j] * inv_keep : dD;
                    v *= (dv[j] > 0.f) ? 1.f : 0.2f;
                    ds[j] = in ? v : 0.f;
                    bw_split(ds[j], h[j], l[j]);
                }
                dbv.x += ds[0]; dbv.y += ds[1]; dbv.z += ds[2]; dbv.w += ds[3];
                const uint32_t ok = (uint32_t)(tx >> 3) * kBwBlk + bw_sw(rr, tx & 7), om = (uint32_t)(tx >> 3) * kBwBlk + bw_sw32(rr, tx & 7);
                bw_st4(aS + ok, h); bw_st4(aS + kBwHalf + ok, l);
                bw_st4(aM + om, h); bw_st4(aM + kBwHalf + om, l);
            }
        };
        // next tile's rows into L2 while this one is processed (bulk prefetch: the [N,64] streams are contiguous over a tile)
        auto prefetch_tile = [&](int rn0) {
            if (rn0 >= N) return;
            const int nrows = (N - rn0 < 128) ? N - rn0 : 128;
            const float* base = nullptr;
            if (bt == 0) base = D; else if (bt == 1) base = dD_ext; else if (bt == 2) base = keep; else if (bt == 3) base = side; else if (bt == 4) base = E;
            if (base) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + (size_t)rn0 * 64), "r"(nrows * 256) : "memory");
            if (bt >= 128 && bt - 128 < nrows) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(dO + (size_t)(rn0 + bt - 128) * dO_stride), "r"(256) : "memory");
        };
        float4 ad4[2], ao4[2], ax4[2], ak4[2];     // first round of the tile: loaded one phase ahead (previous tile's epilogue)
        load_round(blockIdx.x * 128, ad4, ao4, ax4, ak4);
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int r0 = t * 128;
            prefetch_tile(r0 + (int)gridDim.x * 128);
            // ---- dS: four rounds of 32 rows, each round's loads in flight while the one before it is built; then the Z loads
            mbar_wait(ds_empty, (it & 1) ^ 1);                    // the previous tile's dW product has read dS
            asm volatile("bar.sync 1, 256;" ::: "memory");       // ... and every builder has left the previous epilogue's staging
            float4 bd4[2], bo4[2], bx4[2], bk4[2];
            float4 sd4[8], e4[8];
            auto load_z = [&](int i0) {
#pragma unroll
                for (int i = i0; i < i0 + 4; ++i) {
                    const int r = r0 + i * 16 + ty;
                    sd4[i] = f4zero(); e4[i] = f4zero();
                    if (r < N) { sd4[i] = ldg4(side + (size_t)r * 64 + tx * 4); e4[i] = ldg4(E + (size_t)r * 64 + tx * 4); }
                }
            };
            load_round(r0 + 32, bd4, bo4, bx4, bk4);
            build_round(r0, 0, ad4, ao4, ax4, ak4);
            load_round(r0 + 64, ad4, ao4, ax4, ak4);
            build_round(r0, 32, bd4, bo4, bx4, bk4);
            load_round(r0 + 96, bd4, bo4, bx4, bk4);
            build_round(r0, 64, ad4, ao4, ax4, ak4);
            load_z(0);
            build_round(r0, 96, bd4, bo4, bx4, bk4);
            load_z(4);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(ds_full);
            if (it > 0) {   // dWcat of the previous tile (complete: ds_empty was waited on above)
                mbar_wait(d2_full, (it - 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int hc = 0; hc < 2; ++hc) {     // 16 columns at a time: the Z rows are live in registers here
                    uint32_t raw[16], run[16];
                    tmem_ld16(tq + 128u + (uint32_t)(half * 32 + hc * 16), raw);
                    tmem_ld16(tq + 192u + (uint32_t)(half * 32 + hc * 16), run);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) run[j] = __float_as_uint(__uint_as_float(run[j]) + __uint_as_float(raw[j]));
                    tmem_st16(tq + 192u + (uint32_t)(half * 32 + hc * 16), run);
                }
                tmem_st_wait();
                tc_fence_before();
            }
            // ---- Z = [side | E (*) side]: eight 16-row chunks through the ring (loads issued above)
#pragma unroll
            for (int c = 0; c < kBwZChunks; ++c) {
                const int g = it * kBwZChunks + c, s = g % kBwZStages;
                mbar_wait(z_empty + s, ((g / kBwZStages) & 1) ^ 1);
                const uint32_t zs = aZ + (uint32_t)s * kBwZStage;
                const float z1[4] = {sd4[c].x, sd4[c].y, sd4[c].z, sd4[c].w};
                const float z2[4] = {e4[c].x * sd4[c].x, e4[c].y * sd4[c].y, e4[c].z * sd4[c].z, e4[c].w * sd4[c].w};
                const uint32_t o = (uint32_t)(tx >> 3) * kBwZBlk + bw_sw32(ty, tx & 7);      // feature block tx/8 (side), 2 + tx/8 (E*side)
                bw_split_store(zs + o, zs + kBwZHalf + o, z1);
                bw_split_store(zs + 2 * kBwZBlk + o, zs + kBwZHalf + 2 * kBwZBlk + o, z2);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(z_full + s);
            }
            // ---- epilogue: dZ from TMEM (thread <-> row, this warp's column half) staged through the K-major dS image -- free once
            // the dZ product has completed -- so that the outputs leave as full 256 B rows; E and side are still in registers
            mbar_wait(d1_full, it & 1);
            tc_fence_after();
            {
                const int row = q * 32 + lane;
                const uint32_t rowbase = aS + (uint32_t)row * 256u + (uint32_t)half * 128u;
#pragma unroll
                for (int part = 0; part < 2; ++part) {     // dZ1 then dZ2: 32 registers at a time (the Z rows are live in registers here)
                    uint32_t zr[32];
                    tmem_ld32(tq + (uint32_t)(part * 64 + half * 32), zr);
                    tmem_ld_wait();
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {     // 16-byte chunk (half*8 + j4) of the row, low three bits XOR row: conflict-free both ways
                        const uint32_t o = rowbase + (uint32_t)part * kBwHalf + (uint32_t)(((j4 ^ row) & 7) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o), "r"(zr[j4 * 4]), "r"(zr[j4 * 4 + 1]), "r"(zr[j4 * 4 + 2]),
                                     "r"(zr[j4 * 4 + 3]) : "memory");
                    }
                }
                tc_fence_before();   // the next dZ product overwrites columns [0,128): ordered through the arrive on ds_full
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (t + (int)gridDim.x < ntiles) load_round((t + (int)gridDim.x) * 128, ad4, ao4, ax4, ak4);     // next tile's first round
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = i * 16 + ty, r = r0 + rr;
                const uint32_t o = aS + (uint32_t)rr * 256u + (uint32_t)(((tx & 8) | ((tx ^ rr) & 7)) << 4);
                float4 a, b;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(o));
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(o + kBwHalf));
                if (r < N) {
                    st4(dside + (size_t)r * 64 + tx * 4, make_float4(a.x + b.x * e4[i].x, a.y + b.y * e4[i].y, a.z + b.z * e4[i].z, a.w + b.w * e4[i].w));
                    st4(dE_direct + (size_t)r * 64 + tx * 4, make_float4(b.x * sd4[i].x, b.y * sd4[i].y, b.z * sd4[i].z, b.w * sd4[i].w));
                }
            }
        }
        // ---- dWcat partial of this CTA: last tile's accumulator, then the registers out (row = Wcat row, this warp's 32 columns)
        mbar_wait(d2_full, (it - 1) & 1);
        tc_fence_after();
        {
            uint32_t raw[32], run[32];
            tmem_ld32(tq + 128u + (uint32_t)(half * 32), raw);
            tmem_ld32(tq + 192u + (uint32_t)(half * 32), run);
            tmem_ld_wait();
            float* wp = dW_part + ((size_t)blockIdx.x * 128 + q * 32 + lane) * 64 + half * 32;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
                st4(wp + j4 * 4, make_float4(__uint_as_float(run[j4 * 4]) + __uint_as_float(raw[j4 * 4]), __uint_as_float(run[j4 * 4 + 1]) + __uint_as_float(raw[j4 * 4 + 1]),
                                             __uint_as_float(run[j4 * 4 + 2]) + __uint_as_float(raw[j4 * 4 + 2]), __uint_as_float(run[j4 * 4 + 3]) + __uint_as_float(raw[j4 * 4 + 3])));
        }
    }
    tc_fence_before();
    __syncthreads();   // every product has completed (d2_full): the ring is free to carry the db partials
    float* scratch = reinterpret_cast<float*>(sZ);
    if (warp >= 1) {
        const int bt = tid - 32, ty = bt >> 4, tx = bt & 15;
        *reinterpret_cast<float4*>(scratch + ty * 64 + tx * 4) = dbv;
    }
    __syncthreads();
    if (tid < 64) {
        float a = 0.f;
#pragma unroll
        for (int y = 0; y < 16; ++y) a += scratch[y * 64 + tid];
        db_part[(size_t)blockIdx.x * 64 + tid] = a;
    }
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kBwTmemCols) : "memory");
    }
}

// launched by idg_ngcf_dense_bwd (csrc/ngcf.cu); returns the number of per-CTA partials written (the reduce kernel's n_parts)
int ngcf_dense_bwd_tc(const float* E, const float* side, const float* Wg, const float* Wb, const float* keep, float inv_keep, const float* S_pre,
                      const float* D, const float* dO, int dO_stride, const float* dD_ext, int N, float* dside, float* dE_direct, float* dW_part,
                      float* db_part, int max_parts, int* n_parts, cudaStream_t stream) {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        IDG_CUDA(cudaGetDevice(&dev));
        IDG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int ntiles = (N + 127) / 128;
    int grid = sms < ntiles ? sms : ntiles;
    if (grid > max_parts) grid = max_parts;
    const size_t smem = (size_t)kBwSmem + 128;
    IDG_CUDA(cudaFuncSetAttribute(ngcf_dense_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ngcf_dense_bwd_tc_kernel<<<grid, 288, smem, stream>>>(E, side, Wg, Wb, keep, inv_keep, S_pre, D, dO, dO_stride, dD_ext, N, dside, dE_direct, dW_part,
                                                         db_part);
    IDG_LAUNCH_CHECK("ngcf_dense_bwd_tc_kernel");
    *n_parts = grid;
    return 0;
}

}  // namespace idg
