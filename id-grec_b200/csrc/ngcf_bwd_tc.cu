// NGCF dense layer, backward, on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Autograd of models/NGCF.py:87-106 for one layer (same algebra as ngcf_dense_bwd_kernel in csrc/ngcf.cu, which stays as the
// CUDA-core cross-check, IDG_NGCF_BWD=fma):
//   dD = dD_ext + (dO - O <O,dO>)/|D| ;  dS = dD (*) keep/(1-p) (*) (S > 0 ? 1 : 0.2)                 [N,64]   CUDA cores, per row
//   dZ = dS . Wcat^T                      [N,64] x [64,128]   (Wcat = [W_gcn ; W_bi], Z = [side | E (*) side])   tensor cores
//   dWcat = Z^T . dS                      [128,N] x [N,64]                                                        tensor cores
//   dside = dZ1 + dZ2 (*) E ;  dE_direct = dZ2 (*) side ;  db = colsum(dS)                                        CUDA cores
// Every product is the 3xTF32 split (xh yh + xh yl + xl yh with xh = x rounded to nearest tf32, fp32 accumulation in TMEM).
//
// Persistent CTAs (one per SM) walk 128-row tiles.
//   * The six input streams of a tile arrive through a four-slot cp.async ring (16-row units, three units = 48 KB in flight per
//     SM): every thread copies exactly the 16-byte pieces it later consumes, so the ring needs no barrier -- it is a register
//     file extension that keeps HBM busy while the previous units are processed (the first version loaded into registers and sat
//     at 26 % of the HBM peak on `long_scoreboard`).
//   * dS (hi, lo) is written in TWO shared-memory images of the same row-major [rows x 32 fp32] blocks: for dZ the rows are the N
//     index and the block is a K-major operand (128-byte swizzle); for dWcat = Z^T . dS the rows are the CONTRACTION index and the
//     blocks are MN-major operands -- tcgen05 takes 32-bit MN-major operands only in the "128-byte swizzle with 32-byte
//     atomicity" layout (cute: Layout_MN_SW128_32B_Atom, descriptor layout type 1).  Z is only ever contracted over rows: its
//     16-row chunks go through a two-stage ring in the same MN-major layout.  No transposes anywhere on the way in.
//   * Wcat (hi, lo) lives in TMEM for the whole kernel and is the A operand of the dZ product, which is therefore computed
//     transposed: dZ^T[128 features x 128 rows] = Wcat . dS^T.  The epilogue un-transposes it through the (by then idle) K-major
//     dS image so that dside / dE_direct leave as full 256 B rows.
//   * dWcat: the tensor core accumulates ONE tile (48 k-steps) at a time; the tiles are added on the CUDA cores in fp32
//     round-to-nearest into a running sum in TMEM.  (Accumulating every tile of the CTA in the MMA accumulator lost 8e-6 relative
//     at the amazon-book shape -- the accumulator truncates; per tile it stays at the ~1e-6 of the split.)  The running sum leaves
//     as a per-CTA partial, summed in CTA order by ngcf_reduce_kernel: deterministic.
//
//   warp 0      : TMEM allocation, one thread issues every tcgen05.mma
//   warps 1..8  : everything else (16 lanes per row for the streaming phases; thread <-> TMEM lane for the TMEM reads)
#include <math.h>

#include "tc_common.cuh"

namespace idg {

constexpr int kBwBuilders = 512;              // warps 1..16 (8 warps left the SM issue-bound: two warps per scheduler)
constexpr uint32_t kBwBlk = 128 * 128;        // bytes of one [128 rows x 32 fp32] block
constexpr uint32_t kBwHalf = 2 * kBwBlk;      // hi (or lo) of dS [128 x 64]: two blocks
constexpr int kBwZRows = 16;                  // rows per Z ring stage (2 k-steps of the dW product)
constexpr int kBwZChunks = 128 / kBwZRows;    // stages per tile
constexpr uint32_t kBwZBlk = kBwZRows * 128;  // [16 rows x 32 fp32]
constexpr uint32_t kBwZHalf = 4 * kBwZBlk;    // hi (or lo) of one stage: the four 32-feature blocks of Z
constexpr uint32_t kBwZStage = 2 * kBwZHalf;
constexpr int kBwZStages = 2;
constexpr uint32_t kBwUnit = 2 * 32 * 256;    // one staging slot: two [32 rows x 64 fp32] pieces
constexpr int kBwSlots = 4, kBwAhead = 3;     // units in flight per thread = kBwAhead
constexpr int kBwUnits = 12;                  // per tile: for k = 0..3, rows 32k..32k+31: (D | dO), (dD_ext | mask), (side | E)
// TMEM columns: dZ^T [0,128) | dWcat of the current tile [128,192) | running dWcat [192,256) | Wcat hi [256,320) | Wcat lo [320,384)
constexpr uint32_t kBwTmemCols = 512, kTcDz = 0, kTcD2 = 128, kTcRun = 192, kTcWh = 256, kTcWl = 320;
constexpr uint32_t kBwSmem = 4 * kBwHalf + kBwZStages * kBwZStage + kBwSlots * kBwUnit;   // dS K-major | dS MN-major | Z ring | staging = 224 KB

__device__ __forceinline__ uint32_t bw_sw32(int r, int cc) {   // 128 B-row block, 32-byte-atom swizzle: chunk pair (cc>>1) ^ (r & 3)
    return (uint32_t)r * 128u + (uint32_t)((((cc >> 1) ^ (r & 3)) << 5) | ((cc & 1) << 4));
}
__device__ __forceinline__ uint32_t bw_sw(int r, int cc) {     // 128 B-row block, 128-byte swizzle: chunk cc ^ (r & 7)
    return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((cc ^ (r & 7)) << 4);
}
__device__ __forceinline__ void bw_st4(uint32_t addr, const float* v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}
__device__ __forceinline__ float4 bw_ld4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
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

// KEEP: 0 = no dropout, 1 = float mask rows (keep), 2 = 64 mask bits per row (kbits); a template so that each variant stays inside
// the 168-register budget of a 288-thread CTA (one generic kernel spilled in the streaming loop: 110 -> 124 us)
template <int KEEP>
__global__ void __launch_bounds__(kBwBuilders + 32, 1) ngcf_dense_bwd_tc_kernel(const float* __restrict__ E, const float* __restrict__ side,
                                                                   const float* __restrict__ Wg, const float* __restrict__ Wb,
                                                                   const float* __restrict__ keep, const uint32_t* __restrict__ kbits, float inv_keep,
                                                                   const float* __restrict__ D,
                                                                   const float* __restrict__ dO, int dO_stride, const float* __restrict__ dD_ext, int N,
                                                                   float* __restrict__ dside, float* __restrict__ dE_direct,
                                                                   float* __restrict__ dW_part, float* __restrict__ db_part) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sS = smem;                          // dS hi | lo, K-major image (dZ product); epilogue staging afterwards
    unsigned char* sM = sS + 2 * kBwHalf;              // dS hi | lo, MN-major image (dWcat product)
    unsigned char* sZ = sM + 2 * kBwHalf;              // ring: stage s = Z chunk hi | lo
    unsigned char* sG = sZ + kBwZStages * kBwZStage;   // cp.async staging slots
    uint64_t* bars = reinterpret_cast<uint64_t*>(sG + kBwSlots * kBwUnit);
    uint64_t *w_full = bars, *ds_full = bars + 1, *d1_full = bars + 2;
    uint64_t *z_full = bars + 3, *z_empty = z_full + kBwZStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(z_empty + kBwZStages);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ntiles = (N + 127) / 128;

    if (tid == 0) {
        mbar_init(w_full, kBwBuilders); mbar_init(ds_full, kBwBuilders); mbar_init(d1_full, 1);
        for (int s = 0; s < kBwZStages; ++s) { mbar_init(z_full + s, kBwBuilders / 2); mbar_init(z_empty + s, 1); }   // a 16-row chunk is built by half the builders
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
    const uint32_t aS = smem_u32(sS), aM = smem_u32(sM), aZ = smem_u32(sZ), aG = smem_u32(sG);
    float4 dbv = f4zero();   // builders: column sums of dS over this thread's rows (columns tx*4 .. tx*4+3)

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t idesc1 = umma_idesc_tf32(128, 128);                             // dZ^T = Wcat . dS^T: A from TMEM, B K-major
            const uint32_t idesc2 = umma_idesc_tf32(128, 64) | (1u << 15) | (1u << 16);    // dWcat = Z^T . dS: both operands MN-major
            mbar_wait(w_full, 0);
            tc_fence_after();
            int it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
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
                            umma_tf32(tmem_base + kTcD2, umma_desc_mn(a + ks * 1024u, kBwZBlk), umma_desc_mn(b + ks * 1024u, kBwBlk), idesc2,
                                      (c | p | ks) != 0);
                    }
                    umma_commit(z_empty + s);
                }
                mbar_wait(ds_full, it & 1);
                tc_fence_after();
#pragma unroll
                for (int p = 0; p < 3; ++p) {      // Wh dSh + Wl dSh + Wh dSl
                    const uint32_t a = tmem_base + (p == 1 ? kTcWl : kTcWh), b = aS + (p == 2 ? kBwHalf : 0u);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_tf32_ts(tmem_base + kTcDz, a + (uint32_t)ks * 8u, umma_desc(b + (uint32_t)(ks >> 2) * kBwBlk + (uint32_t)(ks & 3) * 32u), idesc1,
                                     (p | ks) != 0);
                }
                umma_commit(d1_full);   // every product of the tile has completed: dZ^T and dWcat readable, both dS images free
            }
        }
    } else {
        const int bt = tid - 32, ty = bt >> 4, tx = bt & 15;    // 16 lanes per row, 32 rows per pass
        const int q = warp & 3, cq = (warp - 1) >> 2;           // TMEM role: lane quadrant, column quarter
        const uint32_t tq = tmem_base + (((uint32_t)(q * 32)) << 16);
        {   // Wcat (hi, lo) into TMEM, once: lane = Wcat row (0..63 W_gcn, 64..127 W_bi), this warp's 16 of its 64 columns; running dWcat = 0
            const int kk = q * 32 + lane;
            const float* wr = (kk < 64 ? Wg + (size_t)kk * 64 : Wb + (size_t)(kk - 64) * 64) + cq * 16;
            uint32_t h[16], l[16];
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 w = ldg4(wr + j4 * 4);
                const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float hh, ll;
                    bw_split(wv[j], hh, ll);
                    h[j4 * 4 + j] = __float_as_uint(hh); l[j4 * 4 + j] = __float_as_uint(ll);
                }
            }
            tmem_st16(tq + kTcWh + (uint32_t)(cq * 16), h);
            tmem_st16(tq + kTcWl + (uint32_t)(cq * 16), l);
#pragma unroll
            for (int j = 0; j < 16; ++j) h[j] = 0u;
            tmem_st16(tq + kTcRun + (uint32_t)(cq * 16), h);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(w_full);
        }
        // unit w of tile tt into staging slot w & 3 (kBwUnits is a multiple of kBwSlots); one commit group per unit, even when empty.
        // Units of the 32-row group k = w / 3: (D | dO), (dD_ext | mask), (side | E); a piece is [32 rows x 64 fp32] = 8 KB.
        auto issue = [&](int tt, int w) {
            if (tt < ntiles) {
                const uint32_t dst = aG + (uint32_t)(w & 3) * kBwUnit + (uint32_t)ty * 256u + (uint32_t)tx * 16u;
                const int k = w / 3, j = w % 3;
                const int r = tt * 128 + k * 32 + ty;
                const int ok = r < N ? 16 : 0;
                const size_t rc = r < N ? (size_t)r : 0;
                if (j == 0) {
                    cp_async16(dst, D + rc * 64 + tx * 4, ok);
                    cp_async16(dst + 8192u, dO + rc * dO_stride + tx * 4, ok);
                } else if (j == 1) {
                    if (dD_ext) cp_async16(dst, dD_ext + rc * 64 + tx * 4, ok);
                    if (KEEP == 1) cp_async16(dst + 8192u, keep + rc * 64 + tx * 4, ok);
                    else if (KEEP == 2)   // 64 bits per row: every lane of the row fetches the same 8 bytes into its own piece (the ring is thread-private)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst + 8192u), "l"(kbits + rc * 2), "r"(ok >> 1) : "memory");
                } else {
                    cp_async16(dst, side + rc * 64 + tx * 4, ok);
                    cp_async16(dst + 8192u, E + rc * 64 + tx * 4, ok);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
#pragma unroll
        for (int w = 0; w < kBwAhead; ++w) issue((int)blockIdx.x, w);

        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int r0 = t * 128;
            if (it > 0) asm volatile("bar.sync 1, %0;" ::"n"(kBwBuilders) : "memory");   // every builder has left the previous epilogue's staging (= the K-major image)
            float4 sd4[4], e4[4], d4, o4;
#pragma unroll
            for (int w = 0; w < kBwUnits; ++w) {
                if (w + kBwAhead < kBwUnits) issue(t, w + kBwAhead); else issue(t + (int)gridDim.x, w + kBwAhead - kBwUnits);
                asm volatile("cp.async.wait_group %0;" ::"n"(kBwAhead) : "memory");
                const uint32_t src = aG + (uint32_t)(w & 3) * kBwUnit + (uint32_t)ty * 256u + (uint32_t)tx * 16u;
                const int k = w / 3, j = w % 3;
                if (j == 0) {
                    d4 = bw_ld4(src); o4 = bw_ld4(src + 8192u);
                } else if (j == 1) {
                    // ---- dS, row 32k + ty (16 lanes per row): the row norm and <D,dO> by shuffles.  S_pre is not read: where keep = 1
                    // the sign of S is the sign of D (D = LeakyReLU(S)/(1-p)), where keep = 0 the gradient is zero either way
                    const int rr = k * 32 + ty;
                    const bool in = r0 + rr < N;
                    const float4 x4 = dD_ext ? bw_ld4(src) : f4zero();
                    float4 k4 = KEEP == 1 ? bw_ld4(src + 8192u) : make_float4(1.f, 1.f, 1.f, 1.f);
                    if (KEEP == 2) {
                        uint32_t w0, w1;
                        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w0), "=r"(w1) : "r"(src + 8192u) : "memory");
                        const uint32_t wd = ((tx >> 3) ? w1 : w0) >> ((tx & 7) * 4);
                        k4 = make_float4((float)(wd & 1u), (float)((wd >> 1) & 1u), (float)((wd >> 2) & 1u), (float)((wd >> 3) & 1u));
                    }
                    float ss = d4.x * d4.x + d4.y * d4.y + d4.z * d4.z + d4.w * d4.w;
                    float dot = d4.x * o4.x + d4.y * o4.y + d4.z * o4.z + d4.w * o4.w;   // <D, dO>
#pragma unroll
                    for (int m = 8; m >= 1; m >>= 1) { ss += __shfl_xor_sync(0xffffffffu, ss, m); dot += __shfl_xor_sync(0xffffffffu, dot, m); }
                    const float nrm = fmaxf(sqrtf(ss), 1e-12f);
                    const float proj = dot / (nrm * nrm);            // <O,dO>/|D| with O = D/|D|
                    const float dv[4] = {d4.x, d4.y, d4.z, d4.w}, ov[4] = {o4.x, o4.y, o4.z, o4.w};
                    const float xv[4] = {x4.x, x4.y, x4.z, x4.w}, kv[4] = {k4.x, k4.y, k4.z, k4.w};
                    float ds[4], h[4], l[4];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const float dD = xv[jj] + (ov[jj] - dv[jj] * proj) / nrm;
                        float v = KEEP ? dD * kv[jj] * inv_keep : dD;
                        v *= (dv[jj] > 0.f) ? 1.f : 0.2f;
                        ds[jj] = in ? v : 0.f;
                        bw_split(ds[jj], h[jj], l[jj]);
                    }
                    dbv.x += ds[0]; dbv.y += ds[1]; dbv.z += ds[2]; dbv.w += ds[3];
                    const uint32_t ok = (uint32_t)(tx >> 3) * kBwBlk + bw_sw(rr, tx & 7), om = (uint32_t)(tx >> 3) * kBwBlk + bw_sw32(rr, tx & 7);
                    bw_st4(aS + ok, h); bw_st4(aS + kBwHalf + ok, l);
                    bw_st4(aM + om, h); bw_st4(aM + kBwHalf + om, l);
                } else {
                    // ---- Z = [side | E (*) side], rows 32k .. 32k+31 = chunks 2k (the lanes with ty < 16) and 2k+1 (the others), each
                    // into its own ring stage (the dS rows they are contracted with were written above; the fence covers them)
                    sd4[k] = bw_ld4(src);
                    e4[k] = bw_ld4(src + 8192u);
                    const int s = ty >> 4;                      // chunk 2k + s lives in stage s (kBwZChunks is even)
                    mbar_wait(z_empty + s, ((it * (kBwZChunks / 2) + k) & 1) ^ 1);
                    const uint32_t zs = aZ + (uint32_t)s * kBwZStage;
                    const float z1[4] = {sd4[k].x, sd4[k].y, sd4[k].z, sd4[k].w};
                    const float z2[4] = {e4[k].x * sd4[k].x, e4[k].y * sd4[k].y, e4[k].z * sd4[k].z, e4[k].w * sd4[k].w};
                    const uint32_t o = (uint32_t)(tx >> 3) * kBwZBlk + bw_sw32(ty & 15, tx & 7);      // feature block tx/8 (side), 2 + tx/8 (E*side)
                    bw_split_store(zs + o, zs + kBwZHalf + o, z1);
                    bw_split_store(zs + 2 * kBwZBlk + o, zs + kBwZHalf + 2 * kBwZBlk + o, z2);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(z_full + s);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(ds_full);
            // ---- epilogue, after every product of the tile has completed
            mbar_wait(d1_full, it & 1);
            tc_fence_after();
            {   // running dWcat += this tile's (row q*32+lane of Wcat, this warp's 16 columns)
                uint32_t raw[16], run[16];
                tmem_ld16(tq + kTcD2 + (uint32_t)(cq * 16), raw);
                tmem_ld16(tq + kTcRun + (uint32_t)(cq * 16), run);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) run[j] = __float_as_uint(__uint_as_float(run[j]) + __uint_as_float(raw[j]));
                tmem_st16(tq + kTcRun + (uint32_t)(cq * 16), run);
            }
            {   // dZ^T: lane = feature kk (0..63 -> dZ1, 64..127 -> dZ2), columns = rows; un-transpose through the K-major dS image:
                // plane kk/64, row-major [128 rows x 64], 16-byte chunk index XOR row (conflict-free for these stores and the reads below)
                const int kk = q * 32 + lane;
                const uint32_t pbase = aS + (uint32_t)(kk >> 6) * kBwHalf + (uint32_t)(kk & 3) * 4u;
                const int chunk = (kk & 63) >> 2;
                uint32_t zr[32];
                tmem_ld32(tq + kTcDz + (uint32_t)(cq * 32), zr);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int row = cq * 32 + j;
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(pbase + (uint32_t)row * 256u + (uint32_t)(((chunk & 8) | ((chunk ^ row) & 7)) << 4)), "r"(zr[j]) : "memory");
                }
            }
            tmem_st_wait();
            tc_fence_before();   // the next tile's products overwrite dZ^T and the tile accumulator: ordered through the arrives on z_full / ds_full
            asm volatile("bar.sync 1, %0;" ::"n"(kBwBuilders) : "memory");
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rr = i * 32 + ty, r = r0 + rr;
                const uint32_t o = aS + (uint32_t)rr * 256u + (uint32_t)(((tx & 8) | ((tx ^ rr) & 7)) << 4);
                const float4 a = bw_ld4(o), b = bw_ld4(o + kBwHalf);
                if (r < N) {
                    st4(dside + (size_t)r * 64 + tx * 4, make_float4(a.x + b.x * e4[i].x, a.y + b.y * e4[i].y, a.z + b.z * e4[i].z, a.w + b.w * e4[i].w));
                    st4(dE_direct + (size_t)r * 64 + tx * 4, make_float4(b.x * sd4[i].x, b.y * sd4[i].y, b.z * sd4[i].z, b.w * sd4[i].w));
                }
            }
        }
        // ---- dWcat partial of this CTA: the running sum (row = Wcat row, this warp's 16 columns)
        tc_fence_after();
        {
            uint32_t run[16];
            tmem_ld16(tq + kTcRun + (uint32_t)(cq * 16), run);
            tmem_ld_wait();
            float* wp = dW_part + ((size_t)blockIdx.x * 128 + q * 32 + lane) * 64 + cq * 16;
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4)
                st4(wp + j4 * 4, make_float4(__uint_as_float(run[j4 * 4]), __uint_as_float(run[j4 * 4 + 1]), __uint_as_float(run[j4 * 4 + 2]), __uint_as_float(run[j4 * 4 + 3])));
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();   // every product has completed (the builders waited on d1_full of the last tile): the ring is free to carry the db partials
    float* scratch = reinterpret_cast<float*>(sZ);
    if (warp >= 1) {
        const int bt = tid - 32, ty = bt >> 4, tx = bt & 15;
        *reinterpret_cast<float4*>(scratch + ty * 64 + tx * 4) = dbv;     // [32][64]
    }
    __syncthreads();
    if (tid < 64) {
        float a = 0.f;
#pragma unroll
        for (int y = 0; y < kBwBuilders / 16; ++y) a += scratch[y * 64 + tid];
        db_part[(size_t)blockIdx.x * 64 + tid] = a;
    }
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kBwTmemCols) : "memory");
    }
}

// launched by idg_ngcf_dense_bwd (csrc/ngcf.cu); returns the number of per-CTA partials written (the reduce kernel's n_parts)
int ngcf_dense_bwd_tc(const float* E, const float* side, const float* Wg, const float* Wb, const float* keep, const uint32_t* keep_bits, float inv_keep,
                      const float* S_pre,
                      const float* D, const float* dO, int dO_stride, const float* dD_ext, int N, float* dside, float* dE_direct, float* dW_part,
                      float* db_part, int max_parts, int* n_parts, cudaStream_t stream) {
    (void)S_pre;   // the sign of S is read off D (see the kernel)
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
    auto launch = [&](auto kern) -> int {
        IDG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kBwBuilders + 32, smem, stream>>>(E, side, Wg, Wb, keep, keep_bits, inv_keep, D, dO, dO_stride, dD_ext, N, dside, dE_direct, dW_part, db_part);
        return 0;
    };
    if (int rc = keep_bits ? launch(ngcf_dense_bwd_tc_kernel<2>) : (keep ? launch(ngcf_dense_bwd_tc_kernel<1>) : launch(ngcf_dense_bwd_tc_kernel<0>))) return rc;
    IDG_LAUNCH_CHECK("ngcf_dense_bwd_tc_kernel");
    *n_parts = grid;
    return 0;
}

}  // namespace idg
