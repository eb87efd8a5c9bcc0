// tcgen05 (5th-gen tensor core) building blocks for sm_100a, hand-written inline PTX.
//
// Numerics: "3xTF32" error-compensated fp32 GEMM.  Every fp32 operand v is split into
//   hi = v with the low 13 mantissa bits cleared (exactly representable in TF32)
//   lo = v - hi (exact in fp32; the tensor core keeps its top 11 significant bits)
// and  A.B ~= A_hi.B_hi + A_lo.B_hi + A_hi.B_lo  accumulated in fp32 in tensor memory.
// The dropped A_lo.B_lo term and the truncation of lo are O(2^-21) relative per product, the
// same order as fp32 rounding, so results stay within the 1e-4 parity bound with margin.
//
// Operand layout: the canonical K-major SWIZZLE_128B shared-memory layout of UMMA
// (cute::UMMA::Layout_K_SW128_Atom): a "K atom" holds 32 fp32 (128 bytes) of K for all rows,
// rows at a 128-byte pitch, 8-row groups 1024 bytes apart, and inside each 128-byte row the
// 16-byte chunk index is XOR-ed with (row & 7).  Buffers must be 1024-byte aligned.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace gnnb {
namespace tc {

constexpr int ATOM_K = 32;                 // fp32 elements of K per 128-byte swizzle atom
constexpr int ROW_BYTES = 128;
constexpr int MMA_K = 8;                   // K per tcgen05.mma.kind::tf32

// byte offset of element (row, k) inside an operand buffer of `rows` rows (rows % 8 == 0)
__host__ __device__ __forceinline__ uint32_t canon_offset(int row, int k, int rows)
{
    const int atom = k >> 5, kk = k & 31;
    const int chunk = (kk >> 2) ^ (row & 7);
    return (uint32_t)atom * (uint32_t)rows * ROW_BYTES + (uint32_t)row * ROW_BYTES +
           (uint32_t)chunk * 16u + (uint32_t)(kk & 3) * 4u;
}
// byte offset of the 16-byte chunk holding (row, k..k+3), k % 4 == 0
__host__ __device__ __forceinline__ uint32_t canon_chunk_offset(int row, int k, int rows)
{
    return canon_offset(row, k, rows);
}

__host__ __device__ __forceinline__ float tf32_hi(float v)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(__float_as_uint(v) & 0xffffe000u);
#else
    union { float f; uint32_t u; } x;
    x.f = v;
    x.u &= 0xffffe000u;
    return x.f;
#endif
}

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- bulk async copy global -> shared (TMA engine, no tensor map) -------------------------
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes,
                                         uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
            "r"(smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void bulk_g2s_addr(uint32_t smem_dst_addr, const void *gsrc, uint32_t bytes,
                                              uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
            "r"(smem_dst_addr),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void *gptr)
{
    asm volatile("prefetch.global.L2 [%0];\n" ::"l"(gptr));
}

// generic-proxy writes (st.shared) -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}

// ---- tensor memory ------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_slot, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     smem_u32(smem_slot)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// ---- descriptors ----------------------------------------------------------------------------
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major, 1) |
//   [32,46) SBO >> 4 = 1024 B between 8-row groups | [46,48) version = 1 | [61,64) layout = 2
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) |
           (2ull << 61);
}
// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, A/B = TF32,
// both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued MMAs have completed
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::
                     "r"(smem_u32(bar))
                 : "memory");
}

// 32 consecutive fp32 accumulator columns of this thread's row (TMEM lane = 32*(warp%4) + lane)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

#endif  // __CUDACC__

// ==== second-generation building blocks: aggregation on the tensor cores ======================
// The per-tile aggregation  AGG = ADJ . X  (ADJ = in-edge multiplicities of the tile's rows, a
// block-diagonal 128x128 matrix of small integers) runs as kind::f16 MMAs on bf16 operands:
//   ADJ   bf16, exact (entries <= 255), A operand, K-major SWIZZLE_128B: 2 K atoms (64 source
//         nodes each) of 128 rows x 128 bytes
//   X     split into three bf16 planes hi + mid + lo (8 + 8 + 8 mantissa bits, truncation, every
//         residual exact in fp32 => |x - (h+m+l)| <= 2^-24 |x|), B operand, MN-major
//         SWIZZLE_128B: storage [node][feature], blocks of 64 features, 128-byte rows per node
// so AGG = ADJ.Xh + ADJ.Xm + ADJ.Xl accumulated in fp32 is fp32-grade.  The node-transform GEMMs
// then take their A operand (hi / lo TF32 parts) from TENSOR MEMORY, written by tcgen05.st from
// the thread that owns the row, so no shared memory is spent on activations.
constexpr int PLANE_BLOCK_BYTES = 128 * ROW_BYTES;   // 64 features (or 64 sources) x 128 rows
constexpr int PLANE_BYTES = 2 * PLANE_BLOCK_BYTES;   // 128 features

// byte offset of the 16-byte chunk holding features [f, f+8) of `node` in a bf16 plane (f % 8 == 0)
__host__ __device__ __forceinline__ uint32_t plane_chunk_offset(int node, int f)
{
    const int block = f >> 6, chunk = (f & 63) >> 3;
    return (uint32_t)block * PLANE_BLOCK_BYTES + (uint32_t)node * ROW_BYTES +
           (uint32_t)((chunk ^ (node & 7)) << 4);
}
// byte offset of the 16-byte chunk holding sources [s, s+8) of destination row `dst` (s % 8 == 0)
__host__ __device__ __forceinline__ uint32_t adj_chunk_offset(int dst, int s)
{
    return plane_chunk_offset(dst, s);   // same physical atom shape: 128-byte rows, XOR (row & 7)
}

#ifdef __CUDACC__
__device__ __forceinline__ float bf16_hi_part(float v)
{
    return __uint_as_float(__float_as_uint(v) & 0xffff0000u);
}
// 8 fp32 values -> three packed chunks of 8 bf16 (hi, mid, lo planes)
__device__ __forceinline__ void split3_pack8(const float (&v)[8], uint4 &h, uint4 &m, uint4 &l)
{
    uint32_t hb[8], mb[8], lb[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float hf = bf16_hi_part(v[j]);
        const float r1 = v[j] - hf;
        const float mf = bf16_hi_part(r1);
        const float r2 = r1 - mf;
        hb[j] = __float_as_uint(hf); mb[j] = __float_as_uint(mf); lb[j] = __float_as_uint(r2);
    }
    h = make_uint4(__byte_perm(hb[0], hb[1], 0x7632), __byte_perm(hb[2], hb[3], 0x7632),
                   __byte_perm(hb[4], hb[5], 0x7632), __byte_perm(hb[6], hb[7], 0x7632));
    m = make_uint4(__byte_perm(mb[0], mb[1], 0x7632), __byte_perm(mb[2], mb[3], 0x7632),
                   __byte_perm(mb[4], mb[5], 0x7632), __byte_perm(mb[6], mb[7], 0x7632));
    l = make_uint4(__byte_perm(lb[0], lb[1], 0x7632), __byte_perm(lb[2], lb[3], 0x7632),
                   __byte_perm(lb[4], lb[5], 0x7632), __byte_perm(lb[6], lb[7], 0x7632));
}
// the inverse: three chunks -> 8 fp32 values (exact)
__device__ __forceinline__ void join3_unpack8(const uint4 &h, const uint4 &m, const uint4 &l,
                                              float (&v)[8])
{
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, mw[4] = {m.x, m.y, m.z, m.w},
                   lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        v[2 * j] = (__uint_as_float(hw[j] << 16) + __uint_as_float(mw[j] << 16)) +
                   __uint_as_float(lw[j] << 16);
        v[2 * j + 1] = (__uint_as_float(hw[j] & 0xffff0000u) + __uint_as_float(mw[j] & 0xffff0000u)) +
                       __uint_as_float(lw[j] & 0xffff0000u);
    }
}

// MN-major SWIZZLE_128B descriptor (cute::UMMA::make_umma_desc<Major::MN>): LBO = byte distance
// between 64-element blocks along MN, SBO = byte distance between groups of 8 rows along K
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes,
                                                 uint32_t sbo_bytes)
{
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: fp32 accumulate, A = B = BF16, A K-major, B K- or MN-major
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N, int b_mn_major)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(b_mn_major & 1) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand from tensor memory (row = lane, K element j = column j of the 32-bit A region)
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::f16 with the A operand in tensor memory (16-bit elements; see tc_test.cu's probe of how
// they are laid out in the 32-bit TMEM cells)
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 consecutive 32-bit columns of this thread's row: registers -> tensor memory
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::
            "r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
        "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
        "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])),
        "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])),
        "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
        "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])),
        "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
// one lane of a converged warp (warp-uniform code keeps descriptors in uniform registers; only the
// tcgen05 / bulk-copy instruction itself is predicated on the elected lane)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.b32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
// 16 consecutive 32-bit columns of this thread's row: registers -> tensor memory
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
        "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
        "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void tmem_st(uint32_t taddr, const float (&v)[32]) { tmem_st32(taddr, v); }
__device__ __forceinline__ void tmem_st(uint32_t taddr, const float (&v)[16]) { tmem_st16(taddr, v); }
__device__ __forceinline__ void tmem_st_wait()
{
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}
// tcgen05.ld without the wait (issue several, then tmem_ld_wait once)
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_nowait(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32_nowait(taddr, r); }
__device__ __forceinline__ void tmem_ld_nowait(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16_nowait(taddr, r); }
__device__ __forceinline__ void tmem_ld_wait()
{
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
#endif  // __CUDACC__ (second-generation blocks)


}  // namespace tc
}  // namespace gnnb
