// gram_tc.cuh -- tcgen05 / TMEM / TMA kernels of the Gram path (sm_100a only).
//
//   forward   G_b = F_b F_b^T          (style_transfer.py:31-35)   both operands K-major (HW contiguous)
//   backward  dF_b = S_b F_b           (autograd of the above)     A = F^T tile (MN-major), B = S (K-major)
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA
// issuer, warps 2-5 = epilogue (TMEM -> registers -> global).  Operands are fp32 in shared memory,
// 128-byte swizzled exactly as TMA writes them, consumed by tcgen05.mma kind::tf32; accumulation is
// fp32 in TMEM.  Every mbarrier wait is bounded: a protocol bug traps instead of hanging the GPU.
#pragma once
#include <atomic>
#include <cuda.h>
#include <stdlib.h>

#include "gram_common.cuh"

namespace st3d {
namespace tc {

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    long long start = 0;
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        const long long now = clock64();
        if (start == 0) start = now;
        if (now - start > 4000000000ll) __trap();  // ~2 s: protocol error, fail loudly
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y), "r"(z)
        : "memory");
}
// multicast variants: the box lands at the same shared-memory offset of every CTA in `mask`, and signals the
// mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y, int z,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y), "r"(z), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void mma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// per-thread asynchronous 16-byte copies global -> shared (the fused backward epilogue prefetches with them)
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- CTA-pair (cta_group::2) variants: two CTAs of a cluster on one TPC share every MMA ----------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// the box lands in THIS CTA's shared memory; the bytes are counted on the mbarrier at `bar_cluster` (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int x, int y,
                                                 int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(x), "r"(y), "r"(z)
        : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t slot_smem) {  // one warp of EACH CTA of the pair, same warp id
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// M = 256 across the pair: each CTA supplies its 128 rows of A and half of the N rows of B from the same
// shared-memory offsets, and receives its 128 rows of D in its own TMEM.  Issued by the leader CTA only.
__device__ __forceinline__ void mma_tf32_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit_pair(uint32_t bar, uint16_t mask) {  // arrives in every CTA of `mask`
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, fp32 operands read as tf32
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive columns: r[j] = TMEM[lane_base + laneid][col + j]
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&r)[32]) {
    uint32_t* u = reinterpret_cast<uint32_t*>(r);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
          "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
          "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
          "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- descriptors (bit layout: PTX ISA "tcgen05 shared memory / instruction descriptor") ---------------
// 128-byte swizzled operand tile.  lbo / sbo in bytes.  layout: 2 = SWIZZLE_128B (16-byte chunks, the
// K-major operands), 1 = SWIZZLE_128B_BASE32B (32-byte chunks; the only layout tcgen05 accepts for an
// MN-major tf32 operand; TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint64_t layout = 2) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) /* sm_100 descriptor version */ |
           (layout << 61);
}
// kind::tf32, fp32 accumulate; a_mn / b_mn = 1 selects the MN-major operand layout
constexpr uint32_t instr_desc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int kThreads = 192;
constexpr int kScratchFloats = 4 * 32 * 36;  // one padded (16-byte aligned rows) 32x32 transpose tile per epilogue warp

// =====================================================================================================
// forward
// =====================================================================================================
template <int C>
struct FwdCfg {
    static constexpr int M = C == 64 ? 64 : 128;
    static constexpr int PANELS = C == 256 ? 2 : 1;  // row panels accumulated by one CTA
    static constexpr int GROUPS = C == 512 ? 4 : 1;  // CTAs per (image, split)
    static constexpr int NHALF = C == 512 ? 2 : 1;   // an MMA covers at most N = 256 columns
    static constexpr int NMMA = C / NHALF;
    static constexpr int TMEM_COLS = C == 64 ? 64 : (C == 128 ? 128 : 512);
    // k-blocks (32 fp32 columns) per pipeline stage: the boxes of one stage are issued back to back, so DRAM
    // sees KPS*128 contiguous bytes per feature row instead of isolated 128-byte bursts 4*HW bytes apart
    static constexpr int KPS = C == 64 ? 4 : (C == 128 ? 2 : 1);
    static constexpr int SLAB_BYTES = C * 128;       // C rows x 32 fp32
    static constexpr int STAGE_BYTES = KPS * SLAB_BYTES;
    static constexpr int STAGES = C == 64 ? 5 : (C == 128 ? 5 : (C == 256 ? 6 : 3));
    // C = 512: the four panel CTAs of one (image, split) form a thread-block cluster; each loads its own quarter of
    // the slab (its 128 channels) and TMA-multicasts it to all four, so the slab leaves L2 once instead of four times
    static constexpr bool CLUSTER = C == 512;
    static constexpr int BOX_ROWS = CLUSTER ? 128 : (C < 256 ? C : 256);
    static constexpr int BOXES = C / BOX_ROWS;
    static constexpr size_t SMEM = 1024 + (size_t)STAGES * STAGE_BYTES + kScratchFloats * 4 + (2 * STAGES + 2) * 8 + 16;
};

// grid (GROUPS, splits, B).  partials [B][splits][C][C].
// NHWC = false: features (B, C, HW), HW contiguous  -> both operands K-major, plain 128-byte swizzle.
// NHWC = true : features (B, HW, C), C contiguous (torch channels_last, what cuDNN's tensor-core convolutions
//               produce) -> the contraction runs over the row index, both operands are MN-major and use the
//               32-byte-atom swizzle; a stage is ROWS = 32*KPS pixel rows, stored as C/32 boxes of
//               [ROWS][32 channels] (one fully contiguous ROWS*C*4-byte region of DRAM).
template <int C, bool NHWC>
__global__ void __launch_bounds__(kThreads, 1)
k_gram_tc_fwd(const __grid_constant__ CUtensorMap map, float* __restrict__ partials, int splits, int64_t k_chunk,
              int64_t HW, const GramEpilogue ep) {
    using Cfg = FwdCfg<C>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* stages = smem;
    float* scratch = reinterpret_cast<float*>(smem + (size_t)Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + kScratchFloats);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 1);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + Cfg::STAGES),
                   tmem_full = smem_u32(bars + 2 * Cfg::STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x, s = blockIdx.y, b = blockIdx.z;
    const int64_t k0 = (int64_t)s * k_chunk, k1 = min(HW, k0 + k_chunk);
    const int nkb = (int)((k1 - k0 + 31) / 32);
    const int nst = (nkb + Cfg::KPS - 1) / Cfg::KPS;  // pipeline stages (KPS k-blocks each)

    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::STAGES; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, Cfg::CLUSTER ? 4 : 1);  // cluster: every CTA's MMAs must have released the slot
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
        tma_prefetch_desc(&map);
    }
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    if (Cfg::CLUSTER) cluster_sync_all();  // peers' barriers are initialised before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < nst; ++it) {
                const int st = it % Cfg::STAGES;
                const uint32_t ph = (it / Cfg::STAGES) & 1;
                mbar_wait(empty0 + 8 * st, ph ^ 1);
                mbar_arrive_expect_tx(full0 + 8 * st, Cfg::STAGE_BYTES);
                const uint32_t dst = smem_u32(stages + (size_t)st * Cfg::STAGE_BYTES);
                if (NHWC) {
                    constexpr int ROWS = 32 * Cfg::KPS;
                    const int x = (int)(k0 + (int64_t)it * ROWS);  // rows >= HW are zero-filled by the 3D map
                    if (Cfg::CLUSTER) {
#pragma unroll
                        for (int cg = 4 * g; cg < 4 * g + 4; ++cg)
                            tma_load_3d_mc(dst + cg * ROWS * 128, &map, full0 + 8 * st, cg * 32, x, b, 0xF);
                    } else {
#pragma unroll
                        for (int cg = 0; cg < C / 32; ++cg)
                            tma_load_3d(dst + cg * ROWS * 128, &map, full0 + 8 * st, cg * 32, x, b);
                    }
                } else {
#pragma unroll
                    for (int kk = 0; kk < Cfg::KPS; ++kk) {
                        // the split owns [k0, k1), a whole number of stages except for the last split of an
                        // image, whose tail boxes lie beyond HW and come back zero-filled
                        const int x = (int)(k0 + ((int64_t)it * Cfg::KPS + kk) * 32);
                        if (Cfg::CLUSTER) {
                            tma_load_2d_mc(dst + kk * Cfg::SLAB_BYTES + g * Cfg::BOX_ROWS * 128, &map, full0 + 8 * st, x,
                                           b * C + g * Cfg::BOX_ROWS, 0xF);
                        } else {
#pragma unroll
                            for (int bx = 0; bx < Cfg::BOXES; ++bx)
                                tma_load_2d(dst + kk * Cfg::SLAB_BYTES + bx * Cfg::BOX_ROWS * 128, &map, full0 + 8 * st, x,
                                            b * C + bx * Cfg::BOX_ROWS);
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc(Cfg::M, Cfg::NMMA, NHWC ? 1 : 0, NHWC ? 1 : 0);
            for (int it = 0; it < nst; ++it) {
                const int st = it % Cfg::STAGES;
                const uint32_t ph = (it / Cfg::STAGES) & 1;
                mbar_wait(full0 + 8 * st, ph);
                tc_fence_after();
                if (NHWC) {
                    constexpr uint32_t BOX = 32 * Cfg::KPS * 128;  // one [ROWS][32 channels] box
                    const uint32_t sbase = smem_u32(stages + (size_t)st * Cfg::STAGE_BYTES);
#pragma unroll
                    for (int kg = 0; kg < 4 * Cfg::KPS; ++kg) {  // 8 pixel rows per MMA
#pragma unroll
                        for (int p = 0; p < Cfg::PANELS; ++p) {
                            const uint32_t grp_a = (C == 512 ? g : p) * 4;  // first 32-channel group of the panel
                            const uint64_t ad = smem_desc(sbase + grp_a * BOX + kg * 1024, BOX, 512, 1);
#pragma unroll
                            for (int h = 0; h < Cfg::NHALF; ++h) {
                                const uint64_t bd = smem_desc(sbase + h * 8 * BOX + kg * 1024, BOX, 512, 1);
                                mma_tf32(tmem_base + p * 256 + h * 256, ad, bd, idesc, (it > 0 || kg > 0) ? 1u : 0u);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int kk = 0; kk < Cfg::KPS; ++kk) {
                        const uint32_t sbase = smem_u32(stages + (size_t)st * Cfg::STAGE_BYTES) + kk * Cfg::SLAB_BYTES;
#pragma unroll
                        for (int p = 0; p < Cfg::PANELS; ++p) {
                            const uint32_t row_a = (C == 512 ? g : p) * 128;  // 0 when C <= 128
#pragma unroll
                            for (int h = 0; h < Cfg::NHALF; ++h) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint64_t ad = smem_desc(sbase + row_a * 128 + k * 32, 16, 1024);
                                    const uint64_t bd = smem_desc(sbase + h * 256 * 128 + k * 32, 16, 1024);
                                    mma_tf32(tmem_base + p * 256 + h * 256, ad, bd, idesc,
                                             (it > 0 || kk > 0 || k > 0) ? 1u : 0u);
                                }
                            }
                        }
                    }
                }
                // frees the smem slot once these MMAs have read it (in every CTA of the cluster when multicasting)
                if (Cfg::CLUSTER) mma_commit_mc(empty0 + 8 * st, 0xF); else mma_commit(empty0 + 8 * st);
            }
            mma_commit(tmem_full);
        }
        __syncwarp();
    } else {
        const int q = warp & 3;  // TMEM lane quarter this warp may read
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        float* P = partials + ((int64_t)b * splits + s) * C * C;
        float* sc = scratch + q * 32 * 36;
        float r[32];
#pragma unroll 1
        for (int p = 0; p < Cfg::PANELS; ++p) {
            const int panel_row0 = (C == 512 ? g : p) * 128;
            // M = 128: TMEM lane = row.  M = 64: rows 16q..16q+15 live in lanes 32q..32q+15.
            const int rows = Cfg::M == 128 ? 32 : 16;
            const int row0 = Cfg::M == 128 ? panel_row0 + q * 32 : q * 16;
#pragma unroll 1
            for (int c0 = 0; c0 < C; c0 += 32) {
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p * 256 + c0), r);
                // transpose through padded shared memory with 128-bit accesses: one warp store then covers
                // 4 rows x 128 contiguous bytes of the partial tile
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(sc + lane * 36 + 4 * j) =
                        make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int t = 4 * i + (lane >> 3);
                    if (t < rows)
                        *reinterpret_cast<float4*>(P + (int64_t)(row0 + t) * C + c0 + 4 * (lane & 7)) =
                            *reinterpret_cast<const float4*>(sc + t * 36 + 4 * (lane & 7));
                }
                __syncwarp();
            }
        }
        if (ep.fused) __threadfence();  // this CTA's partial tile is visible device-wide before its arrival is counted
    }
    tc_fence_before();
    __syncthreads();
    if (Cfg::CLUSTER) cluster_sync_all();  // no CTA leaves while peers may still multicast into it / arrive on its barriers
    if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    if (!ep.fused) return;

    // ---- split-K reduction + style-loss epilogue, inside the GEMM kernel --------------------------------------------
    // The n = splits * GROUPS CTAs of image b are all resident (the launcher sizes the grid to at most one CTA per SM),
    // so each one counts its arrival, waits for the others, and then reduces ITS 1/n slice of the C x C tile over the
    // splits in a fixed order (deterministic sum): no second kernel, no host-visible seam between GEMM and loss.
    const int n = splits * Cfg::GROUPS;
    if (n > 1) {
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(ep.counters + b, 1u);
            long long start = 0;
            while (*reinterpret_cast<volatile unsigned*>(ep.counters + b) < (unsigned)n) {
                __nanosleep(40);
                const long long now = clock64();
                if (start == 0) start = now;
                if (now - start > 4000000000ll) __trap();  // ~2 s: a CTA of this image never ran, fail loudly
            }
            __threadfence();
        }
        __syncthreads();
    }
    constexpr int64_t CC = (int64_t)C * C;
    const int j = s * Cfg::GROUPS + g;
    const int64_t per = ((CC / 4 + n - 1) / n) * 4;  // elements per CTA, whole float4s
    const int64_t e_lo = min(CC, (int64_t)j * per), e_hi = min(CC, e_lo + per);
    const float* P0 = partials + (int64_t)b * splits * CC;
    const float* tg = ep.target ? ep.target + (ep.Bt == 1 ? 0 : (int64_t)b * CC) : nullptr;
    float lsum = 0.0f;
    // U independent float4 columns per thread and round: U x splits loads in flight (the partials sit in L2; one
    // column at a time would pay one L2 round trip per split and column, the whole tail latency-bound)
    constexpr int U = 4;
    for (int64_t e = e_lo + 4 * (int64_t)threadIdx.x; e < e_hi; e += 4 * kThreads * U) {
        float4 acc[U], tv[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t ee = e + (int64_t)u * 4 * kThreads;
            ok[u] = ee < e_hi;
            acc[u] = ok[u] ? __ldcg(reinterpret_cast<const float4*>(P0 + ee)) : make_float4(0.f, 0.f, 0.f, 0.f);
            tv[u] = (ok[u] && tg) ? __ldg(reinterpret_cast<const float4*>(tg + ee)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int sp = 1; sp < splits; ++sp) {
            float4 t[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                t[u] = ok[u] ? __ldcg(reinterpret_cast<const float4*>(P0 + (int64_t)sp * CC + e + (int64_t)u * 4 * kThreads))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < U; ++u) { acc[u].x += t[u].x; acc[u].y += t[u].y; acc[u].z += t[u].z; acc[u].w += t[u].w; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const int64_t ee = e + (int64_t)u * 4 * kThreads;
            if (ep.gram) *reinterpret_cast<float4*>(ep.gram + (int64_t)b * CC + ee) = acc[u];
            if (tg) {
                float4 d = make_float4(acc[u].x - tv[u].x, acc[u].y - tv[u].y, acc[u].z - tv[u].z, acc[u].w - tv[u].w);
                lsum += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
                const float s2 = 2.0f * ep.scale;
                d.x *= s2; d.y *= s2; d.z *= s2; d.w *= s2;
                if (ep.dgram) *reinterpret_cast<float4*>(ep.dgram + (int64_t)b * CC + ee) = d;
            }
        }
    }
    if (tg && ep.loss_out) {
        lsum = warp_sum(lsum);
        float* red = scratch;  // the transpose scratch is free by now
        if (lane == 0) red[warp] = lsum;
        __syncthreads();
        if (threadIdx.x == 0) {
            float v = 0.0f;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) v += red[w];
            if (v != 0.0f) atomicAdd(ep.loss_out, v * ep.scale);
        }
    }
}

// =====================================================================================================
// backward: D[x (128 lanes)][c (C columns)] = sum_j F[j][x] S[c][j]
// =====================================================================================================
// RING: the epilogue owns a cp.async prefetch ring (fused accumulate / ReLU-mask tail).  Always at C <= 128; at C = 256
// the ring takes the place of two of the four TMA stages, so the launcher picks it only when the tail is requested.
template <int C, bool RING = (C <= 128)>
struct BwdCfg {
    static_assert(!RING || C <= 256, "no shared memory left for a prefetch ring at C = 512");
    static constexpr int NHALF = C == 512 ? 2 : 1;
    static constexpr int NMMA = C / NHALF;
    static constexpr int ACC = C <= 256 ? 2 : 1;  // TMEM accumulator buffers (epilogue overlaps the next chunk)
    static constexpr int TMEM_COLS = C * ACC < 32 ? 32 : C * ACC;  // 128, 256, 512, 512
    static constexpr int F_BYTES = 4 * 4096;      // four [32 j][32 x] boxes = 128 x's
    static constexpr int S_BYTES = C * 128;       // [C rows c][32 j]
    static constexpr int STAGE_BYTES = F_BYTES + S_BYTES;
    // The fused epilogue (ST3D_GRAM_ACCUMULATE / ST3D_GRAM_RELU_MASK) prefetches the incoming gradient and the mask
    // values through a per-warp cp.async ring, RING_SLOTS batches of 8 KB deep.  At C = 256 the ring costs two TMA
    // stages: with the tail an item's epilogue (eight 32-channel groups, each a read-modify-write of global memory)
    // bounds the kernel and its 2 us of MMAs are fed well enough by two stages (0.128 -> 0.083 ms at 8 x 256 x 16384);
    // without the tail four stages are faster (58 vs 73 us), hence the second configuration.
    static constexpr int RING_SLOTS = RING ? 3 : 0;
    static constexpr int RING_BYTES = 4 * RING_SLOTS * 8192;
    static constexpr int STAGES = C == 64 ? 4 : (C == 128 ? 3 : (C == 256 ? (RING ? 2 : 4) : 2));
    static constexpr int KB = C / 32;             // k-blocks (32 channels j each) per chunk
    static constexpr int S_BOX_ROWS = C < 256 ? C : 256;
    static constexpr int S_BOXES = C / S_BOX_ROWS;
    static constexpr int TR_FLOATS = 4 * 32 * 36;  // NHWC epilogue transpose tiles (one per epilogue warp)
    static constexpr size_t SMEM = 1024 + (size_t)STAGES * STAGE_BYTES + TR_FLOATS * 4 + RING_BYTES +
                                   (2 * STAGES + 2 * ACC) * 8 + 16;
};

// NHWC = false: A = F^T tile from rows j of (B, C, HW): MN-major, 32-byte-atom swizzle, four [32 j][32 x] boxes.
// NHWC = true : A = F tile [128 x rows][32 j] of (B, HW, C): K-major, plain 128-byte swizzle, one box; the
//               epilogue thread owns one pixel row and stores 32 consecutive channels (128 contiguous bytes).
template <int C, bool NHWC, bool RING>
__global__ void __launch_bounds__(kThreads, 1)
k_gram_tc_bwd(const __grid_constant__ CUtensorMap map_f, const __grid_constant__ CUtensorMap map_s,
              const float* __restrict__ feat, float* __restrict__ grad_feat, int B, int64_t HW, int accumulate,
              float out_scale, const float* __restrict__ out_scale_dev) {
    using Cfg = BwdCfg<C, RING>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* stages = smem;
    float* tr_scratch = reinterpret_cast<float*>(smem + (size_t)Cfg::STAGES * Cfg::STAGE_BYTES);
    uint8_t* ring = reinterpret_cast<uint8_t*>(tr_scratch + Cfg::TR_FLOATS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + Cfg::RING_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 2 * Cfg::ACC);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + Cfg::STAGES),
                   accf0 = smem_u32(bars + 2 * Cfg::STAGES), acce0 = smem_u32(bars + 2 * Cfg::STAGES + Cfg::ACC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t chunks = (HW + 127) / 128, items = (int64_t)B * chunks;

    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::STAGES; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, 1);
        }
        for (int i = 0; i < Cfg::ACC; ++i) {
            mbar_init(accf0 + 8 * i, 1);
            mbar_init(acce0 + 8 * i, 4);  // one arrival per epilogue warp
        }
        fence_barrier_init();
        tma_prefetch_desc(&map_f);
        tma_prefetch_desc(&map_s);
    }
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
                const int b = (int)(item / chunks);
                const int x0 = (int)(item % chunks) * 128;
                for (int kb = 0; kb < Cfg::KB; ++kb, ++it) {
                    const int st = it % Cfg::STAGES;
                    const uint32_t ph = (it / Cfg::STAGES) & 1;
                    mbar_wait(empty0 + 8 * st, ph ^ 1);
                    mbar_arrive_expect_tx(full0 + 8 * st, Cfg::STAGE_BYTES);
                    const uint32_t dst = smem_u32(stages + (size_t)st * Cfg::STAGE_BYTES);
                    if (NHWC) {
                        tma_load_3d(dst, &map_f, full0 + 8 * st, kb * 32, x0, b);  // [128 x][32 j], rows >= HW zero
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            tma_load_2d(dst + i * 4096, &map_f, full0 + 8 * st, x0 + 32 * i, b * C + kb * 32);
                    }
#pragma unroll
                    for (int bx = 0; bx < Cfg::S_BOXES; ++bx)
                        tma_load_2d(dst + Cfg::F_BYTES + bx * Cfg::S_BOX_ROWS * 128, &map_s, full0 + 8 * st, kb * 32,
                                    b * C + bx * Cfg::S_BOX_ROWS);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc(128, Cfg::NMMA, NHWC ? 0 : 1, 0);  // NCHW: A = F^T tile, MN-major
            uint32_t it = 0, li = 0;
            for (int64_t item = blockIdx.x; item < items; item += gridDim.x, ++li) {
                const uint32_t a = li % Cfg::ACC, aph = (li / Cfg::ACC) & 1;
                mbar_wait(acce0 + 8 * a, aph ^ 1);  // epilogue has drained this accumulator
                tc_fence_after();
                for (int kb = 0; kb < Cfg::KB; ++kb, ++it) {
                    const int st = it % Cfg::STAGES;
                    const uint32_t ph = (it / Cfg::STAGES) & 1;
                    mbar_wait(full0 + 8 * st, ph);
                    tc_fence_after();
                    const uint32_t sF = smem_u32(stages + (size_t)st * Cfg::STAGE_BYTES), sS = sF + Cfg::F_BYTES;
#pragma unroll
                    for (int kg = 0; kg < 4; ++kg) {  // 8 channels j per MMA
                        // NCHW: 4 M-atoms (32 x's) 4096 B apart; each MMA spans two 4-row K-atoms 512 B apart
                        // NHWC: K-major rows of 128 B, 8 channels (32 B) per MMA
                        const uint64_t ad = NHWC ? smem_desc(sF + kg * 32, 16, 1024) : smem_desc(sF + kg * 1024, 4096, 512, 1);
#pragma unroll
                        for (int h = 0; h < Cfg::NHALF; ++h) {
                            const uint64_t bd = smem_desc(sS + h * 256 * 128 + kg * 32, 16, 1024);
                            mma_tf32(tmem_base + a * C + h * 256, ad, bd, idesc, (kb > 0 || kg > 0) ? 1u : 0u);
                        }
                    }
                    mma_commit(empty0 + 8 * st);
                }
                mma_commit(accf0 + 8 * a);
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        uint32_t li = 0;
        float r[32];
        const float osc = out_scale * (out_scale_dev ? __ldg(out_scale_dev) : 1.0f);
        // Fused tail, C <= 128: batch k = (k / G)-th item of this CTA, channel group k % G.  The warp keeps
        // RING_SLOTS - 1 batches of its own operands (old gradient, mask values: 8 x 16 B of each per lane) in
        // flight with cp.async, so the HBM round trip of the read-modify-write is paid once, not per batch.
        constexpr int G = C / 32;
        constexpr bool kRing = NHWC && Cfg::RING_SLOTS > 0;
        const bool use_ring = kRing && accumulate != 0;
        const uint32_t ring_q = smem_u32(ring) + (uint32_t)q * (Cfg::RING_SLOTS * 8192);
        const uint32_t my_items = blockIdx.x < items ? (uint32_t)((items - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
        const uint32_t total_k = my_items * G;
        const uint32_t chunks32 = (uint32_t)chunks;  // items = B * chunks < 2^31 (checked by the launcher)
        auto prefetch = [&](uint32_t k) {
            if (k < total_k) {
                const uint32_t it2 = blockIdx.x + (k / G) * gridDim.x;
                const uint32_t b2 = it2 / chunks32;
                const int64_t xw2 = (int64_t)(it2 - b2 * chunks32) * 128 + q * 32;
                // this lane's first row (lane >> 3) and 16-byte column (lane & 7); the eight copies are 4 rows apart
                const int64_t off = ((int64_t)b2 * HW + xw2 + (lane >> 3)) * C + (int)(k % G) * 32 + 4 * (lane & 7);
                const uint32_t slot = ring_q + (k % (Cfg::RING_SLOTS > 0 ? Cfg::RING_SLOTS : 1)) * 8192 + lane * 16;
                const float* po = grad_feat + off;
                const float* pf = feat + off;
                if (accumulate == 3 && xw2 + 32 <= HW) {  // the common case: both operands, no ragged last chunk
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        cp_async16(slot + i * 512, po + i * 4 * C);
                        cp_async16(slot + 4096 + i * 512, pf + i * 4 * C);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (xw2 + 4 * i + (lane >> 3) < HW) {
                            if (accumulate & 1) cp_async16(slot + i * 512, po + i * 4 * C);
                            if (accumulate & 2) cp_async16(slot + 4096 + i * 512, pf + i * 4 * C);
                        }
                    }
                }
            }
            cp_async_commit();  // an empty group keeps the group count in step with k
        };
        uint32_t kbatch = 0;
        if (use_ring) {
#pragma unroll
            for (int d = 0; d < Cfg::RING_SLOTS - 1; ++d) prefetch((uint32_t)d);
        }
        for (int64_t item = blockIdx.x; item < items; item += gridDim.x, ++li) {
            const uint32_t a = li % Cfg::ACC, aph = (li / Cfg::ACC) & 1;
            const int b = (int)((uint32_t)item / chunks32);
            const int64_t xchunk = (int64_t)((uint32_t)item - (uint32_t)b * chunks32) * 128;
            const int64_t x = xchunk + q * 32 + lane;
            mbar_wait(accf0 + 8 * a, aph);
            tc_fence_after();
            float* out = NHWC ? grad_feat + ((int64_t)b * HW + x) * C : grad_feat + (int64_t)b * C * HW + x;
#pragma unroll 1
            for (int c0 = 0; c0 < C; c0 += 32) {
                if (use_ring) {  // batch kbatch is the oldest of the RING_SLOTS - 1 groups in flight once this one is queued
                    prefetch(kbatch + Cfg::RING_SLOTS - 1);
                    cp_async_wait<(Cfg::RING_SLOTS > 0 ? Cfg::RING_SLOTS - 1 : 0)>();
                }
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * C + c0), r);
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] *= osc;  // 1 unless S is the raw, symmetric dG (ST3D_GRAM_DGRAM_SYMMETRIC)
                if (NHWC) {
                    // r[j] = D[pixel row = lane][channel c0 + j].  Transpose through padded shared memory so that
                    // one warp store covers 4 pixel rows x 128 contiguous bytes instead of 32 rows x 16 bytes.
                    float* sc = tr_scratch + q * 32 * 36;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(sc + lane * 36 + 4 * j) =
                            make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                    __syncwarp();
                    const int64_t xw = xchunk + q * 32;  // first pixel row of this warp
                    float* ow = grad_feat + ((int64_t)b * HW + xw) * C + c0 + 4 * (lane & 7);
                    if (accumulate == 0) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int t = 4 * i + (lane >> 3);
                            if (xw + t < HW)
                                *reinterpret_cast<float4*>(ow + (int64_t)t * C) =
                                    *reinterpret_cast<const float4*>(sc + t * 36 + 4 * (lane & 7));
                        }
                    } else if (use_ring && accumulate == 3 && xw + 32 <= HW) {
                        // fused tail, common case: operands from this lane's ring slot, constant offsets throughout
                        const float4* so = reinterpret_cast<const float4*>(
                            ring + ((size_t)q * (kRing ? Cfg::RING_SLOTS : 1) + kbatch % (kRing ? Cfg::RING_SLOTS : 1)) * 8192 +
                            lane * 16);
                        const float* st = sc + (lane >> 3) * 36 + 4 * (lane & 7);
                        float* og = ow + (int64_t)(lane >> 3) * C;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float4 v = *reinterpret_cast<const float4*>(st + i * 4 * 36);
                            const float4 o = so[i * 32], f = so[256 + i * 32];
                            v.x = f.x <= 0.0f ? 0.0f : v.x + o.x;  // ReLU backward of (incoming gradient + dF)
                            v.y = f.y <= 0.0f ? 0.0f : v.y + o.y;
                            v.z = f.z <= 0.0f ? 0.0f : v.z + o.z;
                            v.w = f.w <= 0.0f ? 0.0f : v.w + o.w;
                            *reinterpret_cast<float4*>(og + i * 4 * C) = v;
                        }
                    } else {
                        // fused elementwise tail: all global loads of the eight row groups are issued before the
                        // first store, so one memory round trip covers them (a load after a store to the same
                        // array could not be hoisted over it)
                        float4 oldv[8], fv[8];
                        const float* fw = feat + (ow - grad_feat);
                        const uint8_t* slot = ring + ((size_t)q * (kRing ? Cfg::RING_SLOTS : 1) +
                                                      kbatch % (kRing ? Cfg::RING_SLOTS : 1)) * 8192 + lane * 16;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int t = 4 * i + (lane >> 3);
                            const bool ok = xw + t < HW;
                            if (use_ring) {  // this lane's own copies, complete after cp_async_wait above
                                oldv[i] = (ok && (accumulate & 1)) ? *reinterpret_cast<const float4*>(slot + i * 512)
                                                                   : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                                fv[i] = (ok && (accumulate & 2)) ? *reinterpret_cast<const float4*>(slot + 4096 + i * 512)
                                                                 : make_float4(1.0f, 1.0f, 1.0f, 1.0f);
                            } else {
                                oldv[i] = (ok && (accumulate & 1)) ? *reinterpret_cast<const float4*>(ow + (int64_t)t * C)
                                                                   : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                                fv[i] = (ok && (accumulate & 2)) ? __ldg(reinterpret_cast<const float4*>(fw + (int64_t)t * C))
                                                                 : make_float4(1.0f, 1.0f, 1.0f, 1.0f);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int t = 4 * i + (lane >> 3);
                            if (xw + t < HW) {
                                float4 v = *reinterpret_cast<const float4*>(sc + t * 36 + 4 * (lane & 7));
                                if (accumulate & 1) {
                                    v.x += oldv[i].x; v.y += oldv[i].y; v.z += oldv[i].z; v.w += oldv[i].w;
                                }
                                // ReLU backward of the layer that produced feat: zero where feat <= 0
                                if (fv[i].x <= 0.0f) v.x = 0.0f;
                                if (fv[i].y <= 0.0f) v.y = 0.0f;
                                if (fv[i].z <= 0.0f) v.z = 0.0f;
                                if (fv[i].w <= 0.0f) v.w = 0.0f;
                                *reinterpret_cast<float4*>(ow + (int64_t)t * C) = v;
                            }
                        }
                    }
                    ++kbatch;
                    __syncwarp();
                } else if (x < HW) {
                    const float* fin = feat + (out - grad_feat);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int64_t o = (int64_t)(c0 + j) * HW;
                        float v = (accumulate & 1) ? out[o] + r[j] : r[j];
                        if ((accumulate & 2) && __ldg(fin + o) <= 0.0f) v = 0.0f;
                        out[o] = v;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acce0 + 8 * a);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// =====================================================================================================
// backward at C = 512 on CTA pairs (cta_group::2)
// =====================================================================================================
// One CTA streams 80 KB per k-block for C = 512 (16 KB of F, 64 KB of S): two stages fit, and two stages cannot
// cover the TMA round trip (tensor pipe 38-44 % active).  A CTA PAIR shares the S slab: each CTA holds its own
// 128 pixel rows of F and HALF of the S rows of every N = 256 MMA, 48 KB per k-block, four stages deep, and the
// slab leaves L2 once per 256 pixels instead of once per 128.  Item = (image, two consecutive 128-pixel chunks).
// RING: the epilogue warps own a cp.async prefetch ring for the fused accumulate / ReLU-mask tail (as BwdCfg at C <= 256),
// paid for with two of the four TMA stages.  With the tail an item's epilogue -- sixteen 32-channel groups, each a
// read-modify-write of global memory -- is 3/4 of the item's time (the single 512-column accumulator serialises MMAs and
// epilogue), so that is where the shared memory earns more; without the tail the four-stage configuration stays.
template <bool RING>
struct BwdPairCfgT {
    static constexpr int C = 512;
    static constexpr int STAGES = RING ? 2 : 4;
    static constexpr int F_BYTES = 4 * 4096;         // this CTA's [128 x][32 j] tile of F
    static constexpr int S_HALF_BYTES = 128 * 128;   // [128 rows c][32 j] of S for one N = 256 MMA
    static constexpr int STAGE_BYTES = F_BYTES + 2 * S_HALF_BYTES;  // 48 KB per CTA
    static constexpr int KB = C / 32;
    static constexpr int TR_FLOATS = 4 * 32 * 36;
    static constexpr int RING_SLOTS = RING ? 3 : 0;
    static constexpr int RING_BYTES = 4 * RING_SLOTS * 8192;
    static constexpr int NBARS = 2 * STAGES + 2;     // full[], empty[], acc_full, acc_empty
    static constexpr size_t SMEM = 1024 + (size_t)STAGES * STAGE_BYTES + TR_FLOATS * 4 + RING_BYTES + NBARS * 8 + 16;
};
using BwdPairCfg = BwdPairCfgT<false>;

template <bool NHWC, bool RING>
__global__ void __launch_bounds__(kThreads, 1)
k_gram_tc_bwd_pair(const __grid_constant__ CUtensorMap map_f, const __grid_constant__ CUtensorMap map_s,
                   const float* __restrict__ feat, float* __restrict__ grad_feat, int B, int64_t HW, int accumulate,
                   float out_scale, const float* __restrict__ out_scale_dev) {
    using Cfg = BwdPairCfgT<RING>;
    static_assert(!RING || NHWC, "the prefetch ring serves channels_last features only");
    constexpr int C = Cfg::C;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* stages = smem;
    float* tr_scratch = reinterpret_cast<float*>(smem + (size_t)Cfg::STAGES * Cfg::STAGE_BYTES);
    uint8_t* ring = reinterpret_cast<uint8_t*>(tr_scratch + Cfg::TR_FLOATS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + Cfg::RING_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::NBARS);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + Cfg::STAGES),
                   accf = smem_u32(bars + 2 * Cfg::STAGES), acce = smem_u32(bars + 2 * Cfg::STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();  // 0 = leader: issues every MMA and owns the full / acc_empty barriers
    const int64_t chunks = (HW + 127) / 128, pairs = (chunks + 1) / 2, items = (int64_t)B * pairs;
    const int64_t cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::STAGES; ++i) {
            mbar_init(full0 + 8 * i, 1);   // leader's copy is used: one arrive.expect_tx + the bytes of both CTAs
            mbar_init(empty0 + 8 * i, 1);  // one multicast commit per use, in each CTA
        }
        mbar_init(accf, 1);
        mbar_init(acce, 8);                // four epilogue warps of each CTA (leader's copy is used)
        fence_barrier_init();
        tma_prefetch_desc(&map_f);
        tma_prefetch_desc(&map_s);
    }
    if (warp == 1) tmem_alloc_pair<512>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers exist before anything can arrive on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t item = cluster_id; item < items; item += nclusters) {
                const int b = (int)(item / pairs);
                const int x0 = (int)((item % pairs) * 2 + rank) * 128;  // a chunk past HW reads zeros
                for (int kb = 0; kb < Cfg::KB; ++kb, ++it) {
                    const int st = it % Cfg::STAGES;
                    const uint32_t ph = (it / Cfg::STAGES) & 1;
                    mbar_wait(empty0 + 8 * st, ph ^ 1);
                    if (rank == 0) mbar_arrive_expect_tx(full0 + 8 * st, 2 * Cfg::STAGE_BYTES);
                    const uint32_t lead_full = mapa_shared(full0 + 8 * st, 0);
                    const uint32_t dst = smem_u32(stages + (size_t)st * Cfg::STAGE_BYTES);
                    if (NHWC) {
                        tma_load_3d_pair(dst, &map_f, lead_full, kb * 32, x0, b);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            tma_load_2d_pair(dst + i * 4096, &map_f, lead_full, x0 + 32 * i, b * C + kb * 32);
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h)  // this CTA's half of the S rows of MMA h
                        tma_load_2d_pair(dst + Cfg::F_BYTES + h * Cfg::S_HALF_BYTES, &map_s, lead_full, kb * 32,
                                         b * C + h * 256 + (int)rank * 128);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = instr_desc(256, 256, NHWC ? 0 : 1, 0);
            uint32_t it = 0, li = 0;
            for (int64_t item = cluster_id; item < items; item += nclusters, ++li) {
                mbar_wait(acce, (li & 1) ^ 1);  // both epilogues have drained the accumulator
                tc_fence_after();
                for (int kb = 0; kb < Cfg::KB; ++kb, ++it) {
                    const int st = it % Cfg::STAGES;
                    const uint32_t ph = (it / Cfg::STAGES) & 1;
                    mbar_wait(full0 + 8 * st, ph);
                    tc_fence_after();
                    const uint32_t sF = smem_u32(stages + (size_t)st * Cfg::STAGE_BYTES), sS = sF + Cfg::F_BYTES;
#pragma unroll
                    for (int kg = 0; kg < 4; ++kg) {
                        const uint64_t ad = NHWC ? smem_desc(sF + kg * 32, 16, 1024) : smem_desc(sF + kg * 1024, 4096, 512, 1);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const uint64_t bd = smem_desc(sS + h * Cfg::S_HALF_BYTES + kg * 32, 16, 1024);
                            mma_tf32_pair(tmem_base + h * 256, ad, bd, idesc, (kb > 0 || kg > 0) ? 1u : 0u);
                        }
                    }
                    mma_commit_pair(empty0 + 8 * st, 0x3);  // frees the slot in both CTAs
                }
                mma_commit_pair(accf, 0x3);
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const uint32_t lead_acce = mapa_shared(acce, 0);
        uint32_t li = 0;
        float r[32];
        const float osc = out_scale * (out_scale_dev ? __ldg(out_scale_dev) : 1.0f);
        // Fused tail with a ring (RING): batch k = (k / G)-th item of this CTA pair, channel group k % G.  Each warp keeps
        // RING_SLOTS - 1 batches of its own operands (old gradient, mask values: 8 x 16 B of each per lane) in flight with
        // cp.async -- also across the MMA phase of the next item, during which the epilogue warps would otherwise idle.
        constexpr int G = C / 32;
        constexpr int kSlots = RING ? Cfg::RING_SLOTS : 1;
        const bool use_ring = RING && accumulate != 0;
        const uint32_t ring_q = smem_u32(ring) + (uint32_t)q * (kSlots * 8192);
        const int64_t my_items = cluster_id < items ? (items - cluster_id + nclusters - 1) / nclusters : 0;
        const uint32_t total_k = (uint32_t)(my_items * G);
        auto prefetch = [&](uint32_t k) {
            if (RING && k < total_k) {
                const int64_t it2 = cluster_id + (int64_t)(k / G) * nclusters;
                const int64_t b2 = it2 / pairs;
                const int64_t xw2 = ((it2 % pairs) * 2 + rank) * 128 + q * 32;
                // this lane's first row (lane >> 3) and 16-byte column (lane & 7); the eight copies are 4 rows apart
                const int64_t off = (b2 * HW + xw2 + (lane >> 3)) * C + (int)(k % G) * 32 + 4 * (lane & 7);
                const uint32_t slot = ring_q + (k % kSlots) * 8192 + lane * 16;
                const float* po = grad_feat + off;
                const float* pf = feat + off;
                if (accumulate == 3 && xw2 + 32 <= HW) {  // the common case: both operands, no ragged last chunk
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        cp_async16(slot + i * 512, po + i * 4 * C);
                        cp_async16(slot + 4096 + i * 512, pf + i * 4 * C);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (xw2 + 4 * i + (lane >> 3) < HW) {
                            if (accumulate & 1) cp_async16(slot + i * 512, po + i * 4 * C);
                            if (accumulate & 2) cp_async16(slot + 4096 + i * 512, pf + i * 4 * C);
                        }
                    }
                }
            }
            cp_async_commit();  // an empty group keeps the group count in step with k
        };
        uint32_t kbatch = 0;
        if (use_ring) {
#pragma unroll
            for (int d = 0; d < kSlots - 1; ++d) prefetch((uint32_t)d);
        }
        for (int64_t item = cluster_id; item < items; item += nclusters, ++li) {
            const int b = (int)(item / pairs);
            const int64_t xc = ((item % pairs) * 2 + rank) * 128;  // first pixel of this CTA's chunk
            const int64_t x = xc + q * 32 + lane;
            mbar_wait(accf, li & 1);
            tc_fence_after();
            float* out = grad_feat + (int64_t)b * C * HW + x;  // NCHW
#pragma unroll 1
            for (int c0 = 0; c0 < C; c0 += 32) {
                if (use_ring) {  // batch kbatch is the oldest of the kSlots - 1 groups in flight once this one is queued
                    prefetch(kbatch + kSlots - 1);
                    cp_async_wait<(RING ? Cfg::RING_SLOTS - 1 : 0)>();
                }
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] *= osc;  // 1 unless S is the raw, symmetric dG (ST3D_GRAM_DGRAM_SYMMETRIC)
                if (NHWC) {
                    float* sc = tr_scratch + q * 32 * 36;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(sc + lane * 36 + 4 * j) =
                            make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                    __syncwarp();
                    const int64_t xw = xc + q * 32;
                    float* ow = grad_feat + ((int64_t)b * HW + xw) * C + c0 + 4 * (lane & 7);
                    if (accumulate == 0) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int t = 4 * i + (lane >> 3);
                            if (xw + t < HW)
                                *reinterpret_cast<float4*>(ow + (int64_t)t * C) =
                                    *reinterpret_cast<const float4*>(sc + t * 36 + 4 * (lane & 7));
                        }
                    } else if (use_ring && accumulate == 3 && xw + 32 <= HW) {
                        // fused tail, common case: operands from this lane's ring slot, constant offsets throughout
                        const float4* so = reinterpret_cast<const float4*>(
                            ring + ((size_t)q * kSlots + kbatch % kSlots) * 8192 + lane * 16);
                        const float* st = sc + (lane >> 3) * 36 + 4 * (lane & 7);
                        float* og = ow + (int64_t)(lane >> 3) * C;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float4 v = *reinterpret_cast<const float4*>(st + i * 4 * 36);
                            const float4 o = so[i * 32], f = so[256 + i * 32];
                            v.x = f.x <= 0.0f ? 0.0f : v.x + o.x;  // ReLU backward of (incoming gradient + dF)
                            v.y = f.y <= 0.0f ? 0.0f : v.y + o.y;
                            v.z = f.z <= 0.0f ? 0.0f : v.z + o.z;
                            v.w = f.w <= 0.0f ? 0.0f : v.w + o.w;
                            *reinterpret_cast<float4*>(og + i * 4 * C) = v;
                        }
                    } else {
                        // fused elementwise tail: all loads of the eight row groups are issued before the first store,
                        // so one memory round trip covers them (a load after a store to the same array could not be
                        // hoisted over it); with the ring they are this lane's own copies, complete after the wait above
                        float4 oldv[8], fv[8];
                        const float* fw = feat + (ow - grad_feat);
                        const uint8_t* slot = ring + ((size_t)q * kSlots + kbatch % kSlots) * 8192 + lane * 16;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int t = 4 * i + (lane >> 3);
                            const bool ok = xw + t < HW;
                            if (use_ring) {
                                oldv[i] = (ok && (accumulate & 1)) ? *reinterpret_cast<const float4*>(slot + i * 512)
                                                                   : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                                fv[i] = (ok && (accumulate & 2)) ? *reinterpret_cast<const float4*>(slot + 4096 + i * 512)
                                                                 : make_float4(1.0f, 1.0f, 1.0f, 1.0f);
                            } else {
                                oldv[i] = (ok && (accumulate & 1)) ? *reinterpret_cast<const float4*>(ow + (int64_t)t * C)
                                                                   : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                                fv[i] = (ok && (accumulate & 2)) ? __ldg(reinterpret_cast<const float4*>(fw + (int64_t)t * C))
                                                                 : make_float4(1.0f, 1.0f, 1.0f, 1.0f);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int t = 4 * i + (lane >> 3);
                            if (xw + t < HW) {
                                float4 v = *reinterpret_cast<const float4*>(sc + t * 36 + 4 * (lane & 7));
                                if (accumulate & 1) {
                                    v.x += oldv[i].x; v.y += oldv[i].y; v.z += oldv[i].z; v.w += oldv[i].w;
                                }
                                // ReLU backward of the layer that produced feat: zero where feat <= 0
                                if (fv[i].x <= 0.0f) v.x = 0.0f;
                                if (fv[i].y <= 0.0f) v.y = 0.0f;
                                if (fv[i].z <= 0.0f) v.z = 0.0f;
                                if (fv[i].w <= 0.0f) v.w = 0.0f;
                                *reinterpret_cast<float4*>(ow + (int64_t)t * C) = v;
                            }
                        }
                    }
                    ++kbatch;
                    __syncwarp();
                } else if (x < HW) {
                    const float* fin = feat + (out - grad_feat);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int64_t o = (int64_t)(c0 + j) * HW;
                        float v = (accumulate & 1) ? out[o] + r[j] : r[j];
                        if ((accumulate & 2) && __ldg(fin + o) <= 0.0f) v = 0.0f;
                        out[o] = v;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_acce);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // neither CTA frees tensor memory or exits while the other may still use the pair
    if (warp == 1) tmem_dealloc_pair<512>(tmem_base);
}

// ---- host side -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled is a DRIVER entry point: it needs a context current on the calling thread, and unlike a runtime
// call it does not make one so.  A thread whose first CUDA work is one of these Gram ops -- autograd's backward thread,
// when the node it starts with is a style tap -- has none yet (CUDA_ERROR_INVALID_CONTEXT, 201): one runtime call per
// thread and device binds the primary context.
static inline void bind_primary_context() {
    thread_local int bound_device = -1;
    int d = -1;
    if (cudaGetDevice(&d) == cudaSuccess && d != bound_device) {
        cudaFree(nullptr);
        bound_device = d;
    }
}

static inline EncodeTiledFn encode_fn() {
    bind_primary_context();
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// 2D fp32 tensor [rows][cols] (cols contiguous), box [box_rows][32 cols], 128-byte swizzle, zero fill
static inline int make_map(CUtensorMap* m, const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                           CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        st3d_set_error("cuTensorMapEncodeTiled is not available from the driver");
        return ST3D_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * sizeof(float)};
    const cuuint32_t box[2] = {32, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    // TFLOAT32: the TMA unit rounds each fp32 value to tf32 (round-to-nearest) while it copies, so the
    // tensor core's operand truncation introduces no systematic bias (plain FLOAT32 would leave the low
    // 13 mantissa bits to be chopped: a -5e-4 relative bias per operand, fatal in G - G_target).
    static const bool plain = [] { const char* e = getenv("ST3D_TMA_PLAIN_FP32"); return e && e[0] == '1'; }();
    const CUresult r = fn(m, plain ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2,
                          const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        st3d_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu)", (int)r,
                       (unsigned long long)rows, (unsigned long long)cols);
        return ST3D_ERR_CUDA;
    }
    return ST3D_OK;
}

// 3D fp32 tensor [B][HW][C] (C contiguous), box [1][box_rows][32 channels]; rows past HW read as zeros
static inline int make_map_nhwc(CUtensorMap* m, const float* base, uint64_t B, uint64_t HW, uint64_t C, uint32_t box_rows,
                                CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        st3d_set_error("cuTensorMapEncodeTiled is not available from the driver");
        return ST3D_ERR_CUDA;
    }
    const cuuint64_t dims[3] = {C, HW, B};
    const cuuint64_t strides[2] = {C * sizeof(float), HW * C * sizeof(float)};
    const cuuint32_t box[3] = {32, box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    static const bool plain = [] { const char* e = getenv("ST3D_TMA_PLAIN_FP32"); return e && e[0] == '1'; }();
    const CUresult r = fn(m, plain ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3,
                          const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        st3d_set_error("cuTensorMapEncodeTiled (NHWC) failed with CUresult %d (B=%llu HW=%llu C=%llu)", (int)r,
                       (unsigned long long)B, (unsigned long long)HW, (unsigned long long)C);
        return ST3D_ERR_CUDA;
    }
    return ST3D_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: remember it per device (bit = device ordinal),
// not per process, so that a second GPU driven from the same process can launch the > 48 KB kernels too
template <typename K>
static int ensure_smem_attr(K kernel, int bytes, std::atomic<uint64_t>& done) {
    int dev = 0;
    ST3D_CUDA_OK(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    if (!(done.load(std::memory_order_acquire) & bit)) {
        ST3D_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        done.fetch_or(bit, std::memory_order_release);
    }
    return ST3D_OK;
}

// SMs of the current device (cached per device ordinal)
static int sm_count() {
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int v = cached[dev & 63].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        cached[dev & 63].store(v, std::memory_order_relaxed);
    }
    return v;
}

template <int C, bool NHWC>
static int launch_fwd(const float* feat, const GramPlan& p, GramEpilogue ep, int* fused_out, cudaStream_t s) {
    using Cfg = FwdCfg<C>;
    CUtensorMap map;
    int rc = NHWC ? make_map_nhwc(&map, feat, p.B, p.HW, C, 32 * Cfg::KPS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
                  : make_map(&map, feat, (uint64_t)p.B * C, (uint64_t)p.HW, Cfg::BOX_ROWS);
    if (rc != ST3D_OK) return rc;
    static std::atomic<uint64_t> attr_done{0};
    rc = ensure_smem_attr(k_gram_tc_fwd<C, NHWC>, (int)Cfg::SMEM, attr_done);
    if (rc != ST3D_OK) return rc;
    // The in-kernel reduction waits for the other CTAs of its image: only when they are certain to be resident --
    // one CTA per SM (the kernel's shared memory allows no more), the whole grid within the SM count (gram_plan sizes
    // it for 148), or no cross-CTA wait at all beyond a hardware-co-scheduled cluster (splits == 1).
    // C = 512 (the 4-CTA cluster) keeps the two-kernel form: a launch that is cooperative AND clustered runs on its own
    // but fails under Nsight Compute's kernel replay (LaunchFailed, CUDA 12.9 / driver 580), and a kernel that cannot be
    // profiled is not worth the 5-10 us the fusion saves on the two smallest layers
    if (Cfg::CLUSTER) ep.fused = 0;
    if (ep.fused) {
        const int64_t grid = (int64_t)Cfg::GROUPS * p.splits * p.B;
        if (p.splits > 1 && grid > sm_count()) ep.fused = 0;
        const uintptr_t bits = (uintptr_t)ep.target | (uintptr_t)ep.gram | (uintptr_t)ep.dgram | (uintptr_t)p.partials;
        if (bits & 15) ep.fused = 0;
    }
    const bool waits = ep.fused && p.splits * Cfg::GROUPS > 1;  // CTAs of an image wait for one another
    if (waits) ST3D_CUDA_OK(cudaMemsetAsync(p.counters, 0, (size_t)p.B * sizeof(unsigned), s));
    float* partials = p.partials;
    int splits = p.splits;
    int64_t k_chunk = p.k_chunk, HW = p.HW;
    for (int attempt = 0; attempt < 2; ++attempt) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(Cfg::GROUPS, p.splits, p.B);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = Cfg::SMEM;
        cfg.stream = s;
        cudaLaunchAttribute attr[2];
        int na = 0;
        if (Cfg::CLUSTER) {
            attr[na].id = cudaLaunchAttributeClusterDimension;
            attr[na].val.clusterDim.x = Cfg::GROUPS;
            attr[na].val.clusterDim.y = 1;
            attr[na].val.clusterDim.z = 1;
            ++na;
        }
        if (waits && ep.fused) {
            // a COOPERATIVE launch: the driver places the whole grid at once or refuses, so the in-kernel wait can
            // neither starve behind another kernel nor deadlock against a second instance of itself on another stream
            attr[na].id = cudaLaunchAttributeCooperative;
            attr[na].val.cooperative = 1;
            ++na;
        }
        cfg.attrs = attr;
        cfg.numAttrs = na;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, k_gram_tc_fwd<C, NHWC>, map, partials, splits, k_chunk, HW, ep);
        if (e == cudaSuccess) break;
        if (attempt == 0 && waits && ep.fused) {  // co-residency refused: two-kernel form (GEMM, then k_gram_finalize)
            (void)cudaGetLastError();
            ep.fused = 0;
            continue;
        }
        st3d_set_error("launch of k_gram_tc_fwd: %s", cudaGetErrorString(e));
        return ST3D_ERR_CUDA;
    }
    *fused_out = ep.fused;
    ST3D_LAUNCH_OK("k_gram_tc_fwd");
    return ST3D_OK;
}

// The B operand of the backward GEMM: S = s (dG + dG^T) from k_gram_symmetrize (scale 1), or -- when the caller vouches
// that dG is symmetric, ST3D_GRAM_DGRAM_SYMMETRIC -- dG itself, with 2 s [* *scale_dev] applied to the accumulator
struct GramS {
    const float* s;
    float scale;
    const float* scale_dev;
};

template <bool NHWC, bool RING>
static int launch_bwd_pair_cfg(const float* feat, const GramPlan& p, GramS S, int accumulate, float* grad_feat, cudaStream_t s) {
    using Cfg = BwdPairCfgT<RING>;
    constexpr int C = Cfg::C;
    CUtensorMap map_f, map_s;
    int rc = NHWC ? make_map_nhwc(&map_f, feat, p.B, p.HW, C, 128, CU_TENSOR_MAP_SWIZZLE_128B)
                  : make_map(&map_f, feat, (uint64_t)p.B * C, (uint64_t)p.HW, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc != ST3D_OK) return rc;
    rc = make_map(&map_s, S.s, (uint64_t)p.B * C, (uint64_t)C, 128);
    if (rc != ST3D_OK) return rc;
    static std::atomic<uint64_t> attr_done{0};
    rc = ensure_smem_attr(k_gram_tc_bwd_pair<NHWC, RING>, (int)Cfg::SMEM, attr_done);
    if (rc != ST3D_OK) return rc;
    const int64_t chunks = (p.HW + 127) / 128, items = (int64_t)p.B * ((chunks + 1) / 2);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * (unsigned)std::min<int64_t>(items, 74));  // one CTA pair per TPC
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int B = p.B;
    int64_t HW = p.HW;
    float out_scale = S.scale;
    const float* out_scale_dev = S.scale_dev;
    ST3D_CUDA_OK(cudaLaunchKernelEx(&cfg, k_gram_tc_bwd_pair<NHWC, RING>, map_f, map_s, feat, grad_feat, B, HW, accumulate,
                                    out_scale, out_scale_dev));
    ST3D_LAUNCH_OK("k_gram_tc_bwd_pair");
    return ST3D_OK;
}

template <bool NHWC>
static int launch_bwd_pair(const float* feat, const GramPlan& p, GramS S, int accumulate, float* grad_feat, cudaStream_t s) {
    // the prefetch-ring configuration when the fused tail is requested on channels_last features (0.090 -> 0.078 ms at
    // 8 x 512 x 4096); ST3D_GRAM_BWD512_NO_RING=1 keeps the four-stage one for A/B timing
    static const bool no_ring = [] { const char* e = getenv("ST3D_GRAM_BWD512_NO_RING"); return e && e[0] == '1'; }();
    // (only for the full tail, accumulate + mask, which has the constant-offset fast path: with the mask alone -- conv5_1,
    // nothing arrives from deeper layers -- the ring configuration measured 0.055 ms against 0.040 ms)
    if (NHWC && accumulate == 3 && !no_ring) return launch_bwd_pair_cfg<NHWC, NHWC>(feat, p, S, accumulate, grad_feat, s);
    return launch_bwd_pair_cfg<NHWC, false>(feat, p, S, accumulate, grad_feat, s);
}

template <int C, bool NHWC, bool RING>
static int launch_bwd_cfg(const float* feat, const GramPlan& p, GramS S, int accumulate, float* grad_feat, cudaStream_t s) {
    using Cfg = BwdCfg<C, RING>;
    CUtensorMap map_f, map_s;
    int rc = NHWC ? make_map_nhwc(&map_f, feat, p.B, p.HW, C, 128, CU_TENSOR_MAP_SWIZZLE_128B)
                  : make_map(&map_f, feat, (uint64_t)p.B * C, (uint64_t)p.HW, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc != ST3D_OK) return rc;
    rc = make_map(&map_s, S.s, (uint64_t)p.B * C, (uint64_t)C, Cfg::S_BOX_ROWS);
    if (rc != ST3D_OK) return rc;
    static std::atomic<uint64_t> attr_done{0};
    rc = ensure_smem_attr(k_gram_tc_bwd<C, NHWC, RING>, (int)Cfg::SMEM, attr_done);
    if (rc != ST3D_OK) return rc;
    const int64_t items = (int64_t)p.B * ((p.HW + 127) / 128);
    ST3D_REQUIRE(items < (1ll << 31), "gram_backward: B * ceil(HW / 128) = %lld work items exceed 2^31", (long long)items);
    const int grid = (int)std::min<int64_t>(items, 148);
    k_gram_tc_bwd<C, NHWC, RING><<<grid, kThreads, Cfg::SMEM, s>>>(map_f, map_s, feat, grad_feat, p.B, p.HW, accumulate, S.scale,
                                                                   S.scale_dev);
    ST3D_LAUNCH_OK("k_gram_tc_bwd");
    return ST3D_OK;
}

template <int C, bool NHWC>
static int launch_bwd(const float* feat, const GramPlan& p, GramS S, int accumulate, float* grad_feat, cudaStream_t s) {
    if (C == 512) {  // CTA pairs (cta_group::2); ST3D_GRAM_BWD512_SINGLE=1 keeps the one-CTA kernel for comparison
        static const bool single = [] { const char* e = getenv("ST3D_GRAM_BWD512_SINGLE"); return e && e[0] == '1'; }();
        if (!single) return launch_bwd_pair<NHWC>(feat, p, S, accumulate, grad_feat, s);
    }
    // C = 256: the prefetch-ring configuration when the fused tail is requested on channels_last features (the only
    // layout the ring serves), the four-stage one otherwise
    if (C == 256 && NHWC && accumulate != 0) return launch_bwd_cfg<C, NHWC, (C <= 256)>(feat, p, S, accumulate, grad_feat, s);
    return launch_bwd_cfg<C, NHWC, (C <= 128)>(feat, p, S, accumulate, grad_feat, s);
}

}  // namespace tc

static inline bool gram_tc_supported(int C, int64_t HW) {
    return (C == 64 || C == 128 || C == 256 || C == 512) && HW % 4 == 0 && HW >= 32;
}

// *fused_out: whether the kernel also reduced the split-K partials and applied `ep` (else k_gram_finalize must run)
static inline int gram_tc_forward(const float* feat, const GramPlan& p, bool nhwc, GramEpilogue ep, int* fused_out,
                                  cudaStream_t s) {
#define ST3D_FWD(CC) \
    return nhwc ? tc::launch_fwd<CC, true>(feat, p, ep, fused_out, s) : tc::launch_fwd<CC, false>(feat, p, ep, fused_out, s)
    switch (p.C) {
        case 64: ST3D_FWD(64);
        case 128: ST3D_FWD(128);
        case 256: ST3D_FWD(256);
        case 512: ST3D_FWD(512);
    }
#undef ST3D_FWD
    return ST3D_ERR_UNSUPPORTED;
}

static inline int gram_tc_backward(const float* feat, const GramPlan& p, bool nhwc, tc::GramS S, int accumulate,
                                   float* grad_feat, cudaStream_t s) {
#define ST3D_BWD(CC)                                                                          \
    return nhwc ? tc::launch_bwd<CC, true>(feat, p, S, accumulate, grad_feat, s)              \
                : tc::launch_bwd<CC, false>(feat, p, S, accumulate, grad_feat, s)
    switch (p.C) {
        case 64: ST3D_BWD(64);
        case 128: ST3D_BWD(128);
        case 256: ST3D_BWD(256);
        case 512: ST3D_BWD(512);
    }
#undef ST3D_BWD
    return ST3D_ERR_UNSUPPORTED;
}

}  // namespace st3d
