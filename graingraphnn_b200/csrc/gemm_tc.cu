// gemm_tc.cu — kernel family (c) on the 5th-generation tensor cores: tcgen05.mma (kind::tf32) with TMEM accumulators,
// operands staged in shared memory by TMA (cp.async.bulk.tensor, 128-byte swizzle), warp-specialised persistent CTAs.
//
//   gg_split_tf32   : [X | h]  ->  A_hi, A_lo   (a = hi + lo, both representable in TF32; K padded to 32-float chunks)
//   gg_node_proj_tc : P = A W^T + b  as  A_lo W_hi^T + A_hi W_lo^T + A_hi W_hi^T  ("3xTF32": fp32-grade accuracy,
//                     error ~2^-22 relative per product, needed for the 1e-4 per-step parity bar that plain TF32 misses)
//
// Per CTA (one per SM): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warp 2 = TMEM allocator,
// warps 4-7 = epilogue (tcgen05.ld -> +bias -> 128-bit global stores).  Tile = 128 rows x 256 columns, K streamed in
// 32-float (128-byte) chunks through two mbarrier rings (A tiles, W tiles: every operand tile is fetched once per chunk); two 256-column TMEM accumulators double-buffer the
// epilogue against the next tile's MMAs.  All three split terms accumulate into the same TMEM tile, smallest first.
//
// Roofline: tensor pipe. Algorithmic flops = 2 M N K; the kernel issues 3x that in TF32 MMAs.  The fp32 output
// (N*4 bytes per row) makes it co-limited by HBM writes: 128 KB per tile vs 6144 MMA cycles.
#include "common.cuh"
#include <mutex>
#include <cuda.h>
#include <stdlib.h>
#include <stdio.h>

namespace {

constexpr int BM = 128;          // rows per tile  (UMMA M)
constexpr int BN = 256;          // columns per tile (max UMMA N)
constexpr int BK = 32;           // floats per K chunk = 128 bytes = one swizzle-128B row
constexpr int A_BYTES = BM * BK * 4;     // 16 KB
constexpr int B_BYTES = BN * BK * 4;     // 32 KB
constexpr int kRingA = 4, kRingB = 4;    // node projection: separate rings of A (16 KB) and W (32 KB) chunk tiles
constexpr int kOutBufs = 2;              // node projection: store staging buffers (the TMA store of a chunk takes ~1 us to drain)
constexpr int OUT_CHUNK = 32;            // columns per epilogue chunk = one 128-byte swizzled row of the store box
constexpr int OUT_BYTES = BM * OUT_CHUNK * 4;            // 16 KB staging per TMA store, double-buffered
constexpr int SMEM_BYTES = kRingA * A_BYTES + kRingB * B_BYTES + kOutBufs * OUT_BYTES + 1024 /*align*/ + 256 /*barriers*/;
static_assert(SMEM_BYTES <= 227 * 1024, "node projection shared memory");
constexpr int kThreads = 256;
constexpr uint32_t TMEM_COLS = 512;

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Plain spin.  A spin counter with a clock-based trap (as the gather's waits have) was tried in round 2: the gate GEMM, whose single
// MMA-issuing thread sits on these waits, became 35 % slower (1.61 -> 2.19 ms per step), so the GEMM waits stay unguarded; a lost
// TMA copy or commit here would hang the launch (the host-side argument checks are what guards them).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// One lane of a converged warp.  The MMA / TMA issuing warps run their loops with all 32 lanes (warp-uniform control flow), so the
// compiler keeps TMEM addresses, shared-memory descriptors and barrier addresses in uniform registers; issued from inside
// `if (lane == 0)` every operand sits in a vector register and ptxas wraps each tcgen05.mma in an ELECT / R2UR.BROADCAST x3 /
// BRA.U.ANY loop (~60 clk per MMA: at N = 96, 48 tensor clk per MMA, the issuing thread and not the tensor pipe set the pace).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled shared-memory operand descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1) |
//   [32,46) stride byte offset >> 4 = 1024 B between 8-row groups | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=TF32 [7,10)=2, b=TF32 [10,13)=2, K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ the GEMM kernel
__global__ void __launch_bounds__(kThreads, 1)
node_proj_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                    const __grid_constant__ CUtensorMap tmOut, const float* __restrict__ bias, int M, int N, int Kp, int k_first_steps) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;           // swizzle-128B tiles need 1024-B alignment
    // Operand rings: the three split terms of a K chunk are issued as (A_lo, W_hi), (A_hi, W_hi), (A_hi, W_lo), so every operand
    // tile is fetched ONCE per chunk (A_lo, W_hi, A_hi, W_lo: 96 KB instead of 3 x 48 KB) — the kernel is bound by the ~60 B/clk an
    // SM pulls from L2 through TMA, not by the tensor pipe.  A tiles and W tiles live in separate rings with their own barriers.
    const uint32_t ringA = smem_base, ringB = smem_base + kRingA * A_BYTES;
    const uint32_t out_base = ringB + kRingB * B_BYTES;                          // 2 x 16 KB store staging (1024-B aligned)
    const uint32_t bar_base = out_base + kOutBufs * OUT_BYTES;
    auto fullA = [&](int s) { return bar_base + 8u * s; };
    auto emptyA = [&](int s) { return bar_base + 8u * (kRingA + s); };
    auto fullB = [&](int s) { return bar_base + 8u * (2 * kRingA + s); };
    auto emptyB = [&](int s) { return bar_base + 8u * (2 * kRingA + kRingB + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kRingA + 2 * kRingB + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kRingA + 2 * kRingB + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kRingA + 2 * kRingB + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blocks = (M + BM - 1) / BM, n_blocks = (N + BN - 1) / BN;
    const int n_tiles = m_blocks * n_blocks;
    const int chunks = Kp / BK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kRingA; ++s) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), 1); }
        for (int s = 0; s < kRingB; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            auto load_a = [&](const CUtensorMap* map, int k0, int m0) {
                mbar_wait(emptyA(sa), pa ^ 1u);
                mbar_expect_tx(fullA(sa), A_BYTES);
                tma_load_2d(ringA + sa * A_BYTES, map, fullA(sa), k0, m0);
                if (++sa == kRingA) { sa = 0; pa ^= 1u; }
            };
            auto load_b = [&](const CUtensorMap* map, int k0, int n0) {
                mbar_wait(emptyB(sb), pb ^ 1u);
                mbar_expect_tx(fullB(sb), B_BYTES);
                tma_load_2d(ringB + sb * B_BYTES, map, fullB(sb), k0, n0);
                if (++sb == kRingB) { sb = 0; pb ^= 1u; }
            };
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int m0 = (t / n_blocks) * BM, n0 = (t % n_blocks) * BN;
                for (int c = 0; c < chunks; ++c) {
                    const int k0 = c * BK;
                    load_a(&tmA_lo, k0, m0);
                    load_b(&tmW_hi, k0, n0);
                    load_a(&tmA_hi, k0, m0);
                    load_b(&tmW_lo, k0, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (single thread) =====
        if (lane == 0) {
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int n0 = (t % n_blocks) * BN;
                int ncols = N - n0; ncols = ncols > BN ? BN : ncols; ncols = (ncols + 15) & ~15;
                const uint32_t idesc = make_idesc(BM, ncols);
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);          // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                for (int c = 0; c < chunks; ++c) {
                    // UMMA K = 8 tf32 = 32 bytes: advance the start address by 2 (x16 B).  The first chunk holds the node features
                    // zero-padded to 32 columns: only its k_first_steps leading K steps can contribute.
                    const int ks = c == 0 ? k_first_steps : BK / 8;
                    auto group = [&](uint32_t a_addr, uint32_t b_addr, bool first) {
                        const uint64_t adesc = make_desc(a_addr), bdesc = make_desc(b_addr);
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k)
                            if (k < ks) umma_tf32(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (first && k == 0) ? 0u : 1u);
                    };
                    // (A_lo, W_hi)
                    const int a_lo = sa; mbar_wait(fullA(sa), pa); if (++sa == kRingA) { sa = 0; pa ^= 1u; }
                    const int b_hi = sb; mbar_wait(fullB(sb), pb); if (++sb == kRingB) { sb = 0; pb ^= 1u; }
                    tc_fence_after();
                    group(ringA + a_lo * A_BYTES, ringB + b_hi * B_BYTES, c == 0);
                    umma_commit(emptyA(a_lo));                        // A_lo is free when these MMAs retire
                    // (A_hi, W_hi)
                    const int a_hi = sa; mbar_wait(fullA(sa), pa); if (++sa == kRingA) { sa = 0; pa ^= 1u; }
                    tc_fence_after();
                    group(ringA + a_hi * A_BYTES, ringB + b_hi * B_BYTES, false);
                    umma_commit(emptyB(b_hi));
                    // (A_hi, W_lo)
                    const int b_lo = sb; mbar_wait(fullB(sb), pb); if (++sb == kRingB) { sb = 0; pb ^= 1u; }
                    tc_fence_after();
                    group(ringA + a_hi * A_BYTES, ringB + b_lo * B_BYTES, false);
                    umma_commit(emptyA(a_hi));
                    umma_commit(emptyB(b_lo));
                }
                umma_commit(tfull_bar(acc));                          // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> (+bias) -> swizzled smem staging -> TMA store =====
        // Each thread owns one accumulator row; a direct global store would scatter 32 x 16 B per instruction
        // (measured: ~1.5 TB/s chip-wide), so 32-column chunks are staged in shared memory in the 128-byte-swizzle
        // layout of the store tensor map and written with one cp.async.bulk.tensor per chunk (OOB rows/cols clipped).
        const int q = warp & 3;                                       // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;                                  // row inside the tile
        int acc = 0; uint32_t acc_phase = 0;
        int obuf = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int m0 = (t / n_blocks) * BM, n0 = (t % n_blocks) * BN;
            int ncols = N - n0; ncols = ncols > BN ? BN : ncols;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            for (int c0 = 0; c0 < ncols; c0 += OUT_CHUNK) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), v);
                // the staging buffer we are about to overwrite was read by the TMA store issued kOutBufs chunks ago
                if (threadIdx.x == 128) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kOutBufs - 1) : "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const uint32_t srow = out_base + obuf * OUT_BYTES + (uint32_t)r * 128u;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                           __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    if (bias && c0 + 4 * j < ncols) {
                        const float4 b = ldg4(bias + n0 + c0 + 4 * j);
                        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                    }
                    const uint32_t dst = srow + (uint32_t)((j ^ (r & 7)) << 4);       // 128-byte swizzle: chunk ^= row % 8
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes -> visible to TMA
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (threadIdx.x == 128) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(&tmOut), "r"(out_base + obuf * OUT_BYTES), "r"(n0 + c0), "r"(m0) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (++obuf == kOutBufs) obuf = 0;
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (threadIdx.x == 128) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ projection, fused split
// gg_node_proj_fused: the same GEMM without the split pass and without A_hi / A_lo in HBM.  TMA lands the fp32 chunk of
// [X (zero-extended to 32 columns by the TMA box) | h] in the first half of an A slot; the tensor core reads a TF32 operand's
// upper 19 bits only, so that half IS A_hi (= trunc_tf32(a)).  Four converter warps (thread = row) write
// a - trunc_tf32(a) (exact in fp32) into the second half of the slot, in the same 128-byte-swizzled layout.  Term order
// (A_hi, W_hi), (A_hi, W_lo), (A_lo, W_hi): the converters work while the first two groups are issued.  Per chunk and tile
// 16 + 64 KB come through TMA instead of 32 + 64 KB, and gg_split_tf32 (one read of [X | h], two writes of 32 + C floats per
// node) is gone.
constexpr int kFusedThreads = 384;       // warps 0-3 TMA / MMA / TMEM / idle, 4-7 epilogue, 8-11 converters
#ifndef GG_PROJ_RING_A
#define GG_PROJ_RING_A 2
#endif
constexpr int kFRingA = GG_PROJ_RING_A;  // A slots of [fp32 chunk | lo chunk] (2 x 16 KB)
#ifndef GG_PROJ_RING_B
#define GG_PROJ_RING_B 4
#endif
#ifndef GG_PROJ_OUT_BUFS
#define GG_PROJ_OUT_BUFS 2
#endif
constexpr int kFRingB = GG_PROJ_RING_B;  // W tiles (32 KB): W_hi and W_lo of two chunks
constexpr int kFOutBufs = GG_PROJ_OUT_BUFS;      // store staging buffers (16 KB each)
constexpr int FUSED_SMEM_BYTES = kFRingA * 2 * A_BYTES + kFRingB * B_BYTES + kFOutBufs * OUT_BYTES + 1024 + 256;
static_assert(FUSED_SMEM_BYTES <= 227 * 1024, "fused projection shared memory");

__global__ void __launch_bounds__(kFusedThreads, 1)
node_proj_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmH,
                       const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                       const __grid_constant__ CUtensorMap tmOut, const float* __restrict__ bias, int M, int N, int chunks, int k_first_steps) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t ringA = smem_base, ringB = smem_base + kFRingA * 2 * A_BYTES;
    const uint32_t out_base = ringB + kFRingB * B_BYTES;
    const uint32_t bar_base = out_base + kFOutBufs * OUT_BYTES;
    auto fullA = [&](int s) { return bar_base + 8u * s; };
    auto convA = [&](int s) { return bar_base + 8u * (kFRingA + s); };
    auto emptyA = [&](int s) { return bar_base + 8u * (2 * kFRingA + s); };
    auto fullB = [&](int s) { return bar_base + 8u * (3 * kFRingA + s); };
    auto emptyB = [&](int s) { return bar_base + 8u * (3 * kFRingA + kFRingB + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (3 * kFRingA + 2 * kFRingB + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (3 * kFRingA + 2 * kFRingB + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (3 * kFRingA + 2 * kFRingB + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blocks = (M + BM - 1) / BM, n_blocks = (N + BN - 1) / BN;
    const int n_tiles = m_blocks * n_blocks;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kFRingA; ++s) { mbar_init(fullA(s), 1); mbar_init(convA(s), 4); mbar_init(emptyA(s), 1); }
        for (int s = 0; s < kFRingB; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmH) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===== TMA producer: per chunk the fp32 A tile, then W_hi, then W_lo (warp-uniform loops, one lane issues) =====
        {
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            auto load_b = [&](const CUtensorMap* map, int k0, int n0) {
                mbar_wait(emptyB(sb), pb ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(fullB(sb), B_BYTES);
                    tma_load_2d(ringB + sb * B_BYTES, map, fullB(sb), k0, n0);
                }
                if (++sb == kFRingB) { sb = 0; pb ^= 1u; }
            };
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int m0 = (t / n_blocks) * BM, n0 = (t % n_blocks) * BN;
                for (int c = 0; c < chunks; ++c) {
                    mbar_wait(emptyA(sa), pa ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(fullA(sa), A_BYTES);
                        if (c == 0) tma_load_2d(ringA + sa * 2 * A_BYTES, &tmX, fullA(sa), 0, m0);
                        else tma_load_2d(ringA + sa * 2 * A_BYTES, &tmH, fullA(sa), (c - 1) * BK, m0);
                    }
                    if (++sa == kFRingA) { sa = 0; pa ^= 1u; }
                    load_b(&tmW_hi, c * BK, n0);
                    load_b(&tmW_lo, c * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: warp-uniform loops, one elected lane issues (see elect_one) =====
        {
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int n0 = (t % n_blocks) * BN;
                int ncols = N - n0; ncols = ncols > BN ? BN : ncols; ncols = (ncols + 15) & ~15;
                const uint32_t idesc = make_idesc(BM, ncols);
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                for (int c = 0; c < chunks; ++c) {
                    const int ks = c == 0 ? k_first_steps : BK / 8;      // MMAs of the zero padding behind the features are not issued
                    auto group = [&](uint32_t a_addr, uint32_t b_addr, bool first) {
                        const uint64_t adesc = make_desc(a_addr), bdesc = make_desc(b_addr);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BK / 8; ++k)
                                if (k < ks) umma_tf32(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (first && k == 0) ? 0u : 1u);
                        }
                    };
                    const int a_s = sa; const uint32_t a_par = pa;
                    if (++sa == kFRingA) { sa = 0; pa ^= 1u; }
                    const uint32_t a_hi = ringA + a_s * 2 * A_BYTES, a_lo = a_hi + A_BYTES;
                    // (A_hi, W_hi)
                    mbar_wait(fullA(a_s), a_par);
                    const int b_hi = sb; mbar_wait(fullB(sb), pb); if (++sb == kFRingB) { sb = 0; pb ^= 1u; }
                    tc_fence_after();
                    group(a_hi, ringB + b_hi * B_BYTES, c == 0);
                    // (A_hi, W_lo)
                    const int b_lo = sb; mbar_wait(fullB(sb), pb); if (++sb == kFRingB) { sb = 0; pb ^= 1u; }
                    tc_fence_after();
                    group(a_hi, ringB + b_lo * B_BYTES, false);
                    if (elect_one()) umma_commit(emptyB(b_lo));
                    // (A_lo, W_hi): the converters have written the lo half by now
                    mbar_wait(convA(a_s), a_par);
                    tc_fence_after();
                    group(a_lo, ringB + b_hi * B_BYTES, false);
                    if (elect_one()) { umma_commit(emptyA(a_s)); umma_commit(emptyB(b_hi)); }
                }
                if (elect_one()) umma_commit(tfull_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp >= 8) {
        // ===== converters: lo = a - trunc_tf32(a), row r of the chunk, same swizzled layout, second half of the slot =====
        const int r = threadIdx.x - 256;
        int sa = 0; uint32_t pa = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            for (int c = 0; c < chunks; ++c) {
                mbar_wait(fullA(sa), pa);
                const uint32_t row = ringA + sa * 2 * A_BYTES + (uint32_t)r * 128u;
#pragma unroll
                for (int j = 0; j < 8; ++j) {                    // 16-byte chunk j of the row sits at (j ^ (r % 8)) * 16: same place in both halves
                    uint32_t a0, a1, a2, a3;
                    const uint32_t off = (uint32_t)((j ^ (r & 7)) << 4);
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(row + off));
                    // the residual has 13 significant bits; rounding it to TF32 here (rna) keeps the hardware's truncation from biasing it
                    auto lo_of = [](uint32_t a) -> uint32_t {
                        uint32_t u;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(__uint_as_float(a) - __uint_as_float(a & 0xFFFFE000u)));
                        return u;
                    };
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + A_BYTES + off), "r"(lo_of(a0)), "r"(lo_of(a1)), "r"(lo_of(a2)), "r"(lo_of(a3)) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(convA(sa));
                if (++sa == kFRingA) { sa = 0; pa ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> (+bias) -> swizzled smem staging -> TMA store (as in node_proj_tc_kernel) =====
        // (Tried in round 2: tcgen05.ld.16x256b, which hands four lanes 32 contiguous bytes of a row, and 8-byte global stores
        // straight from the registers - eight whole sectors per warp store, no staging, no barriers.  Correct, and twice as slow:
        // 1.2 ms against 0.60 ms for the decoder projection of 2.5e5 joints.  More staging buffers (4 x 16 KB with a 3-slot W
        // ring) changed nothing: profiles/r2_gemm_ablations.txt.)
        const int q = warp & 3;
        const int r = q * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        int obuf = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int m0 = (t / n_blocks) * BM, n0 = (t % n_blocks) * BN;
            int ncols = N - n0; ncols = ncols > BN ? BN : ncols;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            for (int c0 = 0; c0 < ncols; c0 += OUT_CHUNK) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), v);
                if (threadIdx.x == 128) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kFOutBufs - 1) : "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const uint32_t srow = out_base + obuf * OUT_BYTES + (uint32_t)r * 128u;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                           __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    if (bias && c0 + 4 * j < ncols) {
                        const float4 b = ldg4(bias + n0 + c0 + 4 * j);
                        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                    }
                    const uint32_t dst = srow + (uint32_t)((j ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (threadIdx.x == 128) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(&tmOut), "r"(out_base + obuf * OUT_BYTES), "r"(n0 + c0), "r"(m0) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (++obuf == kFOutBufs) obuf = 0;
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (threadIdx.x == 128) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ gate GEMM + LSTM
// gg_gate_update_tc: per 128-node tile and gate g, D_g[128, C] (TMEM columns [g*C, (g+1)*C)) accumulates
//   sum over K chunks of  A_chunk (agg of every incoming edge type, then X, then h)  x  Wall[g*C + n, k]   (3xTF32),
// then the epilogue adds the rank-1 terms (lin_edge, lin_l2 bias, gate bias) and applies the LSTM update.
//
// All A operands arrive as plain fp32 (the aggregate written by gg_pgat_gather, the node features, the hidden state):
// the TF32 hi/lo split happens INSIDE the kernel and the split operand lives in TENSOR MEMORY.  TMA lands the fp32
// [128 x 32] chunk in a ring slot; eight converter warps (thread = half a row) read it and tcgen05.st  hi = the fp32 word
// itself (the tensor core reads only the upper 19 bits of a TF32 operand) and lo = a - trunc_tf32(a) into the 64 TMEM
// columns of that stage; ONE stage then feeds 12 MMAs in the A-from-TMEM form (tcgen05.mma [d], [a_tmem], b_desc):
// (lo, W_hi), (hi, W_lo), (hi, W_hi) x 4 k-steps.  The rank-1 terms of the pre-activation (lin_edge x ea, lin_l2 bias x
// [deg > 0], biases) are injected by the converters into the unused columns of the feature chunk, so they are added by the
// tensor core too and the epilogue sees finished pre-activations.
// Why A in TMEM: with 4-byte operands and N = C = 96 an SS-form MMA reads 7 KB of shared memory per 48 tensor cycles, more
// than the 128 B/clk an SM delivers (measured: the SS version scaled with the SM count at ~20 B/clk/SM of TMA traffic, a
// per-SM limit).  With A in TMEM the tensor core only fetches W from shared memory, the converters write no shared memory,
// and no agg_lo / A_hi / A_lo tensors exist in HBM.  Measured limits now (scripts/kernel_bench.py, clock64 traces): the
// single MMA-issuing thread (~60 clk per N = 96 MMA + ~400 clk of wait/fence/commit per stage vs 576 tensor clk).
//
// Warp roles (512 threads): 0 TMA producer | 1 MMA issuer | 2 TMEM allocator | 4-7 epilogue | 8-15 hi/lo converters
// (two warps per TMEM lane quarter, 16 of the chunk's 32 columns each).  setmaxnreg moves registers to the epilogue warps,
// which keep one row of the LSTM state (C values) in registers across the four gate passes.
constexpr int G_A_BYTES = BM * BK * 4;                  // 16 KB: fp32 chunk, rewritten in place as its TF32 hi part
constexpr int kGateThreads = 512;
constexpr int kMaxIn = 2;

struct GateMaps {
    CUtensorMap agg[kMaxIn];        // fp32 [M, G*C]
    CUtensorMap x, h;               // fp32 [M, K1] (box 32 columns, zero-filled beyond K1) and [M, C]
    CUtensorMap w_hi, w_lo;         // Wall [G*C, Ktot]
    CUtensorMap out_h, out_c;       // stores: [M, C] (LSTM modes) or [M, G*C] (RAW / RELU), box 128 x 32
};
struct GateEpi {
    const float* ea[kMaxIn]; const int* rowptr[kMaxIn]; int weighted[kMaxIn];
    const float* c_in; const float* out_c;
    int n_in, M, G, C, has_h, mode, stages;

};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
// A-from-TMEM form: A = [128 lanes x 8 columns] of 32-bit TF32 values at a_tmem (row m in lane m, k in column k)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 ::"r"(tmem_d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}

// sigmoid / tanh through ex2.approx + rcp.approx (relative error ~2^-22 each; saturate correctly at +-inf)
__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sigmoid(float x) { return fast_rcp(1.0f + fast_ex2(-1.4426950408889634f * x)); }
__device__ __forceinline__ float fast_tanh(float x) { return fmaf(-2.0f, fast_rcp(1.0f + fast_ex2(2.8853900817779268f * x)), 1.0f); }

// Gates are processed one per PASS; pass p accumulates in TMEM buffer p & 1 (C columns each), so the epilogue of one gate
// overlaps the MMAs of the next and the rest of tensor memory (512 - 2C columns) holds a deep ring of split A chunks.
// LSTM (i,f,c,o) runs i, c~, f, o:  u = s(i);  u *= tanh(c~);  u = s(f) c + u (= c');  h' = s(o) tanh(u).
// LSTM0 (i,c~,o) runs i, c~, o with c = 0.  u lives in the epilogue thread's registers (one row, C values).
template <int G, int MODE> struct GatePasses {
    static constexpr bool lstm = MODE == GG_GATE_LSTM, lstm0 = MODE == GG_GATE_LSTM0;
    static constexpr int n = G;
    __host__ __device__ static constexpr int gate(int p) { return lstm ? (p == 1 ? 2 : (p == 2 ? 1 : p)) : p; }
};

template <int G, int MODE, int NV>
__global__ void __launch_bounds__(kGateThreads, 1)
gate_update_tc_kernel(const __grid_constant__ GateMaps maps, const GateEpi ep) {
    using PS = GatePasses<G, MODE>;
    static_assert(G == 1 || MODE == GG_GATE_LSTM || MODE == GG_GATE_LSTM0, "RAW / RELU run one gate");
    extern __shared__ uint8_t smem_raw[];
    constexpr int C = 32 * NV, CC = NV;
    const int S = ep.stages;
    const uint32_t w_bytes = (uint32_t)C * BK * 4;                        // one W tile (hi or lo)
    const uint32_t slot_bytes = G_A_BYTES + 2 * w_bytes;                  // [A fp32 -> hi | W_hi | W_lo], 1024-B multiples
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stage_base = smem_base + S * slot_bytes;               // epilogue staging: one [128 x 32] output chunk (16 KB)
    const uint32_t bar_base = stage_base + OUT_BYTES;
    constexpr uint32_t a_col0 = 2u * C;                                   // split-A ring behind the two accumulator buffers
    // Split-A ring slot s belongs to shared-memory stage s (S <= (512 - 2C) / 64): full[s] can only fire after the producer saw
    // empty[s], i.e. after the MMAs that read TMEM slot s retired, so the converters need no barrier of their own.
    // barriers: full[S] | conv[S] | empty[S] | tfull[2] | tempty[2]
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto conv_bar = [&](int s) { return bar_base + 8u * (S + s); };
    auto empty_bar = [&](int s) { return bar_base + 8u * (2 * S + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (3 * S + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (3 * S + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (3 * S + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_in = ep.n_in;
    const int n_tiles = (ep.M + BM - 1) / BM;
    const int nck = n_in * CC + 1 + (ep.has_h ? CC : 0);                 // K chunks per gate
    const uint32_t stage_tx = (uint32_t)G_A_BYTES + 2u * w_bytes;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(conv_bar(s), 8); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 0) {
        // ===== TMA producer: one fp32 A chunk + the matching W_hi / W_lo tiles per stage (warp-uniform loops, one lane issues) =====
        {
            int stage = 0; uint32_t phase = 0;
            auto a_chunk = [&](int g, int ck, const CUtensorMap*& am, int& acol) {
                if (ck < n_in * CC) { am = &maps.agg[ck / CC]; acol = g * C + (ck % CC) * BK; }
                else if (ck == n_in * CC) { am = &maps.x; acol = 0; }
                else { am = &maps.h; acol = (ck - n_in * CC - 1) * BK; }
            };
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int m0 = t * BM;
#pragma unroll
                for (int p = 0; p < PS::n; ++p) {
                    {
                        const int g = PS::gate(p);
                        for (int ck = 0; ck < nck; ++ck) {
                            const CUtensorMap* am; int acol;
                            a_chunk(g, ck, am, acol);
                            mbar_wait(empty_bar(stage), phase ^ 1u);
                            const uint32_t sa = smem_base + stage * slot_bytes;
                            if (elect_one()) {
                                mbar_expect_tx(full_bar(stage), stage_tx);
                                tma_load_2d(sa, am, full_bar(stage), acol, m0);
                                tma_load_2d(sa + G_A_BYTES, &maps.w_hi, full_bar(stage), ck * BK, g * C);
                                tma_load_2d(sa + G_A_BYTES + w_bytes, &maps.w_lo, full_bar(stage), ck * BK, g * C);
                            }
                            if (++stage == S) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: warp-uniform loops, one elected lane issues (see elect_one) =====
        {
            int stage = 0; uint32_t phase = 0;
            int buf = 0; uint32_t bphase = 0;                           // bit b = phase of TMEM buffer b
            const uint32_t idesc = make_idesc(BM, C);
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
#pragma unroll
                for (int p = 0; p < PS::n; ++p) {
                    mbar_wait(tempty_bar(buf), ((bphase >> buf) & 1u) ^ 1u);      // epilogue has drained this buffer
                    tc_fence_after();
                    {
                        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * C);
                        for (int ck = 0; ck < nck; ++ck) {
                            mbar_wait(conv_bar(stage), phase);          // TMA landed AND the chunk is split into hi / lo
                            tc_fence_after();
                            const uint32_t sa = smem_base + stage * slot_bytes;
                            const uint32_t ah = tmem_base + a_col0 + (uint32_t)(stage * 64), al = ah + 32u;
                            const uint64_t whdesc = make_desc(sa + G_A_BYTES), wldesc = make_desc(sa + G_A_BYTES + w_bytes);
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < BK / 8; ++k) umma_tf32_ts(tmem_d, al + 8u * k, whdesc + 2u * k, idesc, (ck | k) ? 1u : 0u);
#pragma unroll
                                for (int k = 0; k < BK / 8; ++k) umma_tf32_ts(tmem_d, ah + 8u * k, wldesc + 2u * k, idesc, 1u);
#pragma unroll
                                for (int k = 0; k < BK / 8; ++k) umma_tf32_ts(tmem_d, ah + 8u * k, whdesc + 2u * k, idesc, 1u);
                                umma_commit(empty_bar(stage));      // frees the smem slot AND the TMEM ring slot
                            }
                            if (++stage == S) { stage = 0; phase ^= 1u; }
                        }
                    }
                    if (elect_one()) umma_commit(tfull_bar(buf));
                    bphase ^= 1u << buf;
                    buf ^= 1;
                }
            }
        }
    } else if (warp >= 8) {
        // ===== converters: thread = half a row (16 columns) of the fp32 chunk -> TF32 hi / lo -> tensor memory (tcgen05.st).
        // hi is the fp32 word itself: the tensor core reads only the upper 19 bits of a TF32 operand, i.e. hi = trunc_tf32(a);
        // lo = a - trunc_tf32(a) is exact in fp32 (13 significant bits) and is in turn truncated by the hardware.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
        const int q = warp & 3, r = q * 32 + lane, hf = (warp - 8) >> 2;
        int stage = 0; uint32_t phase = 0;
        constexpr int RB = 32 - (2 * G + 3);                   // rank-1 columns of the X chunk: [ea0[G] | cnt0 | ea1[G] | cnt1 | 1]
        static_assert(RB >= 16, "the rank-1 columns must sit in the upper half of the feature chunk");
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            // per-row inputs of the rank-1 terms (lin_edge, lin_l2 bias x [deg > 0], biases): they ride in the unused columns
            // of the feature chunk, so the tensor core adds them with 3xTF32 accuracy and the epilogue only sees pre-activations
            const int m = t * BM + r;
            float ea0[G], ea1[G], cnt0 = 0.f, cnt1 = 0.f;
#pragma unroll
            for (int g = 0; g < G; ++g) { ea0[g] = 0.f; ea1[g] = 0.f; }
            if (hf == 1 && m < ep.M) {
                const int d0 = __ldg(&ep.rowptr[0][m + 1]) - __ldg(&ep.rowptr[0][m]);
                cnt0 = ep.weighted[0] ? (d0 > 0 ? 1.f : 0.f) : (float)d0;
#pragma unroll
                for (int g = 0; g < G; ++g) ea0[g] = __ldg(&ep.ea[0][(size_t)m * G + g]);
                if (n_in > 1) {
                    const int d1 = __ldg(&ep.rowptr[1][m + 1]) - __ldg(&ep.rowptr[1][m]);
                    cnt1 = ep.weighted[1] ? (d1 > 0 ? 1.f : 0.f) : (float)d1;
#pragma unroll
                    for (int g = 0; g < G; ++g) ea1[g] = __ldg(&ep.ea[1][(size_t)m * G + g]);
                }
            }
            for (int it = 0; it < G * nck; ++it) {
                const bool xchunk = hf == 1 && (it % nck) == n_in * CC;
                mbar_wait(full_bar(stage), phase);
                const uint32_t srow = smem_base + stage * slot_bytes + (uint32_t)r * 128u;
                uint32_t a[16], lw[16];
#pragma unroll
                for (int c = 0; c < 4; ++c)                    // logical 16-byte chunk 4*hf + c sits at (4*hf + c) ^ (row % 8)
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a[4 * c]), "=r"(a[4 * c + 1]), "=r"(a[4 * c + 2]), "=r"(a[4 * c + 3])
                                 : "r"(srow + (uint32_t)(((4 * hf + c) ^ (r & 7)) << 4)));
                if (xchunk) {
#pragma unroll
                    for (int g = 0; g < G; ++g) { a[RB - 16 + g] = __float_as_uint(ea0[g]); a[RB - 16 + G + 1 + g] = __float_as_uint(ea1[g]); }
                    a[RB - 16 + G] = __float_as_uint(cnt0); a[RB - 16 + 2 * G + 1] = __float_as_uint(cnt1); a[15] = 0x3F800000u;
                }
#pragma unroll
                for (int c = 0; c < 16; ++c) lw[c] = __float_as_uint(__uint_as_float(a[c]) - __uint_as_float(a[c] & 0xFFFFE000u));
                tc_fence_after();
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + a_col0 + (uint32_t)(stage * 64 + 16 * hf);
                tmem_st16(ta, a);
                tmem_st16(ta + 32u, lw);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(conv_bar(stage));
                if (++stage == S) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: one accumulator row per thread; outputs leave through 128-byte-swizzled shared-memory staging and
        // TMA stores (a thread-per-row global store scatters 16 B over 32 lines per instruction: measured LSU-bound at
        // ~16 B/clk/SM, 3x the MMA time of a tile).  u = s(i) tanh(c~) of pass 0 stays in registers until pass 1. =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int q = warp & 3, r = q * 32 + lane;
        int buf = 0; uint32_t bphase = 0;
        const float* __restrict__ c_in = ep.c_in;
        const bool want_c = ep.out_c != nullptr;
        const int M = ep.M;
        const uint32_t stage_row = stage_base + (uint32_t)r * 128u;
        const bool leader = threadIdx.x == 128;
        constexpr bool kLstmAny = PS::lstm || PS::lstm0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int m0 = t * BM, m = m0 + r;
            const bool ok = m < M;
            float u[kLstmAny ? C : 1];
#pragma unroll
            for (int p = 0; p < PS::n; ++p) {
                mbar_wait(tfull_bar(buf), (bphase >> buf) & 1u);
                tc_fence_after();
                const uint32_t tbuf = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * C);
                // what this pass writes: LSTM  p=2 -> c' (if wanted), p=3 -> h';  LSTM0  p=1 -> c', p=2 -> h';  RAW/RELU -> out
                const bool is_c = (PS::lstm && p == 2) || (PS::lstm0 && p == 1);
                const bool is_h = !kLstmAny || p == PS::n - 1;
                const bool stores = ((is_c && want_c) || is_h);
#pragma unroll
                for (int c = 0; c < NV; ++c) {                         // 32-column chunk = one swizzled staging row
                    if (stores) {
                        if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging is free again
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                    }
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int c0 = 32 * c + 16 * half;
                        float pre[16], o[16];
                        {
                            uint32_t v[16];
                            tmem_ld16(tbuf + (uint32_t)c0, v);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int i = 0; i < 16; ++i) pre[i] = __uint_as_float(v[i]);
                        }
                        if (!kLstmAny) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[i] = MODE == GG_GATE_RELU ? fmaxf(pre[i], 0.f) : pre[i];
                        } else if (p == 0) {                            // input gate
#pragma unroll
                            for (int i = 0; i < 16; ++i) u[c0 + i] = fast_sigmoid(pre[i]);
                        } else if (p == 1) {                            // candidate: u = s(i) tanh(c~)  (= c' when c = 0)
#pragma unroll
                            for (int i = 0; i < 16; ++i) { u[c0 + i] *= fast_tanh(pre[i]); o[i] = u[c0 + i]; }
                        } else if (PS::lstm && p == 2) {                // forget gate: c' = s(f) c + u
                            if (c_in && ok) {
#pragma unroll
                                for (int i = 0; i < 16; i += 4) {
                                    const float4 c4 = ldg4(c_in + (size_t)m * C + c0 + i);
                                    u[c0 + i] = fmaf(fast_sigmoid(pre[i]), c4.x, u[c0 + i]);
                                    u[c0 + i + 1] = fmaf(fast_sigmoid(pre[i + 1]), c4.y, u[c0 + i + 1]);
                                    u[c0 + i + 2] = fmaf(fast_sigmoid(pre[i + 2]), c4.z, u[c0 + i + 2]);
                                    u[c0 + i + 3] = fmaf(fast_sigmoid(pre[i + 3]), c4.w, u[c0 + i + 3]);
                                }
                            }
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[i] = u[c0 + i];
                        } else {                                        // output gate: h' = s(o) tanh(c')
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[i] = fast_sigmoid(pre[i]) * fast_tanh(u[c0 + i]);
                        }
                        if (stores) {
#pragma unroll
                            for (int i = 0; i < 4; ++i)                 // 16-byte chunk (4*half + i) of the row, 128-byte swizzle
                                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stage_row + (uint32_t)(((4 * half + i) ^ (r & 7)) << 4)),
                                             "f"(o[4 * i]), "f"(o[4 * i + 1]), "f"(o[4 * i + 2]), "f"(o[4 * i + 3]) : "memory");
                        }
                    }
                    if (stores) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                        if (leader) {
                            // RAW / RELU output is [M, G*C] (G = 1 here); the LSTM outputs are [M, C]
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                         ::"l"(is_h ? &maps.out_h : &maps.out_c), "r"(stage_base), "r"(32 * c), "r"(m0) : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(tempty_bar(buf));
                bphase ^= 1u << buf;
                buf ^= 1;
            }
        }
        if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ split kernel
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// one float4 of the padded row per thread: columns [0,K1p32) from X (zero beyond K1), then K2 columns of H
__global__ void split_tf32_kernel(const float* __restrict__ X, int ldx, int K1, const float* __restrict__ H, int ldh, int K2,
                                  int M, float* __restrict__ Ahi, float* __restrict__ Alo, int Kp, int K1p32) {
    const int q4 = Kp >> 2;
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)M * q4) return;
    const int m = (int)(t / q4), c = (int)(t % q4) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < K1p32) {
        if (c < K1) v = ldg4(X + (size_t)m * ldx + c);            // K1 % 4 == 0 (padded feature rows)
    } else if (H && c - K1p32 < K2) {
        v = ldg4(H + (size_t)m * ldh + (c - K1p32));
    }
    float4 hi = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
    float4 lo = make_float4(to_tf32(v.x - hi.x), to_tf32(v.y - hi.y), to_tf32(v.z - hi.z), to_tf32(v.w - hi.w));
    *reinterpret_cast<float4*>(Ahi + (size_t)m * Kp + c) = hi;
    *reinterpret_cast<float4*>(Alo + (size_t)m * Kp + c) = lo;
}

// ------------------------------------------------------------------------------------------------ host helpers
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp32 row-major [rows, cols] with row stride ld (floats); box = [box_rows x 32 floats], 128-byte swizzle
// Descriptors are cached per (pointer, shape): the engine's buffers are resident, so after the first step every launch finds its 5-9
// maps here instead of calling the driver (paid on every step of the eager multi-GPU path; free under graph replay either way).
struct MapKey { const void* base; int64_t rows, cols, ld; int box_rows; };
struct MapSlot { MapKey key; CUtensorMap map; bool used; };
constexpr int kMapSlots = 256;
MapSlot g_map_cache[kMapSlots];
std::mutex g_map_mutex;

int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    if ((ld & 3) || !gg_aligned16(base)) return GG_EALIGN;
    const uint64_t h = ((uint64_t)(uintptr_t)base >> 4) * 0x9e3779b97f4a7c15ull ^ (uint64_t)rows * 0xff51afd7ed558ccdull ^ (uint64_t)cols * 0xc4ceb9fe1a85ec53ull ^
                       (uint64_t)ld * 31 ^ (uint64_t)box_rows;
    MapSlot& slot = g_map_cache[(h >> 17) % kMapSlots];
    {
        std::lock_guard<std::mutex> lock(g_map_mutex);
        if (slot.used && slot.key.base == base && slot.key.rows == rows && slot.key.cols == cols && slot.key.ld == ld && slot.key.box_rows == box_rows) {
            *map = slot.map;
            return 0;
        }
    }
    EncodeTiledFn enc = get_encode();
    if (!enc) return GG_EARCH;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return GG_EINVAL;
    std::lock_guard<std::mutex> lock(g_map_mutex);
    slot.key = MapKey{base, rows, cols, ld, box_rows};
    slot.map = *map;
    slot.used = true;
    return 0;
}

}  // namespace

extern "C" int gg_tc_supported(void) { return gg_device_is_sm100() && get_encode() != nullptr ? 1 : 0; }

extern "C" int gg_split_tf32(const float* X, int32_t ldx, int32_t K1, const float* H, int32_t ldh, int32_t K2,
                             int32_t M, float* A_hi, float* A_lo, int32_t Kp, int32_t K1p32, void* stream) {
    if (M < 0 || K1 < 0 || K2 < 0 || Kp <= 0 || (Kp % BK) || (K1p32 % BK) || K1 > K1p32 || K1p32 + (H ? K2 : 0) > Kp) return GG_EINVAL;
    if (M == 0) return 0;
    if (!X || !A_hi || !A_lo) return GG_EINVAL;
    if ((K1 & 3) || (K2 & 3) || (ldx & 3) || (H && (ldh & 3))) return GG_EALIGN;
    if (!gg_aligned16(X) || (H && !gg_aligned16(H)) || !gg_aligned16(A_hi) || !gg_aligned16(A_lo)) return GG_EALIGN;
    const int64_t n = (int64_t)M * (Kp / 4);
    split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, GG_STREAM(stream)>>>(X, ldx, K1, H, ldh, K2, M, A_hi, A_lo, Kp, K1p32);
    GG_LAUNCH_OK();
    return 0;
}

extern "C" int gg_node_proj_tc(const float* A_hi, const float* A_lo, int32_t Kp, int32_t k_first, const float* W_hi, const float* W_lo,
                               int32_t N, const float* bias, float* out, int32_t ldo, int32_t M, int32_t n_sms, void* stream) {
    if (M < 0 || N < 0 || Kp <= 0 || (Kp % BK) || k_first < 1 || k_first > BK) return GG_EINVAL;
    if (M == 0 || N == 0) return 0;
    if (!A_hi || !A_lo || !W_hi || !W_lo || !out) return GG_EINVAL;
    if ((N & 3) || (ldo & 3) || !gg_aligned16(out) || (bias && !gg_aligned16(bias))) return GG_EALIGN;
    if (!gg_aligned16(A_hi) || !gg_aligned16(A_lo) || !gg_aligned16(W_hi) || !gg_aligned16(W_lo)) return GG_EALIGN;
    CUtensorMap mA_hi, mA_lo, mW_hi, mW_lo;
    int rc;
    if ((rc = make_map(&mA_hi, A_hi, M, Kp, Kp, BM))) return rc;
    if ((rc = make_map(&mA_lo, A_lo, M, Kp, Kp, BM))) return rc;
    if ((rc = make_map(&mW_hi, W_hi, N, Kp, Kp, BN))) return rc;
    if ((rc = make_map(&mW_lo, W_lo, N, Kp, Kp, BN))) return rc;
    CUtensorMap mOut;
    if ((rc = make_map(&mOut, out, M, N, ldo, BM))) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(node_proj_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    if (n_sms <= 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int grid = tiles < n_sms ? tiles : n_sms;
    node_proj_tc_kernel<<<grid, kThreads, SMEM_BYTES, GG_STREAM(stream)>>>(mA_hi, mA_lo, mW_hi, mW_lo, mOut, bias, M, N, Kp, (k_first + 7) / 8);
    GG_LAUNCH_OK();
    return 0;
}

// Projection with the TF32 split fused in: out[M, N] = [X (K1 <= 32 columns, zero-extended to 32) | H (K2 columns)] W^T + bias,
// W_hi / W_lo [N, 32 + K2] as for gg_node_proj_tc.  No A_hi / A_lo arrays, no gg_split_tf32 pass.
extern "C" int gg_node_proj_fused(const float* X, int32_t ldx, int32_t K1, const float* H, int32_t ldh, int32_t K2,
                                  const float* W_hi, const float* W_lo, int32_t N, const float* bias, float* out, int32_t ldo,
                                  int32_t M, int32_t n_sms, void* stream) {
    if (M < 0 || N < 0 || K1 < 4 || K1 > BK || (K1 & 3) || K2 < 0 || (K2 % BK)) return GG_EINVAL;
    if (M == 0 || N == 0) return 0;
    if (!X || !W_hi || !W_lo || !out || (K2 > 0 && !H)) return GG_EINVAL;
    if ((N & 3) || (ldo & 3) || !gg_aligned16(out) || (bias && !gg_aligned16(bias))) return GG_EALIGN;
    const int Kp = BK + K2;
    CUtensorMap mX, mH, mW_hi, mW_lo, mOut;
    int rc;
    if ((rc = make_map(&mX, X, M, K1, ldx, BM))) return rc;
    if (K2 > 0) { if ((rc = make_map(&mH, H, M, K2, ldh, BM))) return rc; } else mH = mX;
    if ((rc = make_map(&mW_hi, W_hi, N, Kp, Kp, BN))) return rc;
    if ((rc = make_map(&mW_lo, W_lo, N, Kp, Kp, BN))) return rc;
    if ((rc = make_map(&mOut, out, M, N, ldo, BM))) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(node_proj_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    if (n_sms <= 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int grid = tiles < n_sms ? tiles : n_sms;
    node_proj_fused_kernel<<<grid, kFusedThreads, FUSED_SMEM_BYTES, GG_STREAM(stream)>>>(mX, mH, mW_hi, mW_lo, mOut, bias, M, N, Kp / BK, (K1 + 7) / 8);
    GG_LAUNCH_OK();
    return 0;
}

// Tensor-core variant of gg_gate_update.  All A operands are plain fp32 and split into TF32 hi / lo inside the kernel:
// inputs[e].agg [M, G*C] (gg_pgat_gather), X [M, K1] (zero-extended to 32 columns by the TMA box), H [M, C] or NULL.
// W_hi/W_lo: Wall [G*C, Ktot], Ktot = n_inputs*C + 32 (+ C with H), K layout
//   [W2 of input 0 | W2 of input 1 | feature chunk (32) | skip(h)]; the feature chunk of gate g holds the summed lin_skip
//   columns of X in [0, K1) and, from RB = 32 - (2G + 3): We_0[g] at RB + g (other gates' slots 0), b2_0[g] at RB + G,
//   We_1[g] at RB + G + 1 + g, b2_1[g] at RB + 2G + 1, btot[g] at 31 - the kernel feeds [ea_0[0..G) | cnt_0 | ea_1[0..G) | cnt_1 | 1]
//   in those columns (packing.tc_gate_weight_layout builds it).
extern "C" int gg_gate_update_tc(const gg_agg_input* inputs, int32_t n_inputs,
                                 const float* X, int32_t ldx, int32_t K1, const float* H, int32_t ldh,
                                 const float* W_hi, const float* W_lo, int32_t Ktot,
                                 const float* c_in, float* out_h, float* out_c,
                                 int32_t M, int32_t G, int32_t C, int32_t mode, int32_t n_sms, void* stream) {
    if (M < 0 || G < 1 || G > 4 || C % 32 || C < 32 || C > 128 || n_inputs < 1 || n_inputs > kMaxIn) return GG_EINVAL;
    if (mode < GG_GATE_RAW || mode > GG_GATE_LSTM0) return GG_EINVAL;
    if ((mode == GG_GATE_LSTM && G != 4) || (mode == GG_GATE_LSTM0 && G != 3) || (mode == GG_GATE_RELU && G != 1)) return GG_EINVAL;
    if (mode == GG_GATE_RAW && G != 1) return GG_EINVAL;      // multi-gate RAW output: use gg_gate_update (fp32 SIMT)
    if (K1 < 4 || K1 > 32 - (2 * G + 3) || (K1 & 3) || Ktot != n_inputs * C + 32 + (H ? C : 0)) return GG_EINVAL;
    if (M == 0) return 0;
    if (!inputs || !X || !W_hi || !W_lo || !out_h) return GG_EINVAL;
    if (!gg_aligned16(out_h) || (out_c && !gg_aligned16(out_c)) || (c_in && !gg_aligned16(c_in))) return GG_EALIGN;
    GateMaps maps;
    GateEpi ep;
    int rc;
    for (int e = 0; e < n_inputs; ++e) {
        const gg_agg_input& s = inputs[e];
        if (!s.agg || !s.ea || !s.rowptr) return GG_EINVAL;
        if ((rc = make_map(&maps.agg[e], s.agg, M, (int64_t)G * C, s.ld_agg, BM))) return rc;
        ep.ea[e] = s.ea; ep.rowptr[e] = s.rowptr; ep.weighted[e] = s.weighted;
    }
    for (int e = n_inputs; e < kMaxIn; ++e) {
        maps.agg[e] = maps.agg[0];
        ep.ea[e] = nullptr; ep.rowptr[e] = nullptr; ep.weighted[e] = 0;
    }
    if ((rc = make_map(&maps.x, X, M, K1, ldx, BM))) return rc;
    if (H) { if ((rc = make_map(&maps.h, H, M, C, ldh, BM))) return rc; } else maps.h = maps.x;
    if ((rc = make_map(&maps.w_hi, W_hi, (int64_t)G * C, Ktot, Ktot, C))) return rc;
    if ((rc = make_map(&maps.w_lo, W_lo, (int64_t)G * C, Ktot, Ktot, C))) return rc;
    const bool lstm_any = mode == GG_GATE_LSTM || mode == GG_GATE_LSTM0;
    if ((rc = make_map(&maps.out_h, out_h, M, lstm_any ? C : G * C, lstm_any ? C : G * C, BM))) return rc;
    if (out_c) { if ((rc = make_map(&maps.out_c, out_c, M, C, C, BM))) return rc; } else maps.out_c = maps.out_h;
    ep.c_in = c_in; ep.out_c = lstm_any ? out_c : nullptr;
    ep.n_in = n_inputs; ep.M = M; ep.G = G; ep.C = C; ep.has_h = H ? 1 : 0; ep.mode = mode;
    if (n_sms <= 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    // ring depth: as many [A | W_hi | W_lo] slots as fit beside the epilogue vectors and the barriers
    const int slot_bytes = G_A_BYTES + 2 * C * BK * 4;
    const int fixed = 1024 + OUT_BYTES + 512;
    int stages = (227 * 1024 - fixed) / slot_bytes;
    if (stages > ((int)TMEM_COLS - 2 * C) / 64) stages = ((int)TMEM_COLS - 2 * C) / 64;   // TMEM: 2 accumulators + 64 columns per stage
    if (stages < 2) return GG_EINVAL;
    ep.stages = stages;
    const int smem_bytes = fixed + stages * slot_bytes;
    const int tiles = (M + BM - 1) / BM;
    const int grid = tiles < n_sms ? tiles : n_sms;
    cudaStream_t st = GG_STREAM(stream);
    cudaError_t err = cudaSuccess;
#define GG_LAUNCH_GATE(GV, MV, NVV)                                                                                              \
    do {                                                                                                                         \
        err = cudaFuncSetAttribute(gate_update_tc_kernel<GV, MV, NVV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
        if (err != cudaSuccess) return (int)err;                                                                                 \
        gate_update_tc_kernel<GV, MV, NVV><<<grid, kGateThreads, smem_bytes, st>>>(maps, ep);                                    \
    } while (0)
#define GG_LAUNCH_GATE_NV(GV, MV)                                                        \
    switch (C / 32) {                                                                    \
        case 1: GG_LAUNCH_GATE(GV, MV, 1); break;                                        \
        case 2: GG_LAUNCH_GATE(GV, MV, 2); break;                                        \
        case 3: GG_LAUNCH_GATE(GV, MV, 3); break;                                        \
        default: GG_LAUNCH_GATE(GV, MV, 4); break;                                       \
    }
    if (mode == GG_GATE_LSTM) GG_LAUNCH_GATE_NV(4, GG_GATE_LSTM)
    else if (mode == GG_GATE_LSTM0) GG_LAUNCH_GATE_NV(3, GG_GATE_LSTM0)
    else if (mode == GG_GATE_RELU) GG_LAUNCH_GATE_NV(1, GG_GATE_RELU)
    else GG_LAUNCH_GATE_NV(1, GG_GATE_RAW)
#undef GG_LAUNCH_GATE_NV
    GG_LAUNCH_OK();
    return 0;
}
