// gemm_tcgen05.cuh — C[M,N] = A[M,K] · W[N,K]^T on the 5th-generation tensor cores (sm_100a only).
//
//   * operands staged global -> shared by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) into a 4-stage ring,
//   * tcgen05.mma.cta_group::1.kind::f16 (UMMA 128 x 256 x 16, bf16 in / fp32 accumulate) issued by ONE thread,
//   * accumulators in TMEM, double buffered (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1,
//   * warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue (one TMEM lane
//     quadrant each), synchronised only by mbarriers (full/empty per stage, full/empty per accumulator),
//   * persistent CTAs, one per SM, striding over 128x256 output tiles; the live row count M is read from device memory.
//
// Same epilogues as gemm_mma.cuh (EPI_QKV / EPI_RESID / EPI_GELU / EPI_SAMPLE).  For the three storing epilogues each
// epilogue warp transposes its 32x32 fp32 chunk through a padded shared-memory tile so that global traffic is
// row-contiguous (8 lanes x 16 B per row); the sampling epilogue works straight from registers, one row per thread.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "gemm_mma.cuh"

namespace lmrl {

constexpr int TC_BM = 128, TC_BN = 256, TC_BK = 64, TC_STAGES = 4, TC_UMMA_K = 16;
constexpr int TC_THREADS = 192;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;          // 16 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 2;          // 32 KB
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr int TC_EPI_PITCH = 36;                       // floats per staged row (32 + 4 pad: 16-byte aligned, conflict free)
constexpr int TC_EPI_BYTES = 4 * 32 * TC_EPI_PITCH * 4;
constexpr int TC_SMEM_BYTES = 1024 /*align slack*/ + TC_STAGES * TC_STAGE_BYTES + TC_EPI_BYTES + 256 /*barriers*/;
constexpr uint32_t TC_TMEM_COLS = 512;

// kind::f16 instruction descriptor: fp32 accumulate, bf16 A and B, both K-major, N = 256, M = 128
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // bounded wait (2 s of wall clock): a protocol bug traps with a context error instead of hanging the GPU
    uint32_t done = 0;
    uint64_t t0 = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && (spin & 1023u) == 1023u) {
            const uint64_t now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) {
                printf("lmrl gemm_tcgen05: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
                __trap();
            }
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 B apart, descriptor version 1
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t r[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// row-contiguous store of one transposed 32x32 chunk: this lane owns 4 consecutive columns of row `row`
template <int EPI>
__device__ __forceinline__ void epi_store4(const GemmArgs& g, int row, int col, float4 v) {
    if (g.bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + col));
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if constexpr (EPI == EPI_RESID) {
        float4* p = reinterpret_cast<float4*>(g.resid + (size_t)row * g.N + col);
        float4 r = *p;
        r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w;
        *p = r;
    } else {
        __nv_bfloat162 lo, hi;
        if constexpr (EPI == EPI_GELU) { lo = __floats2bfloat162_rn(gelu_new(v.x), gelu_new(v.y)); hi = __floats2bfloat162_rn(gelu_new(v.z), gelu_new(v.w)); }
        else { lo = __floats2bfloat162_rn(v.x, v.y); hi = __floats2bfloat162_rn(v.z, v.w); }
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&lo); pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        if constexpr (EPI == EPI_GELU) {
            *reinterpret_cast<uint2*>(g.out + (size_t)row * g.N + col) = pk;
        } else {   // EPI_QKV
            const int d = g.d_model;
            const int part = col / d, c = col - part * d;
            if (part == 0) *reinterpret_cast<uint2*>(g.out + (size_t)row * d + c) = pk;
            else {
                const int e = g.row_env[row], p = g.row_pos[row];
                __nv_bfloat16* cache = part == 1 ? g.kcache : g.vcache;
                *reinterpret_cast<uint2*>(cache + (((size_t)e * g.n_head + (c >> 6)) * g.t_cap + p) * 64 + (c & 63)) = pk;
            }
        }
    }
}

template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                         const __grid_constant__ CUtensorMap tmW, const GemmArgs g) {
    extern __shared__ unsigned char tc_smem_raw[];
    const int M = *g.M_dev;
    if (M <= 0) return;
    const uint32_t raw = smem_u32(tc_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                        // 128B swizzle atoms need 1024-byte alignment
    unsigned char* base_ptr = tc_smem_raw + (base - raw);
    const uint32_t epi_base = base + TC_STAGES * TC_STAGE_BYTES;
    float* epi_ptr = reinterpret_cast<float*>(base_ptr + TC_STAGES * TC_STAGE_BYTES);
    const uint32_t bar_base = epi_base + TC_EPI_BYTES;
    // barrier layout (8 bytes each): full[4], empty[4], tmem_full[2], tmem_empty[2], then the TMEM base address word
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (TC_STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * TC_STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * TC_STAGES + 2 + a); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + TC_STAGES * TC_STAGE_BYTES + TC_EPI_BYTES + 8 * (2 * TC_STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = g.N / TC_BN, m_tiles = (M + TC_BM - 1) / TC_BM, total = n_tiles * m_tiles, kt = g.K / TC_BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int m0 = (tile / n_tiles) * TC_BM, n0 = (tile % n_tiles) * TC_BN;
                for (int kb = 0; kb < kt; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    mbar_expect_tx(full_bar(stage), TC_STAGE_BYTES);
                    const uint32_t sA = base + stage * TC_STAGE_BYTES, sB = sA + TC_A_BYTES;
                    tma_load_2d(sA, &tmA, full_bar(stage), kb * TC_BK, m0);
                    tma_load_2d(sB, &tmW, full_bar(stage), kb * TC_BK, n0);
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);               // epilogue has drained this accumulator
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_BN);
                for (int kb = 0; kb < kt; ++kb) {
                    mbar_wait(full_bar(stage), phase);                    // TMA bytes have landed
                    tcgen05_fence_after();
                    const uint32_t sA = base + stage * TC_STAGE_BYTES, sB = sA + TC_A_BYTES;
                    const uint64_t da = umma_desc_sw128(sA), db = umma_desc_sw128(sB);
#pragma unroll
                    for (int k = 0; k < TC_BK / TC_UMMA_K; ++k)           // +32 bytes per UMMA_K step inside the swizzle atom
                        tcgen05_mma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), TC_IDESC, (kb | k) ? 1u : 0u);
                    tcgen05_commit(empty_bar(stage));                     // frees the smem slot when these MMAs retire
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
                }
                tcgen05_commit(tfull_bar(acc));                           // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ===== epilogue warps 2..5: TMEM lane quadrant = warp % 4 =====
        const int quad = warp & 3;
        float* stg = epi_ptr + (warp - 2) * 32 * TC_EPI_PITCH;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            const int m0 = (tile / n_tiles) * TC_BM, n0 = (tile % n_tiles) * TC_BN;
            mbar_wait(tfull_bar(acc), acc_phase);
            tcgen05_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * TC_BN);
            if constexpr (EPI == EPI_SAMPLE) {
                const int row = m0 + quad * 32 + lane;
                const bool live = row < M;
                const uint64_t key = (live && g.do_sample) ? g.row_key[row] : 0ull;
                float best = -INFINITY;
                int best_col = -1;
                for (int c0 = 0; c0 < TC_BN; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(t_addr + c0, r);
                    if (live) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int col = n0 + c0 + j;
                            if (col < g.n_vocab) {
                                const float logit = __uint_as_float(r[j]);
                                if (g.logits_out) g.logits_out[(size_t)row * g.n_vocab + col] = logit;
                                float s = logit * g.inv_temp;
                                if (g.do_sample) s += gumbel_from_bits(noise_bits(key, (uint32_t)col));
                                if (s > best) { best = s; best_col = col; }
                            }
                        }
                    }
                }
                if (live && best_col >= 0) atomicMax(g.best + row, pack_best(best, best_col));
            } else {
                for (int c0 = 0; c0 < TC_BN; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(t_addr + c0, r);
                    // lane = row of the quadrant: park the 32 columns, then re-read with lane = (row group, column quad)
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(stg + lane * TC_EPI_PITCH + 4 * j) =
                            make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                    __syncwarp();
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rr = it * 4 + (lane >> 3), cq = lane & 7;
                        const float4 v = *reinterpret_cast<const float4*>(stg + rr * TC_EPI_PITCH + 4 * cq);
                        const int row = m0 + quad * 32 + rr;
                        if (row < M) epi_store4<EPI>(g, row, n0 + c0 + 4 * cq, v);
                    }
                    __syncwarp();
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

}  // namespace lmrl
