// gemm_tcgen05_2cta.cuh — the CTA-pair form of gemm_tcgen05.cuh for the large (prefill) GEMMs.
//
// Two CTAs of a cluster (the two SMs of a TPC) execute ONE tcgen05.mma.cta_group::2 of shape 256 x 256 x 16:
//   * each CTA stages its own 128 rows of A and only HALF of the W tile (128 of the 256 N rows): 32 KB per k-block
//     instead of 48 KB, and each UMMA reads 8 KB of a CTA's shared memory instead of 12 KB — the single-CTA kernel is
//     bounded by shared-memory read bandwidth at ~72 % tensor-pipe activity (profiles/r01_v3_ncu_full_summary.txt);
//   * the accumulator of CTA r (its 128 rows x 256 columns, fp32) lives in CTA r's own TMEM, double buffered;
//   * only the leader CTA (cluster rank 0) issues MMAs; both CTAs run a TMA producer (signalling the LEADER's full
//     barrier: cp.async.bulk.tensor ... .cta_group::2) and their own epilogue warps;
//   * tcgen05.commit ... .multicast::cluster arrives on the same barrier in both CTAs (smem slot free / accumulator ready);
//     the epilogue warps of both CTAs arrive on the leader's tmem-empty barrier (remote mbarrier.arrive for rank 1).
// 6-stage ring (6 x 32 KB).  Plain GEMMs only (one K segment, one accumulator): EPI_QKV / EPI_RESID / EPI_GELU / EPI_RELU.
#pragma once
#include "gemm_tcgen05.cuh"

namespace lmrl {

constexpr int TC2_STAGES = 6;
constexpr int TC2_B_ROWS = TC_BN / 2;                                  // W rows staged per CTA
constexpr int TC2_STAGE_BYTES = TC_A_BYTES + TC2_B_ROWS * TC_BK * 2;   // 16 KB + 16 KB
constexpr int TC2_SMEM_BYTES = 1024 + TC2_STAGES * TC2_STAGE_BYTES + TC_EPI_BYTES + TC_BIAS_BYTES + 256;
static_assert(TC2_SMEM_BYTES <= 232448, "shared memory of one CTA (227 KB)");
constexpr int TC2_THREADS = 64 + 32 * 8;                                // producer, MMA issuer, 8 epilogue warps
// instruction descriptor of the pair: M = 256, N = 256
constexpr uint32_t TC2_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(addr), "r"(rank));
    return out;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory, completing on a barrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_f16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_2cta(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

template <int EPI>
__global__ void __launch_bounds__(TC2_THREADS, 1) gemm_bf16_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                                const __grid_constant__ CUtensorMap tmW,
                                                                                const __grid_constant__ CUtensorMap tmOut, const GemmArgs g) {
    extern __shared__ unsigned char tc2_smem_raw[];
    pdl_trigger();
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t raw = smem_u32(tc2_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* base_ptr = tc2_smem_raw + (base - raw);
    const uint32_t epi_base = base + TC2_STAGES * TC2_STAGE_BYTES;
    const uint32_t bias_base = epi_base + TC_EPI_BYTES;                  // the current tile's 256 bias values (staged once per tile, as in the one-CTA kernel)
    const uint32_t bar_base = bias_base + TC_BIAS_BYTES;
    // barriers (8 bytes each): full[6], empty[6], tmem_full[2], tmem_empty[2], then the TMEM base address word
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (TC2_STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * TC2_STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * TC2_STAGES + 2 + a); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + TC2_STAGES * TC2_STAGE_BYTES + TC_EPI_BYTES + TC_BIAS_BYTES + 8 * (2 * TC2_STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kt = g.K / TC_BK;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC2_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 16); }   // 8 epilogue warps x 2 CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();                                                  // barriers of both CTAs are initialised, TMEM is allocated
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                                          // first access to the predecessor's global memory is below
    const int M = *g.M_dev;
    const int n_tiles = g.N / TC_BN, m_tiles = (M + 2 * TC_BM - 1) / (2 * TC_BM), total = M > 0 ? n_tiles * m_tiles : 0;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): my 128 rows of A, my 128 rows of the W tile; bytes complete on the LEADER's barrier =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = cluster_id; tile < total; tile += n_clusters) {
                const int m0 = (tile / n_tiles) * 2 * TC_BM + (int)rank * TC_BM, n0 = (tile % n_tiles) * TC_BN + (int)rank * TC2_B_ROWS;
                for (int kb = 0; kb < kt; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);             // multicast commit: this slot is free in BOTH CTAs
                    if (leader) mbar_expect_tx(full_bar(stage), 2 * TC2_STAGE_BYTES);
                    const uint32_t sA = base + stage * TC2_STAGE_BYTES, sB = sA + TC_A_BYTES;
                    const uint32_t fb = mapa_shared(full_bar(stage), 0);
                    tma_load_2d_2cta(sA, &tmA, fb, kb * TC_BK, m0);
                    tma_load_2d_2cta(sB, &tmW, fb, kb * TC_BK, n0);
                    if (++stage == TC2_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the leader CTA drives both SMs' tensor cores =====
        if (leader && lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = cluster_id; tile < total; tile += n_clusters) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);              // the epilogues of both CTAs have drained this accumulator
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_BN);
                for (int kb = 0; kb < kt; ++kb) {
                    mbar_wait(full_bar(stage), phase);                   // both CTAs' bytes have landed
                    tcgen05_fence_after();
                    const uint32_t sA = base + stage * TC2_STAGE_BYTES, sB = sA + TC_A_BYTES;
                    const uint64_t da = umma_desc_sw128(sA), db = umma_desc_sw128(sB);
#pragma unroll
                    for (int k = 0; k < TC_BK / TC_UMMA_K; ++k)
                        tcgen05_mma_f16_2cta(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), TC2_IDESC, (kb | k) ? 1u : 0u);
                    tcgen05_commit_2cta(empty_bar(stage));               // frees the slot in both CTAs when these MMAs retire
                    if (++stage == TC2_STAGES) { stage = 0; phase ^= 1u; }
                }
                tcgen05_commit_2cta(tfull_bar(acc));                     // accumulators of both CTAs complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ===== epilogue warps 2..9 of both CTAs: my CTA's 128 rows x 256 columns, two warps per lane quadrant =====
        const int quad = warp & 3;
        const uint32_t stg = epi_base + (uint32_t)(warp - 2) * TC_EPI_BUF;
        const int half = (warp - 2) >> 2;
        const int et = (int)threadIdx.x - 64;                                     // thread index inside the epilogue group (256 threads)
        const uint32_t sbias = g.bias ? bias_base : 0u;
        auto epi_bar = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };      // named barrier of the eight epilogue warps
        float bias_next = 0.f;                                                     // this thread's bias value of the cluster's next tile
        if (sbias && cluster_id < total) bias_next = __ldg(g.bias + (cluster_id % n_tiles) * TC_BN + et);
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = cluster_id; tile < total; tile += n_clusters) {
            const int m0 = (tile / n_tiles) * 2 * TC_BM + (int)rank * TC_BM, n0 = (tile % n_tiles) * TC_BN;
            if (sbias) {
                epi_bar();                                                         // nobody still reads the previous tile's values
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_base + 4u * (uint32_t)et), "f"(bias_next) : "memory");
                epi_bar();
                const int nt_b = tile + n_clusters;
                if (nt_b < total) bias_next = __ldg(g.bias + (nt_b % n_tiles) * TC_BN + et);
            }
            mbar_wait(tfull_bar(acc), acc_phase);
            tcgen05_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * TC_BN);
            tc_store_tile<EPI>(g, tmOut, t_addr, M, m0, n0, quad, lane, stg, half * (TC_BN / 2), sbias);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (lane == 0) bulk_wait_all();
    }
    tcgen05_fence_before();
    cluster_sync_all();                                                  // the peer may still be reading my shared memory / TMEM
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

}  // namespace lmrl
