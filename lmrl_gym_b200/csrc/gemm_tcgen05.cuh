// gemm_tcgen05.cuh — C[M,N] = A[M,K] · W[N,K]^T on the 5th-generation tensor cores (sm_100a only).
//
//   * operands staged global -> shared by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) into a 4-stage ring,
//   * tcgen05.mma.cta_group::1.kind::f16 (UMMA 128 x 256 x 16, bf16 in / fp32 accumulate) issued by ONE thread,
//   * accumulators in TMEM, double buffered (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1,
//   * warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue (one TMEM lane
//     quadrant each), synchronised only by mbarriers (full/empty per stage, full/empty per accumulator),
//   * persistent CTAs, one per SM, striding over 128x256 output tiles; the live row count M is read from device memory.
//
// The mainloop is a list of K segments (GemmArgs::seg): each segment streams k-blocks of A against one of up to three
// W matrices into one of up to two TMEM accumulators, which is how the ILQL / MC value-perturbed sampler
// (logits = pi_beta + beta * min(q1, q2)) runs as ONE kernel: acc0 = [h_pi | b*r1]·[wte | W2_1]^T, acc1 = [h_pi | b*r2]·[wte | W2_2]^T.
//
// Same epilogues as gemm_mma.cuh (EPI_QKV / EPI_RESID / EPI_GELU / EPI_SAMPLE) plus EPI_RELU.  Dense outputs leave through TMA:
// each epilogue warp parks a 32-row x 128-byte box in swizzled shared memory and one lane issues
//   * cp.reduce.async.bulk.tensor ... .add  (fp32 residual stream: x += tile + bias, no read of x by the SM),
//   * cp.async.bulk.tensor store            (bf16 GeLU output, bf16 Q),
// double buffered per warp.  K / V tiles of the QKV GEMM are scattered to the KV slab straight from registers
// (64 contiguous bytes per thread per chunk).  The sampling epilogue works from registers, one row per thread.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "gemm_mma.cuh"

namespace lmrl {

constexpr int TC_BM = 128, TC_BN = 256, TC_BK = 64, TC_STAGES = 4, TC_UMMA_K = 16;
// warps: 0 = TMA producer, 1 = MMA issuer, 2..9 = epilogue.  Eight epilogue warps: warps w and w+4 share a TMEM lane quadrant and
// split the tile's columns.  With K = 768 a tile's mainloop is only ~6 k cycles, and four warps draining 128 x 256 fp32 (bias,
// GeLU / Gumbel hash, shared-memory staging) took about as long — tensor activity of the K = 768 GEMMs was epilogue-bound.
template <int EPI> struct TcCfg {
    static constexpr int EPI_WARPS = 8;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
};
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;          // 16 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 2;          // 32 KB
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr int TC_EPI_BUF = 32 * 128;                   // one staged box: 32 rows x 128 bytes (128B-swizzled, TMA store source)
constexpr int TC_EPI_BYTES = 8 * TC_EPI_BUF;           // one staging box per epilogue warp
constexpr int TC_BIAS_BYTES = 256 * 4;                 // the tile's bias values, staged once per tile by the epilogue warps
constexpr int TC_SMEM_BYTES = 1024 /*align slack*/ + TC_STAGES * TC_STAGE_BYTES + TC_EPI_BYTES + TC_BIAS_BYTES + 256 /*barriers*/;
constexpr uint32_t TC_TMEM_COLS = 512;
// per tile width BN (256 by default; 192 for the N = 768 GEMMs of a decode pass: 4 x 32 = 128 tiles fill 148 SMs in one wave of
// 3/4-size tiles where 256-wide tiles give 96 full-size ones)
template <int BN> struct TcTile {
    static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "tile width: a multiple of 64 up to 256 (2 x BN TMEM columns)");
    static constexpr int B_BYTES = BN * TC_BK * 2;
    static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
    static constexpr int SMEM_BYTES = 1024 /*align slack*/ + TC_STAGES * STAGE_BYTES + TC_EPI_BYTES + TC_BIAS_BYTES + 256 /*barriers*/;
    static_assert(SMEM_BYTES <= 232448, "shared memory of one CTA (227 KB)");
    // kind::f16 instruction descriptor: fp32 accumulate, bf16 A and B, both K-major, N = BN, M = 128
    static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
};

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
// wall-clock bound of one mbarrier wait; LMRL_MBAR_TIMEOUT_S (read at lmrl_create) raises it for tools that slow kernels down by orders of
// magnitude (compute-sanitizer racecheck: a tile's MMAs can take seconds)
__device__ unsigned long long g_mbar_timeout_ns = 2000000000ull;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // bounded wait (2 s of wall clock by default): a protocol bug traps with a context error instead of hanging the GPU
    uint32_t done = 0;
    uint64_t t0 = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && (spin & 1023u) == 1023u) {
            const uint64_t now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > g_mbar_timeout_ns) {
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
// issue only: the destination registers are valid after the next tmem_ld_wait()
__device__ __forceinline__ void tmem_ld_32x32b_x32_issue(uint32_t taddr, uint32_t r[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t r[32]) {
    tmem_ld_32x32b_x32_issue(taddr, r);
    tmem_ld_wait();
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2_tc(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 32 contiguous bytes per thread (sm_100: 256-bit global store): one full sector per lane
__device__ __forceinline__ void st_global_v8(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f, uint32_t g, uint32_t h) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- storing epilogues of one 128 x BN accumulator (shared by the 1-CTA and the 2-CTA kernel) ------------------------
// t_addr: TMEM address of this warp's lane quadrant; m0 / n0: origin of the CTA's tile; c_beg: first of this warp's BN / 2 columns
// of the tile; stg: this warp's staging box (32 rows x 128 bytes, 128B-swizzled).  The TMEM loads are software pipelined: the
// tcgen05.ld of 32-column chunk c + 1 is in flight while chunk c gets its bias / activation / packing (round 1 issued the wait
// right behind every load: 30 % of the kernel's stall samples).  The staging box is single buffered: the warp waits for the TMA
// engine to have read the previous box only when it is about to refill it.  sbias != 0: shared-memory address of the tile's BN
// bias values (staged once per tile by the kernel); sbias == 0: bias read from global memory (the 2-CTA kernel).
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 tc_bias4(const GemmArgs& g, uint32_t sbias, int n0, int c) {      // bias of tile columns c .. c + 3
    if (sbias) return ld_shared_f4(sbias + 4u * (uint32_t)c);
    return __ldg(reinterpret_cast<const float4*>(g.bias + n0 + c));
}

// XB residual epilogue: coalesced read of a warp's 32-row x 32-column fp32 box of x (lane -> row 4i + lane / 8, 16-byte chunk lane % 8)
__device__ __forceinline__ void tc_xb_load_box(const GemmArgs& g, int M, int brow0, int col, int lane, float4 xin[8]) {
    const int lr = lane >> 3, lc = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = brow0 + 4 * i + lr;
        xin[i] = r < M ? __ldcg(reinterpret_cast<const float4*>(g.resid + (size_t)r * g.N + col) + lc) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <int EPI, int BN = TC_BN, bool XB = false>
__device__ __forceinline__ void tc_store_tile(const GemmArgs& g, const CUtensorMap& tmOut, uint32_t t_addr, int M, int m0, int n0, int quad,
                                              int lane, uint32_t stg, int c_beg, uint32_t sbias, float rs = 1.0f, float4* xin = nullptr, int nm0 = -1, int nn0 = 0) {
    // xin (XB only): the first two chunks' boxes of x, loaded BEFORE the accumulator wait (by the caller, or by the previous tile's call when
    // nm0 / nn0 = origin of this CTA's next tile was passed: with an even chunk count the two-deep x prefetch runs on across tiles)
    // rs: per-row scale of the accumulator (1 / sigma of the row when the LayerNorm is folded into this GEMM's weights, else 1)
    constexpr int NCH = BN / 64;                                                   // 32-column chunks per warp (half a tile)
    const int row = m0 + quad * 32 + lane;                                         // this thread's output row
    uint32_t rr[2][32];
    tmem_ld_32x32b_x32_issue(t_addr + c_beg, rr[0]);
    if (EPI == EPI_QKV && n0 >= g.d_model) {
        // K / V tile: scatter this thread's row segment into the KV page of (env, pos)
        const bool live = row < M;
        const int part = n0 / g.d_model;                                   // 1 = K, 2 = V (tiles never straddle: d % BN == 0)
        __nv_bfloat16* cache = part == 1 ? g.kcache : g.vcache;
        size_t base_off = 0;                                               // (page, head 0, position inside the page)
        if (live) {
            const int pos = g.row_pos[row];
            base_off = ((size_t)g.kv_table[(size_t)g.row_env[row] * g.kv_ppe + (pos >> 5)] * g.n_head * 32 + (size_t)(pos & 31)) * 64;
        }
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
            const int c0 = c_beg + 32 * ci;
            tmem_ld_wait();
            if (ci + 1 < NCH) tmem_ld_32x32b_x32_issue(t_addr + c0 + 32, rr[(ci + 1) & 1]);
            const uint32_t* r = rr[ci & 1];
            if (live) {
                const int c = n0 + c0 - part * g.d_model;                  // column inside K (or V): head = c / 64
                __nv_bfloat16* dst = cache + base_off + (size_t)(c >> 6) * (32 * 64) + (c & 63);
                uint32_t pk[16];                                           // 32 columns = 64 bytes: two 32-byte stores (one full sector per lane each)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b = tc_bias4(g, sbias, n0, c0 + 4 * j);
                    pk[2 * j] = pack_bf16x2_tc(fmaf(rs, __uint_as_float(r[4 * j]), b.x), fmaf(rs, __uint_as_float(r[4 * j + 1]), b.y));
                    pk[2 * j + 1] = pack_bf16x2_tc(fmaf(rs, __uint_as_float(r[4 * j + 2]), b.z), fmaf(rs, __uint_as_float(r[4 * j + 3]), b.w));
                }
                st_global_v8(dst, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
                st_global_v8(dst + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
            }
        }
    } else if constexpr (EPI == EPI_RESID && XB) {
        // residual GEMM that also feeds a LayerNorm-folded consumer.  Per 32-column chunk the warp
        //   1. reads its 32-row x 128-byte box of x with COALESCED 16-byte loads (lane -> row 4i + lane / 8, chunk lane % 8: four full lines
        //      per instruction; a first version had every thread read its own row, 32 lines per instruction, and was L1-wavefront bound:
        //      +17 ms per rollout), one chunk ahead of its use, and parks it in the swizzled staging box,
        //   2. every thread updates ITS row in place (x + acc + bias) and the box leaves through a plain TMA store,
        //   3. the bf16 copy of the row goes out as two 32-byte stores per thread (full sectors), and the row's partial (sum, sum of squares)
        //      over this warp's columns lands in a fixed slot: no atomics, deterministic.
        const bool live = row < M;
        const int lr = lane >> 3, lc = lane & 7;
        const int brow0 = m0 + quad * 32;                                  // first row of the warp's box
        __nv_bfloat16* brow = g.xb_out + (size_t)row * g.N + n0;
        float s1 = 0.f, s2 = 0.f;
        // x boxes are loaded TWO chunks ahead (xin[8 * (ci & 1)]).  The TMEM chunk is single buffered here (registers): the next chunk's load is
        // issued as soon as this chunk's accumulator values are consumed, and is in flight during the stores and the next box's staging
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
            const int c0 = c_beg + 32 * ci;
            float4* xc = xin + 8 * (ci & 1);
            if (lane == 0) bulk_wait_read<0>();                            // the TMA engine has read the previous box
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int br = 4 * i + lr;
                st_shared_v4(stg + br * 128 + ((lc ^ (br & 7)) << 4), __float_as_uint(xc[i].x), __float_as_uint(xc[i].y), __float_as_uint(xc[i].z), __float_as_uint(xc[i].w));
            }
            __syncwarp();
            if (ci + 2 < NCH) tc_xb_load_box(g, M, brow0, n0 + c0 + 64, lane, xc);
            else if (NCH % 2 == 0 && nm0 >= 0) tc_xb_load_box(g, M, nm0 + quad * 32, nn0 + c_beg + 32 * (ci + 2 - NCH), lane, xc);
            tmem_ld_wait();
            const uint32_t* r = rr[0];
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {                                  // 4 columns of this thread's row per step
                const uint32_t addr = stg + lane * 128 + ((j ^ (lane & 7)) << 4);
                const float4 x4 = ld_shared_f4(addr);
                float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g.bias) b = tc_bias4(g, sbias, n0, c0 + 4 * j);
                const float4 v = make_float4(x4.x + (__uint_as_float(r[4 * j]) + b.x), x4.y + (__uint_as_float(r[4 * j + 1]) + b.y),
                                             x4.z + (__uint_as_float(r[4 * j + 2]) + b.z), x4.w + (__uint_as_float(r[4 * j + 3]) + b.w));
                st_shared_v4(addr, __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
                s1 += (v.x + v.y) + (v.z + v.w);
                s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
                pk[2 * j] = pack_bf16x2_tc(v.x, v.y); pk[2 * j + 1] = pack_bf16x2_tc(v.z, v.w);
            }
            if (ci + 1 < NCH) tmem_ld_32x32b_x32_issue(t_addr + c0 + 32, rr[0]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { tma_store_2d(&tmOut, stg, n0 + c0, brow0); bulk_commit(); }
            if (live) {
                st_global_v8(brow + c0, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
                st_global_v8(brow + c0 + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
            }
        }
        if (live) g.stat_out[(size_t)row * 8 + 2 * (n0 / BN) + (c_beg != 0 ? 1 : 0)] = make_float2(s1, s2);
    } else if constexpr (EPI == EPI_RESID) {
        // fp32 residual: 32-column boxes, x += acc + bias by TMA reduce-add
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
            const int c0 = c_beg + 32 * ci;
            tmem_ld_wait();
            if (ci + 1 < NCH) tmem_ld_32x32b_x32_issue(t_addr + c0 + 32, rr[(ci + 1) & 1]);
            uint32_t* r = rr[ci & 1];
            if (g.bias) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b = tc_bias4(g, sbias, n0, c0 + 4 * j);
                    r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + b.x); r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + b.y);
                    r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + b.z); r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + b.w);
                }
            }
            if (lane == 0) bulk_wait_read<0>();                            // the TMA engine has read the previous box
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) st_shared_v4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { tma_reduce_add_2d(&tmOut, stg, n0 + c0, m0 + quad * 32); bulk_commit(); }
        }
    } else {
        // bf16 outputs (GeLU, ReLU, Q): 64-column boxes by TMA store
        static_assert(EPI == EPI_RESID || (BN / 2) % 64 == 0, "bf16 epilogues store 64-column boxes: BN / 2 must be a multiple of 64");
        uint32_t pk[32];                                                   // 64 columns of this row, packed bf16
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
            const int c0 = c_beg + 32 * ci;
            tmem_ld_wait();
            if (ci + 1 < NCH) tmem_ld_32x32b_x32_issue(t_addr + c0 + 32, rr[(ci + 1) & 1]);
            const uint32_t* r = rr[ci & 1];
#pragma unroll
            for (int j = 0; j < 4; ++j) {                                  // 8 columns -> one 16-byte chunk
                const float4 b0 = tc_bias4(g, sbias, n0, c0 + 8 * j), b1 = tc_bias4(g, sbias, n0, c0 + 8 * j + 4);
                float v[8] = {fmaf(rs, __uint_as_float(r[8 * j]), b0.x), fmaf(rs, __uint_as_float(r[8 * j + 1]), b0.y), fmaf(rs, __uint_as_float(r[8 * j + 2]), b0.z),
                              fmaf(rs, __uint_as_float(r[8 * j + 3]), b0.w), fmaf(rs, __uint_as_float(r[8 * j + 4]), b1.x), fmaf(rs, __uint_as_float(r[8 * j + 5]), b1.y),
                              fmaf(rs, __uint_as_float(r[8 * j + 6]), b1.z), fmaf(rs, __uint_as_float(r[8 * j + 7]), b1.w)};
                if constexpr (EPI == EPI_GELU) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] = gelu_new(v[q]);
                }
                if constexpr (EPI == EPI_RELU) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] = g.out_scale * fmaxf(v[q], 0.f);
                }
                const int ch = (ci & 1) * 4 + j;
                pk[4 * ch] = pack_bf16x2_tc(v[0], v[1]); pk[4 * ch + 1] = pack_bf16x2_tc(v[2], v[3]);
                pk[4 * ch + 2] = pack_bf16x2_tc(v[4], v[5]); pk[4 * ch + 3] = pack_bf16x2_tc(v[6], v[7]);
            }
            if (ci & 1) {                                                  // two chunks = one 64-column box
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) st_shared_v4(stg + lane * 128 + ((ch ^ (lane & 7)) << 4), pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) { tma_store_2d(&tmOut, stg, g.out_col0 + n0 + c0 - 32, m0 + quad * 32); bulk_commit(); }
            }
        }
    }
}

// LayerNorm of the 128 rows of a finished m-tile of the fp32 residual stream (EPI_RESID with GemmArgs::ln_out): the eight epilogue
// warps take 16 rows each, two rows in flight per warp (the rows were just reduce-added by TMA: they come from L2).  Same two-pass
// arithmetic as layernorm_kernel.
__device__ __forceinline__ void tc_ln_rows(const GemmArgs& g, int m0, int M, int w, int lane) {
    const int d = g.N, nv = d >> 7;                                        // float4 per lane (d % 128 == 0, d <= 1024)
#pragma unroll 1
    for (int i = 0; i < 16; i += 2) {
        const int r0 = m0 + w * 16 + i;
        if (r0 >= M) break;
        const bool two = r0 + 1 < M;
        float4 v[2][8];
        float s[2] = {0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float4* src = reinterpret_cast<const float4*>(g.resid + (size_t)(r0 + (two ? k : 0)) * d);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < nv) { v[k][j] = __ldcg(src + lane + 32 * j); s[k] += v[k][j].x + v[k][j].y + v[k][j].z + v[k][j].w; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { s[0] += __shfl_xor_sync(0xFFFFFFFFu, s[0], o); s[1] += __shfl_xor_sync(0xFFFFFFFFu, s[1], o); }
        const float mean[2] = {s[0] / d, s[1] / d};
        float q[2] = {0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < nv) {
                    const float a = v[k][j].x - mean[k], b = v[k][j].y - mean[k], c = v[k][j].z - mean[k], e = v[k][j].w - mean[k];
                    q[k] += a * a + b * b + c * c + e * e;
                }
#pragma unroll
        for (int o = 16; o; o >>= 1) { q[0] += __shfl_xor_sync(0xFFFFFFFFu, q[0], o); q[1] += __shfl_xor_sync(0xFFFFFFFFu, q[1], o); }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (k == 1 && !two) break;
            const float rstd = rsqrtf(q[k] / d + g.ln_eps);
            __nv_bfloat16* dst = g.ln_out + (size_t)(r0 + k) * d;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < nv) {
                    const int c = (lane + 32 * j) * 4;
                    const float4 g4 = __ldg(reinterpret_cast<const float4*>(g.ln_gamma + c)), b4 = __ldg(reinterpret_cast<const float4*>(g.ln_beta + c));
                    uint2 pk;
                    pk.x = pack_bf16x2_tc((v[k][j].x - mean[k]) * rstd * g4.x + b4.x, (v[k][j].y - mean[k]) * rstd * g4.y + b4.y);
                    pk.y = pack_bf16x2_tc((v[k][j].z - mean[k]) * rstd * g4.z + b4.z, (v[k][j].w - mean[k]) * rstd * g4.w + b4.w);
                    *reinterpret_cast<uint2*>(dst + c) = pk;
                }
        }
    }
}

// one logit of the sampling epilogue: score = logit/T (+ Gumbel noise), running arg-max (lowest column wins ties)
template <bool NOISE, bool MASK>
__device__ __forceinline__ void sample_update(float logit, int col, float inv_temp, uint64_t key, int n_vocab, float& best, int& best_col) {
    float s = logit * inv_temp;
    if constexpr (NOISE) s += gumbel_from_bits(noise_bits(key, (uint32_t)col));
    if constexpr (MASK) s = col < n_vocab ? s : -INFINITY;
    const bool take = s > best;
    best = take ? s : best;
    best_col = take ? col : best_col;
}
template <bool NOISE, bool MASK>
__device__ __forceinline__ void sample_chunk(const uint32_t r[32], int col0, float inv_temp, uint64_t key, int n_vocab, float best[4], int best_col[4]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) sample_update<NOISE, MASK>(__uint_as_float(r[j]), col0 + j, inv_temp, key, n_vocab, best[j & 3], best_col[j & 3]);
}

// XB (EPI_RESID only): the residual epilogue that also writes the bf16 copy + row statistics for a LayerNorm-folded consumer
// (GemmArgs::xb_out / stat_out); its own instantiation so that its register allocation does not touch the plain kernels
template <int EPI, int BN = TC_BN, bool XB = false>
// 168 registers per thread: the pipelined epilogues (two TMEM chunks in flight) and the fused LayerNorm (two rows in flight) without spills
__global__ void __maxnreg__(168) gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                                   const __grid_constant__ CUtensorMap tmW,
                                                                                   const __grid_constant__ CUtensorMap tmW1,
                                                                                   const __grid_constant__ CUtensorMap tmW2,
                                                                                   const __grid_constant__ CUtensorMap tmOut,
                                                                                   const GemmArgs g) {
    extern __shared__ unsigned char tc_smem_raw[];
    pdl_trigger();                                                       // the next kernel may start its prologue behind my tail
    // the live row count: issued FIRST so that its global-memory round trip (~0.7 us, on the critical path of every launch: no TMA can be
    // issued before the tile count is known) overlaps barrier initialisation and the TMEM allocation.  With programmatic dependent
    // launch the predecessor must be complete before the load (griddepcontrol.wait); PDL is off by default (measured slower).
    pdl_wait();
    const int M = __ldg(g.M_dev);
    const uint32_t raw = smem_u32(tc_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                        // 128B swizzle atoms need 1024-byte alignment
    unsigned char* base_ptr = tc_smem_raw + (base - raw);
    const uint32_t epi_base = base + TC_STAGES * TcTile<BN>::STAGE_BYTES;
    const uint32_t bias_base = epi_base + TC_EPI_BYTES;                  // BN floats: the current tile's bias
    const uint32_t bar_base = bias_base + TC_BIAS_BYTES;
    // barrier layout (8 bytes each): full[4], empty[4], tmem_full[2], tmem_empty[2], then the TMEM base address word and the
    // "this CTA finishes the m-tile's LayerNorm" flag
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (TC_STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * TC_STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * TC_STAGES + 2 + a); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + TC_STAGES * TcTile<BN>::STAGE_BYTES + TC_EPI_BYTES + TC_BIAS_BYTES + 8 * (2 * TC_STAGES + 4));
    volatile uint32_t* ln_flag = tmem_slot + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_acc = g.n_acc, n_stage = n_acc == 2 ? 1 : 2;          // TMEM: 512 columns = n_stage x n_acc accumulators of 256

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), TcCfg<EPI>::EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
        if (g.n_seg > 1) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW2) : "memory");
        }
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_tiles = g.N / BN, m_tiles = (M + TC_BM - 1) / TC_BM, total = M > 0 ? n_tiles * m_tiles : 0;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int m0 = (tile / n_tiles) * TC_BM, n0 = (tile % n_tiles) * BN;
                for (int sg = 0; sg < g.n_seg; ++sg) {
                    const CUtensorMap* tw = g.seg[sg].w_idx == 0 ? &tmW : (g.seg[sg].w_idx == 1 ? &tmW1 : &tmW2);
                    const int a_kb0 = g.seg[sg].a_kb0, nkb = g.seg[sg].kb_count;
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(empty_bar(stage), phase ^ 1u);
                        mbar_expect_tx(full_bar(stage), TcTile<BN>::STAGE_BYTES);
                        const uint32_t sA = base + stage * TcTile<BN>::STAGE_BYTES, sB = sA + TC_A_BYTES;
                        tma_load_2d(sA, &tmA, full_bar(stage), (a_kb0 + kb) * TC_BK, m0);
                        tma_load_2d(sB, tw, full_bar(stage), kb * TC_BK, n0);
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);               // epilogue has drained this accumulator stage
                tcgen05_fence_after();
                uint32_t started = 0;                                     // bit a: accumulator a already holds a partial sum
                for (int sg = 0; sg < g.n_seg; ++sg) {
                    const int a = g.seg[sg].acc, nkb = g.seg[sg].kb_count;
                    const uint32_t d_tmem = tmem_base + (uint32_t)((acc * n_acc + a) * BN);
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(full_bar(stage), phase);                // TMA bytes have landed
                        tcgen05_fence_after();
                        const uint32_t sA = base + stage * TcTile<BN>::STAGE_BYTES, sB = sA + TC_A_BYTES;
                        const uint64_t da = umma_desc_sw128(sA), db = umma_desc_sw128(sB);
#pragma unroll
                        for (int k = 0; k < TC_BK / TC_UMMA_K; ++k)       // +32 bytes per UMMA_K step inside the swizzle atom
                            tcgen05_mma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), TcTile<BN>::IDESC, (((started >> a) & 1u) | (uint32_t)kb | (uint32_t)k) ? 1u : 0u);
                        tcgen05_commit(empty_bar(stage));                 // frees the smem slot when these MMAs retire
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
                    }
                    started |= 1u << a;
                }
                tcgen05_commit(tfull_bar(acc));                           // accumulators complete -> epilogue
                if (++acc == n_stage) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ===== epilogue warps 2..: TMEM lane quadrant = warp % 4 =====
        const int quad = warp & 3;
        const uint32_t stg = epi_base + (uint32_t)(warp - 2) * TC_EPI_BUF;         // this warp's 4 KB staging box, 1024-byte aligned
        const int half = (warp - 2) >> 2;                                          // which half of the tile's columns this warp drains
        const int et = (int)threadIdx.x - 64;                                      // thread index inside the epilogue group (256 threads)
        constexpr bool STORING = EPI != EPI_SAMPLE && EPI != EPI_LOGPROB;
        const uint32_t sbias = (STORING && g.bias) ? bias_base : 0u;
        auto epi_bar = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };      // named barrier of the eight epilogue warps
        int acc = 0; uint32_t acc_phase = 0;
        float4 xin[16];                                                            // XB residual epilogue: two boxes of x in flight
        bool x_ready = false;
        float bias_next = 0.f;                                                     // this thread's bias value of the CTA's next tile
        if (sbias && et < BN && (int)blockIdx.x < total) bias_next = __ldg(g.bias + ((int)blockIdx.x % n_tiles) * BN + et);
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            const int m0 = (tile / n_tiles) * TC_BM, n0 = (tile % n_tiles) * BN;
            if (sbias) {
                // the tile's bias goes to shared memory once (while the MMAs of this tile are still running) instead of
                // every thread re-loading all BN values from global memory for its own row.  The value was fetched into a register
                // during the PREVIOUS tile's epilogue: no global-memory round trip between the two barriers
                epi_bar();                                                         // nobody still reads the previous tile's values
                if (et < BN) asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_base + 4u * (uint32_t)et), "f"(bias_next) : "memory");
                epi_bar();
                const int nt_b = tile + gridDim.x;
                if (nt_b < total && et < BN) bias_next = __ldg(g.bias + (nt_b % n_tiles) * BN + et);
            }
            int nm0 = -1, nn0 = 0;
            if constexpr (EPI == EPI_RESID && XB) {
                // the NEXT tile's part of x on its way into L2 (the plain epilogue's reduce-add never waits for x; this one reads it)
                const int nt = tile + gridDim.x;
                if (nt < total) {
                    nm0 = (nt / n_tiles) * TC_BM; nn0 = (nt % n_tiles) * BN;
                    const int prow = nm0 + quad * 32 + lane;
                    if (prow < M) {
                        const float* px = g.resid + (size_t)prow * g.N + nn0 + half * (BN / 2);
#pragma unroll
                        for (int k = 0; k < BN / 2; k += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(px + k));
                    }
                }
                if (!x_ready) {                                                    // the first two x boxes of this warp: under the tile's MMAs
                    tc_xb_load_box(g, M, m0 + quad * 32, n0 + half * (BN / 2), lane, xin);
                    tc_xb_load_box(g, M, m0 + quad * 32, n0 + half * (BN / 2) + 32, lane, xin + 8);
                }
                x_ready = (BN / 64) % 2 == 0 && nm0 >= 0;                          // even chunk count: this tile's epilogue loads the next tile's first boxes
            }
            // LayerNorm-folded consumer: the row's partial (sum, sum of squares) pairs, also fetched under the MMAs
            float4 st4[4];
            bool has_stat = false;
            if constexpr (EPI == EPI_QKV || EPI == EPI_GELU) {
                has_stat = g.stat_in != nullptr && m0 + quad * 32 + lane < M;
                if (has_stat) {
                    const float4* sp = reinterpret_cast<const float4*>(g.stat_in + (size_t)(m0 + quad * 32 + lane) * 8);
#pragma unroll
                    for (int i = 0; i < 4; ++i) st4[i] = __ldcg(sp + i);
                }
            }
            mbar_wait(tfull_bar(acc), acc_phase);
            tcgen05_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * n_acc * BN);
            const int row = m0 + quad * 32 + lane;                                 // this thread's output row
            if constexpr (EPI == EPI_SAMPLE || EPI == EPI_LOGPROB) {
                // warps w and w+4 share the lane quadrant; each takes one 128-column half of the tile
                const bool live = row < M;
                const uint64_t key = (live && EPI == EPI_SAMPLE) ? g.row_key[row] : 0ull;
                const bool masked = n0 + BN > g.n_vocab;                         // only the last N tile has padded columns
                float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                int best_col[4] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF};
                float lp_m = -INFINITY, lp_s = 0.f;                                 // EPI_LOGPROB: running max / sum of exp
                const int lp_tgt = (EPI == EPI_LOGPROB && live) ? g.lp_target[row] : -1;
#pragma unroll 1
                for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(t_addr + c0, r);
                    const int col0 = n0 + c0;
                    if (g.bias) {                                                   // value-perturbed sampler: + beta * b2 of the Q head(s)
#pragma unroll
                        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __ldg(g.bias + col0 + j));
                    }
                    if (n_acc == 2) {                                               // logit = min over the two Q heads
                        uint32_t r2[32];
                        tmem_ld_32x32b_x32(t_addr + BN + c0, r2);
#pragma unroll
                        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fminf(__uint_as_float(r[j]), __uint_as_float(r2[j]) + __ldg(g.bias2 + col0 + j)));
                    }
                    if (g.logits_out && live) {                                     // parity hook only
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < g.n_vocab) g.logits_out[(size_t)row * g.n_vocab + col0 + j] = __uint_as_float(r[j]);
                    }
                    if constexpr (EPI == EPI_LOGPROB) {
                        float cm = -INFINITY;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float v = (col0 + j < g.n_vocab) ? __uint_as_float(r[j]) * g.inv_temp : -INFINITY;
                            r[j] = __float_as_uint(v);
                            cm = fmaxf(cm, v);
                            if (col0 + j == lp_tgt) g.lp_tlogit[row] = v;
                        }
                        if (cm > -INFINITY) {
                            const float m_new = fmaxf(lp_m, cm);
                            float add = 0.f;
#pragma unroll
                            for (int j = 0; j < 32; ++j) add += __expf(__uint_as_float(r[j]) - m_new);
                            lp_s = lp_s * __expf(lp_m - m_new) + add;
                            lp_m = m_new;
                        }
                    } else if (g.do_sample) {
                        if (masked) sample_chunk<true, true>(r, col0, g.inv_temp, key, g.n_vocab, best, best_col);
                        else sample_chunk<true, false>(r, col0, g.inv_temp, key, g.n_vocab, best, best_col);
                    } else {
                        sample_chunk<false, true>(r, col0, g.inv_temp, key, g.n_vocab, best, best_col);
                    }
                }
                if constexpr (EPI == EPI_LOGPROB) {
                    if (live) g.lp_partial[(size_t)row * (2 * n_tiles) + 2 * (n0 / BN) + half] = make_float2(lp_m, lp_s);
                }
#pragma unroll
                for (int q = 1; q < 4; ++q)
                    if (best[q] > best[0] || (best[q] == best[0] && best_col[q] < best_col[0])) { best[0] = best[q]; best_col[0] = best_col[q]; }
                if (EPI == EPI_SAMPLE && live && best_col[0] != 0x7FFFFFFF) atomicMax(g.best + row, pack_best(best[0], best_col[0]));
            } else {
                float rs = 1.0f;
                if constexpr (EPI == EPI_QKV || EPI == EPI_GELU) {
                    if (has_stat) {                                                // 1 / sigma of this thread's row from the producer's partial sums
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (2 * i < g.stat_n) { s1 += st4[i].x; s2 += st4[i].y; }
                            if (2 * i + 1 < g.stat_n) { s1 += st4[i].z; s2 += st4[i].w; }
                        }
                        const float mean = s1 / g.K;
                        rs = rsqrtf(fmaxf(s2 / g.K - mean * mean, 0.f) + g.ln_eps);
                    }
                }
                tc_store_tile<EPI, BN, XB>(g, tmOut, t_addr, M, m0, n0, quad, lane, stg, half * (BN / 2), sbias, rs, xin, nm0, nn0);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == n_stage) { acc = 0; acc_phase ^= 1u; }
            if constexpr (EPI == EPI_RESID && !XB) {
                if (g.ln_out) {
                    // fused LayerNorm: once every N tile of these 128 rows has been reduce-added into the residual stream (arrival
                    // counter per m-tile, device scope), the CTA that arrives last normalises the rows for the next GEMM
                    if (lane == 0) { bulk_wait_all(); asm volatile("fence.proxy.async;" ::: "memory"); __threadfence(); }
                    epi_bar();                                                     // all eight warps' reductions of this tile have landed
                    if (et == 0) {
                        int* cnt = g.ln_cnt + m0 / TC_BM;
                        const int old = atomicAdd(cnt, 1);
                        const uint32_t last = old + 1 == n_tiles ? 1u : 0u;
                        if (last) *cnt = 0;                                        // ready for the next launch (same stream: ordered)
                        __threadfence();
                        *ln_flag = last;
                    }
                    epi_bar();
                    if (*ln_flag) tc_ln_rows(g, m0, M, warp - 2, lane);
                }
            }
        }
        // the staging boxes must have been read before the CTA's shared memory goes away; the global writes themselves complete with
        // the grid (the fused LayerNorm above is the one consumer inside the kernel, and it waits for full completion itself)
        if (lane == 0) bulk_wait_read<0>();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

}  // namespace lmrl
