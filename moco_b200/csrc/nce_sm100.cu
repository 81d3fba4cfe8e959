// tcgen05 / TMEM / TMA kernels of the InfoNCE head (sm_100a only).
//
//  nce_stats_kernel<G,KPS,DENSE>  S = q . Queue^T on tcgen05 (cta_group::G), accumulators double-buffered in
//                        TMEM, epilogue = x/T, online log-sum-exp per row (and optional dense logits).
//                        Replaces torch.mm + cat + div + CrossEntropyLoss + softmax
//                        (moco/NCE/Contrast.py:25-27, NCECriterion.py:11-13, train.py:264).
//  (the dq pass / the one-sweep head kernel live in nce_dq2_sm100.cu)
//
// Data layout: q [N, C] bf16 and queue [K, C] bf16 are row-major in HBM ("K-major" for the S GEMM).
// TMA stages [rows x 64 elements] boxes (128 B per row, 128B swizzle) into shared memory; a tile of
// R rows is stored as C/64 slabs of R x 128 B.
//
// What the measurements on B200 say about this shape of kernel (profiles/README.md):
//  * one thread can keep the tensor pipe 100 % busy, but only just: a tcgen05.mma costs the issuing thread
//    ~105 cycles and the issue queue is ~3 MMAs deep, so every other instruction in the issue loop counts;
//  * tcgen05.ld drains 175-460 B/clk/SM (4-16 warps), MUFU.EX2 sustains 15.8/clk/SM;
//  * L2 -> SM delivery saturates near 6.3 KB/clk chip-wide; TMA multicast across <= 4 CTAs did not relieve it
//    (measured time-neutral in round 1 and removed), CTA pairs (cta_group::2) do.
#include <cuda.h>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tc_common.cuh"

namespace moco {

// =====================================================================================
// Kernel A: per-slice softmax statistics (and optional dense logits)
// =====================================================================================
constexpr int kStatsBN = 256;          // queue rows per tile (UMMA N)
// warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps 4-19 epilogue (two ping-pong groups of 8 warps)

struct StatsArgs {
    int N, C, K;
    int mblks, slices, n_pad, num_tiles, stages;
    float inv_T;
    float* logits;        // optional [N, K+1]
    float2* part_ms;      // [slices, n_pad]
    int debug;            // bring-up only (env MOCO_DEBUG_MODE & 1): skip the epilogue math (tools/pipe_probe.py)
};

// G  = tcgen05 cta_group (1: M = 128 per CTA; 2: M = 256 per CTA pair, B tile split across the pair)
// KPS = 64-wide K chunks per shared-memory stage (1 or 2).  The MMA-issuing thread needs ~105 cycles per
//       tcgen05.mma it issues plus ~150 cycles of barrier wait / fence / commit per stage (tools/umma_bench.cu,
//       tools/trace_probe.py); with 4 MMAs (512 tensor cycles) per stage that is more than the stage holds, with
//       8 it is not.
// DENSE = 1: instantiation for the dense-logits API -- the [N, K+1] fp32 logits are written with full-line
//       coalesced stores (each 32 x 32 register block is transposed across the warp with shuffles first), not
//       with one 4-byte store per row per lane.  DENSE = 0 keeps a plain per-lane store for the rarely used
//       variant kernels and never pays registers for the transpose on the fused path.
template <int G, int KPS, int DENSE>
__global__ void __launch_bounds__(128 + 16 * 32, 1)
nce_stats_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_queue,
                 const StatsArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kChunkBytes = (kStatsBN / G) * 128;         // one 64-wide K chunk of this CTA's B rows
    constexpr int kStageBytes = kChunkBytes * KPS;
    constexpr int kEpiWarps = 16;
    // two groups of 8 warps in ping-pong -- group p owns accumulator buffer p and drains the tiles of parity p,
    // so the exps of tile t overlap the drain of t+1.
    constexpr int kEpiCols = 128;                          // accumulator columns per epilogue thread per tile
    constexpr int kEpiChunks = kEpiCols / 32;              // 32-column tcgen05.ld per thread per tile
    constexpr int kTileStep = 2;
    const int kchunks = a.C >> 6;
    const int NS = a.stages;
    uint8_t* q_s = smem;
    uint8_t* b_s = q_s + kchunks * kSlab;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_s + (size_t)NS * kStageBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + NS;
    uint64_t* tfull = bars + 2 * NS;
    uint64_t* tempty = bars + 2 * NS + 2;
    uint64_t* qfull = bars + 2 * NS + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 5);
    float2* red_s = reinterpret_cast<float2*>(bars + 2 * NS + 6);   // [kEpiWarps/4 - 1][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool kClustered = G > 1;
    const uint32_t rank = kClustered ? cluster_ctarank() : 0u;  // rank inside the MMA pair
    const int cluster_id = blockIdx.x / G;
    const int mblk = cluster_id % a.mblks;
    const int slice = cluster_id / a.mblks;
    const int t0 = (int)(((long long)slice * a.num_tiles) / a.slices);
    const int t1 = (int)(((long long)(slice + 1) * a.num_tiles) / a.slices);
    const int row0 = (mblk * G + (int)rank) * kRowsPerCta;

    pdl_launch_dependents();
    // set-up that touches no global memory: overlaps the predecessor kernel (prep) under PDL
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_queue);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 8 * G); }
        mbar_init(qfull, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc<G>(tmem_slot, 512);
        tmem_relinquish<G>();
    }
    pdl_wait();                                  // predecessor complete: q_bf16 and the queue are final
    tc_fence_before();
    if (kClustered) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            // ------------------------------------------------ TMA producer
            // NOTE (measured, tools/trace_probe.py): the producer and the MMA issuer are single threads whose
            // scalar instruction stream runs at ~4-6 cycles per dependent instruction; a runtime `it % NS`,
            // `it / NS` or a rebuilt 64-bit descriptor per stage costs hundreds of cycles -- more than the
            // 512 tensor cycles a stage holds.  Everything in these loops is therefore strength-reduced to
            // running counters and adds.
            const uint32_t qfull_addr = (G == 2) ? mapa_shared(smem_u32(qfull), 0) : smem_u32(qfull);
            if (rank == 0) mbar_arrive_expect_tx(qfull, (uint32_t)(kchunks * kSlab * G));
            for (int kc = 0; kc < kchunks; ++kc) {
                if (G == 2) tma_load_2d_2sm(&tm_q, qfull_addr, q_s + kc * kSlab, kc * 64, row0);
                else        tma_load_2d(&tm_q, qfull, q_s + kc * kSlab, kc * 64, row0);
            }
            int st = 0;
            uint32_t ph = 0;
            uint8_t* dst = b_s;
            const uint32_t full0 = (G == 2) ? mapa_shared(smem_u32(&full[0]), 0) : 0u;
            int brow = t0 * kStatsBN + (int)rank * (kStatsBN / G);
            for (int t = t0; t < t1; ++t, brow += kStatsBN) {
                for (int kc = 0; kc < kchunks; kc += KPS) {
                    mbar_wait(&empty[st], ph ^ 1u);
                    if (rank == 0) mbar_arrive_expect_tx(&full[st], (uint32_t)(kStageBytes * G));
#pragma unroll
                    for (int kk = 0; kk < KPS; ++kk) {
                        uint8_t* d = dst + kk * kChunkBytes;
                        if (G == 2) tma_load_2d_2sm(&tm_queue, full0 + (uint32_t)st * 8u, d, (kc + kk) * 64, brow);
                        else        tma_load_2d(&tm_queue, &full[st], d, (kc + kk) * 64, brow);
                    }
                    dst += kStageBytes;
                    if (++st == NS) { st = 0; ph ^= 1u; dst -= (size_t)NS * kStageBytes; }
                }
            }
        }
    } else if (warp == 1) {
        // elect.sync (not `lane == 0`): ptxas then knows a single thread runs the region and issues the tcgen05.mma
        // instructions back to back from uniform registers; under a threadIdx predicate it wraps EVERY MMA in an
        // elect/branch loop (~65-90 cycles per MMA, the "issue-bound" limit of round 1)
        if (rank == 0 && elect_one()) {
            // ------------------------------------------------ MMA issuer (pair leader only)
            const uint32_t idesc = make_idesc_bf16(128 * G, kStatsBN, 0, 0);
            // descriptors differ from tile to tile only in the 14-bit start-address field: build once, then add
            const uint64_t a_desc0 = make_sw128_desc(smem_u32(q_s), 0, 1024);
            const uint64_t b_desc0 = make_sw128_desc(smem_u32(b_s), 0, 1024);
            constexpr uint64_t kStageUnits = (uint64_t)(kStageBytes >> 4), kSlabUnits = (uint64_t)(kSlab >> 4);
            constexpr uint64_t kChunkUnits = (uint64_t)(kChunkBytes >> 4);
            mbar_wait(qfull, 0);
            tc_fence_after();
            int st = 0;
            uint32_t ph = 0;
            uint64_t b_desc = b_desc0;
            uint32_t acc = 0, aph = 0;
            for (int t = t0; t < t1; ++t) {
                mbar_wait(&tempty[acc], aph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)kStatsBN;
                uint64_t a_desc = a_desc0;
                for (int kc = 0; kc < kchunks; kc += KPS) {
                    mbar_wait(&full[st], ph);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < KPS; ++kk) {
                        const uint64_t ad = a_desc + (uint64_t)kk * kSlabUnits, bd = b_desc + (uint64_t)kk * kChunkUnits;
                        umma_ss<G>(d_tmem, ad, bd, idesc, (uint32_t)((kc | kk) != 0));
                        umma_ss<G>(d_tmem, ad + 2, bd + 2, idesc, 1u);
                        umma_ss<G>(d_tmem, ad + 4, bd + 4, idesc, 1u);
                        umma_ss<G>(d_tmem, ad + 6, bd + 6, idesc, 1u);
                    }
                    umma_commit<G>(&empty[st]);
                    a_desc += kSlabUnits * KPS;
                    b_desc += kStageUnits;
                    if (++st == NS) { st = 0; ph ^= 1u; b_desc = b_desc0; }
                }
                umma_commit<G>(&tfull[acc]);
                acc ^= 1u;
                aph ^= (acc == 0u) ? 1u : 0u;
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------- epilogue (kEpiWarps warps)
        // Warp w may only touch TMEM lanes 32*(w%4)..+31.  A thread owns one q row and 128 accumulator columns
        // of a tile: four 32-column tcgen05.ld, each folded into the thread's running (max, sum) in the log2
        // domain; the accumulator buffer goes back to the MMA as soon as the last load has landed.  With 16
        // warps, warps 4-11 serve the even tiles (buffer 0) and warps 12-19 the odd tiles (buffer 1).
        const int quarter = warp & 3;
        const int cgrp = (warp - 4) >> 2;
        const int row_local = quarter * 32 + lane;
        const int grow = row0 + row_local;
        const float scale2 = a.inv_T * kLog2e;
        float m = -INFINITY, s = 0.f;
        float* lrow = (a.logits != nullptr && grow < a.N) ? a.logits + (size_t)grow * (a.K + 1) + 1 : nullptr;
        auto fold = [&](const uint32_t (&r)[32], int col0) {
            const int valid = a.K - col0;
            if (valid >= 32) {
                float c0 = fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1]));
                float c1 = fmaxf(__uint_as_float(r[2]), __uint_as_float(r[3]));
#pragma unroll
                for (int j = 4; j < 32; j += 4) {
                    c0 = fmaxf(c0, fmaxf(__uint_as_float(r[j + 0]), __uint_as_float(r[j + 1])));
                    c1 = fmaxf(c1, fmaxf(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
                }
                const float cm = fmaxf(c0, c1) * scale2;
                if (cm > m) { s *= ex2(m - cm); m = cm; }
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    s0 += ex2(fmaf(__uint_as_float(r[j + 0]), scale2, -m));
                    s1 += ex2(fmaf(__uint_as_float(r[j + 1]), scale2, -m));
                    s2 += ex2(fmaf(__uint_as_float(r[j + 2]), scale2, -m));
                    s3 += ex2(fmaf(__uint_as_float(r[j + 3]), scale2, -m));
                }
                s += (s0 + s1) + (s2 + s3);
                if constexpr (DENSE) {
                    if (a.logits != nullptr) {          // warp-uniform
                        // 32 x 32 transpose: afterwards v[k] of lane L is logit (row block + k, col0 + L)
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * a.inv_T;
#pragma unroll
                        for (int sft = 16; sft >= 1; sft >>= 1) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                if ((j & sft) == 0) {
                                    const bool hi = (lane & sft) != 0;
                                    const float send = hi ? v[j] : v[j | sft];
                                    const float recv = __shfl_xor_sync(0xffffffffu, send, sft);
                                    if (hi) v[j] = recv; else v[j | sft] = recv;
                                }
                            }
                        }
                        const int rbase = row0 + quarter * 32;
                        float* out = a.logits + (size_t)rbase * (a.K + 1) + 1 + col0 + lane;
#pragma unroll
                        for (int k2 = 0; k2 < 32; ++k2)
                            if (rbase + k2 < a.N) out[(size_t)k2 * (a.K + 1)] = v[k2];
                    }
                } else if (lrow) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) lrow[col0 + j] = __uint_as_float(r[j]) * a.inv_T;
                }
            } else if (valid > 0) {
                float cm = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; ++j) if (j < valid) cm = fmaxf(cm, __uint_as_float(r[j]));
                cm *= scale2;
                if (cm > m) { s *= ex2(m - cm); m = cm; }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (j < valid) {
                        s += ex2(fmaf(__uint_as_float(r[j]), scale2, -m));
                        if (lrow) lrow[col0 + j] = __uint_as_float(r[j]) * a.inv_T;
                    }
                }
            }
        };
        int lt = cgrp >> 1;
        for (int t = t0 + lt; t < t1; t += kTileStep, lt += kTileStep) {
            const int acc = lt & 1;
            const uint32_t aph = (uint32_t)(lt >> 1) & 1u;
            mbar_wait(&tfull[acc], aph);
            tc_fence_after();
            const int col = (cgrp & 1) * kEpiCols;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kStatsBN + col);
            {
#pragma unroll 1
                for (int ch = 0; ch < kEpiChunks; ++ch) {
                    uint32_t r[32];
                    tmem_ld32(taddr + (uint32_t)(ch * 32), r);
                    tmem_ld_wait();
                    if (ch == kEpiChunks - 1) {
                        // the whole accumulator slice of this warp is in registers: hand the buffer back to the MMA
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (G == 2) mbar_arrive_cluster_relaxed(&tempty[acc], 0);
                            else        mbar_arrive(&tempty[acc]);
                        }
                    }
                    if (!(a.debug & 1)) fold(r, t * kStatsBN + col + ch * 32);
                }
            }
        }
        // combine the column groups of each row (fixed order), then publish the slice partial
        if (cgrp > 0) red_s[(cgrp - 1) * kRowsPerCta + row_local] = make_float2(m, s);
        named_bar_sync(1, kEpiWarps * 32);
        if (cgrp == 0) {
            float M = m;
#pragma unroll
            for (int gq = 0; gq < kEpiWarps / 4 - 1; ++gq) M = fmaxf(M, red_s[gq * kRowsPerCta + row_local].x);
            float S = (m != -INFINITY) ? s * ex2(m - M) : 0.f;
#pragma unroll
            for (int gq = 0; gq < kEpiWarps / 4 - 1; ++gq) {
                float2 o = red_s[gq * kRowsPerCta + row_local];
                if (o.x != -INFINITY) S += o.y * ex2(o.x - M);
            }
            a.part_ms[(size_t)slice * a.n_pad + grow] = make_float2(M, S);
        }
    }

    __syncwarp();
    tc_fence_before();
    if (kClustered) cluster_sync_all(); else __syncthreads();
    if (warp == 2) tmem_dealloc<G>(tmem_base, 512);
}

// =====================================================================================
// Host side
// =====================================================================================
cudaError_t launch_nce_tc(NceTcParams& p, const NceWorkspace& ws, cudaStream_t stream) {
    if (p.C % 64 != 0 || p.C < 64 || p.C > 256 || p.N < 1 || p.K < 1) return cudaErrorNotSupported;
    const int G = p.cta_group;
    const int kchunks = p.C / 64;
    const int mblks = (p.N + 128 * G - 1) / (128 * G);
    if (mblks * G > p.num_sms) return cudaErrorNotSupported;
    const int num_tiles = (p.K + kStatsBN - 1) / kStatsBN;
    const int n_pad = mblks * G * 128;
    p.n_pad = n_pad;

    CUtensorMap tm_q, tm_queue;
    if (!make_tmap(&tm_q, p.q_bf16, p.N, p.C, 128)) return cudaErrorUnknown;
    if (!make_tmap(&tm_queue, p.queue, p.K, p.C, kStatsBN / G)) return cudaErrorUnknown;

    // CTA pairs: two 64-wide K chunks per smem stage (8 MMAs per barrier round trip; 57.4 -> 55.2 us at configs[4])
    const int KPS = (G == 2 && kchunks % 2 == 0) ? 2 : 1;
    const int stage_bytes = (kStatsBN / G) * 128 * KPS;
    const int fixed = kchunks * kSlab + 4096;     // q tile + barriers/red_s
    int stages = (kSmemBudget - fixed) / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages < 2) return cudaErrorNotSupported;
    const int smem = fixed + stages * stage_bytes + 1024;

    StatsArgs a;
    a.N = p.N; a.C = p.C; a.K = p.K;
    a.mblks = mblks; a.slices = 0; a.n_pad = n_pad; a.num_tiles = num_tiles; a.stages = stages;
    a.inv_T = p.inv_T;
    a.logits = p.logits;
    a.part_ms = ws.part_ms;
    a.debug = debug_mode();
    auto fill = [](StatsArgs& x, int slices) { x.slices = slices; };
    constexpr int kThreads = 128 + 16 * 32;
    const int per_slice = mblks * G;
    if (G == 2 && KPS == 2)
        return plan_and_launch(nce_stats_kernel<2, 2, 0>, kernel_cache(0), kThreads, smem, 2, mblks, per_slice, num_tiles,
                               n_pad, &p.slices, stream, tm_q, tm_queue, a, fill, true);
    if (G == 2)
        return plan_and_launch(nce_stats_kernel<2, 1, 0>, kernel_cache(1), kThreads, smem, 2, mblks, per_slice, num_tiles,
                               n_pad, &p.slices, stream, tm_q, tm_queue, a, fill, true);
    if (p.logits != nullptr)
        return plan_and_launch(nce_stats_kernel<1, 1, 1>, kernel_cache(2), kThreads, smem, 1, mblks, per_slice,
                               num_tiles, n_pad, &p.slices, stream, tm_q, tm_queue, a, fill, true);
    return plan_and_launch(nce_stats_kernel<1, 1, 0>, kernel_cache(3), kThreads, smem, 1, mblks, per_slice, num_tiles,
                           n_pad, &p.slices, stream, tm_q, tm_queue, a, fill, true);
}

}  // namespace moco
