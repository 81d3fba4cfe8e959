// tcgen05 / TMEM / TMA kernels of the InfoNCE head (sm_100a only).
//
//  nce_stats_kernel<G,CS,EW,KPS>  S = q . Queue^T on tcgen05 (cta_group::G), accumulators double-buffered in
//                        TMEM, epilogue = x/T, online log-sum-exp per row (and optional dense logits).
//                        Replaces torch.mm + cat + div + CrossEntropyLoss + softmax
//                        (moco/NCE/Contrast.py:25-27, NCECriterion.py:11-13, train.py:264).
//  nce_dq_kernel<CS>     first-generation dq pass (P through shared memory); the default dq pass is
//                        nce_dq2_sm100.cu.  Replaces autograd's backward GEMM (train.py:273) and the queue
//                        clone it needs (Contrast.py:24-25).
//
// Data layout: q [N, C] bf16 and queue [K, C] bf16 are row-major in HBM ("K-major" for the S GEMM).
// TMA stages [rows x 64 elements] boxes (128 B per row, 128B swizzle) into shared memory; a tile of
// R rows is stored as C/64 slabs of R x 128 B.
//
// What the measurements on B200 say about this shape of kernel (profiles/README.md):
//  * one thread can keep the tensor pipe 100 % busy, but only just: a tcgen05.mma costs the issuing thread
//    ~105 cycles and the issue queue is ~3 MMAs deep, so every other instruction in the issue loop counts;
//  * tcgen05.ld drains 175-460 B/clk/SM (4-16 warps), MUFU.EX2 sustains 15.8/clk/SM;
//  * L2 -> SM delivery saturates near 6.3 KB/clk chip-wide; TMA multicast across <= 4 CTAs does not relieve it,
//    CTA pairs (cta_group::2) do.
#include <cuda.h>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tc_common.cuh"

namespace moco {

// =====================================================================================
// Kernel A: per-slice softmax statistics (and optional dense logits)
// =====================================================================================
constexpr int kStatsBN = 256;          // queue rows per tile (UMMA N)
// warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps 4.. epilogue (EW = 8 or 16 epilogue warps)

struct StatsArgs {
    int N, C, K;
    int mblks, slices, n_pad, num_tiles, stages;
    float inv_T;
    float* logits;        // optional [N, K+1]
    float2* part_ms;      // [slices, n_pad]
    int debug;            // bring-up only (env MOCO_DEBUG_MODE & 1): skip the epilogue math (tools/pipe_probe.py)
};

// G  = tcgen05 cta_group (1: M = 128 per CTA; 2: M = 256 per CTA pair, B tile split across the pair)
// CS = CTAs per cluster that handle DIFFERENT q row blocks but the SAME queue tiles (G == 1 only): each
//      loads 1/CS of every queue tile and TMA-multicasts it to all CS CTAs, dividing L2->SM traffic by CS.
// KPS = 64-wide K chunks per shared-memory stage (1 or 2).  The MMA-issuing thread needs ~105 cycles per
//       tcgen05.mma it issues plus ~150 cycles of barrier wait / fence / commit per stage (tools/umma_bench.cu,
//       tools/trace_probe.py); with 4 MMAs (512 tensor cycles) per stage that is more than the stage holds, with
//       8 it is not.
// DENSE = 1: instantiation for the dense-logits API -- the [N, K+1] fp32 logits are written with full-line
//       coalesced stores (each 32 x 32 register block is transposed across the warp with shuffles first), not
//       with one 4-byte store per row per lane.  DENSE = 0 keeps a plain per-lane store for the rarely used
//       variant kernels and never pays registers for the transpose on the fused path.
template <int G, int CS, int EW, int KPS, int DENSE>
__global__ void __launch_bounds__(128 + EW * 32, 1)
nce_stats_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_queue,
                 const StatsArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kChunkBytes = (kStatsBN / G) * 128;         // one 64-wide K chunk of this CTA's B rows
    constexpr int kStageBytes = kChunkBytes * KPS;
    constexpr int kEpiWarps = EW;
    // EW == 8 : one epilogue group, every tile.   EW == 16: two groups of 8 warps in ping-pong -- group p owns
    // accumulator buffer p and drains the tiles of parity p, so the exps of tile t overlap the drain of t+1.
    constexpr int kEpiCols = 128;                          // accumulator columns per epilogue thread per tile
    constexpr int kEpiChunks = kEpiCols / 32;              // 32-column tcgen05.ld per thread per tile
    constexpr int kTileStep = (EW == 16) ? 2 : 1;
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
    static_assert(G == 1 || CS == 1, "multicast sharing is implemented for cta_group::1 only");
    constexpr int kCluster = G * CS;
    constexpr bool kClustered = kCluster > 1;
    constexpr uint16_t kMask = (uint16_t)((1u << CS) - 1u);
    const uint32_t crank = kClustered ? cluster_ctarank() : 0u;
    const uint32_t rank = (G == 2) ? crank : 0u;               // rank inside the MMA pair
    const int cluster_id = blockIdx.x / kCluster;
    const int mgroups = a.mblks / CS;                           // host guarantees CS | mblks
    const int mblk = (cluster_id % mgroups) * CS + ((CS > 1) ? (int)crank : 0);
    const int slice = cluster_id / mgroups;
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
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CS); }
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
        if (lane == 0) {
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
            uint8_t* dst = b_s + ((CS > 1) ? (size_t)crank * (kStatsBN / CS) * 128 : 0);
            const uint32_t full0 = (G == 2) ? mapa_shared(smem_u32(&full[0]), 0) : 0u;
            int brow = t0 * kStatsBN + (int)rank * (kStatsBN / G) + ((CS > 1) ? (int)crank * (kStatsBN / CS) : 0);
            for (int t = t0; t < t1; ++t, brow += kStatsBN) {
                for (int kc = 0; kc < kchunks; kc += KPS) {
                    mbar_wait(&empty[st], ph ^ 1u);
                    if (rank == 0) mbar_arrive_expect_tx(&full[st], (uint32_t)(kStageBytes * G));
#pragma unroll
                    for (int kk = 0; kk < KPS; ++kk) {
                        uint8_t* d = dst + kk * kChunkBytes;
                        if (G == 2)      tma_load_2d_2sm(&tm_queue, full0 + (uint32_t)st * 8u, d, (kc + kk) * 64, brow);
                        else if (CS > 1) tma_load_2d_mc(&tm_queue, &full[st], d, (kc + kk) * 64, brow, kMask);
                        else             tma_load_2d(&tm_queue, &full[st], d, (kc + kk) * 64, brow);
                    }
                    dst += kStageBytes;
                    if (++st == NS) { st = 0; ph ^= 1u; dst -= (size_t)NS * kStageBytes; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
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
                    if (CS > 1) umma_commit_mc(&empty[st], kMask); else umma_commit<G>(&empty[st]);
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
        int lt = (EW == 16) ? (cgrp >> 1) : 0;
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
// Kernel B: dq pass.  O[128, C] += P[128, 128] . Queue_tile[128, C] with P = 2^(S*scale2 - lse2)
// =====================================================================================
constexpr int kDqBN = 128;
constexpr int kDqThreads = 384;

struct DqArgs {
    int N, C, K;
    int mblks, slices, n_pad, num_tiles, stages;
    float inv_T;
    const float* lse;     // [N] natural log
    float* part_o;        // [slices, n_pad, C]
};

template <int CS>
__global__ void __launch_bounds__(kDqThreads, 1)
nce_dq_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_queue,
              const DqArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int kchunks = a.C >> 6;
    const int NS = a.stages;
    const int tile_bytes = kchunks * kSlab;          // [128 rows x C] as C/64 slabs
    uint8_t* q_s = smem;
    uint8_t* p_s = q_s + tile_bytes;                 // P tile: 2 slabs (128 rows x 128 cols bf16)
    uint8_t* v_s = p_s + 2 * kSlab;                  // NS queue tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + (size_t)NS * tile_bytes);
    uint64_t* kv_full = bars;
    uint64_t* kv_empty = bars + NS;
    uint64_t* s_full = bars + 2 * NS;        // [2]
    uint64_t* s_empty = bars + 2 * NS + 2;   // [2]
    uint64_t* p_full = bars + 2 * NS + 4;
    uint64_t* p_empty = bars + 2 * NS + 5;
    uint64_t* o_full = bars + 2 * NS + 6;
    uint64_t* qfull = bars + 2 * NS + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool kClustered = CS > 1;
    constexpr uint16_t kMask = (uint16_t)((1u << CS) - 1u);
    const uint32_t crank = kClustered ? cluster_ctarank() : 0u;
    const int cluster_id = blockIdx.x / CS;
    const int mgroups = a.mblks / CS;                           // host guarantees CS | mblks
    const int mblk = (cluster_id % mgroups) * CS + (int)crank;
    const int slice = cluster_id / mgroups;
    const int t0 = (int)(((long long)slice * a.num_tiles) / a.slices);
    const int t1 = (int)(((long long)(slice + 1) * a.num_tiles) / a.slices);
    const int ntiles = t1 - t0;
    const int row0 = mblk * kRowsPerCta;
    // TMEM columns: S buffers at [0,128) and [128,256); O at [256, 256 + C)
    constexpr uint32_t kOCol = 256;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_queue);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], CS); }
        for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], 8); }
        mbar_init(p_full, 8);
        mbar_init(p_empty, 1);
        mbar_init(o_full, 1);
        mbar_init(qfull, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc<1>(tmem_slot, 512);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    if (kClustered) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------ TMA producer
            mbar_arrive_expect_tx(qfull, (uint32_t)tile_bytes);
            for (int kc = 0; kc < kchunks; ++kc) tma_load_2d(&tm_q, qfull, q_s + kc * kSlab, kc * 64, row0);
            int st = 0;
            uint32_t ph = 0;
            for (int i = 0; i < ntiles; ++i, st = (st + 1 == NS) ? 0 : st + 1, ph ^= (st == 0) ? 1u : 0u) {
                mbar_wait(&kv_empty[st], ph ^ 1u);
                mbar_arrive_expect_tx(&kv_full[st], (uint32_t)tile_bytes);
                for (int kc = 0; kc < kchunks; ++kc) {
                    if (CS > 1) {
                        constexpr int kPart = kDqBN / CS;      // this CTA's rows of the tile, multicast to the cluster
                        tma_load_2d_mc(&tm_queue, &kv_full[st],
                                       v_s + (size_t)st * tile_bytes + kc * kSlab + (size_t)crank * kPart * 128, kc * 64,
                                       (t0 + i) * kDqBN + (int)crank * kPart, kMask);
                    } else {
                        tma_load_2d(&tm_queue, &kv_full[st], v_s + (size_t)st * tile_bytes + kc * kSlab, kc * 64,
                                    (t0 + i) * kDqBN);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------ MMA issuer
            const uint32_t idesc_s = make_idesc_bf16(128, kDqBN, 0, 0);            // S = q . tile^T  (both K-major)
            const uint32_t idesc_o = make_idesc_bf16(128, (uint32_t)a.C, 0, 1);    // O += P . tile   (B MN-major)
            mbar_wait(qfull, 0);
            tc_fence_after();
            // lean single-thread loops (see the note in nce_stats_kernel): descriptors are built once and
            // advanced by adds, stage / phase are running counters
            const uint64_t q_desc0 = make_sw128_desc(smem_u32(q_s), 0, 1024);
            const uint64_t p_desc0 = make_sw128_desc(smem_u32(p_s), 0, 1024);
            const uint64_t vk_desc0 = make_sw128_desc(smem_u32(v_s), 0, 1024);         // tile as K-major B (S MMA)
            const uint64_t vm_desc0 = make_sw128_desc(smem_u32(v_s), kSlab, 1024);     // tile as MN-major B (PV MMA)
            constexpr uint64_t kSlabUnits = (uint64_t)(kSlab >> 4);
            const uint64_t tile_units = (uint64_t)(tile_bytes >> 4);
            int s_st = 0; uint32_t s_ph = 0; uint64_t s_vdesc = vk_desc0;             // S side ring cursor
            int o_st = 0; uint64_t o_vdesc = vm_desc0;                                 // PV side ring cursor
            auto issue_s = [&](int i) {
                const uint32_t b = (uint32_t)i & 1u;
                mbar_wait(&kv_full[s_st], s_ph);
                mbar_wait(&s_empty[b], (((uint32_t)i >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d = tmem_base + b * (uint32_t)kDqBN;
                uint64_t qd = q_desc0, vd = s_vdesc;
                for (int kc = 0; kc < kchunks; ++kc) {
                    umma_ss<1>(d, qd, vd, idesc_s, (uint32_t)(kc != 0));
                    umma_ss<1>(d, qd + 2, vd + 2, idesc_s, 1u);
                    umma_ss<1>(d, qd + 4, vd + 4, idesc_s, 1u);
                    umma_ss<1>(d, qd + 6, vd + 6, idesc_s, 1u);
                    qd += kSlabUnits;
                    vd += kSlabUnits;
                }
                umma_commit<1>(&s_full[b]);
                s_vdesc += tile_units;
                if (++s_st == NS) { s_st = 0; s_ph ^= 1u; s_vdesc = vk_desc0; }
            };
            if (ntiles > 0) issue_s(0);
            for (int i = 0; i < ntiles; ++i) {
                if (i + 1 < ntiles) issue_s(i + 1);
                mbar_wait(p_full, (uint32_t)i & 1u);
                tc_fence_after();
                // A = P[:, 16kk .. 16kk+16) (K-major slab kk/4, 32-byte step kk%4)
                // B = tile rows [16kk, 16kk+16) x C (MN-major: 64-element chunks LBO = slab apart, 8-row groups
                //     SBO = 1024 B apart): 2048 B (= 128 descriptor units) per 16 rows
#pragma unroll
                for (int kk = 0; kk < kDqBN / 16; ++kk) {
                    umma_ss<1>(tmem_base + kOCol, p_desc0 + (uint64_t)((kk >> 2) * (kSlab >> 4) + (kk & 3) * 2),
                               o_vdesc + (uint64_t)(kk * 128), idesc_o, (uint32_t)((i | kk) != 0));
                }
                if (CS > 1) umma_commit_mc(&kv_empty[o_st], kMask); else umma_commit<1>(&kv_empty[o_st]);
                umma_commit<1>(p_empty);
                o_vdesc += tile_units;
                if (++o_st == NS) { o_st = 0; o_vdesc = vm_desc0; }
            }
            umma_commit<1>(o_full);
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------- softmax warps (8) + O epilogue
        const int quarter = warp & 3;
        const int half = (warp - 4) >> 2;
        const int row_local = quarter * 32 + lane;
        const int grow = row0 + row_local;
        const float scale2 = a.inv_T * kLog2e;
        const float lse2 = (grow < a.N) ? a.lse[grow] * kLog2e : 0.f;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        uint8_t* p_row = p_s + half * kSlab + row_local * 128;     // this thread's 128-byte P row (64 cols)
        const int sw = row_local & 7;
        for (int i = 0; i < ntiles; ++i) {
            const int b = i & 1;
            mbar_wait(&s_full[b], (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            uint32_t packed[32];       // 64 bf16
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                uint32_t r[32];
                tmem_ld32(lane_base + (uint32_t)(b * kDqBN + half * 64 + ch * 32), r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    float e0 = ex2(fmaf(__uint_as_float(r[j]), scale2, -lse2));
                    float e1 = ex2(fmaf(__uint_as_float(r[j + 1]), scale2, -lse2));
                    __nv_bfloat162 h = __floats2bfloat162_rn(e0, e1);
                    packed[ch * 16 + (j >> 1)] = *reinterpret_cast<uint32_t*>(&h);
                }
            }
            // S buffer b may now be overwritten by S(i+2)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[b]);
            // P smem is free once PV(i-1) has completed
            mbar_wait(p_empty, ((uint32_t)i & 1u) ^ 1u);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                uint4 v = make_uint4(packed[u * 4], packed[u * 4 + 1], packed[u * 4 + 2], packed[u * 4 + 3]);
                *reinterpret_cast<uint4*>(p_row + ((u ^ sw) << 4)) = v;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // O epilogue: each thread stores its row's C/2 columns of the slice partial
        mbar_wait(o_full, 0);
        tc_fence_after();
        const int ccols = a.C >> 1;
        float* orow = a.part_o + ((size_t)slice * a.n_pad + grow) * a.C + half * ccols;
        for (int c = 0; c < ccols; c += 32) {
            uint32_t r[32];
            if (ntiles > 0) {
                tmem_ld32(lane_base + kOCol + (uint32_t)(half * ccols + c), r);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<uint4*>(orow + c + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        }
    }

    __syncwarp();
    tc_fence_before();
    if (kClustered) cluster_sync_all(); else __syncthreads();
    if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

// =====================================================================================
// Host side
// =====================================================================================
cudaError_t launch_nce_tc(NceTcParams& p, const NceWorkspace& ws, cudaStream_t stream) {
    if (p.C % 64 != 0 || p.C < 64 || p.C > 256 || p.N < 1 || p.K < 1) return cudaErrorNotSupported;
    const int G = p.cta_group;
    const int kchunks = p.C / 64;
    const int mblks = (p.N + 128 * G - 1) / (128 * G);
    const int CS = (G == 1) ? pick_share(mblks, p.max_share) : 1;
    if (mblks * G > p.num_sms) return cudaErrorNotSupported;
    const int num_tiles = (p.K + kStatsBN - 1) / kStatsBN;
    const int n_pad = mblks * G * 128;
    p.n_pad = n_pad;

    CUtensorMap tm_q, tm_queue;
    if (!make_tmap(&tm_q, p.q_bf16, p.N, p.C, 128)) return cudaErrorUnknown;
    if (!make_tmap(&tm_queue, p.queue, p.K, p.C, kStatsBN / (G * CS))) return cudaErrorUnknown;

    const int KPS = (G == 2 && kchunks % 2 == 0 && p.epi_warps == 16 && !p.kps1) ? 2 : 1;
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
    static KernelCache kc[4];
    const int mgroups = mblks / CS, per_slice = mblks * G;
    static KernelCache kc16[4];
    if (G == 2 && KPS == 2)
        return plan_and_launch(nce_stats_kernel<2, 1, 16, 2, 0>, kc16[2], 128 + 16 * 32, smem, 2, mgroups, per_slice, num_tiles,
                               n_pad, &p.slices, stream, tm_q, tm_queue, a, fill, true);
    if (G == 2 && p.epi_warps == 16)
        return plan_and_launch(nce_stats_kernel<2, 1, 16, 1, 0>, kc16[0], 128 + 16 * 32, smem, 2, mgroups, per_slice, num_tiles,
                               n_pad, &p.slices, stream, tm_q, tm_queue, a, fill, true);
    if (G == 2)
        return plan_and_launch(nce_stats_kernel<2, 1, 8, 1, 0>, kc[0], 128 + 8 * 32, smem, 2, mgroups, per_slice, num_tiles,
                               n_pad, &p.slices, stream, tm_q, tm_queue, a, fill, true);
    if (CS == 4)
        return plan_and_launch(nce_stats_kernel<1, 4, 8, 1, 0>, kc[1], 128 + 8 * 32, smem, 4, mgroups, per_slice, num_tiles,
                               n_pad, &p.slices, stream, tm_q, tm_queue, a, fill, true);
    if (CS == 2)
        return plan_and_launch(nce_stats_kernel<1, 2, 8, 1, 0>, kc[2], 128 + 8 * 32, smem, 2, mgroups, per_slice, num_tiles,
                               n_pad, &p.slices, stream, tm_q, tm_queue, a, fill, true);
    static KernelCache kc_dense;
    if (p.epi_warps == 16 && p.logits != nullptr)
        return plan_and_launch(nce_stats_kernel<1, 1, 16, 1, 1>, kc_dense, 128 + 16 * 32, smem, 1, mgroups, per_slice,
                               num_tiles, n_pad, &p.slices, stream, tm_q, tm_queue, a, fill, true);
    if (p.epi_warps == 16)
        return plan_and_launch(nce_stats_kernel<1, 1, 16, 1, 0>, kc16[1], 128 + 16 * 32, smem, 1, mgroups, per_slice, num_tiles,
                               n_pad, &p.slices, stream, tm_q, tm_queue, a, fill, true);
    return plan_and_launch(nce_stats_kernel<1, 1, 8, 1, 0>, kc[3], 128 + 8 * 32, smem, 1, mgroups, per_slice, num_tiles, n_pad,
                           &p.slices, stream, tm_q, tm_queue, a, fill, true);
}

cudaError_t launch_nce_dq_tc(const __nv_bfloat16* q_bf16, const __nv_bfloat16* queue, int N, int C, int K,
                             float inv_T, const float* lse, int num_sms, int max_share, int* slices_out,
                             int* n_pad_out, const NceWorkspace& ws, cudaStream_t stream) {
    if (C % 64 != 0 || C < 64 || C > 256) return cudaErrorNotSupported;
    const int kchunks = C / 64;
    const int mblks = (N + 127) / 128;
    const int CS = pick_share(mblks, max_share);
    if (mblks > num_sms) return cudaErrorNotSupported;
    const int num_tiles = (K + kDqBN - 1) / kDqBN;
    const int n_pad = mblks * 128;
    *n_pad_out = n_pad;

    CUtensorMap tm_q, tm_queue;
    if (!make_tmap(&tm_q, q_bf16, N, C, 128)) return cudaErrorUnknown;
    if (!make_tmap(&tm_queue, queue, K, C, kDqBN / CS)) return cudaErrorUnknown;

    const int tile_bytes = kchunks * kSlab;
    const int fixed = tile_bytes + 2 * kSlab + 1024;      // q + P + barriers
    int stages = (kSmemBudget - fixed) / tile_bytes;
    if (stages > 4) stages = 4;
    if (stages < 2) return cudaErrorNotSupported;
    const int smem = fixed + stages * tile_bytes + 1024;

    DqArgs a;
    a.N = N; a.C = C; a.K = K;
    a.mblks = mblks; a.slices = 0; a.n_pad = n_pad; a.num_tiles = num_tiles; a.stages = stages;
    a.inv_T = inv_T;
    a.lse = lse;
    a.part_o = ws.part_o;
    auto fill = [](DqArgs& x, int slices) { x.slices = slices; };
    static KernelCache kc[3];
    const int mgroups = mblks / CS;
    if (CS == 4)
        return plan_and_launch(nce_dq_kernel<4>, kc[0], kDqThreads, smem, 4, mgroups, mblks, num_tiles, n_pad,
                               slices_out, stream, tm_q, tm_queue, a, fill);
    if (CS == 2)
        return plan_and_launch(nce_dq_kernel<2>, kc[1], kDqThreads, smem, 2, mgroups, mblks, num_tiles, n_pad,
                               slices_out, stream, tm_q, tm_queue, a, fill);
    return plan_and_launch(nce_dq_kernel<1>, kc[2], kDqThreads, smem, 1, mgroups, mblks, num_tiles, n_pad, slices_out,
                           stream, tm_q, tm_queue, a, fill);
}

}  // namespace moco
