// dq pass, second generation ("TS" form): both A operands live in TMEM.
//
//   q[128, C]  is loaded ONCE by the softmax warps (global -> registers -> tcgen05.st) and stays in TMEM as
//              the A operand of every S = q . tile^T MMA: no smem footprint and no smem read bandwidth for q.
//   P[128, 64] = 2^(S*log2e/T - lse) is written back into the TMEM columns its S tile came from (bf16 pairs,
//              tcgen05.st) and is the A operand of O += P . tile (tile re-used from smem as MN-major B).
//   O[128, C]  accumulates in TMEM for the whole queue slice.
//
// TMEM map (512 columns): q [0, C/2) | O [128, 128 + C) | S/P buffers 0 and 1 in the top 2*BN columns.
// BN (queue rows per tile) is 128 when C <= 128 and 64 when C = 192/256 (O then needs up to 256 columns);
// tcgen05.mma with N = 64 only reaches ~48 % of peak (tools/umma_bench.cu), so the wider tile matters.
// Queue tiles are BN rows x C (C/64 slabs of BN x 128 B, 128B swizzle), 4-8 stage TMA ring; optional
// TMA-multicast sharing across a cluster of CS CTAs that own different q row blocks.
// tcgen05 MMAs issued by one thread execute in order, so S(i+2) overwriting the buffer P(i) was read from
// needs no barrier: the issue order S(i+1), PV(i), S(i+2), PV(i+1), ... is the dependency.
//
// Replaces autograd's backward GEMM (train.py:273 of bl0/moco) and the queue clone it needs
// (moco/NCE/Contrast.py:24-25).
#include <cuda.h>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tc_common.cuh"

namespace moco {

constexpr int kDq2Threads = 384;          // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps 4-11 softmax
constexpr uint32_t kQCol = 0, kOCol2 = 128;

struct Dq2Args {
    int N, C, K;
    int mblks, slices, n_pad, num_tiles, stages;
    float inv_T;
    const __nv_bfloat16* q;   // [N, C]
    const float* lse;         // [N] natural log
    float* part_o;            // [slices, n_pad, C]
    int debug;                // bring-up only: 1 = no exps, 2 = no MMA issue, 4 = no TMA
};

template <int CS, int BN>
__global__ void __launch_bounds__(kDq2Threads, 1)
nce_dq2_kernel(const __grid_constant__ CUtensorMap tm_queue, const __grid_constant__ CUtensorMap tm_unused,
               const Dq2Args a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int kchunks = a.C >> 6;
    const int NS = a.stages;
    constexpr int kDq2BN = BN;
    constexpr uint32_t kSCol = 512 - 2 * BN;
    constexpr int kSlab64 = kDq2BN * 128;                 // one [BN rows x 64 bf16] slab
    const int tile_bytes = kchunks * kSlab64;
    uint8_t* v_s = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + (size_t)NS * tile_bytes);
    uint64_t* kv_full = bars;
    uint64_t* kv_empty = bars + NS;
    uint64_t* s_full = bars + 2 * NS;        // [2]
    uint64_t* p_full = bars + 2 * NS + 2;    // [2]
    uint64_t* o_full = bars + 2 * NS + 4;
    uint64_t* q_ready = bars + 2 * NS + 5;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool kClustered = CS > 1;
    constexpr uint16_t kMask = (uint16_t)((1u << CS) - 1u);
    const uint32_t crank = kClustered ? cluster_ctarank() : 0u;
    const int cluster_id = blockIdx.x / CS;
    const int mgroups = a.mblks / CS;
    const int mblk = (cluster_id % mgroups) * CS + (int)crank;
    const int slice = cluster_id / mgroups;
    const int t0 = (int)(((long long)slice * a.num_tiles) / a.slices);
    const int t1 = (int)(((long long)(slice + 1) * a.num_tiles) / a.slices);
    const int ntiles = t1 - t0;
    const int row0 = mblk * kRowsPerCta;

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_queue);
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], CS); }
        for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&p_full[b], 8); }
        mbar_init(o_full, 1);
        mbar_init(q_ready, 4);
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
            // ------------------------------------------------ TMA producer (queue tiles only)
            int st = 0;
            uint32_t ph = 0;
            for (int i = 0; i < ntiles; ++i, st = (st + 1 == NS) ? 0 : st + 1, ph ^= (st == 0) ? 1u : 0u) {
                mbar_wait(&kv_empty[st], ph ^ 1u);
                if (a.debug & 4) { mbar_arrive(&kv_full[st]); continue; }
                mbar_arrive_expect_tx(&kv_full[st], (uint32_t)tile_bytes);
                for (int kc = 0; kc < kchunks; ++kc) {
                    uint8_t* dst = v_s + (size_t)st * tile_bytes + kc * kSlab64;
                    if (CS > 1) {
                        constexpr int kPart = kDq2BN / CS;
                        tma_load_2d_mc(&tm_queue, &kv_full[st], dst + (size_t)crank * kPart * 128, kc * 64,
                                       (t0 + i) * kDq2BN + (int)crank * kPart, kMask);
                    } else {
                        tma_load_2d(&tm_queue, &kv_full[st], dst, kc * 64, (t0 + i) * kDq2BN);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------ MMA issuer
            const uint32_t idesc_s = make_idesc_bf16(128, kDq2BN, 0, 0);            // S = q . tile^T   (A: TMEM, B: K-major)
            const uint32_t idesc_o = make_idesc_bf16(128, (uint32_t)a.C, 0, 1);     // O += P . tile    (A: TMEM, B: MN-major)
            mbar_wait(q_ready, 0);
            tc_fence_after();
            // lean single-thread loops: descriptors built once and advanced by adds, running stage / phase counters
            const uint64_t vk_desc0 = make_sw128_desc(smem_u32(v_s), 0, 1024);          // tile as K-major B (S MMA)
            const uint64_t vm_desc0 = make_sw128_desc(smem_u32(v_s), kSlab64, 1024);    // tile as MN-major B (PV MMA)
            constexpr uint64_t kSlabUnits = (uint64_t)(kSlab64 >> 4);
            const uint64_t tile_units = (uint64_t)(tile_bytes >> 4);
            int s_st = 0; uint32_t s_ph = 0; uint64_t s_vdesc = vk_desc0;
            int o_st = 0; uint64_t o_vdesc = vm_desc0;
            auto issue_s = [&](int i) {
                const uint32_t b = (uint32_t)i & 1u;
                mbar_wait(&kv_full[s_st], s_ph);
                tc_fence_after();
                const uint32_t d = tmem_base + kSCol + b * (uint32_t)kDq2BN;
                uint32_t qa = tmem_base + kQCol;
                uint64_t vd = s_vdesc;
                for (int kc = 0; kc < kchunks; ++kc) {
                    if (a.debug & 2) break;
                    umma_ts<1>(d, qa, vd, idesc_s, (uint32_t)(kc != 0));
                    umma_ts<1>(d, qa + 8, vd + 2, idesc_s, 1u);
                    umma_ts<1>(d, qa + 16, vd + 4, idesc_s, 1u);
                    umma_ts<1>(d, qa + 24, vd + 6, idesc_s, 1u);
                    qa += 32;
                    vd += kSlabUnits;
                }
                umma_commit<1>(&s_full[b]);
                s_vdesc += tile_units;
                if (++s_st == NS) { s_st = 0; s_ph ^= 1u; s_vdesc = vk_desc0; }
            };
            if (ntiles > 0) issue_s(0);
            if (ntiles > 1) issue_s(1);
            for (int i = 0; i < ntiles; ++i) {
                const uint32_t b = (uint32_t)i & 1u;
                mbar_wait(&p_full[b], ((uint32_t)i >> 1) & 1u);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < kDq2BN / 16; ++kk) {
                    if (a.debug & 2) break;
                    // A = P[:, 16kk..16kk+16) : 8 packed TMEM columns;  B = tile rows [16kk, 16kk+16) x C (MN-major)
                    umma_ts<1>(tmem_base + kOCol2, tmem_base + kSCol + b * (uint32_t)kDq2BN + (uint32_t)(kk * 8),
                               o_vdesc + (uint64_t)(kk * 128), idesc_o, (uint32_t)((i | kk) != 0));
                }
                if (CS > 1) umma_commit_mc(&kv_empty[o_st], kMask); else umma_commit<1>(&kv_empty[o_st]);
                o_vdesc += tile_units;
                if (++o_st == NS) { o_st = 0; o_vdesc = vm_desc0; }
                if (i + 2 < ntiles) issue_s(i + 2);      // overwrites buffer b: ordered after PV(i) by the pipe
            }
            umma_commit<1>(o_full);
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------- softmax warps (8) + q staging + O epilogue
        const int quarter = warp & 3;
        const int chalf = (warp - 4) >> 2;
        const int row_local = quarter * 32 + lane;
        const int grow = row0 + row_local;
        const float scale2 = a.inv_T * kLog2e;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        if (chalf == 0) {
            // q row -> TMEM (A operand layout: lane = row, one 32-bit column = two consecutive bf16 of K)
            const uint4* src = reinterpret_cast<const uint4*>(a.q + (size_t)(grow < a.N ? grow : 0) * a.C);
            for (int c = 0; c < a.C; c += 64) {
                uint32_t r[32];
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    uint4 u = (grow < a.N) ? __ldg(src + (c >> 3) + v) : make_uint4(0u, 0u, 0u, 0u);
                    r[v * 4 + 0] = u.x; r[v * 4 + 1] = u.y; r[v * 4 + 2] = u.z; r[v * 4 + 3] = u.w;
                }
                tmem_st32(lane_base + kQCol + (uint32_t)(c >> 1), r);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_ready);
        }
        const float lse2 = (grow < a.N) ? a.lse[grow] * kLog2e : 0.f;
        for (int i = 0; i < ntiles; ++i) {
            const int b = i & 1;
            mbar_wait(&s_full[b], (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            constexpr int kHalf = BN / 2;                 // S columns per thread (32 or 64)
            uint32_t r[kHalf / 32][32];
#pragma unroll
            for (int h = 0; h < kHalf / 32; ++h)
                tmem_ld32(lane_base + kSCol + (uint32_t)(b * kDq2BN + chalf * kHalf + h * 32), r[h]);
            tmem_ld_wait();
            // both column halves of this lane quarter must have read S before either overwrites it with P
            named_bar_sync(2 + quarter, 64);
#pragma unroll
            for (int h = 0; h < kHalf / 32; ++h) {
                uint32_t p[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    if (a.debug & 1) { p[j >> 1] = r[h][j]; continue; }
                    float e0 = ex2(fmaf(__uint_as_float(r[h][j]), scale2, -lse2));
                    float e1 = ex2(fmaf(__uint_as_float(r[h][j + 1]), scale2, -lse2));
                    __nv_bfloat162 hh = __floats2bfloat162_rn(e0, e1);
                    p[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
                }
                tmem_st16(lane_base + kSCol + (uint32_t)(b * kDq2BN + chalf * (kHalf / 2) + h * 16), p);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[b]);
        }
        // O epilogue
        mbar_wait(o_full, 0);
        tc_fence_after();
        const int ccols = a.C >> 1;
        float* orow = a.part_o + ((size_t)slice * a.n_pad + grow) * a.C + chalf * ccols;
        for (int c = 0; c < ccols; c += 32) {
            uint32_t r[32];
            tmem_ld32(lane_base + kOCol2 + (uint32_t)(chalf * ccols + c), r);
            tmem_ld_wait();
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

cudaError_t launch_nce_dq2_tc(const __nv_bfloat16* q_bf16, const __nv_bfloat16* queue, int N, int C, int K,
                              float inv_T, const float* lse, int num_sms, int max_share, int* slices_out,
                              int* n_pad_out, const NceWorkspace& ws, cudaStream_t stream) {
    if (C % 64 != 0 || C < 64 || C > 256) return cudaErrorNotSupported;
    const int kchunks = C / 64;
    const int mblks = (N + 127) / 128;
    const int CS = pick_share(mblks, max_share);
    if (mblks > num_sms) return cudaErrorNotSupported;
    const int BN = (C <= 128) ? 128 : 64;
    const int num_tiles = (K + BN - 1) / BN;
    const int n_pad = mblks * 128;
    *n_pad_out = n_pad;

    CUtensorMap tm_queue;
    if (!make_tmap(&tm_queue, queue, K, C, BN / CS)) return cudaErrorUnknown;

    const int tile_bytes = kchunks * BN * 128;
    int stages = (kSmemBudget - 1024) / tile_bytes;
    if (stages > 8) stages = 8;
    if (stages < 2) return cudaErrorNotSupported;
    const int smem = stages * tile_bytes + 1024 + 1024;

    Dq2Args a;
    a.N = N; a.C = C; a.K = K;
    a.mblks = mblks; a.slices = 0; a.n_pad = n_pad; a.num_tiles = num_tiles; a.stages = stages;
    a.inv_T = inv_T;
    a.q = q_bf16;
    a.lse = lse;
    a.part_o = ws.part_o;
    a.debug = debug_mode();
    auto fill = [](Dq2Args& x, int slices) { x.slices = slices; };
    static KernelCache kc[6];
    const int mgroups = mblks / CS;
#define MOCO_DQ2_LAUNCH(CS_, BN_, IDX)                                                                             \
    return plan_and_launch(nce_dq2_kernel<CS_, BN_>, kc[IDX], kDq2Threads, smem, CS_, mgroups, mblks, num_tiles,   \
                           n_pad, slices_out, stream, tm_queue, tm_queue, a, fill)
    if (BN == 128) {
        if (CS == 4) MOCO_DQ2_LAUNCH(4, 128, 0);
        if (CS == 2) MOCO_DQ2_LAUNCH(2, 128, 1);
        MOCO_DQ2_LAUNCH(1, 128, 2);
    }
    if (CS == 4) MOCO_DQ2_LAUNCH(4, 64, 3);
    if (CS == 2) MOCO_DQ2_LAUNCH(2, 64, 4);
    MOCO_DQ2_LAUNCH(1, 64, 5);
#undef MOCO_DQ2_LAUNCH
}

}  // namespace moco
