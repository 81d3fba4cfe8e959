// q . Queue^T statistics kernel, "TS" variant: the q block lives in TMEM as the A operand of every tcgen05.mma
// (staged once by the epilogue warps: global -> registers -> tcgen05.st), so the whole shared memory is a
// 9-stage TMA ring of 192-row queue tiles.
//
// Status: selectable alternative (MOCO_NCE_STATS_TS), not the default.  It was built on the hypothesis that the
// 4-stage ring of the SS kernel starved the tensor pipe; the measurements (profiles/README.md) showed otherwise
// -- the deep ring did not help and the SS kernel with lean issue loops + ping-pong epilogue groups is faster
// (56-58 us vs 69.5 us at N=512, C=256, K=262144).  It stays because it is parity-green and is the natural
// starting point for a kernel that also keeps P in TMEM (see nce_dq2_sm100.cu).
//
// TMEM map (512 columns): q [0, C/2) (bf16 pairs, <= 128 cols) | accumulator 0 [128, 320) | accumulator 1 [320, 512)
// Tile = 192 queue rows (UMMA 128 x 192 x 16); stage = one 64-wide K chunk of a tile (192 rows x 128 B).
//
// Replaces torch.mm + cat + div + CrossEntropyLoss + softmax of the reference
// (moco/NCE/Contrast.py:25-27, NCECriterion.py:11-13, train.py:264).
#include <cuda.h>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tc_common.cuh"

namespace moco {

constexpr int kS3BN = 192;
constexpr int kS3StageBytes = kS3BN * 128;
constexpr uint32_t kS3QCol = 0, kS3AccCol = 128;

struct Stats3Args {
    int N, C, K;
    int mblks, slices, n_pad, num_tiles, stages;
    float inv_T;
    const __nv_bfloat16* q;   // [N, C]
    float* logits;            // optional [N, K+1]
    float2* part_ms;          // [slices, n_pad]
};

// EW = number of epilogue warps (8 or 16)
template <int EW>
__global__ void __launch_bounds__(128 + EW * 32, 1)
nce_stats3_kernel(const __grid_constant__ CUtensorMap tm_queue, const __grid_constant__ CUtensorMap tm_unused,
                  const Stats3Args a) {
    constexpr int kGroups = EW / 4;                 // column groups
    constexpr int kCols = kS3BN / kGroups;          // accumulator columns per thread per tile (96 or 48)
    constexpr int kChunks = kCols / 16;             // 16-column tcgen05.ld per thread per tile
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int kchunks = a.C >> 6;
    const int NS = a.stages;
    uint8_t* b_s = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_s + (size_t)NS * kS3StageBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + NS;
    uint64_t* tfull = bars + 2 * NS;
    uint64_t* tempty = bars + 2 * NS + 2;
    uint64_t* q_ready = bars + 2 * NS + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 5);
    float2* red_s = reinterpret_cast<float2*>(bars + 2 * NS + 6);   // [kGroups - 1][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mblk = blockIdx.x % a.mblks;
    const int slice = blockIdx.x / a.mblks;
    const int t0 = (int)(((long long)slice * a.num_tiles) / a.slices);
    const int t1 = (int)(((long long)(slice + 1) * a.num_tiles) / a.slices);
    const int row0 = mblk * kRowsPerCta;

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_queue);
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], EW); }
        mbar_init(q_ready, 4);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc<1>(tmem_slot, 512);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------ TMA producer (queue only; starts immediately)
            int st = 0;
            uint32_t ph = 0;
            uint8_t* dst = b_s;
            int brow = t0 * kS3BN;
            for (int t = t0; t < t1; ++t, brow += kS3BN) {
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&empty[st], ph ^ 1u);
                    mbar_arrive_expect_tx(&full[st], (uint32_t)kS3StageBytes);
                    tma_load_2d(&tm_queue, &full[st], dst, kc * 64, brow);
                    dst += kS3StageBytes;
                    if (++st == NS) { st = 0; ph ^= 1u; dst = b_s; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------ MMA issuer
            const uint32_t idesc = make_idesc_bf16(128, kS3BN, 0, 0);
            mbar_wait(q_ready, 0);
            tc_fence_after();
            const uint64_t b_desc0 = make_sw128_desc(smem_u32(b_s), 0, 1024);
            constexpr uint64_t kStageUnits = (uint64_t)(kS3StageBytes >> 4);
            int st = 0;
            uint32_t ph = 0, acc = 0, aph = 0;
            uint64_t b_desc = b_desc0;
            for (int t = t0; t < t1; ++t) {
                mbar_wait(&tempty[acc], aph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + kS3AccCol + acc * (uint32_t)kS3BN;
                uint32_t qa = tmem_base + kS3QCol;
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&full[st], ph);
                    tc_fence_after();
                    umma_ts<1>(d_tmem, qa, b_desc, idesc, (uint32_t)(kc != 0));
                    umma_ts<1>(d_tmem, qa + 8, b_desc + 2, idesc, 1u);
                    umma_ts<1>(d_tmem, qa + 16, b_desc + 4, idesc, 1u);
                    umma_ts<1>(d_tmem, qa + 24, b_desc + 6, idesc, 1u);
                    umma_commit<1>(&empty[st]);
                    qa += 32;
                    b_desc += kStageUnits;
                    if (++st == NS) { st = 0; ph ^= 1u; b_desc = b_desc0; }
                }
                umma_commit<1>(&tfull[acc]);
                acc ^= 1u;
                aph ^= (acc == 0u) ? 1u : 0u;
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------- epilogue warps
        const int quarter = warp & 3;
        const int cgrp = (warp - 4) >> 2;
        const int row_local = quarter * 32 + lane;
        const int grow = row0 + row_local;
        const float scale2 = a.inv_T * kLog2e;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        if (cgrp == 0) {
            // stage this thread's q row into TMEM: lane = row, one 32-bit column = two consecutive bf16 of K
            const uint4* src = reinterpret_cast<const uint4*>(a.q + (size_t)(grow < a.N ? grow : 0) * a.C);
            for (int c = 0; c < a.C; c += 64) {
                uint32_t r[32];
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    uint4 u = (grow < a.N) ? __ldg(src + (c >> 3) + v) : make_uint4(0u, 0u, 0u, 0u);
                    r[v * 4 + 0] = u.x; r[v * 4 + 1] = u.y; r[v * 4 + 2] = u.z; r[v * 4 + 3] = u.w;
                }
                tmem_st32(lane_base + kS3QCol + (uint32_t)(c >> 1), r);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_ready);
        }
        float m = -INFINITY, s = 0.f;
        float* lrow = (a.logits != nullptr && grow < a.N) ? a.logits + (size_t)grow * (a.K + 1) + 1 : nullptr;
        auto fold = [&](const uint32_t (&r)[16], int col0) {
            const int valid = a.K - col0;
            if (valid >= 16) {
                float c0 = fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1]));
                float c1 = fmaxf(__uint_as_float(r[2]), __uint_as_float(r[3]));
#pragma unroll
                for (int j = 4; j < 16; j += 4) {
                    c0 = fmaxf(c0, fmaxf(__uint_as_float(r[j + 0]), __uint_as_float(r[j + 1])));
                    c1 = fmaxf(c1, fmaxf(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
                }
                const float cm = fmaxf(c0, c1) * scale2;
                if (cm > m) { s *= ex2(m - cm); m = cm; }
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    s0 += ex2(fmaf(__uint_as_float(r[j + 0]), scale2, -m));
                    s1 += ex2(fmaf(__uint_as_float(r[j + 1]), scale2, -m));
                    s2 += ex2(fmaf(__uint_as_float(r[j + 2]), scale2, -m));
                    s3 += ex2(fmaf(__uint_as_float(r[j + 3]), scale2, -m));
                }
                s += (s0 + s1) + (s2 + s3);
                if (lrow) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) lrow[col0 + j] = __uint_as_float(r[j]) * a.inv_T;
                }
            } else if (valid > 0) {
                float cm = -INFINITY;
#pragma unroll
                for (int j = 0; j < 16; ++j) if (j < valid) cm = fmaxf(cm, __uint_as_float(r[j]));
                cm *= scale2;
                if (cm > m) { s *= ex2(m - cm); m = cm; }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (j < valid) {
                        s += ex2(fmaf(__uint_as_float(r[j]), scale2, -m));
                        if (lrow) lrow[col0 + j] = __uint_as_float(r[j]) * a.inv_T;
                    }
                }
            }
        };
        int lt = 0;
        for (int t = t0; t < t1; ++t, ++lt) {
            const int acc = lt & 1;
            mbar_wait(&tfull[acc], (uint32_t)(lt >> 1) & 1u);
            tc_fence_after();
            const int col = cgrp * kCols;
            const uint32_t taddr = lane_base + kS3AccCol + (uint32_t)(acc * kS3BN + col);
            // pull the whole slice into registers first, hand the buffer back, then do the exps
            uint32_t r[kChunks][16];
#pragma unroll
            for (int ch = 0; ch < kChunks; ++ch) tmem_ld16(taddr + (uint32_t)(ch * 16), r[ch]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
#pragma unroll
            for (int ch = 0; ch < kChunks; ++ch) fold(r[ch], t * kS3BN + col + ch * 16);
        }
        if (cgrp > 0) red_s[(cgrp - 1) * kRowsPerCta + row_local] = make_float2(m, s);
        named_bar_sync(1, EW * 32);
        if (cgrp == 0) {
            float M = m;
#pragma unroll
            for (int gq = 0; gq < kGroups - 1; ++gq) M = fmaxf(M, red_s[gq * kRowsPerCta + row_local].x);
            float S = (m != -INFINITY) ? s * ex2(m - M) : 0.f;
#pragma unroll
            for (int gq = 0; gq < kGroups - 1; ++gq) {
                float2 o = red_s[gq * kRowsPerCta + row_local];
                if (o.x != -INFINITY) S += o.y * ex2(o.x - M);
            }
            a.part_ms[(size_t)slice * a.n_pad + grow] = make_float2(M, S);
        }
    }

    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

cudaError_t launch_nce_stats3(NceTcParams& p, int epi_warps, const NceWorkspace& ws, cudaStream_t stream) {
    if (p.C % 64 != 0 || p.C < 64 || p.C > 256 || p.N < 1 || p.K < 1) return cudaErrorNotSupported;
    const int mblks = (p.N + 127) / 128;
    if (mblks > p.num_sms) return cudaErrorNotSupported;
    const int num_tiles = (p.K + kS3BN - 1) / kS3BN;
    const int n_pad = mblks * 128;
    p.n_pad = n_pad;

    CUtensorMap tm_queue;
    if (!make_tmap(&tm_queue, p.queue, p.K, p.C, kS3BN)) return cudaErrorUnknown;

    const int fixed = 4096;     // barriers + red_s
    int stages = (kSmemBudget - fixed) / kS3StageBytes;
    if (stages > 12) stages = 12;
    const int smem = fixed + stages * kS3StageBytes + 1024;

    Stats3Args a;
    a.N = p.N; a.C = p.C; a.K = p.K;
    a.mblks = mblks; a.slices = 0; a.n_pad = n_pad; a.num_tiles = num_tiles; a.stages = stages;
    a.inv_T = p.inv_T;
    a.q = p.q_bf16;
    a.logits = p.logits;
    a.part_ms = ws.part_ms;
    auto fill = [](Stats3Args& x, int slices) { x.slices = slices; };
    static KernelCache kc[2];
    if (epi_warps == 16)
        return plan_and_launch(nce_stats3_kernel<16>, kc[0], 128 + 16 * 32, smem, 1, mblks, mblks, num_tiles, n_pad,
                               &p.slices, stream, tm_queue, tm_queue, a, fill);
    return plan_and_launch(nce_stats3_kernel<8>, kc[1], 128 + 8 * 32, smem, 1, mblks, mblks, num_tiles, n_pad,
                           &p.slices, stream, tm_queue, tm_queue, a, fill);
}

}  // namespace moco
