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
// Queue tiles are BN rows x C (C/64 slabs of BN x 128 B, 128B swizzle), 4-8 stage TMA ring.
// Two MMA-issuing threads (S and P.V), see the kernel body; S(i+2) overwriting the buffer PV(i) reads P from is
// ordered by the s_free mbarrier that PV(i)'s tcgen05.commit arrives on.
//
// Replaces autograd's backward GEMM (train.py:273 of bl0/moco) and the queue clone it needs
// (moco/NCE/Contrast.py:24-25).
//
// ONE-PASS mode (template FUSED = true): the same kernel also produces the softmax statistics, so the separate
// statistics pass (nce_sm100.cu) is not run at all: 4NCK FLOP and NK exponentials for loss + gradient instead
// of 6NCK and 2NK.  There is no lse to normalise with yet, so each (CTA, row) stabilises with the row maximum
// of the CTA's FIRST tile, m, and keeps it:  P~ = 2^(x - m), O~ = sum_j P~_j queue_j, l = sum_j P~_j.  The
// partial (m, l) pairs have exactly the format the statistics kernel writes, so combine_kernel merges them
// unchanged, and dq_reduce rescales each slice's O~ by 2^(m_slice - lse).  fp32 has 127 binades of headroom:
// this is exact unless a later logit exceeds the first tile's maximum by more than ~88 nats, in which case the
// row's sum overflows to inf and the loss is inf/NaN -- loud, never silently wrong.  moco_nce_fwd therefore
// selects this mode only when 1/T is small enough that L2-normalised features cannot get there (capi.cu).
#include <cuda.h>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tc_common.cuh"

namespace moco {

#ifdef MOCO_TRACE
// Lab-only timeline trace (tools/trace_probe.py builds a separate library with -DMOCO_TRACE; never in the product
// build): clock64 stamps of CTA 0's MMA thread (role 0) and one softmax thread per tile group (roles 1, 2).
__device__ long long g_dq2_trace[4][64][8];   // role 3, row 0: kernel-level milestones
#define MOCO_TR(role, tile, slot) do { if (blockIdx.x == 0 && (tile) < 64) g_dq2_trace[role][tile][slot] = clock64(); } while (0)
#else
#define MOCO_TR(role, tile, slot) do { } while (0)
#endif

constexpr int kDq2Threads = 640;          // warp0 TMA, warp1 S-MMA, warp2 TMEM alloc, warp3 PV-MMA, warps 4-19 softmax
constexpr uint32_t kQCol = 0, kOCol2 = 128;

struct Dq2Args {
    int N, C, K;
    int mblks, slices, n_pad, num_tiles, stages;
    float inv_T;
    const __nv_bfloat16* q;   // [N, C]
    const float* lse;         // [N] natural log (two-pass mode)
    float* part_o;            // [slices, n_pad, C]
    float2* part_ms;          // [slices, n_pad] (stabiliser, sum) in the log2 domain (one-pass mode)
};

template <int BN, bool FUSED, bool ISS2>
__global__ void __launch_bounds__(kDq2Threads, 1)
nce_dq2_kernel(const __grid_constant__ CUtensorMap tm_queue, const __grid_constant__ CUtensorMap tm_unused,
               const Dq2Args a) {
    // The __align__(1024) makes the dynamic window start 1 KB aligned (that is the kernel's 1 KB of "static" smem),
    // so no alignment slack is budgeted; a misaligned base traps instead of running off the end.
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
    if (threadIdx.x == 0) MOCO_TR(3, 0, 0);                       // kernel entry
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
    uint64_t* s_free = bars + 2 * NS + 6;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 8);
    // one-pass mode: [3][128] floats the four softmax threads of a row exchange through (tile-0 maxima of the two
    // column halves first, the three non-leading partial sums at the end); 256 + 1536 B <= the 2 KB after the tiles
    float* exch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mblk = blockIdx.x % a.mblks;
    const int slice = blockIdx.x / a.mblks;
    const int t0 = (int)(((long long)slice * a.num_tiles) / a.slices);
    const int t1 = (int)(((long long)(slice + 1) * a.num_tiles) / a.slices);
    const int ntiles = t1 - t0;
    const int row0 = mblk * kRowsPerCta;

    pdl_launch_dependents();
    // ---- set-up that touches no global memory (overlaps the predecessor kernel under PDL) ----
    if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_queue);     // kernel parameter space, not global data
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&p_full[b], 8); mbar_init(&s_free[b], 1); }
        mbar_init(o_full, 1);
        mbar_init(q_ready, 4);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc<1>(tmem_slot, 512);
        tmem_relinquish<1>();
    }
    pdl_wait();                                                   // predecessor complete: q / lse / the queue are final
    // q staging warps (4-7): put the global loads of the first 128 columns of their q row in flight before the
    // set-up barrier -- at small K (two tiles per CTA) this load's latency was 1.6 us of the kernel's critical path
    uint4 qpre[16];
    const int qpre_n = (a.C < 128 ? a.C : 128) >> 3;
    if (warp >= 4 && warp < 8) {
        const int grow_q = row0 + (warp & 3) * 32 + lane;
        const uint4* src = reinterpret_cast<const uint4*>(a.q + (size_t)(grow_q < a.N ? grow_q : 0) * a.C);
        // unconditional loads from a clamped row (nothing may consume them before the barrier); rows >= N are
        // zeroed when they are stored to TMEM
#pragma unroll
        for (int v = 0; v < 16; ++v)
            if (v < qpre_n) qpre[v] = __ldg(src + v);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) MOCO_TR(3, 0, 1);                       // barriers initialised, TMEM allocated

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------ TMA producer (queue tiles only)
            int st = 0;
            uint32_t ph = 0;
            for (int i = 0; i < ntiles; ++i, st = (st + 1 == NS) ? 0 : st + 1, ph ^= (st == 0) ? 1u : 0u) {
                mbar_wait(&kv_empty[st], ph ^ 1u);
                mbar_arrive_expect_tx(&kv_full[st], (uint32_t)tile_bytes);
                for (int kc = 0; kc < kchunks; ++kc) {
                    uint8_t* dst = v_s + (size_t)st * tile_bytes + kc * kSlab64;
                    tma_load_2d(&tm_queue, &kv_full[st], dst, kc * 64, (t0 + i) * kDq2BN);
                }
            }
        }
    } else if (warp == 1 && !ISS2) {
        if (lane == 0) {
            // ------------------------------------------------ single MMA issuer (C <= 128: latency-chain bound, where
            // the extra s_free hand-off of the two-issuer form costs more than the issue overlap gains; measured)
            const uint32_t idesc_s = make_idesc_bf16(128, kDq2BN, 0, 0);            // S = q . tile^T   (A: TMEM, B: K-major)
            const uint32_t idesc_o = make_idesc_bf16(128, (uint32_t)a.C, 0, 1);     // O += P . tile    (A: TMEM, B: MN-major)
            mbar_wait(q_ready, 0);
            tc_fence_after();
            const uint64_t vk_desc0 = make_sw128_desc(smem_u32(v_s), 0, 1024);          // tile as K-major B (S MMA)
            const uint64_t vm_desc0 = make_sw128_desc(smem_u32(v_s), kSlab64, 1024);    // tile as MN-major B (PV MMA)
            constexpr uint64_t kSlabUnits = (uint64_t)(kSlab64 >> 4);
            const uint64_t tile_units = (uint64_t)(tile_bytes >> 4);
            int s_st = 0; uint32_t s_ph = 0; uint64_t s_vdesc = vk_desc0;
            int o_st = 0; uint64_t o_vdesc = vm_desc0;
            // tcgen05 MMAs issued by one thread execute in order, so S(i+2) overwriting the buffer PV(i) reads P from
            // needs no barrier: the issue order S(i+1), PV(i), S(i+2), PV(i+1), ... is the dependency.
            auto issue_s = [&](int i) {
                const uint32_t b = (uint32_t)i & 1u;
                MOCO_TR(0, i, 4);
                mbar_wait(&kv_full[s_st], s_ph);
                tc_fence_after();
                MOCO_TR(0, i, 5);
                const uint32_t d = tmem_base + kSCol + b * (uint32_t)kDq2BN;
                uint32_t qa = tmem_base + kQCol;
                uint64_t vd = s_vdesc;
                for (int kc = 0; kc < kchunks; ++kc) {
                    umma_ts<1>(d, qa, vd, idesc_s, (uint32_t)(kc != 0));
                    umma_ts<1>(d, qa + 8, vd + 2, idesc_s, 1u);
                    umma_ts<1>(d, qa + 16, vd + 4, idesc_s, 1u);
                    umma_ts<1>(d, qa + 24, vd + 6, idesc_s, 1u);
                    qa += 32;
                    vd += kSlabUnits;
                }
                MOCO_TR(0, i, 6);
                umma_commit<1>(&s_full[b]);
                MOCO_TR(0, i, 7);
                s_vdesc += tile_units;
                if (++s_st == NS) { s_st = 0; s_ph ^= 1u; s_vdesc = vk_desc0; }
            };
            if (ntiles > 0) issue_s(0);
            if (ntiles > 1) issue_s(1);
            for (int i = 0; i < ntiles; ++i) {
                const uint32_t b = (uint32_t)i & 1u;
                MOCO_TR(0, i, 0);
                mbar_wait(&p_full[b], ((uint32_t)i >> 1) & 1u);
                tc_fence_after();
                MOCO_TR(0, i, 1);
#pragma unroll
                for (int kk = 0; kk < kDq2BN / 16; ++kk) {
                    constexpr int kHalfRows = kDq2BN / 2;
                    const uint32_t hh = (uint32_t)((kk * 16) / kHalfRows), off = (uint32_t)(((kk * 16) % kHalfRows) >> 1);
                    umma_ts<1>(tmem_base + kOCol2, tmem_base + kSCol + b * (uint32_t)kDq2BN + hh * (uint32_t)kHalfRows + off,
                               o_vdesc + (uint64_t)(kk * 128), idesc_o, (uint32_t)((i | kk) != 0));
                }
                MOCO_TR(0, i, 2);
                umma_commit<1>(&kv_empty[o_st]);
                MOCO_TR(0, i, 3);
                o_vdesc += tile_units;
                if (++o_st == NS) { o_st = 0; o_vdesc = vm_desc0; }
                if (i + 2 < ntiles) issue_s(i + 2);
            }
            umma_commit<1>(o_full);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------ MMA issuer 1 of 2: S = q . tile^T
            // Two issuing threads (C > 128, BN = 64) because one cannot feed the pipe: a tcgen05.mma costs its issuing
            // thread ~62-72 cycles whatever N <= 128 is (tools/umma_bench.cu), so the 16 N = 64 S-MMAs + 4 PV MMAs
            // of a C = 256 tile take one thread ~1,570 cycles against 1,024 cycles of tensor-pipe work; two threads
            // overlap (umma_bench: 66 -> 53 cycles/MMA aggregate at N = 64, N = 128 reaches 100 %).
            // The order S(i+2)-after-PV(i) that one in-order thread gave for free is now the s_free barrier.
            const uint32_t idesc_s = make_idesc_bf16(128, kDq2BN, 0, 0);            // A: TMEM (q), B: K-major tile
            mbar_wait(q_ready, 0);
            // lean single-thread loop: descriptors built once and advanced by adds, running stage / phase counters
            const uint64_t vk_desc0 = make_sw128_desc(smem_u32(v_s), 0, 1024);
            constexpr uint64_t kSlabUnits = (uint64_t)(kSlab64 >> 4);
            const uint64_t tile_units = (uint64_t)(tile_bytes >> 4);
            int s_st = 0; uint32_t s_ph = 0; uint64_t s_vdesc = vk_desc0;
            for (int i = 0; i < ntiles; ++i) {
                const uint32_t b = (uint32_t)i & 1u;
                MOCO_TR(0, i, 4);
                mbar_wait(&kv_full[s_st], s_ph);
                if (i >= 2) mbar_wait(&s_free[b], (((uint32_t)i >> 1) - 1u) & 1u);   // PV(i-2) has consumed P in buffer b
                tc_fence_after();
                MOCO_TR(0, i, 5);
                const uint32_t d = tmem_base + kSCol + b * (uint32_t)kDq2BN;
                uint32_t qa = tmem_base + kQCol;
                uint64_t vd = s_vdesc;
                for (int kc = 0; kc < kchunks; ++kc) {
                    umma_ts<1>(d, qa, vd, idesc_s, (uint32_t)(kc != 0));
                    umma_ts<1>(d, qa + 8, vd + 2, idesc_s, 1u);
                    umma_ts<1>(d, qa + 16, vd + 4, idesc_s, 1u);
                    umma_ts<1>(d, qa + 24, vd + 6, idesc_s, 1u);
                    qa += 32;
                    vd += kSlabUnits;
                }
                MOCO_TR(0, i, 6);
                umma_commit<1>(&s_full[b]);
                MOCO_TR(0, i, 7);
                s_vdesc += tile_units;
                if (++s_st == NS) { s_st = 0; s_ph ^= 1u; s_vdesc = vk_desc0; }
            }
        }
    } else if (warp == 3 && ISS2) {
        if (lane == 0) {
            // ------------------------------------------------ MMA issuer 2 of 2: O += P . tile
            const uint32_t idesc_o = make_idesc_bf16(128, (uint32_t)a.C, 0, 1);     // A: TMEM (P), B: MN-major tile
            const uint64_t vm_desc0 = make_sw128_desc(smem_u32(v_s), kSlab64, 1024);
            const uint64_t tile_units = (uint64_t)(tile_bytes >> 4);
            int o_st = 0; uint32_t o_ph = 0; uint64_t o_vdesc = vm_desc0;
            for (int i = 0; i < ntiles; ++i) {
                const uint32_t b = (uint32_t)i & 1u;
                MOCO_TR(0, i, 0);
                mbar_wait(&p_full[b], ((uint32_t)i >> 1) & 1u);
                mbar_wait(&kv_full[o_st], o_ph);          // complete long ago (S(i) read the tile); observed for visibility
                tc_fence_after();
                MOCO_TR(0, i, 1);
#pragma unroll
                for (int kk = 0; kk < kDq2BN / 16; ++kk) {
                    // P rows [16kk, 16kk+16) of the tile: column half hh wrote them at the start of ITS S columns
                    constexpr int kHalfRows = kDq2BN / 2;
                    const uint32_t hh = (uint32_t)((kk * 16) / kHalfRows), off = (uint32_t)(((kk * 16) % kHalfRows) >> 1);
                    umma_ts<1>(tmem_base + kOCol2, tmem_base + kSCol + b * (uint32_t)kDq2BN + hh * (uint32_t)kHalfRows + off,
                               o_vdesc + (uint64_t)(kk * 128), idesc_o, (uint32_t)((i | kk) != 0));
                }
                MOCO_TR(0, i, 2);
                umma_commit<1>(&kv_empty[o_st]);
                if (i + 2 < ntiles) umma_commit<1>(&s_free[b]);
                MOCO_TR(0, i, 3);
                o_vdesc += tile_units;
                if (++o_st == NS) { o_st = 0; o_ph ^= 1u; o_vdesc = vm_desc0; }
            }
            umma_commit<1>(o_full);
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------- softmax warps (16) + q staging + O epilogue
        // Two tile groups (grp = tile parity = S/P buffer) x two column halves x four TMEM lane quarters.  The
        // per-tile chain (commit -> mbarrier wake -> tcgen05.ld -> exps -> tcgen05.st -> arrive) is ~1,000+ cycles
        // of latency; with one group it was exposed once per tile and bounded the kernel at C = 128
        // (profiles/README.md); two groups on alternating tiles overlap it.
        const int sw = warp - 4;
        const int quarter = warp & 3;                     // TMEM lanes [32 * quarter, +32)
        const int chalf = (sw >> 2) & 1;
        const int grp = sw >> 3;
        const int row_local = quarter * 32 + lane;
        const int grow = row0 + row_local;
        const float scale2 = a.inv_T * kLog2e;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        if (sw < 4) {
            // q row -> TMEM (A operand layout: lane = row, one 32-bit column = two consecutive bf16 of K)
            const uint4* src = reinterpret_cast<const uint4*>(a.q + (size_t)(grow < a.N ? grow : 0) * a.C);
            {                                             // columns [0, 64): loaded before the set-up barrier
                uint32_t r[32];
#pragma unroll
                for (int v = 0; v < 8; ++v) { r[v * 4 + 0] = qpre[v].x; r[v * 4 + 1] = qpre[v].y; r[v * 4 + 2] = qpre[v].z; r[v * 4 + 3] = qpre[v].w; }
                if (grow >= a.N) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
                tmem_st32(lane_base + kQCol, r);
            }
            if (a.C > 64) {                               // columns [64, 128): idem
                uint32_t r[32];
#pragma unroll
                for (int v = 0; v < 8; ++v) { r[v * 4 + 0] = qpre[8 + v].x; r[v * 4 + 1] = qpre[8 + v].y; r[v * 4 + 2] = qpre[8 + v].z; r[v * 4 + 3] = qpre[8 + v].w; }
                if (grow >= a.N) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
                tmem_st32(lane_base + kQCol + 32u, r);
            }
            for (int c = 128; c < a.C; c += 64) {
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
            if (sw == 0 && lane == 0) MOCO_TR(3, 0, 2);           // q staged into TMEM
        }
        constexpr int kHalf = BN / 2;                     // S columns per thread (32 or 64), in 32-column chunks
        const bool ragged = (a.K % BN) != 0;
        // this thread's columns of S buffer `grp`; its P (bf16 pairs) goes into the FIRST kHalf/2 of those same
        // columns -- chunk h of P lands on columns chunk h/2 of S occupied, which this thread has already read,
        // so no thread ever overwrites S another thread still needs (no barrier between the column halves).
        const uint32_t own = lane_base + kSCol + (uint32_t)(grp * kDq2BN + chalf * kHalf);
        // stabiliser in the log2 domain: the row's lse (two-pass) or the first tile's row maximum (one-pass)
        float lse2 = (!FUSED && grow < a.N) ? a.lse[grow] * kLog2e : 0.f;
        float lsum = 0.f;
        if (FUSED) {
            if (grp == 0) {                               // tile 0 lives in buffer 0
                mbar_wait(&s_full[0], 0);
                tc_fence_after();
                const int valid = (ragged && t0 == a.num_tiles - 1) ? (a.K - (t0 * kDq2BN + chalf * kHalf)) : kHalf;
                float cm = -INFINITY;
#pragma unroll
                for (int h = 0; h < kHalf / 32; ++h) {
                    uint32_t r[32];
                    tmem_ld32(own + (uint32_t)(h * 32), r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (h * 32 + j < valid) cm = fmaxf(cm, __uint_as_float(r[j]));
                }
                exch[chalf * kRowsPerCta + row_local] = cm;
            }
            named_bar_sync(2 + quarter, 128);             // the 4 warps (2 groups x 2 halves) of this lane quarter
            // the first tile always has >= 1 valid column, so at least one of the two maxima is finite
            lse2 = fmaxf(exch[row_local], exch[kRowsPerCta + row_local]) * scale2;
            named_bar_sync(2 + quarter, 128);             // everyone has read: exch may be reused for the sums
        }
        const bool tracer = (quarter == 0 && chalf == 0 && lane == 0);
        for (int i = grp; i < ntiles; i += 2) {
            if (tracer) MOCO_TR(1 + grp, i, 0);
            mbar_wait(&s_full[grp], (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            if (tracer) MOCO_TR(1 + grp, i, 1);
            // queue rows beyond K (last tile only) arrive as zeros: they must not enter the statistics
            const int col0 = (t0 + i) * kDq2BN + chalf * kHalf;
            const int valid = (ragged && t0 + i == a.num_tiles - 1) ? (a.K - col0) : kHalf;
#pragma unroll
            for (int h = 0; h < kHalf / 32; ++h) {
                uint32_t r[32];
                tmem_ld32(own + (uint32_t)(h * 32), r);
                tmem_ld_wait();
                if (tracer && h == 0) MOCO_TR(1 + grp, i, 2);
                // branch-free straight-line block: 32 FFMA, 32 MUFU.EX2, 16 packs (+ 2 sum chains) the scheduler can
                // interleave freely -- a per-element branch or select here costs 2x (profiles/README.md)
                float e[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) e[j] = ex2(fmaf(__uint_as_float(r[j]), scale2, -lse2));
                if (FUSED && valid < kHalf) {             // ragged last tile only: mask (also keeps inf * 0 out of O)
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (h * 32 + j >= valid) e[j] = 0.f;
                }
                uint32_t p[16];
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    if (FUSED) { s0 += e[j]; s1 += e[j + 1]; }
                    __nv_bfloat162 hh = __floats2bfloat162_rn(e[j], e[j + 1]);
                    p[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
                }
                if (FUSED) lsum += s0 + s1;
                if (tracer && h == 0) MOCO_TR(1 + grp, i, 3);
                tmem_st16(own + (uint32_t)(h * 16), p);
            }
            if (tracer) MOCO_TR(1 + grp, i, 4);
            tmem_st_wait();
            if (tracer) MOCO_TR(1 + grp, i, 5);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[grp]);
            if (tracer) MOCO_TR(1 + grp, i, 6);
        }
        if (FUSED) {
            // publish (stabiliser, sum) of this (slice, row): the four partial sums are added in a fixed order
            const int part = grp * 2 + chalf;
            if (part > 0) exch[(part - 1) * kRowsPerCta + row_local] = lsum;
            named_bar_sync(2 + quarter, 128);
            if (part == 0)
                a.part_ms[(size_t)slice * a.n_pad + grow] =
                    make_float2(lse2, ((lsum + exch[row_local]) + exch[kRowsPerCta + row_local]) + exch[2 * kRowsPerCta + row_local]);
        }
        // O epilogue: C/4 columns per warp of a lane quarter when that is a multiple of 32, else C/2 on group 0
        mbar_wait(o_full, 0);
        tc_fence_after();
        if (sw == 0 && lane == 0) MOCO_TR(3, 0, 3);               // last P.V MMA complete
        const bool four = (a.C & 127) == 0;
        if (four || grp == 0) {
            const int ccols = four ? (a.C >> 2) : (a.C >> 1);
            const int cbeg = (four ? (grp * 2 + chalf) : chalf) * ccols;
            // A thread owns one row of O in TMEM; storing it directly makes every warp store hit 32 different
            // 512-byte-strided rows (measured 3.2 us at C = 128).  Each 32 x 32 block goes through a padded
            // (stride 33, conflict-free both ways) buffer in the now idle tile ring, then out as 128-byte rows.
            float* tbuf = reinterpret_cast<float*>(v_s) + sw * (32 * 33);
            float* oblk = a.part_o + ((size_t)slice * a.n_pad + row0 + quarter * 32) * a.C + cbeg + lane;
            for (int c = 0; c < ccols; c += 32) {
                uint32_t r[32];
                tmem_ld32(lane_base + kOCol2 + (uint32_t)(cbeg + c), r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) tbuf[lane * 33 + j] = __uint_as_float(r[j]);
                __syncwarp();
#pragma unroll
                for (int k2 = 0; k2 < 32; ++k2) oblk[(size_t)k2 * a.C + c] = tbuf[k2 * 33 + lane];
                __syncwarp();
            }
        }
    }

    if (warp == 4 && lane == 0) MOCO_TR(3, 0, 4);                 // O written
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) MOCO_TR(3, 0, 5);                       // all roles done
    if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

// lse == nullptr selects the one-pass mode: the kernel also writes ws.part_ms (see the header comment).
// plan_only: launch nothing, just report the slice count / padded rows this shape gets (sharded one-pass finish).
cudaError_t launch_nce_dq2_tc(const __nv_bfloat16* q_bf16, const __nv_bfloat16* queue, int N, int C, int K,
                              float inv_T, const float* lse, int num_sms, int* slices_out,
                              int* n_pad_out, const NceWorkspace& ws, cudaStream_t stream, bool plan_only) {
    const bool fused = (lse == nullptr);
    if (C != 192 && C != 256) return cudaErrorNotSupported;      // C <= 128 runs on nce_head128_sm100.cu
    const int kchunks = C / 64;
    const int mblks = (N + 127) / 128;
    if (mblks > num_sms) return cudaErrorNotSupported;
    const int BN = 64;
    const int num_tiles = (K + BN - 1) / BN;
    const int n_pad = mblks * 128;
    *n_pad_out = n_pad;

    CUtensorMap tm_queue;
    if (!make_tmap(&tm_queue, queue, K, C, BN)) return cudaErrorUnknown;

    const int tile_bytes = kchunks * BN * 128;
    int stages = (kSmemBudget - 2048) / tile_bytes;      // 2 KB: barriers + the one-pass exchange array
    if (stages > 8) stages = 8;
    if (stages < 2) return cudaErrorNotSupported;
    const int smem = stages * tile_bytes + 2048;          // + 1 KB static = 227 KB at C = 256 (7 stages)

    Dq2Args a;
    a.N = N; a.C = C; a.K = K;
    a.mblks = mblks; a.slices = 0; a.n_pad = n_pad; a.num_tiles = num_tiles; a.stages = stages;
    a.inv_T = inv_T;
    a.q = q_bf16;
    a.lse = lse;
    a.part_o = ws.part_o;
    a.part_ms = ws.part_ms;
    auto fill = [](Dq2Args& x, int slices) { x.slices = slices; };
#define MOCO_DQ2_LAUNCH(BN_, IDX)                                                                                  \
    do {                                                                                                           \
        constexpr bool kIss2 = (BN_ == 64);                                                                        \
        if (fused)                                                                                                 \
            return plan_and_launch(nce_dq2_kernel<BN_, true, kIss2>, kernel_cache(2 + IDX), kDq2Threads, smem, 1,  \
                                   mblks, mblks, num_tiles, n_pad, slices_out, stream, tm_queue, tm_queue, a,      \
                                   fill, true, plan_only);                                                         \
        return plan_and_launch(nce_dq2_kernel<BN_, false, kIss2>, kernel_cache(IDX), kDq2Threads, smem, 1, mblks,  \
                               mblks, num_tiles, n_pad, slices_out, stream, tm_queue, tm_queue, a, fill, true,     \
                               plan_only);                                                                         \
    } while (0)
    MOCO_DQ2_LAUNCH(64, 1);
#undef MOCO_DQ2_LAUNCH
}

#ifdef MOCO_TRACE
extern "C" int moco_debug_dq2_trace(long long* host_buf) {
    return (int)cudaMemcpyFromSymbol(host_buf, g_dq2_trace, sizeof(g_dq2_trace));
}
#endif

}  // namespace moco
