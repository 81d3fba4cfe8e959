// InfoNCE head kernel for feat_dim 192 / 256 (BASELINE configs[4]: C = 256, the tensor-bound stress shape): ONE sweep
// over the queue on tcgen05 produces the softmax statistics AND the unnormalised gradient partials of a 128-row block
// of queries against a slice of the queue (two-pass mode: the gradient partials only, normalised with a given lse).
//
//   S[128, 64]  = q . tile^T     tcgen05.mma kind::f16, N = 64.  q is split: its first 128 columns live in TMEM as the
//                                A operand (8 "TS" MMAs per tile), the remaining 64 / 128 columns in shared memory
//                                (4 / 8 "SS" MMAs per tile).
//   P           = 2^(S log2e/T - m) -> bf16 pairs written over the start of each thread's OWN S columns (tcgen05.st)
//   O[128, C]  += P . tile       P as the TMEM-resident A operand, the SAME smem tile as MN-major B (4 MMAs, N = C)
//
// Why this shape.  TMEM has 512 columns and O takes C (256) of them.  Round 1 kept all of q in TMEM (128 columns),
// which left room for only TWO 64-wide S/P buffers; the per-buffer dependency chain S(i) -> softmax(i) -> P.V(i) ->
// S(i+2) (~2,700-4,000 cycles with its mbarrier wake-ups) was longer than the two tiles of tensor work it has to
// cover, and on top of that every tcgen05.mma sat in a compiler-generated elect/branch loop (the issuing thread was
// selected with `lane == 0`; see the note at the issue loops): 1,640 cycles per 64-row tile against 1,024 of tensor
// work, 54.8 % tensor-pipe utilisation (profiles/r1_onepass_c5_ncu_metrics.csv, profiles/r2a_trace_c5.txt).
// Here HALF of q stays in TMEM (64 columns) and the other half goes to shared memory (32 KB), which makes room for
// THREE S/P buffers (64 + 256 + 3 x 64 = 512 columns) next to six 32 KB ring stages (32 + 6 x 32 + 2 = 226 KB): the
// chain of one buffer now has three tiles of tensor work to hide behind.  The smem ports carry 32 KB (q half) +
// 2 x 32 KB (tile, read by both MMAs) + 32 KB (TMA fill) per 1,024 tensor cycles = 125 B/clk.  Two issuing threads
// (S on warp 1, P.V on warp 3); S(i+3) is ordered after P.V(i) by the second arrival on kv_full[stage of tile i+3]
// that P.V(i)'s tcgen05.commit provides.  (A 96-row tile with two buffers was measured first:
// 99.7 us, chain-bound at ~2,100 cycles per tile against 1,536 -- profiles/r2d_trace_c5_bn96.txt.)
//
// Stabiliser (FUSED): the constant m = log2e / T, see nce_head128_sm100.cu; the tail kernel (nce_tail.cu) detects
// rows whose exponent range it cannot hold and recomputes them exactly.
//
// Replaces torch.mm + cat + div + CrossEntropyLoss + softmax and autograd's backward GEMM with its queue clone
// (moco/NCE/Contrast.py:23-27, NCECriterion.py:11-13, train.py:264,273).
#include <cuda.h>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tc_common.cuh"

namespace moco {

#ifdef MOCO_TRACE
__device__ long long g_dq2_trace[4][64][8];
#define MOCO_TR(role, tile, slot) do { if (blockIdx.x == 0 && (tile) < 64) g_dq2_trace[role][tile][slot] = clock64(); } while (0)
#else
#define MOCO_TR(role, tile, slot) do { } while (0)
#endif

constexpr int kH2Threads = 640;           // warp0 TMA, warp1 S-MMA, warp2 TMEM alloc, warp3 PV-MMA, warps 4-19 softmax
constexpr int kH2BN = 64;                 // queue rows per tile
constexpr int kH2Bufs = 3;                // S/P buffers in TMEM
constexpr int kH2Slab = kH2BN * 128;      // one [64 rows x 64 bf16] swizzled slab: 8 KB
constexpr int kH2QSlab = 128 * 128;       // one [128 rows x 64 bf16] slab of q: 16 KB
constexpr uint32_t kH2QCol = 0, kH2OCol = 64, kH2SCol = 320;

struct Head256Args {
    int N, C, K;
    int mblks, slices, n_pad, num_tiles, stages;
    float inv_T;
    const __nv_bfloat16* q;   // [N, C] bf16
    const float* lse;         // [N] natural log (two-pass mode)
    float* part_o;            // [slices, n_pad, C]
    float2* part_ms;          // [slices, n_pad] (stabiliser, sum) in the log2 domain (one-sweep mode)
    unsigned int* counters;   // zeroed here for the tail kernel
    unsigned long long* cta_times;   // [grid][2] %globaltimer at entry / exit (profiling hook moco_prof_sweep_window)
};

template <bool FUSED>
__global__ void __launch_bounds__(kH2Threads, 1)
nce_head256_kernel(const __grid_constant__ CUtensorMap tm_queue, const __grid_constant__ CUtensorMap tm_q,
                   const Head256Args a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
    if (threadIdx.x == 0) MOCO_TR(3, 0, 0);
    const int kchunks = a.C >> 6;                          // 3 or 4
    const int qhi_chunks = kchunks - 2;                    // 64-column chunks of q that live in shared memory
    const int NS = a.stages;
    const int tile_bytes = kchunks * kH2Slab;
    uint8_t* qhi_s = smem;
    uint8_t* v_s = qhi_s + qhi_chunks * kH2QSlab;
    uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + (size_t)NS * tile_bytes);
    uint64_t* kv_full = bars;
    uint64_t* kv_empty = bars + NS;
    uint64_t* s_full = bars + 2 * NS;        // [3]
    uint64_t* p_full = bars + 2 * NS + 3;    // [3]
    uint64_t* o_full = bars + 2 * NS + 6;
    uint64_t* q_ready = bars + 2 * NS + 7;   // q's first 128 columns are in TMEM
    uint64_t* qhi_full = bars + 2 * NS + 8;  // q's remaining columns have landed in smem
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 9);
    float* exch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);     // [3][128] floats
    // profiling hook (moco_prof_sweep_window): this slot is written by this kernel only and read by the host only, so the
    // store may precede griddepcontrol.wait
    if (threadIdx.x == 0 && a.cta_times != nullptr) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.cta_times[2 * blockIdx.x] = t;
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mblk = blockIdx.x % a.mblks;
    const int slice = blockIdx.x / a.mblks;
    const int t0 = (int)(((long long)slice * a.num_tiles) / a.slices);
    const int t1 = (int)(((long long)(slice + 1) * a.num_tiles) / a.slices);
    const int ntiles = t1 - t0;
    const int row0 = mblk * kRowsPerCta;

    pdl_launch_dependents();
    // ---- set-up that touches no global memory (overlaps the predecessor kernel under PDL) ----
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_queue); tma_prefetch_desc(&tm_q); }
    if (warp == 1 && lane == 0) {
        // kv_full[s] completes on TWO arrivals: the TMA fill of the tile (expect_tx) AND the tcgen05.commit of the P.V
        // MMA that last read the S/P buffer the tile's S will overwrite -- one wait per tile for the S-issuer instead of
        // two (an mbarrier wait costs its thread ~100-200 cycles even when the phase has long completed)
        for (int s = 0; s < NS; ++s) { mbar_init(&kv_full[s], 2); mbar_init(&kv_empty[s], 1); }
        for (int b = 0; b < kH2Bufs; ++b) { mbar_init(&s_full[b], 1); mbar_init(&p_full[b], 8); }
        mbar_init(o_full, 1);
        mbar_init(q_ready, 4);
        mbar_init(qhi_full, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc<1>(tmem_slot, 512);
        tmem_relinquish<1>();
    }
    pdl_wait();                                            // predecessor complete: q / lse / the queue are final
    if (blockIdx.x == 0 && threadIdx.x < 4 && a.counters != nullptr) a.counters[threadIdx.x] = 0u;
    // q staging warps (4-7): the first 128 columns of their row, raw bits in flight before the set-up barrier;
    // the 16-byte pieces are fetched in an order rotated by the slice index (the CTAs of an m-block read the same rows)
    uint4 qpre[16];
    if (warp >= 4 && warp < 8) {
        const int grow_q = row0 + (warp & 3) * 32 + lane;
        const uint4* src = reinterpret_cast<const uint4*>(a.q + (size_t)(grow_q < a.N ? grow_q : 0) * a.C);
        // (four statically indexed variants: a run-time register index would put qpre[] in local memory)
#define MOCO_LOADQ(ROT)                                                                   \
        _Pragma("unroll") for (int v = 0; v < 16; ++v) qpre[(v + ROT) & 15] = __ldg(src + ((v + ROT) & 15));
        switch (slice & 3) {
            case 0: MOCO_LOADQ(0) break;
            case 1: MOCO_LOADQ(4) break;
            case 2: MOCO_LOADQ(8) break;
            default: MOCO_LOADQ(12) break;
        }
#undef MOCO_LOADQ
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) MOCO_TR(3, 0, 1);

    if (warp == 0) {
        if (elect_one()) {
            // ------------------------------------------------ TMA producer: q's smem part once, then the queue tiles
            mbar_arrive_expect_tx(qhi_full, (uint32_t)(qhi_chunks * kH2QSlab));
            for (int kc = 0; kc < qhi_chunks; ++kc)
                tma_load_2d(&tm_q, qhi_full, qhi_s + kc * kH2QSlab, (2 + kc) * 64, row0);
            int st = 0;
            uint32_t ph = 0;
            for (int i = 0; i < ntiles; ++i, st = (st + 1 == NS) ? 0 : st + 1, ph ^= (st == 0) ? 1u : 0u) {
                mbar_wait(&kv_empty[st], ph ^ 1u);
                if (i < kH2Bufs) mbar_arrive(&kv_full[st]);             // no earlier P.V to wait for
                mbar_arrive_expect_tx(&kv_full[st], (uint32_t)tile_bytes);
                for (int kc = 0; kc < kchunks; ++kc)
                    tma_load_2d(&tm_queue, &kv_full[st], v_s + (size_t)st * tile_bytes + kc * kH2Slab, kc * 64,
                                (t0 + i) * kH2BN);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ------------------------------------------------ MMA issuer 1 of 2: S = q . tile^T
            const uint32_t idesc_s = make_idesc_bf16(128, kH2BN, 0, 0);
            mbar_wait(q_ready, 0);
            mbar_wait(qhi_full, 0);
            tc_fence_after();
            const uint64_t vk_desc0 = make_sw128_desc(smem_u32(v_s), 0, 1024);
            const uint64_t qh_desc0 = make_sw128_desc(smem_u32(qhi_s), 0, 1024);
            constexpr uint64_t kSlabUnits = (uint64_t)(kH2Slab >> 4), kQSlabUnits = (uint64_t)(kH2QSlab >> 4);
            const uint64_t tile_units = (uint64_t)(tile_bytes >> 4);
            int s_st = 0; uint32_t s_ph = 0; uint64_t s_vdesc = vk_desc0; uint32_t b = 0;
            for (int i = 0; i < ntiles; ++i) {
                MOCO_TR(0, i, 4);
                mbar_wait(&kv_full[s_st], s_ph);          // tile landed AND P.V(i-3) has consumed P in buffer b
                tc_fence_after();
                MOCO_TR(0, i, 5);
                const uint32_t d = tmem_base + kH2SCol + b * (uint32_t)kH2BN;
                uint64_t vd = s_vdesc;
                uint32_t qa = tmem_base + kH2QCol;
#pragma unroll
                for (int kc = 0; kc < 2; ++kc) {                       // q columns [0, 128): A from TMEM
                    umma_ts<1>(d, qa, vd, idesc_s, (uint32_t)(kc != 0));
                    umma_ts<1>(d, qa + 8, vd + 2, idesc_s, 1u);
                    umma_ts<1>(d, qa + 16, vd + 4, idesc_s, 1u);
                    umma_ts<1>(d, qa + 24, vd + 6, idesc_s, 1u);
                    qa += 32;
                    vd += kSlabUnits;
                }
                uint64_t qd = qh_desc0;
                for (int kc = 0; kc < qhi_chunks; ++kc) {              // q columns [128, C): A from shared memory
                    umma_ss<1>(d, qd, vd, idesc_s, 1u);
                    umma_ss<1>(d, qd + 2, vd + 2, idesc_s, 1u);
                    umma_ss<1>(d, qd + 4, vd + 4, idesc_s, 1u);
                    umma_ss<1>(d, qd + 6, vd + 6, idesc_s, 1u);
                    qd += kQSlabUnits;
                    vd += kSlabUnits;
                }
                MOCO_TR(0, i, 6);
                umma_commit<1>(&s_full[b]);
                MOCO_TR(0, i, 7);
                s_vdesc += tile_units;
                if (++s_st == NS) { s_st = 0; s_ph ^= 1u; s_vdesc = vk_desc0; }
                if (++b == (uint32_t)kH2Bufs) b = 0;
            }
        }
    } else if (warp == 3) {
        if (elect_one()) {
            // ------------------------------------------------ MMA issuer 2 of 2: O += P . tile
            const uint32_t idesc_o = make_idesc_bf16(128, (uint32_t)a.C, 0, 1);      // A: TMEM (P), B: MN-major tile
            const uint64_t vm_desc0 = make_sw128_desc(smem_u32(v_s), kH2Slab, 1024);
            const uint64_t tile_units = (uint64_t)(tile_bytes >> 4);
            int o_st = 0; uint64_t o_vdesc = vm_desc0; uint32_t b = 0, o_ph = 0;
            for (int i = 0; i < ntiles; ++i) {
                MOCO_TR(0, i, 0);
                // p_full alone orders this thread after the tile's TMA fill: softmax(i) arrived here after it saw
                // s_full, which S(i)'s commit raised after the S-issuer had observed kv_full
                mbar_wait(&p_full[b], o_ph);
                tc_fence_after();
                MOCO_TR(0, i, 1);
#pragma unroll
                for (int kk = 0; kk < kH2BN / 16; ++kk) {
                    // P rows [16kk, 16kk+16) of the tile: column half hh (32 queue rows) wrote them, as 8 columns per 16
                    // rows, at the start of ITS S columns
                    const uint32_t hh = (uint32_t)(kk >> 1), off = (uint32_t)((kk & 1) * 8);
                    umma_ts<1>(tmem_base + kH2OCol, tmem_base + kH2SCol + b * (uint32_t)kH2BN + hh * 32u + off,
                               o_vdesc + (uint64_t)(kk * 128), idesc_o, (uint32_t)((i | kk) != 0));
                }
                MOCO_TR(0, i, 2);
                umma_commit<1>(&kv_empty[o_st]);
                if (i + kH2Bufs < ntiles) {               // second arrival on the barrier S(i+3) waits on (its tile's stage)
                    int st3 = o_st + kH2Bufs;
                    if (st3 >= NS) st3 -= NS;
                    umma_commit<1>(&kv_full[st3]);
                }
                MOCO_TR(0, i, 3);
                o_vdesc += tile_units;
                if (++o_st == NS) { o_st = 0; o_vdesc = vm_desc0; }
                if (++b == (uint32_t)kH2Bufs) { b = 0; o_ph ^= 1u; }
            }
            umma_commit<1>(o_full);
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------- softmax warps (16) + q staging + O epilogue
        // Two tile groups (grp = tile parity = S/P buffer) x two column halves (48 queue rows each) x four TMEM lane
        // quarters.
        const int sw = warp - 4;
        const int quarter = warp & 3;                     // TMEM lanes [32 * quarter, +32)
        const int chalf = (sw >> 2) & 1;
        const int grp = sw >> 3;
        const int row_local = quarter * 32 + lane;
        const int grow = row0 + row_local;
        const float scale2 = a.inv_T * kLog2e;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        if (sw < 4) {
            // q row, columns [0, 128) -> TMEM (A operand layout: lane = row, one 32-bit column = two consecutive bf16)
#pragma unroll
            for (int hblk = 0; hblk < 2; ++hblk) {
                uint32_t r[32];
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    const uint4 u = qpre[hblk * 8 + v];
                    r[v * 4 + 0] = u.x; r[v * 4 + 1] = u.y; r[v * 4 + 2] = u.z; r[v * 4 + 3] = u.w;
                }
                if (grow >= a.N) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
                tmem_st32(lane_base + kH2QCol + (uint32_t)(hblk * 32), r);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_ready);
            if (sw == 0 && lane == 0) MOCO_TR(3, 0, 2);
        }
        constexpr int kHalf = kH2BN / 2;                  // 32 S columns per thread: one tcgen05.ld
        const bool ragged = (a.K % kH2BN) != 0;
        const float lse2 = FUSED ? scale2 : ((grow < a.N) ? a.lse[grow] * kLog2e : 0.f);
        float lsum = 0.f;
        const bool tracer = (quarter == 0 && chalf == 0 && lane == 0);
        uint32_t b = (uint32_t)grp;                       // buffer of tile i = i % 3, advanced by 2 per iteration
        uint32_t use = 0;                                 // i / 3
        for (int i = grp; i < ntiles; i += 2) {
            if (tracer) MOCO_TR(1 + grp, i, 0);
            mbar_wait(&s_full[b], use & 1u);
            tc_fence_after();
            if (tracer) MOCO_TR(1 + grp, i, 1);
            // this thread's columns of S buffer b; its P (bf16 pairs) goes into the FIRST 16 of those same columns
            const uint32_t own = lane_base + kH2SCol + b * (uint32_t)kH2BN + (uint32_t)(chalf * kHalf);
            // queue rows beyond K (last tile only) arrive as zeros: they must not enter the statistics
            const int col0 = (t0 + i) * kH2BN + chalf * kHalf;
            const int valid = (ragged && t0 + i == a.num_tiles - 1) ? (a.K - col0) : kHalf;
            uint32_t r[32];
            tmem_ld32(own, r);
            tmem_ld_wait();
            if (tracer) MOCO_TR(1 + grp, i, 2);
            float e[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) e[j] = ex2(fmaf(__uint_as_float(r[j]), scale2, -lse2));
            if (FUSED && valid < kHalf) {                 // ragged last tile only: mask (also keeps inf * 0 out of O)
#pragma unroll
                for (int j = 0; j < 32; ++j) if (j >= valid) e[j] = 0.f;
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
            if (tracer) MOCO_TR(1 + grp, i, 3);
            tmem_st16(own, p);
            if (tracer) MOCO_TR(1 + grp, i, 4);
            tmem_st_wait();
            if (tracer) MOCO_TR(1 + grp, i, 5);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[b]);
            if (tracer) MOCO_TR(1 + grp, i, 6);
            b += 2;
            if (b >= (uint32_t)kH2Bufs) { b -= (uint32_t)kH2Bufs; ++use; }
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
        if (sw == 0 && lane == 0) MOCO_TR(3, 0, 3);
        const bool four = (a.C & 127) == 0;
        if (four || grp == 0) {
            const int ccols = four ? (a.C >> 2) : (a.C >> 1);
            const int cbeg = (four ? (grp * 2 + chalf) : chalf) * ccols;
            float* tbuf = reinterpret_cast<float*>(v_s) + sw * (32 * 33);
            float* oblk = a.part_o + ((size_t)slice * a.n_pad + row0 + quarter * 32) * a.C + cbeg + lane;
            for (int c = 0; c < ccols; c += 32) {
                uint32_t r[32];
                tmem_ld32(lane_base + kH2OCol + (uint32_t)(cbeg + c), r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) tbuf[lane * 33 + j] = __uint_as_float(r[j]);
                __syncwarp();
#pragma unroll
                for (int k2 = 0; k2 < 32; ++k2) __stcs(oblk + (size_t)k2 * a.C + c, tbuf[k2 * 33 + lane]);
                __syncwarp();
            }
        }
    }

    if (warp == 4 && lane == 0) MOCO_TR(3, 0, 4);
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) MOCO_TR(3, 0, 5);
    if (threadIdx.x == 0 && a.cta_times != nullptr) {      // two plain stores per CTA; read by the bench's profiling hook
        unsigned long long t_exit;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_exit));
        a.cta_times[2 * blockIdx.x + 1] = t_exit;
    }
    if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

// lse == nullptr selects the one-sweep mode (the kernel also writes ws.part_ms).
// plan_only: launch nothing, just report the slice count / padded rows this shape gets (sharded one-sweep finish).
cudaError_t launch_nce_dq2_tc(const __nv_bfloat16* q_bf16, const __nv_bfloat16* queue, int N, int C, int K,
                              float inv_T, const float* lse, int num_sms, int* slices_out,
                              int* n_pad_out, const NceWorkspace& ws, cudaStream_t stream, bool plan_only) {
    const bool fused = (lse == nullptr);
    if (C != 192 && C != 256) return cudaErrorNotSupported;      // C <= 128 runs on nce_head128_sm100.cu
    if ((reinterpret_cast<uintptr_t>(q_bf16) & 15) != 0) return cudaErrorNotSupported;
    const int kchunks = C / 64;
    const int mblks = (N + 127) / 128;
    if (mblks > num_sms) return cudaErrorNotSupported;
    const int num_tiles = (K + kH2BN - 1) / kH2BN;
    const int n_pad = mblks * 128;
    *n_pad_out = n_pad;

    CUtensorMap tm_queue, tm_q;
    if (!make_tmap(&tm_queue, queue, K, C, kH2BN)) return cudaErrorUnknown;
    if (!make_tmap(&tm_q, q_bf16, N, C, 128)) return cudaErrorUnknown;

    const int tile_bytes = kchunks * kH2Slab;
    const int q_bytes = (kchunks - 2) * kH2QSlab;
    int stages = (kSmemBudget - q_bytes - 2048) / tile_bytes;    // 2 KB: barriers + the exchange array
    if (stages > 8) stages = 8;
    if (stages < 2 || stages * tile_bytes < 16 * 32 * 33 * 4) return cudaErrorNotSupported;
    const int smem = q_bytes + stages * tile_bytes + 2048;       // C = 256: 32 + 6 x 32 + 2 KB (+ 1 KB static) = 227 KB

    Head256Args a;
    a.N = N; a.C = C; a.K = K;
    a.mblks = mblks; a.slices = 0; a.n_pad = n_pad; a.num_tiles = num_tiles; a.stages = stages;
    a.inv_T = inv_T;
    a.q = q_bf16;
    a.lse = lse;
    a.part_o = ws.part_o;
    a.part_ms = ws.part_ms;
    a.counters = ws.counters;
    a.cta_times = ws.cta_times;
    auto fill = [](Head256Args& x, int slices) { x.slices = slices; };
    if (fused)
        return plan_and_launch(nce_head256_kernel<true>, kernel_cache(0), kH2Threads, smem, 1, mblks, mblks, num_tiles,
                               n_pad, slices_out, stream, tm_queue, tm_q, a, fill, true, plan_only);
    return plan_and_launch(nce_head256_kernel<false>, kernel_cache(1), kH2Threads, smem, 1, mblks, mblks, num_tiles, n_pad,
                           slices_out, stream, tm_queue, tm_q, a, fill, true, plan_only);
}

#ifdef MOCO_TRACE
extern "C" int moco_debug_dq2_trace(long long* host_buf) {
    return (int)cudaMemcpyFromSymbol(host_buf, g_dq2_trace, sizeof(g_dq2_trace));
}
#endif

}  // namespace moco
