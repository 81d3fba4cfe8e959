// InfoNCE head kernel for feat_dim <= 128 (MoCo's 128): ONE sweep over the queue on tcgen05 produces the softmax
// statistics AND the unnormalised gradient partials of a 128-row block of queries against a slice of the queue.
//
//   q[128, C]   read straight from the caller's tensor (fp32 or bf16, optionally L2-normalised here -- the
//               reference's Normalize layer, moco/models/resnet.py:24-33 -- then rounded to bf16) by all 16 softmax
//               warps with coalesced row loads and written into shared memory in the UMMA K-major / 128B-swizzle
//               layout: no separate cast kernel, no global round trip, nothing for this kernel to wait on but q.
//   S[128, 128] = q . tile^T          tcgen05.mma, A and B from shared memory, fp32 accumulators in TMEM,
//                                     THREE S/P buffers (TMEM: O [0,128) | S0 S1 S2 at 128 + 128 b)
//   P           = 2^(S log2e/T - m)   softmax warps: tcgen05.ld -> ex2 -> bf16 pairs -> tcgen05.st over the start of
//                                     each thread's OWN S columns (no cross-thread hazard)
//   O[128, C]  += P . tile            tcgen05.mma with P as the TMEM-resident A operand and the SAME smem tile as
//                                     MN-major B; O stays in TMEM for the whole slice
//
// Why three S buffers: with two, the per-buffer dependency chain S(i) -> softmax(i) -> P.V(i) -> S(i+2) is exactly
// as long as two tiles of tensor work, so every latency in it (commit -> mbarrier wake-up, tcgen05.ld/st, the
// issuing thread's ~65 cycles per MMA) was exposed: 1,750 cycles per 128-row tile against 1,024 of tensor / MUFU
// work (profiles/r2a_trace_c3.txt).  q used to live in TMEM (64 columns); staging it in shared memory instead makes
// room for the third buffer, so S(i+2) is already waiting when the softmax group finishes tile i.
//
// Two MMA-issuing threads: a tcgen05.mma costs its issuing thread ~65-80 cycles, so the 8 S + 8 P.V MMAs of a tile
// keep ONE thread busy ~1,500-1,850 cycles per tile (profiles/r2b_trace_c4.txt) -- more than the 1,024 cycles of
// tensor work.  Warp 1 issues S, warp 3 issues P.V; S(i+3) overwriting the buffer P.V(i) reads P from is ordered by
// the second arrival on kv_full[stage of tile i+3], which P.V(i)'s tcgen05.commit provides (see the barrier set-up).
// (That ~65-80 cycles was itself an artefact: the issuing threads are now chosen with elect.sync -- with a threadIdx
// predicate ptxas wraps every tcgen05.mma in an elect/branch loop -- and issue is paced by the pipe.)
//
// Stabiliser (FUSED = true, no lse yet): the CONSTANT m = log2e / T -- the largest logit a unit-norm query can have
// against a unit-norm queue row -- so neither a pass over the first S tile (it cost every CTA ~700 cycles of its
// critical path) nor a row norm (40 dependent shuffles per warp in front of the first MMA) is needed.  P~ = 2^(x - m),
// l = sum P~, O~ = sum P~ queue_j; the tail kernel (nce_tail.cu) merges the per-slice (m, l) pairs and rescales each
// slice's O~ by 2^(m - lse).  All terms of a row share the exponent offset, so precision is that of the exponent-free
// bf16 / fp32 formats; what can go wrong is only the exponent RANGE: rows far from unit norm (the reference's
// U(-s, s) initial queue rows, norm <= sqrt(3), are fine: P~ <= 2^16).  The tail kernel detects both directions
// (a slice sum > 2^100, or a merged sum < 2^-80) and recomputes such rows exactly on CUDA cores, so the result is
// never silently wrong and equals the reference's for ANY q.
// FUSED = false: P is normalised with the given lse (two-pass mode, after the statistics kernel).
//
// Replaces torch.mm + cat + div + CrossEntropyLoss + softmax and autograd's backward GEMM with its queue clone
// (moco/NCE/Contrast.py:23-27, NCECriterion.py:11-13, train.py:264,273).
#include <cuda.h>

#include "../../include/moco_b200.h"
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tc_common.cuh"

namespace moco {

#ifdef MOCO_TRACE
__device__ long long g_h128_trace[4][64][8];
__device__ unsigned long long g_h128_cta[160][4];      // per CTA: globaltimer at entry / exit, %smid, tiles
__device__ unsigned long long g_moco_evt[64][4];       // per launch: sweep first entry / last exit
__device__ unsigned int g_moco_launch = 0, g_moco_exits = 0;
__device__ __forceinline__ unsigned long long h128_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define MOCO_TR(role, tile, slot) do { if (blockIdx.x == 0 && (tile) < 64) g_h128_trace[role][tile][slot] = clock64(); } while (0)
#else
#define MOCO_TR(role, tile, slot) do { } while (0)
#endif

constexpr int kH1Threads = 640;           // warp0 TMA, warp1 S-MMA, warp2 TMEM alloc, warp3 PV-MMA, warps 4-19 softmax
constexpr int kH1BN = 128;                // queue rows per tile
constexpr int kH1Bufs = 3;                // S/P buffers in TMEM
constexpr uint32_t kH1OCol = 0, kH1SCol = 128;
constexpr int kH1Slab = kH1BN * 128;      // one [128 rows x 64 bf16] swizzled slab

struct Head128Args {
    int N, C, K;
    int mblks, slices, n_pad, num_tiles, stages;
    float inv_T;
    const void* q;            // [N, C] fp32 or bf16 (qk_dtype)
    int q_dtype;
    int normalize;            // 1: L2-normalise each q row before the bf16 rounding
    const float* lse;         // [N] natural log (two-pass mode)
    float* part_o;            // [slices, n_pad, C]
    float2* part_ms;          // [slices, n_pad] (stabiliser, sum) in the log2 domain (one-sweep mode)
    unsigned int* counters;   // workspace counters the tail kernel's last-block logic uses: zeroed here
    unsigned long long* cta_times;   // [grid][2] %globaltimer at entry / exit (profiling hook moco_prof_sweep_window)
};

template <bool FUSED>
__global__ void __launch_bounds__(kH1Threads, 1)
nce_head128_kernel(const __grid_constant__ CUtensorMap tm_queue, const __grid_constant__ CUtensorMap tm_unused,
                   const Head128Args a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
    if (threadIdx.x == 0) MOCO_TR(3, 0, 0);
#ifdef MOCO_TRACE
    if (threadIdx.x == 0 && blockIdx.x < 160) {
        unsigned int smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_h128_cta[blockIdx.x][0] = h128_gtime(); g_h128_cta[blockIdx.x][2] = smid;
    }
    if (threadIdx.x == 0) atomicMin(&g_moco_evt[g_moco_launch & 63u][0], h128_gtime());
#endif
    const int kchunks = a.C >> 6;
    const int NS = a.stages;
    const int tile_bytes = kchunks * kH1Slab;
    uint8_t* q_s = smem;                                   // kchunks slabs of [128 rows x 128 B]
    uint8_t* v_s = q_s + tile_bytes;                       // NS queue tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + (size_t)NS * tile_bytes);
    uint64_t* kv_full = bars;
    uint64_t* kv_empty = bars + NS;
    uint64_t* s_full = bars + 2 * NS;                      // [3]
    uint64_t* p_full = bars + 2 * NS + 3;                  // [3]
    uint64_t* o_full = bars + 2 * NS + 6;
    uint64_t* q_ready = bars + 2 * NS + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 8);
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
    if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_queue);
    if (warp == 1 && lane == 0) {
        // kv_full[s] completes on TWO arrivals: the TMA fill of the tile (expect_tx) AND the tcgen05.commit of the P.V
        // MMA that last read the S/P buffer the tile's S will overwrite -- one wait per tile for the S-issuer instead of
        // two (an mbarrier wait costs its thread ~100-200 cycles even when the phase has long completed)
        for (int s = 0; s < NS; ++s) { mbar_init(&kv_full[s], 2); mbar_init(&kv_empty[s], 1); }
        for (int b = 0; b < kH1Bufs; ++b) { mbar_init(&s_full[b], 1); mbar_init(&p_full[b], 8); }
        mbar_init(o_full, 1);
        mbar_init(q_ready, 16);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc<1>(tmem_slot, 512);
        tmem_relinquish<1>();
    }
    pdl_wait();                                            // predecessor complete: q / lse / the queue are final
    if (blockIdx.x == 0 && threadIdx.x < 4 && a.counters != nullptr) a.counters[threadIdx.x] = 0u;
    // q rows of this block: 16 warps x 8 rows, one coalesced row load per (warp, row), all 8 in flight before the
    // set-up barrier (their latency is the kernel's critical path at small K)
    if (threadIdx.x == 0) MOCO_TR(3, 0, 1);
    // Each softmax warp owns 8 consecutive q rows = 8 * C * esz contiguous bytes, fetched as 16-byte-per-lane loads
    // (512 B per instruction: 2, 4 or 8 instructions).  RAW bits only: nothing may consume a loaded value before the
    // barrier below, or the load latency lands in front of it.
    uint4 qraw[8];
    const int esz = (a.q_dtype == MOCO_F32) ? 4 : 2;
    const int row_bytes = a.C * esz;                       // 128, 256 or 512
    const int rb_shift = 31 - __clz(row_bytes);            // a power of two: shifts, not divisions, in front of the loads
    const int n_ld = row_bytes >> 6;                       // 8 rows * row_bytes / 512
    // All CTAs of an m-block read the same 128 q rows at the same moment: each starts at a different 8-row block
    // (rotation by the slice index) so that the requests spread over the L2 slices instead of queueing on a few lines.
    const int qblk = (warp - 4 + slice) & 15;              // the 8-row block this softmax warp stages
    if (warp >= 4) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            qraw[t] = make_uint4(0u, 0u, 0u, 0u);
            if (t < n_ld) {
                const int o = t * 512 + lane * 16;         // byte offset inside the warp's 8-row block
                const int grow = row0 + qblk * 8 + (o >> rb_shift);
                const uint8_t* src = static_cast<const uint8_t*>(a.q) + ((size_t)(grow < a.N ? grow : 0) << rb_shift) + (o & (row_bytes - 1));
                qraw[t] = __ldg(reinterpret_cast<const uint4*>(src));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) MOCO_TR(3, 0, 2);

    if (warp == 0) {
        if (elect_one()) {
            // ------------------------------------------------ TMA producer (queue tiles)
            int st = 0;
            uint32_t ph = 0;
            for (int i = 0; i < ntiles; ++i, st = (st + 1 == NS) ? 0 : st + 1, ph ^= (st == 0) ? 1u : 0u) {
                mbar_wait(&kv_empty[st], ph ^ 1u);
                if (i < kH1Bufs) mbar_arrive(&kv_full[st]);             // no earlier P.V to wait for
                mbar_arrive_expect_tx(&kv_full[st], (uint32_t)tile_bytes);
                for (int kc = 0; kc < kchunks; ++kc)
                    tma_load_2d(&tm_queue, &kv_full[st], v_s + (size_t)st * tile_bytes + kc * kH1Slab, kc * 64,
                                (t0 + i) * kH1BN);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ------------------------------------------------ MMA issuer 1 of 2: S = q . tile^T (both operands in smem)
            const uint32_t idesc_s = make_idesc_bf16(128, kH1BN, 0, 0);
            mbar_wait(q_ready, 0);
            tc_fence_after();
            MOCO_TR(3, 1, 0);
            const uint64_t q_desc0 = make_sw128_desc(smem_u32(q_s), 0, 1024);
            const uint64_t vk_desc0 = make_sw128_desc(smem_u32(v_s), 0, 1024);           // tile as K-major B
            constexpr uint64_t kSlabUnits = (uint64_t)(kH1Slab >> 4);
            const uint64_t tile_units = (uint64_t)(tile_bytes >> 4);
            int s_st = 0; uint32_t s_ph = 0; uint64_t s_vdesc = vk_desc0; uint32_t s_b = 0;
            for (int i = 0; i < ntiles; ++i) {
                MOCO_TR(0, i, 4);
                mbar_wait(&kv_full[s_st], s_ph);          // tile landed AND P.V(i-3) has consumed P in buffer s_b
                tc_fence_after();
                MOCO_TR(0, i, 5);
                const uint32_t d = tmem_base + kH1SCol + s_b * (uint32_t)kH1BN;
                uint64_t qd = q_desc0, vd = s_vdesc;
                for (int kc = 0; kc < kchunks; ++kc) {
                    umma_ss<1>(d, qd, vd, idesc_s, (uint32_t)(kc != 0));
                    umma_ss<1>(d, qd + 2, vd + 2, idesc_s, 1u);
                    umma_ss<1>(d, qd + 4, vd + 4, idesc_s, 1u);
                    umma_ss<1>(d, qd + 6, vd + 6, idesc_s, 1u);
                    qd += kSlabUnits;
                    vd += kSlabUnits;
                }
                MOCO_TR(0, i, 6);
                umma_commit<1>(&s_full[s_b]);
                MOCO_TR(0, i, 7);
                s_vdesc += tile_units;
                if (++s_st == NS) { s_st = 0; s_ph ^= 1u; s_vdesc = vk_desc0; }
                if (++s_b == kH1Bufs) s_b = 0;
            }
        }
    } else if (warp == 3) {
        if (elect_one()) {
            // ------------------------------------------------ MMA issuer 2 of 2: O += P . tile (P in TMEM, tile MN-major)
            const uint32_t idesc_o = make_idesc_bf16(128, (uint32_t)a.C, 0, 1);
            const uint64_t vm_desc0 = make_sw128_desc(smem_u32(v_s), kH1Slab, 1024);
            const uint64_t tile_units = (uint64_t)(tile_bytes >> 4);
            int o_st = 0; uint64_t o_vdesc = vm_desc0; uint32_t o_b = 0, o_ph = 0;
            for (int i = 0; i < ntiles; ++i) {
                MOCO_TR(0, i, 0);
                // p_full alone orders this thread after the tile's TMA fill: softmax(i) arrived here after it saw
                // s_full, which S(i)'s commit raised after the S-issuer had observed kv_full
                mbar_wait(&p_full[o_b], o_ph);
                tc_fence_after();
                MOCO_TR(0, i, 1);
#pragma unroll
                for (int kk = 0; kk < kH1BN / 16; ++kk) {
                    // P rows [16kk, 16kk+16) of the tile: column half hh wrote them at the start of ITS S columns
                    const uint32_t hh = (uint32_t)(kk >> 2), off = (uint32_t)((kk & 3) * 8);
                    umma_ts<1>(tmem_base + kH1OCol, tmem_base + kH1SCol + o_b * (uint32_t)kH1BN + hh * 64u + off,
                               o_vdesc + (uint64_t)(kk * 128), idesc_o, (uint32_t)((i | kk) != 0));
                }
                MOCO_TR(0, i, 2);
                umma_commit<1>(&kv_empty[o_st]);
                if (i + kH1Bufs < ntiles) {               // second arrival on the barrier S(i+3) waits on (its tile's stage)
                    int st3 = o_st + kH1Bufs;
                    if (st3 >= NS) st3 -= NS;
                    umma_commit<1>(&kv_full[st3]);
                }
                MOCO_TR(0, i, 3);
                o_vdesc += tile_units;
                if (++o_st == NS) { o_st = 0; o_vdesc = vm_desc0; }
                if (++o_b == kH1Bufs) { o_b = 0; o_ph ^= 1u; }
            }
            umma_commit<1>(o_full);
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------- softmax warps (16): q staging, softmax, O epilogue
        const int sw = warp - 4;
        const int quarter = warp & 3;                     // TMEM lanes [32 * quarter, +32)
        const int chalf = (sw >> 2) & 1;
        const int grp = sw >> 3;                          // tile parity this warp serves
        const int row_local = quarter * 32 + lane;
        const int grow = row0 + row_local;
        const float scale2 = a.inv_T * kLog2e;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        {
            // -> shared memory, K-major 128B-swizzle (row r at r*128 B inside a 64-column slab, 16-byte chunk c at
            // position c ^ (r & 7)): exactly what a TMA load of the bf16 tensor would have written
            const int lanes_per_row = row_bytes >> 4;      // 8, 16 or 32 lanes hold one row
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                if (t < n_ld) {
                    const int o = t * 512 + lane * 16;
                    const int r = qblk * 8 + (o >> rb_shift);                  // row inside the CTA's 128-row block
                    const int col0 = (o & (row_bytes - 1)) >> (esz == 4 ? 2 : 1);
                    const bool pad = row0 + r >= a.N;
                    if (t == 0 && sw == 0 && lane == 0) MOCO_TR(3, 0, 3);     // first q data in registers
                    if (a.q_dtype == MOCO_F32) {           // 4 fp32 -> 4 bf16 (8 bytes)
                        float v0 = __uint_as_float(qraw[t].x), v1 = __uint_as_float(qraw[t].y);
                        float v2 = __uint_as_float(qraw[t].z), v3 = __uint_as_float(qraw[t].w);
                        if (a.normalize) {                 // x / sqrt(sum x^2): resnet.py:31-32
                            float ss = v0 * v0 + v1 * v1 + v2 * v2 + v3 * v3;
                            for (int w = lanes_per_row >> 1; w > 0; w >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, w);
                            const float n = sqrtf(ss);
                            v0 = v0 / n; v1 = v1 / n; v2 = v2 / n; v3 = v3 / n;
                        }
                        if (pad) { v0 = v1 = v2 = v3 = 0.f; }                 // (0/0 of a padding row must not reach the MMA)
                        const __nv_bfloat162 lo = __floats2bfloat162_rn(v0, v1), hi = __floats2bfloat162_rn(v2, v3);
                        uint8_t* dst = q_s + (col0 >> 6) * kH1Slab + r * 128 + ((((col0 & 63) >> 3) ^ (r & 7)) << 4) + (col0 & 7) * 2;
                        *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo),
                                                                    *reinterpret_cast<const uint32_t*>(&hi));
                    } else {                               // 8 bf16 = one 16-byte swizzle chunk
                        uint4 u = qraw[t];
                        if (a.normalize) {
                            float f[8];
                            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
                            float ss = 0.f;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 x = __bfloat1622float2(h[e]);
                                f[2 * e] = x.x; f[2 * e + 1] = x.y;
                                ss = fmaf(x.x, x.x, fmaf(x.y, x.y, ss));
                            }
                            for (int w = lanes_per_row >> 1; w > 0; w >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, w);
                            const float n = sqrtf(ss);
                            __nv_bfloat162 o2[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) o2[e] = __floats2bfloat162_rn(f[2 * e] / n, f[2 * e + 1] / n);
                            u = *reinterpret_cast<const uint4*>(o2);
                        }
                        if (pad) u = make_uint4(0u, 0u, 0u, 0u);
                        uint8_t* dst = q_s + (col0 >> 6) * kH1Slab + r * 128 + ((((col0 & 63) >> 3) ^ (r & 7)) << 4);
                        *reinterpret_cast<uint4*>(dst) = u;
                    }
                }
            }
            if (sw == 0 && lane == 0) MOCO_TR(3, 2, 1);                   // smem stores issued
            fence_proxy_async();                          // generic-proxy smem writes -> visible to tcgen05.mma
            __syncwarp();
            if (lane == 0) mbar_arrive(q_ready);
            if (sw == 0 && lane == 0) MOCO_TR(3, 0, 4);
        }
        constexpr int kHalf = kH1BN / 2;                  // 64 S columns per thread, in two 32-column chunks
        const bool ragged = (a.K % kH1BN) != 0;
        float lse2 = (!FUSED && grow < a.N) ? a.lse[grow] * kLog2e : 0.f;
        float lsum = 0.f;
        if (FUSED) lse2 = scale2;                         // the stabiliser of unit-norm rows (see the header comment)
        const bool tracer = (quarter == 0 && chalf == 0 && lane == 0);
        uint32_t b = (uint32_t)grp;                       // buffer of tile i = i % 3, advanced by 2 per iteration
        uint32_t use = 0;                                 // i / 3
        for (int i = grp; i < ntiles; i += 2) {
            if (tracer) MOCO_TR(1 + grp, i, 0);
            mbar_wait(&s_full[b], use & 1u);
            tc_fence_after();
            if (tracer) MOCO_TR(1 + grp, i, 1);
            const uint32_t own = lane_base + kH1SCol + b * (uint32_t)kH1BN + (uint32_t)(chalf * kHalf);
            const int col0 = (t0 + i) * kH1BN + chalf * kHalf;
            const int valid = (ragged && t0 + i == a.num_tiles - 1) ? (a.K - col0) : kHalf;
#pragma unroll
            for (int h = 0; h < kHalf / 32; ++h) {
                uint32_t r[32];
                tmem_ld32(own + (uint32_t)(h * 32), r);
                tmem_ld_wait();
                if (tracer && h == 0) MOCO_TR(1 + grp, i, 2);
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
            if (lane == 0) mbar_arrive(&p_full[b]);
            if (tracer) MOCO_TR(1 + grp, i, 6);
            b += 2;
            if (b >= (uint32_t)kH1Bufs) { b -= (uint32_t)kH1Bufs; ++use; }
        }
        if (FUSED) {
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
        if (sw == 0 && lane == 0) MOCO_TR(3, 0, 5);
        const bool four = (a.C & 127) == 0;
        if (four || grp == 0) {
            const int ccols = four ? (a.C >> 2) : (a.C >> 1);
            const int cbeg = (four ? (grp * 2 + chalf) : chalf) * ccols;
            // each 32 x 32 block goes through a padded (stride 33) buffer in the now idle tile ring, then out as
            // 128-byte rows (a direct store would hit 32 rows 512 B apart per instruction)
            float* tbuf = reinterpret_cast<float*>(v_s) + sw * (32 * 33);
            float* oblk = a.part_o + ((size_t)slice * a.n_pad + row0 + quarter * 32) * a.C + cbeg + lane;
            for (int c = 0; c < ccols; c += 32) {
                uint32_t r[32];
                tmem_ld32(lane_base + kH1OCol + (uint32_t)(cbeg + c), r);
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

    if (warp == 4 && lane == 0) MOCO_TR(3, 0, 6);
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) MOCO_TR(3, 0, 7);
#ifdef MOCO_TRACE
    if (threadIdx.x == 0 && blockIdx.x < 160) { g_h128_cta[blockIdx.x][1] = h128_gtime(); g_h128_cta[blockIdx.x][3] = (unsigned long long)ntiles; }
    if (threadIdx.x == 0) {
        atomicMax(&g_moco_evt[g_moco_launch & 63u][1], h128_gtime());
        __threadfence();
        if (atomicAdd(&g_moco_exits, 1u) == gridDim.x - 1) { g_moco_exits = 0u; __threadfence(); g_moco_launch = g_moco_launch + 1u; }
    }
#endif
    if (threadIdx.x == 0 && a.cta_times != nullptr) {      // two plain stores per CTA; read by the bench's profiling hook
        unsigned long long t_exit;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_exit));
        a.cta_times[2 * blockIdx.x + 1] = t_exit;
    }
    if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

// lse == nullptr selects the one-sweep mode (the kernel also writes ws.part_ms).
// plan_only: launch nothing, just report the slice count / padded rows this shape gets.
cudaError_t launch_nce_head128(const void* q, int q_dtype, int normalize, const __nv_bfloat16* queue, int N, int C, int K,
                               float inv_T, const float* lse, int num_sms, int* slices_out, int* n_pad_out,
                               const NceWorkspace& ws, cudaStream_t stream, bool plan_only) {
    const bool fused = (lse == nullptr);
    if (C != 64 && C != 128) return cudaErrorNotSupported;
    if ((reinterpret_cast<uintptr_t>(q) & 15) != 0) return cudaErrorNotSupported;
    const int kchunks = C / 64;
    const int mblks = (N + 127) / 128;
    if (mblks > num_sms) return cudaErrorNotSupported;
    const int num_tiles = (K + kH1BN - 1) / kH1BN;
    const int n_pad = mblks * 128;
    *n_pad_out = n_pad;

    CUtensorMap tm_queue;
    if (!make_tmap(&tm_queue, queue, K, C, kH1BN)) return cudaErrorUnknown;

    const int tile_bytes = kchunks * kH1Slab;
    int stages = (kSmemBudget - tile_bytes - 2048) / tile_bytes;     // q tile + 2 KB barriers / exchange array
    if (stages > 8) stages = 8;
    if (stages * tile_bytes < 16 * 32 * 33 * 4) return cudaErrorNotSupported;   // the O epilogue stages through the ring
    const int smem = (stages + 1) * tile_bytes + 2048;

    Head128Args a;
    a.N = N; a.C = C; a.K = K;
    a.mblks = mblks; a.slices = 0; a.n_pad = n_pad; a.num_tiles = num_tiles; a.stages = stages;
    a.inv_T = inv_T;
    a.q = q; a.q_dtype = q_dtype; a.normalize = normalize;
    a.lse = lse;
    a.part_o = ws.part_o;
    a.part_ms = ws.part_ms;
    a.counters = ws.counters;
    a.cta_times = ws.cta_times;
    auto fill = [](Head128Args& x, int slices) { x.slices = slices; };
    if (fused)
        return plan_and_launch(nce_head128_kernel<true>, kernel_cache(0), kH1Threads, smem, 1, mblks, mblks, num_tiles,
                               n_pad, slices_out, stream, tm_queue, tm_queue, a, fill, true, plan_only);
    return plan_and_launch(nce_head128_kernel<false>, kernel_cache(1), kH1Threads, smem, 1, mblks, mblks, num_tiles, n_pad,
                           slices_out, stream, tm_queue, tm_queue, a, fill, true, plan_only);
}

#ifdef MOCO_TRACE
extern "C" int moco_debug_h128_trace(long long* host_buf) {
    return (int)cudaMemcpyFromSymbol(host_buf, g_h128_trace, sizeof(g_h128_trace));
}
extern "C" int moco_debug_evt(unsigned long long* host_buf) {
    return (int)cudaMemcpyFromSymbol(host_buf, g_moco_evt, sizeof(g_moco_evt));
}
extern "C" int moco_debug_evt_reset() {
    static unsigned long long init[64][4];
    for (int i = 0; i < 64; ++i) { init[i][0] = ~0ull; init[i][1] = 0; init[i][2] = ~0ull; init[i][3] = 0; }
    unsigned int z = 0;
    cudaMemcpyToSymbol(g_moco_launch, &z, sizeof(z));
    return (int)cudaMemcpyToSymbol(g_moco_evt, init, sizeof(init));
}
extern "C" int moco_debug_h128_cta(unsigned long long* host_buf) {
    return (int)cudaMemcpyFromSymbol(host_buf, g_h128_cta, sizeof(g_h128_cta));
}
#endif

}  // namespace moco
