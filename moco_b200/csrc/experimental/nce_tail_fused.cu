// EXPERIMENTAL -- not compiled, never run (see README.md in this directory).
//
// One-pass chain tail: combine_kernel + dq_reduce_kernel (nce_support.cu) fused, one block per q row.
//
//   part_ms[s][i] = (m_s, l_s)   per-slice stabiliser / sum (log2 domain), written by nce_dq2_kernel<FUSED>
//   part_o[s][i][:]              per-slice unnormalised O~_s = sum_j 2^(x_ij - m_s) queue_j
//
//   M     = max(x0, max_s m_s)                      x0 = lpos_i * scale2 (positive logit, log2 domain)
//   l     = 2^(x0 - M) + sum_s l_s 2^(m_s - M)      (fixed order)
//   lse2  = M + log2 l ;  prob = 2^(x0 - lse2) ;  loss_i = lse2 ln2 - lpos_i / T
//   dq_i  = (1 / (T N)) ( sum_s 2^(m_s - lse2) O~_s[i] + (prob - 1) k_i )      (fixed order: groups, then lanes)
//   last block: loss_prob = {mean loss_i, mean prob_i}  (finish_mean: deterministic, self re-arming counter)
//
// Wiring (round 2): declare launch_tail_fused in common.cuh; in capi.cu's one-pass branch replace
// launch_combine + launch_dq_reduce by it; _lib.py: one-pass launches 4 -> 3; tests: every "onepass"/"auto"
// parity case covers it; measure with tools/gpu_lab.py op_c2 (fwd_dq_us).  Must keep: pdl_launch_dependents /
// pdl_wait at the top, no float atomics, rows >= N untouched.
#include "../common.cuh"
#include "../sm100_ptx.cuh"

namespace moco {

// finish_mean, warp_sum, warp_max, load_as_float are file-local helpers of nce_support.cu today: move them to a
// shared header (or move this kernel into nce_support.cu) when wiring.
__device__ void finish_mean(unsigned int* counter, int N, const float* loss_rows, const float* prob_rows, float* loss_prob);
__device__ float warp_sum(float v);
__device__ float warp_max(float v);
__device__ float load_as_float(const void* p, int dtype, size_t idx);

__global__ void __launch_bounds__(256)
tail_fused_kernel(int N, int C, int slices, int n_pad, float inv_T, const float* __restrict__ lpos,
                  const float2* __restrict__ part_ms, const float* __restrict__ part_o, const void* __restrict__ k,
                  int k_dtype, float* __restrict__ lse, float* __restrict__ loss_rows, float* __restrict__ prob_rows,
                  float* __restrict__ loss_prob, float* __restrict__ dq, unsigned int* __restrict__ counters) {
    __shared__ float4 s_part[256];
    __shared__ float s_stat[2];                 // lse2, prob of this row
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x;                   // grid = N blocks
    const float scale2 = inv_T * kLog2e;

    // ---- statistics of row i (warp 0; slices <= 160)
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const float x0 = lpos[i] * scale2;
        float m = x0;
        for (int s = lane; s < slices; s += 32) m = fmaxf(m, part_ms[(size_t)s * n_pad + i].x);
        m = warp_max(m);
        float l = 0.f;
        for (int s = lane; s < slices; s += 32) {
            const float2 ms = part_ms[(size_t)s * n_pad + i];
            l += ms.y * ex2(ms.x - m);
        }
        l = warp_sum(l);                        // butterfly: identical in every lane, order fixed by the shuffle tree
        l += ex2(x0 - m);
        const float lse2 = m + log2f(l);
        const float prob = exp2f(x0 - lse2);
        if (lane == 0) {
            const float lse_nat = lse2 * kLn2;
            lse[i] = lse_nat;
            loss_rows[i] = lse_nat - lpos[i] * inv_T;
            prob_rows[i] = prob;
            s_stat[0] = lse2;
            s_stat[1] = prob;
        }
    }
    __syncthreads();
    const float lse2 = s_stat[0], pm1 = s_stat[1] - 1.f;

    // ---- dq row i: float4 lanes x slice groups, groups then added in index order (as dq_reduce_kernel)
    const int lanes = C >> 2;                   // C % 4 == 0, C <= 1024
    const int groups = 256 / lanes;
    const int lane4 = threadIdx.x % lanes, grp = threadIdx.x / lanes;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (grp < groups) {
        for (int s = grp; s < slices; s += groups) {
            const float w = ex2(part_ms[(size_t)s * n_pad + i].x - lse2);
            const float4 v = __ldcs(reinterpret_cast<const float4*>(part_o + ((size_t)s * n_pad + i) * C) + lane4);
            acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
            acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
        }
        s_part[grp * lanes + lane4] = acc;
    }
    __syncthreads();
    if (grp == 0) {
        for (int g = 1; g < groups; ++g) {
            const float4 v = s_part[g * lanes + lane4];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        const size_t base = (size_t)i * C + lane4 * 4;
        const float gscale = inv_T / (float)N;
        float4 o;
        o.x = gscale * (acc.x + pm1 * load_as_float(k, k_dtype, base + 0));
        o.y = gscale * (acc.y + pm1 * load_as_float(k, k_dtype, base + 1));
        o.z = gscale * (acc.z + pm1 * load_as_float(k, k_dtype, base + 2));
        o.w = gscale * (acc.w + pm1 * load_as_float(k, k_dtype, base + 3));
        *reinterpret_cast<float4*>(dq + base) = o;
    }
    finish_mean(counters + 0, N, loss_rows, prob_rows, loss_prob);   // contains the __syncthreads it needs
}

}  // namespace moco
