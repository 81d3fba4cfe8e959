// Training-mode batch normalisation of channels_last (NHWC) bf16 activations with the ReLU and the residual add of the
// ResNet blocks folded in: the consumer of ShuffleBN's output (ShuffleBN exists so that these batch statistics cannot
// leak the positive pair) and, measured, 64 % of the GPU time of a MoCo step when left to ATen's kernels
// (profiles/r2_bench_launches_by_kernel.csv: batch_norm_collect_statistics_channels_last 30 %, transform_input 18 %,
// backward_reduce 11 %, backward_elemt 5 %, + the separate ReLU and add passes).
//
// Reference call sites: moco/models/resnet.py:42-63 (BasicBlock), :74-102 (Bottleneck: bn -> relu, bn -> relu,
// bn -> += residual -> relu), :114,156-157 (stem), :139-143 (downsample conv -> bn).  Semantics = torch.nn.BatchNorm2d
// in training mode: per-channel mean and BIASED variance over the N*H*W rows, y = (x - mean) * rsqrt(var + eps) * gamma
// + beta, running_mean / running_var updated with `momentum` (running_var from the UNBIASED variance),
// num_batches_tracked += 1; backward = the standard three-term formula.  Everything is computed in fp32 from the bf16
// activations; outputs are rounded to bf16 once.
//
// The activation tensor is a row-major [M, C] matrix (M = N*H*W, C = channels, C * 2 bytes per row).  All four kernels
// are HBM-bound streaming passes:
//   bn_stats_kernel      reads x once        -> mean, invstd (+ running stats)          2 B / element
//   bn_apply_kernel      reads x (+ residual), writes y = relu(x^ * gamma + beta (+ r)) 4 (6) B / element
//   bn_bwd_reduce_kernel reads dy, x (+ y)   -> sum(g), sum(g * x^) = dbeta, dgamma     4 (6) B / element
//   bn_bwd_apply_kernel  reads dy, x (+ y), writes dx (+ d residual)                    6 (10) B / element
// where g = dy masked by the ReLU (the mask is recomputed from x when there is no residual, read from y otherwise).
//
// Reductions: a CTA owns a 64-channel slab (128 contiguous bytes of every row = one 16-byte vector per lane of an
// 8-lane group) and a contiguous range of rows, 32 rows per pass, 4 or 8 passes in flight (see the table of measured
// settings below); per-CTA partials go to a small workspace and the LAST CTA of a slab to finish (ticket counter)
// adds them in a fixed order in fp64 -- deterministic, no atomics on the data.  Sums are taken of (x - x[0, c]) so
// that the variance does not cancel.
// Element-wise passes: thread t keeps the coefficients of its 8 channels in registers (its channel group never
// changes because the grid stride is a multiple of the row length in vectors) and streams 16-byte vectors linearly.
#include "common.cuh"

#include <cuda_bf16.h>

namespace moco {

constexpr int kBnThreads = 256;
constexpr int kBnSlab = 64;                          // channels per reduction CTA
constexpr int kBnLanes = kBnSlab / 8;                // 16-byte vectors per slab row
constexpr int kBnRows = kBnThreads / kBnLanes;       // rows per pass
constexpr int kBnMaxSlabs = 32;                      // C <= 2048
constexpr int kBnPartial = 2 * kBnSlab;              // floats per CTA partial
// (16-byte loads in flight per thread and operand, resident CTAs per SM) of each kernel; every grid is one resident
// wave.  Measured over the 53 layers of ResNet-50 at batch 256 (tools/bn_kernel_times.py, us per encoder pass):
//   statistics  (4,4) 1652  (8,3) 1748  (8,2) 1610  (2,4) 1773      apply      (4,3) 2527  (4,4) 2372  (8,2) 2145  (2,4) 2531
//   bwd reduce  (4,2) 2809  (4,3) 4586  (2,4) 3689  (2,3) 3294      bwd apply  (4,2) 4818  (4,3) 6816  (2,4) 5342  (2,3) 4601
constexpr int kBnStatsUnroll = 8, kBnStatsCtas = 2;
constexpr int kBnApplyUnroll = 8, kBnApplyCtas = 2;
constexpr int kBnBwdReduceUnroll = 4, kBnBwdReduceCtas = 2;
constexpr int kBnBwdApplyUnroll = 2, kBnBwdApplyCtas = 3;
constexpr int kBnSms = 148;
constexpr int kBnMaxCtas = kBnSms * 4;               // workspace sizing: no reduction grid is larger

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 t = __bfloat1622float2(h[k]);
        f[2 * k] = t.x;
        f[2 * k + 1] = t.y;
    }
}

__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
    return u;
}

// Sum the 16 per-thread accumulators over the CTA's 32 row groups.  Returns, in threads j < 128, element j of the CTA
// partial: j = v * 16 + k with v = 16-byte lane (8 channels), k < 8 the first sum, k >= 8 the second.
__device__ __forceinline__ float slab_reduce(float (&acc)[16], float* red /*[8 * 128]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 8);
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 16);
    }
    if (lane < 8) {
#pragma unroll
        for (int k = 0; k < 16; ++k) red[warp * 128 + lane * 16 + k] = acc[k];
    }
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 128) {
#pragma unroll
        for (int w = 0; w < kBnThreads / 32; ++w) t += red[w * 128 + threadIdx.x];
    }
    return t;
}

// Publishes this CTA's partial, and in the last CTA of the slab to arrive returns true with the slab totals in
// tot[128] (same element order as slab_reduce).  The R partials are added in a fixed order: warp w takes partials
// w, w + 8, ... (lane l owns floats 4l .. 4l+3 of each, one coalesced 512-byte read per partial, 16 reads in flight),
// then the 8 warps' sums are added in warp order.
__device__ __forceinline__ bool slab_finish(float part, float* partial, unsigned int* counter, int slab, int r, int R,
                                            double* tot /*[8 * 128] shared*/, int* flag /*shared*/) {
    float* mine = partial + ((size_t)slab * R + r) * kBnPartial;
    if (threadIdx.x < kBnPartial) mine[threadIdx.x] = part;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(counter, 1u);
        *flag = (ticket == (unsigned int)(R - 1));
    }
    __syncthreads();
    if (!*flag) return false;
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4* base = reinterpret_cast<const float4*>(partial + (size_t)slab * R * kBnPartial) + lane;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    constexpr int kBatch = 16;
    for (int q0 = warp; q0 < R; q0 += 8 * kBatch) {
        float4 v[kBatch];
#pragma unroll
        for (int t = 0; t < kBatch; ++t) {
            const int q = q0 + 8 * t;
            v[t] = (q < R) ? __ldcg(base + (size_t)q * (kBnPartial / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int t = 0; t < kBatch; ++t) { s0 += v[t].x; s1 += v[t].y; s2 += v[t].z; s3 += v[t].w; }
    }
    double* mine_tot = tot + warp * kBnPartial + lane * 4;
    mine_tot[0] = s0; mine_tot[1] = s1; mine_tot[2] = s2; mine_tot[3] = s3;
    __syncthreads();
    if (threadIdx.x < kBnPartial) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += tot[w * kBnPartial + threadIdx.x];
        __syncwarp();
        tot[threadIdx.x] = t;                        // warp w' < 4 overwrites only slots its own lanes read last
    }
    if (threadIdx.x == 0) *counter = 0u;             // re-armed for the next launch on this stream
    __syncthreads();
    return true;
}

struct BnStatsArgs {
    const uint4* x;                  // [M, C / 8] vectors
    long long M, passes, ppc;        // ppc = passes per row chunk
    int C, R;
    float eps, momentum;
    float* partial;
    unsigned int* counters;
    float* mean;
    float* invstd;
    float* running_mean;             // nullable
    float* running_var;
    long long* num_batches_tracked;  // nullable
};

template <int kUnroll, int kCtas>
__global__ void __launch_bounds__(kBnThreads, kCtas)
bn_stats_kernel(const BnStatsArgs a) {
    __shared__ float red[8 * 128];
    __shared__ double tot[8 * kBnPartial];
    __shared__ int flag;
    const int slab = blockIdx.x, r = blockIdx.y;
    const int v = threadIdx.x & (kBnLanes - 1), rl = threadIdx.x >> 3;
    const int vec_per_row = a.C >> 3;
    const int colv = slab * kBnLanes + v;
    float sh[8];
    {
        const uint4 u = __ldg(a.x + colv);           // row 0: the shift
        unpack8(u, sh);
    }
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
    const long long p0 = (long long)r * a.ppc;
    const long long p1 = (p0 + a.ppc < a.passes) ? p0 + a.ppc : a.passes;
    for (long long p = p0; p < p1; p += kUnroll) {
        uint4 u[kUnroll];
        bool live[kUnroll];
#pragma unroll
        for (int t = 0; t < kUnroll; ++t) {
            const long long row = (p + t) * kBnRows + rl;
            live[t] = (p + t < p1) && row < a.M;
            u[t] = make_uint4(0u, 0u, 0u, 0u);
            if (live[t]) u[t] = __ldg(a.x + row * vec_per_row + colv);
        }
#pragma unroll
        for (int t = 0; t < kUnroll; ++t) {
            if (live[t]) {
                float f[8];
                unpack8(u[t], f);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float d = f[k] - sh[k];
                    acc[k] += d;
                    acc[8 + k] = fmaf(d, d, acc[8 + k]);
                }
            }
        }
    }
    const float part = slab_reduce(acc, red);
    if (!slab_finish(part, a.partial, a.counters + slab, slab, r, a.R, tot, &flag)) return;
    if (threadIdx.x < kBnSlab) {
        const int c8 = threadIdx.x >> 3, k = threadIdx.x & 7;
        const int c = slab * kBnSlab + threadIdx.x;
        const double s1 = tot[c8 * 16 + k], s2 = tot[c8 * 16 + 8 + k];
        const double inv_m = 1.0 / (double)a.M;
        const double md = s1 * inv_m;
        double var = s2 * inv_m - md * md;
        if (var < 0.0) var = 0.0;
        const float shift = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.x)[c]);
        const float mean = (float)((double)shift + md);
        a.mean[c] = mean;
        a.invstd[c] = (float)(1.0 / sqrt(var + (double)a.eps));
        if (a.running_mean != nullptr) {
            const double unbiased = a.M > 1 ? var * ((double)a.M / (double)(a.M - 1)) : var;
            a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * mean;
            a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (float)unbiased;
        }
    }
    if (slab == 0 && threadIdx.x == 0 && a.num_batches_tracked != nullptr) *a.num_batches_tracked += 1;
}

struct BnApplyArgs {
    const uint4* x;
    const uint4* res;                // nullable
    uint4* y;
    long long V;                     // M * C / 8
    int C, relu;
    const float* mean;
    const float* invstd;
    const float* gamma;
    const float* beta;
};

template <int kUnroll, int kCtas>
__global__ void __launch_bounds__(kBnThreads, kCtas)
bn_apply_kernel(const BnApplyArgs a) {
    const int lanes = a.C >> 3;                      // a power of two <= 256: this thread's channel group is fixed
    const int v = threadIdx.x & (lanes - 1);
    float ca[8], cb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = v * 8 + k;
        ca[k] = __ldg(a.gamma + c) * __ldg(a.invstd + c);
        cb[k] = fmaf(-__ldg(a.mean + c), ca[k], __ldg(a.beta + c));
    }
    const long long stride = (long long)gridDim.x * kBnThreads;
    const bool has_res = a.res != nullptr;
    for (long long i = (long long)blockIdx.x * kBnThreads + threadIdx.x; i < a.V; i += stride * kUnroll) {
        uint4 u[kUnroll], w[kUnroll];
#pragma unroll
        for (int t = 0; t < kUnroll; ++t) {
            const long long j = i + t * stride;
            u[t] = make_uint4(0u, 0u, 0u, 0u);
            w[t] = make_uint4(0u, 0u, 0u, 0u);
            if (j < a.V) {
                u[t] = __ldg(a.x + j);
                if (has_res) w[t] = __ldg(a.res + j);
            }
        }
#pragma unroll
        for (int t = 0; t < kUnroll; ++t) {
            const long long j = i + t * stride;
            if (j < a.V) {
                float f[8], g[8];
                unpack8(u[t], f);
                unpack8(w[t], g);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float z = fmaf(f[k], ca[k], cb[k]) + g[k];
                    if (a.relu) z = fmaxf(z, 0.f);
                    f[k] = z;
                }
                a.y[j] = pack8(f);
            }
        }
    }
}

// how the ReLU mask of the backward is obtained
enum { kMaskNone = 0, kMaskFromY = 1, kMaskFromX = 2 };

struct BnBwdReduceArgs {
    const uint4* dy;
    const uint4* x;
    const uint4* y;                  // kMaskFromY only
    long long M, passes, ppc;
    int C, R, mask;
    float* partial;
    unsigned int* counters;
    const float* mean;
    const float* invstd;
    const float* gamma;
    const float* beta;
    float* sum_dy;                   // = dbeta
    float* sum_dy_xhat;              // = dgamma
};

template <int kUnroll, int kCtas>
__global__ void __launch_bounds__(kBnThreads, kCtas)
bn_bwd_reduce_kernel(const BnBwdReduceArgs a) {
    __shared__ float red[8 * 128];
    __shared__ double tot[8 * kBnPartial];
    __shared__ int flag;
    const int slab = blockIdx.x, r = blockIdx.y;
    const int v = threadIdx.x & (kBnLanes - 1), rl = threadIdx.x >> 3;
    const int vec_per_row = a.C >> 3;
    const int colv = slab * kBnLanes + v;
    float mu[8], ca[8], cb[8];                       // sum(g x^) = invstd * sum(g (x - mean)): invstd applied at the end
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = colv * 8 + k;
        mu[k] = __ldg(a.mean + c);
        ca[k] = __ldg(a.gamma + c) * __ldg(a.invstd + c);
        cb[k] = fmaf(-mu[k], ca[k], __ldg(a.beta + c));
    }
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
    const long long p0 = (long long)r * a.ppc;
    const long long p1 = (p0 + a.ppc < a.passes) ? p0 + a.ppc : a.passes;
    for (long long p = p0; p < p1; p += kUnroll) {
        uint4 ud[kUnroll], ux[kUnroll], uy[kUnroll];
        bool live[kUnroll];
#pragma unroll
        for (int t = 0; t < kUnroll; ++t) {
            const long long row = (p + t) * kBnRows + rl;
            live[t] = (p + t < p1) && row < a.M;
            ud[t] = ux[t] = uy[t] = make_uint4(0u, 0u, 0u, 0u);
            if (live[t]) {
                const long long j = row * vec_per_row + colv;
                ud[t] = __ldg(a.dy + j);
                ux[t] = __ldg(a.x + j);
                if (a.mask == kMaskFromY) uy[t] = __ldg(a.y + j);
            }
        }
#pragma unroll
        for (int t = 0; t < kUnroll; ++t) {
            if (live[t]) {
                float d[8], f[8], yy[8];
                unpack8(ud[t], d);
                unpack8(ux[t], f);
                unpack8(uy[t], yy);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    bool on = true;
                    if (a.mask == kMaskFromY) on = yy[k] > 0.f;
                    else if (a.mask == kMaskFromX) on = fmaf(f[k], ca[k], cb[k]) > 0.f;
                    const float g = on ? d[k] : 0.f;
                    acc[k] += g;
                    acc[8 + k] = fmaf(g, f[k] - mu[k], acc[8 + k]);
                }
            }
        }
    }
    const float part = slab_reduce(acc, red);
    if (!slab_finish(part, a.partial, a.counters + slab, slab, r, a.R, tot, &flag)) return;
    if (threadIdx.x < kBnSlab) {
        const int c8 = threadIdx.x >> 3, k = threadIdx.x & 7;
        const int c = slab * kBnSlab + threadIdx.x;
        a.sum_dy[c] = (float)tot[c8 * 16 + k];
        a.sum_dy_xhat[c] = (float)(tot[c8 * 16 + 8 + k] * (double)a.invstd[c]);
    }
}

struct BnBwdApplyArgs {
    const uint4* dy;
    const uint4* x;
    const uint4* y;                  // kMaskFromY only
    uint4* dx;
    uint4* dres;                     // nullable: the masked gradient, for the residual branch
    long long V;
    int C, mask;
    float inv_m;
    const float* mean;
    const float* invstd;
    const float* gamma;
    const float* beta;
    const float* sum_dy;
    const float* sum_dy_xhat;
};

template <int kUnroll, int kCtas>
__global__ void __launch_bounds__(kBnThreads, kCtas)
bn_bwd_apply_kernel(const BnBwdApplyArgs a) {
    const int lanes = a.C >> 3;
    const int v = threadIdx.x & (lanes - 1);
    // dx = k1 * (g - m1 - x^ * m2),  k1 = gamma * invstd, m1 = sum(g) / M, m2 = sum(g x^) / M, x^ = (x - mean) * invstd
    //    = cA * g + cB * x + cD
    float cA[8], cB[8], cD[8], ca[8], cb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = v * 8 + k;
        const float mu = __ldg(a.mean + c), is = __ldg(a.invstd + c);
        const float k1 = __ldg(a.gamma + c) * is;
        const float m1 = __ldg(a.sum_dy + c) * a.inv_m, m2 = __ldg(a.sum_dy_xhat + c) * a.inv_m;
        cA[k] = k1;
        cB[k] = -k1 * m2 * is;
        cD[k] = fmaf(-cB[k], mu, -k1 * m1);
        ca[k] = k1;
        cb[k] = fmaf(-mu, k1, __ldg(a.beta + c));
    }
    const long long stride = (long long)gridDim.x * kBnThreads;
    for (long long i = (long long)blockIdx.x * kBnThreads + threadIdx.x; i < a.V; i += stride * kUnroll) {
        uint4 ud[kUnroll], ux[kUnroll], uy[kUnroll];
#pragma unroll
        for (int t = 0; t < kUnroll; ++t) {
            const long long j = i + t * stride;
            ud[t] = ux[t] = uy[t] = make_uint4(0u, 0u, 0u, 0u);
            if (j < a.V) {
                ud[t] = __ldg(a.dy + j);
                ux[t] = __ldg(a.x + j);
                if (a.mask == kMaskFromY) uy[t] = __ldg(a.y + j);
            }
        }
#pragma unroll
        for (int t = 0; t < kUnroll; ++t) {
            const long long j = i + t * stride;
            if (j < a.V) {
                float d[8], f[8], yy[8], o[8];
                unpack8(ud[t], d);
                unpack8(ux[t], f);
                unpack8(uy[t], yy);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    bool on = true;
                    if (a.mask == kMaskFromY) on = yy[k] > 0.f;
                    else if (a.mask == kMaskFromX) on = fmaf(f[k], ca[k], cb[k]) > 0.f;
                    const float g = on ? d[k] : 0.f;
                    d[k] = g;
                    o[k] = fmaf(cA[k], g, fmaf(cB[k], f[k], cD[k]));
                }
                a.dx[j] = pack8(o);
                if (a.dres != nullptr) a.dres[j] = pack8(d);
            }
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------

size_t bn_workspace_bytes() { return 256 + (size_t)kBnMaxCtas * kBnPartial * sizeof(float); }

static bool bn_shape_ok(long long M, int C) {
    if (M < 1 || C < kBnSlab || C > kBnSlab * kBnMaxSlabs) return false;
    return (C & (C - 1)) == 0;                       // 64 .. 2048, a power of two (256 % (C / 8) == 0)
}

static void bn_reduce_plan(long long M, int C, int unroll, int ctas_per_sm, long long* passes, long long* ppc, int* R) {
    const int slabs = C / kBnSlab;
    const long long np = (M + kBnRows - 1) / kBnRows;
    long long r = kBnSms * ctas_per_sm / slabs;
    const long long want = (np + unroll - 1) / unroll;            // at least one unrolled trip per CTA
    if (r > want) r = want;
    if (r < 1) r = 1;
    long long per = (np + r - 1) / r;
    per = (per + unroll - 1) / unroll * unroll;
    r = (np + per - 1) / per;
    *passes = np; *ppc = per; *R = (int)r;
}

static int bn_apply_grid(long long V, int unroll, int ctas_per_sm) {
    long long g = (V + (long long)kBnThreads * unroll - 1) / ((long long)kBnThreads * unroll);
    if (g > kBnSms * ctas_per_sm) g = kBnSms * ctas_per_sm;
    if (g < 1) g = 1;
    return (int)g;
}

template <int U, int CT>
static cudaError_t run_stats(BnStatsArgs& s, cudaStream_t stream) {
    bn_reduce_plan(s.M, s.C, U, CT, &s.passes, &s.ppc, &s.R);
    bn_stats_kernel<U, CT><<<dim3(s.C / kBnSlab, s.R), kBnThreads, 0, stream>>>(s);
    return cudaGetLastError();
}
template <int U, int CT>
static cudaError_t run_apply(BnApplyArgs& p, cudaStream_t stream) {
    bn_apply_kernel<U, CT><<<bn_apply_grid(p.V, U, CT), kBnThreads, 0, stream>>>(p);
    return cudaGetLastError();
}
template <int U, int CT>
static cudaError_t run_bwd_reduce(BnBwdReduceArgs& s, cudaStream_t stream) {
    bn_reduce_plan(s.M, s.C, U, CT, &s.passes, &s.ppc, &s.R);
    bn_bwd_reduce_kernel<U, CT><<<dim3(s.C / kBnSlab, s.R), kBnThreads, 0, stream>>>(s);
    return cudaGetLastError();
}
template <int U, int CT>
static cudaError_t run_bwd_apply(BnBwdApplyArgs& p, cudaStream_t stream) {
    bn_bwd_apply_kernel<U, CT><<<bn_apply_grid(p.V, U, CT), kBnThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_bn_fwd_train(const void* x, const void* res, void* y, long long M, int C, const float* gamma,
                                const float* beta, float* running_mean, float* running_var, long long* nbt, float momentum,
                                float eps, int relu, float* save_mean, float* save_invstd, void* ws, cudaStream_t stream) {
    if (!bn_shape_ok(M, C)) return cudaErrorNotSupported;
    BnStatsArgs s{};
    s.x = static_cast<const uint4*>(x);
    s.M = M; s.C = C;
    s.eps = eps; s.momentum = momentum;
    s.counters = static_cast<unsigned int*>(ws);
    s.partial = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + 256);
    s.mean = save_mean; s.invstd = save_invstd;
    s.running_mean = running_mean; s.running_var = running_var; s.num_batches_tracked = nbt;
    cudaError_t e = run_stats<kBnStatsUnroll, kBnStatsCtas>(s, stream);
    if (e != cudaSuccess) return e;
    BnApplyArgs p{};
    p.x = static_cast<const uint4*>(x); p.res = static_cast<const uint4*>(res); p.y = static_cast<uint4*>(y);
    p.V = M * (C >> 3); p.C = C; p.relu = relu;
    p.mean = save_mean; p.invstd = save_invstd; p.gamma = gamma; p.beta = beta;
    return run_apply<kBnApplyUnroll, kBnApplyCtas>(p, stream);
}

cudaError_t launch_bn_bwd(const void* dy, const void* x, const void* y, long long M, int C, const float* gamma,
                          const float* beta, const float* save_mean, const float* save_invstd, int relu, int has_residual,
                          void* dx, void* dres, float* dgamma, float* dbeta, void* ws, cudaStream_t stream) {
    if (!bn_shape_ok(M, C)) return cudaErrorNotSupported;
    const int mask = !relu ? kMaskNone : (has_residual ? kMaskFromY : kMaskFromX);
    if (mask == kMaskFromY && y == nullptr) return cudaErrorInvalidValue;
    BnBwdReduceArgs s{};
    s.dy = static_cast<const uint4*>(dy); s.x = static_cast<const uint4*>(x); s.y = static_cast<const uint4*>(y);
    s.M = M; s.C = C; s.mask = mask;
    s.counters = static_cast<unsigned int*>(ws);
    s.partial = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + 256);
    s.mean = save_mean; s.invstd = save_invstd; s.gamma = gamma; s.beta = beta;
    s.sum_dy = dbeta; s.sum_dy_xhat = dgamma;
    cudaError_t e = run_bwd_reduce<kBnBwdReduceUnroll, kBnBwdReduceCtas>(s, stream);
    if (e != cudaSuccess) return e;
    BnBwdApplyArgs p{};
    p.dy = s.dy; p.x = s.x; p.y = s.y; p.dx = static_cast<uint4*>(dx); p.dres = static_cast<uint4*>(dres);
    p.V = M * (C >> 3); p.C = C; p.mask = mask; p.inv_m = (float)(1.0 / (double)M);
    p.mean = save_mean; p.invstd = save_invstd; p.gamma = gamma; p.beta = beta;
    p.sum_dy = dbeta; p.sum_dy_xhat = dgamma;
    return run_bwd_apply<kBnBwdApplyUnroll, kBnBwdApplyCtas>(p, stream);
}

}  // namespace moco
