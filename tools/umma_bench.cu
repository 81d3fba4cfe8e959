// Microbenchmark: issue-limited throughput of tcgen05.mma (kind::f16, bf16, M=128, cta_group::1) per SM for
// the operand forms / tile widths the NCE kernels can use.  Operands are whatever is in smem/TMEM (values do not
// matter for timing).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I moco_b200/csrc -o tools/umma_bench tools/umma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"

using namespace moco;

// MODE 0: SS (A and B from smem, both K-major)   MODE 1: TS (A from TMEM, B from smem K-major)
// MODE 2: SS with B MN-major (the P.V form)      MODE 3: TS with B MN-major
template <int MODE, int N, int CE, int BG = 0>
__global__ void __launch_bounds__(384, 1) k(long long* cycles, int iters) {
    __shared__ __align__(8) uint64_t spin_bar;
    __shared__ volatile int done_flag;
    __shared__ __align__(8) uint64_t dummy[4];
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (64 + 128) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (warp == 0) { tmem_alloc<1>(&slot, 512); tmem_relinquish<1>(); }
    if (threadIdx.x == 32) { mbar_init(&spin_bar, 1); done_flag = 0; mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); for (int j = 0; j < 4; ++j) mbar_init(&dummy[j], 1); fence_mbar_init(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1 && lane == 0) {
        const uint32_t a_addr = smem_u32(smem);                 // A: 128 rows x 256 K (4 slabs of 16 KB)
        const uint32_t b_addr = smem_u32(smem + 64 * 1024);     // B: up to 256 rows x 256 K (4 slabs of 32 KB)
        const uint32_t idesc = make_idesc_bf16(128, N, 0, (MODE >= 2) ? 1 : 0);
        // `ready` completes once here and stays complete for parity 0: models waiting on an already-full stage
        __shared__ __align__(8) uint64_t ready;
        mbar_init(&ready, 1); fence_mbar_init(); mbar_arrive(&ready);
        long long issue_cycles = 0;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int b = i & 1;
            if (i >= 2) mbar_wait(&bar[b], (uint32_t)((i >> 1) - 1) & 1u);
            const uint32_t d = tmem + 256 + (N > 128 ? 0u : (uint32_t)(b * 128));
            long long c0 = clock64();
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) {
                if (CE == 3 && (ks & 3) == 0) { mbar_wait(&ready, 0); tc_fence_after(); }
                uint64_t bdesc = (MODE >= 2) ? make_sw128_desc(b_addr + (ks & 7) * 2048, 16384, 1024)
                                             : make_sw128_desc(b_addr + (ks >> 2) * 32768 + (ks & 3) * 32, 0, 1024);
                if (MODE == 0 || MODE == 2)
                    umma_ss<1>(d, make_sw128_desc(a_addr + (ks >> 2) * 16384 + (ks & 3) * 32, 0, 1024), bdesc, idesc, ks != 0);
                else
                    umma_ts<1>(d, tmem + (uint32_t)(ks * 8), bdesc, idesc, ks != 0);
                if (CE < 16 && ((ks % (CE == 3 ? 4 : CE)) == (CE == 3 ? 4 : CE) - 1) && ks != 15) umma_commit<1>(&dummy[(ks / (CE == 3 ? 4 : CE)) & 3]);   // extra commits (stage releases)
            }
            umma_commit<1>(&bar[b]);
            if (i == 0) issue_cycles = clock64() - c0;
        }
        cycles[148 + blockIdx.x] = issue_cycles;
        done_flag = 1;
        mbar_arrive(&spin_bar);
        mbar_wait(&bar[(iters - 1) & 1], (uint32_t)((iters - 1) >> 1) & 1u);
        if (iters > 1) mbar_wait(&bar[(iters - 2) & 1], (uint32_t)((iters - 2) >> 1) & 1u);
        long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    if (warp >= 4) {
        if (BG == 1) {                       // epilogue-style pollers
            mbar_wait(&spin_bar, 0);
        } else if (BG == 2) {                // epilogue-style TMEM drain
            uint32_t acc = 0, r[32];
            const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + (uint32_t)((warp >> 2) & 1) * 64;
            while (!done_flag) { tmem_ld32(base, r); tmem_ld_wait(); acc ^= r[0] ^ r[31]; }
            if (acc == 0x12345678u) cycles[0] = 0;
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<1>(tmem, 512);
}

template <int MODE, int N, int CE = 16, int BG = 0>
void run(const char* name) {
    int sms = 148, iters = 400;
    long long* cyc;
    cudaMalloc(&cyc, sizeof(long long) * sms * 2);
    int smem = (64 + 128) * 1024 + 1024;
    cudaFuncSetAttribute(k<MODE, N, CE, BG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<MODE, N, CE, BG><<<sms, 384, smem>>>(cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-34s ERROR %s\n", name, cudaGetErrorString(e)); exit(1); }
    k<MODE, N, CE, BG><<<sms, 384, smem>>>(cyc, iters);
    cudaDeviceSynchronize();
    long long h[296];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < sms; ++i) c += h[i]; c /= sms;
    double per = c / (iters * 16.0);
    printf("%-34s commit/%2d N=%3d  %.1f cycles/MMA (ideal %d)  -> %.0f%% of tensor peak; first 16 issues returned after %lld cycles\n", name, CE, N, per, N / 2, 100.0 * (N / 2) / per, h[148]);
    cudaFree(cyc);
}


// Two issuing threads (warps 1 and 2), each with its own accumulator and commit barriers: does the ~66-cycle
// per-instruction issue cost of small-N MMAs belong to the thread or to the SM?
template <int N, int NI>
__global__ void __launch_bounds__(384, 1) k2(long long* cycles, int iters) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bar[2][2];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (64 + 128) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (warp == 0) { tmem_alloc<1>(&slot, 512); tmem_relinquish<1>(); }
    if (threadIdx.x == 32) { for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) mbar_init(&bar[a][b], 1); fence_mbar_init(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if ((warp == 1 || (warp == 2 && NI == 2)) && lane == 0) {
        const int me = warp - 1;
        const uint32_t a_addr = smem_u32(smem);
        const uint32_t b_addr = smem_u32(smem + 64 * 1024);
        const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int b = i & 1;
            if (i >= 2) mbar_wait(&bar[me][b], (uint32_t)((i >> 1) - 1) & 1u);
            const uint32_t d = tmem + (uint32_t)(me * 256 + b * 128);
#pragma unroll
            for (int ks = 0; ks < 16; ++ks)
                umma_ss<1>(d, make_sw128_desc(a_addr + (ks >> 2) * 16384 + (ks & 3) * 32, 0, 1024),
                           make_sw128_desc(b_addr + (ks >> 2) * 32768 + (ks & 3) * 32, 0, 1024), idesc, ks != 0);
            umma_commit<1>(&bar[me][b]);
        }
        mbar_wait(&bar[me][(iters - 1) & 1], (uint32_t)((iters - 1) >> 1) & 1u);
        if (iters > 1) mbar_wait(&bar[me][(iters - 2) & 1], (uint32_t)((iters - 2) >> 1) & 1u);
        cycles[me * 148 + blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<1>(tmem, 512);
}

template <int N, int NI>
void run2(const char* name) {
    int sms = 148, iters = 400;
    long long* cyc;
    cudaMalloc(&cyc, sizeof(long long) * sms * 2);
    cudaMemset(cyc, 0, sizeof(long long) * sms * 2);
    int smem = (64 + 128) * 1024 + 1024;
    cudaFuncSetAttribute(k2<N, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; ++rep) {
        k2<N, NI><<<sms, 384, smem>>>(cyc, iters);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-34s ERROR %s\n", name, cudaGetErrorString(e)); exit(1); }
    }
    long long h[296];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < sms; ++i) c += (h[i] > h[148 + i] ? h[i] : h[148 + i]); c /= sms;
    double per = c / (iters * 16.0 * NI);
    printf("%-34s issuers=%d N=%3d  %.1f cycles/MMA aggregate (ideal %d) -> %.0f%% of tensor peak\n", name, NI, N, per, N / 2,
           100.0 * (N / 2) / per);
    cudaFree(cyc);
}

int main() {
    run2<64, 1>("one issuing thread");
    run2<64, 2>("two issuing threads");
    run2<128, 1>("one issuing thread");
    run2<128, 2>("two issuing threads");
    run2<256, 2>("two issuing threads");
    if (getenv("UMMA_BENCH_ONLY2")) return 0;
    run<0, 256, 3, 1>("SS + 8 warps polling mbarrier");
    run<0, 256, 3, 2>("SS + 8 warps streaming tcgen05.ld");
    run<0, 256, 3>("SS + wait/fence every 4 MMAs");
    run<0, 256, 4>("SS  A,B K-major");
    run<0, 256, 8>("SS  A,B K-major");
    run<0, 256, 2>("SS  A,B K-major");
    run<1, 256, 4>("TS  A tmem, B K-major");
    run<0, 256>("SS  A,B K-major");
    run<0, 128>("SS  A,B K-major");
    run<0, 64>("SS  A,B K-major");
    run<1, 256>("TS  A tmem, B K-major");
    run<1, 128>("TS  A tmem, B K-major");
    run<1, 64>("TS  A tmem, B K-major");
    run<2, 256>("SS  B MN-major");
    run<2, 128>("SS  B MN-major");
    run<3, 256>("TS  B MN-major");
    run<3, 128>("TS  B MN-major");
    return 0;
}
