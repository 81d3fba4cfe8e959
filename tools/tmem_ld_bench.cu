// Microbenchmark: TMEM -> register drain bandwidth per SM for the tcgen05.ld shapes the NCE epilogue could use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_ld_bench tools/tmem_ld_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int SHAPE>
__device__ __forceinline__ uint32_t ld(uint32_t taddr) {
    uint32_t acc = 0;
    if constexpr (SHAPE == 0) {          // 32x32b.x32 : 32 lanes x 32 columns (4 KB per warp instruction)
        uint32_t r[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
              "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= r[j];
    } else if constexpr (SHAPE == 1) {   // 32x32b.x16 : 2 KB per warp instruction
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) acc ^= r[j];
    } else if constexpr (SHAPE == 2) {   // 16x256b.x8 : 16 lanes x 256 bit x 8 = 4 KB per warp instruction
        uint32_t r[32];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
              "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= r[j];
    } else {                             // 32x32b.x32 issued twice before one wait (8 KB in flight per warp)
        uint32_t r[32], q[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
              "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(q[0]),"=r"(q[1]),"=r"(q[2]),"=r"(q[3]),"=r"(q[4]),"=r"(q[5]),"=r"(q[6]),"=r"(q[7]),"=r"(q[8]),"=r"(q[9]),"=r"(q[10]),"=r"(q[11]),"=r"(q[12]),"=r"(q[13]),"=r"(q[14]),"=r"(q[15]),
              "=r"(q[16]),"=r"(q[17]),"=r"(q[18]),"=r"(q[19]),"=r"(q[20]),"=r"(q[21]),"=r"(q[22]),"=r"(q[23]),"=r"(q[24]),"=r"(q[25]),"=r"(q[26]),"=r"(q[27]),"=r"(q[28]),"=r"(q[29]),"=r"(q[30]),"=r"(q[31])
            : "r"(taddr + 32));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= r[j] ^ q[j];
    }
    return acc;
}

template <int SHAPE>
__global__ void k(uint32_t* out, int iters, long long* cycles) {
    __shared__ uint32_t slot;
    int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) acc ^= ld<SHAPE>(base + (uint32_t)(((i * 64) + (warp >> 2) * 64) & 255));
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

template <int SHAPE>
void run(const char* name, int warps, int bytes_per_instr) {
    int sms = 148, iters = 2000;
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, sizeof(uint32_t) * sms * warps * 32);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    k<SHAPE><<<sms, warps * 32>>>(out, iters, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-26s warps=%2d  ERROR %s\n", name, warps, cudaGetErrorString(e)); return; }
    k<SHAPE><<<sms, warps * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < sms; ++i) c += h[i]; c /= sms;
    double bytes = (double)warps * iters * bytes_per_instr;
    printf("%-26s warps=%2d  %.0f cycles  %.1f B/clk/SM\n", name, warps, c, bytes / c);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int warps : {4, 8, 16}) {
        run<0>("32x32b.x32 (+wait each)", warps, 4096);
        run<1>("32x32b.x16 (+wait each)", warps, 2048);
        run<2>("16x256b.x8 (+wait each)", warps, 4096);
        run<3>("2 x 32x32b.x32 per wait", warps, 8192);
    }
    return 0;
}
