// ubench_latency.cu -- dependent-chain latencies of the instructions the FGK kernels are built from
// (one warp alone on an SM, clock64 around a chain of N dependent operations).  Design aid only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_latency ubench_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N 256

__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

__global__ void k(uint64_t *out, int mode)
{
    __shared__ uint32_t tab[2048];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) tab[i] = base + 4u * ((i * 17 + 5) & 2047);   // pointer chase
    __syncthreads();
    uint32_t x = base + 4 * lane, acc = 0;
    long long t0 = clock64();
    if (mode == 0) {            // LDS.32 pointer chase
#pragma unroll 16
        for (int i = 0; i < N; i++) x = lds32(x);
    } else if (mode == 1) {     // shfl (variable source lane)
#pragma unroll 16
        for (int i = 0; i < N; i++) x = __shfl_sync(0xffffffffu, x, x & 31);
    } else if (mode == 2) {     // ballot -> dependent compare
#pragma unroll 16
        for (int i = 0; i < N; i++) x = __ballot_sync(0xffffffffu, (x >> lane) & 1) + lane;
    } else if (mode == 3) {     // ballot + clz + shfl (one "tie" decision)
#pragma unroll 16
        for (int i = 0; i < N; i++) {
            uint32_t m = __ballot_sync(0xffffffffu, ((x >> lane) & 1) || lane == 3);
            uint32_t k0 = 31 - __clz(m);
            x = __shfl_sync(0xffffffffu, x + lane, k0) * 3 + 1;
        }
    } else if (mode == 4) {     // LDS.16 table lookup with shift/add address math (pt lookup)
#pragma unroll 16
        for (int i = 0; i < N; i++) { uint32_t e = lds16(base + 2 * ((x >> (8 - (lane & 7))) & 1023)); x = e * 8 + 3; }
    } else if (mode == 5) {     // LDS then __syncwarp then STS then LDS (publish + reread)
#pragma unroll 16
        for (int i = 0; i < N; i++) {
            uint32_t v = lds32(x);
            __syncwarp();
            if (lane == 0) asm volatile("st.shared.u32 [%0], %1;" :: "r"(x), "r"(v) : "memory");
            __syncwarp();
            x = v;
        }
    } else if (mode == 6) {     // dependent integer add chain
#pragma unroll 16
        for (int i = 0; i < N; i++) x = x * 3 + lane;
    } else if (mode == 7) {     // match-free: redux (warp reduce) chain
#pragma unroll 16
        for (int i = 0; i < N; i++) x = __reduce_max_sync(0xffffffffu, x ^ lane) + 1;
    } else if (mode == 8) {     // uniform-address LDS (broadcast) chase
        x = base;
#pragma unroll 16
        for (int i = 0; i < N; i++) x = lds32(x);
    } else if (mode == 9) {     // divergent single-lane section then reconverge
#pragma unroll 4
        for (int i = 0; i < N; i++) {
            if (lane == (x & 7)) { acc += lds32(x); }
            __syncwarp();
            x = x * 5 + 1;
            x = base + 4 * (x & 2047);
        }
    }
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) { out[mode * 2] = (uint64_t)(t1 - t0); out[mode * 2 + 1] = x + acc; }
}

int main()
{
    uint64_t *d, h[32];
    cudaMalloc(&d, sizeof h);
    const char *names[] = {"LDS.32 chase", "SHFL idx", "VOTE.ballot", "ballot+clz+shfl", "LDS.16 lookup (shift,add,lea)", "LDS+warpsync+STS+warpsync",
                           "IMAD chain", "REDUX.max", "LDS.32 uniform chase", "single-lane branch + syncwarp"};
    for (int rep = 0; rep < 2; rep++)
        for (int m = 0; m < 10; m++) k<<<1, 32>>>(d, m);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    for (int m = 0; m < 10; m++) printf("%-36s %7.1f cycles/op\n", names[m], (double)h[2 * m] / N);
    return cudaGetLastError() != cudaSuccess;
}
