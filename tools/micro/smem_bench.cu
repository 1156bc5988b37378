// Microbenchmarks behind two design decisions (run on the GPU box: `make -C tools/micro && gpurun_out/smem_bench`):
//  A. what one SM sustains for lane-private shared-memory histogram updates (the nhood count kernel's inner operation):
//     red.shared.add with an immediate 1 (SASS: ATOMS.POPC.INC), with a register operand (ATOMS.ADD), a non-atomic
//     LDS + STS pair (upper bound if the columns were warp-private), and two packed 16-bit counters per word;
//  B. random single-byte read + write against (1) the CTA's own shared memory, (2) distributed shared memory of a cluster
//     of 8 CTAs (a 1 MB label array spread over 8 SMs), (3) an L2-resident global array: the inner operation of the exact
//     Fisher-Yates replay.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/smem_bench tools/micro/smem_bench.cu
#include <cooperative_groups.h>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t lcg(uint32_t& x) {
    x = x * 1664525u + 1013904223u;
    return x;
}

template <int MODE>
__global__ void __launch_bounds__(1024) hist_kernel(int iters, int nbins, uint32_t* out, long long* cyc) {
    extern __shared__ __align__(16) uint32_t hist[];
    for (int i = threadIdx.x; i < nbins * 32; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(hist) + lane * 4u;
    uint32_t one = 1;
    asm volatile("" : "+r"(one));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = base + (__umulhi(lcg(x), (uint32_t)nbins) << 7);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a[u]));
            if (MODE == 1) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a[u]), "r"(one));
            if (MODE == 2) {
                uint32_t v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a[u]));
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(a[u]), "r"(v + 1));
            }
            if (MODE == 3) {  // ALU only: the address generation without any shared-memory operation
                x ^= a[u];
            }
            if (MODE == 4) asm volatile("red.shared.inc.u32 [%0], %1;" ::"r"(a[u]), "r"(0xFFFFFFFFu));
            if (MODE == 5) {
                uint32_t v;
                asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(v) : "r"(a[u]), "r"(one));
                x ^= v;
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    uint32_t s = x;
    for (int i = threadIdx.x; i < nbins * 32; i += blockDim.x) s += hist[i];
    if (s == 0x12345678u) out[0] = s;
}

// B: random byte read-modify-write.  WHERE 0: own shared memory (128 KB), 1: cluster of 8 x 128 KB distributed shared memory,
// 2: global array of 1 MB per CTA (L2 resident)
template <int WHERE>
__global__ void __launch_bounds__(1024) byte_kernel(int iters, uint8_t* g, uint32_t* out, long long* cyc) {
    extern __shared__ __align__(16) uint8_t sm[];
    constexpr uint32_t SLAB = 128 * 1024;
    for (int i = threadIdx.x; i < (int)SLAB; i += blockDim.x) sm[i] = (uint8_t)i;
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 97u;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
    uint8_t* gp = g + (size_t)blockIdx.x * (1u << 20);
    if (WHERE == 1) cg::this_cluster().sync();
    else __syncthreads();
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t r[8], v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) r[u] = lcg(x) >> 12;  // 20 bits
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (WHERE == 0) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v[u]) : "r"(sbase + (r[u] & (SLAB - 1))));
            if (WHERE == 1) {
                uint32_t ra;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(sbase + (r[u] & (SLAB - 1))), "r"(r[u] >> 17));
                asm volatile("ld.shared::cluster.u8 %0, [%1];" : "=r"(v[u]) : "r"(ra));
                r[u] = ra;
            }
            if (WHERE == 2) asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v[u]) : "l"(gp + r[u]));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            acc += v[u];
            if (WHERE == 0) asm volatile("st.shared.u8 [%0], %1;" ::"r"(sbase + (r[u] & (SLAB - 1))), "r"(v[u] + 1));
            if (WHERE == 1) asm volatile("st.shared::cluster.u8 [%0], %1;" ::"r"(r[u]), "r"(v[u] + 1));
            if (WHERE == 2) asm volatile("st.global.cg.u8 [%0], %1;" ::"l"(gp + r[u]), "r"(v[u] + 1));
        }
    }
    const long long t1 = clock64();
    if (WHERE == 1) cg::this_cluster().sync();
    else __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) out[0] = acc;
}

static double avg_cycles(long long* d, int n) {
    long long h[1024];
    cudaMemcpy(h, d, n * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < n; ++i) s += (double)h[i];
    return s / n;
}

int main() {
    uint32_t* out;
    long long* cyc;
    uint8_t* g;
    cudaMalloc(&out, 4);
    cudaMalloc(&cyc, 1024 * 8);
    cudaMalloc(&g, (size_t)160 << 20);
    cudaMemset(g, 0, (size_t)160 << 20);
    const int nbins = 900, iters = 2000;
    const size_t hs = (size_t)nbins * 128;
    const char* names[] = {"red.add imm 1 (ATOMS.POPC.INC)", "red.add reg (ATOMS.ADD)", "LDS+STS non-atomic", "ALU only", "red.inc", "atom.add with return"};
#define RUN_H(M)                                                                                                        \
    {                                                                                                                   \
        cudaFuncSetAttribute(hist_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hs);                     \
        hist_kernel<M><<<148, 1024, hs>>>(10, nbins, out, cyc);                                                         \
        hist_kernel<M><<<148, 1024, hs>>>(iters, nbins, out, cyc);                                                      \
        cudaError_t e = cudaDeviceSynchronize();                                                                        \
        const double c = avg_cycles(cyc, 148);                                                                          \
        printf("A%d %-34s: %.2f cycles per warp-op per SM (%s)\n", M, names[M], c / ((double)iters * 8 * 32), cudaGetErrorString(e)); \
    }
    RUN_H(0) RUN_H(1) RUN_H(2) RUN_H(3) RUN_H(4) RUN_H(5)
    const size_t bs = 128 * 1024;
    {
        cudaFuncSetAttribute(byte_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bs);
        byte_kernel<0><<<148, 1024, bs>>>(10, g, out, cyc);
        byte_kernel<0><<<148, 1024, bs>>>(iters, g, out, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        printf("B0 own shared memory byte RMW      : %.2f cycles per warp-RMW per SM (%s)\n", avg_cycles(cyc, 148) / ((double)iters * 8 * 32), cudaGetErrorString(e));
    }
    {
        cudaFuncSetAttribute(byte_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bs);
        cudaFuncSetAttribute(byte_kernel<1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(144), cfg.blockDim = dim3(1024), cfg.dynamicSmemBytes = bs;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 8, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
        cfg.attrs = at, cfg.numAttrs = 1;
        int it0 = 10;
        cudaLaunchKernelEx(&cfg, byte_kernel<1>, it0, g, out, cyc);
        int it1 = iters;
        cudaLaunchKernelEx(&cfg, byte_kernel<1>, it1, g, out, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        printf("B1 DSMEM cluster of 8 byte RMW     : %.2f cycles per warp-RMW per SM (%s)\n", avg_cycles(cyc, 144) / ((double)iters * 8 * 32), cudaGetErrorString(e));
    }
    {
        cudaFuncSetAttribute(byte_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bs);
        byte_kernel<2><<<148, 1024, bs>>>(10, g, out, cyc);
        byte_kernel<2><<<148, 1024, bs>>>(iters, g, out, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        printf("B2 global (L2 resident) byte RMW   : %.2f cycles per warp-RMW per SM (%s)\n", avg_cycles(cyc, 148) / ((double)iters * 8 * 32), cudaGetErrorString(e));
    }
    return 0;
}
