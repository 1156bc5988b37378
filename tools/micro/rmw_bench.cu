// Microbenchmark: what does the B200 memory system sustain for the access pattern of an exact Fisher-Yates replay?
// Every "step" = read one random byte + one sequential byte of a private 1 MB array, write both back (a swap), for
// P arrays (P MB total), fully parallel and with no dependencies other than load->store of the same step.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/rmw_bench tools/micro/rmw_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// each warp owns one array; lane handles steps lane, lane+32, ...; UNR independent swaps in flight per lane
template <int UNR>
__global__ void rmw_kernel(uint8_t* __restrict__ a, int n, int steps_per_part, int n_arrays, int parts) {
    const int lane = threadIdx.x & 31;
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    for (int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < n_arrays * parts; v += warps_total) {
        const int arr = v % n_arrays, part = v / n_arrays;
        uint8_t* p = a + (size_t)arr * n;
        for (int s0 = part * steps_per_part; s0 < (part + 1) * steps_per_part; s0 += 32 * UNR) {
            uint32_t i[UNR], j[UNR]; uint8_t vi[UNR], vj[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int s = s0 + u * 32 + lane;
                i[u] = (uint32_t)(n - 1 - s) % (uint32_t)n;             // descending, coalesced
                j[u] = hash32((uint32_t)s * 2654435761u + arr) % (uint32_t)n;  // random target
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) { vi[u] = __ldcg(p + i[u]); vj[u] = __ldcg(p + j[u]); }
#pragma unroll
            for (int u = 0; u < UNR; ++u) { p[i[u]] = vj[u]; p[j[u]] = vi[u]; }
        }
    }
}

int main() {
    const int n = 1 << 20;
    for (int P : {16, 32, 48, 64, 96, 128, 160, 192, 256, 384, 512, 1000}) {
        uint8_t* a; cudaMalloc(&a, (size_t)P * n); cudaMemset(a, 1, (size_t)P * n);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int wpa : {16}) {            // warps per array: 1 = P warps in flight, 8 = 8P (arrays revisited by 8 warps)
            for (int unr : {4}) {
                const int n_arr_eff = P * wpa;
                const int blocks = (n_arr_eff * 32 + 255) / 256;
                const int steps = n / wpa;
                float best = 1e9f;
                for (int rep = 0; rep < 5; ++rep) {
                    cudaEventRecord(e0);
                    if (unr == 4) rmw_kernel<4><<<blocks, 256>>>(a, n, steps, P, wpa);
                    else rmw_kernel<16><<<blocks, 256>>>(a, n, steps, P, wpa);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
                }
                // note: with wpa = 8 the grid has 8x the warps but they stride over the same P arrays -> each array visited once per warp-stride; count work actually done:
                const double warps = (double)blocks * 8, total = (double)P * n;
                printf("P=%d arrays(MB) warps=%.0f unr=%d : %.3f ms -> %.2f Gsteps/s (%.2f ms per 1e9 steps)\n", P, warps, unr, best,
                       total / best / 1e6, best * 1e9 / total);
            }
        }
        cudaFree(a);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
