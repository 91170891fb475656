// random_access.cu -- measured ceiling for the hash-table kernels: independent random reads of
// B bytes (8/16/32/64) over a footprint of F bytes, fully occupied grid, U loads in flight per
// thread.  Prints G accesses/s and GB/s.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x;
}
template <int B, int U, bool ATOMIC>
__global__ void __launch_bounds__(256) probe(const uint64_t* __restrict__ t, uint64_t mask_units, int iters, uint64_t* out) {
    uint64_t acc = 0;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        uint64_t v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t r = mix(tid * 1315423911ull + (uint64_t)it * U + u) & mask_units;  // unit = B bytes
            const uint64_t* p = t + r * ((B == 256 ? 32 : B) / 8);
            if (ATOMIC) {
                v[u] = atomicCAS((unsigned long long*)p, 0ull, (unsigned long long)(tid | 1));
            } else if (B == 256) {  // one 256-bit load instruction (sm_100)
                uint64_t a0, a1, a2, a3;
                asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a0), "=l"(a1), "=l"(a2), "=l"(a3) : "l"(p));
                v[u] = a0 ^ a1 ^ a2 ^ a3;
            } else if (B == 8) v[u] = __ldcg(p);
            else if (B == 16) { ulonglong2 q = __ldcg((const ulonglong2*)p); v[u] = q.x ^ q.y; }
            else { ulonglong2 q = __ldcg((const ulonglong2*)p); ulonglong2 q2 = __ldcg((const ulonglong2*)p + 1); v[u] = q.x ^ q.y ^ q2.x ^ q2.y;
                   if (B == 64) { ulonglong2 q3 = __ldcg((const ulonglong2*)p + 2); ulonglong2 q4 = __ldcg((const ulonglong2*)p + 3); v[u] ^= q3.x ^ q4.y; } }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u];
    }
    if (acc == 0x1234567) out[0] = acc;
}
template <int B, int U, bool ATOMIC>
void run(uint64_t* t, uint64_t bytes, uint64_t* out, int sms) {
    const int blocks = sms * 8, iters = 256 / U;
    const uint64_t units = bytes / (B == 256 ? 32 : B);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    probe<B, U, ATOMIC><<<blocks, 256>>>(t, units - 1, 8, out);
    cudaEventRecord(a);
    probe<B, U, ATOMIC><<<blocks, 256>>>(t, units - 1, iters, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double n = (double)blocks * 256 * iters * U;
    printf("{\"footprint_GiB\": %.2f, \"bytes\": %d, \"inflight_per_thread\": %d, \"atomic\": %d, \"G_access_per_s\": %.2f, \"GB_per_s\": %.1f}\n",
           bytes / 1073741824.0, B, U, (int)ATOMIC, n / ms * 1e-6, n * B / ms * 1e-6);
}
int main(int argc, char** argv) {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    uint64_t* out; cudaMalloc(&out, 8);
    for (uint64_t gib : {4ull, 32ull}) {
        uint64_t bytes = gib << 30; uint64_t* t;
        if (cudaMalloc(&t, bytes) != cudaSuccess) break;
        cudaMemset(t, 0, bytes);
        run<8, 1, false>(t, bytes, out, p.multiProcessorCount);
        run<32, 1, false>(t, bytes, out, p.multiProcessorCount);
        run<32, 4, false>(t, bytes, out, p.multiProcessorCount);
        run<64, 4, false>(t, bytes, out, p.multiProcessorCount);
        run<16, 8, false>(t, bytes, out, p.multiProcessorCount);
        run<8, 4, true>(t, bytes, out, p.multiProcessorCount);
        run<256, 1, false>(t, bytes, out, p.multiProcessorCount);
        run<256, 4, false>(t, bytes, out, p.multiProcessorCount);
        cudaFree(t);
    }
    return 0;
}
