// Memory-pattern probe: what does the B200 memory system deliver for "gather ~95-byte packed pieces from a
// 3.1 GB array, write 2x as much sequentially" with a trivial kernel?  (scratch; not part of the product)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <random>
#include <cmath>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)

template <int F> __device__ __forceinline__ uint2 ldf(const uint2* p) {
    uint2 v;
    if (F == 0) v = __ldg(p);
    else if (F == 1) asm volatile("ld.global.ca.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    else if (F == 2) asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    else if (F == 3) asm volatile("ld.global.cs.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    else if (F == 4) asm volatile("ld.global.lu.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    else if (F == 5) asm volatile("ld.global.cv.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    else if (F == 6) asm volatile("ld.global.nc.L2::64B.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    else if (F == 7) asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    else if (F == 8) asm volatile("ld.global.nc.L1::evict_first.L2::64B.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    else if (F == 9) { uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol)); }
    else if (F == 10) asm volatile("ld.global.nc.L2::256B.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
template <int F>
__global__ void k_gatherF(const uint2* __restrict__ src, const int64_t* __restrict__ idx, int64_t n, uint4* __restrict__ out) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
        const uint2 v = ldf<F>(src + idx[c]);
        uint4 o = make_uint4(v.x, v.y, v.x ^ 0x55555555u, v.y ^ 0x33333333u);
        asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(out + c), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
    }
}
__global__ void k_gather(const uint2* __restrict__ src, const int64_t* __restrict__ idx, int64_t n, uint4* __restrict__ out) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
        const uint2 v = __ldg(src + idx[c]);
        uint4 o = make_uint4(v.x, v.y, v.x ^ 0x55555555u, v.y ^ 0x33333333u);
        asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(out + c), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
    }
}
__global__ void k_stream(const uint2* __restrict__ src, int64_t n, uint4* __restrict__ out) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
        const uint2 v = __ldg(src + c);
        uint4 o = make_uint4(v.x, v.y, v.x ^ 0x55555555u, v.y ^ 0x33333333u);
        asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(out + c), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
    }
}
int main(int argc, char** argv) {
    const int64_t genome_bytes = 3100000000ll;            // two planes of 1.55 GB
    const int64_t out_bytes = 365000000ll;                // exon launch of the bench
    const double mean_piece = argc > 1 ? atof(argv[1]) : 190.0;   // output bytes per piece
    const int64_t n = out_bytes / 16;
    std::vector<int64_t> idx(n);
    std::mt19937_64 rng(4);
    std::lognormal_distribution<double> ln(std::log(140.0), 1.3);
    int64_t c = 0;
    while (c < n) {
        double L = mean_piece > 0 ? std::max(3.0, ln(rng)) : 1e18;
        int64_t chunks = std::max<int64_t>(1, (int64_t)(L / 16));
        int64_t start = (int64_t)(rng() % (uint64_t)(genome_bytes / 8 - chunks - 8));
        for (int64_t k = 0; k < chunks && c < n; k++) idx[c++] = start + k;
    }
    uint2* d_src; int64_t* d_idx; uint4* d_out;
    CK(cudaMalloc(&d_src, genome_bytes)); CK(cudaMemset(d_src, 1, genome_bytes));
    CK(cudaMalloc(&d_idx, n * 8)); CK(cudaMemcpy(d_idx, idx.data(), n * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_out, n * 16));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (argc > 2) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[2]));
    for (int F = 0; F <= 10; F++) {
        float best = 1e9;
        for (int it = 0; it < 5; it++) {
            cudaEventRecord(e0);
            switch (F) {
#define C(F) case F: k_gatherF<F><<<148 * 16, 256>>>(d_src, d_idx, n, d_out); break;
            C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10)
            }
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (it >= 2 && ms < best) best = ms;
        }
        printf("flavor %d: %.1f us\n", F, best * 1e3);
    }
    for (int mode = 0; mode < 2; mode++) {
        float best = 1e9;
        for (int it = 0; it < 8; it++) {
            cudaEventRecord(e0);
            if (mode == 0) k_gather<<<148 * 16, 256>>>(d_src, d_idx, n, d_out);
            else k_stream<<<148 * 16, 256>>>(d_src, n, d_out);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (it >= 2 && ms < best) best = ms;
        }
        const double alg = n * (8.0 + 16.0) + (mode == 0 ? n * 8.0 : 0);
        printf("%s: %.1f us for %.0f MB out  -> %.0f GB/s useful (8 B read + 16 B written per chunk%s)\n", mode == 0 ? "gather" : "stream",
               best * 1e3, out_bytes / 1e6, alg / best / 1e6, mode == 0 ? " + 8 B index" : "");
    }
    return 0;
}
