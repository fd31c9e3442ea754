// mg_scan.cu -- K1 building block: exclusive prefix sum int32 -> int64 with warp/block scans
// (three launches: per-block reduce, scan of block sums, per-block scan + offset).  This is the
// device replacement of the implicit length bookkeeping of `"".join(seq_list)` (genome.py:705).
#include "mg_common.cuh"

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ int64_t warp_incl_scan(int64_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// inclusive scan across the block; *total receives the block sum.  smem: 8 x int64 + 1
__device__ __forceinline__ int64_t block_incl_scan(int64_t v, int64_t *s_warp, int64_t *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_incl_scan(v);
    if (lane == 31) s_warp[wid] = v;
    __syncthreads();
    if (wid == 0) {
        int64_t w = lane < (SCAN_THREADS / 32) ? s_warp[lane] : 0;
        w = warp_incl_scan(w);
        if (lane < (SCAN_THREADS / 32)) s_warp[lane] = w;
    }
    __syncthreads();
    const int64_t off = wid ? s_warp[wid - 1] : 0;
    *total = s_warp[SCAN_THREADS / 32 - 1];
    __syncthreads();
    return v + off;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const int32_t *__restrict__ in, int64_t n, int64_t *__restrict__ bsum) {
    __shared__ int64_t s_warp[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int64_t v = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        const int64_t i = base + j * SCAN_THREADS + threadIdx.x;
        if (i < n) v += in[i];
    }
    int64_t total;
    block_incl_scan(v, s_warp, &total);
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

// single block: exclusive scan of nb block sums in place, bsum[nb] = grand total
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_bsums(int64_t *__restrict__ bsum, int64_t nb) {
    __shared__ int64_t s_warp[SCAN_THREADS / 32];
    int64_t carry = 0;
    for (int64_t b0 = 0; b0 < nb; b0 += SCAN_THREADS) {
        const int64_t i = b0 + threadIdx.x;
        const int64_t v = i < nb ? bsum[i] : 0;
        int64_t total;
        const int64_t inc = block_incl_scan(v, s_warp, &total);
        if (i < nb) bsum[i] = carry + inc - v;
        carry += total;
    }
    if (threadIdx.x == 0) bsum[nb] = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const int32_t *__restrict__ in, int64_t n,
                                                             const int64_t *__restrict__ bsum, int64_t nb,
                                                             int64_t *__restrict__ out) {
    __shared__ int64_t s_warp[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int64_t carry = bsum[blockIdx.x];
#pragma unroll 1
    for (int j = 0; j < SCAN_ITEMS; j++) {
        const int64_t i = base + j * SCAN_THREADS + threadIdx.x;
        const int64_t v = i < n ? in[i] : 0;
        int64_t total;
        const int64_t inc = block_incl_scan(v, s_warp, &total);
        if (i < n) out[i] = carry + inc - v;
        carry += total;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = bsum[nb];
}

int64_t mg_scan_tmp_elems(int64_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 2; }

int mg_scan_i32(const int32_t *d_in, int64_t *d_out, int64_t n, int64_t *d_tmp, int64_t tmp_cap, cudaStream_t st) {
    const int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    MG_REQUIRE(tmp_cap >= nb + 1, "scan scratch too small");
    if (n == 0) {
        MG_CUDA(cudaMemsetAsync(d_out, 0, sizeof(int64_t), st));
        return MG_OK;
    }
    k_scan_reduce<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(d_in, n, d_tmp);
    MG_LAUNCH_CHECK();
    k_scan_bsums<<<1, SCAN_THREADS, 0, st>>>(d_tmp, nb);
    MG_LAUNCH_CHECK();
    k_scan_apply<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(d_in, n, d_tmp, nb, d_out);
    MG_LAUNCH_CHECK();
    return MG_OK;
}
