// mg_scan.cu -- K1 building block: exclusive prefix sum int32 -> int64 with warp/block scans
// (one launch: per-tile warp/block scan + decoupled look-back across tiles, mg_lookback.cuh).  This is the
// device replacement of the implicit length bookkeeping of `"".join(seq_list)` (genome.py:705).
#include "mg_common.cuh"
#include "mg_lookback.cuh"

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_lookback(const int32_t *__restrict__ in, int64_t n, unsigned long long *tmp,
                                                                int64_t *__restrict__ out) {
    __shared__ int64_t s_warp[SCAN_THREADS / 32];
    __shared__ int64_t s_prefix;
    __shared__ unsigned int s_tile;
    const int64_t tile = mg_next_tile(tmp, &s_tile);
    const int64_t base = tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int32_t v[SCAN_ITEMS];
    if (base + SCAN_ITEMS <= n) {
        const int4 a = __ldg(reinterpret_cast<const int4 *>(in + base)), b = __ldg(reinterpret_cast<const int4 *>(in + base) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; j++) v[j] = base + j < n ? in[base + j] : 0;
    }
    int64_t mine = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) mine += v[j];
    int64_t total;
    const int64_t incl = mg_block_incl_scan(mine, s_warp, &total);
    const int64_t prefix = mg_lookback(tmp, tile, total, &s_prefix);
    int64_t run = prefix + incl - mine;
    if (base + SCAN_ITEMS <= n) {
        int64_t o[SCAN_ITEMS];
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; j++) { o[j] = run; run += v[j]; }
        longlong2 *dst = reinterpret_cast<longlong2 *>(out + base);
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS / 2; j++) dst[j] = make_longlong2(o[2 * j], o[2 * j + 1]);
    } else {
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; j++) { if (base + j < n) out[base + j] = run; run += v[j]; }
    }
    if (tile == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) out[n] = prefix + total;
}

int64_t mg_scan_tmp_elems(int64_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 2; }

int mg_scan_i32(const int32_t *d_in, int64_t *d_out, int64_t n, int64_t *d_tmp, int64_t tmp_cap, cudaStream_t st) {
    const int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    MG_REQUIRE(tmp_cap >= nb + 1, "scan scratch too small");
    if (n == 0) {
        MG_CUDA(cudaMemsetAsync(d_out, 0, sizeof(int64_t), st));
        return MG_OK;
    }
    MG_CUDA(cudaMemsetAsync(d_tmp, 0, (nb + 1) * sizeof(int64_t), st));
    k_scan_lookback<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(d_in, n, reinterpret_cast<unsigned long long *>(d_tmp), d_out);
    MG_LAUNCH_CHECK();
    return MG_OK;
}
