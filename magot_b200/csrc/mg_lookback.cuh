// mg_lookback.cuh -- block-level inclusive scan and decoupled look-back across tiles (single-pass prefix sums).
// Tile status words live in st[0..n_tile) as (status << 62) | (value as a 62-bit two's complement number): status 1 = the tile's own sum, 2 = inclusive prefix up
// to and including the tile; tmp[0] is a ticket counter that hands out tiles in launch order, so a tile only ever waits
// for tiles that are already running.  tmp[] must be zeroed before the launch.
#pragma once
#include <stdint.h>

#define MG_ST_SUM (1ull << 62)
#define MG_ST_PREFIX (2ull << 62)
#define MG_ST_MASK (3ull << 62)

__device__ __forceinline__ int64_t mg_warp_incl_scan(int64_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// inclusive scan across a block of 256 threads; *total receives the block sum.  s_warp: 8 x int64
__device__ __forceinline__ int64_t mg_block_incl_scan(int64_t v, int64_t *s_warp, int64_t *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = mg_warp_incl_scan(v);
    if (lane == 31) s_warp[wid] = v;
    __syncthreads();
    if (wid == 0) {
        int64_t w = lane < 8 ? s_warp[lane] : 0;
        w = mg_warp_incl_scan(w);
        if (lane < 8) s_warp[lane] = w;
    }
    __syncthreads();
    const int64_t off = wid ? s_warp[wid - 1] : 0;
    *total = s_warp[7];
    __syncthreads();
    return v + off;
}

// next tile in launch order (whole block must call; s_tile is a shared scratch word)
__device__ __forceinline__ int64_t mg_next_tile(unsigned long long *tmp, unsigned int *s_tile) {
    if (threadIdx.x == 0) *s_tile = (unsigned int)atomicAdd(tmp, 1ull);
    __syncthreads();
    return *s_tile;
}

// Publishes this tile's sum and returns the sum of all earlier tiles (whole block must call; ends with a barrier).
__device__ __forceinline__ int64_t mg_lookback(unsigned long long *tmp, int64_t tile, int64_t total, int64_t *s_prefix) {
    volatile unsigned long long *st = tmp + 1;
    if (threadIdx.x == 0) st[tile] = (tile == 0 ? MG_ST_PREFIX : MG_ST_SUM) | ((unsigned long long)total & ~MG_ST_MASK);
    if (threadIdx.x < 32) {                                     // warp 0 looks back over 32 predecessors at a time
        int64_t prefix = 0;
        int64_t j = tile - 1 - (int64_t)threadIdx.x;
        while (true) {
            unsigned long long w = MG_ST_PREFIX;                // before tile 0: an empty inclusive prefix
            if (j >= 0) { do { w = st[j]; } while ((w & MG_ST_MASK) == 0); }
            const unsigned int done = __ballot_sync(0xffffffffu, (w & MG_ST_MASK) == MG_ST_PREFIX);
            const int first = done ? __ffs(done) - 1 : 32;      // nearest predecessor that already has its inclusive prefix
            int64_t add = (int)threadIdx.x <= first ? ((int64_t)(w << 2) >> 2) : 0;   // 62-bit two's complement: sums may be negative
#pragma unroll
            for (int d = 16; d; d >>= 1) add += __shfl_xor_sync(0xffffffffu, add, d);
            prefix += add;
            if (done) break;
            j -= 32;
        }
        if (threadIdx.x == 0) {
            if (tile > 0) st[tile] = MG_ST_PREFIX | ((unsigned long long)(prefix + total) & ~MG_ST_MASK);
            *s_prefix = prefix;
        }
    }
    __syncthreads();
    return *s_prefix;
}
