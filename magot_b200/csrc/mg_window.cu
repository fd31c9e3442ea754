// mg_window.cu -- K6: per-base flags and sliding-window sums, the device side of `position_dic` (genome.py:981-1100).
//
//   k_at_flags     position_dic.at_content (genome.py:1030-1034): one byte per base, 1 where the base is one of "ATat".
//                  Reads the packed genome (0.5 B/base) and writes 1 B/base; the per-position Python loop of the reference
//                  disappears.  Nibble codes A=0, T=3, a=4, t=7 are exactly the codes < 8 whose two low bits are equal.
//   k_prefix_*     exclusive prefix sums E[i] = sum(values[0:i]) (int64, n + 1 entries) in ONE pass: per-thread serial
//                  scan of 16 (uint8) or 4 (int64) items, warp/block scan of the thread totals, decoupled look-back across
//                  tiles (mg_lookback.cuh).
//   k_window_sums  numpy.sum(values[s : s + window]) for s = k * jump (sliding_window_calculate, genome.py:1055) as
//                  E[s + window] - E[s]: O(n) instead of the reference's O(n * window / jump).
// All of it is HBM-bound integer work (1 B + 8 B per element for the scan, 16 B per window): no tensor cores.
#include <algorithm>
#include "mg_common.cuh"
#include "mg_lookback.cuh"

__global__ void __launch_bounds__(256) k_at_flags(const uint32_t *__restrict__ packed, int64_t g0, int64_t n, uint8_t *__restrict__ out) {
    // one thread = 8 bases = one packed word (g0 is a multiple of 8: contigs start on multiples of 32) = 8 output bytes
    const int64_t nw = (n + 7) >> 3;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nw; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t x = __ldg(packed + (g0 >> 3) + i);
        // per nibble: bit3 == 0 and bit1 == bit0
        const uint32_t hit = ~(x >> 3) & ~((x >> 1) ^ x) & 0x11111111u;
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            lo |= ((hit >> (4 * k)) & 1u) << (8 * k);
            hi |= ((hit >> (4 * k + 16)) & 1u) << (8 * k);
        }
        const int64_t b = i << 3;
        if (b + 8 <= n) {
            *reinterpret_cast<uint2 *>(out + b) = make_uint2(lo, hi);
        } else {
            const uint64_t v = ((uint64_t)hi << 32) | lo;
            for (int k = 0; b + k < n; k++) out[b + k] = (uint8_t)(v >> (8 * k));
        }
    }
}

template <typename T, int ITEMS>
__global__ void __launch_bounds__(256) k_prefix(const T *__restrict__ in, int64_t n, unsigned long long *tmp, int64_t *__restrict__ out) {
    __shared__ int64_t s_warp[8];
    __shared__ int64_t s_prefix;
    __shared__ unsigned int s_tile;
    const int64_t tile = mg_next_tile(tmp, &s_tile);
    const int64_t base = (tile * 256 + (int64_t)threadIdx.x) * ITEMS;
    int64_t v[ITEMS];
    if (base + ITEMS <= n) {
        if (sizeof(T) == 1) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(in + base));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < ITEMS; j++) v[j] = (w[j >> 2] >> ((j & 3) * 8)) & 0xFFu;
        } else {
#pragma unroll
            for (int j = 0; j < ITEMS; j++) v[j] = (int64_t)in[base + j];
        }
    } else {
#pragma unroll
        for (int j = 0; j < ITEMS; j++) v[j] = base + j < n ? (int64_t)in[base + j] : 0;
    }
    int64_t mine = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) mine += v[j];
    int64_t total;
    const int64_t incl = mg_block_incl_scan(mine, s_warp, &total);
    const int64_t prefix = mg_lookback(tmp, tile, total, &s_prefix);
    int64_t run = prefix + incl - mine;
    if (base + ITEMS <= n) {
        longlong2 *dst = reinterpret_cast<longlong2 *>(out + base);
#pragma unroll
        for (int j = 0; j < ITEMS; j += 2) {
            const int64_t a = run, b = run + v[j];
            dst[j >> 1] = make_longlong2(a, b);
            run = b + v[j + 1];
        }
    } else {
#pragma unroll
        for (int j = 0; j < ITEMS; j++) { if (base + j < n) out[base + j] = run; run += v[j]; }
    }
    if (tile == gridDim.x - 1 && threadIdx.x == 255) out[n] = prefix + total;
}

__global__ void __launch_bounds__(256) k_window_sums(const int64_t *__restrict__ E, int64_t n, int64_t window, int64_t jump,
                                                     int64_t n_windows, int64_t *__restrict__ sums) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n_windows; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = min(k * jump, n), e = min(s + window, n);      // numpy slices clamp at the end of the array
        sums[k] = __ldg(E + e) - __ldg(E + s);
    }
}

extern "C" int mg_genome_at_flags(mg_genome *g, int64_t contig, int64_t lo, int64_t hi, uint8_t *out_host, void *stream) {
    MG_REQUIRE(g != nullptr, "genome handle is NULL");
    MG_REQUIRE(g->finalized, "mg_genome_finalize has not been called");
    MG_REQUIRE(contig >= 0 && contig < g->n_contigs, "contig index out of range");
    MG_REQUIRE(lo >= 0 && lo <= hi && hi <= g->h_contig_len[contig], "range outside the contig");
    MG_REQUIRE(lo % 8 == 0, "lo must be a multiple of 8 bases");
    if (hi == lo) return MG_OK;
    MG_REQUIRE(out_host != nullptr, "out_host is NULL");
    MG_CUDA(cudaSetDevice(g->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t STAGE = 256ll << 20;
    int rc = mg_ensure_stage(g, std::min<int64_t>(STAGE, (hi - lo + 255) / 256 * 256));
    if (rc) return rc;
    for (int64_t done = lo; done < hi;) {
        const int64_t m = std::min<int64_t>(g->stage_cap / 8 * 8, hi - done);
        const int blocks = (int)std::min<int64_t>(((m + 7) / 8 + 255) / 256, 148 * 16);
        k_at_flags<<<std::max(blocks, 1), 256, 0, st>>>(g->d_packed, g->h_contig_base[contig] + done, m, g->d_stage);
        MG_LAUNCH_CHECK();
        MG_CUDA(cudaMemcpyAsync(out_host + (done - lo), g->d_stage, m, cudaMemcpyDeviceToHost, st));
        MG_CUDA(cudaStreamSynchronize(st));           // the staging buffer is reused by the next trip
        done += m;
    }
    return MG_OK;
}

extern "C" int mg_window_sums(int device, const void *values_host, int elem_size, int64_t n, int64_t window, int64_t jump,
                              int64_t n_windows, int64_t *sums_host, void *stream) {
    MG_REQUIRE(elem_size == 1 || elem_size == 8, "elem_size must be 1 (uint8 / bool) or 8 (int64)");
    MG_REQUIRE(n >= 0 && window >= 0 && jump >= 1 && n_windows >= 0, "bad window arguments");
    int ndev = 0;
    int rc = mg_device_count(&ndev);
    if (rc) return rc;
    if (device < 0 || device >= ndev) {
        mg_set_error("device %d not available (%d CUDA devices); libmagot_b200 has no CPU fallback", device, ndev);
        return MG_ECUDA;
    }
    if (n_windows == 0) return MG_OK;
    MG_REQUIRE(sums_host != nullptr && (n == 0 || values_host != nullptr), "NULL buffer");
    MG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const int items = elem_size == 1 ? 16 : 4;
    const int64_t tile = 256 * items;
    const int64_t nt = std::max<int64_t>((n + tile - 1) / tile, 1);
    auto al = [](int64_t x) { return (x + 255) / 256 * 256; };
    const int64_t b_in = al(n * elem_size + 16), b_E = al((n + 1) * 8), b_tmp = al((nt + 2) * 8), b_out = al(n_windows * 8);
    uint8_t *d = nullptr;
    MG_CUDA(cudaMallocAsync((void **)&d, b_in + b_E + b_tmp + b_out, st));
    uint8_t *d_in = d;
    int64_t *d_E = (int64_t *)(d + b_in);
    unsigned long long *d_tmp = (unsigned long long *)(d + b_in + b_E);
    int64_t *d_out = (int64_t *)(d + b_in + b_E + b_tmp);
    if (n) MG_CUDA(cudaMemcpyAsync(d_in, values_host, n * elem_size, cudaMemcpyHostToDevice, st));
    MG_CUDA(cudaMemsetAsync(d_tmp, 0, (nt + 1) * 8, st));
    if (elem_size == 1) k_prefix<uint8_t, 16><<<(unsigned)nt, 256, 0, st>>>(d_in, n, d_tmp, d_E);
    else k_prefix<int64_t, 4><<<(unsigned)nt, 256, 0, st>>>((const int64_t *)d_in, n, d_tmp, d_E);
    MG_LAUNCH_CHECK();
    const int blocks = (int)std::min<int64_t>((n_windows + 255) / 256, 148 * 16);
    k_window_sums<<<std::max(blocks, 1), 256, 0, st>>>(d_E, n, window, jump, n_windows, d_out);
    MG_LAUNCH_CHECK();
    MG_CUDA(cudaMemcpyAsync(sums_host, d_out, n_windows * 8, cudaMemcpyDeviceToHost, st));
    MG_CUDA(cudaStreamSynchronize(st));
    MG_CUDA(cudaFreeAsync(d, st));
    return MG_OK;
}
