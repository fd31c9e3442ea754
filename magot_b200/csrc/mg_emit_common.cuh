// mg_emit_common.cuh -- device helpers shared by the emit kernels (mg_emit.cu: K2/K3 with per-lane global loads;
// mg_emit_tma.cu: the bulk-copy staged K2 and the fused splice+translate kernel).
#pragma once
#include "mg_common.cuh"
#include "mg_gather.cuh"

#ifndef NUC_LD64
#define NUC_LD64 1
#endif
#ifndef NUC_LD128
#define NUC_LD128 1                 // measured on config 4: CDS launch 0.1053 -> 0.1037 ms, exon launch 0.1514 -> 0.1505 ms against three 8-byte loads
#endif
#ifndef FRAME_LANES4
#define FRAME_LANES4 1
#endif

#define BIG 0x7fffffff

// expand the low 4 bits of x into a byte mask (bit k -> byte k = 0xFF)
__device__ __forceinline__ uint32_t expand4(uint32_t x) {
    return ((x & 1u) | ((x & 2u) << 7) | ((x & 4u) << 14) | ((x & 8u) << 21)) * 0xFFu;
}

// nibble mask for positions [lo, hi) of a 16-nibble word, 0 <= lo < hi <= 16
__device__ __forceinline__ uint64_t nib_range_mask(int lo, int hi) {
    return ((~0ull) >> (64 - 4 * (hi - lo))) << (4 * lo);
}

// 16 bytes starting at byte index a of `lit` (a may be unaligned; the buffer is padded on both sides)
__device__ __forceinline__ void ld_lit16(const uint8_t *__restrict__ lit, int64_t a, uint32_t w[4]) {
    const uint32_t *p = reinterpret_cast<const uint32_t *>(lit) + (a >> 2);
    const uint32_t sh = ((uint32_t)a & 3u) << 3;
    const uint32_t x0 = __ldg(p), x1 = __ldg(p + 1), x2 = __ldg(p + 2), x3 = __ldg(p + 3), x4 = __ldg(p + 4);
    w[0] = __funnelshift_r(x0, x1, sh);
    w[1] = __funnelshift_r(x1, x2, sh);
    w[2] = __funnelshift_r(x2, x3, sh);
    w[3] = __funnelshift_r(x3, x4, sh);
}

// 32 consecutive nibbles starting at global base index g, as four words
__device__ __forceinline__ void ld_nib32(const uint32_t *__restrict__ pk, int64_t g, uint32_t n[4]) {
    const uint32_t *q = pk + (g >> 3);
    const uint32_t sh = ((uint32_t)g & 7u) << 2;
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2), w3 = __ldg(q + 3), w4 = __ldg(q + 4);
    n[0] = __funnelshift_r(w0, w1, sh);
    n[1] = __funnelshift_r(w1, w2, sh);
    n[2] = __funnelshift_r(w2, w3, sh);
    n[3] = __funnelshift_r(w3, w4, sh);
}

// one word of the packed genome.  L2::64B: a piece is ~100 packed bytes at a random address; without the hint L2 fills
// whole 128-byte lines from DRAM on a sector miss (measured: 1.6 x the sectors the SMs asked for), with it 64-byte halves.
__device__ __forceinline__ uint32_t ld_pk(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// five consecutive packed words starting at the (4-byte aligned) word p, fetched as three 8-byte loads from the enclosing
// 8-byte aligned window and selected by the odd/even word address: 3 instead of 5 load instructions and 48 instead of 80 L1
// sectors per warp (the lanes of a warp read 16-byte strided windows, so every load instruction touches all 16 sectors
// of the 512-byte span whatever its width).  Reads at most 12 bytes before p and 12 bytes past p + 20: inside the front
// padding and the tail slack of the buffer.
__device__ __forceinline__ void ld_pk5(const uint32_t *p, uint32_t r[5]) {
#if NUC_LD128
    // two 16-byte loads from the enclosing 16-byte aligned window + a two-level select on the word offset: 2 instead of 3 load
    // instructions per window (each one touches every 128-byte line the warp's windows lie in, whatever its width)
    const uint64_t a = (uint64_t)p;
    const uint64_t a16 = a & ~15ull;
    uint32_t v[8];
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "l"(a16));
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(a16 + 16));
    const bool two = (a & 8ull) != 0, one = (a & 4ull) != 0;
    uint32_t t[6];
#pragma unroll
    for (int k = 0; k < 6; k++) t[k] = two ? v[k + 2] : v[k];
#pragma unroll
    for (int k = 0; k < 5; k++) r[k] = one ? t[k + 1] : t[k];
#elif NUC_LD64
    const uint64_t a = (uint64_t)p;
    const uint64_t a8 = a & ~7ull;
    uint32_t v[6];
    asm volatile("ld.global.nc.L2::64B.v2.u32 {%0,%1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "l"(a8));
    asm volatile("ld.global.nc.L2::64B.v2.u32 {%0,%1}, [%2];" : "=r"(v[2]), "=r"(v[3]) : "l"(a8 + 8));
    asm volatile("ld.global.nc.L2::64B.v2.u32 {%0,%1}, [%2];" : "=r"(v[4]), "=r"(v[5]) : "l"(a8 + 16));
    const bool odd = (a & 4ull) != 0;
#pragma unroll
    for (int k = 0; k < 5; k++) r[k] = odd ? v[k + 1] : v[k];
#else
#pragma unroll
    for (int k = 0; k < 5; k++) r[k] = ld_pk(p + k);
#endif
}

// word mask with the low 4*t bits set, t clamped to [0, 8] nibbles: one max and one clamped funnel shift
__device__ __forceinline__ uint32_t low_nibbles(int t) {
    return __funnelshift_lc(0xFFFFFFFFu, 0u, (uint32_t)(4 * max(t, 0)));
}

__device__ __forceinline__ void st32(uint8_t *p, const uint32_t w[8]) {          // one 256-bit store (STG.E.ENL2.256)
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}

// The same 32 bytes as two 128-bit stores, for the out-of-line slow paths: ptxas 12.9 turns the v8 store into a single 32-bit
// STG when it clones such a helper for an entry that takes its arguments as a __grid_constant__ struct (seen in the SASS of
// nuc_chunk_slow / nuc_chunk_generic under k_emit_nuc: only the first word of a chunk with out-of-alphabet bytes was written).
__device__ __forceinline__ void st32_2x16(uint8_t *p, const uint32_t w[8]) {
    asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
    asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p + 16), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}

// 32 nibbles -> 32 ASCII bytes.  Chunks without N / IUPAC / '-' codes (bit 3 clear in every nibble: A C G T a c g t) need two
// table look-ups per word; the general decode needs four plus the two select masks (and, at 32 registers, re-materialises
// its four table constants: 70 instructions per chunk in the round-1 SASS against ~20 here).
__device__ __forceinline__ void mg_decode32(const uint32_t n[4], uint32_t w[8]) {
    if (((n[0] | n[1] | n[2] | n[3]) & 0x88888888u) == 0) {
        const uint32_t LA = 0x54474341u, LB = 0x74676361u;    // "ACGT", "acgt"
#pragma unroll
        for (int k = 0; k < 4; k++) {
            w[2 * k] = __byte_perm(LA, LB, n[k]);
            w[2 * k + 1] = __byte_perm(LA, LB, n[k] >> 16);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) mg_decode8(n[k], w[2 * k], w[2 * k + 1]);
    }
}

// ---- generic (slow, always correct) assembly of one 32-byte chunk straight from global memory -------------------
// Walks the piece table from piece j (piece_off[j] <= P), genome and literal pieces alike.  Used by the literal-chunk
// kernel (every chunk that contains framing bytes) and by K2 for tiles whose piece list overflows its staging.
static __device__ __noinline__ void nuc_chunk_generic(const uint32_t *__restrict__ packed, const int64_t *__restrict__ piece_off,
                                               const int64_t *__restrict__ piece_src, int64_t j, int64_t P, int64_t total,
                                               int64_t T, const uint8_t *__restrict__ lit, const int64_t *__restrict__ exc_pos,
                                               const uint8_t *__restrict__ exc_byte, int64_t n_exc, uint8_t *__restrict__ out) {
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t off_j = __ldg(piece_off + j), off_n = __ldg(piece_off + j + 1);
    const int end = (int)min((int64_t)32, total - P);
    int t = 0;
    while (t < end) {
        while (off_n <= P + t) { j++; off_j = off_n; off_n = __ldg(piece_off + j + 1); }
        const uint64_t sk = (uint64_t)__ldg(piece_src + j);
        const int64_t src = (int64_t)(sk & MG_SRC_MASK) + (P - off_j);           // source index of chunk position 0
        const int hi = (int)min((int64_t)end, off_n - P);
        const uint32_t m = (hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u)) & ~((1u << t) - 1u);   // chunk positions [t, hi)
        uint32_t b[8];
        if ((sk >> MG_KIND_SHIFT) == MG_KIND_LIT) {
            ld_lit16(lit, src, b);
            ld_lit16(lit, src + 16, b + 4);
        } else {
            uint32_t n[4];
            ld_nib32(packed, src, n);
#pragma unroll
            for (int k = 0; k < 4; k++) mg_decode8(n[k], b[2 * k], b[2 * k + 1]);
            if (n_exc > 0 && src < T) {
                for (int q = t; q < hi; q++) {
                    if (((n[q >> 3] >> ((q & 7) * 4)) & 15u) == MG_CODE_EXC) {
                        const uint32_t c = mg_exc_byte(exc_pos, exc_byte, n_exc, src + q);
                        b[q >> 2] = (b[q >> 2] & ~(0xFFu << ((q & 3) * 8))) | (c << ((q & 3) * 8));
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t mk = expand4(m >> (4 * k));
            w[k] = (w[k] & ~mk) | (b[k] & mk);
        }
        t = hi;
    }
    st32_2x16(out + P, w);
}

#define PIECE_G 0
#define PIECE_E 1        // empty (clamped-away segment, empty literal)
#define PIECE_L 2        // non-empty literal

// ---- K3 framing bytes (">ID\n" prefixes, "\n" suffixes) of one protein tile, one thread per record -----------------------
// Runs at the end of the CTA that wrote the tile's residues, over the zero bytes it left at those positions: the lines are
// still in L2, so this costs no DRAM traffic (as a separate kernel after K3 it was 0.025 ms per launch of config 4).
__device__ __forceinline__ void prot_write_framing(int64_t r_lo, int64_t r_hi, int64_t P0, int tile_len,
                                                   const int64_t *__restrict__ prot_off, const int32_t *__restrict__ rec_aa,
                                                   const int64_t *__restrict__ rec_lit_off, const int32_t *__restrict__ rec_pre,
                                                   const int32_t *__restrict__ rec_suf, const uint8_t *__restrict__ lit,
                                                   uint8_t *__restrict__ out) {
#if FRAME_LANES4
    // four lanes per record: a header of <= 32 bytes is one batch of loads (lane s takes bytes s, s + 4, ..), the suffix goes
    // with the last lane, so the CTA's epilogue is two load latencies (record fields, literal bytes) whatever the header length
    for (int64_t idx = threadIdx.x; idx < (r_hi - r_lo) * 4; idx += blockDim.x) {
        const int64_t r = r_lo + (idx >> 2);
        const int sub = (int)(idx & 3);
        const int pre = rec_pre[r], suf = rec_suf[r];
        int32_t naa = rec_aa[r];
        if (naa < 0) naa = 0;
        const int64_t a = __ldg(prot_off + r) - P0;            // tile-relative start of the record
        const uint8_t *src = lit + rec_lit_off[r];
        const int64_t e = a + pre + naa;                       // suffix position
        const int q0 = (int)max(a, (int64_t)0), q1 = (int)min(a + pre, (int64_t)tile_len);
        for (int q = q0 + sub; q < q1; q += 32) {
            uint8_t b[8];
#pragma unroll
            for (int j = 0; j < 8; j++) b[j] = q + 4 * j < q1 ? __ldg(src + (q + 4 * j - a)) : (uint8_t)0;
#pragma unroll
            for (int j = 0; j < 8; j++) if (q + 4 * j < q1) out[P0 + q + 4 * j] = b[j];
        }
        if (sub == 3)
            for (int64_t q = max(e, (int64_t)0); q < min(e + suf, (int64_t)tile_len); q++) out[P0 + q] = __ldg(src + pre + (q - e));
    }
#else
    for (int64_t r = r_lo + threadIdx.x; r < r_hi; r += blockDim.x) {
        const int pre = rec_pre[r], suf = rec_suf[r];
        int32_t naa = rec_aa[r];
        if (naa < 0) naa = 0;
        const int64_t a = __ldg(prot_off + r) - P0;            // tile-relative start of the record
        const uint8_t *src = lit + rec_lit_off[r];
        const int64_t e = a + pre + naa;                       // suffix position
        {   // prefix in batches of eight bytes: eight loads in flight, then eight stores
            const int q0 = (int)max(a, (int64_t)0), q1 = (int)min(a + pre, (int64_t)tile_len);
            for (int q = q0; q < q1; q += 8) {
                uint8_t b[8];
#pragma unroll
                for (int j = 0; j < 8; j++) b[j] = q + j < q1 ? __ldg(src + (q + j - a)) : (uint8_t)0;
#pragma unroll
                for (int j = 0; j < 8; j++) if (q + j < q1) out[P0 + q + j] = b[j];
            }
        }
        for (int64_t q = max(e, (int64_t)0); q < min(e + suf, (int64_t)tile_len); q++) out[P0 + q] = __ldg(src + pre + (q - e));
    }
#endif
}

// ---- K3 -------------------------------------------------------------------------------------------------------
// generic fallback: one residue at a time from global memory (literal positions skipped)
static __device__ __noinline__ void prot_chunk_generic(const uint32_t *__restrict__ packed, const int64_t *__restrict__ piece_off,
                                                const int64_t *__restrict__ piece_src, const int64_t *__restrict__ rec_seg_off,
                                                const int64_t *__restrict__ prot_off, const int32_t *__restrict__ rec_aa,
                                                const int8_t *__restrict__ rec_skip, const int32_t *__restrict__ rec_pre, int64_t r,
                                                int64_t P, int64_t total, const uint8_t *__restrict__ aa4096, uint8_t *__restrict__ out) {
    uint32_t w[4] = {0, 0, 0, 0};
    int64_t off_r = __ldg(prot_off + r), off_n = __ldg(prot_off + r + 1);
    for (int t = 0; t < 16 && P + t < total; t++) {
        const int64_t pos = P + t;
        while (off_n <= pos) { r++; off_r = off_n; off_n = __ldg(prot_off + r + 1); }
        const int64_t q = pos - off_r, pre = rec_pre[r];
        int64_t naa = rec_aa[r];
        if (naa < 0) naa = 0;
        if (q < pre || q >= pre + naa) continue;
        const int64_t f0 = __ldg(rec_seg_off + r) + 2 * r, f1 = __ldg(rec_seg_off + r + 1) + 2 * (r + 1);
        const int64_t S = __ldg(piece_off + f0 + 1) + rec_skip[r] + 3 * (q - pre);
        const int64_t j = mg_search_le(piece_off, f0 + 1, f1 - 1, S);
        uint64_t acc[3];
        mg_gather_nib(packed, piece_off, piece_src, j, S, 3, acc);
        w[t >> 2] |= (uint32_t)__ldg(aa4096 + ((uint32_t)acc[0] & 0xFFFu)) << ((t & 3) * 8);
    }
    mg_st16(out + P, w[0], w[1], w[2], w[3]);
}
