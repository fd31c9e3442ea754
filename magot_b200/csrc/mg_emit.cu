// mg_emit.cu -- K2 (segmented interval gather + per-segment reverse complement + FASTA framing) and
// K3 (codon translation of the spliced sequence).  Both are flat passes over OUTPUT bytes: one thread
// owns one aligned 16-byte chunk of the final text and issues exactly one 16-byte streaming store, so
// stores are perfectly coalesced and the load balance is independent of transcript/exon lengths.
//
//   K2 replaces ParentAnnotation.get_fasta seq_type="nucleotide" (genome.py:687-710),
//      BaseAnnotation.get_seq (genome.py:603-608) and Sequence.reverse_compliment (genome.py:784-793).
//   K3 replaces Sequence.translate(frame=0, strand='+', trimX=True) (genome.py:795-822) as called on
//      the spliced sequence at genome.py:707.
//
// Algorithmic HBM bytes: K2 = 0.5 B/base packed read + 1 B/base text written (+ tables);
// K3 = 0.5 B/base read + 1/3 B/base written.  No dense contraction exists -> no tensor cores.
//
// Round-1 ncu finding (profiles/r1a_*): the first version of these kernels was issue-bound (~520 / ~1900
// warp instructions per 32 chunks, 18 of 32 lanes active), not HBM-bound.  This version therefore
//   * stages the tile's piece table in shared memory in TILE-RELATIVE 32-bit form, with one 64-bit
//     "base" per piece chosen so that  source index = base +/- (position in tile):  a chunk loads the 16
//     nibbles that are ALREADY ALIGNED with its 16 output positions and only masks them -- no per-piece
//     shifting, no 64-bit offset arithmetic in the inner loop;
//   * replaces the per-chunk binary search by a 64-byte-unit -> piece lookup table built per tile;
//   * copies literal bytes (FASTA headers) with 5 aligned word loads + funnel shifts instead of byte loops;
//   * patches bytes outside the packed alphabet (code 15) in a rare tail path.
#include <algorithm>
#include "mg_common.cuh"
#include "mg_gather.cuh"

#define NUC_THREADS 256
#define NUC_CHUNKS (MG_NUC_TILE / 16 / NUC_THREADS)     // 4 chunks of 16 B per thread
#define NUC_CAP 1024                                     // pieces cached in shared memory per tile
#define NUC_UNITS (MG_NUC_TILE / 64)

#define PROT_THREADS 256
#define PROT_CHUNKS (MG_PROT_TILE / 16 / PROT_THREADS)  // 2
#define PROT_RCAP 256                                    // records cached per tile
#define PROT_PCAP 1536                                   // pieces cached per tile
#define PROT_UNITS (MG_PROT_TILE / 64)

#define KIND_FWD 0
#define KIND_RC 1
#define KIND_LIT 2

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

// ---- generic (slow, always correct) chunk assembly straight from global memory ------------------------------
// Used for tiles whose piece list does not fit the shared-memory cache (thousands of tiny pieces per 16 KB).
__device__ __noinline__ void nuc_chunk_generic(const uint32_t *__restrict__ packed, const int64_t *__restrict__ piece_off,
                                               const int64_t *__restrict__ piece_src, int64_t n_piece, int64_t j, int64_t P,
                                               int64_t total, const uint8_t *__restrict__ lit, const int64_t *__restrict__ exc_pos,
                                               const uint8_t *__restrict__ exc_byte, int64_t n_exc, uint8_t *__restrict__ out) {
    uint32_t w[4] = {0, 0, 0, 0};
    int64_t off_j = __ldg(piece_off + j), off_n = __ldg(piece_off + j + 1);
    for (int t = 0; t < 16 && P + t < total; t++) {
        const int64_t pos = P + t;
        while (off_n <= pos) { j++; off_j = off_n; off_n = __ldg(piece_off + j + 1); }
        const uint64_t sk = (uint64_t)__ldg(piece_src + j);
        const uint64_t kind = sk >> MG_KIND_SHIFT;
        const int64_t src = (int64_t)(sk & MG_SRC_MASK), o = pos - off_j;
        uint32_t b;
        if (kind == MG_KIND_LIT) {
            b = __ldg(lit + src + o);
        } else {
            const int64_t gi = kind == MG_KIND_FWD ? src + o : src + (off_n - off_j) - 1 - o;
            uint32_t code = (__ldg(packed + (gi >> 3)) >> (((uint32_t)gi & 7u) * 4)) & 15u;
            if (kind == MG_KIND_RC) code = code < 8 ? (code ^ 3u) : (code > 10 ? 9u : code);
            uint32_t d0, d1;
            mg_decode8(code, d0, d1);
            b = d0 & 0xFFu;
            if (code == MG_CODE_EXC) b = mg_exc_byte(exc_pos, exc_byte, n_exc, gi);
        }
        w[t >> 2] |= b << ((t & 3) * 8);
    }
    mg_st16(out + P, w[0], w[1], w[2], w[3]);
}

// ---- K2 ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NUC_THREADS) k_emit_nuc(
    const uint32_t *__restrict__ packed, const int64_t *__restrict__ piece_off, const int64_t *__restrict__ piece_src,
    int64_t n_piece, const int64_t *__restrict__ tile_first, int64_t total, const uint8_t *__restrict__ lit,
    const int64_t *__restrict__ exc_pos, const uint8_t *__restrict__ exc_byte, int64_t n_exc, uint8_t *__restrict__ out) {
    __shared__ int64_t s_base[NUC_CAP];               // source index of tile position 0 (see header comment)
    __shared__ int32_t s_rel[NUC_CAP + 1];            // piece start relative to the tile, clamped to [.., TILE]
    __shared__ uint8_t s_kind[NUC_CAP];
    __shared__ uint16_t s_unit[NUC_UNITS];            // piece holding byte 64*u of the tile
    const int64_t P0 = (int64_t)blockIdx.x * MG_NUC_TILE;
    const int64_t p_lo = tile_first[blockIdx.x];
    int64_t p_hi = tile_first[blockIdx.x + 1] + 1;    // one past the last piece this tile can touch
    if (p_hi > n_piece) p_hi = n_piece;
    const int ncache = (int)min((int64_t)NUC_CAP, p_hi - p_lo);
    for (int i = threadIdx.x; i <= ncache; i += NUC_THREADS) {
        const int64_t off = __ldg(piece_off + p_lo + i);
        const int64_t rel = off - P0;                 // > -2^31: piece lengths are int32
        s_rel[i] = rel > MG_NUC_TILE ? MG_NUC_TILE : (int32_t)rel;
        if (i < ncache) {
            const uint64_t sk = (uint64_t)__ldg(piece_src + p_lo + i);
            const int kind = (int)(sk >> MG_KIND_SHIFT);
            const int64_t src = (int64_t)(sk & MG_SRC_MASK);
            int64_t base;
            if (kind == KIND_RC) base = src + (__ldg(piece_off + p_lo + i + 1) - off) - 1 + rel;   // index = base - q
            else base = src - rel;                                                               // index = base + q
            s_base[i] = base;
            s_kind[i] = (uint8_t)kind;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ncache; i += NUC_THREADS) {
        const int r0 = s_rel[i] < 0 ? 0 : s_rel[i], r1 = s_rel[i + 1] < 0 ? 0 : s_rel[i + 1];
        const int u1 = min((r1 + 63) >> 6, NUC_UNITS);
        for (int u = (r0 + 63) >> 6; u < u1; u++) s_unit[u] = (uint16_t)i;
    }
    __syncthreads();
    const int cached_end = s_rel[ncache];             // tile-relative position where the cached pieces end
    const int tile_len = (int)min((int64_t)MG_NUC_TILE, total - P0);

#pragma unroll 1
    for (int cidx = 0; cidx < NUC_CHUNKS; cidx++) {
        const int p = (cidx * NUC_THREADS + (int)threadIdx.x) << 4;
        if (p >= tile_len) break;
        if (p + 16 > cached_end && cached_end < tile_len) {      // piece list overflowed the cache: slow path
            const int64_t j = mg_search_le(piece_off, p_lo, n_piece, P0 + p);
            nuc_chunk_generic(packed, piece_off, piece_src, n_piece, j, P0 + p, total, lit, exc_pos, exc_byte, n_exc, out);
            continue;
        }
        int j = s_unit[p >> 6];
        while (s_rel[j + 1] <= p) j++;
        uint64_t nacc = 0;                             // nibble codes of the 16 output positions
        uint32_t bw0 = 0, bw1 = 0, bw2 = 0, bw3 = 0;  // raw bytes (literals)
        uint32_t bm = 0;                               // which of the 16 bytes are raw
        const int end = min(16, tile_len - p);
        int lo = 0;
        for (;;) {
            const int hi = min(s_rel[j + 1] - p, end);
            const int kind = s_kind[j];
            const int64_t base = s_base[j];
            if (kind == KIND_LIT) {
                uint32_t w[4];
                ld_lit16(lit, base + p, w);
                const uint32_t m = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
                const uint32_t m0 = expand4(m), m1 = expand4(m >> 4), m2 = expand4(m >> 8), m3 = expand4(m >> 12);
                bw0 |= w[0] & m0; bw1 |= w[1] & m1; bw2 |= w[2] & m2; bw3 |= w[3] & m3;
                bm |= m;
            } else {
                uint64_t v = kind == KIND_FWD ? mg_ld_nib16(packed, base + p) : mg_rc_nib16(mg_ld_nib16(packed, base - p - 15));
                if (hi - lo < 16) v &= nib_range_mask(lo, hi);
                nacc |= v;
            }
            if (hi >= end) break;
            lo = hi;
            do { j++; } while (s_rel[j + 1] <= p + lo);
        }
        uint32_t w0, w1, w2, w3;
        mg_decode8((uint32_t)nacc, w0, w1);
        mg_decode8((uint32_t)(nacc >> 32), w2, w3);
        if (bm) {
            const uint32_t m0 = expand4(bm), m1 = expand4(bm >> 4), m2 = expand4(bm >> 8), m3 = expand4(bm >> 12);
            w0 = (w0 & ~m0) | bw0; w1 = (w1 & ~m1) | bw1; w2 = (w2 & ~m2) | bw2; w3 = (w3 & ~m3) | bw3;
        }
        // code 15 = byte outside the packed alphabet on a '+' piece (reverse pieces already turned it into 'n',
        // genome.py:791-792): fetch the exact byte the FASTA had (genome.py:606 keeps it).  Rare.
        uint64_t e = nacc & (nacc >> 1) & (nacc >> 2) & (nacc >> 3) & 0x1111111111111111ull;
        if (e) {
            uint32_t w[4] = {w0, w1, w2, w3};
            while (e) {
                const int t = (__ffsll((long long)e) - 1) >> 2;
                e &= e - 1;
                int jj = s_unit[p >> 6];
                while (s_rel[jj + 1] <= p + t) jj++;
                const uint32_t b = mg_exc_byte(exc_pos, exc_byte, n_exc, s_base[jj] + p + t);
                w[t >> 2] = (w[t >> 2] & ~(0xFFu << ((t & 3) * 8))) | (b << ((t & 3) * 8));
            }
            w0 = w[0]; w1 = w[1]; w2 = w[2]; w3 = w[3];
        }
        mg_st16(out + P0 + p, w0, w1, w2, w3);
    }
}

// ---- K3 -------------------------------------------------------------------------------------------------------
// generic fallback: one byte at a time from global memory
__device__ __noinline__ void prot_chunk_generic(const uint32_t *__restrict__ packed, const int64_t *__restrict__ piece_off,
                                                const int64_t *__restrict__ piece_src, const int64_t *__restrict__ rec_seg_off,
                                                const int64_t *__restrict__ prot_off, const int32_t *__restrict__ rec_aa,
                                                const int8_t *__restrict__ rec_skip, const int64_t *__restrict__ rec_lit_off,
                                                const int32_t *__restrict__ rec_pre, int64_t r, int64_t P, int64_t total,
                                                const uint8_t *__restrict__ lit, const uint8_t *__restrict__ aa4096,
                                                uint8_t *__restrict__ out) {
    uint32_t w[4] = {0, 0, 0, 0};
    int64_t off_r = __ldg(prot_off + r), off_n = __ldg(prot_off + r + 1);
    for (int t = 0; t < 16 && P + t < total; t++) {
        const int64_t pos = P + t;
        while (off_n <= pos) { r++; off_r = off_n; off_n = __ldg(prot_off + r + 1); }
        const int64_t q = pos - off_r, pre = rec_pre[r];
        int64_t naa = rec_aa[r];
        if (naa < 0) naa = 0;
        uint32_t b;
        if (q < pre) b = __ldg(lit + rec_lit_off[r] + q);
        else if (q < pre + naa) {
            const int64_t f0 = __ldg(rec_seg_off + r) + 2 * r, f1 = __ldg(rec_seg_off + r + 1) + 2 * (r + 1);
            const int64_t S = __ldg(piece_off + f0 + 1) + rec_skip[r] + 3 * (q - pre);
            const int64_t j = mg_search_le(piece_off, f0 + 1, f1 - 1, S);
            uint64_t acc[3];
            mg_gather_nib(packed, piece_off, piece_src, j, S, 3, acc);
            b = __ldg(aa4096 + ((uint32_t)acc[0] & 0xFFFu));
        } else b = __ldg(lit + rec_lit_off[r] + pre + (q - pre - naa));
        w[t >> 2] |= b << ((t & 3) * 8);
    }
    mg_st16(out + P, w[0], w[1], w[2], w[3]);
}

// Output chunk = 16 bytes of protein text.  Amino acid a of record r is the codon at spliced offset
// skip[r] + 3a; the 4096-entry nibble-triplet table (case-insensitive, anything non-ACGT -> 'X') sits in
// shared memory.  Stop codons are emitted as '*' and translation continues (genome.py:811-818).
__global__ void __launch_bounds__(PROT_THREADS) k_emit_prot(
    const uint32_t *__restrict__ packed, const int64_t *__restrict__ piece_off, const int64_t *__restrict__ piece_src,
    const int64_t *__restrict__ rec_seg_off, const int64_t *__restrict__ prot_off, const int32_t *__restrict__ rec_aa,
    const int8_t *__restrict__ rec_skip, const int64_t *__restrict__ rec_lit_off, const int32_t *__restrict__ rec_pre,
    int64_t n_rec, const int64_t *__restrict__ tile_first, int64_t total, const uint8_t *__restrict__ lit,
    const uint8_t *__restrict__ aa4096, uint8_t *__restrict__ out) {
    __shared__ __align__(16) uint8_t s_aa[4096];
    // records of the tile
    __shared__ int32_t s_pstart[PROT_RCAP + 1];       // record start in the protein text, relative to the tile (may be < 0)
    __shared__ int32_t s_pre[PROT_RCAP], s_naa[PROT_RCAP];
    __shared__ int32_t s_q0[PROT_RCAP];               // nucleotide-text position of the first codon, relative to origin O
    __shared__ int16_t s_j0[PROT_RCAP], s_j1[PROT_RCAP];   // first / one-past-last segment piece (tile-local)
    __shared__ int64_t s_lbase[PROT_RCAP];            // literal index of tile position 0 for the prefix
    // pieces of those records
    __shared__ int64_t s_base[PROT_PCAP];
    __shared__ int32_t s_rel[PROT_PCAP + 1];          // relative to O = piece_off[first cached piece]
    __shared__ uint8_t s_kind[PROT_PCAP];
    __shared__ uint16_t s_unit[PROT_UNITS];
    reinterpret_cast<uint4 *>(s_aa)[threadIdx.x] = __ldg(reinterpret_cast<const uint4 *>(aa4096) + threadIdx.x);
    const int64_t P0 = (int64_t)blockIdx.x * MG_PROT_TILE;
    const int64_t r_lo = tile_first[blockIdx.x];
    int64_t r_hi = tile_first[blockIdx.x + 1] + 1;
    if (r_hi > n_rec) r_hi = n_rec;
    const int nrec = (int)(r_hi - r_lo);
    const int64_t pc_lo = __ldg(rec_seg_off + r_lo) + 2 * r_lo;
    const int64_t pc_hi = __ldg(rec_seg_off + r_hi) + 2 * r_hi;
    const int64_t O = __ldg(piece_off + pc_lo);
    const int npc = (int)min((int64_t)PROT_PCAP + 1, pc_hi - pc_lo);
    const bool fits = nrec <= PROT_RCAP && npc <= PROT_PCAP && (__ldg(piece_off + pc_hi) - O) < 0x7fffffffll;
    const int tile_len = (int)min((int64_t)MG_PROT_TILE, total - P0);
    if (!fits) {                                      // rare: whole tile through the generic path
        for (int cidx = 0; cidx < PROT_CHUNKS; cidx++) {
            const int p = (cidx * PROT_THREADS + (int)threadIdx.x) << 4;
            if (p >= tile_len) break;
            const int64_t r = mg_search_le(prot_off, r_lo, n_rec, P0 + p);
            prot_chunk_generic(packed, piece_off, piece_src, rec_seg_off, prot_off, rec_aa, rec_skip, rec_lit_off, rec_pre, r,
                               P0 + p, total, lit, aa4096, out);
        }
        return;
    }
    for (int i = threadIdx.x; i <= npc; i += PROT_THREADS) {
        const int64_t off = __ldg(piece_off + pc_lo + i);
        const int32_t rel = (int32_t)(off - O);
        s_rel[i] = rel;
        if (i < npc) {
            const uint64_t sk = (uint64_t)__ldg(piece_src + pc_lo + i);
            const int kind = (int)(sk >> MG_KIND_SHIFT);
            const int64_t src = (int64_t)(sk & MG_SRC_MASK);
            s_base[i] = kind == KIND_RC ? src + (__ldg(piece_off + pc_lo + i + 1) - off) - 1 + rel : src - rel;
            s_kind[i] = (uint8_t)kind;
        }
    }
    for (int i = threadIdx.x; i <= nrec; i += PROT_THREADS) {
        const int64_t r = r_lo + i;
        const int64_t ps = __ldg(prot_off + r) - P0;
        s_pstart[i] = ps > MG_PROT_TILE ? MG_PROT_TILE : (int32_t)ps;
        if (i < nrec) {
            const int64_t f0 = __ldg(rec_seg_off + r) + 2 * r, f1 = __ldg(rec_seg_off + r + 1) + 2 * (r + 1);
            int32_t naa = rec_aa[r];
            s_pre[i] = rec_pre[r];
            s_naa[i] = naa < 0 ? 0 : naa;
            s_q0[i] = (int32_t)(__ldg(piece_off + f0 + 1) - O) + rec_skip[r];
            s_j0[i] = (int16_t)(f0 + 1 - pc_lo);
            s_j1[i] = (int16_t)(f1 - 1 - pc_lo);
            s_lbase[i] = rec_lit_off[r] - ps;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nrec; i += PROT_THREADS) {
        const int r0 = s_pstart[i] < 0 ? 0 : s_pstart[i], r1 = s_pstart[i + 1] < 0 ? 0 : s_pstart[i + 1];
        const int u1 = min((r1 + 63) >> 6, PROT_UNITS);
        for (int u = (r0 + 63) >> 6; u < u1; u++) s_unit[u] = (uint16_t)i;
    }
    __syncthreads();

#pragma unroll 1
    for (int cidx = 0; cidx < PROT_CHUNKS; cidx++) {
        const int p = (cidx * PROT_THREADS + (int)threadIdx.x) << 4;
        if (p >= tile_len) break;
        int r = s_unit[p >> 6];
        while (s_pstart[r + 1] <= p) r++;
        uint32_t bw0 = 0, bw1 = 0, bw2 = 0, bw3 = 0;
        const int end = min(16, tile_len - p);
        int lo = 0;
        for (;;) {                                    // pieces of this chunk: prefix | residues | suffix of record r, ...
            const int qq = p + lo - s_pstart[r];      // position inside the record's text
            const int pre = s_pre[r], naa = s_naa[r];
            int hi;
            uint32_t w[4];
            if (qq < pre) {                           // literal prefix
                hi = min(end, lo + (pre - qq));
                ld_lit16(lit, s_lbase[r] + p, w);
            } else if (qq < pre + naa) {              // residues: position t of the chunk is amino acid (p+t-pstart-pre)
                hi = min(end, lo + (pre + naa - qq));
                // nucleotide-text position (relative to O) of the codon that lands on chunk position 0
                const int q = s_q0[r] + 3 * (p - s_pstart[r] - pre);
                uint64_t acc[3] = {0, 0, 0};
                int j = s_j0[r];
                const int j1 = s_j1[r];
                {   // locate the segment holding the first needed nibble
                    const int first = q + 3 * lo;
                    int a = j, b = j1;
                    while (b - a > 1) {
                        const int mid = (a + b) >> 1;
                        if (s_rel[mid] <= first) a = mid; else b = mid;
                    }
                    j = a;
                }
                const int need_lo = 3 * lo, need_hi = 3 * hi;           // nibbles [need_lo, need_hi) of the 48
#pragma unroll
                for (int g = 0; g < 3; g++) {
                    int nlo = max(need_lo - 16 * g, 0);
                    const int nend = min(need_hi - 16 * g, 16);
                    if (nend <= nlo) continue;
                    const int qg = q + 16 * g;        // position of nibble 0 of this group
                    while (s_rel[j + 1] <= qg + nlo) j++;
                    uint64_t a = 0;
                    for (;;) {
                        const int nhi = min(s_rel[j + 1] - qg, nend);
                        const int64_t base = s_base[j];
                        uint64_t v = s_kind[j] == KIND_FWD ? mg_ld_nib16(packed, base + qg) : mg_rc_nib16(mg_ld_nib16(packed, base - qg - 15));
                        if (nhi - nlo < 16) v &= nib_range_mask(nlo, nhi);
                        a |= v;
                        if (nhi >= nend) break;
                        nlo = nhi;
                        do { j++; } while (s_rel[j + 1] <= qg + nlo);
                    }
                    acc[g] = a;
                }
                w[0] = w[1] = w[2] = w[3] = 0;
#pragma unroll
                for (int k = 0; k < 16; k++) {        // codon k sits at bit 12k of acc[2]:acc[1]:acc[0]
                    const int bit = 12 * k, ww = bit >> 6, sh = bit & 63;
                    uint32_t idx = (uint32_t)(acc[ww] >> sh);
                    if (sh > 52) idx |= (uint32_t)(acc[ww + 1] << (64 - sh));
                    w[k >> 2] |= (uint32_t)s_aa[idx & 0xFFFu] << ((k & 3) * 8);
                }
            } else {                                  // literal suffix
                hi = min(end, s_pstart[r + 1] - p);
                ld_lit16(lit, s_lbase[r] - naa + p, w);
            }
            const uint32_t m = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
            if (m == 0xFFFFu) { bw0 = w[0]; bw1 = w[1]; bw2 = w[2]; bw3 = w[3]; }
            else {
                bw0 |= w[0] & expand4(m); bw1 |= w[1] & expand4(m >> 4); bw2 |= w[2] & expand4(m >> 8); bw3 |= w[3] & expand4(m >> 12);
            }
            if (hi >= end) break;
            lo = hi;
            while (s_pstart[r + 1] <= p + lo) r++;
        }
        mg_st16(out + P0 + p, bw0, bw1, bw2, bw3);
    }
}

// ---- host API -----------------------------------------------------------------------------------------------

static int ensure_out(mg_plan *p, int64_t bytes, cudaStream_t st) {
    if (p->out_cap >= bytes) return MG_OK;
    if (p->d_out) MG_CUDA(cudaFreeAsync(p->d_out, st));
    p->d_out = nullptr;
    p->out_cap = 0;
    MG_CUDA(cudaMallocAsync((void **)&p->d_out, bytes, st));
    p->out_cap = bytes;
    return MG_OK;
}

extern "C" int mg_emit_nuc_device(mg_plan *p, uint8_t *out_dev, void *stream) {
    MG_REQUIRE(p != nullptr, "plan handle is NULL");
    if (!p->prepared) { mg_set_error("mg_plan_prepare has not been called"); return MG_ESTATE; }
    if (p->nuc_total == 0) return MG_OK;
    MG_REQUIRE(out_dev != nullptr && ((uintptr_t)out_dev & 15) == 0, "out_dev must be a 16-byte aligned device pointer");
    MG_CUDA(cudaSetDevice(p->device));
    mg_genome *g = p->g;
    p->last_stream = (cudaStream_t)stream;
    k_emit_nuc<<<(unsigned)p->n_nuc_tile, NUC_THREADS, 0, (cudaStream_t)stream>>>(
        g->d_packed, p->d_piece_off, p->d_piece_src, p->n_piece, p->d_nuc_tile, p->nuc_total, p->d_lit, g->d_exc_pos,
        g->d_exc_byte, g->n_exc, out_dev);
    MG_LAUNCH_CHECK();
    return MG_OK;
}

extern "C" int mg_emit_prot_device(mg_plan *p, uint8_t *out_dev, void *stream) {
    MG_REQUIRE(p != nullptr, "plan handle is NULL");
    if (!p->prepared) { mg_set_error("mg_plan_prepare has not been called"); return MG_ESTATE; }
    if (p->prot_total == 0) return MG_OK;
    MG_REQUIRE(out_dev != nullptr && ((uintptr_t)out_dev & 15) == 0, "out_dev must be a 16-byte aligned device pointer");
    MG_CUDA(cudaSetDevice(p->device));
    mg_genome *g = p->g;
    p->last_stream = (cudaStream_t)stream;
    k_emit_prot<<<(unsigned)p->n_prot_tile, PROT_THREADS, 0, (cudaStream_t)stream>>>(
        g->d_packed, p->d_piece_off, p->d_piece_src, p->d_rec_seg_off, p->d_prot_off, p->d_rec_aa, p->d_rec_skip,
        p->d_rec_lit_off, p->d_rec_pre, p->n_rec, p->d_prot_tile, p->prot_total, p->d_lit, g->d_aa4096, out_dev);
    MG_LAUNCH_CHECK();
    return MG_OK;
}

extern "C" int mg_emit_nuc_host(mg_plan *p, uint8_t *out_host, void *stream) {
    MG_REQUIRE(p != nullptr, "plan handle is NULL");
    if (!p->prepared) { mg_set_error("mg_plan_prepare has not been called"); return MG_ESTATE; }
    if (p->nuc_total == 0) return MG_OK;
    MG_REQUIRE(out_host != nullptr, "out_host is NULL");
    MG_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_out(p, (p->nuc_total + 15) / 16 * 16, st);
    if (rc) return rc;
    rc = mg_emit_nuc_device(p, p->d_out, stream);
    if (rc) return rc;
    MG_CUDA(cudaMemcpyAsync(out_host, p->d_out, p->nuc_total, cudaMemcpyDeviceToHost, st));
    return MG_OK;
}

extern "C" int mg_emit_prot_host(mg_plan *p, uint8_t *out_host, void *stream) {
    MG_REQUIRE(p != nullptr, "plan handle is NULL");
    if (!p->prepared) { mg_set_error("mg_plan_prepare has not been called"); return MG_ESTATE; }
    if (p->prot_total == 0) return MG_OK;
    MG_REQUIRE(out_host != nullptr, "out_host is NULL");
    MG_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_out(p, (p->prot_total + 15) / 16 * 16, st);
    if (rc) return rc;
    rc = mg_emit_prot_device(p, p->d_out, stream);
    if (rc) return rc;
    MG_CUDA(cudaMemcpyAsync(out_host, p->d_out, p->prot_total, cudaMemcpyDeviceToHost, st));
    return MG_OK;
}
