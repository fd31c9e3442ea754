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
// How the kernels got here (ncu evidence under profiles/):
//   r1a  first version: 16 bytes/thread, per-chunk binary search, per-piece shifting, reverse complement in
//        registers.  Issue-bound: ~520 (K2) / ~1900 (K3) warp instructions per 32 chunks, 17 of 32 lanes
//        active, DRAM at 20 % -- nowhere near the HBM roofline the algorithm allows.
//   r1b  tile-relative 32-bit tables in shared memory, position-aligned loads, 64-byte-unit lookup table:
//        ~370 instructions per 32 chunks; what was left was divergence (second piece of a chunk executed by
//        2 lanes, literal bytes by 1 lane) and the strand branch.
//   r1c  (this file) * the genome keeps a second, reverse-complemented plane, so a '-' interval is a plain
//        forward read: no strand branch, no register reversal (mg_common.cuh);
//        * the first two GENOME pieces of every chunk are handled branch-free by all lanes (literal and
//          clamped-away empty pieces are skipped by flag: at most two sit between two genome pieces in the
//          common case); anything else in 16 bytes is rare and takes a loop;
//        * literal bytes (">ID\n", "\n") are not touched here at all: a tiny second kernel writes them
//          afterwards, one thread per record.
#include <algorithm>
#include "mg_common.cuh"
#include "mg_gather.cuh"

#define NUC_THREADS 256
#define NUC_CHUNKS (MG_NUC_TILE / 16 / NUC_THREADS)     // 4 chunks of 16 B per thread
#define NUC_CAP 1024                                     // pieces staged per tile
#define NUC_UNITS (MG_NUC_TILE / 64)

#define PROT_THREADS 256
#define PROT_CHUNKS (MG_PROT_TILE / 16 / PROT_THREADS)  // 4
#define PROT_RCAP 512                                    // records staged per tile
#define PROT_PCAP 2048                                   // pieces staged per tile
#define PROT_UNITS (MG_PROT_TILE / 64)

#define BIG 0x7fffffff
// 1: K2 patches literal bytes itself (1 lane, ~1.5 chunks per record); 0: k_emit_lit writes them afterwards.
// Measured on B200 (config 4): inline 0.300 ms vs separate 0.204 + 0.027 ms per exon launch -> separate.
#define MG_NUC_INLINE_LIT 0

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
// Used for tiles whose piece list does not fit the shared-memory staging (thousands of tiny pieces per tile).
__device__ __noinline__ void nuc_chunk_generic(const uint32_t *__restrict__ packed, const int64_t *__restrict__ piece_off,
                                               const int64_t *__restrict__ piece_src, int64_t j, int64_t P, int64_t total,
                                               int64_t T, const uint8_t *__restrict__ lit, const int64_t *__restrict__ exc_pos, const uint8_t *__restrict__ exc_byte,
                                               int64_t n_exc, uint8_t *__restrict__ out) {
    uint32_t w[4] = {0, 0, 0, 0};
    int64_t off_j = __ldg(piece_off + j), off_n = __ldg(piece_off + j + 1);
    for (int t = 0; t < 16 && P + t < total; t++) {
        const int64_t pos = P + t;
        while (off_n <= pos) { j++; off_j = off_n; off_n = __ldg(piece_off + j + 1); }
        const uint64_t sk = (uint64_t)__ldg(piece_src + j);
        if ((sk >> MG_KIND_SHIFT) == MG_KIND_LIT) {
            w[t >> 2] |= (uint32_t)__ldg(lit + (int64_t)(sk & MG_SRC_MASK) + (pos - off_j)) << ((t & 3) * 8);
            continue;
        }
        const int64_t gi = (int64_t)(sk & MG_SRC_MASK) + (pos - off_j);
        const uint32_t code = (__ldg(packed + (gi >> 3)) >> (((uint32_t)gi & 7u) * 4)) & 15u;
        uint32_t d0, d1;
        mg_decode8(code, d0, d1);
        uint32_t b = d0 & 0xFFu;
        if (code == MG_CODE_EXC && gi < T) b = mg_exc_byte(exc_pos, exc_byte, n_exc, gi);
        w[t >> 2] |= b << ((t & 3) * 8);
    }
    mg_st16(out + P, w[0], w[1], w[2], w[3]);
}

// ---- K2 ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NUC_THREADS) k_emit_nuc(
    const uint32_t *__restrict__ packed, const int64_t *__restrict__ piece_off, const int64_t *__restrict__ piece_src,
    int64_t n_piece, const int64_t *__restrict__ tile_first, int64_t total, int64_t T, const uint8_t *__restrict__ lit,
    const int64_t *__restrict__ exc_pos, const uint8_t *__restrict__ exc_byte, int64_t n_exc, uint8_t *__restrict__ out) {
    // pieces of the tile: piece i covers text [s_rel[i], s_rel[i+1]) relative to the tile; the nibble index of
    // tile position q is s_base[i] + q (byte index into lit[] for a literal piece); s_skip[i] = 1 for clamped-away
    // empty pieces, 3 for literal pieces (bit 0 = "not a genome piece", bit 1 = literal)
    __shared__ int64_t s_base[NUC_CAP + 4];
    __shared__ int32_t s_rel[NUC_CAP + 5];
    __shared__ uint8_t s_skip[NUC_CAP + 4];
    __shared__ uint16_t s_unit[NUC_UNITS];            // piece holding byte 64*u of the tile
    const int64_t P0 = (int64_t)blockIdx.x * MG_NUC_TILE;
    const int64_t p_lo = tile_first[blockIdx.x];
    int64_t p_hi = tile_first[blockIdx.x + 1] + 1;    // one past the last piece this tile can touch
    if (p_hi > n_piece) p_hi = n_piece;
    const int ncache = (int)min((int64_t)NUC_CAP, p_hi - p_lo);
    const int tile_len = (int)min((int64_t)MG_NUC_TILE, total - P0);
    for (int i = threadIdx.x; i < ncache + 4; i += NUC_THREADS) {
        if (i <= ncache) {
            const int64_t rel = __ldg(piece_off + p_lo + i) - P0;      // > -2^31: piece lengths are int32
            s_rel[i] = rel > MG_NUC_TILE ? MG_NUC_TILE : (int32_t)rel;
        } else {
            s_rel[i] = BIG;
        }
        if (i < ncache) {
            const int64_t rel = __ldg(piece_off + p_lo + i) - P0;
            const uint64_t sk = (uint64_t)__ldg(piece_src + p_lo + i);
            s_base[i] = (int64_t)(sk & MG_SRC_MASK) - rel;
            s_skip[i] = (sk >> MG_KIND_SHIFT) == MG_KIND_LIT ? 3 : (__ldg(piece_off + p_lo + i + 1) - P0 == rel ? 1 : 0);
        } else {
            s_base[i] = 0;
            s_skip[i] = 1;                            // nothing beyond the staged pieces may be selected as A or B
        }
    }
    if (threadIdx.x == 0) s_rel[ncache + 4] = BIG;
    __syncthreads();
    for (int i = threadIdx.x; i < ncache; i += NUC_THREADS) {
        const int r0 = s_rel[i] < 0 ? 0 : s_rel[i], r1 = s_rel[i + 1] < 0 ? 0 : s_rel[i + 1];
        const int u1 = min((r1 + 63) >> 6, NUC_UNITS);
        for (int u = (r0 + 63) >> 6; u < u1; u++) s_unit[u] = (uint16_t)i;
    }
    __syncthreads();
    const int covered = s_rel[ncache];                // tile-relative position where the staged pieces end

#pragma unroll 1
    for (int cidx = 0; cidx < NUC_CHUNKS; cidx++) {
        const int p = (cidx * NUC_THREADS + (int)threadIdx.x) << 4;
        if (p >= tile_len) break;
        if (p + 16 > covered && covered < tile_len) {  // staging overflowed: slow path
            const int64_t j = mg_search_le(piece_off, p_lo, n_piece, P0 + p);
            nuc_chunk_generic(packed, piece_off, piece_src, j, P0 + p, total, T, lit, exc_pos, exc_byte, n_exc, out);
            continue;
        }
        int j = s_unit[p >> 6];
        while (s_rel[j + 1] <= p) j++;
        // first two genome pieces of the chunk, branch-free; up to two skippable pieces (suffix + prefix literal)
        // may sit in front of each
        int A = j;
        A += s_skip[A] & 1;
        A += s_skip[A] & 1;
        int B = A + 1;
        B += s_skip[B] & 1;
        B += s_skip[B] & 1;
        const int sA = s_rel[A], eA = s_rel[A + 1], sB = s_rel[B], eB = s_rel[B + 1];
        const bool okA = !s_skip[A] && sA < p + 16, okB = okA && !s_skip[B] && sB < p + 16;
        const uint64_t vA = mg_ld_nib16(packed, okA ? s_base[A] + p : (int64_t)MG_FRONT_PAD);
        const uint64_t vB = mg_ld_nib16(packed, okB ? s_base[B] + p : (int64_t)MG_FRONT_PAD);
        uint64_t nacc = 0;
        if (okA) nacc = vA & nib_range_mask(max(sA - p, 0), min(eA - p, 16));   // literal positions stay 0 and are patched below
        if (okB) nacc |= vB & nib_range_mask(sB - p, min(eB - p, 16));
        // anything beyond that inside these 16 bytes (a third genome piece, chains of blank records): rare
        const bool more = (s_skip[A] && sA < p + 16) || (okA && s_skip[B] && sB < p + 16) || (okB && eB < p + 16);
        if (more) {
            nacc = 0;
            for (int k = j; k < ncache && s_rel[k] < p + 16; k++) {
                if (!s_skip[k]) nacc |= mg_ld_nib16(packed, s_base[k] + p) & nib_range_mask(max(s_rel[k] - p, 0), min(s_rel[k + 1] - p, 16));
            }
        }
        uint32_t w0, w1, w2, w3;
        mg_decode8((uint32_t)nacc, w0, w1);
        mg_decode8((uint32_t)(nacc >> 32), w2, w3);
        // literal bytes (">ID\n", "\n") inside this chunk: skipped pieces in front of A, or right behind A inside the
        // chunk.  About 1.5 chunks per record take this path; a suffix and the following prefix that are adjacent in
        // lit[] share one 16-byte window load.
        if (MG_NUC_INLINE_LIT && (A != j || (okA && B != A + 1 && eA < p + 16) || more)) {
            uint32_t lw[4] = {0, 0, 0, 0};
            int64_t last = INT64_MIN;
            for (int k = j; k < ncache && s_rel[k] < p + 16; k++) {
                if ((s_skip[k] & 2) && s_rel[k + 1] > s_rel[k]) {
                    if (s_base[k] != last) { ld_lit16(lit, s_base[k] + p, lw); last = s_base[k]; }
                    const int lo = max(s_rel[k] - p, 0), hi = min(s_rel[k + 1] - p, 16);
                    const uint32_t m = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
                    const uint32_t m0 = expand4(m), m1 = expand4(m >> 4), m2 = expand4(m >> 8), m3 = expand4(m >> 12);
                    w0 = (w0 & ~m0) | (lw[0] & m0); w1 = (w1 & ~m1) | (lw[1] & m1);
                    w2 = (w2 & ~m2) | (lw[2] & m2); w3 = (w3 & ~m3) | (lw[3] & m3);
                }
            }
        }
        // code 15 = byte outside the packed alphabet on a '+' piece (the reverse plane already holds 'n',
        // genome.py:791-792): fetch the exact byte the FASTA had (genome.py:606 keeps it).  Rare.
        uint64_t e = nacc & (nacc >> 1) & (nacc >> 2) & (nacc >> 3) & 0x1111111111111111ull;
        if (e) {
            uint32_t w[4] = {w0, w1, w2, w3};
            while (e) {
                const int t = (__ffsll((long long)e) - 1) >> 2;
                e &= e - 1;
                int jj = j;
                while (s_rel[jj + 1] <= p + t) jj++;
                const int64_t gi = s_base[jj] + p + t;
                if (gi < T) {
                    const uint32_t b = mg_exc_byte(exc_pos, exc_byte, n_exc, gi);
                    w[t >> 2] = (w[t >> 2] & ~(0xFFu << ((t & 3) * 8))) | (b << ((t & 3) * 8));
                }
            }
            w0 = w[0]; w1 = w[1]; w2 = w[2]; w3 = w[3];
        }
        mg_st16(out + P0 + p, w0, w1, w2, w3);
    }
}

// ---- literal framing bytes of the PROTEIN text (">ID\n" prefixes, "\n" suffixes): one thread per literal piece ---
// Runs AFTER k_emit_prot on the same stream and overwrites the placeholder bytes it left there (K2 writes its own).
// which = 0: nucleotide text (positions from piece_off); which = 1: protein text (positions from prot_off).
__global__ void __launch_bounds__(256) k_emit_lit(int which, int64_t n_rec, const int64_t *__restrict__ rec_seg_off,
                                                  const int64_t *__restrict__ piece_off, const int64_t *__restrict__ prot_off,
                                                  const int32_t *__restrict__ rec_aa, const int64_t *__restrict__ rec_lit_off,
                                                  const int32_t *__restrict__ rec_pre, const int32_t *__restrict__ rec_suf,
                                                  const uint8_t *__restrict__ lit, uint8_t *__restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= 2 * n_rec) return;
    const int64_t r = i >> 1;
    const bool suffix = i & 1;
    const int pre = rec_pre[r];
    const int n = suffix ? rec_suf[r] : pre;
    if (n <= 0) return;
    const uint8_t *src = lit + rec_lit_off[r] + (suffix ? pre : 0);
    int64_t dst;
    if (which == 0) {
        const int64_t f0 = rec_seg_off[r] + 2 * r, f1 = rec_seg_off[r + 1] + 2 * (r + 1);
        dst = suffix ? piece_off[f1 - 1] : piece_off[f0];
    } else {
        int32_t naa = rec_aa[r];
        if (naa < 0) naa = 0;
        dst = prot_off[r] + (suffix ? pre + naa : 0);
    }
    for (int k = 0; k < n; k++) out[dst + k] = __ldg(src + k);
}

// ---- K3 -------------------------------------------------------------------------------------------------------
// generic fallback: one residue at a time from global memory (literal positions skipped)
__device__ __noinline__ void prot_chunk_generic(const uint32_t *__restrict__ packed, const int64_t *__restrict__ piece_off,
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

// Output chunk = 16 bytes of protein text.  Amino acid a of record r is the codon at spliced offset
// skip[r] + 3a; the 4096-entry nibble-triplet table (case-insensitive, anything non-ACGT -> 'X') sits in
// shared memory.  Stop codons are emitted as '*' and translation continues (genome.py:811-818).
__global__ void __launch_bounds__(PROT_THREADS) k_emit_prot(
    const uint32_t *__restrict__ packed, const int64_t *__restrict__ piece_off, const int64_t *__restrict__ piece_src,
    const int64_t *__restrict__ rec_seg_off, const int64_t *__restrict__ prot_off, const int32_t *__restrict__ rec_aa,
    const int8_t *__restrict__ rec_skip, const int32_t *__restrict__ rec_pre, int64_t n_rec,
    const int64_t *__restrict__ tile_first, int64_t total, const uint8_t *__restrict__ aa4096, uint8_t *__restrict__ out) {
    __shared__ __align__(16) uint8_t s_aa[4096];
    // records of the tile: residues occupy protein-text positions [r_s[i], r_e[i]) relative to the tile (empty when the
    // record has none); r_q0 = nucleotide-text position (relative to O) of the codon that would land on tile position 0;
    // r_j0 = first segment piece of the record (tile-local index)
    __shared__ int32_t r_s[PROT_RCAP + 2], r_e[PROT_RCAP + 2], r_q0[PROT_RCAP + 2];
    __shared__ int16_t r_j0[PROT_RCAP + 2];
    // pieces of those records: piece i covers nucleotide-text [s_rel[i], s_rel[i+1]) relative to O
    __shared__ int64_t s_base[PROT_PCAP + 4];
    __shared__ int32_t s_rel[PROT_PCAP + 5];
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
    const int64_t nraw64 = pc_hi - pc_lo;
    const bool fits = nrec <= PROT_RCAP && nraw64 <= PROT_PCAP && (__ldg(piece_off + pc_hi) - O) < (int64_t)BIG;
    const int tile_len = (int)min((int64_t)MG_PROT_TILE, total - P0);
    if (!fits) {                                      // rare: whole tile through the generic path
        for (int cidx = 0; cidx < PROT_CHUNKS; cidx++) {
            const int p = (cidx * PROT_THREADS + (int)threadIdx.x) << 4;
            if (p >= tile_len) break;
            const int64_t r = mg_search_le(prot_off, r_lo, n_rec, P0 + p);
            prot_chunk_generic(packed, piece_off, piece_src, rec_seg_off, prot_off, rec_aa, rec_skip, rec_pre, r, P0 + p, total,
                               aa4096, out);
        }
        return;
    }
    const int nraw = (int)nraw64;
    for (int i = threadIdx.x; i < nraw + 4; i += PROT_THREADS) {
        if (i <= nraw) {
            const int32_t rel = (int32_t)(__ldg(piece_off + pc_lo + i) - O);
            s_rel[i] = rel;
            s_base[i] = i < nraw ? (int64_t)((uint64_t)__ldg(piece_src + pc_lo + i) & MG_SRC_MASK) - rel : 0;
        } else {
            s_rel[i] = BIG;
            s_base[i] = 0;
        }
    }
    if (threadIdx.x == 0) s_rel[nraw + 4] = BIG;
    for (int i = threadIdx.x; i < nrec + 2; i += PROT_THREADS) {
        if (i < nrec) {
            const int64_t r = r_lo + i;
            int32_t naa = rec_aa[r];
            if (naa < 0) naa = 0;
            const int64_t rs = __ldg(prot_off + r) + rec_pre[r] - P0, re = rs + naa;
            const int64_t f0 = __ldg(rec_seg_off + r) + 2 * r;
            r_s[i] = rs < -BIG ? -BIG : (rs > MG_PROT_TILE ? MG_PROT_TILE : (int32_t)rs);
            r_e[i] = re < -BIG ? -BIG : (re > MG_PROT_TILE ? MG_PROT_TILE : (int32_t)re);
            r_q0[i] = (int32_t)(__ldg(piece_off + f0 + 1) - O) + rec_skip[r] - 3 * (int32_t)rs;
            r_j0[i] = (int16_t)(f0 + 1 - pc_lo);
        } else {
            r_s[i] = BIG; r_e[i] = BIG; r_q0[i] = 0; r_j0[i] = (int16_t)(nraw + 1);   // so that r_j0[R+1]-2 is the last record's suffix piece
        }
    }
    __syncthreads();
    // unit u -> first record whose residues end after byte 64*u
    for (int i = threadIdx.x; i <= nrec; i += PROT_THREADS) {
        int e0 = 0;
        if (i) { e0 = r_e[i - 1]; if (e0 < 0) e0 = 0; }
        int e1 = i < nrec ? r_e[i] : MG_PROT_TILE;
        if (e1 < 0) e1 = 0;
        const int u1 = min((e1 + 63) >> 6, PROT_UNITS);
        for (int u = i ? ((e0 + 63) >> 6) : 0; u < u1; u++) s_unit[u] = (uint16_t)i;
    }
    __syncthreads();

#pragma unroll 1
    for (int cidx = 0; cidx < PROT_CHUNKS; cidx++) {
        const int p = (cidx * PROT_THREADS + (int)threadIdx.x) << 4;
        if (p >= tile_len) break;
        int R = s_unit[p >> 6];
        while (r_e[R] <= p) R++;
        uint32_t bw0 = 0, bw1 = 0, bw2 = 0, bw3 = 0;
        for (; r_s[R] < p + 16; R++) {                // records with residues inside this chunk (usually one)
            const int lo = max(r_s[R] - p, 0), hi = min(r_e[R] - p, 16);
            if (hi <= lo) continue;                   // record without residues
            // nucleotide-text position (relative to O) of the codon that lands on chunk position 0
            const int q = r_q0[R] + 3 * p;
            const int need_lo = q + 3 * lo, need_hi = q + 3 * hi;        // nibbles [need_lo, need_hi)
            int j = r_j0[R];
            {   // last piece of the record that starts at or before need_lo (its pieces are contiguous)
                int a = j, b = r_j0[R + 1] - 2;       // one past the record's last segment piece
                while (b - a > 1) {
                    const int mid = (a + b) >> 1;
                    if (s_rel[mid] <= need_lo) a = mid; else b = mid;
                }
                j = a;
            }
            uint64_t acc[3];
#pragma unroll
            for (int g = 0; g < 3; g++) {
                const int qg = q + 16 * g;
                const int w_lo = max(qg, need_lo), w_hi = min(qg + 16, need_hi);
                uint64_t a = 0;
                if (w_hi > w_lo) {
                    while (s_rel[j + 1] <= w_lo) j++;
                    // first two pieces of this 16-nibble group, branch-free (segment pieces of a record are adjacent)
                    const int sA = s_rel[j], eA = s_rel[j + 1], eB = s_rel[j + 2];
                    const bool hasB = eA < w_hi && eB > eA;
                    const uint64_t vA = mg_ld_nib16(packed, s_base[j] + qg);
                    const uint64_t vB = mg_ld_nib16(packed, hasB ? s_base[j + 1] + qg : (int64_t)MG_FRONT_PAD);
                    a = vA & nib_range_mask(max(sA, w_lo) - qg, min(eA, w_hi) - qg);
                    if (hasB) a |= vB & nib_range_mask(eA - qg, min(eB, w_hi) - qg);
                    if (eA < w_hi && (!hasB || eB < w_hi)) {           // more than two pieces in 16 nibbles: rare
                        for (int jj = j + 1; s_rel[jj] < w_hi; jj++) {
                            if (s_rel[jj + 1] > s_rel[jj] && (jj > j + 1 || !hasB))
                                a |= mg_ld_nib16(packed, s_base[jj] + qg) & nib_range_mask(s_rel[jj] - qg, min(s_rel[jj + 1], w_hi) - qg);
                        }
                    }
                }
                acc[g] = a;
            }
            uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
            for (int k = 0; k < 16; k++) {            // codon k sits at bit 12k of acc[2]:acc[1]:acc[0]
                const int bit = 12 * k, ww = bit >> 6, sh = bit & 63;
                uint32_t idx = (uint32_t)(acc[ww] >> sh);
                if (sh > 52) idx |= (uint32_t)(acc[ww + 1] << (64 - sh));
                w[k >> 2] |= (uint32_t)s_aa[idx & 0xFFFu] << ((k & 3) * 8);
            }
            const uint32_t m = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
            if (m == 0xFFFFu) { bw0 = w[0]; bw1 = w[1]; bw2 = w[2]; bw3 = w[3]; }
            else {
                bw0 |= w[0] & expand4(m); bw1 |= w[1] & expand4(m >> 4); bw2 |= w[2] & expand4(m >> 8); bw3 |= w[3] & expand4(m >> 12);
            }
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
    cudaStream_t st = (cudaStream_t)stream;
    p->last_stream = st;
    k_emit_nuc<<<(unsigned)p->n_nuc_tile, NUC_THREADS, 0, st>>>(g->d_packed, p->d_piece_off, p->d_piece_src, p->n_piece, p->d_nuc_tile,
                                                              p->nuc_total, g->total_bases, p->d_lit, g->d_exc_pos, g->d_exc_byte, g->n_exc, out_dev);
    MG_LAUNCH_CHECK();
    if (!MG_NUC_INLINE_LIT && p->n_lit > 0) {
        k_emit_lit<<<(unsigned)((2 * p->n_rec + 255) / 256), 256, 0, st>>>(0, p->n_rec, p->d_rec_seg_off, p->d_piece_off, p->d_prot_off,
                                                                         p->d_rec_aa, p->d_rec_lit_off, p->d_rec_pre, p->d_rec_suf, p->d_lit, out_dev);
        MG_LAUNCH_CHECK();
    }
    return MG_OK;
}

extern "C" int mg_emit_prot_device(mg_plan *p, uint8_t *out_dev, void *stream) {
    MG_REQUIRE(p != nullptr, "plan handle is NULL");
    if (!p->prepared) { mg_set_error("mg_plan_prepare has not been called"); return MG_ESTATE; }
    if (p->prot_total == 0) return MG_OK;
    MG_REQUIRE(out_dev != nullptr && ((uintptr_t)out_dev & 15) == 0, "out_dev must be a 16-byte aligned device pointer");
    MG_CUDA(cudaSetDevice(p->device));
    mg_genome *g = p->g;
    cudaStream_t st = (cudaStream_t)stream;
    p->last_stream = st;
    k_emit_prot<<<(unsigned)p->n_prot_tile, PROT_THREADS, 0, st>>>(g->d_packed, p->d_piece_off, p->d_piece_src, p->d_rec_seg_off,
                                                                 p->d_prot_off, p->d_rec_aa, p->d_rec_skip, p->d_rec_pre, p->n_rec,
                                                                 p->d_prot_tile, p->prot_total, g->d_aa4096, out_dev);
    MG_LAUNCH_CHECK();
    if (p->n_lit > 0) {
        k_emit_lit<<<(unsigned)((2 * p->n_rec + 255) / 256), 256, 0, st>>>(1, p->n_rec, p->d_rec_seg_off, p->d_piece_off, p->d_prot_off,
                                                                         p->d_rec_aa, p->d_rec_lit_off, p->d_rec_pre, p->d_rec_suf, p->d_lit, out_dev);
        MG_LAUNCH_CHECK();
    }
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
