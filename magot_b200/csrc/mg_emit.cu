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
#include <algorithm>
#include "mg_common.cuh"
#include "mg_gather.cuh"

#define NUC_THREADS 256
#define NUC_CHUNKS (MG_NUC_TILE / 16 / NUC_THREADS)     // 4 chunks of 16 B per thread
#define NUC_CAP 768                                      // pieces cached in shared memory per tile

#define PROT_THREADS 256
#define PROT_CHUNKS (MG_PROT_TILE / 16 / PROT_THREADS)  // 2
#define PROT_CAP 256                                     // records cached per tile

// expand the low 4 bits of x into a byte mask (bit k -> byte k = 0xFF)
__device__ __forceinline__ uint32_t expand4(uint32_t x) {
    return ((x & 1u) | ((x & 2u) << 7) | ((x & 4u) << 14) | ((x & 8u) << 21)) * 0xFFu;
}

__global__ void __launch_bounds__(NUC_THREADS) k_emit_nuc(
    const uint32_t *__restrict__ packed, const int64_t *__restrict__ piece_off, const int64_t *__restrict__ piece_src,
    int64_t n_piece, const int64_t *__restrict__ tile_first, int64_t total, const uint8_t *__restrict__ lit,
    const int64_t *__restrict__ exc_pos, const uint8_t *__restrict__ exc_byte, int64_t n_exc, uint8_t *__restrict__ out) {
    __shared__ int64_t s_off[NUC_CAP + 1];
    __shared__ int64_t s_src[NUC_CAP];
    const int64_t tile = blockIdx.x;
    const int64_t P0 = tile * MG_NUC_TILE;
    const int64_t p_lo = tile_first[tile];
    int64_t p_hi = tile_first[tile + 1] + 1;           // one past the last piece this tile can touch
    if (p_hi > n_piece) p_hi = n_piece;
    const int ncache = (int)min((int64_t)NUC_CAP, p_hi - p_lo);
    for (int i = threadIdx.x; i <= ncache; i += NUC_THREADS) {
        s_off[i] = __ldg(piece_off + p_lo + i);
        if (i < ncache) s_src[i] = __ldg(piece_src + p_lo + i);
    }
    __syncthreads();
    const int64_t cached_end = s_off[ncache];          // text offset where the cached pieces end

#pragma unroll 1
    for (int cidx = 0; cidx < NUC_CHUNKS; cidx++) {
        const int64_t P = P0 + ((int64_t)(cidx * NUC_THREADS + threadIdx.x) << 4);
        if (P >= total) break;
        // ---- locate the piece that holds byte P: largest j with off[j] <= P
        int64_t j;
        if (P < cached_end) {
            int lo = 0, hi = ncache;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_off[mid] <= P) lo = mid; else hi = mid;
            }
            j = p_lo + lo;
        } else {
            j = mg_search_le(piece_off, p_lo + ncache, n_piece, P);
        }
        // ---- assemble 16 output bytes piece by piece
        uint64_t nacc = 0;                              // nibble codes, decoded at the end
        uint64_t blo = 0, bhi = 0;                      // raw bytes (literals, exceptions)
        uint32_t bm = 0;                                // which of the 16 bytes are raw
        int filled = 0;
        int64_t pos = P;
        int64_t off_j, off_n;
        {
            const int64_t k = j - p_lo;
            off_j = (k <= ncache) ? s_off[k] : __ldg(piece_off + j);
            off_n = (k + 1 <= ncache) ? s_off[k + 1] : __ldg(piece_off + j + 1);
        }
        while (filled < 16 && pos < total) {
            while (off_n <= pos) {                      // advance over finished / empty pieces
                j++;
                off_j = off_n;
                const int64_t k = j + 1 - p_lo;
                off_n = (k <= ncache) ? s_off[k] : __ldg(piece_off + j + 1);
            }
            const int64_t kk = j - p_lo;
            const uint64_t sk = (uint64_t)((kk < ncache) ? s_src[kk] : __ldg(piece_src + j));
            const uint64_t kind = sk >> MG_KIND_SHIFT;
            const int64_t src = (int64_t)(sk & MG_SRC_MASK);
            const int64_t o = pos - off_j;
            int c = 16 - filled;
            if (off_n - pos < c) c = (int)(off_n - pos);
            if (kind == MG_KIND_LIT) {
                for (int k = 0; k < c; k++) {
                    const uint64_t b = __ldg(lit + src + o + k);
                    const int q = filled + k;
                    if (q < 8) blo |= b << (8 * q); else bhi |= b << (8 * (q - 8));
                }
                bm |= ((1u << c) - 1u) << filled;
            } else {
                uint64_t v;
                if (kind == MG_KIND_FWD) {
                    v = mg_ld_nib16(packed, src + o);
                    // code 15 = byte outside the packed alphabet: fetch the exact byte (genome.py:606 keeps it)
                    uint64_t e = v & (v >> 1) & (v >> 2) & (v >> 3) & 0x1111111111111111ull;
                    if (c < 16) e &= (1ull << (4 * c)) - 1ull;
                    while (e) {
                        const int k = (__ffsll((long long)e) - 1) >> 2;
                        e &= e - 1;
                        const uint64_t b = mg_exc_byte(exc_pos, exc_byte, n_exc, src + o + k);
                        const int q = filled + k;
                        if (q < 8) blo |= b << (8 * q); else bhi |= b << (8 * (q - 8));
                        bm |= 1u << q;
                    }
                } else {
                    v = mg_rc_nib16(mg_ld_nib16(packed, src + (off_n - off_j) - o - 16));
                }
                if (c < 16) v &= (1ull << (4 * c)) - 1ull;
                nacc |= v << (4 * filled);
            }
            filled += c;
            pos += c;
        }
        uint32_t w0, w1, w2, w3;
        mg_decode8((uint32_t)nacc, w0, w1);
        mg_decode8((uint32_t)(nacc >> 32), w2, w3);
        if (bm) {
            const uint32_t m0 = expand4(bm), m1 = expand4(bm >> 4), m2 = expand4(bm >> 8), m3 = expand4(bm >> 12);
            w0 = (w0 & ~m0) | ((uint32_t)blo & m0);
            w1 = (w1 & ~m1) | ((uint32_t)(blo >> 32) & m1);
            w2 = (w2 & ~m2) | ((uint32_t)bhi & m2);
            w3 = (w3 & ~m3) | ((uint32_t)(bhi >> 32) & m3);
        }
        mg_st16(out + P, w0, w1, w2, w3);
    }
}

// ---- K3 -------------------------------------------------------------------------------------------------
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
    __shared__ int64_t s_off[PROT_CAP + 1];
    reinterpret_cast<uint4 *>(s_aa)[threadIdx.x] = __ldg(reinterpret_cast<const uint4 *>(aa4096) + threadIdx.x);
    const int64_t tile = blockIdx.x;
    const int64_t P0 = tile * MG_PROT_TILE;
    const int64_t r_lo = tile_first[tile];
    int64_t r_hi = tile_first[tile + 1] + 1;
    if (r_hi > n_rec) r_hi = n_rec;
    const int ncache = (int)min((int64_t)PROT_CAP, r_hi - r_lo);
    for (int i = threadIdx.x; i <= ncache; i += PROT_THREADS) s_off[i] = __ldg(prot_off + r_lo + i);
    __syncthreads();
    const int64_t cached_end = s_off[ncache];

#pragma unroll 1
    for (int cidx = 0; cidx < PROT_CHUNKS; cidx++) {
        const int64_t P = P0 + ((int64_t)(cidx * PROT_THREADS + threadIdx.x) << 4);
        if (P >= total) break;
        int64_t r;
        if (P < cached_end) {
            int lo = 0, hi = ncache;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_off[mid] <= P) lo = mid; else hi = mid;
            }
            r = r_lo + lo;
        } else {
            r = mg_search_le(prot_off, r_lo + ncache, n_rec, P);
        }
        uint64_t blo = 0, bhi = 0;
        int filled = 0;
        int64_t pos = P;
        int64_t off_r = __ldg(prot_off + r), off_n = __ldg(prot_off + r + 1);
        while (filled < 16 && pos < total) {
            while (off_n <= pos) {
                r++;
                off_r = off_n;
                off_n = __ldg(prot_off + r + 1);
            }
            const int64_t q = pos - off_r;
            const int64_t pre = rec_pre[r];
            int32_t naa = rec_aa[r];
            if (naa < 0) naa = 0;
            int c = 16 - filled;
            if (q < pre) {                                        // literal prefix
                if (pre - q < c) c = (int)(pre - q);
                const uint8_t *lp = lit + rec_lit_off[r] + q;
                for (int k = 0; k < c; k++) {
                    const uint64_t b = __ldg(lp + k);
                    const int t = filled + k;
                    if (t < 8) blo |= b << (8 * t); else bhi |= b << (8 * (t - 8));
                }
            } else if (q < pre + naa) {                           // amino acids
                const int64_t a = q - pre;
                if (naa - a < c) c = (int)(naa - a);
                const int64_t f0 = __ldg(rec_seg_off + r) + 2 * r, f1 = __ldg(rec_seg_off + r + 1) + 2 * (r + 1);
                const int64_t S = __ldg(piece_off + f0 + 1) + rec_skip[r] + 3 * a;
                const int64_t j = mg_search_le(piece_off, f0 + 1, f1 - 1, S);
                uint64_t acc[3];
                mg_gather_nib(packed, piece_off, piece_src, j, S, 3 * c, acc);
                // 16 codons = 192 bits; codon k sits at bit 12k
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    if (k < c) {
                        const int bit = 12 * k, w = bit >> 6, sh = bit & 63;
                        uint32_t idx = (uint32_t)(acc[w] >> sh);
                        if (sh > 52) idx |= (uint32_t)(acc[w + 1] << (64 - sh));
                        const uint64_t b = s_aa[idx & 0xFFFu];
                        const int t = filled + k;
                        if (t < 8) blo |= b << (8 * t); else bhi |= b << (8 * (t - 8));
                    }
                }
            } else {                                              // literal suffix
                const int64_t sq = q - pre - naa;
                if (off_n - pos < c) c = (int)(off_n - pos);
                const uint8_t *lp = lit + rec_lit_off[r] + pre + sq;
                for (int k = 0; k < c; k++) {
                    const uint64_t b = __ldg(lp + k);
                    const int t = filled + k;
                    if (t < 8) blo |= b << (8 * t); else bhi |= b << (8 * (t - 8));
                }
            }
            filled += c;
            pos += c;
        }
        mg_st16(out + P, (uint32_t)blo, (uint32_t)(blo >> 32), (uint32_t)bhi, (uint32_t)(bhi >> 32));
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
