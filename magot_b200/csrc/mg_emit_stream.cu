// mg_emit_stream.cu -- K2 as a streaming gather without shared-memory tables, barriers or a tile prologue.
//
//   replaces ParentAnnotation.get_fasta seq_type="nucleotide" (genome.py:687-710), BaseAnnotation.get_seq (genome.py:603-608)
//   and Sequence.reverse_compliment (genome.py:784-793), FASTA framing included -- as k_emit_nuc (mg_emit.cu) does.
//
// Why (round 2, ncu profiles r2a-r2c in profiles/): every tile-staged variant of K2 -- per-lane global loads (round 1),
// one cp.async.bulk per piece into shared memory, 16-byte cp.async staging, per-chunk descriptors built by the piece's
// thread -- executes 300-340 warp-instructions per KB of text, of which the tile prologue (classify ~200 pieces, block
// scan, unit tables, 2-5 barriers) is ~150 and the "which piece am I in, where does the next one start" arithmetic of
// the main loop another ~90.  All of them end up within +-15 % of each other, co-limited by issue rate, the L1 data pipe
// and barrier/latency bubbles.  This kernel removes the prologue instead of moving it:
//
//   Phase A  one warp = one 1 KB block of the text = 32 chunks of 32 bytes.  K1 leaves the index of the piece that holds the
//            first byte of every 1 KB block (mg_plan.cu: blk1k, filled by scatter).  The warp loads the offsets and sources of
//            the 32 pieces from there on with two coalesced loads, and every lane finds the piece of its chunk by a
//            five-step binary search over the lanes' registers (SHFL).  A chunk that lies inside one piece ("simple",
//            75-85 % of config 4) is fetched, decoded and stored right away; literal (framing) chunks likewise.
//   Phase B  thread per piece of the CTA's blocks: the chunk in which the piece ENDS (if it ends inside a chunk and holds
//            that chunk's first byte) is assembled from every piece that reaches into it, genome and literal alike.
//            One such chunk per piece: the threads are 2/3 to 5/6 busy without any list or compaction.
//   No thread ever waits for another one: no shared memory, no barrier, no write-twice of the framing bytes.
#include <algorithm>
#include "mg_common.cuh"
#include "mg_gather.cuh"
#include "mg_emit_common.cuh"

#ifndef S4_THREADS
#define S4_THREADS 256
#endif
#ifndef S4_MINB
#define S4_MINB 6
#endif
#define S4_ROWS (MG_NUC_TILE / 1024)              // 1 KB blocks per CTA
#define S4_WARPS (S4_THREADS / 32)

__device__ __forceinline__ int64_t shfl64(int64_t v, int src) {
    const int lo = __shfl_sync(0xffffffffu, (int)(uint32_t)v, src);
    const int hi = __shfl_sync(0xffffffffu, (int)(v >> 32), src);
    return ((int64_t)hi << 32) | (uint32_t)lo;
}

__device__ __forceinline__ void s4_window(const uint32_t *__restrict__ packed, int64_t g, uint32_t n[4]) {
    uint32_t r[5];
    ld_pk5(packed + (g >> 3), r);
    const uint32_t bs = ((uint32_t)g & 7u) << 2;
#pragma unroll
    for (int k = 0; k < 4; k++) n[k] = __funnelshift_r(r[k], r[k + 1], bs);
}

__device__ __forceinline__ bool s4_has_code15(const uint32_t n[4]) {
    uint32_t rare = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t e = n[k] & (n[k] >> 1);
        rare |= e & (e >> 2) & 0x11111111u;
    }
    return rare != 0;
}

// 32 nibbles -> 32 ASCII bytes; chunks without N / IUPAC / '-' codes (bit 3 clear in every nibble) need two table look-ups
// per word instead of four plus the select masks
__device__ __forceinline__ void s4_decode32(const uint32_t n[4], uint32_t w[8]) {
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

__global__ void __launch_bounds__(S4_THREADS, S4_MINB) k_emit_nuc_stream(
    const uint32_t *__restrict__ packed, const int64_t *__restrict__ piece_off, const int64_t *__restrict__ piece_src,
    int64_t n_piece, const int32_t *__restrict__ blk1k, const int64_t *__restrict__ total_dev, int64_t cap, int64_t T,
    const uint8_t *__restrict__ lit, const int64_t *__restrict__ exc_pos, const uint8_t *__restrict__ exc_byte, int64_t n_exc,
    uint8_t *__restrict__ out) {
    const int64_t total = min(__ldg(total_dev), cap);
    const int64_t P0 = (int64_t)blockIdx.x * MG_NUC_TILE;
    if (P0 >= total) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t n_rows = (total + 1023) >> 10;
    const int64_t row0 = (int64_t)blockIdx.x * S4_ROWS;

    // ---- Phase A: chunks that lie inside one piece
#pragma unroll 1
    for (int rr = wid; rr < S4_ROWS; rr += S4_WARPS) {
        const int64_t row = row0 + rr;
        if (row >= n_rows) break;
        const int64_t B0 = row << 10;
        const int64_t j0 = __ldg(blk1k + row);                       // piece that holds the block's first byte
        const int64_t jj = j0 + lane;
        const int64_t off = jj <= n_piece ? __ldg(piece_off + jj) : (int64_t)1 << 60;
        const int64_t src = jj < n_piece ? __ldg(piece_src + jj) : 0;
        const int64_t d = off - B0;
        const int r = (int)max(min(d, (int64_t)1 << 30), -((int64_t)1 << 30));    // block-relative start of piece j0 + lane
        const int tL = lane << 5;
        int idx = 0;                                                  // largest i with r_i <= tL (r_0 <= 0; empty pieces come first
#pragma unroll                                                        // among equal starts, so the last one holds the byte)
        for (int step = 16; step; step >>= 1) {
            const int v = __shfl_sync(0xffffffffu, r, idx + step);
            if (v <= tL) idx += step;
        }
        int64_t o_j = shfl64(off, idx), s_j = shfl64(src, idx);
        int r_n = __shfl_sync(0xffffffffu, r, min(idx + 1, 31));
        const int64_t C = B0 + tL;
        if (C >= total) continue;
        int64_t j = j0 + idx;
        if (idx == 31) {                                              // more than 31 pieces start in this block before the chunk: search
            j = mg_search_le(piece_off, j, n_piece, C);
            o_j = __ldg(piece_off + j);
            s_j = __ldg(piece_src + j);
            r_n = (int)min(__ldg(piece_off + j + 1) - B0, (int64_t)1 << 30);
        }
        const int end = (int)min((int64_t)32, total - C);
        if (r_n - tL < end) continue;                                 // a piece ends inside the chunk: Phase B
        uint32_t w[8];
        const int64_t a = (int64_t)((uint64_t)s_j & MG_SRC_MASK) + (C - o_j);
        if (((uint64_t)s_j >> MG_KIND_SHIFT) == MG_KIND_LIT) {
            ld_lit16(lit, a, w);
            ld_lit16(lit, a + 16, w + 4);
        } else {
            uint32_t n[4];
            s4_window(packed, a, n);
            // code 15 = byte outside the packed alphabet on a '+' piece: the exact FASTA byte must come out (genome.py:606)
            if (n_exc > 0 && s4_has_code15(n)) {
                nuc_chunk_generic(packed, piece_off, piece_src, j, C, total, T, lit, exc_pos, exc_byte, n_exc, out);
                continue;
            }
            s4_decode32(n, w);
        }
        st32(out + C, w);
    }

    // ---- Phase B: the chunk in which a piece ends, assembled by the thread of the piece that holds the chunk's first byte
    const int64_t row_end = min(row0 + S4_ROWS, n_rows);
    const int64_t j_lo = __ldg(blk1k + row0);
    const int64_t j_hi = row_end < n_rows ? (int64_t)__ldg(blk1k + row_end) : n_piece - 1;
#pragma unroll 1
    for (int64_t j = j_lo + tid; j <= j_hi; j += S4_THREADS) {
        int64_t ok = __ldg(piece_off + j), ok1 = __ldg(piece_off + j + 1);
        if (ok1 == ok || (ok1 & 31) == 0 || ok1 >= total) continue;
        const int64_t C = ok1 & ~(int64_t)31;
        if (C < P0 || C >= P0 + MG_NUC_TILE || ok > C) continue;      // another CTA's chunk / another piece holds its first byte
        const int end = (int)min((int64_t)32, total - C);
        uint32_t n[4] = {0, 0, 0, 0}, lw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint32_t lmask = 0;
        int64_t k = j;
        while (true) {
            const int lo = (int)max(ok - C, (int64_t)0), hi = (int)min(ok1 - C, (int64_t)end);
            if (hi > lo) {
                const uint64_t sk = (uint64_t)__ldg(piece_src + k);
                const int64_t a = (int64_t)(sk & MG_SRC_MASK) + (C - ok);
                if ((sk >> MG_KIND_SHIFT) == MG_KIND_LIT) {
                    uint32_t b[8];
                    ld_lit16(lit, a, b);
                    ld_lit16(lit, a + 16, b + 4);
                    const uint32_t m = (hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
                    lmask |= m;
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const uint32_t mk = expand4(m >> (4 * q));
                        lw[q] = (lw[q] & ~mk) | (b[q] & mk);
                    }
                } else {                                              // overwrites the chunk from position lo on
                    uint32_t y[4];
                    s4_window(packed, a, y);
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t keep = low_nibbles(lo - 8 * q);
                        n[q] = (n[q] & keep) | (y[q] & ~keep);
                    }
                }
            }
            if (ok1 - C >= end) break;
            k++;
            ok = ok1;
            ok1 = __ldg(piece_off + k + 1);
        }
        if (n_exc > 0 && s4_has_code15(n)) {
            nuc_chunk_generic(packed, piece_off, piece_src, j, C, total, T, lit, exc_pos, exc_byte, n_exc, out);
            continue;
        }
        uint32_t w[8];
        s4_decode32(n, w);
        if (lmask) {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint32_t mk = expand4(lmask >> (4 * q));
                w[q] = (w[q] & ~mk) | (lw[q] & mk);
            }
        }
        st32(out + C, w);
    }
}

int mg_launch_nuc_stream(mg_plan *p, uint8_t *out_dev, cudaStream_t st) {
    mg_genome *g = p->g;
    if (!p->blk1k_ready) {                           // the plan was prepared while another K2 variant was selected
        mg_set_error("the 1 KB block table of this plan was not built: select the streaming variant (mg_tune(\"emit\", 2)) before mg_plan_prepare");
        return MG_ESTATE;
    }
    k_emit_nuc_stream<<<(unsigned)p->n_nuc_tile, S4_THREADS, 0, st>>>(g->d_packed, p->d_piece_off, p->d_piece_src, p->n_piece, p->d_blk1k,
                                                                     p->d_totals, p->nuc_total, g->total_bases, p->d_lit, g->d_exc_pos,
                                                                     g->d_exc_byte, g->n_exc, out_dev);
    MG_LAUNCH_CHECK();
    return MG_OK;
}
