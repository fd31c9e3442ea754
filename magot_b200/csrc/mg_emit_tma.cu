// mg_emit_tma.cu -- K2 with the tile's packed genome bytes staged in shared memory by per-piece bulk copies
// (cp.async.bulk + mbarrier: the copy engine, not the warps, waits for DRAM).  An A/B variant (MAGOT_EMIT=tma / mg_tune("emit", 1)):
// measured slower than the per-lane loads of mg_emit.cu (CDS 0.118-0.122 vs 0.107 ms, exon 0.169-0.173 vs 0.155 ms on config 4;
// profiles/r2a_bulk_copy_ab.jsonl), kept as a tested variant.  (The fused splice + translate kernel K23 lives in mg_emit.cu.)
//
// Why: the round-1 K2 (mg_emit.cu) fetched every 32-nibble window with per-lane global loads (3 x 8 B for the first piece
// of a chunk, 3 x 8 B for the second): ~31 of its ~60 L1 data-pipe wavefronts per KB of text, L1 data pipe 72 % busy,
// 44 % of the stall samples on the load-use scoreboard.  Here a piece is ONE bulk copy of its 16-byte granules into
// shared memory (issued by the thread that classified the piece; ~200 copies in flight per CTA), the lanes then read
// aligned 16-byte granules (conflict-free LDS.128) and shift in registers.
//
//   K2  replaces ParentAnnotation.get_fasta seq_type="nucleotide" (genome.py:687-710), BaseAnnotation.get_seq
//       (genome.py:603-608) and Sequence.reverse_compliment (genome.py:784-793) -- as k_emit_nuc does.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "mg_common.cuh"
#include "mg_gather.cuh"
#include "mg_emit_common.cuh"

#ifndef T2_THREADS
#define T2_THREADS 256
#endif
#define T2_WARPS (T2_THREADS / 32)
#ifndef T2_MINB
#define T2_MINB 4                   // 64 registers: the loop keeps the decode tables, two windows and the chunk in registers
#endif
#ifndef T2_STAGE
#define T2_STAGE 0                  // 0 = one cp.async.bulk (UBLKCP) per piece + mbarrier into shared memory;
                                    // 2 = no staging: the lanes load their windows from global memory (as mg_emit.cu does)
#endif
#ifndef T2_CAP
#define T2_CAP 1024                 // pieces of any kind examined per tile (rounds of T2_THREADS)
#endif
#ifndef T2_GCAP
#define T2_GCAP 640                 // non-empty genome pieces staged per tile (a 32 KB tile of config 4 has ~180)
#endif
#ifndef T2_RAWG
#define T2_RAWG 1408                // 16-byte granules of staged packed bytes (22 KB for the 16 KB a tile needs + ~15 B of
                                    // alignment slack at both ends of every piece)
#endif
#define T2_PADG 2                   // granules of padding in front (windows that start before the first piece)
#ifndef T2_LITCAP
#define T2_LITCAP 160               // literal (framing) pieces listed per tile (config 4: ~40)
#endif
#define T2_UNITS (MG_NUC_TILE / 32)
#define T2_CHUNKS ((MG_NUC_TILE / 32 + T2_THREADS - 1) / T2_THREADS)
#ifndef T2_LDS128
#define T2_LDS128 0
#endif

// ---- mbarrier / bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP) ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MG_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MG_DONE_%=;\n"
        "bra MG_WAIT_%=;\n"
        "MG_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
// global -> shared bulk copy of `bytes` (multiple of 16; both addresses 16-byte aligned), completion counted on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

// 16-byte asynchronous copy global -> shared (LDGSTS), L2 fills 64-byte halves only (a piece is ~100 packed bytes at a random
// address: with whole 128-byte lines DRAM delivers 1.25 x the bytes, measured 453 vs 368 MB on the exon launch of config 4)
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
}

// 32 nibbles -> 32 ASCII bytes; chunks without N / IUPAC / '-' codes (bit 3 clear in every nibble) need two table look-ups
// per word instead of four plus the select masks
__device__ __forceinline__ void decode32(const uint32_t n[4], uint32_t w[8]) {
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

// 32 nibbles starting at nibble index gx of the staged buffer (granule index clamped into the buffer: windows that lie in
// framing placeholders may point anywhere)
__device__ __forceinline__ void lds_window(const uint4 *s_raw, int gx, uint32_t n[4]) {
    int gi = gx >> 5;
    gi = min(max(gi, 0), T2_PADG + T2_RAWG);
    const uint32_t bs = ((uint32_t)gx & 7u) << 2;
#if T2_LDS128
    const uint4 a = s_raw[gi], b = s_raw[gi + 1];
    const bool w2 = (gx & 16) != 0, w1 = (gx & 8) != 0;
    const uint32_t t0 = w2 ? a.z : a.x, t1 = w2 ? a.w : a.y, t2 = w2 ? b.x : a.z, t3 = w2 ? b.y : a.w, t4 = w2 ? b.z : b.x,
                   t5 = w2 ? b.w : b.y;
    const uint32_t r0 = w1 ? t1 : t0, r1 = w1 ? t2 : t1, r2 = w1 ? t3 : t2, r3 = w1 ? t4 : t3, r4 = w1 ? t5 : t4;
#else
    // three 8-byte loads from the enclosing 8-byte aligned window + five selects (as ld_pk5 does from global memory)
    const uint2 *q = reinterpret_cast<const uint2 *>(s_raw + gi) + ((gx >> 4) & 1);
    const uint2 a = q[0], b = q[1], c = q[2];
    const bool w1 = (gx & 8) != 0;
    const uint32_t r0 = w1 ? a.y : a.x, r1 = w1 ? b.x : a.y, r2 = w1 ? b.y : b.x, r3 = w1 ? c.x : b.y, r4 = w1 ? c.y : c.x;
#endif
    n[0] = __funnelshift_r(r0, r1, bs);
    n[1] = __funnelshift_r(r1, r2, bs);
    n[2] = __funnelshift_r(r2, r3, bs);
    n[3] = __funnelshift_r(r3, r4, bs);
}

// ---- K2, v3 ------------------------------------------------------------------------------------------------------
// Tile = MG_NUC_TILE bytes of nucleotide text.
// Prologue (thread per piece, rounds of T2_THREADS): the tile's pieces are classified; non-empty genome pieces are
// compacted by a block scan that (T2_STAGE 0) also hands out their place in the staging buffer, where one bulk copy per
// piece puts their 16-byte granules.  Compact piece k "owns" text [start_k, start_{k+1}).  The thread of piece k then writes
// ONE WORD PER 32-BYTE CHUNK it owns: the source nibble index of the chunk's first byte, or "skip" for chunks that hold
// only framing bytes -- and lists the one chunk in which piece k+1 starts.
// Pass 1: a lane reads that word, fetches its 32-nibble window, decodes and stores: no piece look-ups, no boundary
// arithmetic (the round-1 kernel paid ~90 instructions per chunk for "which piece, where does the next one start, load
// and merge both", and 99 % of its warp-iterations had a boundary in some lane).
// Pass 2: the listed boundary chunks, one per genome piece, densely packed onto the threads: every piece that starts inside
// the chunk overwrites it from its first position on.
// Epilogue: framing bytes (">ID\n", "\n") over the placeholders the two passes left.
struct t2_chunk_ctx {
    const uint32_t *packed; const int64_t *piece_off; const int64_t *piece_src; int64_t n_piece, p_lo, P0, total, T;
    const uint8_t *lit; const int64_t *exc_pos; const uint8_t *exc_byte; int64_t n_exc; uint8_t *out;
};

// exact bytes for a chunk that holds a byte outside the packed alphabet (code 15): generic path from global memory
static __device__ __noinline__ void t2_chunk_exact(const t2_chunk_ctx &c, int p) {
    const int64_t j = mg_search_le(c.piece_off, c.p_lo, c.n_piece, c.P0 + p);
    nuc_chunk_generic(c.packed, c.piece_off, c.piece_src, j, c.P0 + p, c.total, c.T, c.lit, c.exc_pos, c.exc_byte, c.n_exc, c.out);
}

__device__ __forceinline__ bool t2_has_code15(const uint32_t n[4]) {
    uint32_t rare = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t e = n[k] & (n[k] >> 1);
        rare |= e & (e >> 2) & 0x11111111u;
    }
    return rare != 0;
}

#if T2_STAGE == 0
typedef int32_t t2_src_t;                              // nibble index into the staging buffer
#define T2_SKIP ((int32_t)0x80000000)
__device__ __forceinline__ void t2_window(const uint4 *s_raw, const uint32_t *, t2_src_t gx, uint32_t n[4]) { lds_window(s_raw, gx, n); }
#else
typedef int64_t t2_src_t;                              // global nibble index (both planes)
#define T2_SKIP ((int64_t)-1)
__device__ __forceinline__ void t2_window(const uint4 *, const uint32_t *packed, t2_src_t gx, uint32_t n[4]) {
    uint32_t r[5];
    ld_pk5(packed + (gx >> 3), r);
    const uint32_t bs = ((uint32_t)gx & 7u) << 2;
#pragma unroll
    for (int k = 0; k < 4; k++) n[k] = __funnelshift_r(r[k], r[k + 1], bs);
}
#endif

__global__ void __launch_bounds__(T2_THREADS, T2_MINB) k_emit_nuc_tma(
    const uint32_t *__restrict__ packed, const int64_t *__restrict__ piece_off, const int64_t *__restrict__ piece_src,
    int64_t n_piece, const int64_t *__restrict__ tile_first, const int64_t *__restrict__ total_dev, int64_t cap, int64_t T,
    const uint8_t *__restrict__ lit, const int64_t *__restrict__ exc_pos, const uint8_t *__restrict__ exc_byte, int64_t n_exc,
    uint8_t *__restrict__ out) {
    const int64_t total = min(__ldg(total_dev), cap);
    const int64_t P0 = (int64_t)blockIdx.x * MG_NUC_TILE;
    if (P0 >= total) return;
#if T2_STAGE == 0
    __shared__ __align__(128) uint4 s_raw[T2_PADG + T2_RAWG + 2];
    __shared__ __align__(8) uint64_t s_mbar;
#else
    const uint4 *s_raw = nullptr;
#endif
    __shared__ t2_src_t s_base[T2_GCAP + 3];          // source nibble index of tile position 0, per compact piece
    __shared__ int32_t s_start[T2_GCAP + 3];          // tile-relative start of the piece (may be negative), BIG after the last
    __shared__ int32_t s_end[T2_GCAP + 3];            // tile-relative end of its bases inside the tile
    __shared__ uint16_t s_def[T2_GCAP + 3];           // chunk in which the NEXT piece starts (0xFFFF: none)
    __shared__ t2_src_t s_gx[T2_UNITS];               // per 32-byte chunk: source nibble index of its first byte, or T2_SKIP
    __shared__ int32_t s_la[T2_LITCAP], s_lb[T2_LITCAP];
    __shared__ int64_t s_lsrc[T2_LITCAP];             // literal byte of tile position q = lit[s_lsrc + q]
    __shared__ uint32_t s_wtot[T2_WARPS];
    __shared__ int s_nlit, s_bad;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t p_lo = tile_first[blockIdx.x];
    int64_t p_hi = tile_first[blockIdx.x + 1] + 1;    // one past the last piece this tile can touch
    if (p_hi > n_piece) p_hi = n_piece;
    const int tile_len = (int)min((int64_t)MG_NUC_TILE, total - P0);
    const int nraw = (int)min((int64_t)T2_CAP, p_hi - p_lo);
    if (tid == 0) {
        s_nlit = 0;
        s_bad = (p_hi - p_lo > T2_CAP) ? 1 : 0;
#if T2_STAGE == 0
        mbar_init(&s_mbar, 1);
#endif
    }
    __syncthreads();
    uint32_t carry = (uint32_t)T2_PADG << 11;         // (next free granule << 11) | compact pieces so far
    for (int base = 0; base < nraw; base += T2_THREADS) {
        const int i = base + tid;
        uint32_t pack = 0;
        int kind = PIECE_E, lo = 0, hi = 0, relc = 0, ng = 0;
        int64_t g = 0, lsrc = 0;
        if (i < nraw) {
            const int64_t rel = __ldg(piece_off + p_lo + i) - P0, nxt = __ldg(piece_off + p_lo + i + 1) - P0;
            const uint64_t sk = (uint64_t)__ldg(piece_src + p_lo + i);
            lo = (int)max(rel, (int64_t)0);
            hi = (int)min(nxt, (int64_t)tile_len);
            if (hi > lo) {
                if ((sk >> MG_KIND_SHIFT) == MG_KIND_LIT) {
                    kind = PIECE_L;
                    lsrc = (int64_t)(sk & MG_SRC_MASK) - rel;
                } else {
                    kind = PIECE_G;
                    g = (int64_t)(sk & MG_SRC_MASK) + (lo - rel);             // nibble index of tile position lo
                    ng = (int)(((g + (hi - lo) - 1) >> 5) - (g >> 5)) + 1;
                    pack = ((uint32_t)ng << 11) | 1u;
                    relc = (int)max(rel, (int64_t)-(1 << 30));
                }
            }
        }
        uint32_t incl = pack;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_wtot[wid] = incl;
        __syncthreads();
        uint32_t woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < T2_WARPS; w++) {
            const uint32_t t = s_wtot[w];
            if (w < wid) woff += t;
            tot += t;
        }
        const uint32_t excl = carry + woff + incl - pack;
        carry += tot;
        __syncthreads();
        if (kind == PIECE_G) {
            const int k = (int)(excl & 2047u), so = (int)(excl >> 11);
            if (k >= T2_GCAP || (T2_STAGE == 0 && so + ng > T2_PADG + T2_RAWG)) {
                s_bad = 1;
            } else {
                s_start[k] = relc;
                s_end[k] = hi;
#if T2_STAGE == 0
                s_base[k] = so * 32 + (int)(g & 31) - lo;
                const uint32_t bytes = (uint32_t)ng * 16u;
                mbar_expect_tx(&s_mbar, bytes);
                bulk_g2s(&s_raw[so], reinterpret_cast<const uint4 *>(packed) + (g >> 5), bytes, &s_mbar);
#else
                s_base[k] = g - lo;
#endif
            }
        } else if (kind == PIECE_L) {
            const int slot = atomicAdd(&s_nlit, 1);
            if (slot < T2_LITCAP) { s_la[slot] = lo; s_lb[slot] = hi; s_lsrc[slot] = lsrc; }
        }
    }
    const int ngen = min((int)(carry & 2047u), T2_GCAP);
    if (tid < 3) { s_start[ngen + tid] = BIG; s_end[ngen + tid] = 0; s_base[ngen + tid] = 0; }
    __syncthreads();
#if T2_STAGE == 0
    if (tid == 0) mbar_arrive(&s_mbar);                // every expect_tx precedes the one arrival (barrier above)
#endif
    // one word per chunk: piece k owns the chunks that START in [start_k, start_{k+1}) (piece 0 also those before it)
    for (int k = tid; k < max(ngen, 1); k += T2_THREADS) {
        const int a = k ? max(s_start[k], 0) : 0, nb = s_start[k + 1], b = min(nb, tile_len);
        const int e = s_end[k];
        const t2_src_t sb = s_base[k];
        const int u1 = min((b + 31) >> 5, T2_UNITS);
        // the chunk in which the next piece starts belongs to pass 2 (unless that piece starts exactly on a chunk)
        const int ub = (nb < tile_len && (nb & 31)) ? (nb >> 5) : -1;
        const int st = s_start[k];
        for (int u = (a + 31) >> 5; u < u1; u++)       // skip: left to pass 2, or no base of piece k inside the chunk
            s_gx[u] = (u == ub || (u << 5) >= e || (u << 5) + 32 <= st) ? T2_SKIP : sb + (u << 5);
        s_def[k] = (uint16_t)(ub >= 0 && (ub << 5) >= a ? ub : 0xFFFF);   // listed by the piece that owns the chunk's first byte
    }
#if T2_STAGE == 0
    __syncthreads();
    mbar_wait(&s_mbar, 0);                             // all staged bytes have landed (also on the fallback path: the buffer
                                                       // must not be released with copies in flight)
#else
    __syncthreads();
#endif
    const int nlit = s_nlit;
    t2_chunk_ctx ctx = {packed, piece_off, piece_src, n_piece, p_lo, P0, total, T, lit, exc_pos, exc_byte, n_exc, out};
    if (s_bad || nlit > T2_LITCAP) {                   // too many / too scattered pieces for the tables: generic path
        for (int cidx = 0; cidx < T2_CHUNKS; cidx++) {
            const int p = (cidx * T2_THREADS + tid) << 5;
            if (p >= tile_len) break;
            t2_chunk_exact(ctx, p);
        }
        return;
    }

    // pass 1
#pragma unroll 1
    for (int u = tid; (u << 5) < tile_len; u += T2_THREADS) {
        const t2_src_t gx = s_gx[u];
        if (gx == T2_SKIP) continue;
        uint32_t n[4];
        t2_window(s_raw, packed, gx, n);
        // code 15 = byte outside the packed alphabet on a '+' piece: the exact FASTA byte must come out (genome.py:606)
        if (n_exc > 0 && t2_has_code15(n)) { t2_chunk_exact(ctx, u << 5); continue; }
        uint32_t w[8];
        decode32(n, w);
        st32(out + P0 + (u << 5), w);
    }
    // pass 2: the chunk in which piece k+1 starts (further pieces may start in it too)
#pragma unroll 1
    for (int k = tid; k < ngen; k += T2_THREADS) {
        const int u = s_def[k];
        if (u == 0xFFFF) continue;
        const int p = u << 5;
        const int end = min(32, tile_len - p);
        uint32_t n[4] = {0, 0, 0, 0}, y[4];
        if (p < s_end[k]) t2_window(s_raw, packed, s_base[k] + p, n);
#pragma unroll 1
        for (int Z = k + 1; s_start[Z] - p < end; Z++) {
            const int cz = s_start[Z] - p;
            t2_window(s_raw, packed, s_base[Z] + p, y);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t keep = low_nibbles(cz - 8 * q);
                n[q] = (n[q] & keep) | (y[q] & ~keep);
            }
        }
        if (n_exc > 0 && t2_has_code15(n)) { t2_chunk_exact(ctx, p); continue; }
        uint32_t w[8];
        decode32(n, w);
        st32(out + P0 + p, w);
    }
    __syncthreads();

    // framing bytes over the placeholders (same CTA, ordered by the barrier; the lines are still in L2)
    for (int k = tid; k < nlit * 8; k += T2_THREADS) {
        const int slot = k >> 3;
        const int r1 = s_lb[slot];
        const uint8_t *src = lit + s_lsrc[slot];
        for (int q = s_la[slot] + (k & 7); q < r1; q += 32) {
            uint8_t b[4];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = q + 8 * j < r1 ? __ldg(src + q + 8 * j) : (uint8_t)0;
#pragma unroll
            for (int j = 0; j < 4; j++) if (q + 8 * j < r1) out[P0 + q + 8 * j] = b[j];
        }
    }
}

// ---- host API -------------------------------------------------------------------------------------------------
static int g_emit_mode = -1;
int mg_emit_mode() {
    if (g_emit_mode < 0) {
        const char *e = getenv("MAGOT_EMIT");
        g_emit_mode = (e && !strcmp(e, "tma")) ? 1 : (e && !strcmp(e, "stream")) ? 2 : 0;
    }
    return g_emit_mode;
}

// Developer / test knob: select a kernel variant at run time (the variants are bit-identical in their results; tests run all of
// them).  "emit": 0 = k_emit_nuc (default), 1 = k_emit_nuc_tma, 2 = k_emit_nuc_stream; "k1": 0 = piece-parallel launches (default), 1 = k_plan_rec.
void mg_set_k1_mode(int v);
void mg_set_six_mode(int v);
extern "C" int mg_tune(const char *key, int value) {
    MG_REQUIRE(key != nullptr, "key is NULL");
    if (!strcmp(key, "emit")) { MG_REQUIRE(value >= 0 && value <= 2, "emit variant must be 0, 1 or 2"); g_emit_mode = value; return MG_OK; }
    if (!strcmp(key, "k1")) { mg_set_k1_mode(value); return MG_OK; }
    if (!strcmp(key, "six")) { MG_REQUIRE(value >= 0 && value <= 2, "six-frame scan variant must be 0, 1 or 2"); mg_set_six_mode(value); return MG_OK; }
    if (!strcmp(key, "fuse")) { mg_set_fuse(value != 0); return MG_OK; }
    if (!strcmp(key, "multi_lag")) { MG_REQUIRE(value >= 0 && value <= 1000000, "multi_lag is in millionths of a text"); mg_set_multi_lag(value); return MG_OK; }
    mg_set_error("mg_tune: unknown key '%s'", key);
    return MG_EINVAL;
}

int mg_launch_nuc_tma(mg_plan *p, uint8_t *out_dev, cudaStream_t st) {
    mg_genome *g = p->g;
    k_emit_nuc_tma<<<(unsigned)p->n_nuc_tile, T2_THREADS, 0, st>>>(g->d_packed, p->d_piece_off, p->d_piece_src, p->n_piece, p->d_nuc_tile,
                                                                  p->d_totals, p->nuc_total, g->total_bases, p->d_lit, g->d_exc_pos,
                                                                  g->d_exc_byte, g->n_exc, out_dev);
    MG_LAUNCH_CHECK();
    return MG_OK;
}
