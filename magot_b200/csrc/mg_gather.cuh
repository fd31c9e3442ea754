// mg_gather.cuh -- spliced-coordinate gather: the device form of `"".join(seq_list)` (genome.py:705)
// with each child already reverse-complemented by its own strand (genome.py:603-608).
#pragma once
#include "mg_common.cuh"

#ifdef __CUDACC__
// Gather nb (<= 48) nibbles of the spliced sequence starting at nucleotide-text offset S, which lies in
// piece j (piece_off[j] <= S < piece_off[j+1]).  Following pieces are walked as needed; the caller
// guarantees that [S, S+nb) stays inside genome-segment pieces.  Result: nibble k in bits [4k,4k+4)
// of the 192-bit value acc[2]:acc[1]:acc[0]; reverse-strand pieces are forward reads of the genome's
// reverse-complement plane, so the consumer never needs to know the strand.
__device__ __forceinline__ void mg_gather_nib(const uint32_t *__restrict__ packed, const int64_t *__restrict__ piece_off,
                                              const int64_t *__restrict__ piece_src, int64_t j, int64_t S, int nb,
                                              uint64_t acc[3]) {
    acc[0] = acc[1] = acc[2] = 0;
    int f = 0;
    int64_t off_j = __ldg(piece_off + j), off_n = __ldg(piece_off + j + 1);
    while (f < nb) {
        while (off_n <= S) {                      // skip empty pieces (clamped-away segments)
            j++;
            off_j = off_n;
            off_n = __ldg(piece_off + j + 1);
        }
        const uint64_t sk = (uint64_t)__ldg(piece_src + j);
        const int64_t src = (int64_t)(sk & MG_SRC_MASK);
        const int64_t o = S - off_j;
        const int64_t rem = off_n - S;
        int c = nb - f;
        if (c > 16) c = 16;
        if (rem < c) c = (int)rem;
        uint64_t v = mg_ld_nib16(packed, src + o);     // '-' pieces already point into the reverse plane
        if (c < 16) v &= (1ull << (4 * c)) - 1ull;
        const int w = f >> 4, sh = (f & 15) << 2;
        if (w == 0) {
            acc[0] |= v << sh;
            if (sh) acc[1] |= v >> (64 - sh);
        } else if (w == 1) {
            acc[1] |= v << sh;
            if (sh) acc[2] |= v >> (64 - sh);
        } else {
            acc[2] |= v << sh;
        }
        f += c;
        S += c;
    }
}
#endif
