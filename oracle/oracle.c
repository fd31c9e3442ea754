/* TEST INFRASTRUCTURE -- plain C restatement of the byte-level functions of MAGOT's
 * annotation-driven sequence path (reference: /root/reference/genome.py).
 *
 * Used only as (a) the checker for the CUDA kernels at sizes the Python restatement
 * (oracle/magot_oracle.py) would take minutes for and (b) the single-threaded "port" leg of
 * bench.py's cpu_baseline.  Never linked into or called from the product library.
 * Validated against oracle/magot_oracle.py and the reference-generated known-answer vectors
 * (tests/golden/kat.json) by tests/test_oracle_golden.py.
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/liboracle.so oracle/oracle.c   (oracle/Makefile)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* genome.py:787 -- complement table; every byte not listed maps to lower-case 'n' (:791-792) */
static unsigned char comp_of(unsigned char c) {
    switch (c) {
    case 'a': return 't'; case 't': return 'a'; case 'g': return 'c'; case 'c': return 'g';
    case 'A': return 'T'; case 'T': return 'A'; case 'G': return 'C'; case 'C': return 'G';
    case 'n': return 'n'; case 'N': return 'N'; case '-': return '-';
    default:  return 'n';
    }
}

/* genome.py:784-793 Sequence.reverse_compliment */
void mo_revcomp(const unsigned char *in, int64_t n, unsigned char *out) {
    for (int64_t i = 0; i < n; i++) out[i] = comp_of(in[n - 1 - i]);
}

/* genome.py:795-802 -- 64-codon standard table, index = 16*b0+4*b1+b2 with T=0,C=1,A=2,G=3 */
static const char AAS[65] = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
static int base_idx(unsigned char c) {
    switch (c) {                       /* .upper() at genome.py:812 */
    case 'T': case 't': return 0; case 'C': case 'c': return 1;
    case 'A': case 'a': return 2; case 'G': case 'g': return 3;
    default: return -1;
    }
}
static unsigned char codon_aa(const unsigned char *t, int len) {
    if (len != 3) return 'X';          /* partial triplet is never a library key (:816-817) */
    int a = base_idx(t[0]), b = base_idx(t[1]), c = base_idx(t[2]);
    if (a < 0 || b < 0 || c < 0) return 'X';
    return (unsigned char)AAS[a * 16 + b * 4 + c];
}

/* genome.py:795-822 Sequence.translate on an already strand-corrected sequence `seq`.
 * Returns the number of amino acids written, or -1 for the reference's `None` (:810). */
int64_t mo_translate_fwd(const unsigned char *seq, int64_t n, int frame, int trimX, unsigned char *out) {
    if (!(n > 2 + frame)) return -1;
    int64_t m = 0;
    unsigned char trip[3];
    int tl = 0;
    for (int64_t pos = frame; pos < n; pos++) {          /* :811 */
        trip[tl++] = seq[pos];
        if ((pos + frame) % 3 == 2) {                    /* :813 -- the frame quirk lives here */
            out[m++] = codon_aa(trip, tl);
            tl = 0;
        }
    }
    if (trimX && m > 0 && out[0] == 'X') {               /* :819-821, exactly one X */
        memmove(out, out + 1, (size_t)(m - 1));
        m--;
    }
    return m;
}

/* translate with strand handling (:806-809); tmp must hold n bytes when strand == '-' */
int64_t mo_translate(const unsigned char *seq, int64_t n, int frame, int minus, int trimX,
                     unsigned char *tmp, unsigned char *out) {
    if (minus) {
        mo_revcomp(seq, n, tmp);
        return mo_translate_fwd(tmp, n, frame, trimX, out);
    }
    return mo_translate_fwd(seq, n, frame, trimX, out);
}

/* genome.py:677-710 -- splice the intervals of each record, in the emission order the caller
 * already established (sorted by coords, reversed on '-'), each interval reverse-complemented by
 * its own strand (:603-608).  Intervals are 0-based half-open [lo,hi) on contig `cid`, already
 * clamped like a Python slice.  Returns total bytes written. */
int64_t mo_splice(const unsigned char *const *contigs, int64_t n_rec, const int64_t *rec_off,
                  const int32_t *cid, const int64_t *lo, const int64_t *hi, const int8_t *minus,
                  unsigned char *out, int64_t *out_off) {
    int64_t w = 0;
    for (int64_t r = 0; r < n_rec; r++) {
        out_off[r] = w;
        for (int64_t k = rec_off[r]; k < rec_off[r + 1]; k++) {
            int64_t n = hi[k] - lo[k];
            if (n <= 0) continue;
            const unsigned char *src = contigs[cid[k]] + lo[k];
            if (minus[k]) mo_revcomp(src, n, out + w);
            else memcpy(out + w, src, (size_t)n);
            w += n;
        }
    }
    out_off[n_rec] = w;
    return w;
}

/* splice + translate(frame 0, '+', trimX) per record (:705-707).  aa_off[r+1]-aa_off[r] is the
 * protein length, or aa_len[r] = -1 where the reference would return None (spliced length <= 2). */
int64_t mo_splice_translate(const unsigned char *nuc, int64_t n_rec, const int64_t *nuc_off,
                            unsigned char *aa, int64_t *aa_off, int64_t *aa_len) {
    int64_t w = 0;
    for (int64_t r = 0; r < n_rec; r++) {
        aa_off[r] = w;
        int64_t m = mo_translate_fwd(nuc + nuc_off[r], nuc_off[r + 1] - nuc_off[r], 0, 1, aa + w);
        aa_len[r] = m;
        if (m > 0) w += m;
    }
    aa_off[n_rec] = w;
    return w;
}

/* genome.py:824-851 Sequence.get_orfs(longest=False, from_atg=False) with an additional
 * length filter (min_aa = 0 reproduces the reference list, empty strings included).
 * Order: frames 0,1,2; within a frame strand '-' then '+' (:829-830).
 * Emits each kept ORF into `aa` back to back; rec[4*i+0..3] = frame, minus, start index in that
 * frame's translated string, length.  Pass aa == NULL to count only.
 * Returns the number of kept ORFs; *n_bytes receives the total amino-acid bytes. */
int64_t mo_sixframe(const unsigned char *seq, int64_t n, int64_t min_aa,
                    unsigned char *aa, int64_t *rec, int64_t *n_bytes) {
    unsigned char *rc = (unsigned char *)malloc((size_t)(n > 0 ? n : 1));
    unsigned char *tr = (unsigned char *)malloc((size_t)(n / 3 + 4));
    int64_t n_orf = 0, w = 0;
    mo_revcomp(seq, n, rc);
    for (int frame = 0; frame < 3; frame++) {
        for (int s = 0; s < 2; s++) {
            int minus = (s == 0);
            int64_t m = mo_translate_fwd(minus ? rc : seq, n, frame, 1, tr);
            if (m <= 0) continue;                        /* `if translated_seq:` (:832) */
            int64_t start = 0;
            for (int64_t i = 0; i <= m; i++) {
                if (i == m || tr[i] == '*') {            /* str.split('*') (:833) */
                    int64_t len = i - start;
                    if (len >= min_aa) {
                        if (aa) {
                            memcpy(aa + w, tr + start, (size_t)len);
                            rec[4 * n_orf + 0] = frame; rec[4 * n_orf + 1] = minus;
                            rec[4 * n_orf + 2] = start; rec[4 * n_orf + 3] = len;
                        }
                        w += len;
                        n_orf++;
                    }
                    start = i + 1;
                }
            }
        }
    }
    free(rc); free(tr);
    *n_bytes = w;
    return n_orf;
}
