/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/bnp_oracle.py for the header that applies
 * here too).  Plain-C restatement of the reference path for sizes the NumPy oracle is too slow
 * for: FASTQ chunk bytes -> newline split (io/one_line_buffer.py:44-71) -> sequence lines
 * (io/file_buffers.py:315-338) -> LUT (encodings/alphabet_encoding.py:19-46) -> k-mer hash
 * h = sum_j code[i+j]*4^j (sequence/kmers.py:105-126) -> [window minimum, sequence/minimizers.py:15-17]
 * -> histogram of (value mod n_bins) (sequence/count_encoded.py:173-177; bucketed = extension).
 * Pinned by tests/test_oracle_goldens.py against the NumPy oracle, which is itself pinned to the
 * reference's golden values.  Scalar, single thread.
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC)
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

static void make_lut(uint8_t lut[256], const char *alphabet) {
    memset(lut, 255, 256);
    for (int i = 0; alphabet[i]; ++i) {
        unsigned char c = (unsigned char)alphabet[i];
        lut[c] = (uint8_t)i;
        lut[c + 32] = (uint8_t)i; /* lower case, alphabet_encoding.py:24-28 */
    }
}

/* Returns the number of complete records, or -1 - (flat offset) of the first invalid base.
 * out_stats: [0] n_complete_bytes, [1] n_bases, [2] n_values. */
int64_t oracle_fastq_kmer_hist(const uint8_t *chunk, size_t n, int lines_per_entry, const char *alphabet,
                               int k, int window, uint64_t n_bins, int64_t *hist, int64_t *out_stats) {
    uint8_t lut[256];
    make_lut(lut, alphabet);
    /* pass 1: count newlines, keep a multiple of lines_per_entry (one_line_buffer.py:63-69) */
    size_t n_lines = 0;
    for (size_t i = 0; i < n; ++i) n_lines += (chunk[i] == '\n');
    const size_t keep = n_lines - n_lines % (size_t)lines_per_entry;
    const uint64_t kmask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    const int w = window ? window - k + 1 : 1;
    size_t line = 0, line_start = 0, complete_bytes = 0;
    int64_t n_bases = 0, n_values = 0, flat = 0;
    uint64_t ring[1024];
    for (size_t i = 0; i < n && line < keep; ++i) {
        if (chunk[i] != '\n') continue;
        if (line % (size_t)lines_per_entry == 1) { /* the sequence line */
            size_t end = i;
            if (end > line_start && chunk[end - 1] == '\r') end--; /* see cr note in tests */
            const size_t L = end - line_start;
            uint64_t h = 0;
            for (size_t p = 0; p < L; ++p) {
                const uint8_t c = lut[chunk[line_start + p]];
                if (c >= 4) return -1 - (flat + (int64_t)p);
                h = (h >> 2) | ((uint64_t)c << (2 * (k - 1)));
                h &= kmask;
                if (p + 1 >= (size_t)k) {
                    const size_t idx = p + 1 - (size_t)k; /* k-mer index in the row */
                    if (!window) {
                        hist[h % n_bins]++;
                        n_values++;
                    } else {
                        ring[idx % (size_t)w] = h;
                        if (idx + 1 >= (size_t)w) {
                            uint64_t m = ring[0];
                            for (int q = 1; q < w; ++q) if (ring[q] < m) m = ring[q];
                            hist[m % n_bins]++;
                            n_values++;
                        }
                    }
                }
            }
            n_bases += (int64_t)L;
            flat += (int64_t)L;
        }
        line_start = i + 1;
        line++;
        if (line % (size_t)lines_per_entry == 0) complete_bytes = i + 1;
    }
    if (out_stats) { out_stats[0] = (int64_t)complete_bytes; out_stats[1] = n_bases; out_stats[2] = n_values; }
    return (int64_t)(keep / (size_t)lines_per_entry);
}

/* flat hashes of one already-encoded row (codes 0..3): out[i] = sum_j code[i+j]*4^j */
void oracle_row_hashes(const uint8_t *codes, size_t L, int k, int64_t *out) {
    const uint64_t kmask = (1ull << (2 * k)) - 1;
    uint64_t h = 0;
    for (size_t p = 0; p < L; ++p) {
        h = ((h >> 2) | ((uint64_t)codes[p] << (2 * (k - 1)))) & kmask;
        if (p + 1 >= (size_t)k) out[p + 1 - (size_t)k] = (int64_t)h;
    }
}
