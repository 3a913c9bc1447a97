/*
 * hc_oracle.h -- CPU ORACLE for the huffman-codec compression pipeline.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference
 * algorithms (dominiksalvet/huffman-codec, src/transform.cpp, src/headers.cpp,
 * src/huffman.cpp, src/main.cpp).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg may load it; the product
 * library (libhc_b200.so) never links or calls anything in oracle/.
 *
 * Parity pin: the restatement is checked (tests/test_oracle.py) against
 *   (1) the golden vectors in tests/golden/ that were generated from the
 *       unmodified reference binary (tests/golden/make_golden.py),
 *   (2) the hand-checkable vectors of SURVEY.md Appendix A.7,
 *   (3) live differential fuzzing against oracle/_ref/libhcref.so (the
 *       reference's own sources compiled where they lie) when present.
 *
 * All functions return 0 on success or the reference's exit code (6, 8..15)
 * where the reference would have printed an error and called exit().
 * Buffers returned through `uint8_t **out` are malloc()ed; free with hco_free.
 */
#ifndef HC_ORACLE_H
#define HC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

void hco_free(void *p);

/* src/transform.cpp:220-229 / :231-239 (in situ) */
void hco_diff_apply(uint8_t *v, size_t n);
void hco_diff_revert(uint8_t *v, size_t n);

/* src/transform.cpp:241-279; out must hold hco_rle_bound(n) bytes; returns m */
size_t hco_rle_bound(size_t n);
size_t hco_rle_encode(const uint8_t *in, size_t n, uint8_t *out);
/* src/transform.cpp:137-159, :281-292 */
int hco_rle_decode(const uint8_t *in, size_t m, uint8_t **out, size_t *n);

/* src/transform.cpp:410-418 */
uint64_t hco_block_count(uint64_t w, uint64_t h, uint64_t b);
/* src/transform.cpp:97-134 (one fixed block size), header per src/headers.cpp:18-63 */
int hco_adapt_encode_bs(const uint8_t *in, uint64_t w, uint64_t h, uint64_t b,
                        uint8_t **out, size_t *m);
/* src/transform.cpp:294-328 (exhaustive block-size search); 12 if w<8||h<8 */
int hco_adapt_encode(const uint8_t *in, uint64_t w, uint64_t h,
                     uint8_t **out, size_t *m, uint64_t *chosen_b);
/* src/transform.cpp:330-361 + src/headers.cpp:65-105; errors 10,11,13,14,15 */
int hco_adapt_decode(const uint8_t *in, size_t m, uint8_t **out, size_t *n);

/* src/transform.cpp:363-384: returns malloc()ed packed bits (MSB first, zero
 * padded to a byte, src/main.cpp:78-84); *nbytes = ceil(bits/8), *nbits = raw bit count.
 * mode 0 = faithful pointer-tree restatement of src/huffman.cpp (recursive
 * findSuccNode); mode 1 = number-indexed array form (same output, much faster). */
int hco_fgk_encode(const uint8_t *sym, size_t m, int mode,
                   uint8_t **out, size_t *nbytes, uint64_t *nbits);
/* src/transform.cpp:386-406: decode `count` symbols; 9 on bit underrun */
int hco_fgk_decode(const uint8_t *bytes, size_t nbytes, uint64_t count, int mode,
                   uint8_t *sym_out);

/* src/main.cpp:39-87 (huffCompress) and :90-128 (huffDecompress), whole .out files */
int hco_compress(const uint8_t *in, size_t n, int diff, int adapt, uint64_t width,
                 int mode, uint8_t **out, size_t *outlen);
int hco_decompress(const uint8_t *in, size_t n, int mode, uint8_t **out, size_t *outlen);

/* statistics helper for DESIGN/bench: tree levels walked + swaps while encoding */
int hco_fgk_stats(const uint8_t *sym, size_t m, uint64_t *levels, uint64_t *swaps,
                  uint32_t *max_depth);

#ifdef __cplusplus
}
#endif
#endif
