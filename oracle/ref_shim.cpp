// ref_shim.cpp -- extern "C" access to the UNMODIFIED reference stage functions.
//
// TEST INFRASTRUCTURE ONLY (see hc_oracle.h).  This file is ours; it is compiled
// together with the reference's own sources where they lie (/root/reference/src,
// never copied into this repo) into oracle/_ref/libhcref.so by oracle/Makefile.
// The reference reports errors with exit(n); callers that probe error paths run
// the oracle/_ref/huffman-codec binary in a subprocess instead.
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <deque>
#include <vector>

#include "transform.hpp"   // -I/root/reference/src
#include "headers.hpp"
#include "huffman.hpp"

// external linkage in src/transform.cpp:97 but not declared in transform.hpp
std::vector<uint8_t> applyAdaptRLE(const std::vector<uint8_t> &matrix, uint64_t matrixWidth,
                                   uint64_t matrixHeight, uint64_t blockSize);

static uint8_t *dup_out(const std::vector<uint8_t> &v, size_t *n)
{
    uint8_t *p = (uint8_t *)malloc(v.size() ? v.size() : 1);
    if (!v.empty()) memcpy(p, v.data(), v.size());
    *n = v.size();
    return p;
}

extern "C" {

void ref_free(void *p) { free(p); }

void ref_diff_apply(uint8_t *v, size_t n)
{
    std::vector<uint8_t> x(v, v + n);
    applyDiffModel(x);
    if (n) memcpy(v, x.data(), n);
}

void ref_diff_revert(uint8_t *v, size_t n)
{
    std::vector<uint8_t> x(v, v + n);
    revertDiffModel(x);
    if (n) memcpy(v, x.data(), n);
}

uint8_t *ref_rle_encode(const uint8_t *in, size_t n, size_t *m)
{
    return dup_out(applyRLE(std::vector<uint8_t>(in, in + n)), m);
}

uint8_t *ref_rle_decode(const uint8_t *in, size_t m, size_t *n)
{
    return dup_out(revertRLE(std::deque<uint8_t>(in, in + m)), n);
}

uint8_t *ref_adapt_encode(const uint8_t *in, uint64_t w, uint64_t h, size_t *m)
{
    return dup_out(applyAdaptRLE(std::vector<uint8_t>(in, in + w * h), w, h), m);
}

uint8_t *ref_adapt_encode_bs(const uint8_t *in, uint64_t w, uint64_t h, uint64_t b, size_t *m)
{
    return dup_out(applyAdaptRLE(std::vector<uint8_t>(in, in + w * h), w, h, b), m);
}

uint8_t *ref_adapt_decode(const uint8_t *in, size_t m, size_t *n)
{
    std::deque<uint8_t> d(in, in + m);
    return dup_out(revertAdaptRLE(d), n);
}

// packed MSB-first like src/main.cpp:78-84; *nbits is the padded bit count
uint8_t *ref_fgk_encode(const uint8_t *sym, size_t m, size_t *nbytes)
{
    std::vector<bool> bits = applyHuffman(std::vector<uint8_t>(sym, sym + m));
    std::vector<uint8_t> out;
    for (size_t i = 0; i < bits.size(); i += 8) {
        uint8_t b = 0;
        for (int j = 0; j < 8; j++) b = (uint8_t)((b << 1) | bits[i + j]);
        out.push_back(b);
    }
    return dup_out(out, nbytes);
}

void ref_fgk_decode(const uint8_t *bytes, size_t nbytes, uint64_t count, uint8_t *sym_out)
{
    std::deque<bool> bits;
    for (size_t i = 0; i < nbytes; i++)
        for (int k = 8; k > 0; k--) bits.push_back((bytes[i] >> (k - 1)) & 1);
    std::deque<uint8_t> r = revertHuffman(bits, count);
    for (size_t i = 0; i < r.size(); i++) sym_out[i] = r[i];
}

uint64_t ref_block_count(uint64_t w, uint64_t h, uint64_t b) { return getBlockCount(w, h, b); }

}  // extern "C"
