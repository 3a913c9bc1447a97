// hc_host.cpp -- see hc_host.hpp.  Thin: packs host vectors, calls hc_compress_batch /
// hc_decompress_batch, maps per-file status codes to the reference's error behaviour.
#include "hc_host.hpp"

#include <cstring>

#include "hc_b200.h"

namespace hc {

const char *statusMessage(int s)
{
    switch (s) {
    case 6: return "invalid size of input 2D data detected";          // src/main.cpp:56
    case 8: return "invalid or missing Huffman coding header";        // src/main.cpp:101
    case 9: return "invalid Huffman coding file contents";            // src/transform.cpp:396
    case 10: return "invalid or missing adaptive block RLE header";   // src/headers.cpp:69
    case 11: return "invalid adaptive block RLE header";              // src/headers.cpp:96
    case 12: return "too small 2D data dimensions";                   // src/transform.cpp:302
    case 13: return "invalid adaptive block RLE file contents";       // src/transform.cpp:182
    case 14: return "unexpected end of adaptive block RLE data";      // src/transform.cpp:172
    case 15: return "leftover data of adaptive block RLE detected";   // src/transform.cpp:356
    default: return hc_error_string(s);
    }
}

static void throwIf(int rc, const char *what)
{
    if (rc != 0) throw CodecError(rc < 0 ? rc : -1000 - rc, std::string(what) + ": " + hc_error_string(rc));
}

Codec::Codec(int device) : h_(nullptr)
{
    hc_codec *c = nullptr;
    throwIf(hc_codec_create(&c, device), "no usable CUDA device (this build has no CPU path)");
    h_ = c;
}

Codec::~Codec() { hc_codec_destroy((hc_codec *)h_); }

namespace {
struct Packed {
    std::vector<uint8_t> buf;
    std::vector<uint64_t> off, len;
};
Packed pack(const std::vector<std::vector<uint8_t>> &files)
{
    Packed p;
    uint64_t pos = 0;
    for (auto &f : files) {
        p.off.push_back(pos);
        p.len.push_back(f.size());
        pos += (f.size() + 15) / 16 * 16;
    }
    p.buf.assign(pos ? pos : 1, 0);
    for (size_t i = 0; i < files.size(); i++)
        if (!files[i].empty()) memcpy(p.buf.data() + p.off[i], files[i].data(), files[i].size());
    return p;
}
std::vector<std::vector<uint8_t>> unpack(const std::vector<uint8_t> &out, const std::vector<uint64_t> &off,
                                         const std::vector<uint64_t> &len, const std::vector<int32_t> &st)
{
    std::vector<std::vector<uint8_t>> r(off.size());
    for (size_t i = 0; i < off.size(); i++)
        if (st[i] == 0) r[i].assign(out.begin() + off[i], out.begin() + off[i] + len[i]);
    return r;
}
}  // namespace

std::vector<std::vector<uint8_t>> Codec::huffCompressBatch(const std::vector<std::vector<uint8_t>> &files, bool useDiffModel,
                                                            bool useAdaptRLE, uint64_t matrixWidth, std::vector<int> &status)
{
    const uint32_t nf = (uint32_t)files.size();
    Packed p = pack(files);
    uint64_t cap = 0;
    for (auto n : p.len) {
        uint64_t m = useAdaptRLE ? n + n / 3 + n / 8 + 1024 : hc_rle_bound(n);
        cap += (hc_fgk_bound(m) + 31) / 16 * 16;
    }
    std::vector<uint8_t> out(cap ? cap : 16);
    std::vector<uint64_t> off(nf), len(nf), width(nf, matrixWidth);
    std::vector<int32_t> st(nf);
    throwIf(hc_compress_batch((hc_codec *)h_, p.buf.data(), p.off.data(), p.len.data(), nf, useDiffModel, useAdaptRLE, width.data(),
                              out.data(), out.size(), off.data(), len.data(), st.data()),
            "hc_compress_batch");
    status.assign(st.begin(), st.end());
    return unpack(out, off, len, st);
}

std::vector<std::vector<uint8_t>> Codec::huffDecompressBatch(const std::vector<std::vector<uint8_t>> &files, std::vector<int> &status)
{
    const uint32_t nf = (uint32_t)files.size();
    Packed p = pack(files);
    // First try with a generous guess; a file that does not fit comes back with HC_E_CAPACITY and the size it
    // needs in len[f], so ONE second call with the exact total settles it.  Hard ceiling: an MNP-5 token
    // expands to at most 255 bytes, anything beyond 255 x input is reported as that file's error.
    uint64_t in_total = 0;
    for (auto n : p.len) in_total += n;
    uint64_t cap = (1 << 20) + 8 * in_total;
    const uint64_t ceiling = 255 * in_total + 16 * (uint64_t)nf + 4096;
    for (int attempt = 0;; attempt++) {
        std::vector<uint8_t> out(cap);
        std::vector<uint64_t> off(nf), len(nf);
        std::vector<int32_t> st(nf);
        int rc = hc_decompress_batch((hc_codec *)h_, p.buf.data(), p.off.data(), p.len.data(), nf, out.data(), out.size(), off.data(),
                                     len.data(), st.data());
        throwIf(rc, "hc_decompress_batch");
        uint64_t need = 0;
        bool retry = false;
        for (uint32_t f = 0; f < nf; f++) {
            if (st[f] == HC_E_CAPACITY) retry = true;
            if (st[f] == 0 || st[f] == HC_E_CAPACITY) need += (len[f] + 15) / 16 * 16;
        }
        if (retry && attempt == 0 && need <= ceiling) { cap = need + 4096; continue; }
        status.assign(st.begin(), st.end());
        return unpack(out, off, len, st);
    }
}

std::vector<uint8_t> Codec::huffCompress(const std::vector<uint8_t> &inData, bool useDiffModel, bool useAdaptRLE, uint64_t matrixWidth)
{
    std::vector<int> st;
    auto r = huffCompressBatch({inData}, useDiffModel, useAdaptRLE, matrixWidth, st);
    if (st[0] != 0) throw CodecError(st[0], statusMessage(st[0]));
    return r[0];
}

std::vector<uint8_t> Codec::huffDecompress(const std::vector<uint8_t> &inData)
{
    std::vector<int> st;
    auto r = huffDecompressBatch({inData}, st);
    if (st[0] != 0) throw CodecError(st[0], statusMessage(st[0]));
    return r[0];
}

}  // namespace hc
