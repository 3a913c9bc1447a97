// hc_host.hpp -- C++ host mirror of the reference's in-process interface for the hot path,
// implemented on top of the C ABI (include/hc_b200.h).  Same names and argument meaning as
// src/main.cpp:39-128 (huffCompress / huffDecompress); where the reference prints an error and
// calls exit(n), these throw hc::CodecError carrying the same code n and message, so a caller
// (the CLI in main.cpp) reproduces the reference's stderr text and exit status exactly.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace hc {

struct CodecError : std::runtime_error {
    int code;                       // reference exit code or <0 for CUDA failures
    CodecError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

// the reference's message for a per-file status (without the "ERROR: " prefix / newline)
const char *statusMessage(int status);

class Codec {
public:
    explicit Codec(int device = 0);
    ~Codec();
    Codec(const Codec &) = delete;
    Codec &operator=(const Codec &) = delete;

    // huffCompress, src/main.cpp:39-87 (the file is already in memory)
    std::vector<uint8_t> huffCompress(const std::vector<uint8_t> &inData, bool useDiffModel, bool useAdaptRLE,
                                      uint64_t matrixWidth);
    // huffDecompress, src/main.cpp:90-128
    std::vector<uint8_t> huffDecompress(const std::vector<uint8_t> &inData);

    // batch forms: one call, many files; status[i] holds the reference exit code of file i
    std::vector<std::vector<uint8_t>> huffCompressBatch(const std::vector<std::vector<uint8_t>> &files, bool useDiffModel,
                                                        bool useAdaptRLE, uint64_t matrixWidth, std::vector<int> &status);
    std::vector<std::vector<uint8_t>> huffDecompressBatch(const std::vector<std::vector<uint8_t>> &files,
                                                          std::vector<int> &status);

private:
    void *h_;
};

}  // namespace hc
