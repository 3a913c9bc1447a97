// hc_api.cu -- C ABI of libhc_b200.so (see include/hc_b200.h): stage launchers, the batched
// host pipeline (huffCompress / huffDecompress of src/main.cpp:39-128) and small utility kernels.
//
// Product build: nvcc -gencode arch=compute_100a,code=sm_100a -> libhc_b200.so.  There is no CPU
// implementation behind these entry points.  (-DHC_EMU builds the SIMT-emulated test double used
// by tests/emu only; see tests/emu/hc_emu.h.)
#include "../../include/hc_b200.h"

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#ifndef HC_EMU
#include <dlfcn.h>
#endif
#include <cstdlib>
#include <cstring>
#include <vector>

#ifdef HC_EMU
#include "hc_emu.h"
#include "hc_emu_cuda.h"
#endif
#include "hc_common.cuh"
#include "scan.cuh"
#include "diff.cuh"
#include "rle.cuh"
#include "adapt.cuh"
#include "adapt_mask.cuh"
#include "adapt_small.cuh"
#include "adapt_large.cuh"
#include "fgk.cuh"

static std::atomic<uint64_t> g_launches{0};
void hc_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

#define HC_CUDA(x)                                         \
    do {                                                   \
        cudaError_t e_ = (x);                              \
        if (e_ != cudaSuccess) return -(int)e_;            \
    } while (0)
#define HC_CHECK_LAUNCH() HC_CUDA(cudaGetLastError())
#define HC_TRY(x)                  \
    do {                           \
        int r_ = (x);              \
        if (r_ != 0) return r_;    \
    } while (0)
#ifndef HC_ADAPT_GROUP_DEFAULT_MB
#define HC_ADAPT_GROUP_DEFAULT_MB 0
#endif

namespace hcd {

// ---------------------------------------------------------------- utility kernels
HC_KERNEL offsets_scan_kernel(const u64 *HC_RESTRICT len, u64 *HC_RESTRICT off, u64 *HC_RESTRICT total,
                              u32 nf, u32 align)
{
    HC_SHARED u64 wsum[NW];
    const u32 tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    u64 carry = 0;
    for (u32 i0 = 0; i0 < nf; i0 += TPB) {
        u32 i = i0 + tid;
        u64 v = i < nf ? (len[i] + align - 1) / align * align : 0;
        u64 inc = v;
        for (int d = 1; d < 32; d <<= 1) {
            u64 x = shfl_up64(inc, d);
            if (lane >= (u32)d) inc += x;
        }
        if (lane == 31) wsum[wid] = inc;
        syncthreads();
        u64 wbase = 0, tot = 0;
        for (u32 j = 0; j < (u32)NW; j++) { if (j < wid) wbase += wsum[j]; tot += wsum[j]; }
        if (i < nf) off[i] = carry + wbase + inc - v;
        carry += tot;
        syncthreads();
    }
    if (tid == 0 && total) *total = carry;
}

HC_KERNEL HC_LAUNCH_BOUNDS(256, 4)
gather_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT len,
              u8 *HC_RESTRICT out, const u64 *HC_RESTRICT out_off, u32 nf)
{
    const u32 tid = threadIdx.x;
    for (u32 f = blockIdx.y; f < nf; f += gridDim.y) {
        const u64 n = len[f];
        const u8 *src = in + in_off[f];
        u8 *dst = out + out_off[f];
        for (u64 p = ((u64)blockIdx.x * TPB + tid) * 16; p < n; p += (u64)gridDim.x * TPB * 16) {
            uint4 v = ldg16(src + p);
            if (p + 16 > n) v = mask_tail(v, (u32)(n - p));
            stg16(dst + p, v);
        }
    }
}

// off[f] = f * stride
HC_KERNEL strided_offsets_kernel(u64 *off, u64 *cap, u64 stride, u32 nf)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nf) { off[i] = (u64)i * stride; if (cap) cap[i] = stride; }
}

// compress prep: height = len / width, status 6 when -a and len % width != 0 (src/main.cpp:54-59)
HC_KERNEL compress_prep_kernel(const u64 *HC_RESTRICT len, const u64 *HC_RESTRICT width, u64 *HC_RESTRICT w_eff,
                               u64 *HC_RESTRICT h_eff, u64 *HC_RESTRICT len_eff, u8 *HC_RESTRICT flags,
                               i32 *HC_RESTRICT st, u32 nf, int use_diff, int use_adapt)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    u64 w = width ? width[i] : 512, n = len[i];
    i32 s = 0;
    u64 h = 0;
    if (use_adapt) {
        if (w == 0 || n % w != 0) { s = 6; w = 0; }
        else h = n / w;
    }
    // the transform kernels keep per-file positions in 32 bits: a file of 2 GiB or more is refused, not mangled
    if (!s && n >= (1ull << 31)) s = 100;
    w_eff[i] = w; h_eff[i] = h;
    len_eff[i] = s ? 0 : n;
    flags[i] = (u8)(((use_diff ? 1 : 0) << 7) | ((use_adapt ? 1 : 0) << 6));
    st[i] = s;
}

// sym_len of files that failed an earlier stage is forced to 0 before FGK
HC_KERNEL mask_len_kernel(u64 *HC_RESTRICT len, const i32 *HC_RESTRICT st_a, const i32 *HC_RESTRICT st_b, u32 nf)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nf && ((st_a && st_a[i]) || (st_b && st_b[i]))) len[i] = 0;
}

HC_KERNEL merge_status_kernel(const i32 *HC_RESTRICT a, const i32 *HC_RESTRICT b, const i32 *HC_RESTRICT c,
                              i32 *HC_RESTRICT st, u64 *HC_RESTRICT out_len, u32 nf)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    i32 s = (a && a[i]) ? a[i] : ((b && b[i]) ? b[i] : (c ? c[i] : 0));
    st[i] = s;
    if (s && s != 100 && out_len) out_len[i] = 0;
}

// decompress: split the FGK-decoded streams by kind (header flag bit 6 / bit 7)
HC_KERNEL decompress_split_kernel(const u64 *HC_RESTRICT sym_len, const u8 *HC_RESTRICT flags,
                                  const i32 *HC_RESTRICT st_fgk, u64 *HC_RESTRICT len_plain,
                                  u64 *HC_RESTRICT len_adapt, u32 nf)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    bool ok = st_fgk[i] == 0, adapt = (flags[i] >> 6) & 1;
    len_plain[i] = (ok && !adapt) ? sym_len[i] : 0;
    len_adapt[i] = (ok && adapt) ? sym_len[i] : 0;
}

HC_KERNEL decompress_merge_kernel(const u8 *HC_RESTRICT flags, const i32 *HC_RESTRICT st_fgk,
                                  const u64 *HC_RESTRICT n_plain, const i32 *HC_RESTRICT st_plain,
                                  const u64 *HC_RESTRICT n_adapt, const i32 *HC_RESTRICT st_adapt,
                                  u64 *HC_RESTRICT out_len, u64 *HC_RESTRICT len_diff, i32 *HC_RESTRICT st, u32 nf)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    bool adapt = (flags[i] >> 6) & 1, diff = (flags[i] >> 7) & 1;
    i32 s = st_fgk[i];
    u64 n = 0;
    if (!s) {
        s = adapt ? st_adapt[i] : st_plain[i];
        n = adapt ? n_adapt[i] : n_plain[i];
    }
    if (s && s != 100) n = 0;
    out_len[i] = n;
    st[i] = s;
    if (len_diff) len_diff[i] = (!s && diff) ? n : 0;
}

static inline dim3 grid2(u64 x, u32 nf)
{
    if (x == 0) x = 1;
    if (x > 0x7fffffffull) x = 0x7fffffffull;
    return dim3((unsigned)x, nf > 65535u ? 65535u : (nf ? nf : 1u));
}

}  // namespace hcd

using namespace hcd;

// ======================================================================= misc
#ifdef HC_EMU
extern "C" const char *hc_version(void) { return "hc_b200 0.2 (SIMT emulator, tests only)"; }
#else
extern "C" const char *hc_version(void) { return "hc_b200 0.2 (sm_100a)"; }
#endif

extern "C" int hc_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return -(int)e;
    return n;
}

extern "C" const char *hc_error_string(int code)
{
    if (code < 0) return cudaGetErrorString((cudaError_t)(-code));
    switch (code) {
    case 0: return "ok";
    case 6: return "invalid size of input 2D data detected";
    case 8: return "invalid or missing Huffman coding header";
    case 9: return "invalid Huffman coding file contents";
    case 10: return "invalid or missing adaptive block RLE header";
    case 11: return "invalid adaptive block RLE header";
    case 12: return "too small 2D data dimensions";
    case 13: return "invalid adaptive block RLE file contents";
    case 14: return "unexpected end of adaptive block RLE data";
    case 15: return "leftover data of adaptive block RLE detected";
    case 100: return "output capacity too small";
    case 101: return "FGK code longer than 56 bits";
    case 102: return "internal error: inconsistent FGK tree";
    default: return "unknown";
    }
}

extern "C" uint64_t hc_launch_count(void) { return g_launches.load(); }

extern "C" uint64_t hc_rle_bound(uint64_t n) { return n + n / 3 + 4; }
extern "C" uint64_t hc_block_count(uint64_t w, uint64_t h, uint64_t b)
{
    if (b == 0) return 0;
    return (w / b + (w % b != 0)) * (h / b + (h % b != 0));
}
extern "C" uint64_t hc_adapt_bound(uint64_t w, uint64_t h)
{
    // header + direction bytes of the finest tiling + 4/3 expansion + one forced literal per block
    uint64_t nb = hc_block_count(w, h, 8);
    return 24 + (nb + 7) / 8 + hc_rle_bound(w * h) + 2 * nb;
}
extern "C" uint64_t hc_fgk_bound(uint64_t m)
{
    // FGK emits at most 2S + m bits (S = optimal static code length <= 8m) plus 256 escapes
    return 9 + (17 * m + 7) / 8 + 8 * 1024;
}

// ======================================================================= stage level
extern "C" int hc_diff_apply_batch(const uint8_t *in, const uint64_t *in_off, uint8_t *out, const uint64_t *out_off,
                                   const uint64_t *len, uint32_t nf, uint64_t max_len, hc_stream_t stream)
{
    if (nf == 0) return 0;
    u64 tiles = (max_len + TILE_BYTES - 1) / TILE_BYTES;
    // enough CTAs for two full waves (148 SMs x 8 resident), each streaming several tiles of its file
    u64 gx = (2368 + nf - 1) / nf;
    if (gx > tiles) gx = tiles;
    HC_LAUNCH(diff_apply_kernel, grid2(gx, nf), dim3(TPB), 0, stream, in, in_off, out, out_off, len, nf);
    HC_CHECK_LAUNCH();
    return 0;
}

// scratch of the segmented scan: one u32 per (file, segment); 0 when the batch needs no segments
static inline u64 diff_revert_ws_bytes(u32 nf, u64 max_len)
{
    const SegPlan p = plan_segments(nf, max_len);
    return p.nseg > 1 ? (u64)nf * p.nseg * sizeof(u32) : 0;
}

// ws: caller-owned scratch of diff_revert_ws_bytes() bytes, or NULL (stream-ordered allocation)
static int diff_revert_launch(const uint8_t *in, const uint64_t *in_off, uint8_t *out, const uint64_t *out_off,
                              const uint64_t *len, uint32_t nf, uint64_t max_len, void *ws, hc_stream_t stream)
{
    if (nf == 0) return 0;
    SegPlan p = plan_segments(nf, max_len);
    u32 *segsum = nullptr;
    bool own = false;
    if (p.nseg > 1) {
        segsum = (u32 *)ws;
        if (!segsum) {
            HC_CUDA(cudaMallocAsync((void **)&segsum, (size_t)nf * p.nseg * sizeof(u32), (cudaStream_t)stream));
            own = true;
        }
        HC_LAUNCH(diff_segsum_kernel, grid2(p.nseg, nf), dim3(TPB), 0, stream, in, in_off, len, nf, p.nseg,
                  p.seg_bytes, segsum);
        HC_CHECK_LAUNCH();
    }
    HC_LAUNCH(diff_revert_kernel, grid2(p.nseg, nf), dim3(TPB), 0, stream, in, in_off, out, out_off, len, nf,
              p.nseg, p.seg_bytes, (const u32 *)segsum);
    HC_CHECK_LAUNCH();
    if (own) HC_CUDA(cudaFreeAsync(segsum, (cudaStream_t)stream));
    return 0;
}

extern "C" int hc_diff_revert_batch(const uint8_t *in, const uint64_t *in_off, uint8_t *out, const uint64_t *out_off,
                                    const uint64_t *len, uint32_t nf, uint64_t max_len, hc_stream_t stream)
{
    return diff_revert_launch(in, in_off, out, out_off, len, nf, max_len, nullptr, stream);
}

static inline unsigned file_grid(uint32_t nf) { return nf > 0x7fffffffu ? 0x7fffffffu : (nf ? nf : 1u); }

extern "C" int hc_rle_encode_batch(const uint8_t *in, const uint64_t *in_off, const uint64_t *in_len,
                                   uint8_t *out, const uint64_t *out_off, uint64_t *out_len,
                                   uint32_t nf, uint64_t max_len, hc_stream_t stream)
{
    (void)max_len;
    if (nf == 0) return 0;
    HC_LAUNCH(rle_encode_kernel, dim3(file_grid(nf)), dim3(RTPB), 0, stream, in, in_off, in_len, out, out_off,
              out_len, nf);
    HC_CHECK_LAUNCH();
    return 0;
}

extern "C" int hc_rle_decode_batch(const uint8_t *in, const uint64_t *in_off, const uint64_t *in_len,
                                   uint8_t *out, const uint64_t *out_off, const uint64_t *out_cap,
                                   uint64_t *out_len, int32_t *status,
                                   uint32_t nf, uint64_t max_len, hc_stream_t stream)
{
    (void)max_len;
    if (nf == 0) return 0;
    HC_LAUNCH(rle_decode_kernel, dim3(file_grid(nf)), dim3(RTPB), 0, stream, in, in_off, in_len, out, out_off,
              out_cap, out_len, status, nf);
    HC_CHECK_LAUNCH();
    return 0;
}

// scratch layout of the adaptive encoder: block streams of the large-block path (per file) | cost
// tables | block offsets | chosen_b
static inline u64 ad_cost_stride(u64 max_len) { return max_len / 16 + 32; }
static inline u64 ad_off_stride(u64 max_len) { return max_len / 32 + 16; }
static inline u64 ad_large_bytes(u32 nf, u64 max_len) { return max_len >= ADL_MINB * ADL_MINB ? (u64)nf * adl_tmp_stride(max_len) : 0; }

extern "C" uint64_t hc_adapt_encode_ws_bytes(uint32_t nf, uint64_t max_len)
{
    return ad_large_bytes(nf, max_len) + (uint64_t)nf * ((ad_cost_stride(max_len) + ad_off_stride(max_len)) * 4 + 8) + 256;
}

// the adaptive encoder for files [0, nf) of the arrays it is given (the caller passes pointers to a group's first file)
static int adapt_encode_range(const uint8_t *in, const uint64_t *in_off, const uint64_t *width, const uint64_t *height,
                              uint8_t *out, const uint64_t *out_off, uint64_t *out_len, int32_t *status,
                              uint32_t nf, uint64_t max_len, u8 *ltmp, u64 lbytes, u64 tstride, u32 *cost, u64 cs, u32 *boff, u64 os,
                              u64 *cb, hc_stream_t stream)
{
    // work split: enough CTAs to fill 148 SMs, at most one chunk per 64 KiB of image
    u64 chunks = max_len / (64 * 1024);
    if (chunks < 1) chunks = 1;
    u64 want = (592 * 2 + (u64)nf * AD_NCAND - 1) / ((u64)nf * AD_NCAND);
    if (chunks > want) chunks = want < 1 ? 1 : want;
    if (chunks > 4096) chunks = 4096;
    // matrices up to 512 x 512: search from equality bitmasks in shared memory (adapt_mask.cuh);
    // larger ones: generic per-block evaluation
#ifndef HC_EMU
    // per device and per context: set on every call (cheap) rather than once per process
    HC_CUDA(cudaFuncSetAttribute(adapt_cost_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ACM_SMEM));
#endif
    HC_LAUNCH(adapt_cost_mask_kernel, dim3(nf < 1184u ? nf : 1184u), dim3(ACM_TPB), ACM_SMEM, stream, in, in_off, width, height,
              nf, cost, cs);
    HC_CHECK_LAUNCH();
    HC_LAUNCH(adapt_cost_kernel, grid2(chunks * AD_NCAND, nf), dim3(AD_COST_TPB), 0, stream, in, in_off, width,
              height, nf, cost, cs, (u32)chunks, (u64)ACM_MAX, (u64)ACM_MAX);
    HC_CHECK_LAUNCH();
    HC_LAUNCH(adapt_select_kernel, dim3(file_grid(nf)), dim3(256), 0, stream, width, height, nf,
              (const u32 *)cost, cs, boff, os, out, out_off, out_len, cb, status);
    HC_CHECK_LAUNCH();
    u64 echunks = max_len / (64 * 1024);
    if (echunks < 1) echunks = 1;
    u64 ewant = (592 * 2 + nf - 1) / nf;
    if (echunks > ewant) echunks = ewant;
    HC_LAUNCH(adapt_emit_kernel, grid2(echunks, nf), dim3(AD_EMIT_TPB), 0, stream, in, in_off, width, height, nf,
              (const u32 *)cost, cs, (const u32 *)boff, os, (const u64 *)cb, out, out_off, (const i32 *)status, true, tstride);
    HC_CHECK_LAUNCH();
    // files whose winning block size is >= 64: one CTA per block through the streaming coder
    if (lbytes) {
        u64 gx = max_len / (ADL_T * ADL_T) + 1;
        if (gx > 64) gx = 64;
        HC_LAUNCH(adapt_gather_large_kernel, grid2(gx, nf), dim3(ADL_TPB), 0, stream, in, in_off, width, height, nf,
                  (const u32 *)cost, cs, (const u64 *)cb, (const i32 *)status, ltmp, tstride);
        HC_CHECK_LAUNCH();
        u64 gb = max_len / (ADL_MINB * ADL_MINB) + 1;
        if (gb > 16) gb = 16;
        HC_LAUNCH(adapt_emit_large_kernel, grid2(gb, nf), dim3(RTPB), 0, stream, (const u8 *)ltmp, tstride, width, height, nf,
                  (const u32 *)boff, os, (const u64 *)cb, out, out_off, (const i32 *)status, in, in_off, (const u32 *)cost, cs);
        HC_CHECK_LAUNCH();
    }
    // files whose winning block size is 8/16/32: one thread per block over shared-memory staged rows
#ifndef HC_EMU
    HC_CUDA(cudaFuncSetAttribute(adapt_emit_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ADS_SMEM));
#endif
    {
        u64 gx = max_len / ADS_STRIP + 1;
        if (gx > 64) gx = 64;
        HC_LAUNCH(adapt_emit_small_kernel, grid2(gx, nf), dim3(ADS_TPB), ADS_SMEM, stream, in, in_off, width, height, nf,
                  (const u32 *)cost, cs, (const u32 *)boff, os, (const u64 *)cb, out, out_off, (const i32 *)status);
        HC_CHECK_LAUNCH();
    }
    return 0;
}

// L2-sized groups: the search kernel reads every pixel once and the emit kernels read them again; when the files of a
// group (and the streams they produce) fit the 126 MB L2 together, the second read does not go to HBM.  group_bytes = 0:
// the whole batch in one pass.
static u32 adapt_group_files(u32 nf, u64 bytes_per_file)
{
    // HC_ADAPT_GROUP_FILES / HC_ADAPT_GROUP_MB: experiment and test hooks
    static const u64 files = getenv("HC_ADAPT_GROUP_FILES") ? strtoull(getenv("HC_ADAPT_GROUP_FILES"), nullptr, 10) : 0;
    static const u64 bytes = getenv("HC_ADAPT_GROUP_MB") ? strtoull(getenv("HC_ADAPT_GROUP_MB"), nullptr, 10) << 20 : (u64)HC_ADAPT_GROUP_DEFAULT_MB << 20;
    u64 g = nf;
    if (files) g = files;
    else if (bytes && bytes_per_file) {
        g = bytes / bytes_per_file;
        if (g < 296) g = 296;                             // never fewer than two files per SM
    }
    return (u32)(g < 1 ? 1 : (g > nf ? nf : g));
}

extern "C" int hc_adapt_encode_batch(const uint8_t *in, const uint64_t *in_off,
                                     const uint64_t *width, const uint64_t *height,
                                     uint8_t *out, const uint64_t *out_off, uint64_t *out_len,
                                     uint64_t *chosen_b, int32_t *status,
                                     uint32_t nf, uint64_t max_len, void *ws, hc_stream_t stream)
{
    if (nf == 0) return 0;
    const u64 cs = ad_cost_stride(max_len), os = ad_off_stride(max_len);
    const u64 lbytes = ad_large_bytes(nf, max_len), tstride = lbytes ? adl_tmp_stride(max_len) : 0;
    u8 *ltmp = (u8 *)ws;
    u32 *cost = (u32 *)((u8 *)ws + lbytes);
    u32 *boff = cost + (u64)nf * cs;
    u64 *cb = (u64 *)(((uintptr_t)(boff + (u64)nf * os) + 7) & ~(uintptr_t)7);
    const u32 group = adapt_group_files(nf, max_len);
    for (u32 f0 = 0; f0 < nf; f0 += group) {
        const u32 n = nf - f0 < group ? nf - f0 : group;
        HC_TRY(adapt_encode_range(in, in_off + f0, width + f0, height + f0, out, out_off + f0, out_len + f0, status + f0, n, max_len,
                                  ltmp + (u64)f0 * tstride, lbytes, tstride, cost + (u64)f0 * cs, cs, boff + (u64)f0 * os, os, cb + f0, stream));
    }
    if (chosen_b)
        HC_CUDA(cudaMemcpyAsync(chosen_b, cb, (size_t)nf * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

static inline u64 ad_blk_stride(u64 max_out_len) { return max_out_len / 32 + 16; }

extern "C" uint64_t hc_adapt_decode_ws_bytes(uint32_t nf, uint64_t max_out_len)
{
    // block streams of the large-block path | block start table
    return ad_large_bytes(nf, max_out_len) + (uint64_t)nf * ad_blk_stride(max_out_len) * 4 + 256;
}

// the adaptive decoder for files [0, nf) of the arrays it is given; ws = this group's block table
static int adapt_decode_range(const uint8_t *in, const uint64_t *in_off, const uint64_t *in_len,
                              uint8_t *out, const uint64_t *out_off, const uint64_t *out_cap,
                              uint64_t *out_len, int32_t *status, uint32_t nf, uint64_t max_out_len,
                              u8 *ltmp, u64 lbytes, u64 tstride, void *ws, u64 bs, hc_stream_t stream)
{
    // matrices of at least this many bytes (with blocks of 64 and more) are indexed by the CTA-wide kernel
    // (HC_INDEX_WIDE_MIN: test hook, so that the CPU suite reaches that kernel with small images)
    static const u64 wide_min = getenv("HC_INDEX_WIDE_MIN") ? strtoull(getenv("HC_INDEX_WIDE_MIN"), nullptr, 10) : (u64)(4u << 20);
    HC_LAUNCH(adapt_index_warp_kernel, dim3(file_grid(nf)), dim3(32), 0, stream, in, in_off, in_len, out_cap, out != nullptr,
              (u32 *)ws, bs, out_len, status, nf, wide_min);
    HC_CHECK_LAUNCH();
    // big matrices with big blocks: eight warps per file walk the token stream together.  Always launched: which files
    // are "big" is decided by their headers (a crafted header may claim anything), the others return at once
    HC_LAUNCH(adapt_index_cta_kernel, dim3(file_grid(nf)), dim3(AD_IDX_WARPS * 32), 0, stream, in, in_off, in_len, out_cap,
              out != nullptr, (u32 *)ws, bs, out_len, status, nf, wide_min);
    HC_CHECK_LAUNCH();
    if (!out) return 0;
    u64 chunks = max_out_len / (64 * 1024);
    if (chunks < 1) chunks = 1;
    u64 want = (592 * 2 + nf - 1) / nf;
    if (chunks > want) chunks = want;
    HC_LAUNCH(adapt_expand_kernel, grid2(chunks, nf), dim3(AD_EXP_TPB), 0, stream, in, in_off, in_len, (const u32 *)ws, bs,
              out, out_off, (const i32 *)status, nf, true, tstride);
    HC_CHECK_LAUNCH();
    if (lbytes) {
        u64 gb = max_out_len / (ADL_MINB * ADL_MINB) + 1;
        if (gb > 16) gb = 16;
        HC_LAUNCH(adapt_expand_large_kernel, grid2(gb, nf), dim3(RTPB), 0, stream, in, in_off, in_len, (const u32 *)ws, bs,
                  (const i32 *)status, nf, ltmp, tstride, out, out_off);
        HC_CHECK_LAUNCH();
        u64 gx = max_out_len / (ADL_T * ADL_T) + 1;
        if (gx > 64) gx = 64;
        HC_LAUNCH(adapt_scatter_large_kernel, grid2(gx, nf), dim3(ADL_TPB), 0, stream, in, in_off, in_len,
                  (const i32 *)status, nf, (const u8 *)ltmp, tstride, out, out_off);
        HC_CHECK_LAUNCH();
    }
#ifndef HC_EMU
    HC_CUDA(cudaFuncSetAttribute(adapt_expand_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ADS_SMEM));
#endif
    {
        u64 gx = max_out_len / ADS_STRIP + 1;
        if (gx > 64) gx = 64;
        HC_LAUNCH(adapt_expand_small_kernel, grid2(gx, nf), dim3(ADS_TPB), ADS_SMEM, stream, in, in_off, in_len, (const u32 *)ws, bs,
                  out, out_off, (const i32 *)status, nf);
        HC_CHECK_LAUNCH();
    }
    // headers with more blocks than the table holds (block size < 8) take the serial decoder
    HC_LAUNCH(adapt_decode_kernel, dim3(file_grid(nf)), dim3(32), 0, stream, in, in_off, in_len, out, out_off, out_cap,
              out_len, status, nf, (i32)AD_ST_SERIAL);
    HC_CHECK_LAUNCH();
    return 0;
}

extern "C" int hc_adapt_decode_batch(const uint8_t *in, const uint64_t *in_off, const uint64_t *in_len,
                                     uint8_t *out, const uint64_t *out_off, const uint64_t *out_cap,
                                     uint64_t *out_len, int32_t *status,
                                     uint32_t nf, uint64_t max_in_len, uint64_t max_out_len,
                                     void *ws, hc_stream_t stream)
{
    (void)max_in_len;
    if (nf == 0) return 0;
    if (out && !ws) {
        // no scratch: one thread per file (slow, kept for callers that cannot provide ws)
        HC_LAUNCH(adapt_decode_kernel, dim3(file_grid(nf)), dim3(32), 0, stream, in, in_off, in_len, out, out_off,
                  out_cap, out_len, status, nf, (i32)-1);
        HC_CHECK_LAUNCH();
        return 0;
    }
    const u64 bs = ad_blk_stride(max_out_len);
    const u64 lbytes = ad_large_bytes(nf, max_out_len), tstride = lbytes ? adl_tmp_stride(max_out_len) : 0;
    u8 *ltmp = (u8 *)ws;
    u32 *tab = ws ? (u32 *)((u8 *)ws + lbytes) : nullptr;
    const u32 group = out ? adapt_group_files(nf, max_out_len) : nf;   // L2-sized groups: the index pass and the expansion read the same tokens
    for (u32 f0 = 0; f0 < nf; f0 += group) {
        const u32 n = nf - f0 < group ? nf - f0 : group;
        HC_TRY(adapt_decode_range(in, in_off + f0, in_len + f0, out, out ? out_off + f0 : nullptr, out_cap ? out_cap + f0 : nullptr, out_len + f0,
                                  status + f0, n, max_out_len, ltmp ? ltmp + (u64)f0 * tstride : nullptr, lbytes, tstride,
                                  tab ? (void *)(tab + (u64)f0 * bs) : nullptr, bs, stream));
    }
    return 0;
}

// order_ws: nf u32 of device scratch for the longest-first start order, or NULL (file order)
static int fgk_encode_launch(const uint8_t *sym, const uint64_t *sym_off, const uint64_t *sym_len, const uint8_t *flags,
                             uint8_t *out, const uint64_t *out_off, const uint64_t *out_cap, uint64_t *out_len,
                             int32_t *status, uint32_t nf, u32 *order_ws, hc_stream_t stream)
{
    if (nf == 0) return 0;
    if (order_ws) {
        HC_LAUNCH(fgk_order_kernel, dim3(1), dim3(1024), 0, stream, sym_len, nf, order_ws);
        HC_CHECK_LAUNCH();
    }
    HC_LAUNCH(fgk_encode_kernel, dim3((nf + FGK_ENC_WARPS - 1) / FGK_ENC_WARPS), dim3(FGK_ENC_WARPS * 32), 0, stream, sym,
              sym_off, sym_len, flags, out, out_off, out_cap, out_len, status, nf, (const u32 *)order_ws);
    HC_CHECK_LAUNCH();
    return 0;
}

extern "C" int hc_fgk_encode_batch(const uint8_t *sym, const uint64_t *sym_off, const uint64_t *sym_len,
                                   const uint8_t *flags,
                                   uint8_t *out, const uint64_t *out_off, const uint64_t *out_cap,
                                   uint64_t *out_len, int32_t *status,
                                   uint32_t nf, hc_stream_t stream)
{
    return fgk_encode_launch(sym, sym_off, sym_len, flags, out, out_off, out_cap, out_len, status, nf, nullptr, stream);
}

static int fgk_decode_launch(const uint8_t *in, const uint64_t *in_off, const uint64_t *in_len,
                             uint8_t *sym, const uint64_t *sym_off, const uint64_t *sym_cap,
                             uint64_t *sym_len, uint8_t *flags, int32_t *status,
                             uint32_t nf, u32 *order_ws, hc_stream_t stream)
{
    if (nf == 0) return 0;
    if (order_ws) {
        HC_LAUNCH(fgk_order_kernel, dim3(1), dim3(1024), 0, stream, in_len, nf, order_ws);   // compressed bytes ~ work
        HC_CHECK_LAUNCH();
    }
    HC_LAUNCH(fgk_decode_kernel, dim3((nf + FGK_DEC_WARPS - 1) / FGK_DEC_WARPS), dim3(FGK_DEC_WARPS * 32), 0, stream, in,
              in_off, in_len, sym, sym_off, sym_cap, sym_len, flags, status, nf, (const u32 *)order_ws);
    HC_CHECK_LAUNCH();
    return 0;
}

extern "C" int hc_fgk_decode_batch(const uint8_t *in, const uint64_t *in_off, const uint64_t *in_len,
                                   uint8_t *sym, const uint64_t *sym_off, const uint64_t *sym_cap,
                                   uint64_t *sym_len, uint8_t *flags, int32_t *status,
                                   uint32_t nf, hc_stream_t stream)
{
    return fgk_decode_launch(in, in_off, in_len, sym, sym_off, sym_cap, sym_len, flags, status, nf, nullptr, stream);
}

extern "C" int hc_offsets_from_lens(const uint64_t *len, uint64_t *out_off, uint64_t *total,
                                    uint32_t nf, uint32_t align, hc_stream_t stream)
{
    if (align == 0) align = 1;
    HC_LAUNCH(offsets_scan_kernel, dim3(1), dim3(TPB), 0, stream, len, out_off, total, nf, align);
    HC_CHECK_LAUNCH();
    return 0;
}

extern "C" int hc_gather_batch(const uint8_t *in, const uint64_t *in_off, const uint64_t *len,
                               uint8_t *out, const uint64_t *out_off,
                               uint32_t nf, uint64_t max_len, hc_stream_t stream)
{
    if (nf == 0) return 0;
    u64 x = (max_len + TPB * 16 * 4 - 1) / (TPB * 16 * 4);
    HC_LAUNCH(gather_kernel, grid2(x, nf), dim3(TPB), 0, stream, in, in_off, len, out, out_off, nf);
    HC_CHECK_LAUNCH();
    return 0;
}

// ======================================================================= codec (host level)
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t n)
    {
        if (n <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) return -(int)e;
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct HostBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t n)
    {
        if (n <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaHostAlloc(&p, n + 4096, cudaHostAllocDefault);
        if (e != cudaSuccess) return -(int)e;
        cap = n + 4096;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

#define HC_MAX_STAGES 16

struct hc_codec {
    int device = 0;
    cudaStream_t stream = nullptr;
    DevBuf in, a, b, out, tab, ws, ord;  // raw input, stage buffers, strided output, tables, scratch, FGK start order
    DevBuf jt;                           // offset / length / status tables of a host-level call (group pipelines)
    HostBuf htab;
    bool timing = false;
    cudaEvent_t ev[HC_MAX_STAGES + 1];
    cudaEvent_t wait_ev = nullptr;       // blocking-sync event: host threads waiting for a batch sleep instead of spinning
    const char *stage_names[HC_MAX_STAGES];
    int nstages = 0;
    bool ev_ready = false;
    std::vector<hc_codec *> kids;       // group pipelines of the host-level batch calls (own stream + buffers)
};

// the entry points leave the caller's current device as they found it
struct DevGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DevGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) err = cudaSetDevice(dev);
    }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define HC_ON_DEVICE(dev)                                  \
    DevGuard dev_guard_(dev);                              \
    if (dev_guard_.err != cudaSuccess) return -(int)dev_guard_.err

// on every exit of a host-level batch call (errors included) nothing of the call is in flight any more: the
// asynchronous copies into the caller's buffers have completed or were never started
struct KidsSync {
    hc_codec *c;
    explicit KidsSync(hc_codec *c_) : c(c_) {}
    ~KidsSync();
};

static void stage_begin(hc_codec *c)
{
    c->nstages = 0;
    if (c->timing) cudaEventRecord(c->ev[0], c->stream);
}
static void stage_mark(hc_codec *c, const char *name)
{
    if (!c->timing || c->nstages >= HC_MAX_STAGES) return;
    c->stage_names[c->nstages] = name;
    c->nstages++;
    cudaEventRecord(c->ev[c->nstages], c->stream);
}

// Host-side wait for everything queued on a codec's stream.  With eight processes of four pipeline threads each on one
// host (bench at N = 8) spinning waits would occupy every core, so the thread sleeps on a blocking-sync event.
static cudaError_t codec_wait(hc_codec *k)
{
    if (!k->wait_ev) return cudaStreamSynchronize(k->stream);
    cudaError_t e = cudaEventRecord(k->wait_ev, k->stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(k->wait_ev);
}

KidsSync::~KidsSync()
{
    for (hc_codec *k : c->kids) codec_wait(k);
}

extern "C" int hc_codec_create(hc_codec **out, int device)
{
    int n = hc_device_count();
    if (n < 0) return n;
    if (n == 0 || device >= n) return -100;     // cudaErrorNoDevice
    HC_ON_DEVICE(device);
    hc_codec *c = new hc_codec();
    c->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return -(int)e; }
    for (int i = 0; i <= HC_MAX_STAGES; i++) cudaEventCreate(&c->ev[i]);
    cudaEventCreateWithFlags(&c->wait_ev, cudaEventBlockingSync | cudaEventDisableTiming);
    c->ev_ready = true;
    *out = c;
    return 0;
}

extern "C" void hc_codec_destroy(hc_codec *c)
{
    if (!c) return;
    for (hc_codec *k : c->kids) hc_codec_destroy(k);
    c->kids.clear();
    DevGuard dev_guard_(c->device);
    cudaStreamSynchronize(c->stream);
    c->in.release(); c->a.release(); c->b.release(); c->out.release(); c->tab.release(); c->ws.release(); c->ord.release(); c->jt.release();
    c->htab.release();
    if (c->ev_ready) for (int i = 0; i <= HC_MAX_STAGES; i++) cudaEventDestroy(c->ev[i]);
    if (c->wait_ev) cudaEventDestroy(c->wait_ev);
    cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" void *hc_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void hc_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" hc_stream_t hc_codec_stream(hc_codec *c) { return (hc_stream_t)c->stream; }
extern "C" void hc_codec_enable_stage_timing(hc_codec *c, int on) { c->timing = on != 0; }
extern "C" const char *hc_stage_name(hc_codec *c, int i) { return (i >= 0 && i < c->nstages) ? c->stage_names[i] : ""; }
extern "C" int hc_codec_stage_times(hc_codec *c, float *ms, int max_stages)
{
    if (!c->timing) return 0;
    int n = c->nstages < max_stages ? c->nstages : max_stages;
    for (int i = 0; i < n; i++) {
        float t = 0;
        if (cudaEventElapsedTime(&t, c->ev[i], c->ev[i + 1]) != cudaSuccess) return -1;
        ms[i] = t;
    }
    return n;
}

static inline unsigned blocks_for(u32 nf) { return (nf + 255) / 256; }

// device table slots inside c->tab (each nf elements)
struct TabView {
    u8 *base;
    u32 nf;
    u64 *u64_at(int slot) const { return (u64 *)(base + (size_t)slot * nf * 8); }
    i32 *i32_at(int slot) const { return (i32 *)(base + (size_t)slot * nf * 8); }
    u8 *u8_at(int slot) const { return base + (size_t)slot * nf * 8; }
};
enum { T_OFF_A = 0, T_CAP_A, T_LEN_A, T_OFF_B, T_CAP_B, T_LEN_B, T_W, T_H, T_LEN_EFF, T_FLAGS, T_ST0, T_ST1, T_ST2,
       T_LEN_P, T_LEN_AD, T_N_P, T_N_AD, T_LEN_DIFF, T_CB, T_USER0, T_USER1, T_USER2, T_USER3, T_USER4, T_NSLOTS };

// worst-case adaptive output when only w*h = n is known (nb <= n/32 + 2 for w, h >= 8)
static inline u64 adapt_bound_from_len(u64 n) { return 24 + n / 256 + 8 + n + n / 3 + 4 * (n / 32 + 2) + 16; }

extern "C" int hc_compress_device(hc_codec *c, const uint8_t *d_in, const uint64_t *d_in_off,
                                  const uint64_t *d_in_len, const uint64_t *d_width,
                                  uint32_t nf, uint64_t max_len, int use_diff, int use_adapt,
                                  uint8_t *d_out, const uint64_t *d_out_off, const uint64_t *d_out_cap,
                                  uint64_t *d_out_len, int32_t *d_status)
{
    if (nf == 0) return 0;
    HC_ON_DEVICE(c->device);
    cudaStream_t s = c->stream;
    HC_TRY(c->tab.ensure((size_t)T_NSLOTS * nf * 8));
    TabView t{(u8 *)c->tab.p, nf};
    const u64 stride_a = align_up(max_len + 16, HC_ALIGN);
    const u64 bound_b = use_adapt ? adapt_bound_from_len(max_len) : hc_rle_bound(max_len);
    const u64 stride_b = align_up(bound_b + 16, HC_ALIGN);
    if (use_diff) HC_TRY(c->a.ensure((size_t)stride_a * nf));
    HC_TRY(c->b.ensure((size_t)stride_b * nf));
    if (use_adapt) HC_TRY(c->ws.ensure((size_t)hc_adapt_encode_ws_bytes(nf, max_len)));

    stage_begin(c);
    HC_LAUNCH(strided_offsets_kernel, dim3(blocks_for(nf)), dim3(256), 0, s, t.u64_at(T_OFF_A), t.u64_at(T_CAP_A), stride_a, nf);
    HC_LAUNCH(strided_offsets_kernel, dim3(blocks_for(nf)), dim3(256), 0, s, t.u64_at(T_OFF_B), t.u64_at(T_CAP_B), stride_b, nf);
    HC_LAUNCH(compress_prep_kernel, dim3(blocks_for(nf)), dim3(256), 0, s, d_in_len, d_width, t.u64_at(T_W), t.u64_at(T_H),
              t.u64_at(T_LEN_EFF), t.u8_at(T_FLAGS), t.i32_at(T_ST0), nf, use_diff, use_adapt);
    HC_CHECK_LAUNCH();
    stage_mark(c, "prep");

    const u8 *cur = d_in;
    const u64 *cur_off = d_in_off;
    const u64 *len_eff = t.u64_at(T_LEN_EFF);
    if (use_diff) {
        HC_TRY(hc_diff_apply_batch(cur, cur_off, (u8 *)c->a.p, t.u64_at(T_OFF_A), len_eff, nf, max_len, s));
        cur = (const u8 *)c->a.p;
        cur_off = t.u64_at(T_OFF_A);
        stage_mark(c, "diff_apply");
    }
    i32 *st1 = nullptr;
    if (use_adapt) {
        // files rejected by prep (status 6) carry width 0 -> the adaptive stage reports 12 for
        // them; the prep status wins in the merge below
        st1 = t.i32_at(T_ST1);
        HC_TRY(hc_adapt_encode_batch(cur, cur_off, t.u64_at(T_W), t.u64_at(T_H), (u8 *)c->b.p, t.u64_at(T_OFF_B),
                                     t.u64_at(T_LEN_B), nullptr, st1, nf, max_len, c->ws.p, s));
        stage_mark(c, "adapt_encode");
    } else {
        HC_TRY(hc_rle_encode_batch(cur, cur_off, len_eff, (u8 *)c->b.p, t.u64_at(T_OFF_B), t.u64_at(T_LEN_B), nf, max_len, s));
        stage_mark(c, "rle_encode");
    }
    HC_LAUNCH(mask_len_kernel, dim3(blocks_for(nf)), dim3(256), 0, s, t.u64_at(T_LEN_B), (const i32 *)t.i32_at(T_ST0),
              (const i32 *)st1, nf);
    HC_CHECK_LAUNCH();
    HC_TRY(c->ord.ensure((size_t)nf * 4));
    HC_TRY(fgk_encode_launch((const u8 *)c->b.p, t.u64_at(T_OFF_B), t.u64_at(T_LEN_B), t.u8_at(T_FLAGS), d_out, d_out_off,
                               d_out_cap, d_out_len, t.i32_at(T_ST2), nf, (u32 *)c->ord.p, s));
    stage_mark(c, "fgk_encode");
    HC_LAUNCH(merge_status_kernel, dim3(blocks_for(nf)), dim3(256), 0, s, (const i32 *)t.i32_at(T_ST0), (const i32 *)st1,
              (const i32 *)t.i32_at(T_ST2), d_status, d_out_len, nf);
    HC_CHECK_LAUNCH();
    stage_mark(c, "merge");
    return 0;
}

// stage 1 of decompression: header parse + FGK decode into c->a (strided), split by kind
static int dec_fgk(hc_codec *c, const uint8_t *d_in, const uint64_t *d_in_off, const uint64_t *d_in_len,
                   uint32_t nf, uint64_t max_sym_len)
{
    cudaStream_t s = c->stream;
    HC_TRY(c->tab.ensure((size_t)T_NSLOTS * nf * 8));
    TabView t{(u8 *)c->tab.p, nf};
    const u64 stride_a = align_up(max_sym_len + 16, HC_ALIGN);
    HC_TRY(c->a.ensure((size_t)stride_a * nf));
    HC_LAUNCH(strided_offsets_kernel, dim3(blocks_for(nf)), dim3(256), 0, s, t.u64_at(T_OFF_A), t.u64_at(T_CAP_A), stride_a, nf);
    HC_CHECK_LAUNCH();
    HC_TRY(c->ord.ensure((size_t)nf * 4));
    HC_TRY(fgk_decode_launch(d_in, d_in_off, d_in_len, (u8 *)c->a.p, t.u64_at(T_OFF_A), t.u64_at(T_CAP_A), t.u64_at(T_LEN_A),
                               t.u8_at(T_FLAGS), t.i32_at(T_ST0), nf, (u32 *)c->ord.p, s));
    stage_mark(c, "fgk_decode");
    HC_LAUNCH(decompress_split_kernel, dim3(blocks_for(nf)), dim3(256), 0, s, (const u64 *)t.u64_at(T_LEN_A),
              (const u8 *)t.u8_at(T_FLAGS), (const i32 *)t.i32_at(T_ST0), t.u64_at(T_LEN_P), t.u64_at(T_LEN_AD), nf);
    HC_CHECK_LAUNCH();
    return 0;
}

// stage 2: RLE / adaptive expansion (+ diff revert) of the symbol streams left in c->a.
// d_out == NULL only computes sizes and statuses.
static int dec_expand(hc_codec *c, uint32_t nf, uint64_t max_sym_len, uint64_t max_out_len, int kinds,
                      uint8_t *d_out, const uint64_t *d_out_off, const uint64_t *d_out_cap,
                      uint64_t *d_out_len, int32_t *d_status)
{
    cudaStream_t s = c->stream;
    TabView t{(u8 *)c->tab.p, nf};
    HC_CUDA(cudaMemsetAsync(t.u64_at(T_N_P), 0, (size_t)nf * 8, s));
    HC_CUDA(cudaMemsetAsync(t.u64_at(T_N_AD), 0, (size_t)nf * 8, s));
    HC_CUDA(cudaMemsetAsync(t.i32_at(T_ST1), 0, (size_t)nf * 8, s));
    HC_CUDA(cudaMemsetAsync(t.i32_at(T_ST2), 0, (size_t)nf * 8, s));
    if (kinds & HC_KIND_PLAIN) {
        HC_TRY(hc_rle_decode_batch((const u8 *)c->a.p, t.u64_at(T_OFF_A), t.u64_at(T_LEN_P), d_out, d_out_off, d_out_cap,
                                   t.u64_at(T_N_P), t.i32_at(T_ST1), nf, max_sym_len, s));
        stage_mark(c, d_out ? "rle_decode" : "rle_decode_size");
    }
    if (kinds & HC_KIND_ADAPT) {
        // non-adaptive files carry length 0 here (status 10 for them is ignored by the merge)
        if (d_out) HC_TRY(c->ws.ensure((size_t)hc_adapt_decode_ws_bytes(nf, max_out_len)));
        HC_TRY(hc_adapt_decode_batch((const u8 *)c->a.p, t.u64_at(T_OFF_A), t.u64_at(T_LEN_AD), d_out, d_out_off, d_out_cap,
                                     t.u64_at(T_N_AD), t.i32_at(T_ST2), nf, max_sym_len, max_out_len, d_out ? c->ws.p : nullptr, s));
        stage_mark(c, d_out ? "adapt_decode" : "adapt_decode_size");
    }
    HC_LAUNCH(decompress_merge_kernel, dim3(blocks_for(nf)), dim3(256), 0, s, (const u8 *)t.u8_at(T_FLAGS),
              (const i32 *)t.i32_at(T_ST0), (const u64 *)t.u64_at(T_N_P), (const i32 *)t.i32_at(T_ST1),
              (const u64 *)t.u64_at(T_N_AD), (const i32 *)t.i32_at(T_ST2), d_out_len, t.u64_at(T_LEN_DIFF), d_status, nf);
    HC_CHECK_LAUNCH();
    if ((kinds & HC_KIND_DIFF) && d_out) {
        // the adaptive / RLE expansion is done with c->ws by now (same stream): reuse it for the segment sums
        HC_TRY(c->ws.ensure((size_t)diff_revert_ws_bytes(nf, max_out_len) + 16));
        HC_TRY(diff_revert_launch(d_out, d_out_off, d_out, d_out_off, t.u64_at(T_LEN_DIFF), nf, max_out_len, c->ws.p, s));
        stage_mark(c, "diff_revert");
    }
    return 0;
}

extern "C" int hc_decompress_device(hc_codec *c, const uint8_t *d_in, const uint64_t *d_in_off,
                                    const uint64_t *d_in_len, uint32_t nf,
                                    uint64_t max_sym_len, uint64_t max_out_len, int kinds,
                                    uint8_t *d_out, const uint64_t *d_out_off, const uint64_t *d_out_cap,
                                    uint64_t *d_out_len, int32_t *d_status)
{
    if (nf == 0) return 0;
    HC_ON_DEVICE(c->device);
    if (kinds == 0) kinds = HC_KIND_PLAIN | HC_KIND_ADAPT | HC_KIND_DIFF;
    stage_begin(c);
    HC_TRY(dec_fgk(c, d_in, d_in_off, d_in_len, nf, max_sym_len));
    return dec_expand(c, nf, max_sym_len, max_out_len, kinds, d_out, d_out_off, d_out_cap, d_out_len, d_status);
}

// ---------------------------------------------------------------------- host buffers in / out
// Upload nf host files.  If the caller's layout is already vector friendly (16-byte aligned
// starts, dense, non-overlapping 16-byte slots) it is copied with ONE cudaMemcpyAsync and used
// as is; otherwise every file is copied into a 256-byte aligned packed layout.
static int upload_files(hc_codec *c, DevBuf &dst, const uint8_t *base, const uint64_t *off, const uint64_t *len,
                        uint32_t nf, std::vector<u64> &d_off)
{
    u64 lo = ~0ull, hi = 0, sum = 0;
    bool direct = true;
    for (u32 i = 0; i < nf; i++) {
        if (off[i] % 16) direct = false;
        if (off[i] < lo) lo = off[i];
        if (off[i] + len[i] > hi) hi = off[i] + len[i];
        sum += len[i];
        if (i + 1 < nf && (off[i + 1] < off[i] || off[i] + align_up(len[i], 16) > off[i + 1])) direct = false;
    }
    if (nf == 0) lo = hi = 0;
    if (hi - lo > sum + sum / 2 + (u64)nf * 256) direct = false;      // too sparse: re-pack
    d_off.resize(nf);
    if (direct) {
        HC_TRY(dst.ensure((size_t)(hi - lo) + 512));
        if (hi > lo) HC_CUDA(cudaMemcpyAsync(dst.p, base + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, c->stream));
        for (u32 i = 0; i < nf; i++) d_off[i] = off[i] - lo;
    } else {
        u64 pos = 0;
        for (u32 i = 0; i < nf; i++) { d_off[i] = pos; pos += align_up(len[i] + 16, HC_ALIGN); }
        HC_TRY(dst.ensure((size_t)pos + 512));
        for (u32 i = 0; i < nf; i++)
            if (len[i])
                HC_CUDA(cudaMemcpyAsync((u8 *)dst.p + d_off[i], base + off[i], (size_t)len[i], cudaMemcpyHostToDevice, c->stream));
    }
    return 0;
}

// ---- group pipelines -----------------------------------------------------------------------------
// The host-level calls split large batches into up to HC_MAX_GROUPS contiguous groups of files, each with
// its own stream and buffers (a child codec).  Host<->device copies of one group overlap the kernels
// of the others, and the FGK kernels of all groups are resident together -- that stage is bound by the
// latency of its longest stream, so running the groups one after another would multiply its time.
#define HC_MAX_GROUPS 8

static int ensure_kids(hc_codec *c, u32 g)
{
    while (c->kids.size() < g) {
        hc_codec *k = nullptr;
        int rc = hc_codec_create(&k, c->device);
        if (rc) return rc;
        c->kids.push_back(k);
    }
    return 0;
}

static u32 group_count(u32 nf)
{
    const char *env = getenv("HC_GROUPS");                 // tests / experiments: force the group count
    const int forced = env ? atoi(env) : 0;
    // a group should still fill the FGK stage on its own (2960 resident streams + the short ones that
    // follow them): measured on C3, 4096 files in 1 group 4.84 GB/s end to end, in 8 groups 4.19 GB/s
    u32 g = forced > 0 ? (u32)forced : nf / 4096;
    if (g > nf) g = nf;
    return g < 1 ? 1 : (g > HC_MAX_GROUPS ? HC_MAX_GROUPS : g);
}

static void group_range(u32 nf, u32 g, u32 ng, u32 *lo, u32 *hi)
{
    u32 base = nf / ng, extra = nf % ng;
    *lo = g * base + (g < extra ? g : extra);
    *hi = *lo + base + (g < extra ? 1 : 0);
}

struct GroupJob {
    u32 lo = 0, hi = 0;
    u64 *d = nullptr;           // device tables
    u64 max_a = 0, max_b = 0;   // per-call maxima (meaning depends on the direction)
    int kinds = 0;
};

extern "C" int hc_compress_batch(hc_codec *c,
                                 const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_len,
                                 uint32_t nf, int use_diff, int use_adapt, const uint64_t *width_host,
                                 uint8_t *out_base, uint64_t out_cap_total,
                                 uint64_t *out_off, uint64_t *out_len, int32_t *status)
{
    if (nf == 0) return 0;
    HC_ON_DEVICE(c->device);
    const u32 ng = group_count(nf);
    HC_TRY(ensure_kids(c, ng));
    KidsSync sync_on_exit(c);
    std::vector<GroupJob> jobs(ng);
    // phase 1: enqueue upload + the whole compression pipeline of every group
    for (u32 g = 0; g < ng; g++) {
        hc_codec *k = c->kids[g];
        GroupJob &j = jobs[g];
        group_range(nf, g, ng, &j.lo, &j.hi);
        const u32 n = j.hi - j.lo;
        const size_t N = n;
        cudaStream_t s = k->stream;
        std::vector<u64> d_off;
        HC_TRY(upload_files(k, k->in, in_base, in_off + j.lo, in_len + j.lo, n, d_off));
        // pinned host tables, 9 x n: in_off | in_len | width | out_off (strided) | out_cap | compact off | compact len
        //                              | out_len (result) | status (result)
        HC_TRY(k->htab.ensure(N * 8 * 9));
        u64 *h = (u64 *)k->htab.p;
        u64 max_len = 0, pos = 0;
        for (u32 i = 0; i < n; i++) {
            const u64 len = in_len[j.lo + i];
            h[i] = d_off[i];
            h[N + i] = len;
            if (len > max_len) max_len = len;
            h[2 * N + i] = width_host ? width_host[j.lo + i] : 512;
            const u64 m_bound = use_adapt ? adapt_bound_from_len(len) : hc_rle_bound(len);
            const u64 cap = align_up(hc_fgk_bound(m_bound) + 16, HC_ALIGN);
            h[3 * N + i] = pos;
            h[4 * N + i] = cap;
            pos += cap;
        }
        j.max_a = max_len;
        HC_TRY(k->out.ensure((size_t)pos + 512));
        // codec-owned (not cudaMallocAsync: with the default pool's release threshold every call pays a
        // fresh mapping, measured as stalls of up to 200 ms)
        HC_TRY(k->jt.ensure(N * 8 * 9));                        // 7 tables + out_len + status
        j.d = (u64 *)k->jt.p;
        u64 *d = j.d;
        HC_CUDA(cudaMemcpyAsync(d, h, N * 8 * 5, cudaMemcpyHostToDevice, s));
        HC_TRY(hc_compress_device(k, (const u8 *)k->in.p, d, d + N, d + 2 * N, n, max_len, use_diff, use_adapt,
                                  (u8 *)k->out.p, d + 3 * N, d + 4 * N, d + 7 * N, (i32 *)(d + 8 * N)));
        // results come back through PINNED memory: a copy into the caller's (possibly pageable) arrays
        // would block this loop until the group has finished and serialise the groups
        HC_CUDA(cudaMemcpyAsync(h + 7 * N, d + 7 * N, N * 8 * 2, cudaMemcpyDeviceToHost, s));
    }
    // phase 2: in file order, turn the sizes of a finished group into compact offsets, gather its
    // outputs on the device and start ONE device-to-host copy; later groups keep computing meanwhile
    u64 total = 0;
    for (u32 g = 0; g < ng; g++) {
        hc_codec *k = c->kids[g];
        GroupJob &j = jobs[g];
        const u32 n = j.hi - j.lo;
        const size_t N = n;
        cudaStream_t s = k->stream;
        HC_CUDA(codec_wait(k));
        u64 *h = (u64 *)k->htab.p;
        memcpy(out_len + j.lo, h + 7 * N, N * 8);
        memcpy(status + j.lo, h + 8 * N, N * 4);
        const u64 start = total;
        u64 max_out = 0;
        for (u32 i = 0; i < n; i++) {
            const u32 f = j.lo + i;
            const u64 len = status[f] == 0 ? out_len[f] : 0;
            if (status[f] != 0 && status[f] != HC_E_CAPACITY) out_len[f] = 0;
            out_off[f] = total;
            h[5 * N + i] = total - start;
            h[6 * N + i] = len;
            if (len > max_out) max_out = len;
            total += align_up(len, 16);
        }
        if (total > out_cap_total) {
            return HC_E_CAPACITY;                          // (KidsSync waits for the groups)
        }
        const u64 gbytes = total - start;
        HC_TRY(k->a.ensure((size_t)gbytes + 512));
        u64 *d = j.d;
        HC_CUDA(cudaMemcpyAsync(d + 5 * N, h + 5 * N, N * 8 * 2, cudaMemcpyHostToDevice, s));
        HC_TRY(hc_gather_batch((const u8 *)k->out.p, d + 3 * N, d + 6 * N, (u8 *)k->a.p, d + 5 * N, n, max_out, s));
        if (gbytes) HC_CUDA(cudaMemcpyAsync(out_base + start, k->a.p, (size_t)gbytes, cudaMemcpyDeviceToHost, s));
    }
    for (u32 g = 0; g < ng; g++) HC_CUDA(codec_wait(c->kids[g]));
    return 0;
}

extern "C" int hc_decompress_batch(hc_codec *c,
                                   const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_len,
                                   uint32_t nf,
                                   uint8_t *out_base, uint64_t out_cap_total,
                                   uint64_t *out_off, uint64_t *out_len, int32_t *status)
{
    if (nf == 0) return 0;
    HC_ON_DEVICE(c->device);
    const u32 ng = group_count(nf);
    HC_TRY(ensure_kids(c, ng));
    KidsSync sync_on_exit(c);
    std::vector<GroupJob> jobs(ng);
    // phase 1: upload, FGK decode and the size-only expansion pass of every group
    for (u32 g = 0; g < ng; g++) {
        hc_codec *k = c->kids[g];
        GroupJob &j = jobs[g];
        group_range(nf, g, ng, &j.lo, &j.hi);
        const u32 n = j.hi - j.lo;
        const size_t N = n;
        cudaStream_t s = k->stream;
        std::vector<u64> d_off;
        HC_TRY(upload_files(k, k->in, in_base, in_off + j.lo, in_len + j.lo, n, d_off));
        // header peek on the host: symbol counts size the FGK output, flag bytes select the kernels
        u64 max_sym = 0;
        int kinds = 0;
        for (u32 i = 0; i < n; i++) {
            const u64 len = in_len[j.lo + i];
            if (len < 9) continue;
            const u8 *p = in_base + in_off[j.lo + i];
            u64 m = 0;
            for (int b = 0; b < 8; b++) m |= (u64)p[b] << (8 * b);
            if (m > (len - 9) * 8 + 1) m = 0;               // cannot decode: the kernel reports 9
            if (m > max_sym) max_sym = m;
            kinds |= (p[8] & 0x40) ? HC_KIND_ADAPT : HC_KIND_PLAIN;
            if (p[8] & 0x80) kinds |= HC_KIND_DIFF;
        }
        if ((kinds & (HC_KIND_PLAIN | HC_KIND_ADAPT)) == 0) kinds |= HC_KIND_PLAIN;
        j.kinds = kinds;
        j.max_a = max_sym;
        // pinned tables, 6 x n: in_off | in_len | out_off | out_cap | out_len (result) | status (result)
        HC_TRY(k->htab.ensure(N * 8 * 6));
        u64 *h = (u64 *)k->htab.p;
        for (u32 i = 0; i < n; i++) { h[i] = d_off[i]; h[N + i] = in_len[j.lo + i]; }
        HC_TRY(k->jt.ensure(N * 8 * 6));
        j.d = (u64 *)k->jt.p;
        u64 *d = j.d;
        HC_CUDA(cudaMemcpyAsync(d, h, N * 8 * 2, cudaMemcpyHostToDevice, s));
        stage_begin(k);
        HC_TRY(dec_fgk(k, (const u8 *)k->in.p, d, d + N, n, max_sym));
        HC_TRY(dec_expand(k, n, max_sym, 0, kinds, nullptr, nullptr, nullptr, d + 4 * N, (i32 *)(d + 5 * N)));
        HC_CUDA(cudaMemcpyAsync(h + 4 * N, d + 4 * N, N * 8 * 2, cudaMemcpyDeviceToHost, s));   // pinned, see compress
    }
    // phase 2: in file order: offsets of the group, expansion straight into a compact buffer, copy out
    u64 total = 0;
    std::vector<u8> oversize(nf, 0);
    std::vector<u64> need(nf, 0);
    for (u32 g = 0; g < ng; g++) {
        hc_codec *k = c->kids[g];
        GroupJob &j = jobs[g];
        const u32 n = j.hi - j.lo;
        const size_t N = n;
        cudaStream_t s = k->stream;
        HC_CUDA(codec_wait(k));
        u64 *h = (u64 *)k->htab.p;
        memcpy(out_len + j.lo, h + 4 * N, N * 8);
        memcpy(status + j.lo, h + 5 * N, N * 4);
        memcpy(need.data() + j.lo, h + 4 * N, N * 8);
        const u64 start = total;
        u64 max_out = 0;
        for (u32 i = 0; i < n; i++) {
            const u32 f = j.lo + i;
            u64 len = status[f] == 0 ? out_len[f] : 0;
            // a file that does not fit any more fails alone (HC_E_CAPACITY, out_len = the size it needs); the
            // other files of the batch are not affected
            if (total + align_up(len, 16) > out_cap_total || total + align_up(len, 16) < total) { oversize[f] = 1; len = 0; }
            out_off[f] = total;
            h[2 * N + i] = total - start;
            h[3 * N + i] = align_up(len, 16);
            if (len > max_out) max_out = len;
            total += align_up(len, 16);
        }
        const u64 gbytes = total - start;
        HC_TRY(k->out.ensure((size_t)gbytes + 512));
        u64 *d = j.d;
        HC_CUDA(cudaMemcpyAsync(d + 2 * N, h + 2 * N, N * 8 * 2, cudaMemcpyHostToDevice, s));
        HC_TRY(dec_expand(k, n, j.max_a, max_out, j.kinds, (u8 *)k->out.p, d + 2 * N, d + 3 * N, d + 4 * N, (i32 *)(d + 5 * N)));
        HC_CUDA(cudaMemcpyAsync(h + 4 * N, d + 4 * N, N * 8 * 2, cudaMemcpyDeviceToHost, s));
        if (gbytes) HC_CUDA(cudaMemcpyAsync(out_base + start, k->out.p, (size_t)gbytes, cudaMemcpyDeviceToHost, s));
    }
    for (u32 g = 0; g < ng; g++) {
        hc_codec *k = c->kids[g];
        HC_CUDA(codec_wait(k));
        const size_t N = jobs[g].hi - jobs[g].lo;
        const u64 *h = (const u64 *)k->htab.p;
        memcpy(out_len + jobs[g].lo, h + 4 * N, N * 8);
        memcpy(status + jobs[g].lo, h + 5 * N, N * 4);
    }
    for (u32 i = 0; i < nf; i++) {
        if (oversize[i]) { status[i] = HC_E_CAPACITY; out_len[i] = need[i]; }
        else if (status[i] != 0) out_len[i] = 0;
    }
    return 0;
}

// ======================================================================= asynchronous pipeline
struct hc_pipeline {
    struct Slot {
        hc_codec *codec = nullptr;
        std::thread worker;
        std::deque<std::pair<int64_t, std::function<int()>>> jobs;
    };
    std::vector<Slot> slots;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::map<int64_t, int> done;         // finished, not yet collected by hc_pipeline_wait
    int64_t next_ticket = 0, collected_below = 0;
    std::map<int64_t, bool> collected;   // tickets >= collected_below that were already waited for
    bool stop = false;
};

static void pipeline_worker(hc_pipeline *p, size_t si)
{
    for (;;) {
        std::pair<int64_t, std::function<int()>> job;
        {
            std::unique_lock<std::mutex> lk(p->mu);
            p->cv_job.wait(lk, [&] { return p->stop || !p->slots[si].jobs.empty(); });
            if (p->slots[si].jobs.empty()) return;
            job = std::move(p->slots[si].jobs.front());
            p->slots[si].jobs.pop_front();
        }
        const int rc = job.second();
        {
            std::lock_guard<std::mutex> lk(p->mu);
            p->done[job.first] = rc;
        }
        p->cv_done.notify_all();
    }
}

extern "C" int hc_pipeline_create(hc_pipeline **out, int device, int depth)
{
    if (depth < 1) depth = 1;
    if (depth > 8) depth = 8;
    hc_pipeline *p = new hc_pipeline();
    p->slots.resize((size_t)depth);
    for (int i = 0; i < depth; i++) {
        int rc = hc_codec_create(&p->slots[(size_t)i].codec, device);
        if (rc) {
            for (int k = 0; k < i; k++) hc_codec_destroy(p->slots[(size_t)k].codec);
            delete p;
            return rc;
        }
    }
    for (int i = 0; i < depth; i++) p->slots[(size_t)i].worker = std::thread(pipeline_worker, p, (size_t)i);
    *out = p;
    return 0;
}

extern "C" void hc_pipeline_destroy(hc_pipeline *p)
{
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stop = true;
    }
    p->cv_job.notify_all();
    for (auto &s : p->slots) if (s.worker.joinable()) s.worker.join();      // pending jobs are finished first
    for (auto &s : p->slots) hc_codec_destroy(s.codec);
    delete p;
}

static int64_t pipeline_submit(hc_pipeline *p, std::function<int(hc_codec *)> fn)
{
    int64_t t;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        if (p->stop) return -1;
        t = p->next_ticket++;
        hc_pipeline::Slot &s = p->slots[(size_t)(t % (int64_t)p->slots.size())];
        hc_codec *c = s.codec;
        s.jobs.emplace_back(t, [fn, c] { return fn(c); });
    }
    p->cv_job.notify_all();
    return t;
}

extern "C" int64_t hc_pipeline_submit_compress(hc_pipeline *p,
                                               const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_len,
                                               uint32_t nf, int use_diff, int use_adapt, const uint64_t *width_host,
                                               uint8_t *out_base, uint64_t out_cap_total,
                                               uint64_t *out_off, uint64_t *out_len, int32_t *status)
{
    return pipeline_submit(p, [=](hc_codec *c) {
        return hc_compress_batch(c, in_base, in_off, in_len, nf, use_diff, use_adapt, width_host, out_base, out_cap_total, out_off,
                                 out_len, status);
    });
}

extern "C" int64_t hc_pipeline_submit_decompress(hc_pipeline *p,
                                                 const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_len,
                                                 uint32_t nf,
                                                 uint8_t *out_base, uint64_t out_cap_total,
                                                 uint64_t *out_off, uint64_t *out_len, int32_t *status)
{
    return pipeline_submit(p, [=](hc_codec *c) {
        return hc_decompress_batch(c, in_base, in_off, in_len, nf, out_base, out_cap_total, out_off, out_len, status);
    });
}

extern "C" int hc_pipeline_wait(hc_pipeline *p, int64_t ticket)
{
    std::unique_lock<std::mutex> lk(p->mu);
    if (ticket < 0 || ticket >= p->next_ticket) return -1;
    if (ticket < p->collected_below || p->collected.count(ticket)) return -1;     // a ticket is waited for once
    p->cv_done.wait(lk, [&] { return p->done.count(ticket) != 0; });
    const int rc = p->done[ticket];
    p->done.erase(ticket);
    p->collected[ticket] = true;
    while (p->collected.count(p->collected_below)) p->collected.erase(p->collected_below++);
    return rc;
}

// ======================================================================= multi-GPU: size all-gather
namespace hcd {
// padded all-gather result (world x per) -> global file order
HC_KERNEL shard_unpad_kernel(const u64 *HC_RESTRICT padded, u64 *HC_RESTRICT all, u32 n_total, u32 world, u32 per)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    const u32 base = n_total / world, extra = n_total % world;
    // owner of file i under the contiguous split (the first `extra` ranks own base + 1 files)
    const u32 cut = extra * (base + 1u);
    const u32 r = i < cut ? i / (base + 1u) : extra + (base ? (i - cut) / base : 0u);
    const u32 lo = r * base + (r < extra ? r : extra);
    all[i] = padded[(u64)r * per + (i - lo)];
}
HC_KERNEL shard_pad_kernel(const u64 *HC_RESTRICT local, u64 *HC_RESTRICT padded, u32 n_local, u32 per)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < per) padded[i] = i < n_local ? local[i] : 0;
}
}  // namespace hcd

extern "C" uint64_t hc_shard_ws_bytes(uint32_t n_total, int world)
{
    if (world < 1) world = 1;
    const uint64_t per = (n_total + (uint32_t)world - 1) / (uint32_t)world;
    return (per * ((uint64_t)world + 1) + 8) * 8;
}

extern "C" int hc_shard_sizes_allgather(void *nccl_comm, int rank, int world,
                                        const uint64_t *d_local_sizes, uint32_t n_total, uint32_t align,
                                        uint64_t *d_all_sizes, uint64_t *d_offsets, uint64_t *d_total,
                                        void *ws, hc_stream_t stream)
{
    if (world < 1 || rank < 0 || rank >= world) return -1;
    if (n_total == 0) return 0;
    const u32 base = n_total / (u32)world, extra = n_total % (u32)world;
    const u32 n_local = base + ((u32)rank < extra ? 1u : 0u);
    const u32 per = (n_total + (u32)world - 1) / (u32)world;
    u64 *send = (u64 *)ws, *recv = send + per;
    HC_LAUNCH(shard_pad_kernel, dim3((per + 255) / 256), dim3(256), 0, stream, d_local_sizes, send, n_local, per);
    HC_CHECK_LAUNCH();
    if (world == 1) {
        HC_CUDA(cudaMemcpyAsync(recv, send, (size_t)per * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    } else {
#ifdef HC_EMU
        (void)nccl_comm;
        return HC_E_NCCL - 999;
#else
        // ncclAllGather(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t); ncclUint64 = 5
        typedef int (*allgather_fn)(const void *, void *, size_t, int, void *, cudaStream_t);
        static allgather_fn fn = nullptr;
        if (!fn) {
            fn = (allgather_fn)dlsym(RTLD_DEFAULT, "ncclAllGather");       // the NCCL this process already uses
            if (!fn) {
                void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
                if (h) fn = (allgather_fn)dlsym(h, "ncclAllGather");
            }
            if (!fn) return HC_E_NCCL - 999;
        }
        const int rc = fn(send, recv, per, 5, nccl_comm, (cudaStream_t)stream);
        if (rc != 0) return HC_E_NCCL - rc;
#endif
    }
    HC_LAUNCH(shard_unpad_kernel, dim3((n_total + 255) / 256), dim3(256), 0, stream, (const u64 *)recv, d_all_sizes, n_total, (u32)world, per);
    HC_CHECK_LAUNCH();
    if (d_offsets) return hc_offsets_from_lens(d_all_sizes, d_offsets, d_total, n_total, align ? align : 1u, stream);
    return 0;
}
