/*
 * hc_b200.h -- C ABI of libhc_b200.so: the B200-native (sm_100a CUDA) implementation of
 * the huffman-codec compression pipeline.  This is the drop-in boundary for the hot path
 *
 *     input -> [differential model] -> (MNP-5 RLE | adaptive block RLE) -> FGK Huffman -> .out
 *
 * of dominiksalvet/huffman-codec.  The reference has no FFI; its in-process boundary is the
 * set of free functions in src/transform.hpp:23-49, src/headers.hpp:22-37 and the class
 * HuffTree (src/huffman.hpp:40-59), driven by huffCompress/huffDecompress (src/main.cpp:39-128).
 * Each entry point below names the reference interface it replaces.  INTEGRATION.md shows
 * the binding a maintainer of the reference would add.
 *
 * Conventions
 *  - Plain C: pointers and sizes only.  `hc_stream_t` is a cudaStream_t passed as void*.
 *  - BATCHES.  Every stage works on `nf` independent files ("streams") at once.  A batch
 *    buffer is one device allocation; file f occupies [off[f], off[f]+len[f]) with
 *    off[f] a multiple of HC_ALIGN and a capacity (distance to the next file's offset, or
 *    the end of the allocation) that is also a multiple of HC_ALIGN.  off/len/cap/status
 *    arrays are DEVICE arrays of nf elements unless the name says `_host`.
 *  - `max_len` (host value) is an upper bound of every len[f]; it only sizes the grid.
 *  - Stage calls are asynchronous on `stream`; they return 0 or -(cudaError_t).
 *  - The library never calls exit().  Where the reference prints an error and exits with
 *    code n, status[f] is set to n for that file and the other files of
 *    the batch are unaffected.  HC_E_CAPACITY / HC_E_CODELEN are library-only codes.
 *  - No CPU fallback exists: without a CUDA device every call fails with a negative code.
 */
#ifndef HC_B200_H
#define HC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HC_ALIGN 256u

/* per-file status values (reference exit codes) */
#define HC_OK 0
#define HC_E_BAD_2D_SIZE 6      /* src/main.cpp:54-58   -a and size % width != 0           */
#define HC_E_HUFF_HEADER 8      /* src/main.cpp:99-104  short/missing <count><flags>      */
#define HC_E_HUFF_DATA 9        /* src/transform.cpp:394-398 bitstream underrun            */
#define HC_E_ADAPT_HEADER 10    /* src/headers.cpp:67-71  adaptive header < 24 bytes       */
#define HC_E_ADAPT_DIRS 11      /* src/headers.cpp:94-98  direction bytes missing          */
#define HC_E_TOO_SMALL 12       /* src/transform.cpp:300-304 width or height < 8           */
#define HC_E_BLOCK_OVERSHOOT 13 /* src/transform.cpp:180-184                               */
#define HC_E_ADAPT_UNDERRUN 14  /* src/transform.cpp:170-174                               */
#define HC_E_ADAPT_LEFTOVER 15  /* src/transform.cpp:354-358                               */
#define HC_E_CAPACITY 100       /* output region too small (len[f] then holds the need)    */
#define HC_E_CODELEN 101        /* FGK code longer than 56 bits (needs > 2^32 symbols)      */
#define HC_E_INTERNAL 102       /* a tree walk did not end (inconsistent FGK tree): never on valid state */
#define HC_E_NCCL (-1000)       /* hc_shard_sizes_allgather: -(1000 + ncclResult_t); -1999 = NCCL not loadable */

typedef void *hc_stream_t;

const char *hc_version(void);
/* number of CUDA devices visible, or -(cudaError_t) */
int hc_device_count(void);
/* text for a negative return code (CUDA error string) or a per-file status */
const char *hc_error_string(int code);
/* number of kernels launched by this library since load (for bench.py's gpu_launches) */
uint64_t hc_launch_count(void);

/* ---- worst-case sizes (host arithmetic, usable without a device) ------------------- */
/* applyRLE output bound for n input bytes (worst case 4/3, src/transform.cpp:241-279) */
uint64_t hc_rle_bound(uint64_t n);
/* applyAdaptRLE output bound: header + direction bytes + per-block RLE bound */
uint64_t hc_adapt_bound(uint64_t width, uint64_t height);
/* .out bound for m symbols: 9-byte header + FGK bits (<= 2S+m bits) */
uint64_t hc_fgk_bound(uint64_t m);
/* getBlockCount, src/transform.cpp:410-418 */
uint64_t hc_block_count(uint64_t width, uint64_t height, uint64_t block_size);

/* ---- stage level: device pointers --------------------------------------------------- */

/* applyDiffModel / revertDiffModel, src/transform.cpp:220-229 / :231-239.
 * out may alias in only for the revert direction (the forward pass reads a halo byte). */
int hc_diff_apply_batch(const uint8_t *in, const uint64_t *in_off, uint8_t *out, const uint64_t *out_off,
                        const uint64_t *len, uint32_t nf, uint64_t max_len, hc_stream_t stream);
int hc_diff_revert_batch(const uint8_t *in, const uint64_t *in_off, uint8_t *out, const uint64_t *out_off,
                         const uint64_t *len, uint32_t nf, uint64_t max_len, hc_stream_t stream);

/* applyRLE, src/transform.cpp:241-279.  Region capacity of out must be >= hc_rle_bound(len). */
int hc_rle_encode_batch(const uint8_t *in, const uint64_t *in_off, const uint64_t *in_len,
                        uint8_t *out, const uint64_t *out_off, uint64_t *out_len,
                        uint32_t nf, uint64_t max_len, hc_stream_t stream);

/* revertRLE, src/transform.cpp:281-292 (+ revertRLEStep :137-159).
 * out_len[f] always receives the decoded size; if it exceeds out_cap[f] nothing beyond the
 * capacity is written and status[f] = HC_E_CAPACITY.  out == NULL computes sizes only. */
int hc_rle_decode_batch(const uint8_t *in, const uint64_t *in_off, const uint64_t *in_len,
                        uint8_t *out, const uint64_t *out_off, const uint64_t *out_cap,
                        uint64_t *out_len, int32_t *status,
                        uint32_t nf, uint64_t max_len, hc_stream_t stream);

/* applyAdaptRLE (block-size search), src/transform.cpp:294-328 with :25-134 and
 * createAdaptRLEHeader src/headers.cpp:18-63.  width/height are per-file device arrays.
 * chosen_b (nullable) receives the selected block size.  status: HC_E_TOO_SMALL.
 * ws: device scratch of hc_adapt_encode_ws_bytes(nf, max_w*max_h) bytes. */
uint64_t hc_adapt_encode_ws_bytes(uint32_t nf, uint64_t max_len);
int hc_adapt_encode_batch(const uint8_t *in, const uint64_t *in_off,
                          const uint64_t *width, const uint64_t *height,
                          uint8_t *out, const uint64_t *out_off, uint64_t *out_len,
                          uint64_t *chosen_b, int32_t *status,
                          uint32_t nf, uint64_t max_len, void *ws, hc_stream_t stream);

/* revertAdaptRLE, src/transform.cpp:330-361 with extractAdaptRLEHeader src/headers.cpp:65-105.
 * status: 10, 11, 13, 14, 15 or HC_E_CAPACITY.  out == NULL computes sizes/status of the
 * header only (out_len = width*height).  ws: hc_adapt_decode_ws_bytes(nf, max_out) bytes. */
uint64_t hc_adapt_decode_ws_bytes(uint32_t nf, uint64_t max_out_len);
int hc_adapt_decode_batch(const uint8_t *in, const uint64_t *in_off, const uint64_t *in_len,
                          uint8_t *out, const uint64_t *out_off, const uint64_t *out_cap,
                          uint64_t *out_len, int32_t *status,
                          uint32_t nf, uint64_t max_in_len, uint64_t max_out_len,
                          void *ws, hc_stream_t stream);

/* applyHuffman + createHuffHeader + bit packing: src/transform.cpp:363-384,
 * src/headers.cpp:107-125, src/main.cpp:73-84 (HuffTree::encode/update src/huffman.cpp:37-58,95-128).
 * Writes the complete .out of every file: <u64 LE sym_len><u8 flags><bits MSB first, zero padded>.
 * flags[f] is the header flag byte (bit7 diff model, bit6 adaptive RLE). */
int hc_fgk_encode_batch(const uint8_t *sym, const uint64_t *sym_off, const uint64_t *sym_len,
                        const uint8_t *flags,
                        uint8_t *out, const uint64_t *out_off, const uint64_t *out_cap,
                        uint64_t *out_len, int32_t *status,
                        uint32_t nf, hc_stream_t stream);

/* header parse + revertHuffman: src/main.cpp:93-113, src/transform.cpp:386-406
 * (HuffTree::decode src/huffman.cpp:60-93).  in is a batch of complete .out files.
 * sym_len[f] = <64b-byte-count> of the header, flags[f] = its flag byte.
 * status: 8 (in_len < 9), 9 (ran out of bits), HC_E_CAPACITY (count > sym_cap). */
int hc_fgk_decode_batch(const uint8_t *in, const uint64_t *in_off, const uint64_t *in_len,
                        uint8_t *sym, const uint64_t *sym_off, const uint64_t *sym_cap,
                        uint64_t *sym_len, uint8_t *flags, int32_t *status,
                        uint32_t nf, hc_stream_t stream);

/* utilities used between stages ------------------------------------------------------ */
/* out_off[f] = sum_{g<f} align_up(len[g], align) (exclusive scan); total (nullable) gets the sum */
int hc_offsets_from_lens(const uint64_t *len, uint64_t *out_off, uint64_t *total,
                         uint32_t nf, uint32_t align, hc_stream_t stream);
/* gather file regions into another layout (e.g. capacity-strided -> compact) */
int hc_gather_batch(const uint8_t *in, const uint64_t *in_off, const uint64_t *len,
                    uint8_t *out, const uint64_t *out_off,
                    uint32_t nf, uint64_t max_len, hc_stream_t stream);

/* ---- host level: whole files in host memory (huffCompress / huffDecompress) ---------- */

typedef struct hc_codec hc_codec;
#define HC_KIND_PLAIN 1  /* some file uses plain MNP-5 RLE   */
#define HC_KIND_ADAPT 2  /* some file uses adaptive block RLE */
#define HC_KIND_DIFF 4   /* some file uses the diff model     */

/* creates a codec bound to CUDA device `device` (workspace, stream, staging buffers) */
int hc_codec_create(hc_codec **out, int device);
void hc_codec_destroy(hc_codec *c);
/* pinned host memory for fast transfers (optional; any host memory is accepted) */
void *hc_host_alloc(size_t bytes);
void hc_host_free(void *p);

/* huffCompress, src/main.cpp:39-87, for nf files.
 *  in_base/in_off/in_len : host memory; file f = in_base[in_off[f] .. +in_len[f])
 *  use_diff, use_adapt   : the -m / -a switches; width_host: per-file -w (NULL = 512 for all)
 *  out_base              : host buffer of out_cap_total bytes; the files are written back to
 *                          back (each start aligned to 16 bytes); out_off/out_len/status are
 *                          host arrays of nf elements written by the call.
 * status[f]: 0, 6 (-a and len % width != 0), 12 (width/height < 8), HC_E_CAPACITY (also for a file of 2 GiB or
 * more: per-file positions inside the transform kernels are 32-bit), HC_E_CODELEN (more than 2^32 - 16 symbols).
 * Returns 0, -(cudaError_t), or HC_E_CAPACITY if out_cap_total is too small. */
int hc_compress_batch(hc_codec *c,
                      const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_len,
                      uint32_t nf, int use_diff, int use_adapt, const uint64_t *width_host,
                      uint8_t *out_base, uint64_t out_cap_total,
                      uint64_t *out_off, uint64_t *out_len, int32_t *status);

/* huffDecompress, src/main.cpp:90-128, for nf .out files; same buffer conventions.
 * status[f]: 0, 8, 9, 10, 11, 13, 14, 15, or HC_E_CAPACITY for a file that does not fit what is left of
 * out_base (out_len[f] = the size it needs; the other files of the batch are decoded normally -- the call
 * itself still returns 0).  Crafted adaptive headers that promise more than 255 output bytes per payload
 * byte (no MNP-5 token expands further; the reference dies in its allocation, src/transform.cpp:340) are
 * walked for their error code (13 / 14) and never sized. */
int hc_decompress_batch(hc_codec *c,
                        const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_len,
                        uint32_t nf,
                        uint8_t *out_base, uint64_t out_cap_total,
                        uint64_t *out_off, uint64_t *out_len, int32_t *status);

/* ---- asynchronous host level: a pipeline of `depth` codecs, each with its own stream, buffers and
 * worker thread (replaces the one-file-at-a-time `ifs.get()` loading of src/main.cpp:46-51 by batched,
 * overlapped transfers).  A submit returns at once; the job runs hc_compress_batch / hc_decompress_batch
 * on the next slot, so the host<->device copies of one batch overlap the kernels of the batches around
 * it.  Buffers named in a submit must stay valid and untouched until hc_pipeline_wait returns for the
 * ticket.  Jobs start in submit order; each slot runs one job at a time. */
typedef struct hc_pipeline hc_pipeline;
int hc_pipeline_create(hc_pipeline **out, int device, int depth /* 1..8 */);
void hc_pipeline_destroy(hc_pipeline *p);
/* same arguments as hc_compress_batch / hc_decompress_batch; returns a ticket >= 0 or a negative error */
int64_t hc_pipeline_submit_compress(hc_pipeline *p,
                                    const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_len,
                                    uint32_t nf, int use_diff, int use_adapt, const uint64_t *width_host,
                                    uint8_t *out_base, uint64_t out_cap_total,
                                    uint64_t *out_off, uint64_t *out_len, int32_t *status);
int64_t hc_pipeline_submit_decompress(hc_pipeline *p,
                                      const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_len,
                                      uint32_t nf,
                                      uint8_t *out_base, uint64_t out_cap_total,
                                      uint64_t *out_off, uint64_t *out_len, int32_t *status);
/* blocks until the job has finished; returns what the synchronous call would have returned */
int hc_pipeline_wait(hc_pipeline *p, int64_t ticket);

/* ---- multi-GPU: the path's only collective.  The batch of n_total files is cut into
 * contiguous shards (rank r owns files [r*base + min(r, extra), ...), base = n_total / world, extra =
 * n_total % world); every rank passes the device array of its shard's per-file output sizes and gets
 * back, in global file order, all sizes, the offsets of the concatenated container (starts aligned to
 * `align` bytes) and its total size.  `nccl_comm` is an ncclComm_t; the NCCL library already loaded in
 * the process (or libnccl.so.2) is bound at run time, libhc_b200.so does not link it.  `ws` = device
 * scratch of hc_shard_ws_bytes() bytes.  Asynchronous on `stream`.  Returns 0, -(cudaError_t) or HC_E_NCCL - code. */
uint64_t hc_shard_ws_bytes(uint32_t n_total, int world);
int hc_shard_sizes_allgather(void *nccl_comm, int rank, int world,
                             const uint64_t *d_local_sizes, uint32_t n_total, uint32_t align,
                             uint64_t *d_all_sizes, uint64_t *d_offsets, uint64_t *d_total,
                             void *ws, hc_stream_t stream);

/* device-resident variants used by bench.py's kernel-only timing: inputs already in HBM,
 * outputs left in HBM (d_out compact, 256-byte aligned starts); nothing crosses PCIe and
 * nothing synchronises.  d_* are device pointers; *_host are host values. */
int hc_compress_device(hc_codec *c, const uint8_t *d_in, const uint64_t *d_in_off,
                       const uint64_t *d_in_len, const uint64_t *d_width /* NULL = 512 */,
                       uint32_t nf, uint64_t max_len_host, int use_diff, int use_adapt,
                       uint8_t *d_out, const uint64_t *d_out_off, const uint64_t *d_out_cap,
                       uint64_t *d_out_len, int32_t *d_status);
int hc_decompress_device(hc_codec *c, const uint8_t *d_in, const uint64_t *d_in_off,
                         const uint64_t *d_in_len, uint32_t nf,
                         uint64_t max_sym_len_host, uint64_t max_out_len_host,
                         int kinds_hint /* HC_KIND_* bits known to occur, 0 = unknown */,
                         uint8_t *d_out, const uint64_t *d_out_off, const uint64_t *d_out_cap,
                         uint64_t *d_out_len, int32_t *d_status);
/* the stream the codec launches on (cudaStream_t) -- time with events on THIS stream */
hc_stream_t hc_codec_stream(hc_codec *c);
/* per-stage device time of the last hc_*_device call in ms, in launch order; names in
 * hc_stage_name(i); returns the number of stages recorded (needs a prior stream sync) */
int hc_codec_stage_times(hc_codec *c, float *ms, int max_stages);
const char *hc_stage_name(hc_codec *c, int i);
void hc_codec_enable_stage_timing(hc_codec *c, int on);

#ifdef __cplusplus
}
#endif
#endif /* HC_B200_H */
