// adapt.cuh -- adaptive block RLE (reference: src/transform.cpp:25-134, 162-216, 294-361;
// header src/headers.cpp:18-105).
//
// ENCODE = three kernels per batch:
//   adapt_cost_kernel   for every candidate block size B = 8 << k (k = 0..7, B <= W, B <= H) and
//                       every block: the MNP-5 size of the block read row-major (horizontal)
//                       and column-major (vertical); keeps min and the direction bit
//                       (tie -> horizontal, src/transform.cpp:114).  A group of L lanes owns a
//                       block: each lane folds a contiguous chunk of the block's sequence into
//                       a RunSum (runsum.cuh), the group reduces them with warp shuffles.
//   adapt_select_kernel per file: total(B) = 24 + ceil(nb/8) + sum of block sizes; smallest
//                       total wins, ties keep the smaller B (strict '<', :319); writes the
//                       header <W><H><B> big endian + direction bits (32 blocks per warp
//                       ballot, block 0 = MSB) and the exclusive scan of the block sizes.
//   adapt_emit_kernel   RLE-encodes every block of the winning size in its direction at its
//                       scanned offset (same lane-chunk decomposition, RunSum prefix per lane).
// DECODE: block boundaries depend on the decoded data itself (fresh RLE state per block,
// a block ends when bw*bh bytes exist), so v1 walks each file's token stream with one thread.
#pragma once
#include "rle.cuh"
#include "runsum.cuh"
#include "scan.cuh"

namespace hcd {

constexpr int AD_NCAND = 8;                 // INIT_RLE_BLOCK_SIZE 8, MAX_RLE_DOUBLING_STEPS 7
constexpr u32 AD_CHUNK = 16;                // target elements per lane

struct BlockGeom { u64 base; u32 bw, bh; };

HC_HD u64 ad_nblocks(u64 w, u64 h, u64 b) { return ((w + b - 1) / b) * ((h + b - 1) / b); }

HC_HD BlockGeom ad_geom(u64 w, u64 h, u64 b, u64 idx)
{
    u64 bil = (w + b - 1) / b;
    u64 bx = (idx % bil) * b, by = (idx / bil) * b;
    BlockGeom g;
    g.base = by * w + bx;
    g.bw = (u32)(bx + b > w ? w - bx : b);
    g.bh = (u32)(by + b > h ? h - by : b);
    return g;
}

// candidate k is evaluated iff k == 0 or (8<<k) <= min(w, h)   (src/transform.cpp:309-315)
HC_HD bool ad_cand_valid(u64 w, u64 h, int k) { u64 b = 8ull << k; return k == 0 || (b <= w && b <= h); }

// offset (in u32 entries) of candidate k's per-block table inside a file's cost scratch
HC_HD u64 ad_kbase(u64 w, u64 h, int k)
{
    u64 o = 0;
    for (int i = 0; i < k; i++)
        if (ad_cand_valid(w, h, i)) o += ad_nblocks(w, h, 8ull << i);
    return o;
}

// lanes per block for block size b: power of two, ~AD_CHUNK elements per lane, <= 32
HC_HD u32 ad_lanes(u64 b)
{
    u64 n = b * b;
    u32 l = 1;
    while (l < 32 && (u64)l * AD_CHUNK < n) l <<= 1;
    return l;
}

// eligibility of a file for the small-block kernels (adapt_small.cuh); same formula as
// ads_rows_per_group, declared here because the lane-group kernels must skip those files
HC_HD u32 ads_rows_per_group_fwd(u64 w, u64 b)
{
    if (b > 32 || b * w > 32 * 1024) return 0;
    const u64 ncb = (w + b - 1) / b;
    u64 s = 256 / ncb;
    const u64 cap = (32 * 1024) / (b * w);
    if (s > cap) s = cap;
    return (u32)s;
}

// eligibility of a file for the large-block kernels (adapt_large.cuh): one CTA per block through
// the streaming coder of rle.cuh
constexpr u64 ADL_MINB = 64;
constexpr u64 ADL_MAXB = 4096;                 // bw*bh fits 32 bits with room
HC_HD bool adl_eligible(u64 w, u64 h, u64 b)
{
    return b >= ADL_MINB && b <= ADL_MAXB && (b & 15u) == 0 && w <= 0x7fffffffull && h <= 0x7fffffffull && w >= 1 && h >= 1;
}

// sequential reader of a block in horizontal (row-major) or vertical (column-major) order
struct BlockCursor {
    const u8 *p;       // matrix + block base
    u64 w;             // matrix width
    u32 inner, i, o;   // inner extent (bw for hor, bh for ver), inner index, outer index
    bool hor;
};

HC_DEV void bc_seek(BlockCursor &c, const u8 *mat, u64 w, const BlockGeom &g, bool hor, u32 pos)
{
    c.p = mat + g.base;
    c.w = w;
    c.hor = hor;
    c.inner = hor ? g.bw : g.bh;
    c.o = pos / c.inner;
    c.i = pos % c.inner;
}

HC_DEV u32 bc_next(BlockCursor &c)
{
    u64 a = c.hor ? (u64)c.o * c.w + c.i : (u64)c.i * c.w + c.o;
    u32 b = ldg8(c.p + a);
    if (++c.i == c.inner) { c.i = 0; c.o++; }
    return b;
}

HC_DEV RunSum rs_shfl_down(const RunSum &s, unsigned d)
{
    RunSum r;
    r.head = shfl_down(s.head, d); r.tail = shfl_down(s.tail, d);
    r.inner = shfl_down(s.inner, d); r.meta = shfl_down(s.meta, d);
    return r;
}
HC_DEV RunSum rs_shfl_up(const RunSum &s, unsigned d)
{
    RunSum r;
    r.head = shfl_up(s.head, d); r.tail = shfl_up(s.tail, d);
    r.inner = shfl_up(s.inner, d); r.meta = shfl_up(s.meta, d);
    return r;
}

// RunSum of elements [lo, hi) of a block sequence
HC_DEV RunSum ad_chunk_sum(const u8 *mat, u64 w, const BlockGeom &g, bool hor, u32 lo, u32 hi)
{
    RunSum s = rs_empty();
    if (lo < hi) {
        BlockCursor c;
        bc_seek(c, mat, w, g, hor, lo);
        for (u32 i = lo; i < hi; i++) rs_push(s, bc_next(c));
    }
    return s;
}

// ordered reduction over the L lanes of a group; result valid in group lane 0
HC_DEV RunSum ad_group_reduce(RunSum s, u32 L, u32 gl)
{
    for (u32 d = 1; d < L; d <<= 1) {
        RunSum o = rs_shfl_down(s, d);
        if ((gl & (2 * d - 1)) == 0 && gl + d < L) s = rs_combine(s, o);
    }
    return s;
}

constexpr int AD_COST_TPB = 256;

HC_KERNEL HC_LAUNCH_BOUNDS(AD_COST_TPB, 4)
adapt_cost_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT width,
                  const u64 *HC_RESTRICT height, u32 nf, u32 *HC_RESTRICT cost, u64 cost_stride, u32 nchunks,
                  u64 skip_w, u64 skip_h)
{
    const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const int k = (int)(blockIdx.x % AD_NCAND);
    const u32 chunk_id = blockIdx.x / AD_NCAND;
    const u64 b = 8ull << k;
    const u32 L = ad_lanes(b), G = 32u / L, gl = lane % L, gi = lane / L;
    for (u32 f = blockIdx.y; f < nf; f += gridDim.y) {
        const u64 w = width[f], h = height[f];
        if (w < 8 || h < 8 || !ad_cand_valid(w, h, k)) continue;
        if (w <= skip_w && h <= skip_h) continue;          // handled by adapt_cost_mask_kernel
        const u8 *mat = in + in_off[f];
        u32 *tab = cost + (u64)f * cost_stride + ad_kbase(w, h, k);
        const u64 nb = ad_nblocks(w, h, b);
        const u64 ngroups = (nb + G - 1) / G;
        const u64 stride = (u64)nchunks * (AD_COST_TPB / 32);
        for (u64 grp = (u64)chunk_id * (AD_COST_TPB / 32) + wid; grp < ngroups; grp += stride) {
            const u64 blk = grp * G + gi;
            const bool live = blk < nb;
            BlockGeom g = ad_geom(w, h, b, live ? blk : 0);
            const u32 n = live ? g.bw * g.bh : 0u;
            const u32 per = (n + L - 1) / L;
            u32 lo = gl * per, hi = lo + per;
            if (lo > n) lo = n;
            if (hi > n) hi = n;
            RunSum sh = ad_group_reduce(ad_chunk_sum(mat, w, g, true, lo, hi), L, gl);
            RunSum sv = ad_group_reduce(ad_chunk_sum(mat, w, g, false, lo, hi), L, gl);
            if (live && gl == 0) {
                u32 ch = rs_cost_final(sh), cv = rs_cost_final(sv);
                tab[blk] = ch <= cv ? (ch | 0x80000000u) : cv;   // bit 31 = horizontal
            }
        }
    }
}

// one CTA per file
HC_KERNEL HC_LAUNCH_BOUNDS(256, 2)
adapt_select_kernel(const u64 *HC_RESTRICT width, const u64 *HC_RESTRICT height, u32 nf,
                    const u32 *HC_RESTRICT cost, u64 cost_stride, u32 *HC_RESTRICT blk_off, u64 off_stride,
                    u8 *HC_RESTRICT out, const u64 *HC_RESTRICT out_off, u64 *HC_RESTRICT out_len,
                    u64 *HC_RESTRICT chosen_b, i32 *HC_RESTRICT status)
{
    HC_SHARED u64 red[NW];
    HC_SHARED u64 s_total[AD_NCAND];
    HC_SHARED u32 s_wsum[NW];
    HC_SHARED int s_best;
    const u32 tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    for (u32 f = blockIdx.x; f < nf; f += gridDim.x) {
        const u64 w = width[f], h = height[f];
        if (w < 8 || h < 8) {                              // src/transform.cpp:300-304
            if (tid == 0) { out_len[f] = 0; if (chosen_b) chosen_b[f] = 0; status[f] = 12; }
            continue;
        }
        const u32 *ctab = cost + (u64)f * cost_stride;
        for (int k = 0; k < AD_NCAND; k++) {
            if (!ad_cand_valid(w, h, k)) { if (tid == 0) s_total[k] = ~0ull; continue; }
            const u64 nb = ad_nblocks(w, h, 8ull << k);
            const u32 *t = ctab + ad_kbase(w, h, k);
            u64 acc = 0;
            for (u64 i = tid; i < nb; i += 256) acc += t[i] & 0x7fffffffu;
            for (int d = 16; d > 0; d >>= 1) acc += shfl_xor64(acc, d);
            if (lane == 0) red[wid] = acc;
            syncthreads();
            if (tid == 0) {
                u64 tot = 24 + (nb + 7) / 8;               // header + direction bytes
                for (int i = 0; i < NW; i++) tot += red[i];
                s_total[k] = tot;
            }
            syncthreads();
        }
        if (tid == 0) {
            int best = 0;
            for (int k = 1; k < AD_NCAND; k++)
                if (s_total[k] != ~0ull && s_total[k] < s_total[best]) best = k;   // strict: ties keep smaller B
            s_best = best;
            out_len[f] = s_total[best];
            if (chosen_b) chosen_b[f] = 8ull << best;
            status[f] = 0;
        }
        syncthreads();
        const int kb = s_best;
        const u64 b = 8ull << kb;
        const u64 nb = ad_nblocks(w, h, b);
        const u32 *t = ctab + ad_kbase(w, h, kb);
        u8 *o = out + out_off[f];
        // header, big endian (src/headers.cpp:27-37)
        if (tid < 24) {
            u64 v = tid < 8 ? w : (tid < 16 ? h : b);
            o[tid] = (u8)(v >> (8 * (7 - (tid & 7))));
        }
        // direction bits: block 0 = MSB of the first byte, zero padded (src/headers.cpp:43-60)
        const u64 dir_bytes = (nb + 7) / 8;
        for (u64 i0 = (u64)wid * 32; i0 < nb; i0 += 256) {
            u64 i = i0 + lane;
            u32 m = ballot(i < nb && (t[i < nb ? i : 0] >> 31));
            u32 rev = brev(m);
            if (lane < 4) {
                u64 bi = i0 / 8 + lane;
                if (bi < dir_bytes) o[24 + bi] = (u8)(rev >> (24 - 8 * lane));
            }
        }
        // exclusive scan of the block sizes -> offset of each block's RLE stream (after header)
        u32 *bo = blk_off + (u64)f * off_stride;
        u32 carry = 0;
        for (u64 i0 = 0; i0 < nb; i0 += 256) {
            u64 i = i0 + tid;
            u32 v = i < nb ? (t[i] & 0x7fffffffu) : 0u;
            u32 inc = v;
            for (int d = 1; d < 32; d <<= 1) {
                u32 x = shfl_up(inc, d);
                if (lane >= (u32)d) inc += x;
            }
            if (lane == 31) s_wsum[wid] = inc;
            syncthreads();
            u32 wbase = 0, tot = 0;
            for (u32 j = 0; j < NW; j++) { if (j < wid) wbase += s_wsum[j]; tot += s_wsum[j]; }
            if (i < nb) bo[i] = carry + wbase + inc - v;
            carry += tot;
            syncthreads();
        }
    }
}

constexpr int AD_EMIT_TPB = 256;

HC_KERNEL HC_LAUNCH_BOUNDS(AD_EMIT_TPB, 4)
adapt_emit_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT width,
                  const u64 *HC_RESTRICT height, u32 nf, const u32 *HC_RESTRICT cost, u64 cost_stride,
                  const u32 *HC_RESTRICT blk_off, u64 off_stride, const u64 *HC_RESTRICT chosen_b,
                  u8 *HC_RESTRICT out, const u64 *HC_RESTRICT out_off, const i32 *HC_RESTRICT status, bool skip_small,
                  u64 large_max)
{
    const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    for (u32 f = blockIdx.y; f < nf; f += gridDim.y) {
        if (status[f] != 0) continue;
        const u64 w = width[f], h = height[f], b = chosen_b[f];
        if (skip_small && ads_rows_per_group_fwd(w, b)) continue;     // adapt_emit_small_kernel
        if (adl_eligible(w, h, b) && w * h <= large_max) continue;    // adapt_emit_large_kernel
        int k = 0;
        while ((8ull << k) < b) k++;
        const u32 L = ad_lanes(b), G = 32u / L, gl = lane % L, gi = lane / L;
        const u8 *mat = in + in_off[f];
        const u32 *tab = cost + (u64)f * cost_stride + ad_kbase(w, h, k);
        const u32 *bo = blk_off + (u64)f * off_stride;
        const u64 nb = ad_nblocks(w, h, b);
        u8 *data = out + out_off[f] + 24 + (nb + 7) / 8;
        const u64 ngroups = (nb + G - 1) / G;
        const u64 stride = (u64)gridDim.x * (AD_EMIT_TPB / 32);
        for (u64 grp = (u64)blockIdx.x * (AD_EMIT_TPB / 32) + wid; grp < ngroups; grp += stride) {
            const u64 blk = grp * G + gi;
            const bool live = blk < nb;
            BlockGeom g = ad_geom(w, h, b, live ? blk : 0);
            const u32 n = live ? g.bw * g.bh : 0u;
            const bool hor = live ? (tab[blk] >> 31) : true;
            const u32 per = (n + L - 1) / L;
            u32 lo = gl * per, hi = lo + per;
            if (lo > n) lo = n;
            if (hi > n) hi = n;
            // inclusive scan of the lane chunks' RunSums inside the group, then shift
            RunSum inc = ad_chunk_sum(mat, w, g, hor, lo, hi);
            for (u32 d = 1; d < L; d <<= 1) {
                RunSum o = rs_shfl_up(inc, d);
                if (gl >= d) inc = rs_combine(o, inc);
            }
            RunSum pre = rs_shfl_up(inc, 1);
            if (gl == 0) pre = rs_empty();
            if (lo < hi) {
                BlockCursor c;
                bc_seek(c, mat, w, g, hor, lo);
                u32 curb = bc_next(c);
                // run index of the first element of this chunk and bytes emitted before it
                const bool is_final0 = (lo == n - 1);
                const bool cont = rs_nonempty(pre) && rs_last(pre) == curb && !is_final0;
                u32 kidx = cont ? pre.tail : 0u;
                u8 *op = data + bo[blk] + rs_emitted_prefix(pre, !cont);
                u32 q = kidx % 258u;
                for (u32 i = lo; i < hi; i++) {
                    // look one element ahead (may belong to the next lane's chunk)
                    u32 nextb = 0x100u;
                    if (i + 1 < n) nextb = bc_next(c);
                    const bool next_cont = (nextb == curb) && (i + 1 != n - 1);
                    if (q < 3u) *op++ = (u8)curb;
                    if (q == 257u) *op++ = 255;
                    else if (!next_cont && q >= 2u && i != n - 1) *op++ = (u8)(q - 2u);
                    q = next_cont ? (q == 257u ? 0u : q + 1u) : 0u;
                    curb = nextb;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// decode.  Block boundaries depend on the decoded data itself (fresh RLE state per block, a
// block ends when bw*bh bytes exist), so decoding is split in two:
//   adapt_index_kernel   one warp per file walks the token stream 32 tokens per step (decoder
//                        state maps and token lengths by warp scans, as in rle.cuh) and records
//                        where every block starts; validates the stream (status 10..15).
//   adapt_expand_kernel  a group of lanes per block decodes its tokens (state-map scan, length
//                        scan, then expansion) and scatters the bytes in the block's direction.
// adapt_decode_kernel (one thread per file) remains as the fallback for headers whose block
// count exceeds the index table (block sizes < 8, which the reference encoder never emits).
// ------------------------------------------------------------------------------------------
struct AdaptHeader { u64 w, h, b, nb, dir_bytes, total; i32 status; };

// src/headers.cpp:65-105 + the allocation of src/transform.cpp:340
HC_DEV AdaptHeader ad_parse_header(const u8 *src, u64 m)
{
    AdaptHeader hd;
    hd.w = hd.h = hd.b = hd.nb = hd.dir_bytes = hd.total = 0;
    hd.status = 0;
    if (m < 24) { hd.status = 10; return hd; }            // src/headers.cpp:67-71
    for (int i = 0; i < 8; i++) hd.w = (hd.w << 8) | src[i];
    for (int i = 0; i < 8; i++) hd.h = (hd.h << 8) | src[8 + i];
    for (int i = 0; i < 8; i++) hd.b = (hd.b << 8) | src[16 + i];
    if (hd.b == 0) { hd.status = 10; return hd; }         // reference: division by zero (UB)
    const u64 nbx = hd.w / hd.b + (hd.w % hd.b != 0), nby = hd.h / hd.b + (hd.h % hd.b != 0);
    // direction bytes are read lazily by the reference; missing ones -> 11 (src/headers.cpp:94-98)
    const u64 avail_bits = (m - 24) * 8;
    if ((nbx != 0 && nby != 0) && (nbx > avail_bits || nby > avail_bits || nbx * nby > avail_bits)) {
        hd.status = 11;
        return hd;
    }
    hd.nb = nbx * nby;
    hd.dir_bytes = (hd.nb + 7) / 8;
    if (hd.h != 0 && hd.w > ~0ull / hd.h) { hd.status = 100; return hd; }
    hd.total = hd.w * hd.h;
    return hd;
}

constexpr i32 AD_ST_SERIAL = 102;   // internal: block table too small, use the serial decoder

// Which of the two index kernels walks a file: big matrices with big blocks (the 4096 x 4096 images of BASELINE
// config 4: one block holds up to a million tokens) go to the CTA-wide kernel, everything else -- blocks of a few
// dozen tokens, where a step ends at the first block boundary whatever its width -- to the one-warp kernel
// (measured on the 512 x 512 batch: one warp 6.2 ms, eight cooperating warps 37 ms).
constexpr u32 AD_WIN_BYTES = 2048;       // token window of the one-warp index kernel (a multiple of 512)

HC_DEV bool ad_index_wide(const AdaptHeader &hd, u64 wide_min) { return hd.b >= 64u && hd.total >= wide_min; }

HC_KERNEL HC_LAUNCH_BOUNDS(32, 1)
adapt_index_warp_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT in_len,
                   const u64 *HC_RESTRICT out_cap, bool have_out, u32 *HC_RESTRICT blk_start, u64 blk_stride,
                   u64 *HC_RESTRICT out_len, i32 *HC_RESTRICT status, u32 nf, u64 wide_min)
{
    const u32 f = blockIdx.x, lane = threadIdx.x & 31u;
    if (f >= nf) return;
    const u64 m = in_len[f];
    const u8 *src = in + in_off[f];
    const AdaptHeader hd = ad_parse_header(src, m);
    if (hd.status) { if (lane == 0) { out_len[f] = 0; status[f] = hd.status; } return; }
    if (ad_index_wide(hd, wide_min)) return;           // walked by adapt_index_cta_kernel
    // A token yields at most 255 bytes, so a header that promises more than 255 bytes per payload byte
    // cannot be satisfied (a crafted w = h = 2^31 header would otherwise make the caller size its output
    // for 2^62 bytes): such a stream is only walked for its exact error (13 or 14), nothing is sized or written.
    const bool hopeless = hd.total > 255u * (m - 24u - hd.dir_bytes);
    if (!have_out && !hopeless) { if (lane == 0) { out_len[f] = hd.total; status[f] = 0; } return; }
    if (!hopeless) {
        if (hd.total > out_cap[f]) { if (lane == 0) { out_len[f] = hd.total; status[f] = 100; } return; }
        if (hd.nb + 1 > blk_stride || m > 0xfffffff0ull) { if (lane == 0) { out_len[f] = hd.total; status[f] = AD_ST_SERIAL; } return; }
    }
    u32 *tab = hopeless ? nullptr : blk_start + (u64)f * blk_stride;
    u64 pos = 24 + hd.dir_bytes;
    i32 err = 0;
    u64 blk = 0;
    // The walk advances a few dozen tokens per step (one small block) but looks at 128: the tokens are served from a
    // 2 KiB shared-memory window that is refilled with coalesced 16-byte loads when the step's range leaves it, so
    // that a step does not wait for global memory.  Decoder state maps travel as indices into the 40-element monoid
    // (rle.cuh: MapTables): one table lookup composes two maps, classifies a lane's four tokens, applies a map.
    HC_SHARED HC_ALIGNED16 u8 win[AD_WIN_BYTES + 16];
    HC_SHARED HC_ALIGNED16 MapTables mt;
    HC_SMEM_ARENA(win);
    const u32 wina = smem_addr(win);
    {
        const uint4 *g = (const uint4 *)&g_map_tables;
        uint4 *d = (uint4 *)&mt;
        for (u32 i = lane; i < sizeof(MapTables) / 16u; i += 32u) d[i] = g[i];
        syncwarp();
    }
    const u32 t_comp = smem_addr(mt.comp), t_bytes = smem_addr(mt.bytes), t_lane4 = smem_addr(mt.lane4), t_cls4 = smem_addr(mt.cls4);
    const u32 m_id = mt.id, m_eq = mt.eq, m_ne = mt.ne;
    u64 wlo = 0, whi = 0;                          // file positions [wlo, whi) held in the window
    // (crafted headers: block sides beyond 32 bits saturate the block size, offsets that would wrap end the loops)
    for (u64 by = 0; by < hd.h && !err; by = by + hd.b < by ? hd.h : by + hd.b) {
        const u64 bh = hd.h - by < hd.b ? hd.h - by : hd.b;
        for (u64 bx = 0; bx < hd.w && !err; bx = bx + hd.b < bx ? hd.w : bx + hd.b, blk++) {
            const u64 bw = hd.w - bx < hd.b ? hd.w - bx : hd.b;
            const u64 req = (bw >> 32 || bh >> 32) ? ~0ull : bw * bh;
            if (lane == 0 && tab) tab[blk] = (u32)pos;
            u64 produced = 0;
            u32 state = 0, prev_last = 0;
            while (produced < req) {
                // 128 tokens per step: 4 consecutive bytes per lane
                if (pos < wlo || pos + 128u > whi) {
                    syncwarp();                            // every lane is done with the old window
                    wlo = pos & ~(u64)15;
                    whi = wlo + AD_WIN_BYTES;
#pragma unroll
                    for (u32 j = 0; j < AD_WIN_BYTES / 512u; j++) {
                        const u64 o = wlo + 16u * (lane + 32u * j);
                        // the 16 bytes that hold the last token lie inside the file's padded region; beyond them: zeros
                        const uint4 v = o < m ? ldg16(src + o) : make_uint4_zero();
                        sts32(wina + 16u * (lane + 32u * j), v.x);
                        sts32(wina + 16u * (lane + 32u * j) + 4u, v.y);
                        sts32(wina + 16u * (lane + 32u * j) + 8u, v.z);
                        sts32(wina + 16u * (lane + 32u * j) + 12u, v.w);
                    }
                    syncwarp();
                }
                const u32 wrel = (u32)(pos - wlo);                       // window offset of the step's first token
                const u32 rel = wrel + 4u * lane;
                const u32 x = funnel_r(lds32(wina + (rel & ~3u)), lds32(wina + (rel & ~3u) + 4u), 8u * (rel & 3u));   // byte k = token k
                const u64 left = m - pos;
                const u32 av = left > 128u ? 128u : (u32)left;           // valid tokens of the step
                const i32 nvs = (i32)av - 4 * (i32)lane;
                const u32 nv = nvs <= 0 ? 0u : (nvs >= 4 ? 4u : (u32)nvs);
                u32 pb = shfl_up(x >> 24, 1);
                if (lane == 0) pb = prev_last;
                // equality bits of the lane's tokens, their state map, the maps of all tokens up to the lane
                const u32 e4 = dp4a_u(rle_zero_bytes(x ^ ((x << 8) | pb)), 0x08040201u, 0u) >> 7;
                u32 mi;
                if (nv == 4u) {
                    mi = lds8(t_lane4 + e4);
                } else {
                    mi = m_id;
                    for (u32 k = 0; k < nv; k++) mi = lds8(t_comp + 64u * mi + (((e4 >> k) & 1u) ? m_eq : m_ne));
                }
                u32 inc = mi;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const u32 t = shfl_up(inc, d);
                    const u32 c = lds8(t_comp + 64u * (lane >= (u32)d ? t : m_id) + inc);
                    inc = c;
                }
                u32 exm = shfl_up(inc, 1);
                if (lane == 0) exm = m_id;
                // which tokens are counts, output length of every token (one byte each: a count is at most 255)
                const u32 s = (lds32(t_bytes + 4u * exm) >> (8u * state)) & 3u;
                u32 cm4;
                if (nv == 4u) {
                    cm4 = lds8(t_cls4 + 16u * s + e4) & 15u;
                } else {
                    cm4 = 0;
                    u32 st = s;
                    for (u32 k = 0; k < nv; k++) {
                        if (st == 3u) { cm4 |= 1u << k; st = 0; }
                        else st = (st == 0u) ? 1u : (((e4 >> k) & 1u) ? st + 1u : 1u);
                    }
                }
                const u32 lens = (x & spread4(cm4)) | (0x01010101u & spread4(((1u << nv) - 1u) & ~cm4));
                const u32 sum = dp4a_u(lens, 0x01010101u, 0u);
                u32 cum = sum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const u32 t = shfl_up(cum, d);
                    if (lane >= (u32)d) cum += t;
                }
                const u64 rem = req - produced;
                const u32 rem32 = rem > 0xffffffffull ? 0xffffffffu : (u32)rem;   // (a step yields less than 2^15 bytes)
                const u32 hit = ballot(nv != 0u && cum >= rem32);
                if (hit == 0u) {
                    // the step did not complete the block: consume every valid token
                    if (av == 0u) { err = 14; break; }                  // src/transform.cpp:170-174
                    produced += shfl(cum, 31);
                    state = (lds32(t_bytes + 4u * shfl(inc, 31)) >> (8u * state)) & 3u;
                    prev_last = lds8(wina + wrel + av - 1u);            // the last consumed token
                    pos += av;
                    if (av < 128u && produced < req) { err = 14; break; }
                } else {
                    // the block ends inside lane `hl`: find the token
                    const int hl = ffs(hit) - 1;
                    const u32 before = shfl(cum - sum, hl);               // produced by the lanes before hl
                    const u32 lh = shfl(lens, hl);
                    const u32 l0 = lh & 0xffu, l1 = (lh >> 8) & 0xffu, l2 = (lh >> 16) & 0xffu, l3 = lh >> 24;
                    const u32 need = rem32 - before;                      // still missing when lane hl starts (>= 1)
                    u32 tok, got;
                    if (l0 >= need) { tok = 0; got = l0; }
                    else if (l0 + l1 >= need) { tok = 1; got = l0 + l1; }
                    else if (l0 + l1 + l2 >= need) { tok = 2; got = l0 + l1 + l2; }
                    else { tok = 3; got = l0 + l1 + l2 + l3; }
                    if (got > need) { err = 13; break; }                 // src/transform.cpp:180-184
                    pos += 4u * (u32)hl + tok + 1u;
                    produced = req;
                }
            }
        }
    }
    if (lane == 0) {
        if (tab) tab[hd.nb] = (u32)pos;
        out_len[f] = hopeless ? 0 : hd.total;
        status[f] = err ? err : (pos != m ? 15 : 0);                    // src/transform.cpp:354-358
    }
}

constexpr u32 AD_IDX_WARPS = 8;          // warps of a CTA that walk one file's token stream together

// One CTA per file.  A step covers 8 x 128 tokens: every warp takes 128 consecutive tokens (4 per lane),
// scans its decoder state maps and token lengths with shuffles, the per-warp totals are combined through
// shared memory, and the first token that completes the current block ends the step there (the decoder
// state resets at a block boundary, so nothing behind it can be classified yet).  A 1024 x 1024 block of a
// 4096 x 4096 image advances 1024 tokens per step instead of 128.
HC_KERNEL HC_LAUNCH_BOUNDS(AD_IDX_WARPS * 32, 1)
adapt_index_cta_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT in_len,
                       const u64 *HC_RESTRICT out_cap, bool have_out, u32 *HC_RESTRICT blk_start, u64 blk_stride,
                       u64 *HC_RESTRICT out_len, i32 *HC_RESTRICT status, u32 nf, u64 wide_min)
{
    HC_SHARED u32 s_map[2][AD_IDX_WARPS], s_sum[2][AD_IDX_WARPS], s_tv[2][AD_IDX_WARPS], s_last[2][AD_IDX_WARPS], s_hit[2][AD_IDX_WARPS];
    const u32 f = blockIdx.x, lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    if (f >= nf) return;
    const u64 m = in_len[f];
    const u8 *src = in + in_off[f];
    const AdaptHeader hd = ad_parse_header(src, m);
    if (hd.status || !ad_index_wide(hd, wide_min)) return;   // walked (or rejected) by adapt_index_warp_kernel
    // A token yields at most 255 bytes, so a header that promises more than 255 bytes per payload byte
    // cannot be satisfied (a crafted w = h = 2^31 header would otherwise make the caller size its output
    // for 2^62 bytes): such a stream is only walked for its exact error (13 or 14), nothing is sized or written.
    const bool hopeless = hd.total > 255u * (m - 24u - hd.dir_bytes);
    if (!have_out && !hopeless) { if (threadIdx.x == 0) { out_len[f] = hd.total; status[f] = 0; } return; }
    if (!hopeless) {
        if (hd.total > out_cap[f]) { if (threadIdx.x == 0) { out_len[f] = hd.total; status[f] = 100; } return; }
        if (hd.nb + 1 > blk_stride || m > 0xfffffff0ull) { if (threadIdx.x == 0) { out_len[f] = hd.total; status[f] = AD_ST_SERIAL; } return; }
    }
    u32 *tab = hopeless ? nullptr : blk_start + (u64)f * blk_stride;
    const u32 nw = AD_IDX_WARPS;
#define AD_IDX_SYNC() syncthreads()
    u64 pos = 24 + hd.dir_bytes;
    i32 err = 0;
    u64 blk = 0;
    u32 par = 0;                                  // double-buffered exchange arrays: one barrier per step
    // (crafted headers: block sides beyond 32 bits saturate the block size, offsets that would wrap end the loops)
    for (u64 by = 0; by < hd.h && !err; by = by + hd.b < by ? hd.h : by + hd.b) {
        const u64 bh = hd.h - by < hd.b ? hd.h - by : hd.b;
        for (u64 bx = 0; bx < hd.w && !err; bx = bx + hd.b < bx ? hd.w : bx + hd.b, blk++) {
            const u64 bw = hd.w - bx < hd.b ? hd.w - bx : hd.b;
            const u64 req = (bw >> 32 || bh >> 32) ? ~0ull : bw * bh;
            if (threadIdx.x == 0 && tab) tab[blk] = (u32)pos;
            u64 produced = 0;
            u32 state = 0, prev_last = 0;
            while (produced < req) {
                // 128 tokens per warp and step: 4 consecutive bytes per lane
                const u64 idx = pos + 128u * wid + 4u * lane;
                u32 b[4], nv = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const bool v = idx + k < m;
                    b[k] = v ? src[idx + k] : 0u;
                    nv += v ? 1u : 0u;
                }
                u32 pb = shfl_up(b[3], 1);
                if (lane == 0) pb = wid == 0 ? prev_last : (idx - 1 < m ? src[idx - 1] : 0u);
                // decoder state map of the lane's (valid) bytes
                u32 map = MAP_ID, pr = pb;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if ((u32)k < nv) map = map_compose(map, b[k] == pr ? MAP_EQ : MAP_NE);
                    pr = b[k];
                }
                u32 inc = map;
                for (int d = 1; d < 32; d <<= 1) {
                    u32 t = shfl_up(inc, d);
                    if (lane >= (u32)d) inc = map_compose(t, inc);
                }
                u32 exm = shfl_up(inc, 1);
                if (lane == 0) exm = MAP_ID;
                // entry state of this warp = the maps of the warps before it applied to the step's entry state;
                // the token lengths need it, so the maps are exchanged first
                if (lane == 31) s_map[par][wid] = inc;
                AD_IDX_SYNC();
                u32 wstate = state;
                for (u32 w = 0; w < wid; w++) wstate = map_apply(s_map[par][w], wstate);
                // output length of each of the lane's tokens
                u32 s = map_apply(exm, wstate), len[4], sum = 0;
                pr = pb;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    len[k] = 0;
                    if ((u32)k < nv) {
                        if (s == 3u) { len[k] = b[k]; s = 0; }
                        else { len[k] = 1; s = (s == 0u) ? 1u : (b[k] == pr ? s + 1u : 1u); }
                    }
                    pr = b[k];
                    sum += len[k];
                }
                u32 cum = sum;
                for (int d = 1; d < 32; d <<= 1) {
                    u32 t = shfl_up(cum, d);
                    if (lane >= (u32)d) cum += t;
                }
                u32 tv = nv;
                for (int d = 16; d > 0; d >>= 1) tv += shfl_xor(tv, d);
                if (lane == 31) { s_sum[par][wid] = cum; s_tv[par][wid] = tv; }
                {
                    // byte value of the warp's last valid token
                    const u32 slot = (tv - 1u) & 3u;
                    const u32 mine = slot == 0 ? b[0] : slot == 1 ? b[1] : slot == 2 ? b[2] : b[3];
                    const u32 lastv = shfl(mine, (int)((tv ? tv - 1u : 0u) >> 2));
                    if (lane == 0) s_last[par][wid] = lastv;
                }
                AD_IDX_SYNC();
                u64 before_w = 0;                 // bytes produced by the warps before this one
                for (u32 w = 0; w < wid; w++) before_w += s_sum[par][w];
                const u64 rem = req - produced;
                const u32 hit = before_w < rem ? ballot(nv != 0u && before_w + (u64)cum >= rem) : 0u;
                if (lane == 0) s_hit[par][wid] = hit;
                AD_IDX_SYNC();
                u32 hw = nw;                      // first warp in which the block ends
                for (u32 w = 0; w < nw; w++)
                    if (s_hit[par][w] && hw == nw) hw = w;
                if (hw == nw) {
                    // the step did not complete the block: consume every valid token of all warps
                    u32 tvall = 0;
                    u64 sumall = 0;
                    u32 st = state, lastv = prev_last;
                    for (u32 w = 0; w < nw; w++) {
                        if (s_tv[par][w]) lastv = s_last[par][w];
                        tvall += s_tv[par][w];
                        sumall += s_sum[par][w];
                        st = map_apply(s_map[par][w], st);
                    }
                    if (tvall == 0u) { err = 14; break; }                  // src/transform.cpp:170-174
                    produced += sumall;
                    state = st;
                    prev_last = lastv;
                    pos += tvall;
                    if (tvall < 128u * nw && produced < req) { err = 14; break; }
                } else {
                    // the block ends inside warp hw, lane hl: that warp finds the token and publishes the result
                    if (wid == hw) {
                        const int hl = ffs(hit) - 1;
                        const u32 before = shfl(cum - sum, hl);            // produced by the lanes before hl
                        const u32 l0 = shfl(len[0], hl), l1 = shfl(len[1], hl), l2 = shfl(len[2], hl), l3 = shfl(len[3], hl);
                        const u64 need = rem - before_w - before;          // still missing when lane hl starts (>= 1)
                        u32 tok, got;
                        if ((u64)l0 >= need) { tok = 0; got = l0; }
                        else if ((u64)l0 + l1 >= need) { tok = 1; got = l0 + l1; }
                        else if ((u64)l0 + l1 + l2 >= need) { tok = 2; got = l0 + l1 + l2; }
                        else { tok = 3; got = l0 + l1 + l2 + l3; }
                        if (lane == 0) {
                            s_sum[par ^ 1u][0] = (u64)got > need ? 1u : 0u;            // overshoot: src/transform.cpp:180-184
                            s_tv[par ^ 1u][0] = 128u * hw + 4u * (u32)hl + tok + 1u;      // tokens consumed by the block's end
                        }
                    }
                    AD_IDX_SYNC();
                    if (s_sum[par ^ 1u][0]) { err = 13; break; }
                    pos += s_tv[par ^ 1u][0];
                    produced = req;
                    AD_IDX_SYNC();                // the scratch words are reused by the next step's exchange
                }
                par ^= 1u;
            }
        }
    }
    if (threadIdx.x == 0) {
        if (tab) tab[hd.nb] = (u32)pos;
        out_len[f] = hopeless ? 0 : hd.total;
        status[f] = err ? err : (pos != m ? 15 : 0);                    // src/transform.cpp:354-358
    }
#undef AD_IDX_SYNC
}

constexpr int AD_EXP_TPB = 256;

HC_KERNEL HC_LAUNCH_BOUNDS(AD_EXP_TPB, 4)
adapt_expand_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT in_len,
                    const u32 *HC_RESTRICT blk_start, u64 blk_stride, u8 *HC_RESTRICT out,
                    const u64 *HC_RESTRICT out_off, const i32 *HC_RESTRICT status, u32 nf, bool skip_small,
                    u64 large_max)
{
    const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    for (u32 f = blockIdx.y; f < nf; f += gridDim.y) {
        if (status[f] != 0) continue;
        const u8 *src = in + in_off[f];
        const AdaptHeader hd = ad_parse_header(src, in_len[f]);
        if (hd.nb == 0) continue;
        if (skip_small && hd.w <= 0xffffffffull && hd.h <= 0xffffffffull && ads_rows_per_group_fwd(hd.w, hd.b)) continue;
        if (adl_eligible(hd.w, hd.h, hd.b) && hd.total <= large_max) continue;   // adapt_expand_large_kernel
        const u32 *tab = blk_start + (u64)f * blk_stride;
        u8 *mat = out + out_off[f];
        const u32 L = ad_lanes(hd.b), G = 32u / L, gl = lane % L, gi = lane / L;
        const u64 ngroups = (hd.nb + G - 1) / G;
        const u64 stride = (u64)gridDim.x * (AD_EXP_TPB / 32);
        for (u64 grp = (u64)blockIdx.x * (AD_EXP_TPB / 32) + wid; grp < ngroups; grp += stride) {
            const u64 blk = grp * G + gi;
            const bool live = blk < hd.nb;
            const BlockGeom g = ad_geom(hd.w, hd.h, hd.b, live ? blk : 0);
            const u32 t0 = live ? tab[blk] : 0u, t1 = live ? tab[blk + 1] : 0u;
            const u32 ntok = t1 - t0, per = (ntok + L - 1) / L;
            u32 lo = t0 + gl * per, hi = lo + per;
            if (lo > t1) lo = t1;
            if (hi > t1) hi = t1;
            const bool hor = live ? ((src[24 + (blk >> 3)] >> (7 - (blk & 7))) & 1u) : true;
            // pass A: decoder state map of this lane's token chunk
            u32 map = MAP_ID;
            {
                u32 prev = lo > t0 ? src[lo - 1] : 0x100u;
                for (u32 i = lo; i < hi; i++) {
                    u32 b = src[i];
                    map = map_compose(map, b == prev ? MAP_EQ : MAP_NE);
                    prev = b;
                }
            }
            u32 inc = map;
            for (u32 d = 1; d < L; d <<= 1) {
                u32 t = shfl_up(inc, d);
                if (gl >= d) inc = map_compose(t, inc);
            }
            u32 exm = shfl_up(inc, 1);
            if (gl == 0) exm = MAP_ID;
            const u32 st_in = map_apply(exm, 0u);                       // fresh state per block
            // pass B: output bytes produced by the chunk
            u32 cnt = 0;
            {
                u32 s = st_in, prev = lo > t0 ? src[lo - 1] : 0x100u;
                for (u32 i = lo; i < hi; i++) {
                    u32 b = src[i];
                    if (s == 3u) { cnt += b; s = 0; }
                    else { cnt += 1u; s = (s == 0u) ? 1u : (b == prev ? s + 1u : 1u); }
                    prev = b;
                }
            }
            u32 cum = cnt;
            for (u32 d = 1; d < L; d <<= 1) {
                u32 t = shfl_up(cum, d);
                if (gl >= d) cum += t;
            }
            // pass C: expand and scatter
            if (lo < hi) {
                const u32 req = g.bw * g.bh, inner = hor ? g.bw : g.bh;
                u32 o = cum - cnt;
                u32 co = o / inner, ci = o % inner;
                u8 *bp = mat + g.base;
                u32 s = st_in, prev = lo > t0 ? src[lo - 1] : 0u;
                for (u32 i = lo; i < hi; i++) {
                    u32 b = src[i], len, val;
                    if (s == 3u) { len = b; val = prev; s = 0; }
                    else { len = 1; val = b; s = (s == 0u) ? 1u : (b == prev ? s + 1u : 1u); }
                    prev = b;
                    for (u32 j = 0; j < len && o < req; j++, o++) {
                        u64 a = hor ? (u64)co * hd.w + ci : (u64)ci * hd.w + co;
                        bp[a] = (u8)val;
                        if (++ci == inner) { ci = 0; co++; }
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// serial fallback: one thread per file walks the token stream
// ------------------------------------------------------------------------------------------
HC_KERNEL HC_LAUNCH_BOUNDS(32, 1)
adapt_decode_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT in_len,
                    u8 *HC_RESTRICT out, const u64 *HC_RESTRICT out_off, const u64 *HC_RESTRICT out_cap,
                    u64 *HC_RESTRICT out_len, i32 *HC_RESTRICT status, u32 nf, i32 only_status)
{
    const u32 f = blockIdx.x;
    if (f >= nf || threadIdx.x != 0) return;
    if (only_status >= 0 && status[f] != only_status) return;   // fallback pass: selected files only
    const u64 m = in_len[f];
    const u8 *src = in + in_off[f];
    out_len[f] = 0;
    if (m < 24) { status[f] = 10; return; }               // src/headers.cpp:67-71
    u64 w = 0, h = 0, b = 0;
    for (int i = 0; i < 8; i++) w = (w << 8) | src[i];
    for (int i = 0; i < 8; i++) h = (h << 8) | src[8 + i];
    for (int i = 0; i < 8; i++) b = (b << 8) | src[16 + i];
    if (b == 0) { status[f] = 10; return; }               // reference: division by zero (UB)
    const u64 nbx = w / b + (w % b != 0), nby = h / b + (h % b != 0);
    // direction bytes are read lazily by the reference; missing ones -> 11 (src/headers.cpp:94-98)
    const u64 avail_bits = (m - 24) * 8;
    if ((nbx != 0 && nby != 0) && (nbx > avail_bits || nby > avail_bits || nbx * nby > avail_bits)) {
        status[f] = 11;
        return;
    }
    const u64 nb = nbx * nby;
    const u64 dir_bytes = (nb + 7) / 8;
    // w*h cannot overflow here: nb <= 8m blocks of at most b*b... guard explicitly anyway
    if (h != 0 && w > ~0ull / h) { status[f] = 100; return; }
    const u64 total = w * h;
    // more than 255 bytes per payload byte cannot be satisfied (see adapt_index_kernel): walk for the
    // exact error only, size and write nothing
    const bool hopeless = total > 255u * (m - 24u - dir_bytes);
    out_len[f] = hopeless ? 0 : total;
    if (!out && !hopeless) { status[f] = 0; return; }
    if (!hopeless && total > out_cap[f]) { status[f] = 100; return; }
    u8 *mat = hopeless ? nullptr : out + out_off[f];
    u64 pos = 24 + dir_bytes;
    for (u64 i = 0; i < nb; i++) {
        const bool hor = (src[24 + (i >> 3)] >> (7 - (i & 7))) & 1u;
        BlockGeom g = ad_geom(w, h, b, i);
        u64 req = (u64)g.bw * g.bh;
        if (hopeless) {
            // block sides beyond 32 bits (crafted headers only): saturate, the walk ends in 13 / 14 anyway
            const u64 bx = (i % nbx) * b, by = (i / nbx) * b;
            const u64 bw64 = bx + b > w ? w - bx : b, bh64 = by + b > h ? h - by : b;
            req = (bw64 >> 32 || bh64 >> 32) ? ~0ull : bw64 * bh64;
        }
        const u32 inner = hor ? g.bw : g.bh;
        u8 *bp = mat + g.base;
        u64 produced = 0;
        u32 ci = 0, co = 0;
        u32 match_byte = 0, match_count = 0;               // fresh state per block (:166-167)
        while (produced < req) {
            if (pos >= m) { status[f] = 14; return; }      // src/transform.cpp:170-174
            u32 cur = src[pos++];
            u32 len, val;
            if (match_count == 3) { len = cur; val = match_byte; match_count = 0; }
            else {
                len = 1; val = cur;
                if (match_byte == cur) match_count++; else { match_byte = cur; match_count = 1; }
            }
            if (produced + len > req) { status[f] = 13; return; }   // src/transform.cpp:180-184
            for (u32 j = 0; j < len && mat; j++) {
                u64 a = hor ? (u64)co * w + ci : (u64)ci * w + co;
                bp[a] = (u8)val;
                if (++ci == inner) { ci = 0; co++; }
            }
            produced += len;
        }
    }
    status[f] = pos != m ? 15 : 0;                         // src/transform.cpp:354-358
}

}  // namespace hcd
