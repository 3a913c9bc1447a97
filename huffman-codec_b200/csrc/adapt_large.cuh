// adapt_large.cuh -- adaptive block RLE emit / expand for large blocks (B >= 64).
//
// A block of B x B pixels in its scan direction is just an MNP-5 stream of B*B elements
// (src/transform.cpp:97-134 calls the same state machine as applyRLE), so large blocks reuse the
// CTA-wide streaming coder of rle.cuh.  Two kernels per direction of the codec:
//   encode:  adapt_gather_large_kernel   block pixels -> linear stream in scratch (row-major copy
//                                        for horizontal blocks, 64 x 64 shared-memory transpose
//                                        for vertical ones)
//            adapt_emit_large_kernel     one CTA per block: rle_encode_stream(scratch -> output at
//                                        the block's offset from adapt_select_kernel)
//   decode:  adapt_expand_large_kernel   one CTA per block: rle_decode_stream(tokens -> scratch);
//                                        block token ranges come from adapt_index_kernel
//            adapt_scatter_large_kernel  scratch -> matrix (inverse of the gather)
// The scratch holds every block's stream at a 16-byte aligned slot (closed-form offset, below).
// Extra HBM traffic: 2N per direction, all of it coalesced.
#pragma once
#include "adapt.cuh"

namespace hcd {

constexpr u32 ADL_T = 64;                      // tile edge of the gather / scatter
constexpr u32 ADL_TS = 68;                     // shared tile row stride (bytes)
constexpr int ADL_TPB = 256;

// bytes of scratch per file for matrices of at most max_len pixels
HC_HD u64 adl_tmp_stride(u64 max_len) { return (max_len + 255) & ~255ull; }

// offset of block blk's stream inside the file's scratch: blocks packed in raster order.  Every
// block before blk in its block row has b*bh bytes and every earlier block row w*b bytes, so with
// b % 16 == 0 (adl_eligible) all slots are 16-byte aligned and the scratch is exactly w*h bytes.
HC_HD u64 adl_slot(u64 w, u64 h, u64 b, u64 blk)
{
    const u64 ncb = (w + b - 1) / b, nbr = (h + b - 1) / b;
    const u64 r = blk / ncb, c = blk % ncb;
    const u64 bhr = r == nbr - 1 ? h - r * b : b;
    return r * b * w + c * b * bhr;
}

// copies an nl x ll (<= 64 x 64) tile: line i of the source is src + i*sstride.  Not transposed:
// line i of the destination is dst + i*dstride.  Transposed: byte j of source line i goes to byte i
// of destination line j.  All ADL_TPB threads of the CTA call it.
HC_DEV void adl_copy_tile(const u8 *HC_RESTRICT src, u64 sstride, u8 *HC_RESTRICT dst, u64 dstride, u32 nl, u32 ll,
                          bool transpose, u8 *tile)
{
    const u32 tid = threadIdx.x, line = tid >> 2, c0 = (tid & 3u) * 16u;
    if (line < nl && c0 < ll) {
        const u8 *p = src + (u64)line * sstride + c0;
        const u32 cnt = ll - c0 < 16u ? ll - c0 : 16u;
        u32 *tw = (u32 *)(tile + line * ADL_TS + c0);
        if (cnt == 16u && (((uintptr_t)p) & 15u) == 0) {
            const uint4 v = ldg16(p);
            tw[0] = v.x; tw[1] = v.y; tw[2] = v.z; tw[3] = v.w;
        } else {
            for (u32 i = 0; i < cnt; i++) tile[line * ADL_TS + c0 + i] = ldg8(p + i);
        }
    }
    syncthreads();
    const u32 dnl = transpose ? ll : nl, dll = transpose ? nl : ll;
    if (line < dnl && c0 < dll) {
        u8 *p = dst + (u64)line * dstride + c0;
        const u32 cnt = dll - c0 < 16u ? dll - c0 : 16u;
        u32 wv[4] = {0, 0, 0, 0};
        if (transpose) {
#pragma unroll
            for (u32 i = 0; i < 16u; i++)
                if (i < cnt) wv[i >> 2] |= (u32)tile[(c0 + i) * ADL_TS + line] << (8u * (i & 3u));
        } else {
            const u32 *tw = (const u32 *)(tile + line * ADL_TS + c0);
            wv[0] = tw[0]; wv[1] = tw[1]; wv[2] = tw[2]; wv[3] = tw[3];
        }
        if (cnt == 16u && (((uintptr_t)p) & 15u) == 0) {
            uint4 r; r.x = wv[0]; r.y = wv[1]; r.z = wv[2]; r.w = wv[3];
            stg16(p, r);
        } else {
            for (u32 i = 0; i < cnt; i++) p[i] = (u8)(wv[i >> 2] >> (8u * (i & 3u)));
        }
    }
    syncthreads();
}

// moves block blk between the matrix and its scratch slot, 64-row strip `strip` of the block
// to_stream: matrix -> stream order (gather); else stream -> matrix (scatter)
HC_DEV void adl_move_strip(u8 *mat, u64 w, const BlockGeom &g, bool hor, u8 *slot, u32 strip, bool to_stream, u8 *tile)
{
    const u32 y0 = strip * ADL_T;
    const u32 th = g.bh - y0 < ADL_T ? g.bh - y0 : ADL_T;
    for (u32 x0 = 0; x0 < g.bw; x0 += ADL_T) {
        const u32 tw = g.bw - x0 < ADL_T ? g.bw - x0 : ADL_T;
        u8 *mp = mat + g.base + (u64)y0 * w + x0;                            // th lines of tw bytes, stride w
        u8 *sp = hor ? slot + (u64)y0 * g.bw + x0 : slot + (u64)x0 * g.bh + y0;   // hor: th x tw stride bw; ver: tw x th stride bh
        const u64 ss = hor ? g.bw : g.bh;
        if (to_stream) adl_copy_tile(mp, w, sp, ss, th, tw, !hor, tile);
        else if (hor)  adl_copy_tile(sp, ss, mp, w, th, tw, false, tile);
        else           adl_copy_tile(sp, ss, mp, w, tw, th, true, tile);
    }
}

HC_KERNEL HC_LAUNCH_BOUNDS(ADL_TPB, 4)
adapt_gather_large_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT width,
                          const u64 *HC_RESTRICT height, u32 nf, const u32 *HC_RESTRICT cost, u64 cost_stride,
                          const u64 *HC_RESTRICT chosen_b, const i32 *HC_RESTRICT status, u8 *HC_RESTRICT tmp, u64 tstride)
{
    HC_SHARED HC_ALIGNED16 u8 tile[ADL_T * ADL_TS];
    for (u32 f = blockIdx.y; f < nf; f += gridDim.y) {
        if (status[f] != 0) continue;
        const u64 w = width[f], h = height[f], b = chosen_b[f];
        if (!adl_eligible(w, h, b) || w * h > tstride) continue;
        int k = 0;
        while ((8ull << k) < b) k++;
        const u32 *tab = cost + (u64)f * cost_stride + ad_kbase(w, h, k);
        const u64 nb = ad_nblocks(w, h, b);
        const u32 spb1 = (u32)(b / ADL_T) + ((b % ADL_T) ? 1u : 0u);   // 64-row strips per full block
        for (u64 it = blockIdx.x; it < nb * spb1; it += gridDim.x) {
            const u64 blk = it / spb1;
            const u32 strip = (u32)(it % spb1);
            const BlockGeom g = ad_geom(w, h, b, blk);
            if (strip * ADL_T >= g.bh) continue;
            const bool hor = tab[blk] >> 31;
            if (hor && g.bw == w) continue;                    // full-width horizontal block: already a stream
            adl_move_strip((u8 *)(in + in_off[f]), w, g, hor, tmp + (u64)f * tstride + adl_slot(w, h, b, blk),
                           strip, true, tile);
        }
    }
}

HC_KERNEL HC_LAUNCH_BOUNDS(RTPB, HC_RLE_ENC_MINB)
adapt_emit_large_kernel(const u8 *HC_RESTRICT tmp, u64 tstride, const u64 *HC_RESTRICT width, const u64 *HC_RESTRICT height,
                        u32 nf, const u32 *HC_RESTRICT blk_off, u64 off_stride, const u64 *HC_RESTRICT chosen_b,
                        u8 *HC_RESTRICT out, const u64 *HC_RESTRICT out_off, const i32 *HC_RESTRICT status,
                        const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u32 *HC_RESTRICT cost, u64 cost_stride)
{
    rle_enc_init();
    for (u32 f = blockIdx.y; f < nf; f += gridDim.y) {
        if (status[f] != 0) continue;
        const u64 w = width[f], h = height[f], b = chosen_b[f];
        if (!adl_eligible(w, h, b) || w * h > tstride) continue;
        const u32 *bo = blk_off + (u64)f * off_stride;
        const u64 nb = ad_nblocks(w, h, b);
        u8 *data = out + out_off[f] + 24 + (nb + 7) / 8;
        int k = 0;
        while ((8ull << k) < b) k++;
        const u32 *tab = cost + (u64)f * cost_stride + ad_kbase(w, h, k);
        for (u64 blk = blockIdx.x; blk < nb; blk += gridDim.x) {
            const BlockGeom g = ad_geom(w, h, b, blk);
            // a full-width horizontal block is contiguous in the matrix (16-byte aligned: the file slot is
            // 256-byte aligned and the block starts b rows down, b % 16 == 0): encode it in place
            const u8 *src = ((tab[blk] >> 31) && g.bw == w) ? in + in_off[f] + g.base
                                                            : tmp + (u64)f * tstride + adl_slot(w, h, b, blk);
            rle_encode_stream(src, (u64)g.bw * g.bh, data + bo[blk]);
        }
    }
}

HC_KERNEL HC_LAUNCH_BOUNDS(RTPB, HC_RLE_DEC_MINB)
adapt_expand_large_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT in_len,
                          const u32 *HC_RESTRICT blk_start, u64 blk_stride, const i32 *HC_RESTRICT status, u32 nf,
                          u8 *HC_RESTRICT tmp, u64 tstride, u8 *HC_RESTRICT out, const u64 *HC_RESTRICT out_off)
{
    rle_dec_init();
    for (u32 f = blockIdx.y; f < nf; f += gridDim.y) {
        if (status[f] != 0) continue;
        const u8 *src = in + in_off[f];
        const AdaptHeader hd = ad_parse_header(src, in_len[f]);
        if (hd.nb == 0 || !adl_eligible(hd.w, hd.h, hd.b) || hd.total > tstride) continue;
        const u32 *tab = blk_start + (u64)f * blk_stride;
        for (u64 blk = blockIdx.x; blk < hd.nb; blk += gridDim.x) {
            const BlockGeom g = ad_geom(hd.w, hd.h, hd.b, blk);
            const u32 t0 = tab[blk], t1 = tab[blk + 1];
            const bool hor = (src[24 + (blk >> 3)] >> (7 - (blk & 7))) & 1u;
            // a full-width horizontal block is decoded straight into the matrix (see adapt_emit_large_kernel)
            u8 *dst = (hor && g.bw == hd.w) ? out + out_off[f] + g.base : tmp + (u64)f * tstride + adl_slot(hd.w, hd.h, hd.b, blk);
            // adapt_index_kernel has validated the stream: these tokens decode to exactly bw*bh bytes
            rle_decode_stream(src + t0, (u64)(t1 - t0), dst, (u64)g.bw * g.bh);
        }
    }
}

HC_KERNEL HC_LAUNCH_BOUNDS(ADL_TPB, 4)
adapt_scatter_large_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT in_len,
                           const i32 *HC_RESTRICT status, u32 nf, const u8 *HC_RESTRICT tmp, u64 tstride,
                           u8 *HC_RESTRICT out, const u64 *HC_RESTRICT out_off)
{
    HC_SHARED HC_ALIGNED16 u8 tile[ADL_T * ADL_TS];
    for (u32 f = blockIdx.y; f < nf; f += gridDim.y) {
        if (status[f] != 0) continue;
        const u8 *src = in + in_off[f];
        const AdaptHeader hd = ad_parse_header(src, in_len[f]);
        if (hd.nb == 0 || !adl_eligible(hd.w, hd.h, hd.b) || hd.total > tstride) continue;
        const u32 spb1 = (u32)(hd.b / ADL_T) + ((hd.b % ADL_T) ? 1u : 0u);
        for (u64 it = blockIdx.x; it < hd.nb * spb1; it += gridDim.x) {
            const u64 blk = it / spb1;
            const u32 strip = (u32)(it % spb1);
            const BlockGeom g = ad_geom(hd.w, hd.h, hd.b, blk);
            if (strip * ADL_T >= g.bh) continue;
            const bool hor = (src[24 + (blk >> 3)] >> (7 - (blk & 7))) & 1u;
            if (hor && g.bw == hd.w) continue;                 // decoded in place by adapt_expand_large_kernel
            adl_move_strip(out + out_off[f], hd.w, g, hor, (u8 *)tmp + (u64)f * tstride + adl_slot(hd.w, hd.h, hd.b, blk),
                           strip, false, tile);
        }
    }
}

}  // namespace hcd
