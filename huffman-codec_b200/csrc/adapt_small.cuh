// adapt_small.cuh -- adaptive block RLE emit / expand for small blocks (B = 8, 16, 32).
//
// A CTA takes a group of whole block rows (a contiguous range of matrix rows = a contiguous range
// of memory), stages it in shared memory with 128-bit copies and lets ONE THREAD PER BLOCK run the
// MNP-5 state machine over its block in the block's scan direction (src/transform.cpp:241-279 for
// emit, :137-187 for expand).  The variable-length side (RLE bytes) of the group is contiguous in
// the stream as well (blocks are stored in raster order), so it is staged in shared memory too and
// moved with coalesced copies.  HBM traffic = algorithmic (every pixel and every RLE byte once).
// Larger blocks go through the lane-group kernels of adapt.cuh.
#pragma once
#include "adapt.cuh"

namespace hcd {

constexpr int ADS_TPB = 256;
constexpr u32 ADS_STRIP = 32 * 1024;                 // staged matrix rows per group
constexpr u32 ADS_RLE = ADS_STRIP + ADS_STRIP / 3 + 2048;   // RLE bytes of a group (4/3 bound + 1 per block) + phase
constexpr u32 ADS_SMEM = ADS_STRIP + ADS_RLE + 64;
constexpr u64 ADS_MAXB = 32;

// block rows per group: at most ADS_TPB blocks and ADS_STRIP staged bytes; 0 = not eligible
HC_HD u32 ads_rows_per_group(u64 w, u64 b)
{
    if (b > ADS_MAXB || b * w > ADS_STRIP) return 0;
    const u64 ncb = (w + b - 1) / b;
    u64 s = ADS_TPB / ncb;
    const u64 cap = ADS_STRIP / (b * w);
    if (s > cap) s = cap;
    return (u32)s;                                     // may be 0 when one block row has > 256 blocks
}

// coalesced copy global -> shared of n bytes (16-byte path when both sides allow it)
HC_DEV void ads_load(u8 *dst, const u8 *src, u32 n, u32 tid)
{
    if ((((uintptr_t)src) & 15u) == 0) {
        const u32 nv = n >> 4;
        for (u32 i = tid; i < nv; i += ADS_TPB) ((uint4 *)dst)[i] = ldg16(src + 16u * i);
        for (u32 i = (nv << 4) + tid; i < n; i += ADS_TPB) dst[i] = ldg8(src + i);
    } else {
        for (u32 i = tid; i < n; i += ADS_TPB) dst[i] = ldg8(src + i);
    }
}

// coalesced copy shared -> global of bytes [lo, hi) of a staging buffer whose byte i corresponds to
// global byte gbase + i, gbase 16-byte aligned
HC_DEV void ads_store(u8 *gbase, const u8 *stage, u32 lo, u32 hi, u32 tid)
{
    const u32 c0 = lo >> 4, c1 = (hi + 15u) >> 4;
    for (u32 c = c0 + tid; c < c1; c += ADS_TPB) {
        const u32 a = c << 4, b = a + 16u;
        if (a >= lo && b <= hi) {
            stg16(gbase + a, *(const uint4 *)(stage + a));
        } else {
            const u32 x = a < lo ? lo : a, y = b > hi ? hi : b;
            for (u32 i = x; i < y; i++) gbase[i] = stage[i];
        }
    }
}

HC_KERNEL HC_LAUNCH_BOUNDS(ADS_TPB, 2)
adapt_emit_small_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT width,
                        const u64 *HC_RESTRICT height, u32 nf, const u32 *HC_RESTRICT cost, u64 cost_stride,
                        const u32 *HC_RESTRICT blk_off, u64 off_stride, const u64 *HC_RESTRICT chosen_b,
                        u8 *HC_RESTRICT out, const u64 *HC_RESTRICT out_off, const i32 *HC_RESTRICT status)
{
    HC_DYN_SMEM(smem);
    u8 *strip = smem, *stage = smem + ADS_STRIP;
    const u32 tid = threadIdx.x;
    for (u32 f = blockIdx.y; f < nf; f += gridDim.y) {
        if (status[f] != 0) continue;
        const u64 w = width[f], h = height[f], b = chosen_b[f];
        const u32 S = ads_rows_per_group(w, b);
        if (S == 0) continue;                               // lane-group kernel handles this file
        int k = 0;
        while ((8ull << k) < b) k++;
        const u32 W = (u32)w, H = (u32)h, B = (u32)b;
        const u32 ncb = (W + B - 1) / B, nbr = (H + B - 1) / B, nb = ncb * nbr;
        const u8 *mat = in + in_off[f];
        const u32 *tab = cost + (u64)f * cost_stride + ad_kbase(w, h, k);
        const u32 *bo = blk_off + (u64)f * off_stride;
        u8 *data = out + out_off[f] + 24 + (nb + 7) / 8;
        for (u32 g = blockIdx.x; g * S < nbr; g += gridDim.x) {
            const u32 r0 = g * S, r1 = r0 + S < nbr ? r0 + S : nbr;
            const u32 y0 = r0 * B, y1 = r1 * B < H ? r1 * B : H;
            ads_load(strip, mat + (u64)y0 * W, (y1 - y0) * W, tid);
            const u32 blk0 = r0 * ncb, nblk = (r1 - r0) * ncb;
            const u32 o0 = bo[blk0];                          // stream offset of the group's first block
            const u32 phase = (u32)((uintptr_t)(data + o0) & 15u);
            syncthreads();
            u32 my_end = 0;
            if (tid < nblk) {
                const u32 blk = blk0 + tid, c = tid % ncb, rl = tid / ncb;
                const u32 bx = c * B, by = (r0 + rl) * B;
                const u32 bw = bx + B > W ? W - bx : B, bh = by + B > H ? H - by : B;
                const bool hor = tab[blk] >> 31;
                const u32 n = bw * bh, inner = hor ? bw : bh;
                const u8 *bp = strip + (by - y0) * W + bx;
                u8 *op = stage + phase + (bo[blk] - o0);
                // the reference encoder, one element at a time (src/transform.cpp:245-276)
                u32 mb = 0, mc = 0, ci = 0, co = 0;
                for (u32 i = 0; i < n; i++) {
                    const u32 cur = hor ? bp[co * W + ci] : bp[ci * W + co];
                    if (++ci == inner) { ci = 0; co++; }
                    if (cur == mb && mc != 0 && i + 1 != n) {
                        mc++;
                        if (mc <= 3) *op++ = (u8)cur;
                        else if (mc == 258) { *op++ = 255; mc = 0; }
                    } else {
                        if (mc >= 3) *op++ = (u8)(mc - 3);
                        *op++ = (u8)cur;
                        mb = cur;
                        mc = 1;
                    }
                }
                if (tid == nblk - 1) my_end = (u32)(op - stage);
            }
            // the last block's thread knows where the group's stream ends
            HC_SHARED u32 s_end;
            if (tid == nblk - 1) s_end = my_end;
            syncthreads();
            ads_store(data + o0 - phase, stage, phase, s_end, tid);
            syncthreads();
        }
    }
}

HC_KERNEL HC_LAUNCH_BOUNDS(ADS_TPB, 2)
adapt_expand_small_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT in_len,
                          const u32 *HC_RESTRICT blk_start, u64 blk_stride, u8 *HC_RESTRICT out,
                          const u64 *HC_RESTRICT out_off, const i32 *HC_RESTRICT status, u32 nf)
{
    HC_DYN_SMEM(smem);
    u8 *strip = smem, *stage = smem + ADS_STRIP;
    const u32 tid = threadIdx.x;
    for (u32 f = blockIdx.y; f < nf; f += gridDim.y) {
        if (status[f] != 0) continue;
        const u8 *src = in + in_off[f];
        const AdaptHeader hd = ad_parse_header(src, in_len[f]);
        if (hd.nb == 0 || hd.w > 0xffffffffull || hd.h > 0xffffffffull) continue;
        const u32 S = ads_rows_per_group(hd.w, hd.b);
        if (S == 0) continue;
        const u32 W = (u32)hd.w, H = (u32)hd.h, B = (u32)hd.b;
        const u32 ncb = (W + B - 1) / B, nbr = (H + B - 1) / B;
        const u32 *tab = blk_start + (u64)f * blk_stride;
        u8 *mat = out + out_off[f];
        for (u32 g = blockIdx.x; g * S < nbr; g += gridDim.x) {
            const u32 r0 = g * S, r1 = r0 + S < nbr ? r0 + S : nbr;
            const u32 y0 = r0 * B, y1 = r1 * B < H ? r1 * B : H;
            const u32 blk0 = r0 * ncb, nblk = (r1 - r0) * ncb;
            const u32 t0 = tab[blk0], t1 = tab[blk0 + nblk];
            const bool fits = t1 - t0 <= ADS_RLE - 64u;       // always true for streams the encoder produced
            const u32 phase = (u32)((uintptr_t)(src + t0) & 15u);
            if (fits) ads_load(stage, src + t0 - phase, phase + (t1 - t0), tid);
            syncthreads();
            if (tid < nblk) {
                const u32 blk = blk0 + tid, c = tid % ncb, rl = tid / ncb;
                const u32 bx = c * B, by = (r0 + rl) * B;
                const u32 bw = bx + B > W ? W - bx : B, bh = by + B > H ? H - by : B;
                const bool hor = (src[24 + (blk >> 3)] >> (7 - (blk & 7))) & 1u;
                const u32 req = bw * bh, inner = hor ? bw : bh;
                u8 *bp = strip + (by - y0) * W + bx;
                const u8 *tp = fits ? stage + phase + (tab[blk] - t0) : src + tab[blk];
                const u32 ntok = tab[blk + 1] - tab[blk];
                // the reference decoder step (src/transform.cpp:137-159), fresh state per block
                u32 mb = 0, mc = 0, produced = 0, ci = 0, co = 0;
                for (u32 i = 0; i < ntok && produced < req; i++) {
                    const u32 cur = tp[i];
                    u32 len, val;
                    if (mc == 3) { len = cur; val = mb; mc = 0; }
                    else { len = 1; val = cur; if (mb == cur) mc++; else { mb = cur; mc = 1; } }
                    for (u32 j = 0; j < len && produced < req; j++, produced++) {
                        bp[hor ? co * W + ci : ci * W + co] = (u8)val;
                        if (++ci == inner) { ci = 0; co++; }
                    }
                }
            }
            syncthreads();
            // rows y0..y1 are contiguous in the matrix
            u8 *g0 = mat + (u64)y0 * W;
            const u32 ph2 = (u32)((uintptr_t)g0 & 15u);
            if (ph2 == 0) ads_store(g0, strip, 0, (y1 - y0) * W, tid);
            else for (u32 i = tid; i < (y1 - y0) * W; i += ADS_TPB) g0[i] = strip[i];
            syncthreads();
        }
    }
}

}  // namespace hcd
