// diff.cuh -- differential model kernels (reference: src/transform.cpp:220-239).
//
//   apply : out[i] = in[i] - in[i-1]  (mod 256), in[-1] = 0      -- elementwise + 1-byte halo
//   revert: out[i] = sum_{k<=i} in[k] (mod 256)                  -- inclusive byte scan
//
// Both are HBM-bound: algorithmic traffic 2N bytes (read N, write N).  Each thread moves
// 16-byte vectors; a warp instruction covers 512 contiguous bytes.  The halo byte of the
// forward pass comes from the neighbouring lane by warp shuffle (lane 0 reads it from global).
#pragma once
#include "scan.cuh"

namespace hcd {

// ---------------------------------------------------------------- apply
// A CTA walks the tiles blockIdx.x, blockIdx.x + gridDim.x, ... of a file and keeps the loads of the
// next tile in flight while the current one is computed and stored.
HC_KERNEL HC_LAUNCH_BOUNDS(256, 4)
diff_apply_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, u8 *HC_RESTRICT out,
                  const u64 *HC_RESTRICT out_off, const u64 *HC_RESTRICT len, u32 nf)
{
    const u32 tid = threadIdx.x, lane = tid & 31;
    const u64 step = (u64)gridDim.x * TILE_BYTES;
    for (u32 f = blockIdx.y; f < nf; f += gridDim.y) {
        const u64 n = len[f];
        const u8 *src = in + in_off[f];
        u8 *dst = out + out_off[f];
        uint4 cur[UN], nxt[UN];
        u32 hcur[UN], hnxt[UN];                          // lane 0: the byte before its vector
        u64 t0 = (u64)blockIdx.x * TILE_BYTES;
#pragma unroll
        for (int j = 0; j < UN; j++) {
            const u64 p = t0 + (u64)j * SUB_BYTES + tid * 16;
            cur[j] = p < n ? ldg16(src + p) : make_uint4_zero();
            hcur[j] = (lane == 0 && p > 0 && p < n) ? ldg8(src + p - 1) : 0u;
        }
        for (; t0 < n; t0 += step) {
#pragma unroll
            for (int j = 0; j < UN; j++) {
                const u64 p = t0 + step + (u64)j * SUB_BYTES + tid * 16;
                nxt[j] = p < n ? ldg16(src + p) : make_uint4_zero();
                hnxt[j] = (lane == 0 && p < n) ? ldg8(src + p - 1) : 0u;
            }
#pragma unroll
            for (int j = 0; j < UN; j++) {
                const u64 p = t0 + (u64)j * SUB_BYTES + tid * 16;
                const uint4 v = cur[j];
                const u32 up = shfl_up(v.w >> 24, 1);   // last byte of the previous lane's vector
                const u32 prev = lane == 0 ? hcur[j] : up;
                uint4 r;
                r.x = vsub4(v.x, (v.x << 8) | prev);
                r.y = vsub4(v.y, (v.y << 8) | (v.x >> 24));
                r.z = vsub4(v.z, (v.z << 8) | (v.y >> 24));
                r.w = vsub4(v.w, (v.w << 8) | (v.z >> 24));
                if (p < n) {
                    if (p + 16 > n) r = mask_tail(r, (u32)(n - p));   // zero the padding bytes
                    stg16(dst + p, r);
                }
            }
#pragma unroll
            for (int j = 0; j < UN; j++) { cur[j] = nxt[j]; hcur[j] = hnxt[j]; }
        }
    }
}

// ---------------------------------------------------------------- revert
// byte-wise inclusive prefix inside one 16-byte vector; returns the vector total in bits 0..7
HC_DEV u32 prefix16(uint4 &v)
{
    u32 a = v.x;
    a = vadd4(a, a << 8);
    a = vadd4(a, a << 16);
    u32 b = v.y;
    b = vadd4(b, b << 8);
    b = vadd4(b, b << 16);
    b = vadd4(b, (a >> 24) * 0x01010101u);
    u32 c = v.z;
    c = vadd4(c, c << 8);
    c = vadd4(c, c << 16);
    c = vadd4(c, (b >> 24) * 0x01010101u);
    u32 d = v.w;
    d = vadd4(d, d << 8);
    d = vadd4(d, d << 16);
    d = vadd4(d, (c >> 24) * 0x01010101u);
    v.x = a; v.y = b; v.z = c; v.w = d;
    return d >> 24;
}

HC_DEV u32 bytesum16(const uint4 &v)
{
    // sum of 16 bytes mod 256
    u32 s = vadd4(vadd4(v.x, v.y), vadd4(v.z, v.w));
    s = vadd4(s, s >> 16);
    s = vadd4(s, s >> 8);
    return s & 0xffu;
}

// per-segment byte sums (only launched when a file is split into several segments)
HC_KERNEL HC_LAUNCH_BOUNDS(256, 4)
diff_segsum_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT len,
                   u32 nf, u32 nseg, u64 seg_bytes, u32 *HC_RESTRICT segsum)
{
    HC_SHARED u32 wsum[NW];
    const u32 tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (u32 f = blockIdx.y; f < nf; f += gridDim.y) {
        const u64 n = len[f];
        const u64 s0 = (u64)blockIdx.x * seg_bytes;
        u64 s1 = s0 + seg_bytes;
        if (s1 > n) s1 = n;
        const u8 *src = in + in_off[f];
        u32 acc = 0;
        for (u64 p = s0 + tid * 16; p < s1; p += (u64)TPB * 16) {
            uint4 v = ldg16(src + p);
            if (p + 16 > n) v = mask_tail(v, (u32)(n - p));
            acc += bytesum16(v);
        }
        for (int d = 16; d > 0; d >>= 1) acc += shfl_xor(acc, d);
        if (lane == 0) wsum[w] = acc;
        syncthreads();
        if (tid == 0) {
            u32 t = 0;
            for (int i = 0; i < NW; i++) t += wsum[i];
            segsum[(u64)f * nseg + blockIdx.x] = t & 0xffu;
        }
        syncthreads();
    }
}

HC_KERNEL HC_LAUNCH_BOUNDS(256, 3)
diff_revert_kernel(const u8 *in, const u64 *HC_RESTRICT in_off, u8 *out, const u64 *HC_RESTRICT out_off,
                   const u64 *HC_RESTRICT len, u32 nf, u32 nseg, u64 seg_bytes, const u32 *HC_RESTRICT segsum)
{
    HC_SHARED u32 wtot[2][32];
    const u32 tid = threadIdx.x;
    for (u32 f = blockIdx.y; f < nf; f += gridDim.y) {
        const u64 n = len[f];
        const u64 s0 = (u64)blockIdx.x * seg_bytes;
        if (s0 >= n) continue;
        u64 s1 = s0 + seg_bytes;
        if (s1 > n) s1 = n;
        const u8 *src = in + in_off[f];
        u8 *dst = out + out_off[f];
        u32 carry = 0;                               // sum of all bytes before this segment
        if (nseg > 1)
            for (u32 s = 0; s < blockIdx.x; s++) carry += segsum[(u64)f * nseg + s];
        carry &= 0xffu;

        uint4 cur[UN], nxt[UN];
#pragma unroll
        for (int j = 0; j < UN; j++) {
            u64 p = s0 + (u64)j * SUB_BYTES + tid * 16;
            cur[j] = p < s1 ? ldg16_rw(src + p) : make_uint4_zero();
        }
        u32 it = 0;
        for (u64 t0 = s0; t0 < s1; t0 += TILE_BYTES, it++) {
            // prefetch the next tile while this one is scanned
#pragma unroll
            for (int j = 0; j < UN; j++) {
                u64 p = t0 + TILE_BYTES + (u64)j * SUB_BYTES + tid * 16;
                nxt[j] = p < s1 ? ldg16_rw(src + p) : make_uint4_zero();
            }
            u32 tot[UN], excl[UN];
#pragma unroll
            for (int j = 0; j < UN; j++) {
                u64 p = t0 + (u64)j * SUB_BYTES + tid * 16;
                if (p < n && p + 16 > n) cur[j] = mask_tail(cur[j], (u32)(n - p));
                tot[j] = prefix16(cur[j]);
            }
            u32 total = block_scan_striped(tot, excl, 0u, OpAdd(), wtot[it & 1]);
#pragma unroll
            for (int j = 0; j < UN; j++) {
                u64 p = t0 + (u64)j * SUB_BYTES + tid * 16;
                if (p < s1) {
                    u32 add = ((carry + excl[j]) & 0xffu) * 0x01010101u;
                    uint4 r;
                    r.x = vadd4(cur[j].x, add);
                    r.y = vadd4(cur[j].y, add);
                    r.z = vadd4(cur[j].z, add);
                    r.w = vadd4(cur[j].w, add);
                    if (p + 16 > n) r = mask_tail(r, (u32)(n - p));
                    stg16(dst + p, r);
                }
            }
            carry = (carry + total) & 0xffu;
#pragma unroll
            for (int j = 0; j < UN; j++) cur[j] = nxt[j];
        }
        syncthreads();
    }
}

}  // namespace hcd
