// rle.cuh -- MNP-5 run-length encoding kernels (reference: src/transform.cpp:137-159, 241-292).
//
// One CTA streams one file tile by tile (16 KiB tiles, see scan.cuh) carrying a few scalars
// between tiles, so a batch of files needs no inter-CTA communication at all and every byte is
// read from HBM exactly once (algorithmic traffic N + M).
//
// ENCODE.  Per element: k = index inside its maximal run (runs are taken over elements
// 0..n-2, the last element is always its own literal), q = k mod 258.  The element emits
//      [q < 3] its byte, [q == 257] the byte 255, [last of run && 2 <= q < 257] the count q-2.
// k comes from a block-wide max-scan of run-start positions, the output position from a
// block-wide exclusive add-scan of the per-element byte counts (0, 1 or 2).  Output bytes are
// staged in shared memory with the same 16-byte phase as the global destination and copied
// out with 128-bit stores.
//
// DECODE.  Whether an input byte is a literal or a count depends on the decoder state
// c in {0,1,2,3}, whose transition only needs c and e[i] = (in[i] == in[i-1]).  Each position is
// therefore a 4->4 map (8 bits); a block-wide scan under function composition classifies
// every byte.  A second add-scan of the token lengths (literal 1, count = byte value) places
// the output.  Expansion is done per 16 KiB output window: every token drops its value and a
// head flag at its first output position, then each thread propagates the last head value
// over 64 consecutive output bytes (a max-scan finds the head that reaches into its range).
#pragma once
#include "runsum.cuh"
#include "scan.cuh"

namespace hcd {

constexpr u32 ENC_STAGE_BYTES = TILE_BYTES + TILE_BYTES / 3 + 64;   // worst case 4/3 + phase

HC_KERNEL HC_LAUNCH_BOUNDS(256, 2)
rle_encode_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT in_len,
                  u8 *HC_RESTRICT out, const u64 *HC_RESTRICT out_off, u64 *HC_RESTRICT out_len, u32 nf)
{
    HC_SHARED u32 wtot[2][32];
    HC_SHARED u32 s_carry_run;
    HC_SHARED HC_ALIGNED16 u8 sout[ENC_STAGE_BYTES];
    const u32 tid = threadIdx.x, lane = tid & 31;

    for (u32 f = blockIdx.x; f < nf; f += gridDim.x) {
        const u64 n = in_len[f];
        const u8 *src = in + in_off[f];
        u8 *dst = out + out_off[f];
        u64 out_pos = 0;      // bytes emitted by all previous tiles
        u32 carry_run = 0;    // length of the run that ends at the last element of the previous tile

        uint4 cur[UN], nxt[UN];
#pragma unroll
        for (int j = 0; j < UN; j++) {
            u64 p = (u64)j * SUB_BYTES + tid * 16;
            cur[j] = p < n ? ldg16(src + p) : make_uint4_zero();
        }
        for (u64 t0 = 0; t0 < n; t0 += TILE_BYTES) {
#pragma unroll
            for (int j = 0; j < UN; j++) {
                u64 p = t0 + TILE_BYTES + (u64)j * SUB_BYTES + tid * 16;
                nxt[j] = p < n ? ldg16(src + p) : make_uint4_zero();
            }
            // ---- equality bits and run starts ------------------------------------------
            u32 eq[UN];      // bit k (0..16): element k equals element k-1 (bit 16 = next thread's first)
            u32 valid[UN];   // bit k: element exists
            u32 smax[UN], sexcl[UN];
#pragma unroll
            for (int j = 0; j < UN; j++) {
                const u64 p = t0 + (u64)j * SUB_BYTES + tid * 16;
                u32 first = cur[j].x & 0xffu, last = cur[j].w >> 24;
                u32 pb = shfl_up(last, 1), nb = shfl_down(first, 1);
                if (lane == 0) pb = (p > 0 && p < n) ? ldg8(src + p - 1) : 0x100u;
                if (lane == 31) nb = (p + 16 < n) ? ldg8(src + p + 16) : 0x100u;
                u32 e = 0, prev = pb;
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    u32 b = vec_byte(cur[j], k);
                    if (b == prev) e |= 1u << k;
                    prev = b;
                }
                if (nb == prev) e |= 1u << 16;
                u32 vm = p >= n ? 0u : (n - p >= 17 ? 0x1ffffu : ((1u << (u32)(n - p)) - 1u));
                if (p == 0) e &= ~1u;
                // the last element of the file never continues a run (forced literal)
                if (n - 1 >= p && n - 1 - p <= 16) e &= ~(1u << (u32)(n - 1 - p));
                e &= vm;
                eq[j] = e;
                valid[j] = vm & 0xffffu;
                u32 starts = valid[j] & ~e;
                // tile-relative position + 1 of the last run start owned by this thread
                smax[j] = starts ? (u32)j * SUB_BYTES + tid * 16 + (31u - (u32)clz(starts)) + 1u : 0u;
            }
            block_scan_striped(smax, sexcl, 0u, OpMax(), wtot[0]);

            // ---- per-element output counts --------------------------------------------
            u32 cnt[UN], oexcl[UN];
            u32 qfirst[UN];   // q of element 0 of the thread's vector
#pragma unroll
            for (int j = 0; j < UN; j++) {
                const u32 tp = (u32)j * SUB_BYTES + tid * 16;   // tile-relative position
                u32 kidx = sexcl[j] ? tp - (sexcl[j] - 1u) : carry_run + tp;
                if (!(eq[j] & 1u)) kidx = 0;
                u32 q = kidx % 258u;
                qfirst[j] = q;
                u32 c = 0;
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    if (k > 0) q = ((eq[j] >> k) & 1u) ? (q == 257u ? 0u : q + 1u) : 0u;
                    bool v = (valid[j] >> k) & 1u;
                    bool is_end = !((eq[j] >> (k + 1)) & 1u);   // next element starts a run / is final / absent
                    u32 ck = (q < 3u ? 1u : 0u) + (q == 257u ? 1u : 0u) + ((is_end && q >= 2u && q != 257u) ? 1u : 0u);
                    // the final element of the file: exactly one literal (q == 0 there, is_end adds nothing)
                    c += v ? ck : 0u;
                }
                cnt[j] = c;
                if (j == UN - 1 && tid == TPB - 1) {
                    // run length at the end of a full tile: q-chain above ended with element 15
                    // recompute the true (non-modular) length for the carry
                    u32 kk = kidx;
                    for (int k = 1; k < 16; k++) kk = ((eq[j] >> k) & 1u) ? kk + 1u : 0u;
                    s_carry_run = kk + 1u;
                }
            }
            u32 total = block_scan_striped(cnt, oexcl, 0u, OpAdd(), wtot[1]);

            // ---- stage the output bytes -------------------------------------------------
            const u32 shift = (u32)(out_pos & 15u);
#pragma unroll
            for (int j = 0; j < UN; j++) {
                u32 o = shift + oexcl[j];
                u32 q = qfirst[j];
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    if (k > 0) q = ((eq[j] >> k) & 1u) ? (q == 257u ? 0u : q + 1u) : 0u;
                    if ((valid[j] >> k) & 1u) {
                        u32 b = vec_byte(cur[j], k);
                        bool is_end = !((eq[j] >> (k + 1)) & 1u);
                        if (q < 3u) sout[o++] = (u8)b;
                        if (q == 257u) sout[o++] = 255;
                        else if (is_end && q >= 2u) sout[o++] = (u8)(q - 2u);
                    }
                }
            }
            syncthreads();
            carry_run = s_carry_run;
            // ---- copy out: sout[shift .. shift+total) -> dst[out_pos ..) ---------------------
            {
                u8 *gbase = dst + (out_pos - shift);            // 16-byte aligned
                const u32 end = shift + total;
                const u32 nchunk = (end + 15u) / 16u;
                for (u32 c = tid; c < nchunk; c += TPB) {
                    u32 lo = c * 16u, hi = lo + 16u;
                    if (lo >= shift && hi <= end) {
                        stg16(gbase + lo, *(const uint4 *)(sout + lo));
                    } else {
                        u32 a = lo < shift ? shift : lo, b = hi > end ? end : hi;
                        for (u32 i = a; i < b; i++) gbase[i] = sout[i];
                    }
                }
            }
            out_pos += total;
#pragma unroll
            for (int j = 0; j < UN; j++) cur[j] = nxt[j];
            syncthreads();
        }
        if (tid == 0) out_len[f] = out_pos;
        syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// decode
// ------------------------------------------------------------------------------------------
// decoder state maps as 8-bit LUTs: bits [2s+1:2s] = next state for state s
constexpr u32 MAP_ID = 0xE4u;   // 0->0 1->1 2->2 3->3
constexpr u32 MAP_NE = 0x15u;   // byte differs from previous: 0->1 1->1 2->1 3->0
constexpr u32 MAP_EQ = 0x39u;   // byte equals previous:       0->1 1->2 2->3 3->0

HC_DEV u32 map_apply(u32 m, u32 s) { return (m >> (2u * s)) & 3u; }
// first a then b
HC_DEV u32 map_compose(u32 a, u32 b)
{
    return map_apply(b, map_apply(a, 0)) | (map_apply(b, map_apply(a, 1)) << 2) |
           (map_apply(b, map_apply(a, 2)) << 4) | (map_apply(b, map_apply(a, 3)) << 6);
}
struct OpCompose { HC_DEVM u32 operator()(u32 a, u32 b) const { return map_compose(a, b); } };

constexpr u32 DEC_WIN = TPB * 64;   // 16 KiB of output per expansion window

HC_KERNEL HC_LAUNCH_BOUNDS(256, 2)
rle_decode_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT in_len,
                  u8 *HC_RESTRICT out, const u64 *HC_RESTRICT out_off, const u64 *HC_RESTRICT out_cap,
                  u64 *HC_RESTRICT out_len, i32 *HC_RESTRICT status, u32 nf)
{
    HC_SHARED u32 wtot[2][32];
    HC_SHARED u32 wlast[NW];
    HC_SHARED u32 heads[DEC_WIN / 32];
    HC_SHARED HC_ALIGNED16 u8 sval[DEC_WIN];
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    for (u32 f = blockIdx.x; f < nf; f += gridDim.x) {
        const u64 n = in_len[f];
        const u8 *src = in + in_off[f];
        u8 *dst = out ? out + out_off[f] : (u8 *)0;
        const u64 cap = out ? out_cap[f] : 0;
        u64 out_pos = 0;
        u32 carry_state = 0;

        uint4 cur[UN], nxt[UN];
#pragma unroll
        for (int j = 0; j < UN; j++) {
            u64 p = (u64)j * SUB_BYTES + tid * 16;
            cur[j] = p < n ? ldg16(src + p) : make_uint4_zero();
        }
        for (u64 t0 = 0; t0 < n; t0 += TILE_BYTES) {
#pragma unroll
            for (int j = 0; j < UN; j++) {
                u64 p = t0 + TILE_BYTES + (u64)j * SUB_BYTES + tid * 16;
                nxt[j] = p < n ? ldg16(src + p) : make_uint4_zero();
            }
            // ---- scan 1: decoder state maps -----------------------------------------------
            u32 eq[UN], valid[UN], pbyte[UN], fmap[UN], mexcl[UN];
#pragma unroll
            for (int j = 0; j < UN; j++) {
                const u64 p = t0 + (u64)j * SUB_BYTES + tid * 16;
                u32 last = cur[j].w >> 24;
                u32 pb = shfl_up(last, 1);
                if (lane == 0) pb = (p > 0 && p < n) ? ldg8(src + p - 1) : 0u;
                pbyte[j] = pb;
                u32 vm = p >= n ? 0u : (n - p >= 16 ? 0xffffu : ((1u << (u32)(n - p)) - 1u));
                u32 e = 0, prev = pb, m = MAP_ID;
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    u32 b = vec_byte(cur[j], k);
                    bool same = (b == prev);
                    if (same) e |= 1u << k;
                    prev = b;
                    if ((vm >> k) & 1u) m = map_compose(m, same ? MAP_EQ : MAP_NE);
                }
                eq[j] = e;
                valid[j] = vm;
                fmap[j] = m;
            }
            u32 tile_map = block_scan_striped(fmap, mexcl, MAP_ID, OpCompose(), wtot[0]);

            // ---- scan 2: output length of every token -------------------------------------
            u32 st_in[UN], cnt[UN], oexcl[UN];
#pragma unroll
            for (int j = 0; j < UN; j++) {
                u32 s = map_apply(mexcl[j], carry_state);
                st_in[j] = s;
                u32 c = 0;
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    if ((valid[j] >> k) & 1u) {
                        if (s == 3u) { c += vec_byte(cur[j], k); s = 0; }
                        else { c += 1u; s = (s == 0u) ? 1u : (((eq[j] >> k) & 1u) ? s + 1u : 1u); }
                    }
                }
                cnt[j] = c;
            }
            u32 total = block_scan_striped(cnt, oexcl, 0u, OpAdd(), wtot[1]);
            carry_state = map_apply(tile_map, carry_state);

            // ---- expansion, one 16 KiB output window at a time ---------------------------
            if (dst) {
                const u32 shift = (u32)(out_pos & 15u);
                // window coordinates: w = shift + tile-relative output position
                for (u32 w0 = 0; w0 < shift + total; w0 += DEC_WIN) {
                    const u32 w1 = w0 + DEC_WIN;
                    for (u32 i = tid; i < DEC_WIN / 32; i += TPB) heads[i] = 0;
                    syncthreads();
#pragma unroll
                    for (int j = 0; j < UN; j++) {
                        u32 o = shift + oexcl[j];
                        if (o < w1 && o + cnt[j] > w0) {
                            u32 s = st_in[j], prev = pbyte[j];
#pragma unroll
                            for (int k = 0; k < 16; k++) {
                                if ((valid[j] >> k) & 1u) {
                                    u32 b = vec_byte(cur[j], k), len, val;
                                    if (s == 3u) { len = b; val = prev; s = 0; }
                                    else { len = 1; val = b; s = (s == 0u) ? 1u : (((eq[j] >> k) & 1u) ? s + 1u : 1u); }
                                    if (len && o < w1 && o + len > w0) {
                                        u32 pos = (o > w0 ? o : w0) - w0;
                                        sval[pos] = (u8)val;
                                        atomic_or_shared(&heads[pos >> 5], 1u << (pos & 31u));
                                    }
                                    o += len;
                                    prev = b;
                                }
                            }
                        }
                    }
                    syncthreads();
                    // each thread fills 64 consecutive output bytes of the window
                    u32 h0 = heads[2 * tid], h1 = heads[2 * tid + 1];
                    u32 mylast = h1 ? 64u * tid + 32u + (31u - (u32)clz(h1)) + 1u
                                    : (h0 ? 64u * tid + (31u - (u32)clz(h0)) + 1u : 0u);
                    // exclusive max-scan over threads: last head before this thread's range
                    u32 inc = mylast;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        u32 t = shfl_up(inc, d);
                        if (lane >= (u32)d && t > inc) inc = t;
                    }
                    if (lane == 31) wlast[wid] = inc;
                    syncthreads();
                    u32 before = 0;
                    for (u32 i = 0; i < wid; i++) before = wlast[i] > before ? wlast[i] : before;
                    u32 le = shfl_up(inc, 1);
                    if (lane == 0) le = 0;
                    if (le > before) before = le;
                    u32 curv = before ? sval[before - 1u] : 0u;
                    const u32 lim = (shift + total - w0) < DEC_WIN ? (shift + total - w0) : DEC_WIN;  // valid bytes in window
                    const u32 lo_valid = w0 == 0 ? shift : 0u;
                    u8 *gbase = dst + (out_pos - shift) + w0;        // 16-byte aligned
                    const u64 gpos0 = out_pos - shift + w0;          // file-relative position of window byte 0
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        u32 wbase = 64u * tid + 16u * c;
                        u32 hb = ((c < 2 ? h0 : h1) >> (16u * (c & 1))) & 0xffffu;
                        u32 wv[4] = {0, 0, 0, 0};
#pragma unroll
                        for (int k = 0; k < 16; k++) {
                            if ((hb >> k) & 1u) curv = sval[wbase + k];
                            wv[k >> 2] |= curv << (8 * (k & 3));
                        }
                        if (wbase < lim) {
                            u32 a = wbase < lo_valid ? lo_valid : wbase;
                            u32 b = wbase + 16u > lim ? lim : wbase + 16u;
                            // clip against the caller's capacity
                            u64 capw = cap > gpos0 ? cap - gpos0 : 0;
                            if (b > capw) b = (u32)capw;
                            if (a == wbase && b == wbase + 16u) {
                                uint4 r; r.x = wv[0]; r.y = wv[1]; r.z = wv[2]; r.w = wv[3];
                                stg16(gbase + wbase, r);
                            } else {
                                for (u32 i = a; i < b; i++) gbase[i] = (u8)(wv[(i - wbase) >> 2] >> (8 * ((i - wbase) & 3)));
                            }
                        }
                    }
                    syncthreads();
                }
            }
            out_pos += total;
#pragma unroll
            for (int j = 0; j < UN; j++) cur[j] = nxt[j];
            syncthreads();
        }
        if (tid == 0) {
            out_len[f] = out_pos;
            if (status) status[f] = (out && out_pos > cap) ? 100 : 0;
        }
        syncthreads();
    }
}

}  // namespace hcd
