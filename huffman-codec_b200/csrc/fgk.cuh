// fgk.cuh -- batched adaptive Huffman (FGK) encode / decode.
// Reference: HuffTree::{encode,decode,update} src/huffman.cpp:37-128, findSuccNode :157-184,
// swapNodes :186-217, drivers applyHuffman/revertHuffman src/transform.cpp:363-406, container
// header src/headers.cpp:107-125 and bit packing src/main.cpp:78-84.
//
// The tree update is serial inside one stream, so the parallelism is the batch: ONE WARP PER
// FILE, tree in shared memory (4.8 KB per stream), all 32 lanes executing the same control flow.
// The reference's pointer tree with its whole-tree recursive leader search is replaced by the
// node-number indexed form (slot = node number 0..512, siblings adjacent, even slot = left
// child = bit 0; weights are non-decreasing in slot order), see SURVEY.md A.5:
//   code(s)   : collect s&1 while s = parent[s] until the root (slot 512)
//   leader(s) : last slot l >= s with w[l] == w[s]  -- 32 slots per step with one warp ballot
//   swap      : exchange the contents (kid/symbol) of slots s and l, re-point their children
// Encoding walks leaf -> root once, collecting the code bits and updating weights in the same
// pass until the first swap (after which the old path is finished by a pure parent chase).
// Bits are packed MSB-first into 32-bit words; each lane keeps one word and the warp flushes
// 128 bytes at a time (coalesced).  The 9-byte container header goes through the same writer.
#pragma once
#include "hc_common.cuh"

namespace hcd {

constexpr int FGK_WARPS = 4;            // streams per CTA
constexpr int FGK_ROOT = 512;
constexpr int FGK_WPAD = 548;           // w[] padded with sentinels for the 32-wide leader probe
constexpr i16 FGK_NYT = -257;           // kid[] < 0: leaf, symbol = -1 - kid; FGK_NYT: the NYT leaf

struct FgkTree {
    u32 w[FGK_WPAD];
    u16 parent[FGK_ROOT + 2];
    i16 kid[FGK_ROOT + 2];
    u16 slot_of[256];
    u32 nyt;
};

HC_DEV void fgk_init(FgkTree &t, u32 lane)
{
    for (u32 i = lane; i < (u32)FGK_WPAD; i += 32) t.w[i] = i <= (u32)FGK_ROOT ? 0u : 0xffffffffu;
    for (u32 i = lane; i < 256u; i += 32) t.slot_of[i] = 0xffffu;
    for (u32 i = lane; i < (u32)FGK_ROOT + 2u; i += 32) { t.kid[i] = FGK_NYT; t.parent[i] = (u16)FGK_ROOT; }
    if (lane == 0) t.nyt = FGK_ROOT;
    syncwarp();
}

// NYT split (src/huffman.cpp:99-111): returns the slot of the new symbol leaf.
// Only lane 0 writes the tree; the warp barrier publishes the writes to the other lanes.
HC_DEV u32 fgk_split(FgkTree &t, u32 sym, u32 lane)
{
    const u32 n = t.nyt;
    syncwarp();                               // every lane has read nyt before it moves
    if (lane == 0) {
        t.kid[n] = (i16)(n - 2);
        t.parent[n - 2] = (u16)n;
        t.parent[n - 1] = (u16)n;
        t.kid[n - 1] = (i16)(-1 - (i32)sym);
        t.kid[n - 2] = FGK_NYT;
        t.w[n - 1] = 0;
        t.w[n - 2] = 0;
        t.slot_of[sym] = (u16)(n - 1);
        t.nyt = n - 2;
    }
    syncwarp();
    return n - 1;
}

// FGK update from slot s (src/huffman.cpp:113-127).  If `coding`, also collects the code of s
// (pre-update tree) into code/depth: bit d of `code` = bit emitted (depth-d)th, i.e. the value
// `code` printed MSB-first over `depth` bits is the root->leaf path.
template <bool CODING>
HC_DEV void fgk_update(FgkTree &t, u32 s, u32 lane, u64 &code, u32 &depth)
{
    // All lanes walk the same path (uniform control flow) and only READ the tree; lane 0 is the
    // only writer.  Nothing written at one level is read again at a higher level of the same
    // walk (parents, leaders and probes all have larger slot numbers), so one warp barrier at
    // the end is enough to publish the update before the next symbol.
    bool coding = CODING;
    while (s != (u32)FGK_ROOT) {
        const u32 ws = t.w[s];
        const u32 wn = t.w[s + 1 + lane];
        u32 p = t.parent[s];
        if (coding) {
            code |= (u64)(s & 1u) << (depth & 63u);
            depth++;
        }
        u32 m = ~ballot(wn == ws);
        u32 run = m ? (u32)ffs(m) - 1u : 32u;
        u32 l = s + run;
        while (run == 32u) {               // block longer than the probe: keep scanning
            m = ~ballot(t.w[l + 1 + lane] == ws);
            run = m ? (u32)ffs(m) - 1u : 32u;
            l += run;
        }
        if (l != s && l != p) {
            if (coding) {                  // finish the code on the old path
                for (u32 c = p; c != (u32)FGK_ROOT; c = t.parent[c]) {
                    code |= (u64)(c & 1u) << (depth & 63u);
                    depth++;
                }
                coding = false;
            }
            if (lane == 0) {
                const i16 ks = t.kid[s], kl = t.kid[l];
                t.kid[s] = kl;
                t.kid[l] = ks;
                if (kl >= 0) { t.parent[kl] = (u16)s; t.parent[kl + 1] = (u16)s; }
                else if (kl == FGK_NYT) t.nyt = s;
                else t.slot_of[-1 - kl] = (u16)s;
                if (ks >= 0) { t.parent[ks] = (u16)l; t.parent[ks + 1] = (u16)l; }
                else if (ks == FGK_NYT) t.nyt = l;
                else t.slot_of[-1 - ks] = (u16)l;
            }
            s = l;
            p = t.parent[l];
        }
        if (lane == 0) t.w[s] = ws + 1u;
        s = p;
    }
    if (lane == 0) t.w[FGK_ROOT]++;
    syncwarp();
}

// MSB-first bit writer: one 32-bit word per lane, 128-byte coalesced flushes
struct BitWriter {
    u64 acc;
    u32 nacc;      // valid low bits of acc (< 32 between calls)
    u32 widx;      // words produced so far
    u32 mine;      // this lane's word of the current 32-word group
    u32 *dst;      // 128-byte aligned
    u64 cap_words;
    bool overflow;
};

HC_DEV void bw_init(BitWriter &b, u8 *dst, u64 cap_bytes)
{
    b.acc = 0; b.nacc = 0; b.widx = 0; b.mine = 0;
    b.dst = (u32 *)dst;
    b.cap_words = cap_bytes / 4;
    b.overflow = false;
}

HC_DEV void bw_put(BitWriter &b, u32 v, u32 d, u32 lane)   // d <= 32
{
    b.acc = (b.acc << d) | v;
    b.nacc += d;
    if (b.nacc >= 32u) {
        u32 word = (u32)(b.acc >> (b.nacc - 32u));
        b.nacc -= 32u;
        if (lane == (b.widx & 31u)) b.mine = bswap32(word);
        b.widx++;
        if ((b.widx & 31u) == 0u) {
            if ((u64)b.widx <= b.cap_words) stg32_stream(b.dst + (b.widx - 32u) + lane, b.mine);
            else b.overflow = true;
        }
    }
}

// flush the tail; returns the total number of bytes of the stream
HC_DEV u64 bw_finish(BitWriter &b, u32 lane)
{
    u32 rem = b.widx & 31u;
    u32 base = b.widx - rem;
    u32 tail_bytes = (b.nacc + 7u) / 8u;
    u64 total = (u64)b.widx * 4u + tail_bytes;
    if (total > b.cap_words * 4u) { b.overflow = true; return total; }
    if (lane < rem) b.dst[base + lane] = b.mine;
    if (lane == 0 && tail_bytes) {
        u32 word = (u32)(b.acc << (32u - b.nacc));           // left-align, zero padded
        u8 *p = (u8 *)(b.dst + b.widx);
        for (u32 i = 0; i < tail_bytes; i++) p[i] = (u8)(word >> (24u - 8u * i));
    }
    return total;
}

HC_KERNEL HC_LAUNCH_BOUNDS(FGK_WARPS * 32, 1)
fgk_encode_kernel(const u8 *HC_RESTRICT sym, const u64 *HC_RESTRICT sym_off, const u64 *HC_RESTRICT sym_len,
                  const u8 *HC_RESTRICT flags, u8 *HC_RESTRICT out, const u64 *HC_RESTRICT out_off,
                  const u64 *HC_RESTRICT out_cap, u64 *HC_RESTRICT out_len, i32 *HC_RESTRICT status, u32 nf)
{
    HC_SHARED FgkTree trees[FGK_WARPS];
    const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const u32 f = blockIdx.x * FGK_WARPS + wid;
    if (f >= nf) return;
    FgkTree &t = trees[wid];
    fgk_init(t, lane);

    const u64 m = sym_len[f];
    const u32 *src = (const u32 *)(sym + sym_off[f]);     // 256-byte aligned region
    BitWriter bw;
    bw_init(bw, out + out_off[f], out_cap[f] & ~(u64)3);
    // container header <u64 LE count><u8 flags> (src/headers.cpp:107-125) through the bit writer
    bw_put(bw, bswap32((u32)m), 32, lane);
    bw_put(bw, bswap32((u32)(m >> 32)), 32, lane);
    bw_put(bw, flags ? flags[f] : 0u, 8, lane);

    bool too_long = false;
    u32 chunk = 0, chunk_next = 0;                        // 4 symbols per lane, 128 per warp
    if (m > 0) chunk = ldg32(src + lane);
    for (u64 i0 = 0; i0 < m; i0 += 128) {
        if (i0 + 128 < m) chunk_next = ldg32(src + (i0 + 128) / 4 + lane);
        u32 cnt = (m - i0) < 128 ? (u32)(m - i0) : 128u;
        for (u32 i = 0; i < cnt; i++) {
            u32 word = shfl(chunk, (int)(i >> 2));
            u32 y = (word >> (8u * (i & 3u))) & 0xffu;
            u32 s = t.slot_of[y];
            u64 code = 0;
            u32 depth = 0;
            if (s == 0xffffu) {
                // not yet transmitted: NYT code followed by the 8 raw bits (src/huffman.cpp:42-51)
                for (u32 c = t.nyt; c != (u32)FGK_ROOT; c = t.parent[c]) {
                    code |= (u64)(c & 1u) << (depth & 63u);
                    depth++;
                }
                if (depth > 56u) too_long = true;
                if (depth > 32u) bw_put(bw, (u32)(code >> 32), depth - 32u, lane);
                bw_put(bw, (u32)code, depth > 32u ? 32u : depth, lane);
                bw_put(bw, y, 8, lane);
                s = fgk_split(t, y, lane);
                u64 dummy = 0; u32 dd = 0;
                fgk_update<false>(t, s, lane, dummy, dd);
            } else {
                fgk_update<true>(t, s, lane, code, depth);
                if (depth > 56u) too_long = true;
                if (depth > 32u) bw_put(bw, (u32)(code >> 32), depth - 32u, lane);
                bw_put(bw, (u32)code, depth > 32u ? 32u : depth, lane);
            }
        }
        chunk = chunk_next;
    }
    u64 total = bw_finish(bw, lane);
    if (lane == 0) {
        out_len[f] = total;
        status[f] = too_long ? 101 : (bw.overflow ? 100 : 0);
    }
}

// MSB-first bit reader over 128-byte chunks held one word per lane
struct BitReader {
    u64 win;        // next bits, MSB aligned
    u32 wbits;      // valid bits in win
    u32 ridx;       // next word index to pull into the window
    u32 chunk, chunk_next;
    const u32 *src;
    u64 nwords;     // words that may be loaded (region capacity)
    u64 avail;      // bits still available in the file (consumed bits are subtracted)
};

HC_DEV void br_refill(BitReader &r, u32 lane)
{
    // precondition: wbits <= 32
    u32 word = bswap32(shfl(r.chunk, (int)(r.ridx & 31u)));
    r.win |= (u64)word << (32u - r.wbits);
    r.wbits += 32u;
    r.ridx++;
    if ((r.ridx & 31u) == 0u) {
        r.chunk = r.chunk_next;
        u64 nx = (u64)r.ridx + 32u + lane;
        r.chunk_next = nx < r.nwords ? ldg32(r.src + nx) : 0u;
    }
}

HC_DEV void br_init(BitReader &r, const u8 *p, u64 len_bytes, u32 lane)
{
    r.src = (const u32 *)p;
    r.nwords = (len_bytes + 3u) / 4u;      // reads stay inside the 256-byte padded region
    r.avail = len_bytes * 8u;
    r.chunk = lane < r.nwords ? ldg32(r.src + lane) : 0u;
    r.chunk_next = 32u + lane < r.nwords ? ldg32(r.src + 32u + lane) : 0u;
    r.win = 0; r.wbits = 0; r.ridx = 0;
    br_refill(r, lane);
    br_refill(r, lane);
}

// take d (1..32) bits; caller checks r.avail first
HC_DEV u32 br_get(BitReader &r, u32 d, u32 lane)
{
    u32 v = (u32)(r.win >> (64u - d));
    r.win <<= d;
    r.wbits -= d;
    r.avail -= d;
    if (r.wbits <= 32u) br_refill(r, lane);
    return v;
}

HC_KERNEL HC_LAUNCH_BOUNDS(FGK_WARPS * 32, 1)
fgk_decode_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT in_len,
                  u8 *HC_RESTRICT sym, const u64 *HC_RESTRICT sym_off, const u64 *HC_RESTRICT sym_cap,
                  u64 *HC_RESTRICT sym_len, u8 *HC_RESTRICT flags, i32 *HC_RESTRICT status, u32 nf)
{
    HC_SHARED FgkTree trees[FGK_WARPS];
    const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const u32 f = blockIdx.x * FGK_WARPS + wid;
    if (f >= nf) return;
    FgkTree &t = trees[wid];
    fgk_init(t, lane);

    const u64 n = in_len[f];
    if (n < 9) {                                         // src/main.cpp:99-104
        if (lane == 0) { sym_len[f] = 0; if (flags) flags[f] = 0; status[f] = 8; }
        return;
    }
    BitReader br;
    br_init(br, in + in_off[f], n, lane);
    u32 lo = bswap32(br_get(br, 32, lane));
    u32 hi = bswap32(br_get(br, 32, lane));
    const u64 m = ((u64)hi << 32) | lo;
    u32 fl = br_get(br, 8, lane);
    if (lane == 0 && flags) flags[f] = (u8)fl;
    const u64 cap = sym_cap[f];
    if (m > cap || m > br.avail + 1u) {
        // more symbols than capacity; (every symbol after the first costs >= 1 bit, so a count
        // beyond avail+1 is a guaranteed underrun -> the reference exits with 9)
        if (lane == 0) { sym_len[f] = m; status[f] = (m > br.avail + 1u) ? 9 : 100; }
        return;
    }
    u32 *dst = (u32 *)(sym + sym_off[f]);
    u32 mine = 0;
    i32 err = 0;
    u64 i = 0;
    for (; i < m; i++) {
        u32 s = FGK_ROOT;
        i32 k = t.kid[s];
        while (k >= 0) {
            if (br.avail < 1u) { err = 9; break; }
            s = (u32)k + br_get(br, 1, lane);
            k = t.kid[s];
        }
        if (err) break;
        u32 y;
        if (k == FGK_NYT) {
            if (br.avail < 8u) { err = 9; break; }
            y = br_get(br, 8, lane);
            // a raw symbol that is already in the tree is still decoded as that symbol
            // (src/huffman.cpp:74-86); update then starts from its existing leaf
            u32 ex = t.slot_of[y];
            s = ex == 0xffffu ? fgk_split(t, y, lane) : ex;
        } else {
            y = (u32)(-1 - k);
        }
        u64 dummy = 0; u32 dd = 0;
        fgk_update<false>(t, s, lane, dummy, dd);
        u32 li = (u32)(i & 127u);
        if (lane == (li >> 2)) mine |= y << (8u * (li & 3u));
        if (li == 127u) { stg32_stream(dst + (i >> 7) * 32u + lane, mine); mine = 0; }
    }
    if (!err) {
        u32 rem = (u32)(m & 127u);                         // symbols in the last partial group
        if (rem && lane < (rem + 3u) / 4u) dst[(m >> 7) * 32u + lane] = mine;
    }
    if (lane == 0) { sym_len[f] = m; status[f] = err; }
}

}  // namespace hcd
