// rle.cuh -- MNP-5 run-length encoding kernels (reference: src/transform.cpp:137-159, 241-292).
//
// One CTA streams one file tile by tile (16 KiB tiles) carrying a few scalars between tiles, so a
// batch of files needs no inter-CTA communication at all and every byte is read from HBM exactly
// once (algorithmic traffic N + M).  A thread owns 64 CONSECUTIVE bytes of a tile: the run logic is a
// handful of 64-bit mask operations per thread and every block scan carries one value per thread.
// Equality bits: one XOR against the stream shifted by a byte, an exact zero-byte test and a DP4A that
// gathers eight flags per instruction pair.
//
// ENCODE (causal form).  q = index of an element inside its maximal run modulo 258 (runs are taken over
// elements 0..n-2, the last element is a run of its own).  Element k emits
//      [the count q' - 2 of the run that ended at k-1, if k starts a run and 2 <= q' < 257]
//      [its byte, if q < 3]   or   [255, if q == 257]
// so a thread needs its own bytes, the byte before them and the length of the run that reaches into them
// (block-wide max-scan of run-start positions) -- no look-ahead.  keep / pre masks by bit logic (q >= 3
// is E & E<<1 & E<<2); the 258 wrap can only happen in the part of a segment that continues the incoming
// run and is patched by clearing one bit.  Output position: block-wide add-scan of popc(keep) + popc(pre).
// Output assembly: per 4-byte word one table lookup (index = keep nibble | pre nibble << 4) gives two
// PRMT selectors that compact the word and leave a hole where a count goes; the bytes are appended by a
// BRANCH-FREE writer (64-bit funnel, predicated STS.32 of completed words) into a staging buffer whose
// 16-byte chunks are permuted inside every 128-byte line, because threads write at a stride of about
// 64 bytes = 16 banks.  Count values are stored into their holes afterwards (ffs loop over the pre mask).
// The staging buffer is copied out with aligned 128-bit stores.
//
// DECODE.  Whether an input byte is a literal or a count depends on the decoder state c in {0,1,2,3},
// whose transition only needs c and e[i] = (in[i] == in[i-1]).  Each position is therefore a 4->4 map,
// kept one byte per state so that composing two maps is one PRMT; the map of 8 positions comes from a
// table indexed by their equality bits (stored as map and as PRMT selector, so a thread composes its
// eight maps right to left with one PRMT each), and a block-wide scan under composition gives every
// thread its entry state.  A second table turns (state, 8 equality bits) into "which of these bytes are
// counts"; one 64-bit add-scan places the output (literals + sum of the count bytes) and numbers the runs.
// Expansion per output window: window w holds the threads whose output STARTS in [w, w + 16 KiB) with
// all of their output (a thread yields at most 4128 bytes; the staging buffer has that much room behind
// the window), so every thread is expanded exactly once and nothing is clipped.  Pass A: literals through
// the same branch-free writer (a count makes the writer jump; word stores may put zeros into bytes of the
// thread's own runs), runs into a run list at their scan-assigned slot.  Pass B: one thread per run fills
// it (after a barrier, so it overwrites whatever pass A left there).  Pass C: aligned 128-bit copy-out.
#pragma once
#include "runsum.cuh"
#include "scan.cuh"

namespace hcd {

// ------------------------------------------------------------------------------------------
// encode
// ------------------------------------------------------------------------------------------
// A thread owns ENC_W = 64 CONSECUTIVE input bytes per tile, so that the whole run logic is a handful
// of 64-bit mask operations per thread and both block scans carry one value per thread.
#ifndef HC_RLE_TPB
#define HC_RLE_TPB 256
#endif
constexpr int RTPB = HC_RLE_TPB;               // threads of a CTA of the RLE coder (a power of two, 64..256)
constexpr int RNW = RTPB / 32;
constexpr u32 ENC_W = 64;
constexpr u32 ENC_TILE = RTPB * ENC_W;                                    // 16 KiB of input per CTA step
constexpr u32 ENC_STAGE_BYTES = (ENC_TILE + ENC_TILE / 3 + 64 + 15) & ~15u;  // worst case 4/3 + one carried count + phase

// Threads write their output bytes at a stride of about 64 bytes = 16 banks, which would serialise the
// stores of a warp 16-fold; the staging buffer is therefore stored with its 16-byte chunks permuted
// inside every 128-byte line (chunk ^= line number mod 8): 8 consecutive chunks still fill all 32 banks
// for the 128-bit copy-out, and chunks 4 apart land in different banks.  `a` = shared address inside a
// 512-byte aligned buffer.
HC_DEV u32 stage_swz(u32 a) { return a ^ ((a >> 2) & 0x70u); }

struct RleEncShared {
    u8 stage[ENC_STAGE_BYTES];   // first member: 512-byte aligned like the object
    u32 lut[256];
    u32 wtot[2][RNW];
    u32 carry;
};

HC_DEV RleEncShared *rle_enc_shared()
{
    HC_SHARED HC_ALIGNED(512) RleEncShared sh;
    return &sh;
}

// Compaction table of the encoder.  Index = k4 | p4 << 4 for the four elements of one input word:
// k4 = elements that emit their own byte, p4 (subset of k4) = elements whose byte is preceded by the
// count byte of the run that ended just before them.  Entry = two PRMT selectors (low / high output
// word) that move the kept bytes together and leave a zero byte where a count goes (selector nibble
// 4 = byte 0 of the second PRMT operand, which is zero).
// called once per kernel by all RTPB threads before the first rle_encode_stream
HC_DEV void rle_enc_init()
{
    RleEncShared *sh = rle_enc_shared();
    for (u32 idx = threadIdx.x; idx < 256u; idx += RTPB) {
        const u32 k4 = idx & 15u, p4 = idx >> 4;
        u64 sel = 0x4444444444444444ull;
        u32 o = 0;
        for (u32 k = 0; k < 4u; k++) {
            if ((k4 >> k) & 1u) {
                if ((p4 >> k) & 1u) o++;                      // nibble stays 4: the count byte's place
                sel = (sel & ~(0xfull << (4u * o))) | ((u64)k << (4u * o));
                o++;
            }
        }
        sh->lut[idx] = (u32)(sel & 0xffffu) | ((u32)((sel >> 16) & 0xffffu) << 16);
    }
    syncthreads();
}

// 0x80 in every byte of the result where the byte of d is zero (exact, no borrow between bytes)
HC_DEV u32 rle_zero_bytes(u32 d)
{
    const u32 t = (d & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    return ~(t | d) & 0x80808080u;
}

// 128 * (eight equality bits) of the words a, b (a first): bit k = byte k equals the byte before it;
// pw = the word before a
HC_DEV u32 rle_eq8x128(u32 pw, u32 a, u32 b)
{
    const u32 za = rle_zero_bytes(a ^ funnel_l(pw, a, 8)), zb = rle_zero_bytes(b ^ funnel_l(a, b, 8));
    return dp4a_u(za, 0x08040201u, dp4a_u(zb, 0x80402010u, 0u));
}

// equality bits of 32 consecutive bytes v0, v1; pw = the word before them
HC_DEV u32 rle_eq32(u32 pw, const uint4 &v0, const uint4 &v1)
{
    const u32 m0 = rle_eq8x128(pw, v0.x, v0.y), m1 = rle_eq8x128(v0.y, v0.z, v0.w);
    const u32 m2 = rle_eq8x128(v0.w, v1.x, v1.y), m3 = rle_eq8x128(v1.y, v1.z, v1.w);
    return (m0 >> 7) | (m1 << 1) | (m2 << 9) | (m3 << 17);
}

// exclusive block scan of one value per thread; wtot: RNW words that no other scan touches before the
// next barrier after this call.  Returns the fold of the whole block.
template <class Op>
HC_DEV u32 block_scan1(u32 v, u32 &excl, u32 identity, Op op, u32 *wtot)
{
    static_assert(RNW == 2 || RNW == 4 || RNW == 8, "the cross-warp step scans RNW partials inside groups of RNW lanes");
    const u32 lane = lane_id(), w = warp_id();
    u32 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 t = shfl_up(inc, d);
        if (lane >= (u32)d) inc = op(t, inc);
    }
    if (lane == 31) wtot[w] = inc;
    syncthreads();
    u32 pin = wtot[lane & (RNW - 1u)];                    // every group of RNW lanes scans the RNW partials
#pragma unroll
    for (int d = 1; d < RNW; d <<= 1) {
        const u32 t = shfl_up(pin, d);
        if ((lane & (RNW - 1u)) >= (u32)d) pin = op(t, pin);
    }
    const u32 total = shfl(pin, RNW - 1);
    u32 base = shfl(pin, (int)(w ? w - 1u : 0u));
    if (w == 0) base = identity;
    u32 le = shfl_up(inc, 1);
    if (lane == 0) le = identity;
    excl = op(base, le);
    return total;
}

// Branch-free byte appender into the (permuted) staging buffer.  Bytes collect in w0; a word goes out with one
// STS.32 as soon as it is complete.  A thread's first word may hold bytes of the thread before it: that word is
// kept back (`head`) and its bytes are stored one by one at the end, like the incomplete last word.
struct EncWriter {
    u32 w0;        // the word being assembled (fill < 4 valid bytes between calls, including leading filler)
    u32 fill;
    u32 waddr;     // logical shared address of the word being assembled
    u32 first;     // address of the first word if it starts with filler (shared with the thread before), else ~0
    u32 skip;      // filler bytes of the first word
    u32 head;      // the completed first word, if it is a shared one
};

HC_DEV void ew_init(EncWriter &w, u32 stage_addr, u32 o)
{
    w.w0 = 0; w.head = 0;
    w.fill = o & 3u;
    w.skip = o & 3u;
    w.waddr = stage_addr + (o & ~3u);
    w.first = w.skip ? w.waddr : 0xffffffffu;
}

// append the low nb (0..8) bytes of hi:lo; the bytes above nb must be zero
HC_DEV void ew_put8(EncWriter &w, u32 lo, u32 hi, u32 nb)
{
    const u32 s = 8u * w.fill;
    const u32 a0 = w.w0 | (lo << s), a1 = funnel_l(lo, hi, s), a2 = funnel_l(hi, 0u, s);
    const u32 f = w.fill + nb;                      // 0..11
    const bool shared = w.waddr == w.first;
    if (f >= 4u && !shared) sts32(stage_swz(w.waddr), a0);
    if (f >= 4u && shared) w.head = a0;
    if (f >= 8u) sts32(stage_swz(w.waddr + 4u), a1);
    w.w0 = f >= 8u ? a2 : (f >= 4u ? a1 : a0);
    w.waddr += f & ~3u;
    w.fill = f & 3u;
}

// append the low nb (0..4) bytes of x
HC_DEV void ew_put4(EncWriter &w, u32 x, u32 nb)
{
    const u32 s = 8u * w.fill;
    const u32 a0 = w.w0 | (x << s), a1 = funnel_l(x, 0u, s);
    const u32 f = w.fill + nb;                      // 0..7
    const bool shared = w.waddr == w.first;
    if (f >= 4u && !shared) sts32(stage_swz(w.waddr), a0);
    if (f >= 4u && shared) w.head = a0;
    w.w0 = f >= 4u ? a1 : a0;
    w.waddr += f & ~3u;
    w.fill = f & 3u;
}

// append 16 bytes
HC_DEV void ew_put16(EncWriter &w, const uint4 &v)
{
    const u32 s = 8u * w.fill;
    const u32 a0 = w.w0 | (v.x << s);
    if (w.waddr == w.first) w.head = a0;
    else sts32(stage_swz(w.waddr), a0);
    sts32(stage_swz(w.waddr + 4u), funnel_l(v.x, v.y, s));
    sts32(stage_swz(w.waddr + 8u), funnel_l(v.y, v.z, s));
    sts32(stage_swz(w.waddr + 12u), funnel_l(v.z, v.w, s));
    w.waddr += 16u;
    w.w0 = funnel_l(v.w, 0u, s);
}

HC_DEV void ew_finish(EncWriter &w)
{
    if (w.skip && w.waddr != w.first) {
        // bytes skip..3 of the first word (skip = 1..3)
        const u32 a = stage_swz(w.first);
        if (w.skip <= 1u) sts8(a + 1u, w.head >> 8);
        if (w.skip <= 2u) sts8(a + 2u, w.head >> 16);
        sts8(a + 3u, w.head >> 24);
        w.skip = 0;
    }
    // bytes skip..fill-1 of the last word (fill <= 3); skip != 0 only if the first word is the last one too
    const u32 a = stage_swz(w.waddr);
    if (w.skip == 0u && w.fill > 0u) sts8(a, w.w0);
    if (w.skip <= 1u && w.fill > 1u) sts8(a + 1u, w.w0 >> 8);
    if (w.skip <= 2u && w.fill > 2u) sts8(a + 2u, w.w0 >> 16);
}

// Encodes the n-byte stream at src (16-byte aligned) to dst (any alignment); called by all RTPB
// threads of a CTA, returns the number of bytes written.  Ends with a CTA barrier.  The kernel must
// have called rle_enc_init() before.
//
// Causal form of src/transform.cpp:241-279: with q = index of an element inside its maximal run
// modulo 258 (runs are taken over elements 0..n-2, the last element is a run of its own), element k emits
//      [the count q' - 2 of the run that ended at k-1, if k starts a run and 2 <= q' < 257]
//      [its byte, if q < 3]   or   [255, if q == 257]
// so that everything a thread emits follows from its own bytes, the byte before them and the length of
// the run that reaches into them (a block-wide max-scan of run-start positions).
HC_DEV u64 rle_encode_stream(const u8 *HC_RESTRICT src, u64 n, u8 *HC_RESTRICT dst)
{
    RleEncShared *sh = rle_enc_shared();
    HC_SMEM_ARENA(*sh);
    const u32 tid = threadIdx.x, lane = tid & 31;
    const u32 stage = smem_addr(sh->stage);
    const u32 dphase = (u32)((uintptr_t)dst & 15u);
    const u32 base = tid * ENC_W;
    const u32 *lut = sh->lut;
    u64 out_pos = 0;      // bytes emitted by all previous tiles
    u32 carry = 0;        // length modulo 258 of the run that ends at the last element of the previous tile

    // one tile ahead: the 64 bytes of the thread and, for lane 0, the byte before them (the other lanes
    // get it by shuffle)
    uint4 nx[4];
    u32 hn = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const u64 p = (u64)base + 16u * i;
        nx[i] = p < n ? ldg16_l1(src + p) : make_uint4_zero();
    }
    if (lane == 0 && base > 0 && base < n) hn = ldg8(src + base - 1);
    for (u64 t0 = 0; t0 < n; t0 += ENC_TILE) {
        uint4 v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = nx[i];
        const u32 hb = hn;
        const u64 g0 = t0 + base;                     // file position of this thread's element 0
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const u64 p = g0 + ENC_TILE + 16u * i;
            nx[i] = p < n ? ldg16_l1(src + p) : make_uint4_zero();
        }
        if (lane == 0 && g0 + ENC_TILE < n) hn = ldg8(src + g0 + ENC_TILE - 1);

        // ---- equality bits, run starts -----------------------------------------------------
        u32 pb = shfl_up(v[3].w >> 24, 1);
        if (lane == 0) pb = hb;
        u64 E = (u64)rle_eq32(pb << 24, v[0], v[1]) | ((u64)rle_eq32(v[1].w, v[2], v[3]) << 32);
        u64 V = ~0ull;
        if (g0 == 0) E &= ~1ull;                      // nothing before the first element
        if (g0 + ENC_W >= n) {
            // ragged end: the file's last element lies in this segment, or the segment (partly) beyond it
            const u32 cv = g0 < n ? (u32)(n - g0) : 0u;
            V = cv >= 64u ? ~0ull : ((1ull << cv) - 1ull);
            if (cv) E &= ~(1ull << (cv - 1u));        // the last element never continues a run
            E &= V;
        }
        const u64 S = V & ~E;                         // run starts
        // tile-relative position + 1 of the last run start owned by this thread
        const u32 smax = S ? base + (63u - (u32)clzll(S)) + 1u : 0u;
        u32 sexcl;
        block_scan1(smax, sexcl, 0u, OpMax(), sh->wtot[0]);

        // ---- what every element emits ---------------------------------------------------------
        // r_in = length of the run that ends at the element before this segment, z = r_in mod 258 = run
        // index of element 0 if it continues that run
        const u32 r_in = sexcl ? base - (sexcl - 1u) : carry + base;
        const u32 z = r_in % 258u;
        const bool cont = (u32)E & 1u;
        u64 Eq = E;                                   // E with the places cleared where q restarts at 0
        if (cont && z == 0u) Eq &= ~1ull;
        u64 wrap = 0;                                 // the element with q == 257 (emits 255)
        if (cont && z >= 258u - ENC_W) {
            const u32 kw = 257u - z;                  // < 64; only if the run gets that far
            if ((~E & ((2ull << kw) - 1ull)) == 0ull) {
                wrap = 1ull << kw;
                Eq &= ~(2ull << kw);
            }
        }
        const u32 c1 = z >= 2u ? 1u : 0u, c2 = z >= 3u ? 1u : 0u;   // "e" bits of the elements -1, -2
        const u64 e1 = (Eq << 1) | c1, e2 = (Eq << 2) | (c1 << 1) | c2;
        const u64 q3 = Eq & e1 & e2;                  // q >= 3
        const u64 keep = (V & ~q3) | wrap;            // emits its byte (255 for the wrap element, patched below)
        const u64 pre = S & e1 & e2 & ~(wrap << 1);   // starts a run and the run before it has 2 <= q' < 257
        const u64 Sq = V & ~Eq;                       // places where q == 0
        const u32 cnt = (u32)popcll(keep) + (u32)popcll(pre);
        if (tid == RTPB - 1) {
            // run length modulo 258 at the last element of a full tile
            sh->carry = Sq ? (u32)clzll(Sq) + 1u : (z + ENC_W) % 258u;
        }
        u32 oexcl;
        const u32 total = block_scan1(cnt, oexcl, 0u, OpAdd(), sh->wtot[1]);

        // ---- stage the output bytes -------------------------------------------------
        const u32 shift = (u32)((dphase + out_pos) & 15u);
        if (cnt) {
            const u32 o = shift + oexcl;
            EncWriter w;
            ew_init(w, stage, o);
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const u32 k16 = (u32)(keep >> (16 * g)) & 0xffffu, p16 = (u32)(pre >> (16 * g)) & 0xffffu;
                if (k16 == 0u) continue;
                if (k16 == 0xffffu && p16 == 0u) { ew_put16(w, v[g]); continue; }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const u32 idx = ((k16 >> (4 * j)) & 15u) | (((p16 >> (4 * j)) & 15u) << 4);
                    const u32 x = vec_word(v[g], j), e = lut[idx];
                    ew_put8(w, prmt_raw(x, 0u, e), prmt_raw(x, 0u, e >> 16), (u32)popc(idx));
                }
            }
            ew_finish(w);
            // count bytes (into the places the table left open) and the 255 of a wrap, 32 elements at a time
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const u32 keep_h = (u32)(keep >> (32 * h)), pre_h = (u32)(pre >> (32 * h));
                const u32 sq_h = (u32)(Sq >> (32 * h)), wrap_h = (u32)(wrap >> (32 * h));
                const u32 obase = h ? o + (u32)popc((u32)keep) + (u32)popc((u32)pre) : o;
                // q of the element before this half
                const u32 qb = h == 0 ? z - 1u : ((u32)Sq ? (u32)clz((u32)Sq) : z + 31u);
                u32 todo = pre_h | wrap_h;
                while (todo) {
                    const u32 k = (u32)ffs(todo) - 1u;
                    todo &= todo - 1u;
                    const u32 below = (1u << k) - 1u;
                    const u32 pos = obase + (u32)popc(keep_h & below) + (u32)popc(pre_h & below);
                    const u32 sb = sq_h & below;
                    // q of element k-1, minus 2 (255 for the wrap element)
                    u32 val = (sb ? k - 1u - (31u - (u32)clz(sb)) : qb + k) - 2u;
                    if ((wrap_h >> k) & 1u) val = 255u;
                    sts8(stage_swz(stage + pos), val);
                }
            }
        }
        syncthreads();
        carry = sh->carry;
        // ---- copy out: stage[shift .. shift+total) -> dst[out_pos ..) ---------------------
        {
            u8 *gbase = dst + out_pos - shift;              // 16-byte aligned
            const u32 end = shift + total;
            const u32 nchunk = (end + 15u) / 16u;
            for (u32 c = tid; c < nchunk; c += RTPB) {
                const u32 lo = c * 16u, hi = lo + 16u;
                if (lo >= shift && hi <= end) {
                    stg16(gbase + lo, lds128(stage_swz(stage + lo)));
                } else {
                    const u32 a = lo < shift ? shift : lo, b = hi > end ? end : hi;
                    for (u32 i = a; i < b; i++) gbase[i] = (u8)lds8(stage_swz(stage + i));
                }
            }
        }
        out_pos += total;
        syncthreads();
    }
    return out_pos;
}

#ifndef HC_RLE_ENC_MINB
#define HC_RLE_ENC_MINB (768 / HC_RLE_TPB)
#endif
HC_KERNEL HC_LAUNCH_BOUNDS(RTPB, HC_RLE_ENC_MINB)
rle_encode_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT in_len,
                  u8 *HC_RESTRICT out, const u64 *HC_RESTRICT out_off, u64 *HC_RESTRICT out_len, u32 nf)
{
    rle_enc_init();
    for (u32 f = blockIdx.x; f < nf; f += gridDim.x) {
        const u64 m = rle_encode_stream(in + in_off[f], in_len[f], out + out_off[f]);
        if (threadIdx.x == 0) out_len[f] = m;
    }
}

// ------------------------------------------------------------------------------------------
// decode
// ------------------------------------------------------------------------------------------
// decoder state maps, one byte per state: byte s = next state for state s.  Composition is then a
// single byte permute (the bytes of b selected by the values of a)
constexpr u32 MAP_ID = 0x03020100u;   // 0->0 1->1 2->2 3->3
constexpr u32 MAP_NE = 0x00010101u;   // byte differs from previous: 0->1 1->1 2->1 3->0
constexpr u32 MAP_EQ = 0x00030201u;   // byte equals previous:       0->1 1->2 2->3 3->0

HC_DEV u32 map_apply(u32 m, u32 s) { return (m >> (8u * s)) & 3u; }
// a map as a PRMT selector (one nibble per state)
HC_DEV u32 map_sel(u32 a)
{
    const u32 y = (a | (a >> 4)) & 0x00ff00ffu;
    return (y | (y >> 8)) & 0xffffu;
}
// first a then b
HC_DEV u32 map_compose(u32 a, u32 b) { return prmt_raw(b, 0u, map_sel(a)); }
struct OpCompose { HC_DEVM u32 operator()(u32 a, u32 b) const { return map_compose(a, b); } };

// The maps that token strings can produce form a monoid of only 40 elements (closure of {EQ, NE} under composition),
// so a map can also travel as an index and be composed by one table lookup: the walk of the adaptive block index
// (adapt.cuh), whose warp scans maps once per block, uses that form.  Built at compile time.
constexpr u32 MAPT_N = 40;
struct HC_ALIGNED16 MapTables {
    u32 bytes[64];           // index -> the map, one byte per state
    u8 comp[MAPT_N * 64];    // comp[a * 64 + b] = index of "first a, then b"
    u8 lane4[16];            // four equality bits -> index of the map of four tokens
    u8 cls4[64];             // state * 16 + four equality bits -> which of the four tokens are counts | state after << 4
    u8 id, eq, ne, n;
    u8 pad[12];
};
static_assert(sizeof(MapTables) % 16 == 0, "copied to shared memory in 16-byte pieces");

constexpr u32 mapt_compose(u32 a, u32 b)      // byte forms, first a then b
{
    u32 r = 0;
    for (u32 s = 0; s < 4; s++) r |= ((b >> (8 * ((a >> (8 * s)) & 3u))) & 3u) << (8 * s);
    return r;
}

constexpr MapTables make_map_tables()
{
    MapTables t{};
    u32 n = 0;
    t.bytes[n++] = MAP_ID;
    for (u32 i = 0; i < n; i++) {                       // breadth first: append EQ / NE to every known map
        for (u32 g = 0; g < 2; g++) {
            const u32 c = mapt_compose(t.bytes[i], g ? MAP_NE : MAP_EQ);
            bool known = false;
            for (u32 j = 0; j < n; j++) known = known || t.bytes[j] == c;
            if (!known && n < 64) t.bytes[n++] = c;
        }
    }
    t.n = (u8)n;
    for (u32 a = 0; a < n && a < MAPT_N; a++)
        for (u32 b = 0; b < n; b++) {
            const u32 c = mapt_compose(t.bytes[a], t.bytes[b]);
            for (u32 j = 0; j < n; j++) if (t.bytes[j] == c) t.comp[a * 64 + b] = (u8)j;
        }
    for (u32 j = 0; j < n; j++) {
        if (t.bytes[j] == MAP_ID) t.id = (u8)j;
        if (t.bytes[j] == MAP_EQ) t.eq = (u8)j;
        if (t.bytes[j] == MAP_NE) t.ne = (u8)j;
    }
    for (u32 e4 = 0; e4 < 16; e4++) {
        u32 m = MAP_ID;
        for (u32 k = 0; k < 4; k++) m = mapt_compose(m, ((e4 >> k) & 1u) ? MAP_EQ : MAP_NE);
        for (u32 j = 0; j < n; j++) if (t.bytes[j] == m) t.lane4[e4] = (u8)j;
        for (u32 s0 = 0; s0 < 4; s0++) {
            u32 st = s0, cm = 0;
            for (u32 k = 0; k < 4; k++) {
                if (st == 3u) { cm |= 1u << k; st = 0; }
                else st = (st == 0u) ? 1u : (((e4 >> k) & 1u) ? st + 1u : 1u);
            }
            t.cls4[s0 * 16 + e4] = (u8)(cm | (st << 4));
        }
    }
    return t;
}

HC_DEVICE_CONST MapTables g_map_tables = make_map_tables();
static_assert(make_map_tables().n == MAPT_N, "the monoid of decoder state maps has 40 elements");

constexpr u32 DEC_W = 64;                      // token bytes per thread per tile
constexpr u32 DEC_TILE = RTPB * DEC_W;         // 16 KiB of tokens per CTA step
constexpr u32 DEC_WIN = RTPB * 64;             // output bytes per expansion window
constexpr u32 DEC_MAXRUN = DEC_TILE / 4 + 8;   // a count byte needs three literals before it
constexpr u32 DEC_MAXOUT = 16u * 255u + 48u;   // most output bytes of one thread's 64 tokens
constexpr u32 DEC_STAGE = DEC_WIN + ((DEC_MAXOUT + 15u) & ~15u) + 80u;

// Tables of the decoder, indexed by 8 consecutive equality bits (bit k: byte k equals byte k-1):
//   map8[e8]        the state map of those 8 bytes, sel8[e8] the same map as a PRMT selector
//   cls8[s][e8]     entering in state s: bits 0..7 = which of the 8 bytes are COUNT bytes (read in
//                   state 3), bits 8..9 = the state after them
//   lut4[m4]        PRMT selector that moves the bytes m4 of a word together
struct RleDecShared {
    u8 stage[DEC_STAGE];         // first member: 512-byte aligned like the object; permuted as the encoder's
    u32 rpos[DEC_MAXRUN];        // runs of the tile: position (shifted tile coordinates) | length << 22
    u8 rval[DEC_MAXRUN];         // ... and their byte
    u32 map8[256];
    u16 sel8[256];
    u16 cls8[4][256];
    u32 lut4[16];
    u32 wtot[2][RNW];
    u64 wtot64[RNW];
    u32 ctl[2][4];               // per window (alternating): first / last + 1 run of its threads, end of their output
};
static_assert(sizeof(RleDecShared) <= 47u * 1024u, "static shared memory");

HC_DEV RleDecShared *rle_dec_shared()
{
    HC_SHARED HC_ALIGNED(512) RleDecShared sh;
    return &sh;
}

// called once per kernel by all RTPB threads before the first rle_decode_stream
HC_DEV void rle_dec_init()
{
    RleDecShared *t = rle_dec_shared();
    for (u32 e8 = threadIdx.x; e8 < 256u; e8 += RTPB) {
    u32 m = MAP_ID;
    for (u32 k = 0; k < 8u; k++) m = map_compose(m, ((e8 >> k) & 1u) ? MAP_EQ : MAP_NE);
    t->map8[e8] = m;
    t->sel8[e8] = (u16)map_sel(m);
    for (u32 s0 = 0; s0 < 4u; s0++) {
        u32 st = s0, cm = 0;
        for (u32 k = 0; k < 8u; k++) {
            if (st == 3u) { cm |= 1u << k; st = 0; }
            else st = (st == 0u) ? 1u : (((e8 >> k) & 1u) ? st + 1u : 1u);
        }
        t->cls8[s0][e8] = (u16)(cm | (st << 8));
    }
    if (e8 < 16u) {
        u32 sel = 0x4444u, o = 0;
        for (u32 k = 0; k < 4u; k++)
            if ((e8 >> k) & 1u) { sel = (sel & ~(0xfu << (4u * o))) | (k << (4u * o)); o++; }
        t->lut4[e8] = sel;
    }
    }
    syncthreads();
}

// state map of the valid bytes vm of a ragged 64-byte segment (first / last segment of a stream)
HC_DEV_NOINLINE u32 rle_dec_map_partial(u64 e, u64 vm)
{
    u32 m = MAP_ID;
    for (u32 k = 0; k < 64u; k++)
        if ((vm >> k) & 1ull) m = map_compose(m, ((e >> k) & 1ull) ? MAP_EQ : MAP_NE);
    return m;
}

// count-byte mask of the valid bytes of a ragged segment entered in state st
HC_DEV_NOINLINE u64 rle_dec_cls_partial(u64 e, u64 vm, u32 st)
{
    u64 cm = 0;
    for (u32 k = 0; k < 64u; k++) {
        if (!((vm >> k) & 1ull)) continue;
        if (st == 3u) { cm |= 1ull << k; st = 0; }
        else st = (st == 0u) ? 1u : (((e >> k) & 1ull) ? st + 1u : 1u);
    }
    return cm;
}

// spread the low 4 bits of m to the four byte lanes of a word (0xff per set bit)
HC_DEV u32 spread4(u32 m) { return (((m & 15u) * 0x00204081u) & 0x01010101u) * 0xffu; }

template <class Op>
HC_DEV u64 block_scan1_u64(u64 v, u64 &excl, Op op, u64 *wtot)
{
    const u32 lane = lane_id(), w = warp_id();
    u64 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u64 t = shfl_up64(inc, d);
        if (lane >= (u32)d) inc = op(t, inc);
    }
    if (lane == 31) wtot[w] = inc;
    syncthreads();
    u64 pin = wtot[lane & (RNW - 1u)];
#pragma unroll
    for (int d = 1; d < RNW; d <<= 1) {
        const u64 t = shfl_up64(pin, d);
        if ((lane & (RNW - 1u)) >= (u32)d) pin = op(t, pin);
    }
    const u64 total = shfl64(pin, RNW - 1);
    u64 base = shfl64(pin, (int)(w ? w - 1u : 0u));
    if (w == 0) base = 0;
    u64 le = shfl_up64(inc, 1);
    if (lane == 0) le = 0;
    excl = op(base, le);
    return total;
}
struct OpAdd64 { HC_DEVM u64 operator()(u64 a, u64 b) const { return a + b; } };

// position of the writer in bytes from the start of the staging buffer
HC_DEV u32 ew_pos(const EncWriter &w, u32 stage_addr) { return w.waddr - stage_addr + w.fill; }

// leave the next b bytes to a run (pass B of the decoder fills them after the literals): the word under
// assembly goes out as it is if the run reaches its end -- whatever a word store puts into bytes of the thread's
// own runs is overwritten afterwards
HC_DEV void ew_jump(EncWriter &w, u32 b)
{
    const u32 f = w.fill + b;
    const bool shared = w.waddr == w.first;
    if (f >= 4u && !shared) sts32(stage_swz(w.waddr), w.w0);
    if (f >= 4u && shared) w.head = w.w0;
    if (f >= 4u) w.w0 = 0u;
    w.waddr += f & ~3u;
    w.fill = f & 3u;
}

// Decodes the n0-byte token stream at src0 (any alignment) to dst (16-byte aligned, or null to
// only measure), writing at most cap bytes; called by all RTPB threads of a CTA, returns the decoded
// length.  An unaligned stream is read from the 16-byte boundary below it with the leading bytes
// masked out (they belong to the caller's buffer: a header or the previous block's tokens).
// The kernel must have called rle_dec_init() before.
HC_DEV u64 rle_decode_stream(const u8 *HC_RESTRICT src0, u64 n0, u8 *HC_RESTRICT dst, u64 cap)
{
    RleDecShared *sh = rle_dec_shared();
    HC_SMEM_ARENA(*sh);
    const u32 tid = threadIdx.x, lane = tid & 31;
    const u32 lead = n0 ? (u32)((uintptr_t)src0 & 15u) : 0u;
    const u8 *src = src0 - lead;
    const u64 n = n0 + lead;
    const u32 stage = smem_addr(sh->stage);
    const u32 base = tid * DEC_W;
    u64 out_pos = 0;
    u32 carry_state = 0;

    for (u64 t0 = 0; t0 < n; t0 += DEC_TILE) {
        // (no register prefetch of the next tile here: the other resident CTAs cover the load latency, and the
        // expansion below needs the registers)
        const u64 g0 = t0 + base;
        uint4 v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const u64 p = g0 + 16u * i;
            v[i] = p < n ? ldg16_l1(src + p) : make_uint4_zero();
        }
        u32 hb = 0;
        if (lane == 0 && g0 > 0 && g0 < n) hb = ldg8(src + g0 - 1);
        if (g0 + DEC_TILE < n) prefetch_l2(src + g0 + DEC_TILE);

        // ---- scan 1: decoder state maps -----------------------------------------------
        u32 pb = shfl_up(v[3].w >> 24, 1);
        if (lane == 0) pb = hb;
        const u64 E = (u64)rle_eq32(pb << 24, v[0], v[1]) | ((u64)rle_eq32(v[1].w, v[2], v[3]) << 32);
        u64 V = ~0ull;
        if (g0 + DEC_W > n || g0 == 0) {
            const u32 cv = g0 < n ? (n - g0 >= 64u ? 64u : (u32)(n - g0)) : 0u;
            V = cv >= 64u ? ~0ull : ((1ull << cv) - 1ull);
            if (g0 == 0) V &= ~((1ull << lead) - 1ull);
        }
        u32 fmap;
        if (V == ~0ull) {
            // right to left: the table holds every 8-byte map as a PRMT selector as well
            fmap = sh->map8[(u32)(E >> 56)];
#pragma unroll
            for (int i = 6; i >= 0; i--) fmap = prmt_raw(fmap, 0u, sh->sel8[(u32)(E >> (8 * i)) & 0xffu]);
        } else {
            fmap = V ? rle_dec_map_partial(E, V) : MAP_ID;
        }
        u32 mexcl;
        const u32 tile_map = block_scan1(fmap, mexcl, MAP_ID, OpCompose(), sh->wtot[0]);
        // (every thread has left the previous tile's last window; nobody touches ctl[0] before the next barrier)
        if (tid == 0) { sh->ctl[0][0] = 0xffffffffu; sh->ctl[0][1] = 0u; sh->ctl[0][2] = (u32)(out_pos & 15u); }

        // ---- scan 2: which bytes are counts, output length and number of runs of every thread ------
        u64 cm;
        {
            u32 s = map_apply(mexcl, carry_state);
            if (V == ~0ull) {
                cm = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const u32 c = sh->cls8[s][(u32)(E >> (8 * i)) & 0xffu];
                    cm |= (u64)(c & 0xffu) << (8 * i);
                    s = c >> 8;
                }
            } else {
                cm = V ? rle_dec_cls_partial(E, V, s) : 0ull;
            }
        }
        u32 olen = (u32)popcll(V & ~cm);
        if (cm) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const u32 c4 = (u32)(cm >> (4 * i)) & 15u;
                if (c4) olen = dp4a_u(vec_word(v[i >> 2], i & 3) & spread4(c4), 0x01010101u, olen);
            }
        }
        const u32 nruns = (u32)popcll(cm);
        u64 sexcl;
        const u64 stot = block_scan1_u64((u64)olen | ((u64)nruns << 32), sexcl, OpAdd64(), sh->wtot64);
        const u32 total = (u32)stot, oexcl = (u32)sexcl, rexcl = (u32)(sexcl >> 32);
        carry_state = map_apply(tile_map, carry_state);

        // ---- expansion, one output window at a time -----------------------------------
        // Window w0 holds the threads whose output STARTS in [w0, w0 + DEC_WIN) with all of their output (the
        // staging buffer has room for the longest output of one thread behind the window), so every thread
        // is expanded exactly once and nothing is ever clipped.
        if (dst) {
            const u32 shift = (u32)(out_pos & 15u);
            const u32 o = shift + oexcl;               // window coordinates: shift + tile-relative output position
            u32 done = shift;                          // everything below has been written to dst
            // first position (these coordinates) that the caller's capacity does not cover
            const u64 room = cap > out_pos - shift ? cap - (out_pos - shift) : 0;
            const u32 climit = room < 0xfffffff0ull ? (u32)room : 0xfffffff0u;
            u32 kp = 0;                                // the control words of this window
            for (u32 w0 = 0; w0 < shift + total; w0 += DEC_WIN, kp ^= 1u) {
                if (w0) syncthreads();                 // the copy-out of the previous window has read the staging buffer
                // pass A: literals into the window, runs into the run list
                const bool active = (olen | nruns) && o >= w0 && o < w0 + DEC_WIN;
                u32 rlo = 0xffffffffu, rhi = 0u;
                if (active) {
                    if (nruns) { rlo = rexcl; rhi = rexcl + nruns; }
                    u32 ri = rexcl;
                    EncWriter w;
                    ew_init(w, stage, o - w0);
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const u32 v16 = (u32)(V >> (16 * g)) & 0xffffu, c16 = (u32)(cm >> (16 * g)) & 0xffffu;
                        if (v16 == 0u) continue;
                        if (v16 == 0xffffu && c16 == 0u) { ew_put16(w, v[g]); continue; }
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const u32 v4 = (v16 >> (4 * j)) & 15u, c4 = (c16 >> (4 * j)) & 15u;   // at most one count per word
                            const u32 x = vec_word(v[g], j);
                            const u32 lit = v4 & ~c4, below = c4 - 1u;                 // below: all ones without a count
                            const u32 mb = lit & below;
                            ew_put4(w, prmt_raw(x, 0u, sh->lut4[mb]), (u32)popc(mb));
                            if (c4) {
                                // the word before x (its top byte precedes byte 0 of x)
                                const u32 pw = j ? vec_word(v[g], j ? j - 1 : 0) : (g ? v[g ? g - 1 : 0].w : pb << 24);
                                const u32 sp = spread4(c4);
                                const u32 b = dp4a_u(x & sp, 0x01010101u, 0u);
                                const u32 prev = dp4a_u(funnel_l(pw, x, 8) & sp, 0x01010101u, 0u);
                                sh->rpos[ri] = (ew_pos(w, stage) + w0) | (b << 22);
                                sh->rval[ri] = (u8)prev;
                                ri++;
                                ew_jump(w, b);
                                const u32 ma = lit & ~below;
                                ew_put4(w, prmt_raw(x, 0u, sh->lut4[ma]), (u32)popc(ma));
                            }
                        }
                    }
                    ew_finish(w);
                }
                rlo = reduce_min(rlo);
                rhi = reduce_max(rhi);
                const u32 myend = reduce_max(active ? o + olen : 0u);
                if (lane == 0 && myend) {
                    atomic_max_smem(&sh->ctl[kp][2], myend);
                    if (rhi) { atomic_min_smem(&sh->ctl[kp][0], rlo); atomic_max_smem(&sh->ctl[kp][1], rhi); }
                }
                syncthreads();
                const u32 ilo = sh->ctl[kp][0], ihi = sh->ctl[kp][1], wend_abs = sh->ctl[kp][2];
                // the next window's control words: their last readers left them before the barrier above, their next
                // writers come after the next one
                if (tid == 0) { sh->ctl[kp ^ 1u][0] = 0xffffffffu; sh->ctl[kp ^ 1u][1] = 0u; sh->ctl[kp ^ 1u][2] = wend_abs; }
                // pass B: the runs of these threads, one thread per run
                for (u32 i = ilo + tid; i < ihi; i += RTPB) {
                    const u32 e = sh->rpos[i], len = e >> 22;
                    if (len == 0u) continue;
                    const u32 val = sh->rval[i], val4 = val * 0x01010101u;
                    u32 q = (e & 0x3fffffu) - w0;
                    const u32 qe = q + len;
                    while ((q & 3u) && q < qe) { sts8(stage_swz(stage + q), val); q++; }
                    for (; q + 4u <= qe; q += 4u) sts32(stage_swz(stage + q), val4);
                    for (; q < qe; q++) sts8(stage_swz(stage + q), val);
                }
                syncthreads();
                // pass C: copy [done, wend) out, clipped to the caller's capacity
                {
                    const u32 wbeg = done - w0;
                    const u32 wend = (wend_abs < climit ? wend_abs : climit) - w0;     // (done <= climit or nothing is copied)
                    done = wend_abs;
                    u8 *gbase = dst + (out_pos - shift) + w0;        // 16-byte aligned
                    if (wend_abs != 0u && wbeg < wend && climit > w0) {
                        for (u32 c = wbeg / 16u + tid; c * 16u < wend; c += RTPB) {
                            const u32 lo = c * 16u, hi = lo + 16u;
                            if (lo >= wbeg && hi <= wend) {
                                stg16(gbase + lo, lds128(stage_swz(stage + lo)));
                            } else {
                                const u32 a = lo < wbeg ? wbeg : lo, b = hi > wend ? wend : hi;
                                for (u32 i = a; i < b; i++) gbase[i] = (u8)lds8(stage_swz(stage + i));
                            }
                        }
                    }
                }
            }
        }
        // (no barrier here: the next tile's scans separate its expansion from this copy-out)
        out_pos += total;
    }
    syncthreads();
    return out_pos;
}

#ifndef HC_RLE_DEC_MINB
#define HC_RLE_DEC_MINB (768 / HC_RLE_TPB)
#endif
HC_KERNEL HC_LAUNCH_BOUNDS(RTPB, HC_RLE_DEC_MINB)
rle_decode_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT in_len,
                  u8 *HC_RESTRICT out, const u64 *HC_RESTRICT out_off, const u64 *HC_RESTRICT out_cap,
                  u64 *HC_RESTRICT out_len, i32 *HC_RESTRICT status, u32 nf)
{
    rle_dec_init();
    for (u32 f = blockIdx.x; f < nf; f += gridDim.x) {
        const u64 cap = out ? out_cap[f] : 0;
        const u64 m = rle_decode_stream(in + in_off[f], in_len[f], out ? out + out_off[f] : (u8 *)0, cap);
        if (threadIdx.x == 0) {
            out_len[f] = m;
            if (status) status[f] = (out && m > cap) ? 100 : 0;
        }
    }
}

}  // namespace hcd
