// rle.cuh -- MNP-5 run-length encoding kernels (reference: src/transform.cpp:137-159, 241-292).
//
// One CTA streams one file tile by tile (16 KiB tiles, see scan.cuh) carrying a few scalars
// between tiles, so a batch of files needs no inter-CTA communication at all and every byte is
// read from HBM exactly once (algorithmic traffic N + M).
//
// ENCODE.  Per element: k = index inside its maximal run (runs are taken over elements
// 0..n-2, the last element is always its own literal), q = k mod 258.  The element emits
//      [q < 3] its byte, [q == 257] the byte 255, [last of run && 2 <= q < 257] the count q-2.
// A thread owns 16-byte vectors: equality bits by byte-SIMD compares, the two emission masks by
// bit logic on them (rle_vec_masks; k of a vector's first element comes from a block-wide max-scan
// of run-start positions), the output position from a block-wide exclusive add-scan of the
// per-vector byte counts.  Output assembly is per 4-byte word: a count replaces the byte of its
// (non-literal) element, then one PRMT with a selector from a 256-entry table compacts the word.
// Bytes are staged in shared memory with the same 16-byte phase as the global destination and
// copied out with 128-bit stores.
//
// DECODE.  Whether an input byte is a literal or a count depends on the decoder state
// c in {0,1,2,3}, whose transition only needs c and e[i] = (in[i] == in[i-1]).  Each position is
// therefore a 4->4 map, kept one byte per state so that composing two maps is one PRMT; the map
// of 8 positions comes from a table indexed by their equality bits, and a block-wide scan under
// composition gives every vector its entry state.  A second table turns (state, 8 equality bits)
// into "which of these bytes are counts"; a block-wide add-scan of the vector lengths (literals +
// the sum of the count bytes) places the output.  Expansion is done per 16 KiB output window:
// every token drops its value and a head flag at its first output position, then each thread
// fills 64 consecutive output bytes from the last head value (a max-scan finds the head that
// reaches into its range), 16 bytes at a time with shortcuts for all-literal and in-run chunks.
#pragma once
#include "runsum.cuh"
#include "scan.cuh"

namespace hcd {

constexpr u32 ENC_STAGE_BYTES = TILE_BYTES + TILE_BYTES / 3 + 64;   // worst case 4/3 + phase

// bit k = (byte k of v == byte k-1), byte -1 = prev (0x100 = none).  Byte-SIMD: one __vcmpeq4 per
// word against the stream shifted by one byte, then a multiply gathers the 0xFF bytes into bits.
HC_DEV u32 rle_eq_mask16(const uint4 &v, u32 prev)
{
    const u32 K = 0x08040201u;
    const u32 m0 = vcmpeq4(v.x, (v.x << 8) | (prev & 0xffu));
    const u32 m1 = vcmpeq4(v.y, funnel_l(v.x, v.y, 8));
    const u32 m2 = vcmpeq4(v.z, funnel_l(v.y, v.z, 8));
    const u32 m3 = vcmpeq4(v.w, funnel_l(v.z, v.w, 8));
    const u32 lo8 = (((m0 & K) | ((m1 & K) << 4)) * 0x01010101u) >> 24;
    const u32 hi8 = (((m2 & K) | ((m3 & K) << 4)) * 0x01010101u) >> 24;
    u32 e = lo8 | (hi8 << 8);
    if (prev > 0xffu) e &= ~1u;
    return e;
}

// Output of one 16-element vector.  e: equality bits 0..16 (bit 16 = the element after the vector),
// valid: bits of existing elements, k0: run index of element 0 if it continues a run (else unused;
// on return reduced modulo 258).
//   lit bit k  : element k emits its byte          (run index q = k mod 258 < 3)
//   cnt bit k  : element k emits the byte q - 2    (last of its run with 2 <= q < 257, or q == 257:
//                the marker 255 of src/transform.cpp:259-263 is 257 - 2)
// Pure bit logic.  Runs that start inside the vector cannot reach q = 257; only the
// leading segment (the elements that continue the incoming run) can, and is patched arithmetically.
HC_DEV void rle_vec_masks_impl(u32 e, u32 valid, u32 &k0, u32 &lit, u32 &cnt)
{
    const bool cont = e & 1u;
    const bool deep = cont && k0 + 16u >= 257u;
    if (deep) k0 %= 258u;
    // bits 1,0 = "e_-1", "e_-2": whether the incoming run already holds >= 2 / >= 3 elements; inside
    // a long run pretend it does and fix the leading segment below
    const u32 b1 = (cont && (deep || k0 >= 2u)) ? 2u : 0u, b2 = (cont && (deep || k0 >= 3u)) ? 1u : 0u;
    const u32 x = ((e & 0xffffu) << 2) | b1 | b2;     // bit k+2 = e_k
    const u32 q2 = (x >> 2) & (x >> 1);               // e_k & e_k-1           (q >= 2)
    const u32 q3 = q2 & x;                            // ... & e_k-2           (q >= 3)
    lit = valid & ~q3;
    cnt = valid & q2 & ~(e >> 1);                     // run ends here (next element does not continue)
    if (deep) {
        const u32 p = (u32)ffs((~e & 0xffffu) | 0x10000u) - 1u;   // elements 0..p-1 continue the incoming run (p >= 1)
        const u32 lead = (1u << p) - 1u;
        const u32 kw = 257u - k0;                                 // element with q == 257 (may lie beyond the vector)
        u32 litl = k0 < 3u ? (1u << (3u - k0)) - 1u : 0u;         // q < 3 at the start ...
        u32 m255 = 0u;
        if (kw < 16u) { litl |= 7u << (kw + 1u); m255 = 1u << kw; }   // ... and after the wrap
        u32 qe = k0 + p - 1u;                                     // run index of the segment's last element
        if (qe >= 258u) qe -= 258u;
        const u32 endbit = cnt & (1u << (p - 1u));                // set iff the run ends there
        lit = (lit & ~lead) | (litl & lead & valid);
        cnt = (cnt & ~lead) | ((qe >= 2u && qe < 257u) ? endbit : 0u) | (m255 & lead & valid);
    }
}

// out-of-line copy (the streaming kernels are instruction-cache bound when everything is inlined four
// times): lit | cnt << 16 in .x, the reduced k0 in .y
HC_DEV_NOINLINE uint2 rle_vec_masks(u32 e, u32 valid, u32 k0)
{
    u32 lit, cnt;
    rle_vec_masks_impl(e, valid, k0, lit, cnt);
    uint2 r; r.x = lit | (cnt << 16); r.y = k0;
    return r;
}

// run index modulo 258 of element k of a vector (k0: as returned by rle_vec_masks)
HC_DEV u32 rle_run_index(u32 e, u32 k0, u32 k)
{
    const u32 zeros = ~e & ((2u << k) - 1u);          // run starts at or below k
    if (zeros) return k - (31u - (u32)clz(zeros));
    const u32 q = k0 + k;
    return q >= 258u ? q - 258u : q;
}

// byte-granular writer into the shared staging buffer: bytes are assembled in a register pair,
// full words go out as STS.32, the ragged first/last bytes as STS.U8 (neighbouring threads own the
// other bytes of those words)
struct StageWriter {
    u32 lo, hi;    // lo: the word being assembled (fill < 4 valid bytes between calls)
    u32 fill;      // bytes in lo (including the leading filler of the first word)
    u32 waddr;     // shared address of the word being assembled
    u32 skip;      // filler bytes of the first word still to be skipped (0 after the first flush)
};

HC_DEV void sw_init(StageWriter &w, u32 stage_addr, u32 o)
{
    w.lo = 0; w.hi = 0;
    w.fill = o & 3u;
    w.skip = o & 3u;
    w.waddr = stage_addr + (o & ~3u);
}

HC_DEV void sw_flush_word(StageWriter &w)
{
    if (w.skip) {                                      // bytes skip..3 (skip = 1..3), predicated stores
        if (w.skip <= 1u) sts8(w.waddr + 1u, w.lo >> 8);
        if (w.skip <= 2u) sts8(w.waddr + 2u, w.lo >> 16);
        sts8(w.waddr + 3u, w.lo >> 24);
        w.skip = 0;
    } else {
        sts32(w.waddr, w.lo);
    }
    w.waddr += 4u;
    w.lo = w.hi;
    w.hi = 0;
    w.fill -= 4u;
}

// append the low nb (0..4) bytes of x; the bytes of x above nb must be zero
HC_DEV void sw_put(StageWriter &w, u32 x, u32 nb)
{
    const u32 s = 8u * w.fill;
    w.lo |= x << s;
    w.hi = funnel_l(x, 0u, s);          // bytes that spill into the next word (0 when s == 0)
    w.fill += nb;
    if (w.fill >= 4u) sw_flush_word(w);
}

HC_DEV void sw_put_byte(StageWriter &w, u32 b) { sw_put(w, b & 0xffu, 1u); }
HC_DEV void sw_put_word(StageWriter &w, u32 x) { sw_put(w, x, 4u); }

HC_DEV void sw_finish(StageWriter &w)
{
    // bytes skip..fill-1 of the last word (fill <= 3)
    if (w.skip == 0u && w.fill > 0u) sts8(w.waddr, w.lo);
    if (w.skip <= 1u && w.fill > 1u) sts8(w.waddr + 1u, w.lo >> 8);
    if (w.skip <= 2u && w.fill > 2u) sts8(w.waddr + 2u, w.lo >> 16);
}

// 16 bytes to the byte offset o of a shared staging buffer whose neighbouring bytes belong to other
// threads: whole words where possible, byte stores for the two ragged words
HC_DEV void stage_put16(u32 stage_addr, u32 o, const uint4 &v)
{
    const u32 r = o & 3u, a = stage_addr + (o & ~3u);
    if (r == 0u) {
        sts32(a, v.x); sts32(a + 4u, v.y); sts32(a + 8u, v.z); sts32(a + 12u, v.w);
    } else {
        const u32 s = 8u * r;
        const u32 w0 = v.x << s, w4 = v.w >> (32u - s);
        if (r <= 1u) sts8(a + 1u, w0 >> 8);
        if (r <= 2u) sts8(a + 2u, w0 >> 16);
        sts8(a + 3u, w0 >> 24);
        sts32(a + 4u, funnel_l(v.x, v.y, s));
        sts32(a + 8u, funnel_l(v.y, v.z, s));
        sts32(a + 12u, funnel_l(v.z, v.w, s));
        sts8(a + 16u, w4);
        if (r >= 2u) sts8(a + 17u, w4 >> 8);
        if (r >= 3u) sts8(a + 18u, w4 >> 16);
    }
}

// Compaction table of the encoder.  Index = m4 | b4 << 4 for one 4-element word: m4 = elements that
// emit their byte (a literal, or a count already substituted into the word), b4 (subset of m4) =
// elements followed by a count byte 0 (a run of exactly three ends there).  Entry = two PRMT
// selectors (low / high output word) that move the kept bytes together and insert the zero bytes
// (selector nibble 4 = byte 0 of the second PRMT operand, which is zero).
HC_DEV u32 *rle_enc_lut()
{
    HC_SHARED u32 lut[256];
    return lut;
}

// called once per kernel by all TPB threads before the first rle_encode_stream
HC_DEV void rle_enc_init()
{
    u32 *lut = rle_enc_lut();
    const u32 idx = threadIdx.x & 255u, m4 = idx & 15u, b4 = idx >> 4;
    u64 sel = 0x4444444444444444ull;
    u32 o = 0;
    for (u32 k = 0; k < 4u; k++) {
        if ((m4 >> k) & 1u) {
            sel = (sel & ~(0xfull << (4u * o))) | ((u64)k << (4u * o));
            o++;
            if ((b4 >> k) & 1u) o++;                  // nibble stays 4: a zero byte
        }
    }
    lut[idx] = (u32)(sel & 0xffffu) | ((u32)((sel >> 16) & 0xffffu) << 16);
    syncthreads();
}

// writes the output bytes of one vector (masks from rle_vec_masks) to the staging buffer at byte o
HC_DEV_NOINLINE void rle_stage_vector(uint4 v, u32 lit, u32 cbm, u32 eq, u32 k0, u32 stage, u32 o, const u32 *lut)
{
    if (lit == 0xffffu && cbm == 0u) {                    // all literals: the vector goes out verbatim
        stage_put16(stage, o, v);
        return;
    }
    StageWriter w;
    sw_init(w, stage, o);
    {
        // word by word: substitute the count of a run that ends on a non-literal element into its
        // byte (at most one per word: such elements are >= 4 apart), then compact
        const u32 keep = lit | cbm, both = lit & cbm, sub = cbm & ~lit;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            u32 x = i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w;
            const u32 s4 = (sub >> (4 * i)) & 15u;
            if (s4) {
                const u32 kk = (u32)ffs(s4) - 1u;
                const u32 val = rle_run_index(eq, k0, 4u * i + kk) - 2u;
                x = (x & ~(0xffu << (8u * kk))) | (val << (8u * kk));
            }
            const u32 idx = ((keep >> (4 * i)) & 15u) | (((both >> (4 * i)) & 15u) << 4);
            const u32 nb = (u32)popc(idx);
            if (nb == 0u) continue;
            const u32 e = lut[idx];
            sw_put(w, prmt(x, 0u, e & 0xffffu), nb < 4u ? nb : 4u);
            if (nb > 4u) sw_put(w, prmt(x, 0u, e >> 16), nb - 4u);
        }
    }
    sw_finish(w);
}

// Encodes the n-byte stream at src (16-byte aligned) to dst (any alignment); called by all TPB
// threads of a CTA, returns the number of bytes written.  Ends with a CTA barrier.  The kernel must
// have called rle_enc_init() before.
HC_DEV u64 rle_encode_stream(const u8 *HC_RESTRICT src, u64 n, u8 *HC_RESTRICT dst)
{
    HC_SHARED u32 wtot[2][32];
    HC_SHARED u32 s_carry_run;
    HC_SHARED HC_ALIGNED16 u8 sout[ENC_STAGE_BYTES];
    HC_SMEM_ARENA(wtot);
    const u32 tid = threadIdx.x, lane = tid & 31;
    const u32 stage = smem_addr(sout);
    const u32 dphase = (u32)((uintptr_t)dst & 15u);
    const u32 *lut = rle_enc_lut();
    {
        u64 out_pos = 0;      // bytes emitted by all previous tiles
        u32 carry_run = 0;    // length of the run that ends at the last element of the previous tile

        // halo: lane 0 needs the byte before its vector, lane 31 the byte after (other lanes get them
        // by shuffle); fetched one tile ahead like the vectors themselves
        uint4 cur[UN], nxt[UN];
        u32 hcur[UN], hnxt[UN];
#pragma unroll
        for (int j = 0; j < UN; j++) {
            u64 p = (u64)j * SUB_BYTES + tid * 16;
            cur[j] = p < n ? ldg16(src + p) : make_uint4_zero();
            hcur[j] = 0x100u;
            if (lane == 0 && p > 0 && p < n) hcur[j] = ldg8(src + p - 1);
            if (lane == 31 && p + 16 < n) hcur[j] = ldg8(src + p + 16);
        }
        for (u64 t0 = 0; t0 < n; t0 += TILE_BYTES) {
#pragma unroll
            for (int j = 0; j < UN; j++) {
                u64 p = t0 + TILE_BYTES + (u64)j * SUB_BYTES + tid * 16;
                nxt[j] = p < n ? ldg16(src + p) : make_uint4_zero();
                hnxt[j] = 0x100u;
                if (lane == 0 && p < n) hnxt[j] = ldg8(src + p - 1);
                if (lane == 31 && p + 16 < n) hnxt[j] = ldg8(src + p + 16);
            }
            // ---- equality bits and run starts ------------------------------------------
            u32 eq[UN];      // bit k (0..16): element k equals element k-1 (bit 16 = next thread's first)
            u32 valid[UN];   // bit k: element exists
            u32 smax[UN], sexcl[UN];
#pragma unroll
            for (int j = 0; j < UN; j++) {
                const u64 p = t0 + (u64)j * SUB_BYTES + tid * 16;
                const u32 first = cur[j].x & 0xffu, last = cur[j].w >> 24;
                u32 pb = shfl_up(last, 1), nb = shfl_down(first, 1);
                if (lane == 0) pb = hcur[j];
                if (lane == 31) nb = hcur[j];
                u32 e = rle_eq_mask16(cur[j], pb);
                if (nb == last) e |= 1u << 16;
                const u32 vm = p >= n ? 0u : (n - p >= 17 ? 0x1ffffu : ((1u << (u32)(n - p)) - 1u));
                // the last element of the file never continues a run (forced literal)
                if (n - 1 >= p && n - 1 - p <= 16) e &= ~(1u << (u32)(n - 1 - p));
                e &= vm;
                eq[j] = e;
                valid[j] = vm & 0xffffu;
                const u32 starts = valid[j] & ~e;
                // tile-relative position + 1 of the last run start owned by this thread
                smax[j] = starts ? (u32)j * SUB_BYTES + tid * 16 + (31u - (u32)clz(starts)) + 1u : 0u;
            }
            block_scan_striped(smax, sexcl, 0u, OpMax(), wtot[0]);

            // ---- per-vector output masks and counts ----------------------------------------
            u32 cnt[UN], oexcl[UN], k0v[UN], lit[UN], cbm[UN];
#pragma unroll
            for (int j = 0; j < UN; j++) {
                const u32 tp = (u32)j * SUB_BYTES + tid * 16;   // tile-relative position
                u32 k0 = sexcl[j] ? tp - (sexcl[j] - 1u) : carry_run + tp;   // run index of element 0
                if (j == UN - 1 && tid == TPB - 1) {
                    // length of the run that ends at the last element of a full tile (kept below
                    // 2^15 + 258: only its value modulo 258 and "is it long" matter)
                    const u32 st = valid[j] & ~eq[j];
                    u32 cr = st ? 16u - (31u - (u32)clz(st)) : ((eq[j] & 1u) ? k0 + 16u : 16u);
                    if (cr >= 258u * 128u) cr = 258u * 64u + cr % 258u;
                    s_carry_run = cr;
                }
                const uint2 mk = rle_vec_masks(eq[j], valid[j], k0);
                lit[j] = mk.x & 0xffffu;
                cbm[j] = mk.x >> 16;
                k0v[j] = mk.y;
                cnt[j] = (u32)popc(mk.x);
            }
            u32 total = block_scan_striped(cnt, oexcl, 0u, OpAdd(), wtot[1]);

            // ---- stage the output bytes -------------------------------------------------
            const u32 shift = (u32)((dphase + out_pos) & 15u);
#pragma unroll
            for (int j = 0; j < UN; j++) {
                if (cnt[j] == 0) continue;
                rle_stage_vector(cur[j], lit[j], cbm[j], eq[j], k0v[j], stage, shift + oexcl[j], lut);
            }
            syncthreads();
            carry_run = s_carry_run;
            // ---- copy out: sout[shift .. shift+total) -> dst[out_pos ..) ---------------------
            {
                u8 *gbase = dst + out_pos - shift;              // 16-byte aligned
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
            for (int j = 0; j < UN; j++) { cur[j] = nxt[j]; hcur[j] = hnxt[j]; }
            syncthreads();
        }
        return out_pos;
    }
}

HC_KERNEL HC_LAUNCH_BOUNDS(256, 3)
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
// first a then b
HC_DEV u32 map_compose(u32 a, u32 b)
{
    const u32 y = (a | (a >> 4)) & 0x00ff00ffu;       // bytes -> nibbles: the PRMT selector
    return prmt(b, 0u, (y | (y >> 8)) & 0xffffu);
}
struct OpCompose { HC_DEVM u32 operator()(u32 a, u32 b) const { return map_compose(a, b); } };

constexpr u32 DEC_WIN = TPB * 64;   // 16 KiB of output per expansion window

// Tables of the decoder, indexed by 8 consecutive equality bits (bit k: byte k equals byte k-1):
//   map8[e8]        the state map of those 8 bytes
//   cls8[s][e8]     entering in state s: bits 0..7 = which of the 8 bytes are COUNT bytes (read in
//                   state 3), bits 8..9 = the state after them
struct RleDecTables {
    u32 map8[256];
    u16 cls8[4][256];
};

HC_DEV RleDecTables *rle_dec_tables()
{
    HC_SHARED RleDecTables t;
    return &t;
}

// called once per kernel by all TPB threads before the first rle_decode_stream
HC_DEV void rle_dec_init()
{
    RleDecTables *t = rle_dec_tables();
    const u32 e8 = threadIdx.x & 255u;
    u32 m = MAP_ID;
    for (u32 k = 0; k < 8u; k++) m = map_compose(m, ((e8 >> k) & 1u) ? MAP_EQ : MAP_NE);
    t->map8[e8] = m;
    for (u32 s0 = 0; s0 < 4u; s0++) {
        u32 st = s0, cm = 0;
        for (u32 k = 0; k < 8u; k++) {
            if (st == 3u) { cm |= 1u << k; st = 0; }
            else st = (st == 0u) ? 1u : (((e8 >> k) & 1u) ? st + 1u : 1u);
        }
        t->cls8[s0][e8] = (u16)(cm | (st << 8));
    }
    syncthreads();
}

// state map of the valid bytes of a ragged vector (first / last vector of a stream)
HC_DEV_NOINLINE u32 rle_dec_map_partial(u32 e, u32 vm)
{
    u32 m = MAP_ID;
    for (u32 k = 0; k < 16u; k++)
        if ((vm >> k) & 1u) m = map_compose(m, ((e >> k) & 1u) ? MAP_EQ : MAP_NE);
    return m;
}

// count-byte mask of the valid bytes of a ragged vector entered in state st
HC_DEV_NOINLINE u32 rle_dec_cls_partial(u32 e, u32 vm, u32 st)
{
    u32 cm = 0;
    for (u32 k = 0; k < 16u; k++) {
        if (!((vm >> k) & 1u)) continue;
        if (st == 3u) { cm |= 1u << k; st = 0; }
        else st = (st == 0u) ? 1u : (((e >> k) & 1u) ? st + 1u : 1u);
    }
    return cm;
}

// spread the low 4 bits of m to the four byte lanes of a word (0xff per set bit)
HC_DEV u32 spread4(u32 m) { return (((m & 15u) * 0x00204081u) & 0x01010101u) * 0xffu; }

// sum of the bytes of v selected by the 16-bit mask cm
HC_DEV u32 rle_masked_byte_sum(const uint4 &v, u32 cm)
{
    return dp4a_u(v.x & spread4(cm), 0x01010101u, 0u) + dp4a_u(v.y & spread4(cm >> 4), 0x01010101u, 0u) +
           dp4a_u(v.z & spread4(cm >> 8), 0x01010101u, 0u) + dp4a_u(v.w & spread4(cm >> 12), 0x01010101u, 0u);
}

// generic expansion of one vector into the current output window [w0, w1): every token drops its
// value and a head flag at its first output position inside the window
HC_DEV_NOINLINE void rle_dec_scatter(uint4 v, u32 vm, u32 cm, u32 pb, u32 o, u32 w0, u32 w1, u8 *sval, u32 *heads)
{
    u32 prev = pb;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const u32 x = i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w;
        const u32 c4 = (cm >> (4 * i)) & 15u, v4 = (vm >> (4 * i)) & 15u;
        if (c4 == 0u && v4 == 15u && o >= w0 && o + 4u <= w1) {
            // four literals inside the window
            const u32 pos = o - w0, sh = pos & 31u;
            if ((pos & 3u) == 0u) *(u32 *)(sval + pos) = x;
            else { sval[pos] = (u8)x; sval[pos + 1u] = (u8)(x >> 8); sval[pos + 2u] = (u8)(x >> 16); sval[pos + 3u] = (u8)(x >> 24); }
            atomic_or_shared(&heads[pos >> 5], 15u << sh);
            if (sh > 28u) atomic_or_shared(&heads[(pos >> 5) + 1u], 15u >> (32u - sh));
            o += 4u;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const u32 b = (x >> (8 * k)) & 0xffu;
                if ((v4 >> k) & 1u) {
                    const bool is_cnt = (c4 >> k) & 1u;
                    const u32 len = is_cnt ? b : 1u, val = is_cnt ? prev : b;
                    if (len && o < w1 && o + len > w0) {
                        const u32 pos = (o > w0 ? o : w0) - w0;
                        sval[pos] = (u8)val;
                        atomic_or_shared(&heads[pos >> 5], 1u << (pos & 31u));
                    }
                    o += len;
                }
                prev = b;
            }
        }
        prev = x >> 24;
    }
}

// the common case of rle_dec_scatter: 16 literals that all lie inside the window, at window offset pos
HC_DEV_NOINLINE void rle_dec_scatter_literals(uint4 v, u32 pos, u32 sval_addr, u32 *heads)
{
    stage_put16(sval_addr, pos, v);
    const u32 sh = pos & 31u;
    atomic_or_shared(&heads[pos >> 5], 0xffffu << sh);
    if (sh > 16u) atomic_or_shared(&heads[(pos >> 5) + 1u], 0xffffu >> (32u - sh));
}

// fills 16 output bytes from the head flags hb / values at sval + wbase, carrying the current value
struct Fill16 { uint4 r; u32 c; };
HC_DEV_NOINLINE Fill16 rle_dec_fill16(u32 hb, const u8 *sv, u32 c)
{
    u32 wv[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if ((hb >> k) & 1u) c = sv[k];
        wv[k >> 2] |= c << (8 * (k & 3));
    }
    Fill16 f;
    f.r.x = wv[0]; f.r.y = wv[1]; f.r.z = wv[2]; f.r.w = wv[3];
    f.c = c;
    return f;
}

// Decodes the n0-byte token stream at src0 (any alignment) to dst (16-byte aligned, or null to
// only measure), writing at most cap bytes; called by all TPB threads of a CTA, returns the decoded
// length.  An unaligned stream is read from the 16-byte boundary below it with the leading bytes
// masked out (they belong to the caller's buffer: a header or the previous block's tokens).
// The kernel must have called rle_dec_init() before.
HC_DEV u64 rle_decode_stream(const u8 *HC_RESTRICT src0, u64 n0, u8 *HC_RESTRICT dst, u64 cap)
{
    HC_SHARED u32 wtot[2][32];
    HC_SHARED u32 wlast[NW];
    HC_SHARED u32 heads[DEC_WIN / 32 + 1];
    HC_SHARED HC_ALIGNED16 u8 sval[DEC_WIN];
    HC_SMEM_ARENA(wtot);
    const RleDecTables *tb = rle_dec_tables();
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const u32 lead = n0 ? (u32)((uintptr_t)src0 & 15u) : 0u;
    const u8 *src = src0 - lead;
    const u64 n = n0 + lead;
    const u32 sval_addr = smem_addr(sval);
    u64 out_pos = 0;
    u32 carry_state = 0;

    uint4 cur[UN], nxt[UN];
    u32 hcur[UN], hnxt[UN];          // lane 0: the byte before its vector (prefetched with the vector)
#pragma unroll
    for (int j = 0; j < UN; j++) {
        u64 p = (u64)j * SUB_BYTES + tid * 16;
        cur[j] = p < n ? ldg16(src + p) : make_uint4_zero();
        hcur[j] = (lane == 0 && p > 0 && p < n) ? ldg8(src + p - 1) : 0u;
    }
    for (u64 t0 = 0; t0 < n; t0 += TILE_BYTES) {
#pragma unroll
        for (int j = 0; j < UN; j++) {
            u64 p = t0 + TILE_BYTES + (u64)j * SUB_BYTES + tid * 16;
            nxt[j] = p < n ? ldg16(src + p) : make_uint4_zero();
            hnxt[j] = (lane == 0 && p < n) ? ldg8(src + p - 1) : 0u;
        }
        // ---- scan 1: decoder state maps -----------------------------------------------
        u32 eq[UN], valid[UN], pbyte[UN], fmap[UN], mexcl[UN];
#pragma unroll
        for (int j = 0; j < UN; j++) {
            const u64 p = t0 + (u64)j * SUB_BYTES + tid * 16;
            u32 pb = shfl_up(cur[j].w >> 24, 1);
            if (lane == 0) pb = hcur[j];
            pbyte[j] = pb;
            u32 vm = p >= n ? 0u : (n - p >= 16 ? 0xffffu : ((1u << (u32)(n - p)) - 1u));
            if (p == 0) vm &= ~((1u << lead) - 1u);
            const u32 e = rle_eq_mask16(cur[j], pb);
            eq[j] = e;
            valid[j] = vm;
            fmap[j] = vm == 0xffffu ? map_compose(tb->map8[e & 0xffu], tb->map8[(e >> 8) & 0xffu])
                                    : (vm ? rle_dec_map_partial(e, vm) : MAP_ID);
        }
        u32 tile_map = block_scan_striped(fmap, mexcl, MAP_ID, OpCompose(), wtot[0]);

        // ---- scan 2: which bytes are counts, output length of every vector ------------
        u32 cmask[UN], cnt[UN], oexcl[UN];
#pragma unroll
        for (int j = 0; j < UN; j++) {
            const u32 s = map_apply(mexcl[j], carry_state);
            u32 cm;
            if (valid[j] == 0xffffu) {
                const u32 c0 = tb->cls8[s][eq[j] & 0xffu];
                const u32 c1 = tb->cls8[c0 >> 8][(eq[j] >> 8) & 0xffu];
                cm = (c0 & 0xffu) | ((c1 & 0xffu) << 8);
            } else {
                cm = valid[j] ? rle_dec_cls_partial(eq[j], valid[j], s) : 0u;
            }
            cmask[j] = cm;
            cnt[j] = (u32)popc(valid[j] & ~cm) + rle_masked_byte_sum(cur[j], cm);
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
                    const u32 o = shift + oexcl[j];
                    if (cnt[j] == 0u || o >= w1 || o + cnt[j] <= w0) continue;
                    if (cmask[j] == 0u && valid[j] == 0xffffu && o >= w0 && o + 16u <= w1)
                        rle_dec_scatter_literals(cur[j], o - w0, sval_addr, heads);
                    else
                        rle_dec_scatter(cur[j], valid[j], cmask[j], pbyte[j], o, w0, w1, sval, heads);
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
                const u64 capw = cap > gpos0 ? cap - gpos0 : 0;  // clip against the caller's capacity
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const u32 wbase = 64u * tid + 16u * c;
                    const u32 hb = ((c < 2 ? h0 : h1) >> (16u * (c & 1))) & 0xffffu;
                    uint4 r;
                    if (hb == 0xffffu) {                          // 16 literals: the values as they are
                        r = *(const uint4 *)(sval + wbase);
                        curv = r.w >> 24;
                    } else if (hb == 0u) {                        // inside a run
                        r.x = r.y = r.z = r.w = curv * 0x01010101u;
                    } else {
                        const Fill16 f = rle_dec_fill16(hb, sval + wbase, curv);
                        r = f.r;
                        curv = f.c;
                    }
                    if (wbase < lim) {
                        u32 a = wbase < lo_valid ? lo_valid : wbase;
                        u32 b = wbase + 16u > lim ? lim : wbase + 16u;
                        if (b > capw) b = (u32)capw;
                        if (a == wbase && b == wbase + 16u) {
                            stg16(gbase + wbase, r);
                        } else {
                            for (u32 i = a; i < b; i++) gbase[i] = vec_byte(r, (int)(i - wbase));
                        }
                    }
                }
                syncthreads();
            }
        }
        out_pos += total;
#pragma unroll
        for (int j = 0; j < UN; j++) { cur[j] = nxt[j]; hcur[j] = hnxt[j]; }
        syncthreads();
    }
    return out_pos;
}

HC_KERNEL HC_LAUNCH_BOUNDS(256, 3)
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
