// adapt_mask.cuh -- block-size search of the adaptive RLE (src/transform.cpp:294-328, :97-134) for
// matrices up to 512 x 512, evaluated from EQUALITY BITMASKS instead of pixels.
//
// The MNP-5 size of a sequence depends only on which elements equal their predecessor (runsum.cuh).
// Inside a block row (horizontal scan) that predicate does not depend on the block size at all, so:
//   phase 1 (one pass over the pixels, 8 rows per step, staged in shared memory)
//       EH[y] bit x : v[y][x] == v[y][x-1]             512 x 512 bits   (row-major)
//       EV[x] bit y : v[y][x] == v[y-1][x]             512 x 512 bits   (column-major)
//       WH_k[c] bit y : first pixel of block-row segment (y, c) == last pixel of segment (y-1, c)
//       WV_k[r] bit x : top pixel of block column segment (r, x) == bottom pixel of segment (r, x-1)
//     (the "wrap" bits WH/WV are the only ones that depend on the candidate block size 8 << k)
//   phase 2 (all 7 candidates x 2 directions from shared memory, no pixel is read again)
//       with e = the block's equality bits in scan order, per 64 elements:
//           literals = ~(e_j & e_j-1 & e_j-2),  count bytes = e_j & e_j-1 & ~e_j+1      -> 2 popcounts
//       exact while every run is shorter than 258; a block that contains an all-ones word may hold a
//       longer run and gets an exact correction from a serial walk over its words (rare: flat areas).
// The per-block results go to the same cost table as adapt_cost_kernel (min size | horizontal << 31),
// so adapt_select_kernel / adapt_emit_kernel are shared.  HBM traffic: every pixel is read once.
#pragma once
#include "adapt.cuh"

namespace hcd {

constexpr u32 ACM_MAX = 512;                 // largest width / height handled here
constexpr int ACM_TPB = 512;
constexpr int ACM_NK = 7;                    // candidates 8..512
constexpr u32 ACM_EH_STRIDE = 64;            // bytes per row of EH (512 bits)
constexpr u32 ACM_EV_STRIDE = 68;            // bytes per column of EV (padded: conflict-free word access)
constexpr u32 ACM_W_STRIDE = 64;             // bytes per WH / WV entry (512 bits)
constexpr u32 ACM_OFF_EH = 0;
constexpr u32 ACM_OFF_EV = ACM_OFF_EH + ACM_MAX * ACM_EH_STRIDE;
constexpr u32 ACM_OFF_WH = ACM_OFF_EV + ACM_MAX * ACM_EV_STRIDE;
constexpr u32 ACM_OFF_WV = ACM_OFF_WH + 128 * ACM_W_STRIDE;      // 64+32+16+8+4+2+1 = 127 entries
constexpr u32 ACM_OFF_PIX = ACM_OFF_WV + 128 * ACM_W_STRIDE;
constexpr u32 ACM_SMEM = ACM_OFF_PIX + 9 * ACM_MAX;               // 88576 bytes

HC_HD bool acm_eligible(u64 w, u64 h) { return w >= 8 && h >= 8 && w <= ACM_MAX && h <= ACM_MAX; }
// first WH/WV entry of candidate k: entries are indexed by block column / block row, at most 64 >> k
HC_HD u32 acm_wbase(int k) { return 128u - (128u >> k); }          // 0, 64, 96, 112, 120, 124, 126

// n (1..64) bits starting at bit `pos` of the little-endian bit array at shared address `base`
HC_DEV u64 acm_bits(u32 base, u32 pos, u32 n)
{
    const u32 wa = base + ((pos >> 5) << 2), sh = pos & 31u;
    const u32 w0 = lds32(wa), w1 = lds32(wa + 4u), w2 = lds32(wa + 8u);   // arrays are padded: always readable
    const u32 lo = funnel_r(w0, w1, sh), hi = funnel_r(w1, w2, sh);
    const u64 v = ((u64)hi << 32) | lo;
    return n >= 64u ? v : (v & ((1ull << n) - 1ull));
}

HC_DEV u32 acm_bit(u32 base, u32 pos) { return (lds32(base + ((pos >> 5) << 2)) >> (pos & 31u)) & 1u; }

// running evaluation of a bit sequence e[0..n) (e[0] = 0, e[n-1] forced to 0 by the caller):
// cost of all complete information so far; the end-of-run test of the newest bit is deferred
struct SeqCost {
    u32 cost;
    u32 p1, p2;      // e[j-1], e[j-2] of the next word
    u32 pend;        // the last bit seen had q >= 2: a count byte is due if the run ends there
    u32 ones;        // some full word was all ones (a run >= 258 is possible)
};

HC_DEV void sc_init(SeqCost &s) { s.cost = 0; s.p1 = s.p2 = 0; s.pend = 0; s.ones = 0; }

// consume n (1..64) bits
HC_DEV void sc_word(SeqCost &s, u64 w, u32 n)
{
    const u64 valid = n >= 64u ? ~0ull : ((1ull << n) - 1ull);
    if (s.pend && !(w & 1ull)) s.cost++;                       // previous run ended with q >= 2
    const u64 s1 = (w << 1) | s.p1, s2 = (w << 2) | ((u64)s.p1 << 1) | s.p2;
    const u64 q2 = w & s1, q3 = q2 & s2;
    s.cost += (u32)popcll(~q3 & valid);                          // literals
    s.cost += (u32)popcll(q2 & ~(w >> 1) & (valid >> 1));        // runs that end inside the word
    s.pend = (u32)((q2 >> (n - 1u)) & 1ull);
    if (n >= 2u) { s.p1 = (u32)((w >> (n - 1u)) & 1ull); s.p2 = (u32)((w >> (n - 2u)) & 1ull); }
    else { s.p2 = s.p1; s.p1 = (u32)(w & 1ull); }
    if (n >= 64u && w == ~0ull) s.ones = 1;
}

HC_DEV void sc_finish(SeqCost &s) { if (s.pend) s.cost++; }    // cannot happen after a forced final 0, kept for safety

// bit accumulator feeding sc_word with 64-bit words
struct BitAcc { u64 acc; u32 n; };
HC_DEV void ba_push(BitAcc &b, SeqCost &s, u64 bits, u32 nb)     // nb 1..64, bits masked
{
    b.acc |= b.n < 64u ? (bits << b.n) : 0ull;
    if (b.n + nb >= 64u) {
        sc_word(s, b.acc, 64);
        const u32 used = 64u - b.n;
        b.acc = used < 64u ? (bits >> used) : 0ull;
        b.n = nb - used;
    } else {
        b.n += nb;
    }
}
HC_DEV void ba_flush(BitAcc &b, SeqCost &s) { if (b.n) sc_word(s, b.acc, b.n); b.acc = 0; b.n = 0; }

// Equality bits of one scan line of a block: `inner` elements starting at bit `start` of line `line`
// of the (EH or EV) mask, first bit replaced by `first` (the wrap bit, 0 for the first line).
struct LineSrc { u32 mask_base, mask_stride; };

// one word of a block's equality sequence: line i, word j of the line (bits [64j, 64j+nb) of the line)
HC_DEV u64 acm_seq_word(u32 mask_base, u32 mask_stride, u32 wrap_base, u32 line0, u32 start, u32 i, u32 j, u32 nb,
                        bool last_of_block)
{
    u64 w = acm_bits(mask_base + (line0 + i) * mask_stride, start + 64u * j, nb);
    if (j == 0) {
        w &= ~1ull;
        if (i && acm_bit(wrap_base, line0 + i)) w |= 1ull;
    }
    if (last_of_block) w &= ~(1ull << (nb - 1u));               // forced final literal
    return w;
}

// Long-run bookkeeping, O(1) per word.  A run = a zero bit and the ones that follow it; only runs of
// >= 258 elements need a correction (rle_size(L) - 4), and such a run cannot start and end inside one
// word, so per word it is enough to close the run at the word's first zero and to restart from its last.
//   run : length of the still open run;  returns the correction for a run closed by this word
HC_DEV u32 acm_run_step(u64 w, u32 nb, u32 &run)
{
    const u64 valid = nb >= 64u ? ~0ull : ((1ull << nb) - 1ull);
    const u64 z = ~w & valid;
    if (!z) { run += nb; return 0u; }
    const u32 L = run + (u32)ffsll(z) - 1u;                      // ones before the first zero extend the open run
    run = nb - (63u - (u32)clzll(z));                            // the last zero and the ones after it
    return L >= 258u ? rle_size(L) - 4u : 0u;
}

// exact size correction for runs of >= 258 elements: sum of rle_size(L) - 4 over those runs (serial)
HC_DEV u32 acm_long_run_correction(u32 mask_base, u32 mask_stride, u32 wrap_base, u32 line0, u32 nlines,
                                   u32 start, u32 inner)
{
    u32 corr = 0, run = 0;
    for (u32 i = 0; i < nlines; i++)
        for (u32 j = 0; 64u * j < inner; j++) {
            const u32 nb = inner - 64u * j < 64u ? inner - 64u * j : 64u;
            const bool last = i == nlines - 1u && 64u * j + nb == inner;
            corr += acm_run_step(acm_seq_word(mask_base, mask_stride, wrap_base, line0, start, i, j, nb, last), nb, run);
        }
    return corr;                                                  // the forced final 0 closed every long run
}

// the same for lines that are multiples of 64 bits, spread over a warp: every lane takes a contiguous
// range of the block's words; the open run entering a lane's range comes from a warp scan
HC_DEV u32 acm_long_run_correction_warp(u32 mask_base, u32 mask_stride, u32 wrap_base, u32 line0, u32 nlines,
                                        u32 start, u32 inner, u32 lane)
{
    const u32 wpl = inner >> 6, total = nlines * wpl, per = (total + 31u) / 32u;
    u32 lo = lane * per, hi = lo + per;
    if (lo > total) lo = total;
    if (hi > total) hi = total;
    // pass over the own range with an unknown entering run: remember where the first zero is
    u32 corr = 0, run = 0, head = 0;
    bool hasz = false;
    for (u32 t = lo; t < hi; t++) {
        const u32 i = t / wpl, j = t % wpl;
        const u64 w = acm_seq_word(mask_base, mask_stride, wrap_base, line0, start, i, j, 64, t == total - 1u);
        if (!hasz) {
            const u64 z = ~w;
            if (z) { hasz = true; head = run + (u32)ffsll(z) - 1u; run = 64u - (63u - (u32)clzll(z)); }
            else run += 64u;
        } else {
            corr += acm_run_step(w, 64, run);
        }
    }
    // entering run of every lane: carry(l) = hasz(l-1) ? tail(l-1) : carry(l-1) + len(l-1)
    u32 hz = hasz ? 1u : 0u, val = run;                          // val = tail if hasz, else the range length
    for (u32 d = 1; d < 32u; d <<= 1) {
        const u32 phz = shfl_up(hz, d), pval = shfl_up(val, d);
        if (lane >= d && !hz) { val += pval; hz = phz; }
    }
    u32 carry = shfl_up(val, 1);
    if (lane == 0) carry = 0;
    if (hasz) { const u32 L = carry + head; if (L >= 258u) corr += rle_size(L) - 4u; }
    for (int d = 16; d > 0; d >>= 1) corr += shfl_xor(corr, d);
    return corr;
}

// cost of one block in one direction, evaluated serially by the calling thread
HC_DEV u32 acm_block_cost_serial(u32 mask_base, u32 mask_stride, u32 wrap_base, u32 line0, u32 nlines,
                                 u32 start, u32 inner)
{
    SeqCost s;
    BitAcc b;
    sc_init(s);
    b.acc = 0; b.n = 0;
    const u32 n = nlines * inner;
    u32 pos = 0;
    for (u32 i = 0; i < nlines; i++) {
        const u32 lb = mask_base + (line0 + i) * mask_stride;
        for (u32 j = 0; j < inner; j += 64u) {
            const u32 nb = inner - j < 64u ? inner - j : 64u;
            u64 w = acm_bits(lb, start + j, nb);
            if (j == 0) {
                w &= ~1ull;
                if (i && acm_bit(wrap_base, line0 + i)) w |= 1ull;
            }
            if (pos + nb == n) w &= ~(1ull << (nb - 1u));       // forced final literal
            pos += nb;
            ba_push(b, s, w, nb);
        }
    }
    ba_flush(b, s);
    sc_finish(s);
    if (s.ones) s.cost += acm_long_run_correction(mask_base, mask_stride, wrap_base, line0, nlines, start, inner);
    return s.cost;
}

// full B x B block with B in {8, 16, 32}: every scan line is one aligned byte / half word / word of the
// mask, the wrap bits of the B lines sit in one aligned byte / half word / word too
HC_DEV u32 acm_block_cost_small(u32 mask_base, u32 mask_stride, u32 wrap_base, u32 line0, u32 start, u32 B)
{
    SeqCost s;
    sc_init(s);
    const u32 lb = mask_base + line0 * mask_stride + (start >> 3);       // byte address of line 0's bits
    if (B == 8u) {
        const u32 wb = lds8(wrap_base + (line0 >> 3)) & 0xfeu;           // wrap bits of lines 1..7
        u64 w = 0;
#pragma unroll
        for (u32 i = 0; i < 8u; i++)
            w |= (u64)((lds8(lb + i * mask_stride) & 0xfeu) | ((wb >> i) & 1u)) << (8u * i);
        sc_word(s, w & ~(1ull << 63), 64);
        return s.cost;
    }
    if (B == 16u) {
        const u32 wb = lds16(wrap_base + (line0 >> 3)) & 0xfffeu;
        for (u32 j = 0; j < 4u; j++) {
            u64 w = 0;
#pragma unroll
            for (u32 i = 0; i < 4u; i++) {
                const u32 ln = 4u * j + i;
                w |= (u64)((lds16(lb + ln * mask_stride) & 0xfffeu) | ((wb >> ln) & 1u)) << (16u * i);
            }
            if (j == 3u) w &= ~(1ull << 63);
            sc_word(s, w, 64);
        }
        return s.cost;
    }
    const u32 wb = lds32(wrap_base + (line0 >> 3)) & ~1u;                // B == 32
    for (u32 j = 0; j < 16u; j++) {
        const u32 l0 = 2u * j, l1 = l0 + 1u;
        u64 w = (u64)((lds32(lb + l0 * mask_stride) & ~1u) | ((wb >> l0) & 1u)) |
                ((u64)((lds32(lb + l1 * mask_stride) & ~1u) | ((wb >> l1) & 1u)) << 32);
        if (j == 15u) w &= ~(1ull << 63);
        sc_word(s, w, 64);
    }
    if (s.ones) s.cost += acm_long_run_correction(mask_base, mask_stride, wrap_base, line0, B, start, B);
    return s.cost;
}

// cost of one block in one direction with the lines spread over the lanes of a warp (lines >= 64 bits
// wide, 64-bit aligned start): every line is evaluated independently from its own bits plus the two
// last bits of the previous line and the first bit of the next one.
HC_DEV u32 acm_block_cost_warp(u32 mask_base, u32 mask_stride, u32 wrap_base, u32 line0, u32 nlines,
                               u32 start, u32 inner, u32 lane)
{
    u32 cost = 0, ones = 0;
    const u32 nwords = inner >> 6;                                  // inner is a multiple of 64 here
    for (u32 i = lane; i < nlines; i += 32u) {
        const u32 lb = mask_base + (line0 + i) * mask_stride;
        SeqCost s;
        sc_init(s);
        if (i) {                                                    // context from the previous line
            const u32 pb = mask_base + (line0 + i - 1u) * mask_stride;
            const u64 t = acm_bits(pb, start + inner - 2u, 2);
            s.p2 = (u32)(t & 1ull);
            s.p1 = (u32)(t >> 1);
            // pend: was the last element of the previous line at q >= 2?  (its two last bits are set;
            // for a 1-word... lines are >= 64 wide so both bits belong to that line)
            s.pend = s.p1 & s.p2;
        }
        for (u32 j = 0; j < nwords; j++) {
            u64 w = acm_bits(lb, start + 64u * j, 64);
            if (j == 0) {
                w &= ~1ull;
                if (i && acm_bit(wrap_base, line0 + i)) w |= 1ull;
            }
            if (i == nlines - 1u && j == nwords - 1u) w &= ~(1ull << 63);   // forced final literal
            sc_word(s, w, 64);
        }
        // the deferred end test of this line's last bit belongs to the next line (its `pend`); the last
        // line ends with a forced 0, so nothing is pending there
        cost += s.cost;
        ones |= s.ones;
    }
    for (int d = 16; d > 0; d >>= 1) { cost += shfl_xor(cost, d); ones |= shfl_xor(ones, d); }
    if (ones) cost += acm_long_run_correction_warp(mask_base, mask_stride, wrap_base, line0, nlines, start, inner, lane);
    return cost;
}

HC_KERNEL HC_LAUNCH_BOUNDS(ACM_TPB, 2)
adapt_cost_mask_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT width,
                       const u64 *HC_RESTRICT height, u32 nf, u32 *HC_RESTRICT cost, u64 cost_stride)
{
    HC_DYN_SMEM(smem);
    HC_SMEM_ARENA(*smem);
    const u32 sb = smem_addr(smem);
    const u32 tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    u8 *pix = smem + ACM_OFF_PIX;                                   // [9][512]

    for (u32 f = blockIdx.x; f < nf; f += gridDim.x) {
        const u64 w64 = width[f], h64 = height[f];
        if (!acm_eligible(w64, h64)) continue;
        const u32 W = (u32)w64, H = (u32)h64;
        const u8 *mat = in + in_off[f];
        // clear the masks (rows / columns beyond W, H and the padding must read as zero)
        for (u32 i = tid; i < ACM_OFF_PIX / 16u; i += ACM_TPB) {
            uint2 z; z.x = 0; z.y = 0;
            sts64(sb + 16u * i, z);
            sts64(sb + 16u * i + 8u, z);
        }
        syncthreads();

        // ---------------- phase 1: masks, 8 rows per step ----------------
        const u32 x = tid;                                          // this thread's column
        u32 topv[ACM_NK];                                           // top pixel of the current block row, per candidate
#pragma unroll
        for (int k = 0; k < ACM_NK; k++) topv[k] = 0x100u;
        for (u32 y0 = 0; y0 < H; y0 += 8u) {
            // stage rows y0..y0+7 into pix[1..8]; pix[0] keeps row y0-1
            if (y0 && x < W) pix[x] = pix[8u * ACM_MAX + x];
            syncthreads();
            const u32 rows = H - y0 < 8u ? H - y0 : 8u;
            if ((W & 15u) == 0) {
                for (u32 i = tid; i < rows * (W >> 4); i += ACM_TPB) {
                    const u32 r = i / (W >> 4), cx = (i % (W >> 4)) << 4;
                    *(uint4 *)(pix + (r + 1u) * ACM_MAX + cx) = ldg16(mat + (u64)(y0 + r) * W + cx);
                }
            } else {
                for (u32 i = tid; i < rows * W; i += ACM_TPB) {
                    const u32 r = i / W, cx = i % W;
                    pix[(r + 1u) * ACM_MAX + cx] = ldg8(mat + (u64)(y0 + r) * W + cx);
                }
            }
            syncthreads();
            // EH (one ballot per row) and EV (8 rows -> one byte per column)
            u32 evb = 0;
            for (u32 r = 0; r < 8u; r++) {
                const u32 y = y0 + r;
                const bool in_img = y < H && x < W;
                const u32 cur = in_img ? pix[(r + 1u) * ACM_MAX + x] : 0x100u;
                const u32 left = (in_img && x) ? pix[(r + 1u) * ACM_MAX + x - 1u] : 0x200u;
                const u32 up = (in_img && y) ? pix[r * ACM_MAX + x] : 0x300u;
                const u32 ehw = ballot(cur == left);
                if (lane == 0 && y < H) sts32(sb + ACM_OFF_EH + y * ACM_EH_STRIDE + 4u * wid, ehw);
                if (cur == up) evb |= 1u << r;
            }
            if (x < W) sts8(sb + ACM_OFF_EV + x * ACM_EV_STRIDE + (y0 >> 3), evb);
            // vertical wrap bits: block rows start and end on strip boundaries (strips are 8 rows, B >= 8)
            {
                const u32 ylast = y0 + rows - 1u;                   // last row of this strip
                const u32 top_px = x < W ? pix[ACM_MAX + x] : 0x100u;
                const u32 bot_left = (x && x < W) ? pix[rows * ACM_MAX + x - 1u] : 0x200u;
#pragma unroll
                for (int k = 0; k < ACM_NK; k++) {
                    const u32 B = 8u << k;
                    if (B <= W && B <= H) {
                        if ((y0 & (B - 1u)) == 0u) topv[k] = top_px;
                        if (((ylast + 1u) & (B - 1u)) == 0u || ylast == H - 1u) {     // block row ends here (uniform)
                            const u32 wvw = ballot(topv[k] == bot_left);
                            if (lane == 0)
                                sts32(sb + ACM_OFF_WV + (acm_wbase(k) + (ylast >> (3 + k))) * ACM_W_STRIDE + 4u * wid, wvw);
                        }
                    }
                }
            }
            // horizontal wrap bits: one thread per (candidate, block column), 8 rows -> one byte
            if (tid < 128u) {
                int k = 0;
                while (k < ACM_NK - 1 && tid >= acm_wbase(k + 1)) k++;
                const u32 B = 8u << k, c = tid - acm_wbase(k), bx = c * B;
                if (B <= W && B <= H && bx < W) {
                    const u32 xe = (bx + B < W ? bx + B : W) - 1u;
                    u32 wb = 0;
                    for (u32 r = 0; r < rows; r++)
                        if ((y0 + r) && pix[(r + 1u) * ACM_MAX + bx] == pix[r * ACM_MAX + xe]) wb |= 1u << r;
                    sts8(sb + ACM_OFF_WH + tid * ACM_W_STRIDE + (y0 >> 3), wb);
                }
            }
            syncthreads();                                          // pix is overwritten by the next step
        }

        // ---------------- phase 2: every candidate, both directions ----------------
        u32 *ctab = cost + (u64)f * cost_stride;
        u64 kb = 0;
        for (int k = 0; k < ACM_NK; k++) {
            const u32 B = 8u << k;
            if (k && (B > W || B > H)) break;
            const u32 ncb = (W + B - 1u) / B, nbr = (H + B - 1u) / B, nb = ncb * nbr;
            u32 *tab = ctab + kb;
            kb += nb;
            if (k <= 2) {
                // small blocks: one thread per block and direction.  The two directions use different
                // thread -> block mappings (lanes along a block row for the horizontal masks, along a
                // block column for the vertical ones) so that shared-memory accesses stay conflict free;
                // the horizontal result waits in the cost table.
                for (u32 i = tid; i < nb; i += ACM_TPB) {
                    const u32 c = i % ncb, r = i / ncb;
                    const u32 bx = c * B, by = r * B;
                    const u32 bw = bx + B > W ? W - bx : B, bh = by + B > H ? H - by : B;
                    const u32 wrap = sb + ACM_OFF_WH + (acm_wbase(k) + c) * ACM_W_STRIDE;
                    tab[r * ncb + c] = (bw == B && bh == B)
                        ? acm_block_cost_small(sb + ACM_OFF_EH, ACM_EH_STRIDE, wrap, by, bx, B)
                        : acm_block_cost_serial(sb + ACM_OFF_EH, ACM_EH_STRIDE, wrap, by, bh, bx, bw);
                }
                syncthreads();                                      // each thread re-reads another thread's entry
                for (u32 i = tid; i < nb; i += ACM_TPB) {
                    const u32 r = i % nbr, c = i / nbr;
                    const u32 bx = c * B, by = r * B;
                    const u32 bw = bx + B > W ? W - bx : B, bh = by + B > H ? H - by : B;
                    const u32 wrap = sb + ACM_OFF_WV + (acm_wbase(k) + r) * ACM_W_STRIDE;
                    const u32 cv = (bw == B && bh == B)
                        ? acm_block_cost_small(sb + ACM_OFF_EV, ACM_EV_STRIDE, wrap, bx, by, B)
                        : acm_block_cost_serial(sb + ACM_OFF_EV, ACM_EV_STRIDE, wrap, bx, bw, by, bh);
                    const u32 ch = tab[r * ncb + c];
                    tab[r * ncb + c] = ch <= cv ? (ch | 0x80000000u) : cv;
                }
            } else {
                // large blocks: one warp per block, lines spread over the lanes
                for (u32 blk = wid; blk < nb; blk += ACM_TPB / 32) {
                    const u32 c = blk % ncb, r = blk / ncb;
                    const u32 bx = c * B, by = r * B;
                    const u32 bw = bx + B > W ? W - bx : B, bh = by + B > H ? H - by : B;
                    const u32 whb_ = sb + ACM_OFF_WH + (acm_wbase(k) + c) * ACM_W_STRIDE;
                    const u32 wvb_ = sb + ACM_OFF_WV + (acm_wbase(k) + r) * ACM_W_STRIDE;
                    u32 ch, cv;
                    if ((bw & 63u) == 0u) ch = acm_block_cost_warp(sb + ACM_OFF_EH, ACM_EH_STRIDE, whb_, by, bh, bx, bw, lane);
                    else { ch = lane == 0 ? acm_block_cost_serial(sb + ACM_OFF_EH, ACM_EH_STRIDE, whb_, by, bh, bx, bw) : 0u; ch = shfl(ch, 0); }
                    if ((bh & 63u) == 0u) cv = acm_block_cost_warp(sb + ACM_OFF_EV, ACM_EV_STRIDE, wvb_, bx, bw, by, bh, lane);
                    else { cv = lane == 0 ? acm_block_cost_serial(sb + ACM_OFF_EV, ACM_EV_STRIDE, wvb_, bx, bw, by, bh) : 0u; cv = shfl(cv, 0); }
                    if (lane == 0) tab[blk] = ch <= cv ? (ch | 0x80000000u) : cv;
                }
            }
        }
        syncthreads();
    }
}

}  // namespace hcd
