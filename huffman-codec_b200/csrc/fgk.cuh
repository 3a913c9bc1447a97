// fgk.cuh -- batched adaptive Huffman (FGK) encode / decode.
// Reference: HuffTree::{encode,decode,update} src/huffman.cpp:37-128, findSuccNode :157-184,
// swapNodes :186-217, drivers applyHuffman/revertHuffman src/transform.cpp:363-406, container
// header src/headers.cpp:107-125 and bit packing src/main.cpp:78-84.
//
// The tree update is serial inside one stream, so the parallelism is the batch: ONE WARP PER FILE
// with the tree in shared memory (8.7 KB per stream).  A lone stream is bound by the chain of
// dependent instructions of its warp, a full batch by instruction issue (7 warps per scheduler),
// so both the chain and the instruction count matter.  Measured on B200 (tools/ubench_latency.cu):
// LDS 24 cycles, SHFL 33, ballot 25, ballot -> clz -> shfl 73, REDUX 28.
//
//   node handle = byte offset o = 8 * node number: siblings adjacent, even number =
//   left child = bit 0, weights non-decreasing in number order, so the block leader of a node is the
//   last node l >= it with the same weight.
//   rec[o / 8] = { weight, parent offset | down << 16 }      one LDS.64 per node
//       down   = internal: offset of the left child;  leaf: 0x8000 | symbol (NYT: 0x8100)
//   The root's weight is never consulted by the algorithm except as "the node after 511", where a
//   tie can only mean "the leader is my parent" = no swap; it is kept at 0xffffffff (never ties).
//   PATH TABLE  pt[(2 << j) - 2 + p] = offset of the node reached from the root by the (j+1)-bit
//   path p (FGK_ROOT_O: none), for the top FGK_D levels; pfx[node] = depth << 12 | path, the path
//   LEFT ALIGNED in 9 bits (0xffff for deeper nodes).  The table describes node NUMBERS, so the most
//   frequent swap -- two leaves exchange their symbols -- leaves it valid.
//   spf[symbol] = leaf offset | pfx of that leaf << 16: one load gives the encoder a symbol's code.
//
//   One symbol (fgk_update_rounds): lane j owns the node at depth j + 1 of the leaf's path.  ROUND:
//   every lane loads its node, the next node and the weight after that in one batch of independent
//   loads, decides locally "tie with the next node -> leader search / swap needed" and what the swap
//   would be, and ONE REDUX.MAX over (lane << 27 | swap description) tells all lanes the deepest
//   tied level, whether it swaps, and the path of the leader.  The lanes below it store weight + 1,
//   the tied lane does its own swap stores, and only if the walk moved to another branch the lanes
//   look up the new path (one table load) and start the next round.  A run of three or more equal
//   weights (2/3 of all ties on high-entropy data) is resolved by all lanes together: a second
//   REDUX (issued with the first) carries the tied node's offset, 32 lanes probe the run in one
//   load + ballot, and the swap is then described by uniform loads -- no shuffles anywhere.
//   Nodes deeper than the table and first occurrences (NYT split) use sequential code (fgk_update_seq).
//
//   Nothing written in a round is read again in a later round of the same symbol (parents, leaders
//   and probes all have larger numbers than what was written), every load of a round is separated
//   from the round's stores by a warp barrier, and one warp barrier per symbol publishes the stores.
//
// Encoder bit output: a table code (<= 9 bits) is parked in the lane whose index is the symbol
// count mod 32; every 32 symbols the warp packs them at once (shuffle scan of the lengths, shared
// memory atomic OR into a 64-word ring, 128-byte coalesced flushes).  Long codes and the 9-byte
// container header go through the same ring serially.
#pragma once
#include "hc_common.cuh"

namespace hcd {

#ifndef HC_FGK_ENC_WARPS
#define HC_FGK_ENC_WARPS 1
#endif
#ifndef HC_FGK_DEC_WARPS
#define HC_FGK_DEC_WARPS 4
#endif
// streams per CTA (one-warp CTAs give their shared memory back as soon as their stream ends)
constexpr int FGK_ENC_WARPS = HC_FGK_ENC_WARPS;
constexpr int FGK_DEC_WARPS = HC_FGK_DEC_WARPS;
constexpr int FGK_WARPS = FGK_ENC_WARPS > FGK_DEC_WARPS ? FGK_ENC_WARPS : FGK_DEC_WARPS;
constexpr u32 FGK_NREC = 516;                    // nodes 0..512, sentinel 513, two pad records
constexpr u32 FGK_ROOT_O = 8u * 512u;            // also "no node" in pt (the root itself is never listed)
constexpr u32 FGK_SENT_O = 8u * 513u;            // weight 0xffffffff: ends every leader search
constexpr u32 FGK_NONE_O = 8u * 515u;            // leaf offset of a symbol not yet transmitted
constexpr u32 FGK_LEAF = 0x8000u;
constexpr u32 FGK_LEAF_NYT = 0x8100u;
constexpr u32 FGK_D = 9;                         // levels covered by the path table
constexpr u32 FGK_PT_N = (2u << FGK_D) - 2u;     // 2 + 4 + ... + 2^D entries
constexpr u32 FGK_NOPATH = 0xffffu;
constexpr u32 FGK_SPF_NONE = FGK_NONE_O | (FGK_NOPATH << 16);

struct HC_ALIGNED16 FgkTree {
    uint2 rec[FGK_NREC];     // {weight, parent | down << 16}
    u32 spf[256];            // leaf offset | pfx << 16 of a symbol, FGK_SPF_NONE = not yet transmitted
    u16 pfx[FGK_NREC];       // depth << 12 | left-aligned path
    u16 pt[FGK_PT_N + 2];
    u32 ring[64];            // encoder: bit output ring (two halves of 32 words)
    u8 buf[256];             // symbol staging: two halves of 128 (encoder: ring; decoder: first half)
    u8 pad[8];
};

static_assert(sizeof(FgkTree) % 16 == 0, "trees are 16-byte aligned");
static_assert(sizeof(FgkTree) * FGK_WARPS <= 48u * 1024u, "the trees of a CTA are static shared memory");

struct FgkCtx {              // shared addresses (identical in every lane) + per-lane table constants
    u32 rec, spf, pfx, pt, buf, ring;
    u32 nyt;                 // offset of the NYT node
    u32 lev;                 // number of populated levels of pt
    u32 err;                 // set when a walk does not end (cannot happen on a consistent tree): the stream fails with status 102
    u32 ptj, shj;            // lane j: address of its level of pt, shift that turns a left-aligned path into its index
};

HC_DEV u32 fgk_parent(u32 pd) { return pd & 0xffffu; }
HC_DEV u32 fgk_down(u32 pd) { return pd >> 16; }

HC_DEV void fgk_init(FgkCtx &c, FgkTree &t, u32 lane)
{
    c.rec = smem_addr(&t.rec[0]);
    c.spf = smem_addr(&t.spf[0]);
    c.pfx = smem_addr(&t.pfx[0]);
    c.pt = smem_addr(&t.pt[0]);
    c.buf = smem_addr(&t.buf[0]);
    c.ring = smem_addr(&t.ring[0]);
    c.nyt = FGK_ROOT_O;
    c.lev = 0;
    c.err = 0;
    c.ptj = c.pt + (lane < FGK_D ? 2u * ((2u << lane) - 2u) : 0u);
    c.shj = lane < FGK_D ? FGK_D - 1u - lane : 0u;
    for (u32 i = lane; i < 256u; i += 32) sts32(c.spf + 4u * i, FGK_SPF_NONE);
    for (u32 i = lane; i < (FGK_PT_N + 2u) / 2u; i += 32) sts32(c.pt + 4u * i, FGK_ROOT_O | (FGK_ROOT_O << 16));
    for (u32 i = lane; i < FGK_NREC / 2u; i += 32) sts32(c.pfx + 4u * i, 0xffffffffu);
    sts32(c.ring + 4u * lane, 0u);
    sts32(c.ring + 128u + 4u * lane, 0u);
    if (lane == 0) {
        uint2 z; z.x = 0xffffffffu; z.y = FGK_LEAF_NYT << 16;
        sts64(c.rec + FGK_ROOT_O, z);
        z.y = 0;
        sts64(c.rec + FGK_SENT_O, z);
        sts64(c.rec + FGK_SENT_O + 8u, z);
        sts64(c.rec + FGK_SENT_O + 16u, z);
    }
    syncwarp();
}

// per-lane node of a path: lane j < depth gets the offset of the node at depth j + 1.  The other
// lanes get whatever lies along the bits below it, or the root; callers mask them by depth.
HC_DEV u32 fgk_lookup(const FgkCtx &c, u32 pf)
{
    return lds16(c.ptj + 2u * ((pf & 0x1ffu) >> c.shj));
}

// pfx of a node; a leaf's symbol entry carries a copy
HC_DEV void fgk_set_pfx(const FgkCtx &c, u32 o, u32 pf)
{
    sts16(c.pfx + (o >> 2), pf);
    const u32 kd = fgk_down(lds32(c.rec + o + 4u));

    if ((kd & FGK_LEAF) && kd != FGK_LEAF_NYT) sts16(c.spf + 4u * (kd & 0xffu) + 2u, pf);
}

// Rebuilds the path table from the tree, one level per step, the nodes of a level spread over the
// lanes.  Called by all lanes after the tree changed shape.
HC_DEV void fgk_rebuild(FgkCtx &c, u32 lane)
{
    syncwarp();                                           // the tree writes are visible
    for (u32 o = c.nyt + 8u * lane; o < FGK_ROOT_O; o += 256u) fgk_set_pfx(c, o, FGK_NOPATH);
    syncwarp();
    u32 lev = 0;
    for (u32 d = 1; d <= FGK_D; d++) {
        const u32 base = (1u << d) - 2u, pbase = (1u << (d - 1u)) - 2u;
        u32 any = 0;
        for (u32 p = lane; p < (1u << d); p += 32) {
            const u32 pe = d == 1u ? FGK_ROOT_O : lds16(c.pt + 2u * (pbase + (p >> 1)));
            u32 e = FGK_ROOT_O;
            if (d == 1u || pe != FGK_ROOT_O) {
                const u32 kd = fgk_down(lds32(c.rec + pe + 4u));
                if (!(kd & FGK_LEAF)) e = kd + 8u * (p & 1u);                 // the child
            }
            sts16(c.pt + 2u * (base + p), e);
            if (e != FGK_ROOT_O) { fgk_set_pfx(c, e, (d << 12) | (p << (FGK_D - d))); any = 1; }
        }
        syncwarp();
        if (ballot(any != 0u) == 0u) {
            // level d is empty (and written as such); clear what an earlier, deeper tree left below
            for (u32 d2 = d + 1u; d2 <= c.lev; d2++)
                for (u32 p = lane; p < (1u << d2); p += 32) sts16(c.pt + 2u * ((1u << d2) - 2u + p), FGK_ROOT_O);
            break;
        }
        lev = d;
    }
    c.lev = lev;
    syncwarp();
}

// After nodes x and y exchanged their subtrees only the table entries BELOW them change (the nodes
// keep their own paths).  px / py = their pfx entries (FGK_NOPATH: that node lies deeper than the
// table, nothing below it is listed).  First every node listed below either one loses its path,
// then both ranges are re-derived level by level.
HC_DEV void fgk_rebuild_pair(FgkCtx &c, u32 px, u32 py, u32 lane)
{
    syncwarp();                                           // the tree writes are visible
#pragma unroll 1
    for (u32 side = 0; side < 2u; side++) {
        const u32 pf = side ? py : px;
        if (pf == FGK_NOPATH) continue;
        const u32 depth = pf >> 12, path = (pf & 0x1ffu) >> (FGK_D - depth);
        for (u32 d = depth + 1u; d <= FGK_D; d++) {
            const u32 n = 1u << (d - depth), base = (1u << d) - 2u + (path << (d - depth));
            for (u32 i = lane; i < n; i += 32) {
                const u32 e = lds16(c.pt + 2u * (base + i));
                if (e != FGK_ROOT_O) fgk_set_pfx(c, e, FGK_NOPATH);
            }
        }
    }
    syncwarp();
    u32 lev = c.lev;
#pragma unroll 1
    for (u32 side = 0; side < 2u; side++) {
        const u32 pf = side ? py : px;
        if (pf == FGK_NOPATH) continue;
        const u32 depth = pf >> 12, path = (pf & 0x1ffu) >> (FGK_D - depth);
        for (u32 d = depth + 1u; d <= FGK_D; d++) {
            const u32 n = 1u << (d - depth), p0 = path << (d - depth);
            const u32 base = (1u << d) - 2u, pbase = (1u << (d - 1u)) - 2u;
            u32 any = 0;
            for (u32 i = lane; i < n; i += 32) {
                const u32 p = p0 + i;
                const u32 pe = lds16(c.pt + 2u * (pbase + (p >> 1)));
                u32 e = FGK_ROOT_O;
                if (pe != FGK_ROOT_O) {
                    const u32 kd = fgk_down(lds32(c.rec + pe + 4u));
                    if (!(kd & FGK_LEAF)) e = kd + 8u * (p & 1u);
                }
                sts16(c.pt + 2u * (base + p), e);
                if (e != FGK_ROOT_O) { fgk_set_pfx(c, e, (d << 12) | (p << (FGK_D - d))); any = 1; }
            }
            syncwarp();
            if (ballot(any != 0u) && d > lev) lev = d;
        }
    }
    c.lev = lev;
}

// out-of-line form for the update's hot loop (an internal node moves about once per 100 symbols)
HC_DEV_NOINLINE u32 fgk_rebuild_pair_cold(FgkCtx c, u32 px, u32 py, u32 lane)
{
    fgk_rebuild_pair(c, px, py, lane);
    return c.lev;
}

// table entries of the two nodes created by an NYT split of node n; pn = pfx of n before the
// split (depth 0 for the root)
HC_DEV void fgk_table_split(FgkCtx &c, u32 n, u32 pn, u32 lane)
{
    if (pn == FGK_NOPATH) return;                         // deeper than the table: the new nodes stay unlisted
    const u32 depth = pn >> 12, path = (pn & 0x1ffu) >> (FGK_D - depth);
    if (depth >= FGK_D) return;
    const u32 d = depth + 1u;                             // children: n - 16 (bit 0) and n - 8 (bit 1)
    if (lane < 2u) {
        const u32 p = (path << 1) | lane, child = n - 16u + 8u * lane;
        sts16(c.pt + 2u * ((1u << d) - 2u + p), child);
        fgk_set_pfx(c, child, (d << 12) | (p << (FGK_D - d)));
    }
    if (d > c.lev) c.lev = d;
    syncwarp();
}

// NYT split (src/huffman.cpp:99-111): the NYT node n becomes internal with children n-2 (new NYT)
// and n-1 (leaf of `sym`).  Returns the offset of the new leaf.
HC_DEV u32 fgk_split(FgkCtx &c, u32 sym, u32 lane)
{
    const u32 n = c.nyt;
    syncwarp();                               // every lane has done its reads of the old tree
    if (lane == 0) {
        uint2 z; z.x = 0; z.y = n | ((FGK_LEAF | sym) << 16);
        sts64(c.rec + n - 8u, z);             // leaf: weight 0, parent n
        z.y = n | (FGK_LEAF_NYT << 16);
        sts64(c.rec + n - 16u, z);            // new NYT
        sts16(c.rec + n + 6u, n - 16u);       // n is internal now: its left child
        sts32(c.spf + 4u * sym, (n - 8u) | (FGK_NOPATH << 16));
    }
    c.nyt = n - 16u;
    syncwarp();
    return n - 8u;
}

// collect one code bit (node number odd = right child = 1) at the top of the 64-bit accumulator hi:lo
HC_DEV void fgk_code_bit(u32 o, u32 &hi, u32 &lo)
{
    lo = funnel_r(lo, hi, 1);
    hi = (hi >> 1) | ((o << 28) & 0x80000000u);
}

// code of node o in the current tree: parent chase to the root.  hi:lo enter as 0x80000000:0 (the
// marker bit ends up below the last code bit).
HC_DEV void fgk_code_of(const FgkCtx &c, u32 o, u32 &hi, u32 &lo)
{
    for (u32 guard = 0; o < FGK_ROOT_O && guard < 600u; guard++, o = fgk_parent(lds32(c.rec + o + 4u))) fgk_code_bit(o, hi, lo);
}

// last node of the run of weight W: `from` = first candidate (the nodes before it are known to carry W)
HC_DEV u32 fgk_leader(const FgkCtx &c, u32 from, u32 W, u32 lane)
{
    u32 run;
    do {
        u32 pa = from + 8u * lane;
        pa = pa < FGK_SENT_O ? pa : FGK_SENT_O;
        const u32 m = ~ballot(lds32(c.rec + pa) == W);
        run = m ? (u32)ffs(m) - 1u : 32u;
        from += 8u * run;
    } while (run == 32u);
    return from - 8u;
}

// re-attach what hangs below a node that moved to offset `to` (whose pfx is pf_to): a leaf's symbol
// entry, or the parent links of the two children
HC_DEV void fgk_relink(const FgkCtx &c, u32 k, u32 to, u32 pf_to)
{
    if (k & FGK_LEAF) {
        if (k != FGK_LEAF_NYT) sts32(c.spf + 4u * (k & 0xffu), to | (pf_to << 16));
    } else {
        sts16(c.rec + k + 4u, to);
        sts16(c.rec + k + 12u, to);
    }
}

// Sequential FGK update from node a (src/huffman.cpp:113-127): all lanes walk together, lane 0
// writes.  For the rare callers: nodes deeper than the path table, the first occurrence of a symbol,
// the continuation after a swap with such a node.  Out of line (one copy instead of one per call
// site keeps the kernels inside the instruction cache); the context travels by value so that the
// caller's copy can stay in registers.
HC_DEV_NOINLINE FgkCtx fgk_update_seq(FgkCtx c, u32 a, u32 lane)
{
    bool sc = false;
    u32 guard = 0;
    while (a != FGK_ROOT_O) {
        if (++guard > 600u || a > FGK_ROOT_O) { c.err = 1; return c; }     // more levels than nodes: never on a consistent tree
        const uint2 r = lds64(c.rec + a);
        const u32 W = r.x;
        u32 parent = fgk_parent(r.y);
        u32 l = a;
        if (lds32(c.rec + a + 8u) == W) l = fgk_leader(c, a + 16u, W, lane);
        syncwarp();                                   // every lane has read this level
        if (l != a && l != parent) {
            // exchange the contents of a and l (src/huffman.cpp:186-217)
            const u32 ka = fgk_down(r.y), pl = lds32(c.rec + l + 4u), kl = fgk_down(pl);
            const u32 pfa = lds16(c.pfx + (a >> 2)), pfl = lds16(c.pfx + (l >> 2));
            syncwarp();
            if (lane == 0) {
                sts16(c.rec + a + 6u, kl);
                sts16(c.rec + l + 6u, ka);
                fgk_relink(c, kl, a, pfa);
                fgk_relink(c, ka, l, pfl);
            }
            if (kl == FGK_LEAF_NYT) c.nyt = a;
            if (ka == FGK_LEAF_NYT) c.nyt = l;
            if (!(ka & kl & FGK_LEAF)) sc = true;     // an internal node moved: the path table is stale
            a = l;
            parent = fgk_parent(pl);
        }
        if (lane == 0) sts32(c.rec + a, W + 1u);
        a = parent;
        syncwarp();
    }
    if (sc) fgk_rebuild(c, lane);
    return c;
}

// what a tied lane tells the others through the REDUX (larger lane = deeper level wins)
constexpr u32 FGK_I_NOSWAP = 1u << 16;    // the leader is the node's parent: plain increment
constexpr u32 FGK_I_SAMEP = 1u << 17;     // leader and node share the parent: the path above is unchanged
constexpr u32 FGK_I_SC = 1u << 18;        // an internal node moves: table entries below the two nodes change
constexpr u32 FGK_I_LONG = 1u << 20;      // more than two equal weights: leader search (bits 0..15 = the node, not a pfx)
constexpr u32 FGK_I_TIE = 1u << 26;

struct FgkPre {          // what a lane loads about its node A per round
    uint2 n;             // rec[A]
    uint2 n1;            // rec[A + 8]
    u32 w2;              // weight of A + 16
    u32 pfl;             // pfx[A + 8]
};

HC_DEV FgkPre fgk_preload(const FgkCtx &c, u32 A)
{
    FgkPre p;
    p.n = lds64(c.rec + A);
    p.n1 = lds64(c.rec + A + 8u);
    p.w2 = lds32(c.rec + A + 16u);
    p.pfl = lds16(c.pfx + (A >> 2) + 2u);
    return p;
}

// description of the swap of the node with record word y (parent | down << 16) with the leader at
// offset l (record word yl, pfx pfl)
HC_DEV u32 fgk_swap_info(u32 y, u32 yl, u32 l, u32 pfl)
{
    u32 info = pfl;
    info |= ((y ^ l) & 0xffffu) == 0u ? FGK_I_NOSWAP : 0u;               // leader == parent
    info |= ((y ^ yl) & 0xffffu) == 0u ? FGK_I_SAMEP : 0u;               // same parent
    info |= ((~(y & yl)) >> 13) & FGK_I_SC;                              // not both leaves (bit 31 of the words)
    return info;
}
static_assert(FGK_I_SC == (0x80000000u >> 13), "bit trick above");

// The part of the update that is only entered when some level of the path ties with the node after it
// (src/huffman.cpp:115-125: leader search, swap, continue from the leader's parent).  Same arguments
// as fgk_update; `pre` holds this round's loads.
// Returns true if the path table may have changed (an internal node moved, or the walk left the table).
// EARLY: a later round first asks with one VOTE whether any level ties at all.  Used by the encoder (-4.5 % warp
// instructions on high-entropy streams); in the decoder the same test made the symbol loop of every class four
// instructions longer (measured), so it keeps the plain form.
template <bool EARLY>
HC_DEV bool fgk_update_ties(FgkCtx &c, u32 A, u32 pf, u32 lane, FgkPre pre)
{
    const u32 lanebits = (lane << 27) | FGK_I_TIE;
    u32 guard = 0;
    bool changed = false;
    for (;;) {
        if (++guard > 600u) { c.err = 1; return true; }  // more rounds than nodes: never on a consistent tree
        const u32 depth = pf >> 12;
        const u32 W = pre.n.x;
        const bool tie = lane < depth && pre.n1.x == W;
        if (EARLY && guard > 1u && !any(tie)) {
            // a later round (the walk moved to another branch) without a tie -- the usual case: plain increments,
            // without describing swaps that nobody does
            syncwarp();                                   // this round's loads precede its stores
            sts32_if(lane < depth, c.rec + A, W + 1u);
            syncwarp();
            return changed;
        }
        const bool lng = pre.w2 == W;
        const u32 info = tie ? (lanebits | (lng ? (FGK_I_LONG | A) : fgk_swap_info(pre.n.y, pre.n1.y, A + 8u, pre.pfl))) : 0u;
        u32 hi = depth;                                   // lanes [0, hi) still have to add 1 to their node
        bool newpath = false;
        for (;;) {
            u32 r = reduce_max(lane < hi ? info : 0u);
            syncwarp();                                   // this round's loads precede its stores
            if (r == 0u) break;
            const u32 k0 = r >> 27;
            sts32_if(lane > k0 && lane < hi, c.rec + A, W + 1u);     // plain levels below the tie
            hi = k0;
            const bool me = lane == k0;
            // the swap as the tied lane sees it; replaced by uniform values when the run is long
            u32 a = A, l = A + 8u, y = pre.n.y, yl = pre.n1.y, w = W;
            if (r & FGK_I_LONG) {
                // a run of three or more equal weights: every lane probes one node of the run, then all
                // lanes fetch the two nodes of the swap
                a = r & 0xffffu;
                const uint2 na = lds64(c.rec + a);
                u32 pa = a + 24u + 8u * lane;
                pa = pa < FGK_SENT_O ? pa : FGK_SENT_O;
                const u32 m = ~ballot(lds32(c.rec + pa) == na.x);
                l = m ? a + 8u + 8u * (u32)ffs(m) : fgk_leader(c, a + 24u + 256u, na.x, lane);
                y = na.y; w = na.x;
                yl = lds32(c.rec + l + 4u);
                r = fgk_swap_info(y, yl, l, lds16(c.pfx + (l >> 2)));
                syncwarp();                               // these loads precede the tied lane's stores
            }
            const u32 pfl = r & 0xffffu;
            if (me) {
                if (r & FGK_I_NOSWAP) {
                    sts32(c.rec + a, w + 1u);
                } else {
                    // exchange the contents of a and l (src/huffman.cpp:186-217); the node continues at l.
                    // Weights here are >= 1, so neither node is the NYT.
                    const u32 sh = FGK_D - 1u - k0;
                    const u32 pfa = ((k0 + 1u) << 12) | (((pf & 0x1ffu) >> sh) << sh);
                    sts16(c.rec + a + 6u, fgk_down(yl));
                    sts16(c.rec + l + 6u, fgk_down(y));
                    if (r & FGK_I_SC) {
                        fgk_relink(c, fgk_down(yl), a, pfa);
                        fgk_relink(c, fgk_down(y), l, pfl);
                    } else {
                        sts32(c.spf + 4u * (fgk_down(yl) & 0xffu), a | (pfa << 16));
                        sts32(c.spf + 4u * (fgk_down(y) & 0xffu), l | (pfl << 16));
                    }
                    sts32(c.rec + l, w + 1u);
                }
            }
            // the common outcome in one test: two leaves swapped, different parents, the leader is listed in the
            // table and does not hang below the root (pfx 0x2000 .. 0x91ff)
            if ((r & (FGK_I_NOSWAP | FGK_I_SAMEP | FGK_I_SC)) == 0u && pfl - 0x2000u < 0x7200u) {
                pf = pfl - 0x1000u;
                A = fgk_lookup(c, pf);
                newpath = true;
                break;
            }
            if (r & FGK_I_NOSWAP) continue;
            if (r & FGK_I_SC) {
                // the old node is level k0 of this path, the leader's path came with r
                const u32 sh = FGK_D - 1u - k0;
                c.lev = fgk_rebuild_pair_cold(c, ((k0 + 1u) << 12) | (((pf & 0x1ffu) >> sh) << sh), pfl, lane);
                changed = true;
            } else if (r & FGK_I_SAMEP) {
                continue;                                 // same parent: the levels above are as loaded
            }
            if (pfl == FGK_NOPATH) {
                // the leader lies deeper than the table: finish sequentially from its parent
                const u32 pl = shfl(fgk_parent(yl), (int)k0);
                c = fgk_update_seq(c, pl, lane);
                return true;
            }
            if ((pfl >> 12) == 1u) { hi = 0; break; }     // the leader hangs below the root: done
            pf = pfl - 0x1000u;                           // the parent of the leader: one level up, same path bits
            A = fgk_lookup(c, pf);
            newpath = true;
            break;
        }
        if (!newpath) {
            sts32_if(lane < hi, c.rec + A, W + 1u);
            syncwarp();
            return changed;
        }
        pre = fgk_preload(c, A);
    }
}

// FGK update of a node whose path is in the table (src/huffman.cpp:113-127).  A = this lane's node
// (fgk_lookup of pf), pf = depth << 12 | left-aligned path of the node the update starts from, `pre` =
// fgk_preload(A).  Most symbols take the first exit: no level ties, every lane adds 1 to its node.
// Returns 0 on that exit (nothing but weights changed), 1 if leaves may have moved, 3 if the path table may
// have changed as well.
template <bool EARLY>
HC_DEV u32 fgk_update(FgkCtx &c, u32 A, u32 pf, u32 lane, const FgkPre &pre)
{
    const bool valid = lane < (pf >> 12);
    const bool tie = valid && pre.n1.x == pre.n.x;
    const bool some = any(tie);
    syncwarp();                                           // the loads precede the stores
    if (!some) {
        sts32_if(valid, c.rec + A, pre.n.x + 1u);
        syncwarp();
        return 0u;
    }
    return fgk_update_ties<EARLY>(c, A, pf, lane, pre) ? 3u : 1u;
}

#if defined(HC_EMU_DEBUG) || defined(HC_FGK_CHECK)
// debugging aid (emulator debug builds, -DHC_FGK_CHECK device builds): consistency of the whole tree
// after symbol idx; returns false and prints what is wrong
HC_DEV bool fgk_validate(const FgkCtx &c, u32 lane, u32 idx)
{
    if (lane != 0) return true;
    u32 prevw = 0;
    bool bad = false;
    for (u32 o = c.nyt; o < FGK_ROOT_O && !bad; o += 8) {
        const uint2 r = lds64(c.rec + o);
        const u32 k = fgk_down(r.y), par = fgk_parent(r.y);
        const u32 pf = lds16(c.pfx + (o >> 2));
        if (r.x < prevw) { printf("sym %u: weight order broken at %u\n", idx, o); bad = true; }
        prevw = r.x;
        if (par <= o || par > FGK_ROOT_O) { printf("sym %u: parent of %u = %u\n", idx, o, par); bad = true; }
        else {
            const u32 pk = fgk_down(lds32(c.rec + par + 4u));
            if ((pk & FGK_LEAF) || (pk != (o & ~8u))) { printf("sym %u: parent %u of %u has down %x\n", idx, par, o, pk); bad = true; }
        }
        if (k & FGK_LEAF) {
            if (k != FGK_LEAF_NYT) {
                const u32 sp = lds32(c.spf + 4u * (k & 0xffu));
                if (sp != (o | (pf << 16))) { printf("sym %u: spf[%u] = %x, leaf at %u pfx %x\n", idx, k & 0xff, sp, o, pf); bad = true; }
            } else if (o != c.nyt) { printf("sym %u: NYT at %u, c.nyt %u\n", idx, o, c.nyt); bad = true; }
        }
        if (pf != FGK_NOPATH) {
            const u32 d = pf >> 12, path = (pf & 0x1ffu) >> (FGK_D - d);
            const u32 e = lds16(c.pt + 2u * ((1u << d) - 2u + path));
            if (e != o) { printf("sym %u: pfx[%u] = %x but pt there = %u\n", idx, o, pf, e); bad = true; }
        }
    }
    if (bad)
        for (u32 q = c.nyt; q <= FGK_ROOT_O; q += 8) {
            const uint2 z = lds64(c.rec + q);
            printf("   node %u (off %u): w=%u parent=%u down=%x pfx=%x\n", q / 8, q, z.x, fgk_parent(z.y), fgk_down(z.y), lds16(c.pfx + (q >> 2)));
        }
    return !bad;
}
#endif

// The trees of the resident CTAs do not hold the whole batch at once, so streams are started in
// order of decreasing length class (counting sort, 4 classes per octave): the long ones first, the
// short ones fill the slots they free.  One CTA; the order only affects scheduling, never the output.
HC_DEV u32 fgk_len_class(u64 v)
{
    if (!v) return 0;
    const u32 lg = 63u - (u32)clzll(v);
    const u32 cl = 4u * lg + (u32)((lg >= 2u ? v >> (lg - 2u) : v << (2u - lg)) & 3u);
    return cl > 255u ? 255u : cl;
}

HC_KERNEL HC_LAUNCH_BOUNDS(1024, 1)
fgk_order_kernel(const u64 *HC_RESTRICT len, u32 nf, u32 *HC_RESTRICT order)
{
    // stable counting sort (file order inside a class, so the schedule is the same in every run): the
    // files are cut in four quarters, thread (class, quarter) places the files of its class in its quarter
    HC_SHARED u32 hist[4][256];
    HC_SHARED u32 ctot[256];
    const u32 tid = threadIdx.x, cls = tid & 255u, qtr = tid >> 8;
    const u32 per = (nf + 3u) / 4u;
    hist[qtr][cls] = 0;
    syncthreads();
    for (u32 f = tid; f < nf; f += blockDim.x) atomic_add(&hist[f / per][fgk_len_class(len[f])], 1u);
    syncthreads();
    if (tid < 256u) ctot[tid] = hist[0][tid] + hist[1][tid] + hist[2][tid] + hist[3][tid];
    syncthreads();
    if (tid == 0) {
        u32 acc = 0;
        for (int b = 255; b >= 0; b--) { const u32 h = ctot[b]; ctot[b] = acc; acc += h; }   // longest class first
    }
    syncthreads();
    u32 pos = ctot[cls];
    for (u32 q = 0; q < qtr; q++) pos += hist[q][cls];
    const u32 f1 = (qtr + 1u) * per < nf ? (qtr + 1u) * per : nf;
    for (u32 f = qtr * per; f < f1; f++)
        if (fgk_len_class(len[f]) == cls) order[pos++] = f;
}

// MSB-first bit writer over a 64-word shared-memory ring (c.ring): bits are OR-ed into zeroed words,
// a half (32 words = 128 bytes) is written out coalesced and zeroed again as soon as the bit position
// has left it.  Serial puts (lane 0) for long codes and the header, warp-parallel packing for the rest.
struct BitWriter {
    u32 pos;       // bit position inside the ring (0..2047)
    u32 groups;    // 128-byte groups written to `dst`
    u32 *dst;
    u64 cap_words;
    bool overflow;
};

HC_DEV void bw_init(BitWriter &b, u8 *dst, u64 cap_bytes)
{
    b.pos = 0; b.groups = 0;
    b.dst = (u32 *)dst;
    b.cap_words = cap_bytes / 4;
    b.overflow = false;
}

// the position moved from `old` to b.pos: write out the half that was left behind
HC_DEV void bw_advance(const FgkCtx &c, BitWriter &b, u32 old, u32 lane)
{
    if ((old ^ b.pos) & 1024u) {
        syncwarp();                                       // all ORs into that half are done
        const u32 half = c.ring + ((old >> 3) & 128u);
        const u32 word = lds32(half + 4u * lane);
        sts32(half + 4u * lane, 0u);
        if (((u64)b.groups + 1u) * 32u <= b.cap_words) stg32_stream(b.dst + (u64)b.groups * 32u + lane, bswap32(word));
        else b.overflow = true;
        b.groups++;
        syncwarp();
    }
}

HC_DEV void bw_put(const FgkCtx &c, BitWriter &b, u32 v, u32 d, u32 lane)   // serial: 0 <= d <= 32, v < 2^d
{
    if (d == 0u) return;
    const u32 old = b.pos, off = old & 31u;
    const u64 bits = ((u64)v << (64u - d)) >> off;        // left aligned at bit `off` of a two-word window
    syncwarp();
    if (lane == 0) {
        const u32 w0 = c.ring + ((old >> 3) & 252u), w1 = c.ring + (((old >> 3) + 4u) & 252u);
        sts32(w0, lds32(w0) | (u32)(bits >> 32));
        if ((u32)bits) sts32(w1, lds32(w1) | (u32)bits);
    }
    syncwarp();
    b.pos = (old + d) & 2047u;
    bw_advance(c, b, old, lane);
}

// emit the code collected in hi:lo (marker scheme of fgk_code_bit); returns false if the code is
// longer than 56 bits (impossible below 2^32 symbols)
HC_DEV bool bw_put_code(const FgkCtx &c, BitWriter &b, u32 hi, u32 lo, u32 lane)
{
    if (lo == 0u) {                                   // depth <= 31: everything is in hi
        const u32 d = 32u - (u32)ffs(hi);             // marker = lowest set bit of hi
        bw_put(c, b, d ? hi >> (32u - d) : 0u, d, lane);
        return true;
    }
    const u32 d2 = 32u - (u32)ffs(lo);                // bits of the code that live in lo
    bw_put(c, b, hi, 32, lane);
    bw_put(c, b, d2 ? lo >> (32u - d2) : 0u, d2, lane);
    return d2 <= 24u;
}

// warp-parallel: lane j contributes the table code parked in `code` (pfx format, 0 = nothing), in lane order
HC_DEV void bw_pack(const FgkCtx &c, BitWriter &b, u32 code, u32 lane)
{
    const u32 d = code >> 12;
    u32 inc = d;                                          // inclusive scan of the lengths (<= 9 * 32)
#pragma unroll
    for (u32 s = 1; s < 32u; s <<= 1) {
        const u32 x = shfl_up(inc, s);
        if (lane >= s) inc += x;
    }
    const u32 total = shfl(inc, 31);
    if (total == 0u) return;
    const u32 old = b.pos, p = old + inc - d, off = p & 31u;
    if (d) {
        const u64 bits = ((u64)((code & 0x1ffu) >> (FGK_D - d)) << (64u - d)) >> off;
        atomic_or_smem(c.ring + ((p >> 3) & 252u), (u32)(bits >> 32));
        if ((u32)bits) atomic_or_smem(c.ring + (((p >> 3) + 4u) & 252u), (u32)bits);
    }
    b.pos = (old + total) & 2047u;
    bw_advance(c, b, old, lane);
}

// flush the tail; returns the total number of bytes of the stream
HC_DEV u64 bw_finish(const FgkCtx &c, BitWriter &b, u32 lane)
{
    syncwarp();
    const u32 in_half = b.pos & 1023u;                    // bits pending in the current half
    const u32 nbytes = (in_half + 7u) / 8u;
    const u64 total = (u64)b.groups * 128u + nbytes;
    if (total > b.cap_words * 4u) { b.overflow = true; return total; }
    const u32 half = c.ring + ((b.pos >> 3) & 128u);
    const u32 word = bswap32(lds32(half + 4u * lane));    // big-endian word -> bytes in stream order
    u8 *p = (u8 *)(b.dst + (u64)b.groups * 32u + lane);
    if (4u * lane + 4u <= nbytes) *(u32 *)p = word;
    else for (u32 i = 0; 4u * lane + i < nbytes; i++) p[i] = (u8)(word >> (8u * i));
    return total;
}

HC_KERNEL HC_LAUNCH_BOUNDS(FGK_ENC_WARPS * 32, 1)
fgk_encode_kernel(const u8 *HC_RESTRICT sym, const u64 *HC_RESTRICT sym_off, const u64 *HC_RESTRICT sym_len,
                  const u8 *HC_RESTRICT flags, u8 *HC_RESTRICT out, const u64 *HC_RESTRICT out_off,
                  const u64 *HC_RESTRICT out_cap, u64 *HC_RESTRICT out_len, i32 *HC_RESTRICT status, u32 nf,
                  const u32 *HC_RESTRICT order)
{
    HC_SHARED FgkTree trees[FGK_ENC_WARPS];
    HC_SMEM_ARENA(trees);
    const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const u32 fi = blockIdx.x * FGK_ENC_WARPS + wid;
    if (fi >= nf) return;
    const u32 f = order ? order[fi] : fi;               // longest streams first (fgk_order_kernel)
    FgkCtx c;
    fgk_init(c, trees[wid], lane);

    const u64 m64 = sym_len[f];
    const bool too_many = m64 > 0xfffffff0ull;            // weights are 32-bit
    const u32 m = too_many ? 0u : (u32)m64;
    const u32 *src = (const u32 *)(sym + sym_off[f]);     // 256-byte aligned region
    BitWriter bw;
    bw_init(bw, out + out_off[f], out_cap[f] & ~(u64)3);
    // container header <u64 LE count><u8 flags> (src/headers.cpp:107-125) through the bit writer
    bw_put(c, bw, bswap32((u32)m64), 32, lane);
    bw_put(c, bw, bswap32((u32)(m64 >> 32)), 32, lane);
    bw_put(c, bw, flags ? flags[f] : 0u, 8, lane);

    bool too_long = too_many;
    // symbols are staged 128 at a time (4 per lane, the next chunk already on its way from global memory);
    // table codes are parked in lane (symbol index mod 32) and packed by the warp every 32 symbols
    u32 cnext = m > 0 ? ldg32(src + lane) : 0u;
    u32 code = 0;                                         // this lane's parked table code
    for (u32 i0 = 0; i0 < m; i0 += 128) {
        syncwarp();                                       // every lane is done with the previous chunk
        sts32(c.buf + 4u * lane, cnext);
        syncwarp();
        if (i0 + 128u < m) cnext = ldg32(src + (i0 + 128u) / 4u + lane);
        const u32 cnt = (m - i0) < 128u ? (m - i0) : 128u;
        // Lookahead, valid by construction: the entry (spf) and the path nodes (A) of the NEXT symbol are fetched while
        // this symbol is processed.  An update that took the first exit of fgk_update changed nothing but weights, so
        // they are still right; after any other update they are simply fetched again.
        u32 y = lds8(c.buf), yn = lds8(c.buf + 1u);
        u32 sp = lds32(c.spf + 4u * y);
        u32 A = fgk_lookup(c, sp >> 16);
        for (u32 g0 = 0; g0 < cnt; g0 += 32) {
            const u32 gcnt = (cnt - g0) < 32u ? (cnt - g0) : 32u;
            for (u32 j = 0; j < gcnt; j++) {
                const u32 ynn = lds8(c.buf + ((g0 + j + 2u) & 255u));     // (past the chunk: never used, see the re-fetch below)
                u32 spn = lds32(c.spf + 4u * yn);
                u32 An = fgk_lookup(c, spn >> 16);
                const u32 pf = sp >> 16;
                u32 dirty;
                if (pf == FGK_NOPATH) {
                    // Not yet transmitted, or deeper than the table.  Encode precedes update (src/transform.cpp:372-375):
                    // the code is read off the tree first, by a parent chase; a new symbol sends the NYT code + 8 raw
                    // bits (src/huffman.cpp:42-51), then splits.
                    bw_pack(c, bw, code, lane);
                    code = 0;
                    const bool isnew = sp == FGK_SPF_NONE;
                    u32 hi = 0x80000000u, lo = 0u, start = sp & 0xffffu;
                    fgk_code_of(c, isnew ? c.nyt : start, hi, lo);
                    if (!bw_put_code(c, bw, hi, lo, lane)) too_long = true;
                    if (isnew) {
                        bw_put(c, bw, y, 8, lane);
                        const u32 n = c.nyt, pn = n == FGK_ROOT_O ? 0u : lds16(c.pfx + (n >> 2));
                        start = fgk_split(c, y, lane);
                        fgk_table_split(c, n, pn, lane);
                    }
                    c = fgk_update_seq(c, start, lane);
                    dirty = 1;
                } else {
                    // the code of a leaf is its path (encode precedes update, src/transform.cpp:372-375)
                    if (lane == j) code = pf;
                    dirty = fgk_update<true>(c, A, pf, lane, fgk_preload(c, A));
                }
                if (dirty) {
                    spn = lds32(c.spf + 4u * yn);
                    An = fgk_lookup(c, spn >> 16);
                }
#if defined(HC_EMU_DEBUG) || defined(HC_FGK_CHECK)
                syncwarp();
                if (!c.err && ballot(!fgk_validate(c, lane, i0 + g0 + j))) c.err = 2;
                syncwarp();
#endif
                y = yn; yn = ynn;
                sp = spn;
                A = An;
            }
            bw_pack(c, bw, code, lane);
            code = 0;
            if (c.err) { i0 = m; break; }
        }
    }
    bw_pack(c, bw, code, lane);
    u64 total = bw_finish(c, bw, lane);
    if (lane == 0) {
        out_len[f] = total;
        status[f] = c.err ? 102 : (too_long ? 101 : (bw.overflow ? 100 : 0));
    }
}

// MSB-first bit reader over 128-byte chunks held one word per lane.  Reading past the end yields
// zeros; the caller compares br_consumed with the length at the end (a truncated stream decodes
// garbage that is then discarded with status 9).
struct BitReader {
    u64 win;        // next bits, MSB aligned
    u32 wbits;      // valid bits in win (> 32 between calls)
    u32 ridx;       // next word index to pull into the window
    u32 chunk, chunk_next;
    const u32 *src;
    u32 nwords;     // words that may be loaded
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
        const u32 nx = r.ridx + 32u + lane;
        r.chunk_next = nx < r.nwords ? ldg32(r.src + nx) : 0u;
    }
}

HC_DEV void br_init(BitReader &r, const u8 *p, u64 len_bytes, u32 lane)
{
    r.src = (const u32 *)p;
    r.nwords = (u32)((len_bytes + 3u) / 4u);      // reads stay inside the padded region
    r.chunk = lane < r.nwords ? ldg32(r.src + lane) : 0u;
    r.chunk_next = 32u + lane < r.nwords ? ldg32(r.src + 32u + lane) : 0u;
    r.win = 0; r.wbits = 0; r.ridx = 0;
    br_refill(r, lane);
    br_refill(r, lane);
}

// drop d (1..32) bits that the caller has already looked at
HC_DEV void br_skip(BitReader &r, u32 d, u32 lane)
{
    r.win <<= d;
    r.wbits -= d;
    if (r.wbits <= 32u) br_refill(r, lane);
}

HC_DEV u32 br_get(BitReader &r, u32 d, u32 lane)   // 1..32 bits
{
    u32 v = (u32)(r.win >> (64u - d));
    br_skip(r, d, lane);
    return v;
}

HC_DEV u64 br_consumed(const BitReader &r) { return (u64)r.ridx * 32u - r.wbits; }

// cold half of the decoder (src/huffman.cpp:60-93): the root is still a leaf, the code is longer
// than the path table, or the symbol arrives raw behind the NYT code.  Walks down bit by bit (codes
// beyond 32 bits included), then updates sequentially.
HC_DEV u32 fgk_decode_cold(FgkCtx &c, BitReader &br, u32 lane)
{
    u32 d = FGK_ROOT_O, k = fgk_down(lds32(c.rec + d + 4u)), len = 0;
    u32 t = (u32)(br.win >> 32), guard = 0;
    while (!(k & FGK_LEAF)) {
        if (++guard > 600u || k >= FGK_ROOT_O) { c.err = 1; return 0; }
        if (len == 32u) {
            br_skip(br, 32, lane);
            t = (u32)(br.win >> 32);
            len = 0;
        }
        d = k + ((t >> 28) & 8u);
        t <<= 1;
        len++;
        k = fgk_down(lds32(c.rec + d + 4u));
    }
    if (len) br_skip(br, len, lane);
    if (k != FGK_LEAF_NYT) {
        c = fgk_update_seq(c, d, lane);
        return k & 0xffu;
    }
    const u32 y = br_get(br, 8, lane);
    // a raw symbol that is already in the tree is still decoded as that symbol
    // (src/huffman.cpp:74-86); the update then starts from its existing leaf
    const u32 sp = lds32(c.spf + 4u * y);
    u32 start = sp & 0xffffu;
    if (sp == FGK_SPF_NONE) {
        const u32 n = c.nyt, pn = n == FGK_ROOT_O ? 0u : lds16(c.pfx + (n >> 2));
        start = fgk_split(c, y, lane);
        fgk_table_split(c, n, pn, lane);
    }
    c = fgk_update_seq(c, start, lane);
    return y;
}

HC_KERNEL HC_LAUNCH_BOUNDS(FGK_DEC_WARPS * 32, 1)
fgk_decode_kernel(const u8 *HC_RESTRICT in, const u64 *HC_RESTRICT in_off, const u64 *HC_RESTRICT in_len,
                  u8 *HC_RESTRICT sym, const u64 *HC_RESTRICT sym_off, const u64 *HC_RESTRICT sym_cap,
                  u64 *HC_RESTRICT sym_len, u8 *HC_RESTRICT flags, i32 *HC_RESTRICT status, u32 nf,
                  const u32 *HC_RESTRICT order)
{
    HC_SHARED FgkTree trees[FGK_DEC_WARPS];
    HC_SMEM_ARENA(trees);
    const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const u32 fi = blockIdx.x * FGK_DEC_WARPS + wid;
    if (fi >= nf) return;
    const u32 f = order ? order[fi] : fi;               // longest streams first (fgk_order_kernel)
    FgkCtx c;
    fgk_init(c, trees[wid], lane);

    const u64 n = in_len[f];
    if (n < 9) {                                         // src/main.cpp:99-104
        if (lane == 0) { sym_len[f] = 0; if (flags) flags[f] = 0; status[f] = 8; }
        return;
    }
    BitReader br;
    br_init(br, in + in_off[f], n, lane);
    u32 lo = bswap32(br_get(br, 32, lane));
    u32 hi = bswap32(br_get(br, 32, lane));
    const u64 m64 = ((u64)hi << 32) | lo;
    u32 fl = br_get(br, 8, lane);
    if (lane == 0 && flags) flags[f] = (u8)fl;
    const u64 cap = sym_cap[f], avail = n * 8u - 72u;
    if (m64 > cap || m64 > avail + 1u || m64 > 0xfffffff0ull) {
        // every symbol after the first costs >= 1 bit: a count beyond avail+1 is a guaranteed
        // underrun, for which the reference exits with 9 (src/transform.cpp:394-398)
        if (lane == 0) { sym_len[f] = m64; status[f] = (m64 > avail + 1u) ? 9 : 100; }
        return;
    }
    const u32 m = (u32)m64;
    u32 *dst = (u32 *)(sym + sym_off[f]);
    // lanes 0..8 look up levels 1..9; the others read one of the two pad entries behind the table, which always say
    // "no node" (the root, which is internal as soon as the first symbol has arrived)
    const u32 shd = lane < FGK_D ? 31u - lane : 31u;
    const u32 ptd = lane < FGK_D ? c.ptj : c.pt + 2u * FGK_PT_N;
    const u32 lanekey = lane << 16;
    u32 e = lds16(ptd + 2u * ((u32)(br.win >> 32) >> shd));
    for (u32 i0 = 0; i0 < m; i0 += 128) {
        const u32 cnt = (m - i0) < 128u ? (m - i0) : 128u;
        for (u32 i = 0; i < cnt; i++) {
            // The path table resolves the next FGK_D code bits in one step: lane j looks up the node that the first
            // j + 1 bits lead to; the shallowest leaf among them ends the code (REDUX.MIN over lane << 16 | symbol).
            // The loads of the update's first round ride on the same round of loads.  A lane without a node holds
            // the root: a leaf only before the first symbol, which the cold path handles.
            const u32 t = (u32)(br.win >> 32);
            const FgkPre pre = fgk_preload(c, e);
            const u32 r = reduce_min((i32)pre.n.y < 0 ? (lanekey | ((pre.n.y >> 16) & 0x1ffu)) : 0xffffffffu);
            u32 y, dirty;
            if (r == 0xffffffffu || (r & 0x100u)) {
                y = fgk_decode_cold(c, br, lane);
                if (c.err) { i0 = m; break; }
                dirty = 3;
            } else {
                const u32 depth = (r >> 16) + 1u;
                y = r & 0xffu;
                br_skip(br, depth, lane);
                // the next code starts here: its table lookup only depends on this symbol's update if that changes
                // the shape of the tree
                const u32 en = lds16(ptd + 2u * ((u32)(br.win >> 32) >> shd));
                dirty = fgk_update<false>(c, e, (depth << 12) | (t >> 23), lane, pre);
                e = en;
            }
            if (dirty & 2u) e = lds16(ptd + 2u * ((u32)(br.win >> 32) >> shd));
            if (lane == 0) sts8(c.buf + i, y);
        }
        syncwarp();
        if (cnt == 128u) stg32_stream(dst + (i0 >> 2) + lane, lds32(c.buf + 4u * lane));
        else if (lane < (cnt + 3u) / 4u) dst[(i0 >> 2) + lane] = lds32(c.buf + 4u * lane);    // the last, partial group
        syncwarp();
    }
    const bool underrun = br_consumed(br) > n * 8u;
    if (lane == 0) { sym_len[f] = m64; status[f] = c.err ? 102 : (underrun ? 9 : 0); }
}

}  // namespace hcd
