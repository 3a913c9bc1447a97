// fgk.cuh -- batched adaptive Huffman (FGK) encode / decode.
// Reference: HuffTree::{encode,decode,update} src/huffman.cpp:37-128, findSuccNode :157-184,
// swapNodes :186-217, drivers applyHuffman/revertHuffman src/transform.cpp:363-406, container
// header src/headers.cpp:107-125 and bit packing src/main.cpp:78-84.
//
// The tree update is serial inside one stream, so the parallelism is the batch: ONE WARP PER
// FILE with the tree in shared memory (9.9 KB per stream, 20 streams per SM).  All 32 lanes execute
// the same control flow; lane 0 is the only writer of links.  The kernel is bound by the chain of
// dependent instructions of one warp (a lone high-entropy stream is only 15-20 % faster than the
// whole batch), so everything is arranged to shorten that chain: a PATH TABLE (below, at FgkTree)
// lets the 32 lanes look at all levels of a root->leaf path at once, and the layout makes the
// per-level work of the remaining sequential walks as few instructions as possible:
//
//   slot = node number 0..512: siblings adjacent, even slot = left child = bit 0,
//   weights non-decreasing in slot order, so the block leader of s is the last slot l >= s with
//   w[l] == w[s].
//   up[s]   = { weight, SHARED-MEMORY ADDRESS of up[parent(s)] }   one LDS.64, no address math
//   down[s] = internal: shared address of down[left child]; leaf: (symbol << 1) | 1
//   The code bit of slot s is bit 3 of the address of up[s] (8-byte entries, 16-byte aligned base).
//
//   sequential walk, per level: LDS.64 up[s], LDS.32 up[s+1].w -> if the weights differ (common
//   case) there is no leader to find: lane 0 stores w+1 and the walk follows the parent address.
//   Only when they are equal the warp probes 32 slots per ballot for the leader and, if needed,
//   swaps the two subtrees.  Used for nodes deeper than the path table; there encoding collects the
//   code bits on the same walk (shifted in from the top, a marker bit tracks the length) until
//   the first swap, after which the old path is finished alongside the rest of the update.
//
// Bits are packed MSB-first into 32-bit words; each lane keeps one word and the warp flushes
// 128 bytes at a time (coalesced).  The 9-byte container header goes through the same writer.
#pragma once
#include "hc_common.cuh"

namespace hcd {

#ifndef HC_FGK_ENC_WARPS
#define HC_FGK_ENC_WARPS 1
#endif
#ifndef HC_FGK_DEC_WARPS
#define HC_FGK_DEC_WARPS 4
#endif
// streams per CTA.  Measured on C3 (final kernels): encoder 1 / 2 / 3 / 4 -> 152 / 178-186 / 183 / 174 ms, decoder
// 1 / 2 / 4 -> 205 / 201 / 198 ms.  One-warp CTAs give their 9.9 KB back as soon as their stream ends.
constexpr int FGK_ENC_WARPS = HC_FGK_ENC_WARPS;
constexpr int FGK_DEC_WARPS = HC_FGK_DEC_WARPS;
constexpr int FGK_WARPS = FGK_ENC_WARPS > FGK_DEC_WARPS ? FGK_ENC_WARPS : FGK_DEC_WARPS;
constexpr u32 FGK_ROOT = 512;
constexpr u32 FGK_NSLOT = 514;           // slots 0..512 + one sentinel (weight 0xffffffff)
constexpr u32 FGK_LEAF_NYT = (256u << 1) | 1u;
constexpr u32 FGK_D = 9;                         // levels covered by the path table
constexpr u32 FGK_PT_N = (2u << FGK_D) - 2u;     // 2 + 4 + ... + 2^D entries
constexpr u32 FGK_NOPATH = 0xffffu;

// PATH TABLE.  pt[(1 << d) - 2 + p] = slot + 1 of the node reached from the root by the d-bit path p
// (0: no such node), for d = 1..FGK_D; pfx[slot] = d << 12 | p for those nodes (FGK_NOPATH for
// deeper ones).  With it the nodes of a root->leaf path are found by 32 lanes AT ONCE (one lookup
// per level) instead of by a chain of dependent loads: decoding reads the next FGK_D code bits and
// knows every node on the way; encoding takes the code of a leaf straight from pfx; and the
// update's "does this level need a swap" test runs for all levels of the path in one ballot.
// The table describes SLOTS, so the most frequent swap (two leaves exchange their symbols) leaves it
// valid; a swap that moves an internal node re-derives the entries below the two slots (warp-parallel,
// level by level), an NYT split adds its two entries.  Deeper levels fall back to the sequential walks.
struct HC_ALIGNED16 FgkTree {
    uint2 up[FGK_NSLOT];     // {weight, shared address of the parent's up entry}
    u32 down[FGK_NSLOT];     // see above
    u16 slot_of[256];        // leaf slot of a symbol, 0xffff = not yet transmitted
    u16 pt[FGK_PT_N + 2];
    u16 pfx[FGK_NSLOT + 2];
    u8 buf[128];             // staging of 128 symbols (one coalesced transfer)
    u8 pad[8];
};

static_assert(sizeof(FgkTree) * FGK_WARPS <= 48u * 1024u, "the trees of a CTA are static shared memory");

struct FgkCtx {              // shared addresses, identical in every lane
    u32 up, down, slot_of, buf, root, sentinel, nyt;
    u32 pt, pfx, lev;        // lev: number of populated levels of pt
    u32 fl;                  // out-of-line helpers: bit 0 = the watched leaf moved
};

HC_DEV void fgk_init(FgkCtx &c, FgkTree &t, u32 lane)
{
    c.up = smem_addr(&t.up[0]);
    c.down = smem_addr(&t.down[0]);
    c.slot_of = smem_addr(&t.slot_of[0]);
    c.buf = smem_addr(&t.buf[0]);
    c.root = c.up + 8u * FGK_ROOT;
    c.sentinel = c.up + 8u * (FGK_ROOT + 1u);
    c.nyt = c.root;
    c.pt = smem_addr(&t.pt[0]);
    c.pfx = smem_addr(&t.pfx[0]);
    c.lev = 0;
    c.fl = 0;
    for (u32 i = lane; i < 128u; i += 32) sts32(c.slot_of + 4u * i, 0xffffffffu);
    for (u32 i = lane; i < (FGK_PT_N + 2u) / 2u; i += 32) sts32(c.pt + 4u * i, 0u);
    for (u32 i = lane; i < (FGK_NSLOT + 2u) / 2u; i += 32) sts32(c.pfx + 4u * i, 0xffffffffu);
    if (lane == 0) {
        uint2 z; z.x = 0; z.y = 0;
        sts64(c.root, z);
        z.x = 0xffffffffu;
        sts64(c.sentinel, z);
        sts32(c.down + 4u * FGK_ROOT, FGK_LEAF_NYT);
    }
    syncwarp();
}

HC_DEV u32 fgk_down_of(const FgkCtx &c, u32 a) { return c.down + ((a - c.up) >> 1); }   // up address -> down address
HC_DEV u32 fgk_up_of(const FgkCtx &c, u32 d) { return c.up + ((d - c.down) << 1); }

// Rebuilds the path table from the tree, one level per step, the nodes of a level spread over the
// lanes.  Called by all lanes after the tree changed shape.
HC_DEV void fgk_rebuild(FgkCtx &c, u32 lane)
{
    syncwarp();                                           // the tree writes of lane 0 are visible
    for (u32 sl = ((c.nyt - c.up) >> 3) + lane; sl <= FGK_ROOT; sl += 32) sts16(c.pfx + 2u * sl, FGK_NOPATH);
    syncwarp();
    u32 lev = 0;
    for (u32 d = 1; d <= FGK_D; d++) {
        const u32 base = (1u << d) - 2u, pbase = (1u << (d - 1u)) - 2u;
        u32 any = 0;
        for (u32 p = lane; p < (1u << d); p += 32) {
            const u32 pe = d == 1u ? FGK_ROOT + 1u : lds16(c.pt + 2u * (pbase + (p >> 1)));
            u32 e = 0;
            if (pe) {
                const u32 kd = lds32(c.down + 4u * (pe - 1u));
                if (!(kd & 1u)) e = ((kd - c.down) >> 2) + (p & 1u) + 1u;     // slot + 1 of the child
            }
            sts16(c.pt + 2u * (base + p), e);
            if (e) sts16(c.pfx + 2u * (e - 1u), (d << 12) | p);
            any |= e;
        }
        syncwarp();
        if (ballot(any != 0u) == 0u) {
            // level d is empty (and written as such); clear what an earlier, deeper tree left below
            for (u32 d2 = d + 1u; d2 <= c.lev; d2++)
                for (u32 p = lane; p < (1u << d2); p += 32) sts16(c.pt + 2u * ((1u << d2) - 2u + p), 0u);
            break;
        }
        lev = d;
    }
    c.lev = lev;
    syncwarp();
}

// After slots x and y exchanged their subtrees only the table entries BELOW them change (the slots
// keep their own paths).  px / py = their pfx entries (FGK_NOPATH: that slot lies deeper than the
// table, nothing below it is listed).  First every node listed below either slot loses its path,
// then both ranges are re-derived level by level.
HC_DEV void fgk_rebuild_pair(FgkCtx &c, u32 px, u32 py, u32 lane)
{
    syncwarp();                                           // the tree writes of lane 0 are visible
#pragma unroll 1
    for (u32 side = 0; side < 2u; side++) {
        const u32 pf = side ? py : px;
        if (pf == FGK_NOPATH) continue;
        const u32 depth = pf >> 12, path = pf & 0xfffu;
        for (u32 d = depth + 1u; d <= FGK_D; d++) {
            const u32 n = 1u << (d - depth), base = (1u << d) - 2u + (path << (d - depth));
            for (u32 i = lane; i < n; i += 32) {
                const u32 e = lds16(c.pt + 2u * (base + i));
                if (e) sts16(c.pfx + 2u * (e - 1u), FGK_NOPATH);
            }
        }
    }
    syncwarp();
    u32 lev = c.lev;
#pragma unroll 1
    for (u32 side = 0; side < 2u; side++) {
        const u32 pf = side ? py : px;
        if (pf == FGK_NOPATH) continue;
        const u32 depth = pf >> 12, path = pf & 0xfffu;
        for (u32 d = depth + 1u; d <= FGK_D; d++) {
            const u32 n = 1u << (d - depth), p0 = path << (d - depth);
            const u32 base = (1u << d) - 2u, pbase = (1u << (d - 1u)) - 2u;
            u32 any = 0;
            for (u32 i = lane; i < n; i += 32) {
                const u32 p = p0 + i;
                const u32 pe = lds16(c.pt + 2u * (pbase + (p >> 1)));
                u32 e = 0;
                if (pe) {
                    const u32 kd = lds32(c.down + 4u * (pe - 1u));
                    if (!(kd & 1u)) e = ((kd - c.down) >> 2) + (p & 1u) + 1u;
                }
                sts16(c.pt + 2u * (base + p), e);
                if (e) sts16(c.pfx + 2u * (e - 1u), (d << 12) | p);
                any |= e;
            }
            syncwarp();
            if (ballot(any != 0u) && d > lev) lev = d;
        }
    }
    c.lev = lev;
}

// table entries of the two slots created by an NYT split of slot n (up address); pn = pfx of n
// before the split, or depth 0 for the root
HC_DEV void fgk_table_split(FgkCtx &c, u32 n, u32 pn, u32 lane)
{
    if (pn == FGK_NOPATH) return;                         // deeper than the table: the new slots stay unlisted
    const u32 depth = pn >> 12, path = pn & 0xfffu;
    if (depth >= FGK_D) return;
    const u32 d = depth + 1u, sl = (n - c.up) >> 3;       // children: slots sl-2 (bit 0) and sl-1 (bit 1)
    if (lane < 2u) {
        const u32 p = (path << 1) | lane, child = sl - 2u + lane;
        sts16(c.pt + 2u * ((1u << d) - 2u + p), child + 1u);
        sts16(c.pfx + 2u * child, (d << 12) | p);
    }
    if (d > c.lev) c.lev = d;
    syncwarp();
}

// NYT split (src/huffman.cpp:99-111): the NYT slot n becomes internal with children n-2 (new NYT)
// and n-1 (leaf of `sym`).  Returns the up address of the new leaf.
HC_DEV u32 fgk_split(FgkCtx &c, u32 sym, u32 lane)
{
    const u32 n = c.nyt;                      // up address of the current NYT
    syncwarp();                               // every lane has done its slot_of / nyt-path reads
    if (lane == 0) {
        uint2 z; z.x = 0; z.y = n;
        sts64(n - 8u, z);                     // leaf  (slot n-1): weight 0, parent n
        sts64(n - 16u, z);                    // NYT   (slot n-2)
        const u32 dn = fgk_down_of(c, n);
        sts32(dn, dn - 8u);                   // internal: address of the left child's down entry
        sts32(dn - 4u, (sym << 1) | 1u);
        sts32(dn - 8u, FGK_LEAF_NYT);
        sts16(c.slot_of + 2u * sym, ((n - c.up) >> 3) - 1u);
    }
    c.nyt = n - 16u;
    syncwarp();
    return n - 8u;
}

// collect one code bit (bit 3 of the up address) at the top of the 64-bit accumulator hi:lo
HC_DEV void fgk_code_bit(u32 a, u32 &hi, u32 &lo)
{
    lo = funnel_r(lo, hi, 1);
    hi = (hi >> 1) | ((a << 28) & 0x80000000u);
}

// Ordering of lane 0's weight store against the other lanes' reads of the same level.  On the
// GPU a converged warp executes the predicated store after the (earlier) load instruction of all
// lanes, so the product build adds nothing; -DHC_FGK_STRICT (and the emulator) insert a warp
// barrier per level for tools that check the CUDA memory model formally (racecheck).
#if defined(HC_EMU) || defined(HC_FGK_STRICT)
#define FGK_LEVEL_SYNC() syncwarp()
#else
#define FGK_LEVEL_SYNC() ((void)0)
#endif

// `sc` (structure changed) is set when the swap moved an internal node.
// slow path of one level: w[s+1] == w[s].  Finds the block leader (32 slots per ballot) and swaps
// the subtrees if required (src/huffman.cpp:115-122).  Returns true if a swap happened; a / parent
// are updated to the slot the node now occupies.  `watch` is the next symbol to be coded: if its
// leaf moves, *moved is set so that the caller refreshes its prefetched slot.
HC_DEV bool fgk_leader_swap(FgkCtx &c, u32 &a, u32 &parent, u32 ws, u32 lane, u32 watch, bool &moved, bool &sc)
{
    u32 l = a + 8u;
    // the common short block: one more plain load decides it without a ballot
    if (lds32(a + 16u) == ws) {
        u32 run;
        l = a;
        do {
            u32 pa = l + 8u * (1u + lane);
            pa = pa < c.sentinel ? pa : c.sentinel;
            u32 m = ~ballot(lds32(pa) == ws);
            run = m ? (u32)ffs(m) - 1u : 32u;
            l += 8u * run;
        } while (run == 32u);
    }
    if (l == parent) return false;
    // exchange the contents of slots a and l; every lane reads, lane 0 writes
    const u32 da = fgk_down_of(c, a), dl = fgk_down_of(c, l);
    const u32 ka = lds32(da), kl = lds32(dl);
    syncwarp();
    if (lane == 0) {
        sts32(da, kl);
        sts32(dl, ka);
        if (!(kl & 1u)) { u32 cu = fgk_up_of(c, kl); sts32(cu + 4u, a); sts32(cu + 12u, a); }
        else if (kl != FGK_LEAF_NYT) sts16(c.slot_of + (kl & ~1u), (a - c.up) >> 3);
        if (!(ka & 1u)) { u32 cu = fgk_up_of(c, ka); sts32(cu + 4u, l); sts32(cu + 12u, l); }
        else if (ka != FGK_LEAF_NYT) sts16(c.slot_of + (ka & ~1u), (l - c.up) >> 3);
    }
    if (kl == FGK_LEAF_NYT) c.nyt = a;
    if (ka == FGK_LEAF_NYT) c.nyt = l;
    if (!(ka & kl & 1u)) sc = true;                       // an internal node moved: the path table is stale
    const u32 wleaf = (watch << 1) | 1u;
    if (ka == wleaf || kl == wleaf) moved = true;
    a = l;
    parent = lds32(l + 4u);
    return true;
}

// FGK update from the node at up address a (src/huffman.cpp:113-127); `count` = symbols processed
// including this one (= the new root weight).  Nothing written at one level is read again at a
// higher level of the same walk (parents, leaders and probes all have larger slot numbers); the
// barrier at the end publishes lane 0's writes before the next symbol.  The parent's entry is
// fetched one level ahead (software pipelining of the dependent shared-memory loads).
HC_DEV void fgk_update_plain(FgkCtx &c, u32 a, u32 lane, u32 count, u32 watch, bool &moved, bool &sc)
{
    const bool w0 = lane == 0;
    if (a != c.root) {
        uint2 n = lds64(a);
        u32 w1 = lds32(a + 8u);
        // two levels per iteration so that the prefetched entry (pn / n) alternates between two
        // register sets instead of being copied every level
        for (;;) {
            u32 p = n.y;                                   // level A: node a, entry n
            uint2 pn = lds64(p);
            u32 pw1 = lds32(p + 8u);
            if (w1 == n.x && fgk_leader_swap(c, a, p, n.x, lane, watch, moved, sc)) {
                pn = lds64(p);
                pw1 = lds32(p + 8u);
            }
            FGK_LEVEL_SYNC();
            sts32_if(w0, a, n.x + 1u);
            if (p == c.root) break;
            a = pn.y;                                      // level B: node p, entry pn
            n = lds64(a);
            w1 = lds32(a + 8u);
            if (pw1 == pn.x && fgk_leader_swap(c, p, a, pn.x, lane, watch, moved, sc)) {
                n = lds64(a);
                w1 = lds32(a + 8u);
            }
            FGK_LEVEL_SYNC();
            sts32_if(w0, p, pn.x + 1u);
            if (a == c.root) break;
        }
    }
    sts32_if(w0, c.root, count);
    syncwarp();
}

// Out-of-line forms of the sequential walks for the rare callers (nodes deeper than the path table,
// the first occurrence of a symbol): one copy of the code instead of one per call site keeps the
// kernels inside the instruction cache.  The context travels by value so that the caller's copy can
// stay in registers; the full table rebuild is included.
HC_DEV_NOINLINE FgkCtx fgk_update_plain_cold(FgkCtx c, u32 a, u32 lane, u32 count, u32 watch)
{
    bool moved = false, sc = false;
    fgk_update_plain(c, a, lane, count, watch, moved, sc);
    if (sc) fgk_rebuild(c, lane);
    c.fl = moved ? 1u : 0u;
    return c;
}

// FGK update of node a whose path is in the table (pf = its pfx entry): lane j takes the node at
// depth j + 1, one ballot tells which levels need the leader search.  Levels below the deepest such
// level are plain increments and done at once; that level is handled by fgk_leader_swap, after which
// the walk continues from the (possibly new) parent with a fresh table lookup.
// A1 (optional, first round only): the caller already knows the node of every level (decode found
// them while resolving the code) -- lane j holds the up address of the node at depth j + 1.
// W1 / W1n: the weights of that node and of the slot after it, fetched by the caller together with
// its own lookups (lanes without a level hold the root, which never ties).
HC_DEV void fgk_update_fast(FgkCtx &c, u32 a, u32 pf, u32 lane, u32 count, u32 watch, bool &moved, bool have_a1 = false,
                            u32 A1 = 0, u32 W1 = 0, u32 W1n = 0)
{
    for (;;) {
        if (pf == FGK_NOPATH) {                           // deeper than the table: sequential walk
            c = fgk_update_plain_cold(c, a, lane, count, watch);
            if (c.fl & 1u) moved = true;
            return;
        }
        const u32 depth = pf >> 12, path = pf & 0xfffu;
        const bool valid = lane < depth;
        u32 A = c.root;                                   // idle lanes: the root never ties (sentinel above it)
        u32 W, w1;
        if (have_a1) {
            A = A1; W = W1; w1 = W1n;
            have_a1 = false;
        } else {
            if (valid) {
                A = a;
                if (lane + 1u < depth) A = c.up - 8u + 8u * lds16(c.pt + 2u * ((2u << lane) - 2u + (path >> (depth - 1u - lane))));
            }
            W = lds32(A);
            w1 = lds32(A + 8u);
        }
        const u32 tm = ballot(valid && w1 == W);
        FGK_LEVEL_SYNC();
        if (tm == 0u) {
            sts32_if(valid, A, W + 1u);
            sts32_if(lane == 0, c.root, count);
            syncwarp();
            return;
        }
        const u32 k0 = 31u - (u32)clz(tm);                // deepest level with a tie
        sts32_if(valid && lane > k0, A, W + 1u);          // plain levels below it
        u32 ak = shfl(A, (int)k0);
        const u32 wk = shfl(W, (int)k0);
        u32 parent = shfl(A, (int)(k0 ? k0 - 1u : 0u));
        if (k0 == 0u) parent = c.root;
        bool sc = false;
        fgk_leader_swap(c, ak, parent, wk, lane, watch, moved, sc);
        FGK_LEVEL_SYNC();
        sts32_if(lane == 0, ak, wk + 1u);
        if (sc) {
            // an internal node moved: only the entries below the two slots change.  The old slot is
            // the node of level k0 of this path; ak is now the leader's slot
            const u32 pf_a = ((k0 + 1u) << 12) | (path >> (depth - 1u - k0));
            fgk_rebuild_pair(c, pf_a, lds16(c.pfx + ((ak - c.up) >> 2)), lane);
        }
        if (parent == c.root) {
            sts32_if(lane == 0, c.root, count);
            syncwarp();
            return;
        }
        syncwarp();
        a = parent;
        pf = lds16(c.pfx + ((a - c.up) >> 2));            // 2 bytes per slot, 8 bytes per up entry
    }
}

// fgk_update_plain that also finishes the code of the old path (cursor q) on the way: after the
// first swap the update continues on another branch of the tree while the code still has to follow
// the old one.  The two link chases are independent chains, so interleaving them level by level
// hides the latency of one behind the other.  A further swap (rare) could re-link nodes of the old
// path, so the chase is completed before any leader search.
HC_DEV void fgk_update_chase(FgkCtx &c, u32 a, u32 lane, u32 count, u32 q, u32 &hi, u32 &lo, u32 watch, bool &moved, bool &sc)
{
    const bool w0 = lane == 0;
    if (a != c.root) {
        uint2 n = lds64(a);
        u32 w1 = lds32(a + 8u);
        for (;;) {
            u32 p = n.y;
            uint2 pn = lds64(p);
            u32 pw1 = lds32(p + 8u);
            u32 qn = lds32(q + 4u);                        // unused when q is the root
            if (w1 == n.x) {
                for (; q != c.root; q = lds32(q + 4u)) fgk_code_bit(q, hi, lo);
                if (fgk_leader_swap(c, a, p, n.x, lane, watch, moved, sc)) {
                    pn = lds64(p);
                    pw1 = lds32(p + 8u);
                }
            }
            if (q != c.root) { fgk_code_bit(q, hi, lo); q = qn; }
            FGK_LEVEL_SYNC();
            sts32_if(w0, a, n.x + 1u);
            if (p == c.root) break;
            a = p;
            n = pn;
            w1 = pw1;
        }
    }
    for (; q != c.root; q = lds32(q + 4u)) fgk_code_bit(q, hi, lo);
    sts32_if(w0, c.root, count);
    syncwarp();
}

// same walk, also collecting the code of the start node in the PRE-update tree
// (encode precedes update, src/transform.cpp:372-375).  hi:lo must enter as 0x80000000:0.
// The update path leaves the code path at the first swap; the rest of the old path is then
// finished by a pure parent chase (safe: the first swap never re-parents an old-path node).
HC_DEV void fgk_update_coding(FgkCtx &c, u32 a, u32 lane, u32 count, u32 &hi, u32 &lo, u32 watch, bool &moved, bool &sc)
{
    const bool w0 = lane == 0;
    uint2 n = lds64(a);                       // a is a leaf: never the root
    u32 w1 = lds32(a + 8u);
    // two levels per iteration (register sets alternate, see fgk_update_plain)
    for (;;) {
        u32 p = n.y;                                       // level A: node a, entry n
        uint2 pn = lds64(p);
        u32 pw1 = lds32(p + 8u);
        fgk_code_bit(a, hi, lo);
        if (w1 == n.x) {
            const u32 old_parent = p;
            if (fgk_leader_swap(c, a, p, n.x, lane, watch, moved, sc)) {
                FGK_LEVEL_SYNC();
                sts32_if(w0, a, n.x + 1u);
                fgk_update_chase(c, p, lane, count, old_parent, hi, lo, watch, moved, sc);
                return;
            }
        }
        FGK_LEVEL_SYNC();
        sts32_if(w0, a, n.x + 1u);
        if (p == c.root) break;
        a = pn.y;                                          // level B: node p, entry pn
        n = lds64(a);
        w1 = lds32(a + 8u);
        fgk_code_bit(p, hi, lo);
        if (pw1 == pn.x) {
            const u32 old_parent = a;
            if (fgk_leader_swap(c, p, a, pn.x, lane, watch, moved, sc)) {
                FGK_LEVEL_SYNC();
                sts32_if(w0, p, pn.x + 1u);
                fgk_update_chase(c, a, lane, count, old_parent, hi, lo, watch, moved, sc);
                return;
            }
        }
        FGK_LEVEL_SYNC();
        sts32_if(w0, p, pn.x + 1u);
        if (a == c.root) break;
    }
    sts32_if(w0, c.root, count);
    syncwarp();
}

// The trees of the resident CTAs no longer hold the whole batch at once, so streams are started in
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

// MSB-first bit writer: one 32-bit word per lane, 128-byte coalesced flushes
struct BitWriter {
    u64 acc;
    u32 nacc;      // valid low bits of acc (< 32 between calls)
    u32 widx;      // words produced so far
    u32 mine;      // this lane's word of the current 32-word group
    u32 *dst;
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

HC_DEV void bw_put(BitWriter &b, u32 v, u32 d, u32 lane)   // 0 <= d <= 32, v < 2^d
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

// emit the code collected in hi:lo (marker scheme of fgk_code_bit); returns false if the code is
// longer than 56 bits (impossible below 2^32 symbols)
HC_DEV bool bw_put_code(BitWriter &b, u32 hi, u32 lo, u32 lane)
{
    if (lo == 0u) {                                   // depth <= 31: everything is in hi
        const u32 d = 32u - (u32)ffs(hi);             // marker = lowest set bit of hi
        if (d) bw_put(b, hi >> (32u - d), d, lane);
        return true;
    }
    const u32 d2 = 32u - (u32)ffs(lo);                // bits of the code that live in lo
    bw_put(b, hi, 32, lane);
    if (d2) bw_put(b, lo >> (32u - d2), d2, lane);
    return d2 <= 24u;
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

    const u64 m = sym_len[f];
    const u32 *src = (const u32 *)(sym + sym_off[f]);     // 256-byte aligned region
    BitWriter bw;
    bw_init(bw, out + out_off[f], out_cap[f] & ~(u64)3);
    // container header <u64 LE count><u8 flags> (src/headers.cpp:107-125) through the bit writer
    bw_put(bw, bswap32((u32)m), 32, lane);
    bw_put(bw, bswap32((u32)(m >> 32)), 32, lane);
    bw_put(bw, flags ? flags[f] : 0u, 8, lane);

    bool too_long = m > 0xfffffff0ull;                    // weights are 32-bit
    u32 chunk = 0;                                        // 4 symbols per lane, 128 per warp
    if (m > 0) chunk = ldg32(src + lane);
    u32 count = 0;
    for (u64 i0 = 0; i0 < m; i0 += 128) {
        syncwarp();                                       // every lane is done with the previous buffer
        sts32(c.buf + 4u * lane, chunk);
        syncwarp();
        if (i0 + 128 < m) chunk = ldg32(src + (i0 + 128) / 4 + lane);   // prefetch the next 128 symbols
        const u32 cnt = (m - i0) < 128 ? (u32)(m - i0) : 128u;
        u32 y = lds8(c.buf);
        u32 slot = lds16(c.slot_of + 2u * y);
        for (u32 i = 0; i < cnt; i++) {
            // fetch the next symbol and its leaf slot before the update; the update reports if
            // that leaf moved (swap) or was created (same new symbol twice) so the slot is re-read
            const u32 yn = lds8(c.buf + ((i + 1u) & 127u));
            u32 slot_n = lds16(c.slot_of + 2u * yn);
            bool moved = false;
            count++;
            if (slot == 0xffffu) {
                // not yet transmitted: NYT code followed by the 8 raw bits (src/huffman.cpp:42-51)
                u32 hi = 0x80000000u, lo = 0u;
                for (u32 p = c.nyt; p != c.root; p = lds32(p + 4u)) fgk_code_bit(p, hi, lo);
                if (!bw_put_code(bw, hi, lo, lane)) too_long = true;
                bw_put(bw, y, 8, lane);
                const u32 nsl = c.nyt, npf = nsl == c.root ? 0u : lds16(c.pfx + ((nsl - c.up) >> 2));
                const u32 leaf = fgk_split(c, y, lane);
                fgk_table_split(c, nsl, npf, lane);
                c = fgk_update_plain_cold(c, leaf, lane, count, yn);
                if (c.fl & 1u) moved = true;
                if (yn == y) moved = true;
            } else {
                const u32 pf = lds16(c.pfx + 2u * slot);
                if (pf != FGK_NOPATH) {
                    // the code of a leaf is its path (encode precedes update, src/transform.cpp:372-375)
                    bw_put(bw, pf & 0xfffu, pf >> 12, lane);
                    fgk_update_fast(c, c.up + 8u * slot, pf, lane, count, yn, moved);
                } else {
                    u32 hi = 0x80000000u, lo = 0u;
                    bool sc = false;
                    fgk_update_coding(c, c.up + 8u * slot, lane, count, hi, lo, yn, moved, sc);
                    if (sc) fgk_rebuild(c, lane);
                    if (!bw_put_code(bw, hi, lo, lane)) too_long = true;
                }
            }
            if (moved) slot_n = lds16(c.slot_of + 2u * yn);
            y = yn;
            slot = slot_n;
        }
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
    u32 wbits;      // valid bits in win (> 32 between calls)
    u32 ridx;       // next word index to pull into the window
    u32 chunk, chunk_next;
    const u32 *src;
    u64 nwords;     // words that may be loaded
    u64 avail;      // bits of the file not yet consumed
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
    r.nwords = (len_bytes + 3u) / 4u;      // reads stay inside the padded region
    r.avail = len_bytes * 8u;
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
    r.avail -= d;
    if (r.wbits <= 32u) br_refill(r, lane);
}

HC_DEV u32 br_get(BitReader &r, u32 d, u32 lane)   // 1..32 bits; caller checks r.avail first
{
    u32 v = (u32)(r.win >> (64u - d));
    br_skip(r, d, lane);
    return v;
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
    const u64 m = ((u64)hi << 32) | lo;
    u32 fl = br_get(br, 8, lane);
    if (lane == 0 && flags) flags[f] = (u8)fl;
    const u64 cap = sym_cap[f];
    if (m > cap || m > br.avail + 1u || m > 0xfffffff0ull) {
        // every symbol after the first costs >= 1 bit: a count beyond avail+1 is a guaranteed
        // underrun, for which the reference exits with 9 (src/transform.cpp:394-398)
        if (lane == 0) { sym_len[f] = m; status[f] = (m > br.avail + 1u) ? 9 : 100; }
        return;
    }
    u32 *dst = (u32 *)(sym + sym_off[f]);
    const u32 root_down = c.down + 4u * FGK_ROOT;
    i32 err = 0;
    u32 count = 0;
    for (u64 i = 0; i < m; i++) {
        // The path table resolves the next FGK_D code bits in one step: lane j looks up the node that
        // the first j + 1 bits lead to; the shallowest leaf among them ends the code.
        u32 t = (u32)(br.win >> 32), depth = 0, k = 0, d = 0, dk = 0, aleaf = 0;
        u32 e = 0, kj = 0;
        if (lane < FGK_D) e = lds16(c.pt + 2u * ((2u << lane) - 2u + (t >> (31u - lane))));
        if (e) kj = lds32(c.down + 4u * (e - 1u));
        // the update's first question ("does a level need a leader search?") rides on the same round
        // of loads: weights of the node and of the slot after it (no node: the root, which never ties)
        const u32 A1 = e ? c.up - 8u + 8u * e : c.root;
        const u32 W1 = lds32(A1), W1n = lds32(A1 + 8u);
        const u32 leafm = ballot(e != 0u && (kj & 1u));
        const bool via_table = leafm != 0u;
        if (via_table) {
            depth = (u32)ffs(leafm);
            k = shfl(kj, (int)(depth - 1u));
            aleaf = c.up - 8u + 8u * shfl(e, (int)(depth - 1u));
            if (br.avail < depth) { err = 9; break; }    // ran out of bits inside a code
        } else {
            // the root is still a leaf, or the code is longer than the table: walk down on a private
            // copy of the window's top 32 bits, consume them afterwards.  Lane j remembers the node
            // reached after j+1 steps.  No length check inside the loop: past 32 steps `t` only supplies
            // zeros, i.e. the walk keeps taking (valid) left children and still ends at a leaf; such a
            // symbol is redone below.
            d = root_down; k = lds32(d);
            while (!(k & 1u)) {
                d = k + ((t >> 29) & 4u);
                if (lane == depth) dk = d;
                t <<= 1;
                depth++;
                k = lds32(d);
            }
            if (depth > 32u) {
                // code longer than 32 bits (very deep tree): exact walk, 32 bits at a time
                d = root_down; k = lds32(d); t = (u32)(br.win >> 32); depth = 0;
                u32 len = 0;
                while (!(k & 1u)) {
                    if (len == 32u) {
                        if (br.avail < 32u) { err = 9; break; }
                        br_skip(br, 32, lane);
                        t = (u32)(br.win >> 32);
                        len = 0;
                    }
                    d = k + ((t >> 29) & 4u);
                    t <<= 1;
                    len++;
                    depth++;
                    k = lds32(d);
                }
                if (err) break;
                if (len) {
                    if (br.avail < len) { err = 9; break; }
                    br_skip(br, len, lane);
                }
            } else if (depth) {
                if (br.avail < depth) { err = 9; break; }
            }
        }
        const u32 pf_leaf = (depth << 12) | (t >> ((32u - depth) & 31u));   // used on the table path only (t intact, depth >= 1)
        if (depth && depth <= 32u) br_skip(br, depth, lane);
        count++;
        u32 y;
        bool moved = false, sc = false;
        if (k == FGK_LEAF_NYT) {
            if (br.avail < 8u) { err = 9; break; }
            y = br_get(br, 8, lane);
            // a raw symbol that is already in the tree is still decoded as that symbol
            // (src/huffman.cpp:74-86); the update then starts from its existing leaf
            const u32 ex = lds16(c.slot_of + 2u * y);
            if (lane == 0) sts8(c.buf + (u32)(i & 127u), y);
            if (ex == 0xffffu) {
                const u32 nsl = c.nyt, npf = nsl == c.root ? 0u : lds16(c.pfx + ((nsl - c.up) >> 2));
                const u32 a = fgk_split(c, y, lane);
                fgk_table_split(c, nsl, npf, lane);
                c = fgk_update_plain_cold(c, a, lane, count, 0x1ffu);     // ends with a warp barrier
            } else {
                fgk_update_fast(c, c.up + 8u * ex, lds16(c.pfx + 2u * ex), lane, count, 0x1ffu, moved);
            }
        } else {
            y = k >> 1;
            if (lane == 0) sts8(c.buf + (u32)(i & 127u), y);
            if (via_table) {
                fgk_update_fast(c, aleaf, pf_leaf, lane, count, 0x1ffu, moved, true, A1, W1, W1n);
            } else if (depth <= 32u) {
                // all levels at once, one lane per level (lane depth-1 = the leaf): a level whose next
                // slot carries the same weight needs leader search / swap and everything above it may
                // move, so the sequential walk takes over from the deepest such level.
                const bool valid = lane < depth;
                const u32 A = valid ? fgk_up_of(c, dk) : c.sentinel - 8u;
                const u32 W = lds32(A), w1 = lds32(A + 8u);
                const u32 tm = ballot(valid && w1 == W);
                FGK_LEVEL_SYNC();
                if (tm == 0u) {
                    sts32_if(valid, A, W + 1u);
                    sts32_if(lane == 0, c.root, count);
                    syncwarp();
                } else {
                    const u32 k0 = 31u - (u32)clz(tm);                   // deepest level with a tie
                    sts32_if(valid && lane > k0, A, W + 1u);             // plain levels below it
                    c = fgk_update_plain_cold(c, shfl(A, (int)k0), lane, count, 0x1ffu);
                }
            } else {
                c = fgk_update_plain_cold(c, fgk_up_of(c, d), lane, count, 0x1ffu);
            }
        }
        if ((i & 127u) == 127u) {
            stg32_stream(dst + (i >> 7) * 32u + lane, lds32(c.buf + 4u * lane));
            syncwarp();
        }
    }
    if (!err) {
        u32 rem = (u32)(m & 127u);                         // symbols in the last partial group
        if (rem && lane < (rem + 3u) / 4u) dst[(m >> 7) * 32u + lane] = lds32(c.buf + 4u * lane);
    }
    if (lane == 0) { sym_len[f] = m; status[f] = err; }
}

}  // namespace hcd
