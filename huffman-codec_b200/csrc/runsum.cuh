// runsum.cuh -- the MNP-5 run-length cost model and the associative "run summary".
//
// Closed form of the reference encoder (src/transform.cpp:241-279): inside a
// sequence whose LAST element is always emitted as its own literal, every maximal run of L
// equal bytes (taken over elements 0..n-2) costs
//        rle_size(L) = 4*floor(L/258) + (r < 3 ? r : 4),   r = L mod 258.
// Element with index k inside its run, q = k mod 258, emits
//        [q < 3] literal   +   [q == 257] the byte 255   +   [last of run, 2 <= q < 257] count q-2.
// A RunSum describes a byte sequence by its first/last byte, the lengths of the runs touching
// both ends and the cost of the runs strictly inside; rs_combine() is associative, so block
// rows, block columns, lane chunks and tiles can be reduced in any grouping.
#pragma once
#include "hc_common.cuh"

namespace hcd {

HC_HD u32 rle_size(u32 L)
{
    u32 r = L % 258u;
    return 4u * (L / 258u) + (r < 3u ? r : 4u);
}

// bytes already emitted by a run of length L; `ended` adds the pending count byte
HC_HD u32 rle_emitted(u32 L, bool ended)
{
    u32 r = L % 258u;
    return 4u * (L / 258u) + (r < 3u ? r : 3u) + ((ended && r >= 3u) ? 1u : 0u);
}

struct RunSum {
    u32 head;   // length of the run that contains the first element
    u32 tail;   // length of the run that contains the last element
    u32 inner;  // sum of rle_size over runs touching neither end
    u32 meta;   // first | last << 8 | nonempty << 16 | allsame << 17
};

HC_HD u32 rs_first(const RunSum &s) { return s.meta & 0xffu; }
HC_HD u32 rs_last(const RunSum &s) { return (s.meta >> 8) & 0xffu; }
HC_HD bool rs_nonempty(const RunSum &s) { return (s.meta >> 16) & 1u; }
HC_HD bool rs_allsame(const RunSum &s) { return (s.meta >> 17) & 1u; }
HC_HD u32 rs_meta(u32 first, u32 last, bool all) { return first | (last << 8) | (1u << 16) | ((all ? 1u : 0u) << 17); }

HC_HD RunSum rs_empty()
{
    RunSum s;
    s.head = s.tail = s.inner = s.meta = 0;
    return s;
}

// append one element
HC_HD void rs_push(RunSum &s, u32 b)
{
    if (!rs_nonempty(s)) {
        s.head = s.tail = 1;
        s.inner = 0;
        s.meta = rs_meta(b, b, true);
    } else if (b == rs_last(s)) {
        s.tail++;
        if (rs_allsame(s)) s.head++;
    } else {
        if (!rs_allsame(s)) s.inner += rle_size(s.tail);
        s.tail = 1;
        s.meta = rs_meta(rs_first(s), b, false);
    }
}

HC_HD RunSum rs_combine(const RunSum &a, const RunSum &b)
{
    if (!rs_nonempty(a)) return b;
    if (!rs_nonempty(b)) return a;
    const bool aa = rs_allsame(a), ba = rs_allsame(b);
    RunSum r;
    if (rs_last(a) == rs_first(b)) {
        if (aa && ba) {
            r.head = r.tail = a.head + b.head;
            r.inner = 0;
        } else if (aa) {
            r.head = a.head + b.head;
            r.tail = b.tail;
            r.inner = b.inner;
        } else if (ba) {
            r.head = a.head;
            r.tail = a.tail + b.head;
            r.inner = a.inner;
        } else {
            r.head = a.head;
            r.tail = b.tail;
            r.inner = a.inner + b.inner + rle_size(a.tail + b.head);
        }
        r.meta = rs_meta(rs_first(a), rs_last(b), aa && ba);
    } else {
        r.head = a.head;
        r.tail = b.tail;
        r.inner = a.inner + b.inner + (aa ? 0u : rle_size(a.tail)) + (ba ? 0u : rle_size(b.head));
        r.meta = rs_meta(rs_first(a), rs_last(b), false);
    }
    return r;
}

// applyRLE output size of the complete sequence summarised by s (forced-literal last byte)
HC_HD u32 rs_cost_final(const RunSum &s)
{
    if (!rs_nonempty(s)) return 0;
    if (rs_allsame(s)) return rle_size(s.head - 1) + 1;
    return rle_size(s.head) + s.inner + rle_size(s.tail - 1) + 1;
}

// bytes emitted so far by a prefix summarised by s whose last run is still open; `ended`
// tells whether the element that follows the prefix differs (or is the forced final literal)
HC_HD u32 rs_emitted_prefix(const RunSum &s, bool ended)
{
    if (!rs_nonempty(s)) return 0;
    if (rs_allsame(s)) return rle_emitted(s.head, ended);
    return rle_size(s.head) + s.inner + rle_emitted(s.tail, ended);
}

}  // namespace hcd
