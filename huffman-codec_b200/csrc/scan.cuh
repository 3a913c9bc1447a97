// scan.cuh -- block-wide scans over a "striped" tile, shared by the diff/RLE kernels.
//
// Tile geometry used by every streaming kernel: TPB = 256 threads, UN = 4 sub-tiles.
// A tile is UN * TPB * 16 bytes = 16 KiB.  Thread t owns the 16 bytes
//     [ j*TPB*16 + t*16 , +16 )   of sub-tile j   (j = 0..UN-1)
// so every 128-bit load/store instruction of a warp covers 512 contiguous bytes
// (fully coalesced).  Sequence order inside a tile is (sub-tile j, warp w, lane l).
#pragma once
#include "hc_common.cuh"

namespace hcd {

constexpr int TPB = 256;
constexpr int UN = 4;
constexpr int NW = TPB / 32;
constexpr u32 SUB_BYTES = TPB * 16;       // 4 KiB
constexpr u32 TILE_BYTES = UN * SUB_BYTES; // 16 KiB
static_assert(UN * NW == 32, "cross-warp phase assumes exactly 32 (sub-tile, warp) partials");

// Exclusive scan of one u32 per (thread, sub-tile) in sequence order with associative `op`
// (op(earlier, later); need not commute).  wtot: 32 u32 of shared memory that no other
// scan touches until the next __syncthreads after this call returns.
// excl[j] = fold of everything before (j, thread); returns the fold of the whole tile.
template <class Op>
HC_DEV u32 block_scan_striped(const u32 (&v)[UN], u32 (&excl)[UN], u32 identity, Op op, u32 *wtot)
{
    const u32 lane = lane_id(), w = warp_id();
    u32 inc[UN];
#pragma unroll
    for (int j = 0; j < UN; j++) inc[j] = v[j];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int j = 0; j < UN; j++) {
            u32 t = shfl_up(inc[j], d);
            if (lane >= (u32)d) inc[j] = op(t, inc[j]);
        }
    }
    if (lane == 31) {
#pragma unroll
        for (int j = 0; j < UN; j++) wtot[j * NW + w] = inc[j];
    }
    syncthreads();
    // every warp scans the 32 partials redundantly (saves a second barrier)
    u32 p = wtot[lane];
    u32 pin = p;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = shfl_up(pin, d);
        if (lane >= (u32)d) pin = op(t, pin);
    }
    u32 total = shfl(pin, 31);
    u32 pex = shfl_up(pin, 1);
    if (lane == 0) pex = identity;
#pragma unroll
    for (int j = 0; j < UN; j++) {
        u32 base = shfl(pex, j * NW + (int)w);
        u32 le = shfl_up(inc[j], 1);
        if (lane == 0) le = identity;
        excl[j] = op(base, le);
    }
    return total;
}

struct OpAdd { HC_DEVM u32 operator()(u32 a, u32 b) const { return a + b; } };
struct OpMax { HC_DEVM u32 operator()(u32 a, u32 b) const { return a > b ? a : b; } };
struct OpAdd4 { HC_DEVM u32 operator()(u32 a, u32 b) const { return vadd4(a, b); } };

// number of grid segments per file so that nf*nseg CTAs fill the machine (>= ~4 CTAs/SM on
// 148 SMs) while a segment stays >= 4 tiles; seg_bytes is a multiple of TILE_BYTES.
struct SegPlan { u32 nseg; u64 seg_bytes; };
static inline SegPlan plan_segments(u32 nf, u64 max_len)
{
    SegPlan p;
    u64 tiles = (max_len + TILE_BYTES - 1) / TILE_BYTES;
    if (tiles == 0) tiles = 1;
    u64 want = nf ? (592 + nf - 1) / nf : 1;   // 148 SMs x 4 resident CTAs
    u64 max_seg = (tiles + 3) / 4;             // keep >= 4 tiles per segment
    if (max_seg == 0) max_seg = 1;
    u64 nseg = want < max_seg ? want : max_seg;
    if (nseg == 0) nseg = 1;
    if (nseg > 65535) nseg = 65535;
    u64 tps = (tiles + nseg - 1) / nseg;
    p.seg_bytes = tps * TILE_BYTES;
    p.nseg = (u32)((tiles + tps - 1) / tps);
    return p;
}

}  // namespace hcd
