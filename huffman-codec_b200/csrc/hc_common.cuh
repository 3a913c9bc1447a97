// hc_common.cuh -- shared device helpers for the sm_100a kernels of libhc_b200.
//
// Product build: nvcc -gencode arch=compute_100a,code=sm_100a.  Every helper maps 1:1 to a
// CUDA intrinsic.  The -DHC_EMU branch exists only for the test double built by tests/backend.py
// (tests/emu/hc_emu.h, found through that build's -I tests/emu; the product build never sees it).
#pragma once

#include <stddef.h>
#include <stdint.h>

#ifdef HC_EMU
#include "hc_emu.h"
#define HC_KERNEL static void
#define HC_DEV static inline
#define HC_HD static inline
#define HC_DEVM inline
#define HC_DEV_NOINLINE static __attribute__((noinline))
#define HC_SHARED static
#define HC_DYN_SMEM(name) unsigned char *name = (unsigned char *)hc_emu::dyn_smem()
#define HC_LAUNCH(kern, grid, block, smem, stream, ...)                                   \
    do {                                                                                  \
        hc_count_launch();                                                                \
        hc_emu::launch((grid), (block), (smem), [&]() { kern(__VA_ARGS__); });            \
    } while (0)
#define HC_LAUNCH_BOUNDS(t, b)
#define HC_RESTRICT
#define HC_ALIGNED16 __attribute__((aligned(16)))
#define HC_ALIGNED(n) __attribute__((aligned(n)))
#define HC_DEVICE_CONST static constexpr
#else
#include <cuda_runtime.h>
#define HC_KERNEL __global__ void
#define HC_DEV __device__ __forceinline__
#define HC_HD __host__ __device__ __forceinline__
#define HC_DEVM __device__ __forceinline__
#define HC_DEV_NOINLINE __device__ __noinline__
#define HC_SHARED __shared__
#define HC_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define HC_LAUNCH(kern, grid, block, smem, stream, ...)                                   \
    do {                                                                                  \
        hc_count_launch();                                                                \
        kern<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);           \
    } while (0)
#define HC_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
#define HC_RESTRICT __restrict__
#define HC_ALIGNED16 __align__(16)
#define HC_ALIGNED(n) __align__(n)
#define HC_DEVICE_CONST __device__ constexpr
#endif

void hc_count_launch();

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef int16_t i16;

namespace hcd {

constexpr u32 FULL = 0xffffffffu;

#ifdef HC_EMU
// ---------------------------------------------------------------- emulation
#ifdef HC_EMU_DEBUG
#define HC_SITE , __builtin_return_address(0)
#undef HC_DEV
#define HC_DEV static __attribute__((noinline))
#else
#define HC_SITE
#endif
HC_DEV void syncthreads() { hc_emu::syncthreads(); }
HC_DEV void syncwarp() { hc_emu::warp_exchange(0 HC_SITE); }
HC_DEV u32 lane_id() { return threadIdx.x & 31u; }
HC_DEV u64 shfl64(u64 v, int src) { return hc_emu::warp_exchange(v HC_SITE)[src & 31]; }
HC_DEV u64 shfl_up64(u64 v, unsigned d)
{
    unsigned l = hc_emu::lane();
    const u64 *s = hc_emu::warp_exchange(v HC_SITE);
    return l >= d ? s[l - d] : v;
}
HC_DEV u64 shfl_down64(u64 v, unsigned d)
{
    unsigned l = hc_emu::lane();
    const u64 *s = hc_emu::warp_exchange(v HC_SITE);
    return l + d < 32 ? s[l + d] : v;
}
HC_DEV u64 shfl_xor64(u64 v, unsigned m)
{
    unsigned l = hc_emu::lane();
    return hc_emu::warp_exchange(v HC_SITE)[l ^ m];
}
HC_DEV u32 shfl(u32 v, int src) { return (u32)shfl64(v, src); }
HC_DEV u32 shfl_up(u32 v, unsigned d) { return (u32)shfl_up64(v, d); }
HC_DEV u32 shfl_down(u32 v, unsigned d) { return (u32)shfl_down64(v, d); }
HC_DEV u32 shfl_xor(u32 v, unsigned m) { return (u32)shfl_xor64(v, m); }
HC_DEV u32 ballot(bool p)
{
    const u64 *s = hc_emu::warp_exchange(p ? 1 : 0 HC_SITE);
    u32 live = hc_emu::warp_live(), r = 0;
    for (int i = 0; i < 32; i++)
        if (((live >> i) & 1) && s[i]) r |= 1u << i;
    return r;
}
HC_DEV u32 reduce_max(u32 v)
{
    const u64 *s = hc_emu::warp_exchange(v HC_SITE);
    u32 live = hc_emu::warp_live(), r = 0;
    for (int i = 0; i < 32; i++)
        if (((live >> i) & 1) && (u32)s[i] > r) r = (u32)s[i];
    return r;
}
HC_DEV bool any(bool p)
{
    const u64 *s = hc_emu::warp_exchange(p ? 1 : 0 HC_SITE);
    u32 live = hc_emu::warp_live();
    for (int i = 0; i < 32; i++)
        if (((live >> i) & 1) && s[i]) return true;
    return false;
}
HC_DEV u32 reduce_min(u32 v)
{
    const u64 *s = hc_emu::warp_exchange(v HC_SITE);
    u32 live = hc_emu::warp_live(), r = 0xffffffffu;
    for (int i = 0; i < 32; i++)
        if (((live >> i) & 1) && (u32)s[i] < r) r = (u32)s[i];
    return r;
}
HC_DEV int popc(u32 v) { return __builtin_popcount(v); }
HC_DEV int popcll(u64 v) { return __builtin_popcountll(v); }
HC_DEV int clz(u32 v) { return v ? __builtin_clz(v) : 32; }
HC_DEV int clzll(u64 v) { return v ? __builtin_clzll(v) : 64; }
HC_DEV int ffs(u32 v) { return __builtin_ffs((int)v); }
HC_DEV int ffsll(u64 v) { return __builtin_ffsll((long long)v); }
HC_DEV u32 bswap32(u32 v) { return __builtin_bswap32(v); }
HC_DEV u32 dp4a_u(u32 a, u32 b, u32 c)
{
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xffu) * ((b >> (8 * i)) & 0xffu);
    return c;
}
// PRMT in its default mode: result byte i = byte (sel nibble i) of the 8 bytes {b:a}
HC_DEV u32 prmt(u32 a, u32 b, u32 sel)
{
    const u64 src = ((u64)b << 32) | a;
    u32 r = 0;
    for (int i = 0; i < 4; i++) r |= (u32)((src >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
}
HC_DEV u32 prmt_raw(u32 a, u32 b, u32 sel) { return prmt(a, b, sel & 0x7777u); }
HC_DEV u32 brev(u32 v) { u32 r = 0; for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i); return r; }
HC_DEV u64 bswap64(u64 v) { return __builtin_bswap64(v); }
HC_DEV u32 vadd4(u32 a, u32 b)
{
    u32 r = 0;
    for (int i = 0; i < 4; i++) r |= (u32)(u8)((a >> (8 * i)) + (b >> (8 * i))) << (8 * i);
    return r;
}
HC_DEV u32 vsub4(u32 a, u32 b)
{
    u32 r = 0;
    for (int i = 0; i < 4; i++) r |= (u32)(u8)((a >> (8 * i)) - (b >> (8 * i))) << (8 * i);
    return r;
}
HC_DEV u32 vcmpeq4(u32 a, u32 b)
{
    u32 r = 0;
    for (int i = 0; i < 4; i++) if (((a >> (8 * i)) & 0xffu) == ((b >> (8 * i)) & 0xffu)) r |= 0xffu << (8 * i);
    return r;
}
HC_DEV u32 funnel_l(u32 lo, u32 hi, u32 sh) { return sh & 31 ? (hi << (sh & 31)) | (lo >> (32 - (sh & 31))) : hi; }
HC_DEV u32 atomic_add(u32 *p, u32 v) { u32 o = *p; *p = o + v; return o; }
HC_DEV u64 atomic_add64(u64 *p, u64 v) { u64 o = *p; *p = o + v; return o; }
HC_DEV void atomic_max_i32(i32 *p, i32 v) { if (v > *p) *p = v; }
HC_DEV void atomic_min_smem(u32 *p, u32 v) { if (v < *p) *p = v; }
HC_DEV void atomic_max_smem(u32 *p, u32 v) { if (v > *p) *p = v; }
// "shared-space addresses": offsets from an arena base (the kernel's __shared__ object)
inline unsigned char *g_emu_smem_base = nullptr;
// base = start of the 4 GiB window that holds the kernel's static "shared" objects, so that every
// one of them has a positive 32-bit offset
#define HC_SMEM_ARENA(obj) (hcd::g_emu_smem_base = (unsigned char *)((uintptr_t)&(obj) & ~(uintptr_t)0xffffffffull))
HC_DEV u32 smem_addr(const void *p) { return (u32)((const unsigned char *)p - g_emu_smem_base); }
HC_DEV uint2 lds64(u32 a) { uint2 v; memcpy(&v, g_emu_smem_base + a, 8); return v; }
HC_DEV u32 lds32(u32 a) { u32 v; memcpy(&v, g_emu_smem_base + a, 4); return v; }
HC_DEV u32 lds16(u32 a) { u16 v; memcpy(&v, g_emu_smem_base + a, 2); return v; }
HC_DEV uint4 lds128(u32 a) { uint4 v; memcpy(&v, g_emu_smem_base + a, 16); return v; }
HC_DEV void prefetch_l2(const void *) {}
HC_DEV u32 lds8(u32 a) { return g_emu_smem_base[a]; }
HC_DEV void sts64(u32 a, uint2 v) { memcpy(g_emu_smem_base + a, &v, 8); }
HC_DEV void sts32(u32 a, u32 v) { memcpy(g_emu_smem_base + a, &v, 4); }
HC_DEV void sts16(u32 a, u32 v) { u16 t = (u16)v; memcpy(g_emu_smem_base + a, &t, 2); }
HC_DEV void sts8(u32 a, u32 v) { g_emu_smem_base[a] = (u8)v; }
HC_DEV u32 funnel_r(u32 lo, u32 hi, u32 sh) { return (u32)((((u64)hi << 32) | lo) >> (sh & 31)); }
HC_DEV void sts32_if(bool p, u32 a, u32 v) { if (p) sts32(a, v); }
HC_DEV void atomic_or_smem(u32 a, u32 v) { sts32(a, lds32(a) | v); }
HC_DEV uint4 ldg16(const void *p) { uint4 v; memcpy(&v, p, 16); return v; }
HC_DEV uint4 ldg16_rw(const void *p) { uint4 v; memcpy(&v, p, 16); return v; }
HC_DEV uint4 ldg16_l1(const void *p) { uint4 v; memcpy(&v, p, 16); return v; }
HC_DEV void stg16(void *p, uint4 v) { memcpy(p, &v, 16); }
HC_DEV u32 atomic_or_shared(u32 *p, u32 v) { u32 o = *p; *p = o | v; return o; }
HC_DEV u8 ldg8(const u8 *p) { return *p; }
HC_DEV u32 ldg32(const void *p) { u32 v; memcpy(&v, p, 4); return v; }
HC_DEV void stg32_stream(void *p, u32 v) { memcpy(p, &v, 4); }
#else
// ---------------------------------------------------------------- CUDA
HC_DEV void syncthreads() { __syncthreads(); }
HC_DEV void syncwarp() { __syncwarp(); }
HC_DEV u32 lane_id() { return threadIdx.x & 31u; }
HC_DEV u32 shfl(u32 v, int src) { return __shfl_sync(FULL, v, src); }
HC_DEV u32 shfl_up(u32 v, unsigned d) { return __shfl_up_sync(FULL, v, d); }
HC_DEV u32 shfl_down(u32 v, unsigned d) { return __shfl_down_sync(FULL, v, d); }
HC_DEV u32 shfl_xor(u32 v, unsigned m) { return __shfl_xor_sync(FULL, v, m); }
HC_DEV u64 shfl64(u64 v, int src) { return __shfl_sync(FULL, v, src); }
HC_DEV u64 shfl_up64(u64 v, unsigned d) { return __shfl_up_sync(FULL, v, d); }
HC_DEV u64 shfl_down64(u64 v, unsigned d) { return __shfl_down_sync(FULL, v, d); }
HC_DEV u64 shfl_xor64(u64 v, unsigned m) { return __shfl_xor_sync(FULL, v, m); }
HC_DEV u32 ballot(bool p) { return __ballot_sync(FULL, p); }
HC_DEV bool any(bool p) { return __any_sync(FULL, p) != 0; }            // VOTE.ANY with a predicate result
HC_DEV u32 reduce_max(u32 v) { return __reduce_max_sync(FULL, v); }     // REDUX.MAX.U32: one instruction, uniform result
HC_DEV u32 reduce_min(u32 v) { return __reduce_min_sync(FULL, v); }
HC_DEV int popc(u32 v) { return __popc(v); }
HC_DEV int popcll(u64 v) { return __popcll(v); }
HC_DEV int clz(u32 v) { return __clz((int)v); }
HC_DEV int clzll(u64 v) { return __clzll((long long)v); }
HC_DEV int ffs(u32 v) { return __ffs((int)v); }
HC_DEV int ffsll(u64 v) { return __ffsll((long long)v); }
HC_DEV u32 bswap32(u32 v) { return __byte_perm(v, 0, 0x0123); }
HC_DEV u32 prmt(u32 a, u32 b, u32 sel) { return __byte_perm(a, b, sel); }
// PRMT without the selector masking of __byte_perm: only the low 16 bits of sel are read; every selector nibble
// must be 0..7 (bit 3 would ask for sign replication)
HC_DEV u32 prmt_raw(u32 a, u32 b, u32 sel)
{
    u32 r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
HC_DEV u32 dp4a_u(u32 a, u32 b, u32 c) { return __dp4a(a, b, c); }
HC_DEV u32 brev(u32 v) { return __brev(v); }
HC_DEV u64 bswap64(u64 v)
{
    return ((u64)bswap32((u32)v) << 32) | bswap32((u32)(v >> 32));
}
HC_DEV u32 vadd4(u32 a, u32 b) { return __vadd4(a, b); }
HC_DEV u32 vsub4(u32 a, u32 b) { return __vsub4(a, b); }
HC_DEV u32 vcmpeq4(u32 a, u32 b) { return __vcmpeq4(a, b); }
HC_DEV u32 funnel_l(u32 lo, u32 hi, u32 sh) { return __funnelshift_l(lo, hi, sh); }
HC_DEV u32 atomic_add(u32 *p, u32 v) { return atomicAdd(p, v); }
HC_DEV u64 atomic_add64(u64 *p, u64 v) { return atomicAdd((unsigned long long *)p, (unsigned long long)v); }
HC_DEV void atomic_max_i32(i32 *p, i32 v) { atomicMax(p, v); }
HC_DEV void atomic_min_smem(u32 *p, u32 v) { atomicMin(p, v); }
HC_DEV void atomic_max_smem(u32 *p, u32 v) { atomicMax(p, v); }
// shared-space (32-bit) addressing: tree links are stored as shared addresses so that a walk
// needs no address arithmetic between dependent loads
#define HC_SMEM_ARENA(obj) ((void)0)
HC_DEV u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }
HC_DEV uint2 lds64(u32 a)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
HC_DEV uint4 lds128(u32 a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
// ask L2 for the line that holds p (no register, no dependency)
HC_DEV void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
HC_DEV u32 lds32(u32 a) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
HC_DEV u32 lds16(u32 a) { u32 v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
HC_DEV u32 lds8(u32 a) { u32 v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
HC_DEV void sts64(u32 a, uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(a), "r"(v.x), "r"(v.y) : "memory"); }
HC_DEV void sts32(u32 a, u32 v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
HC_DEV void sts16(u32 a, u32 v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
HC_DEV void sts8(u32 a, u32 v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
HC_DEV u32 funnel_r(u32 lo, u32 hi, u32 sh) { return __funnelshift_r(lo, hi, sh); }
// predicated store: no branch, no divergence (used for "lane 0 is the only writer")
HC_DEV void sts32_if(bool p, u32 a, u32 v)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.shared.u32 [%0], %1; }" :: "r"(a), "r"(v), "r"((u32)p) : "memory");
}
// OR into a shared-memory word by shared-space address, no return value (RED)
HC_DEV void atomic_or_smem(u32 a, u32 v) { asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
// streaming 16-byte load: read-only path, do not allocate in L1 (data is touched once)
HC_DEV uint4 ldg16(const void *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// 16-byte load through L1 (read-only path, allocating): for access patterns in which a sector is shared by
// consecutive loads of the same thread (every thread walks its own contiguous bytes)
HC_DEV uint4 ldg16_l1(const void *p)
{
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// same, for buffers that the kernel also writes (in-place scans): coherent path
HC_DEV uint4 ldg16_rw(const void *p)
{
    uint4 v;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
HC_DEV u32 atomic_or_shared(u32 *p, u32 v) { return atomicOr(p, v); }
HC_DEV void stg16(void *p, uint4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
HC_DEV u8 ldg8(const u8 *p) { return __ldg(p); }
HC_DEV u32 ldg32(const void *p) { return __ldg((const u32 *)p); }
HC_DEV void stg32_stream(void *p, u32 v) { __stcs((u32 *)p, v); }
#endif

HC_DEV u32 warp_id() { return threadIdx.x >> 5; }

// byte k (0..15) of a 16-byte vector (little endian element order)
HC_DEV u32 vec_word(const uint4 &v, int j) { return j == 0 ? v.x : j == 1 ? v.y : j == 2 ? v.z : v.w; }
HC_DEV u8 vec_byte(const uint4 &v, int k) { return (u8)(vec_word(v, k >> 2) >> (8 * (k & 3))); }

HC_DEV uint4 make_uint4_zero() { uint4 v; v.x = v.y = v.z = v.w = 0; return v; }
// keep the first k (1..15) bytes of a 16-byte vector, zero the rest
HC_DEV uint4 mask_tail(uint4 v, u32 k)
{
    u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (u32 j = 0; j < 4; j++) {
        if (k <= 4 * j) w[j] = 0;
        else if (k < 4 * j + 4) w[j] &= (1u << (8 * (k - 4 * j))) - 1u;
    }
    v.x = w[0]; v.y = w[1]; v.z = w[2]; v.w = w[3];
    return v;
}

template <typename T> HC_HD T hmin(T a, T b) { return a < b ? a : b; }
template <typename T> HC_HD T hmax(T a, T b) { return a > b ? a : b; }
HC_HD u64 align_up(u64 v, u64 a) { return (v + a - 1) / a * a; }

}  // namespace hcd
