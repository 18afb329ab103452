// wmix_b200 — shared helpers for the sm_100a kernels.
//
// Every algorithm body in this directory is written as `WMX_HD` (host+device) inline code so
// that the *same source* the kernels run can also be compiled by g++ into the lane-by-lane
// emulation harness under tests/emu/ (CPU CI for the kernel logic; it is never part of the
// shipped library, and libwmix_b200.so has no CPU execution path).
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define WMX_HD __host__ __device__ __forceinline__
#define WMX_D __device__ __forceinline__
#else
#define WMX_HD inline
#define WMX_D inline
#endif

namespace wmx {

// ---- integer primitives with the reference's semantics (T:.../signal_processing/include/spl_inl.h) ----
WMX_HD int clz32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
// WebRtcSpl_NormW32 (spl_inl.h:101-121): redundant sign bits; 0 for 0.
WMX_HD int norm_w32(int32_t a)
{
    if (a == 0) return 0;
    if (a < 0) a = ~a;
    return clz32((uint32_t)a) - 1;
}
// WebRtcSpl_NormU32 (spl_inl.h:123-139)
WMX_HD int norm_u32(uint32_t a) { return a ? clz32(a) : 0; }
// WebRtcSpl_GetSizeInBits (spl_inl.h:84-99)
WMX_HD int size_in_bits(uint32_t n) { return 32 - clz32(n); }
// WebRtcSpl_SatW32ToW16 (spl_inl.h:24-33)
WMX_HD int16_t sat16(int32_t v) { return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v)); }
// C's truncating int32 / int32.  There is no divide instruction: nvcc expands every `/` into ~55 instructions, and the
// VAD / AGC bodies have 47 division sites, a quarter of post_kernel's 170 KB of code — which is bound by instruction
// fetch as much as by anything (no_instruction stalls).  One out-of-line copy keeps that code out of the instruction
// cache's way; the quotient is the same.
#if defined(__CUDACC__)
static __device__ __noinline__ int32_t sdiv32_device(int32_t num, int32_t den) { return num / den; }
#endif
WMX_HD int32_t sdiv32(int32_t num, int32_t den)
{
#if defined(__CUDA_ARCH__)
    return sdiv32_device(num, den);
#else
    return num / den;
#endif
}
// WebRtcSpl_DivW32W16 (division_operations.c:37-46)
WMX_HD int32_t div_w32_w16(int32_t num, int16_t den) { return den ? sdiv32(num, den) : (int32_t)0x7FFFFFFF; }
// WebRtcSpl_DivW32W16ResW16 (division_operations.c:48-57)
WMX_HD int16_t div_w32_w16_res16(int32_t num, int16_t den) { return den ? (int16_t)sdiv32(num, den) : (int16_t)0x7FFF; }
// two's-complement helpers: the reference relies on wrap-around where C calls it undefined
WMX_HD int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
WMX_HD int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
WMX_HD int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
WMX_HD int32_t wshl(int32_t a, int s) { return (int32_t)((uint32_t)a << s); }
// WEBRTC_SPL_SHIFT_W32
WMX_HD int32_t shift_w32(int32_t x, int c) { return c >= 0 ? wshl(x, c) : (x >> (-c)); }

// Bulk L2 prefetch (bytes a multiple of 16, address 16-byte aligned): one instruction, no destination, nothing to
// wait for.  A no-op in the host emulation build.
WMX_HD void l2_prefetch_bulk(const void* p, uint32_t bytes)
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#else
    (void)p;
    (void)bytes;
#endif
}

// packed int16 pair <-> 32-bit state word (lo = first field, hi = second field)
WMX_HD int16_t lo16(int32_t w) { return (int16_t)(w & 0xFFFF); }
WMX_HD int16_t hi16(int32_t w) { return (int16_t)(w >> 16); }
WMX_HD int32_t pack16(int16_t lo, int16_t hi) { return (int32_t)(((uint32_t)(uint16_t)hi << 16) | (uint16_t)lo); }

// Structure-of-arrays view of one stream's 32-bit state words: word w of stream s lives at
// base[w * stride + s], so a warp of consecutive streams touches one 128-byte line per word.
struct SoaWords {
    int32_t* p;      // already offset to this stream
    size_t stride;   // streams per word row
    WMX_HD int32_t get(int w) const { return p[(size_t)w * stride]; }
    WMX_HD void set(int w, int32_t v) const { p[(size_t)w * stride] = v; }
    // hint: words [w, w + n) will be read soon (a warp's 32 consecutive streams share one line per word); no-op on the host
    WMX_HD void prefetch(int w, int n) const
    {
#if defined(__CUDA_ARCH__)
        for (int k = 0; k < n; ++k) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + (size_t)(w + k) * stride));
#else
        (void)w;
        (void)n;
#endif
    }
};

}  // namespace wmx
