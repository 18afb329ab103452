/* ORACLE (test infrastructure) — the handful of WebRTC signal-processing primitives the
 * VAD / AGC path needs, restated from T:webrtc/common_audio/signal_processing.
 * All arithmetic is two's-complement with wrap-around on narrowing and arithmetic right
 * shift of negatives, exactly what gcc/x86 gives the reference. */
#include "oracle.h"

static int orc_clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }

/* T:.../include/spl_inl.h:101-121 — leading sign bits minus one; 0 for a==0 */
int16_t orc_norm_w32(int32_t a)
{
    if (a == 0)
        return 0;
    if (a < 0)
        a = ~a;
    return (int16_t)(orc_clz32((uint32_t)a) - 1);
}

/* T:.../include/spl_inl.h:123-139 */
int16_t orc_norm_u32(uint32_t a) { return (int16_t)(a ? orc_clz32(a) : 0); }

/* T:.../include/spl_inl.h:84-99 */
int16_t orc_size_in_bits(uint32_t n) { return (int16_t)(32 - orc_clz32(n)); }

/* T:.../include/spl_inl.h:24-33 */
int16_t orc_sat16(int32_t v) { return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v)); }

/* T:.../division_operations.c:37-46 — C division, and 0x7FFFFFFF on divide-by-zero */
int32_t orc_div_w32_w16(int32_t num, int16_t den)
{
    return den ? (int32_t)(num / den) : (int32_t)0x7FFFFFFF;
}

/* T:.../get_scaling_square.c:20-49 + T:.../energy.c:18-35.
 * |x| is taken in int16, so -32768 stays -32768 and never wins the max. */
int32_t orc_energy(const int16_t *v, int n, int *scale)
{
    int16_t nbits = orc_size_in_bits((uint32_t)n);
    int16_t peak = -1, t;
    int i, sh;
    int32_t e = 0;
    for (i = 0; i < n; ++i) {
        int16_t a = (int16_t)(v[i] > 0 ? v[i] : -v[i]);
        if (a > peak)
            peak = a;
    }
    t = orc_norm_w32((int32_t)peak * peak);
    if (peak == 0)
        sh = 0;
    else
        sh = (t > nbits) ? 0 : nbits - t;
    for (i = 0; i < n; ++i)
        e += ((int32_t)v[i] * v[i]) >> sh;
    *scale = sh;
    return e;
}

/* T:.../spl_sqrt.c:24-66 (SqrtLocal): 5-term series for sqrt(1+x) in Q31 */
static int32_t orc_sqrt_series(int32_t in)
{
    int16_t xh, t;
    int32_t a, b, x2;
    b = in / 2;
    b = (int32_t)((uint32_t)b - 0x40000000u);
    xh = (int16_t)(b >> 16);
    b = (int32_t)((uint32_t)b + 0x40000000u);
    b = (int32_t)((uint32_t)b + 0x40000000u);
    x2 = (int32_t)((uint32_t)((int32_t)xh * xh) * 2u);
    a = (int32_t)(0u - (uint32_t)x2);
    b = (int32_t)((uint32_t)b + (uint32_t)(a >> 1));
    a >>= 16;
    a = (int32_t)((uint32_t)(a * a) * 2u);
    t = (int16_t)(a >> 16);
    b = (int32_t)((uint32_t)b + (uint32_t)(-20480 * t) * 2u);
    a = (int32_t)((uint32_t)((int32_t)xh * t) * 2u);
    t = (int16_t)(a >> 16);
    b = (int32_t)((uint32_t)b + (uint32_t)(28672 * t) * 2u);
    t = (int16_t)(x2 >> 16);
    a = (int32_t)((uint32_t)((int32_t)xh * t) * 2u);
    b = (int32_t)((uint32_t)b + (uint32_t)(a >> 1));
    b = (int32_t)((uint32_t)b + 32768u);
    return b;
}

/* T:.../spl_sqrt.c:71-184 */
int32_t orc_sqrt(int32_t value)
{
    int16_t xn, nshift, t, sh;
    int32_t a = value;
    if (a == 0)
        return 0;
    sh = orc_norm_w32(a);
    a = (int32_t)((uint32_t)a << sh);
    if (a < (0x7fffffff - 32767))
        a = a + 32768;
    else
        a = 0x7fffffff;
    xn = (int16_t)(a >> 16);
    nshift = (int16_t)(sh / 2);
    a = (int32_t)((uint32_t)(int32_t)xn << 16);
    if (a < 0)
        a = (int32_t)(0u - (uint32_t)a);
    a = orc_sqrt_series(a);
    if (2 * nshift == sh) {
        t = (int16_t)(a >> 16);
        a = (int32_t)((uint32_t)(23170 * t) * 2u);
        a = (int32_t)((uint32_t)a + 32768u);
        a &= 0x7fff0000;
        a >>= 15;
    } else {
        a >>= 16;
    }
    a &= 0x0000ffff;
    a >>= nshift;
    return a;
}

/* c + (diff * coef) >> 16 split in high/low halves, result wraps in 32 bits
 * (T:.../include/signal_processing_library.h:79-80 WEBRTC_SPL_SCALEDIFF32) */
static int32_t orc_scalediff(uint16_t coef, int32_t diff, int32_t c)
{
    uint32_t hi = (uint32_t)((diff >> 16) * (int32_t)coef);
    uint32_t lo = ((uint32_t)(diff & 0xFFFF) * coef) >> 16;
    return (int32_t)((uint32_t)c + hi + lo);
}

/* T:.../resample_by_2.c:70-121 — two 3-section all-pass branches, decimate by 2 */
void orc_downsample_by2(const int16_t *in, int len, int16_t *out, int32_t st[8])
{
    static const uint16_t ka[3] = {12199, 37471, 60255};   /* branch fed by even samples */
    static const uint16_t kb[3] = {3284, 24441, 49528};    /* branch fed by odd samples  */
    int i;
    for (i = 0; i < (len >> 1); ++i) {
        int32_t x, d, t1, t2, o;
        x = (int32_t)((uint32_t)(int32_t)in[2 * i] << 10);
        d = (int32_t)((uint32_t)x - (uint32_t)st[1]);
        t1 = orc_scalediff(ka[0], d, st[0]);
        st[0] = x;
        d = (int32_t)((uint32_t)t1 - (uint32_t)st[2]);
        t2 = orc_scalediff(ka[1], d, st[1]);
        st[1] = t1;
        d = (int32_t)((uint32_t)t2 - (uint32_t)st[3]);
        st[3] = orc_scalediff(ka[2], d, st[2]);
        st[2] = t2;

        x = (int32_t)((uint32_t)(int32_t)in[2 * i + 1] << 10);
        d = (int32_t)((uint32_t)x - (uint32_t)st[5]);
        t1 = orc_scalediff(kb[0], d, st[4]);
        st[4] = x;
        d = (int32_t)((uint32_t)t1 - (uint32_t)st[6]);
        t2 = orc_scalediff(kb[1], d, st[5]);
        st[5] = t1;
        d = (int32_t)((uint32_t)t2 - (uint32_t)st[7]);
        st[7] = orc_scalediff(kb[2], d, st[6]);
        st[6] = t2;

        o = (int32_t)((uint32_t)st[3] + (uint32_t)st[7] + 1024u) >> 11;
        out[i] = orc_sat16(o);
    }
}
