// Legacy WebRTC digital AGC (adaptive-digital mode, limiter off, one band), one stream per
// thread, state in structure-of-arrays 32-bit words.
//
// Follows T:webrtc/modules/audio_processing/agc/legacy/digital_agc.c:294-604 (ProcessDigital)
// and :633-771 (ProcessVad) bit for bit for the configuration wmix uses
// (R:src/webrtc.c:694-819): far-end VAD never fed, lowLevelSignal == 0, and
// WebRtcAgc_ProcessAnalog only moving mic-volume bookkeeping that wmix discards
// (SURVEY.md §8 a10), so neither is carried here.
#pragma once
#include "common.cuh"

namespace wmx {
namespace agc {

enum {
    W_CAP_SLOW = 0, W_CAP_FAST = 1, W_GAIN = 2, W_GATE_PREV = 3,
    W_DOWN = 4,        // 8 words : 8k->4k decimator states
    W_HP_CNT = 12,     // (HPstate, counter)
    W_LR_ML = 13,      // (logRatio, meanLongTerm)
    W_VAR_LONG = 14,
    W_SL_MS = 15,      // (stdLongTerm, meanShortTerm)
    W_VAR_SHORT = 16,
    W_STD_SHORT = 17,  // (stdShortTerm, -)
    N_WORDS = 18
};

// WebRtcSpl_Sqrt (T:.../signal_processing/spl_sqrt.c:24-184): normalise, 5-term series for
// sqrt(1+x) in Q31, optional 1/sqrt(2), de-normalise.  Negative arguments occur (variance
// estimates) and must wrap exactly like the C code.
WMX_HD int32_t spl_sqrt(int32_t value)
{
    int32_t a = value;
    if (a == 0) return 0;
    int sh = norm_w32(a);
    a = wshl(a, sh);
    a = (a < (0x7fffffff - 32767)) ? a + 32768 : 0x7fffffff;
    int16_t xn = (int16_t)(a >> 16);
    int nshift = sh / 2;
    a = wshl((int32_t)xn, 16);
    if (a < 0) a = wsub(0, a);
    {   // series
        int32_t b = a / 2;
        b = wsub(b, 0x40000000);
        int16_t xh = (int16_t)(b >> 16);
        b = wadd(b, 0x40000000);
        b = wadd(b, 0x40000000);
        int32_t x2 = wmul(wmul((int32_t)xh, (int32_t)xh), 2);
        int32_t c = wsub(0, x2);
        b = wadd(b, c >> 1);
        c >>= 16;
        c = wmul(wmul(c, c), 2);
        int16_t t = (int16_t)(c >> 16);
        b = wadd(b, wmul(-20480 * t, 2));
        c = wmul((int32_t)xh * t, 2);
        t = (int16_t)(c >> 16);
        b = wadd(b, wmul(28672 * t, 2));
        t = (int16_t)(x2 >> 16);
        c = wmul((int32_t)xh * t, 2);
        b = wadd(b, c >> 1);
        a = wadd(b, 32768);
    }
    if (2 * nshift == sh) {
        int16_t t = (int16_t)(a >> 16);
        a = wmul(23170 * t, 2);
        a = wadd(a, 32768);
        a &= 0x7fff0000;
        a >>= 15;
    } else {
        a >>= 16;
    }
    a &= 0xffff;
    return a >> nshift;
}

// c + ((diff * coef) >> 16) in two halves, 32-bit wrap (WEBRTC_SPL_SCALEDIFF32)
WMX_HD int32_t scalediff_u16(uint32_t coef, int32_t diff, int32_t c)
{
    uint32_t hi = (uint32_t)((diff >> 16) * (int32_t)coef);
    uint32_t lo = ((uint32_t)(diff & 0xFFFF) * coef) >> 16;
    return (int32_t)((uint32_t)c + hi + lo);
}
// AGC_SCALEDIFF32 / AGC_MUL32 (digital_agc.h:21-23): signed low half
WMX_HD int32_t agc_scalediff(int32_t a, int32_t b, int32_t c)
{
    return (int32_t)((uint32_t)c + (uint32_t)((b >> 16) * a) + (uint32_t)(((0xFFFF & b) * a) >> 16));
}
WMX_HD int32_t agc_mul32(int32_t a, int32_t b)
{
    return (int32_t)((uint32_t)wmul(b >> 13, a) + (uint32_t)(wmul(0x1FFF & b, a) >> 13));
}

// one output sample of WebRtcSpl_DownsampleBy2 from an input pair
// (T:.../signal_processing/resample_by_2.c:70-121)
WMX_HD int16_t down2_pair(int x_even, int x_odd, int32_t st[8])
{
    int32_t x = wshl(x_even, 10);
    int32_t d = wsub(x, st[1]);
    int32_t t1 = scalediff_u16(12199, d, st[0]);
    st[0] = x;
    d = wsub(t1, st[2]);
    int32_t t2 = scalediff_u16(37471, d, st[1]);
    st[1] = t1;
    d = wsub(t2, st[3]);
    st[3] = scalediff_u16(60255, d, st[2]);
    st[2] = t2;
    x = wshl(x_odd, 10);
    d = wsub(x, st[5]);
    t1 = scalediff_u16(3284, d, st[4]);
    st[4] = x;
    d = wsub(t1, st[6]);
    t2 = scalediff_u16(24441, d, st[5]);
    st[5] = t1;
    d = wsub(t2, st[7]);
    st[7] = scalediff_u16(49528, d, st[6]);
    st[6] = t2;
    return sat16(wadd(wadd(st[3], st[7]), 1024) >> 11);
}

// Activity estimate on a 4 kHz version of the frame (digital_agc.c:633-771).  Returns logRatio;
// std_long / std_short are handed back for the caller's decay and gate logic.
template <bool FS16>
WMX_HD int16_t process_vad(const SoaWords& st, const int16_t* in, int16_t& std_long, int16_t& std_short)
{
    int32_t ds[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) ds[k] = st.get(W_DOWN + k);
    int32_t w = st.get(W_HP_CNT);
    int16_t hp = lo16(w), counter = hi16(w);
    int32_t nrg = 0;
    for (int sub = 0; sub < 10; ++sub) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int xe, xo;
            if (FS16) {
                const int16_t* p = in + sub * 16 + 4 * k;
                xe = (int16_t)(((int32_t)p[0] + (int32_t)p[1]) >> 1);
                xo = (int16_t)(((int32_t)p[2] + (int32_t)p[3]) >> 1);
            } else {
                xe = in[sub * 8 + 2 * k];
                xo = in[sub * 8 + 2 * k + 1];
            }
            int16_t b = down2_pair(xe, xo, ds);
            int32_t o = b + hp;
            hp = (int16_t)(((600 * o) >> 10) - b);
            nrg = wadd(nrg, wmul(o, o) >> 6);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) st.set(W_DOWN + k, ds[k]);

    // leading zeros of nrg as the reference's open-coded ladder gives them (0 -> 31)
    int zeros = nrg ? clz32((uint32_t)nrg) : 31;
    int16_t dB = (int16_t)((15 - zeros) << 11);
    if (counter < 250) counter++;
    st.set(W_HP_CNT, pack16(hp, counter));

    int32_t w13 = st.get(W_LR_ML), w15 = st.get(W_SL_MS);
    int16_t log_ratio = lo16(w13), mean_long = hi16(w13), mean_short = hi16(w15);
    int32_t var_long = st.get(W_VAR_LONG), var_short = st.get(W_VAR_SHORT);

    int32_t t32 = mean_short * 15 + dB;
    mean_short = (int16_t)(t32 >> 4);
    t32 = (dB * dB) >> 12;
    t32 = wadd(t32, wmul(var_short, 15));
    var_short = t32 / 16;
    t32 = mean_short * mean_short;
    t32 = wsub(wshl(var_short, 12), t32);
    std_short = (int16_t)spl_sqrt(t32);

    int16_t cnt1 = sat16((int32_t)counter + 1);
    t32 = mean_long * counter + dB;
    mean_long = div_w32_w16_res16(t32, cnt1);
    t32 = (dB * dB) >> 12;
    t32 = wadd(t32, wmul(var_long, counter));
    var_long = div_w32_w16(t32, cnt1);
    t32 = mean_long * mean_long;
    t32 = wsub(wshl(var_long, 12), t32);
    std_long = (int16_t)spl_sqrt(t32);

    t32 = (3 << 12) * (int16_t)(dB - mean_long);
    t32 = div_w32_w16(t32, std_long);
    int32_t t32b = (int32_t)log_ratio * 53248;          // (uint16)(13 << 12)
    t32 = wadd(t32, t32b >> 10);
    log_ratio = (int16_t)(t32 >> 6);
    if (log_ratio > 2048) log_ratio = 2048;
    if (log_ratio < -2048) log_ratio = -2048;

    st.set(W_LR_ML, pack16(log_ratio, mean_long));
    st.set(W_VAR_LONG, var_long);
    st.set(W_SL_MS, pack16(std_long, mean_short));
    st.set(W_VAR_SHORT, var_short);
    st.set(W_STD_SHORT, pack16(std_short, 0));
    return log_ratio;
}

// One 10 ms packet of one stream, in place (digital_agc.c:294-604).  `table` is the 32-entry
// Q16 compressor curve (same for all streams of an engine).  L = samples per sub-frame.
template <bool FS16>
WMX_HD void process_packet(const SoaWords& st, int16_t* x, const int32_t* table)
{
    const int L = FS16 ? 16 : 8, L2 = FS16 ? 4 : 3;
    int16_t std_long, std_short;
    int16_t logratio = process_vad<FS16>(st, x, std_long, std_short);

    int16_t decay;
    if (logratio > 1024) decay = -65;
    else if (logratio < 0) decay = 0;
    else decay = (int16_t)(((0 - logratio) * 65) >> 10);
    if (std_long < 4000) decay = 0;
    else if (std_long < 8096) decay = (int16_t)(((std_long - 4000) * decay) >> 12);

    int32_t env[10], gains[11];
    for (int k = 0; k < 10; ++k) {
        int32_t peak = 0;
#pragma unroll 4
        for (int i = 0; i < L; ++i) {                 // (sample loops stay rolled: post_kernel is bound by instruction fetch)
            int32_t v = x[k * L + i];
            int32_t e = v * v;
            if (e > peak) peak = e;
        }
        env[k] = peak;
    }
    int32_t cap_fast = st.get(W_CAP_FAST), cap_slow = st.get(W_CAP_SLOW);
    const int32_t tab0 = table[0];
    gains[0] = st.get(W_GAIN);
    int16_t zeros = 0, frac = 0;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        cap_fast = agc_scalediff(-1000, cap_fast, cap_fast);
        if (env[k] > cap_fast) cap_fast = env[k];
        if (env[k] > cap_slow) cap_slow = agc_scalediff(500, wsub(env[k], cap_slow), cap_slow);
        else cap_slow = agc_scalediff(decay, cap_slow, cap_slow);
        int32_t cur = (cap_fast > cap_slow) ? cap_fast : cap_slow;
        zeros = (int16_t)norm_u32((uint32_t)cur);
        if (cur == 0) zeros = 31;
        int32_t t32 = (int32_t)(((uint32_t)cur << zeros) & 0x7FFFFFFF);
        frac = (int16_t)(t32 >> 19);
        t32 = wmul(table[zeros - 1] - table[zeros], frac);
        gains[k + 1] = table[zeros] + (t32 >> 12);
    }
    st.set(W_CAP_FAST, cap_fast);
    st.set(W_CAP_SLOW, cap_slow);

    // gate (digital_agc.c:470-518)
    zeros = (int16_t)((zeros << 9) - (frac >> 3));
    int16_t zeros_fast = (int16_t)norm_u32((uint32_t)cap_fast);
    if (cap_fast == 0) zeros_fast = 31;
    int32_t t32 = (int32_t)(((uint32_t)cap_fast << zeros_fast) & 0x7FFFFFFF);
    zeros_fast = (int16_t)(zeros_fast << 9);
    zeros_fast = (int16_t)(zeros_fast - (int16_t)(t32 >> 22));
    int16_t gate = (int16_t)(1000 + zeros_fast - zeros - std_short);
    int16_t gate_prev = (int16_t)st.get(W_GATE_PREV);
    if (gate < 0) {
        gate_prev = 0;
    } else {
        gate = (int16_t)((gate + gate_prev * 7) >> 3);
        gate_prev = gate;
    }
    st.set(W_GATE_PREV, gate_prev);
    if (gate > 0) {
        int adj = (gate < 2500) ? ((2500 - gate) >> 5) : 0;
#pragma unroll
        for (int k = 1; k <= 10; ++k) {
            int32_t d = gains[k] - tab0;
            if (d > 8388608) { t32 = d >> 8; t32 = wmul(t32, 178 + adj); }
            else { t32 = wmul(d, 178 + adj); t32 >>= 8; }
            gains[k] = tab0 + t32;
        }
    }
    // limiter loop (digital_agc.c:520-547)
#pragma unroll
    for (int k = 1; k <= 10; ++k) {
        int z = 10;
        if (gains[k] > 47453132) z = 16 - norm_w32(gains[k]);
        int32_t g32 = (gains[k] >> z) + 1;
        g32 = wmul(g32, g32);
        const int32_t lim = shift_w32(32767, 2 * (1 - z + 10));
        while (agc_mul32((env[k - 1] >> 12) + 1, g32) > lim) {
            if (gains[k] > 8388607) gains[k] = (gains[k] / 256) * 253;
            else gains[k] = (gains[k] * 253) / 256;
            g32 = (gains[k] >> z) + 1;
            g32 = wmul(g32, g32);
        }
    }
#pragma unroll
    for (int k = 1; k < 10; ++k)
        if (gains[k] > gains[k + 1]) gains[k] = gains[k + 1];
    st.set(W_GAIN, gains[10]);

    // apply: first sub-frame with saturation test, the rest a plain per-sample ramp
    int32_t delta = wshl(gains[1] - gains[0], 4 - L2);
    int32_t g32 = wshl(gains[0], 4);
    for (int i = 0; i < L; ++i) {
        int32_t v = x[i];
        int32_t o = wmul(v, (g32 + 127) >> 7) >> 16;
        if (o > 4095) x[i] = 32767;
        else if (o < -4096) x[i] = -32768;
        else x[i] = (int16_t)(wmul(v, g32 >> 4) >> 16);
        g32 += delta;
    }
    for (int k = 1; k < 10; ++k) {
        delta = wshl(gains[k + 1] - gains[k], 4 - L2);
        g32 = wshl(gains[k], 4);
#pragma unroll 4
        for (int i = 0; i < L; ++i) {
            int32_t v = x[k * L + i];
            x[k * L + i] = (int16_t)(wmul(v, g32 >> 4) >> 16);
            g32 += delta;
        }
    }
}

}  // namespace agc
}  // namespace wmx
