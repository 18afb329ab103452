/* ORACLE (test infrastructure) — legacy digital AGC restated from
 * T:webrtc/modules/audio_processing/agc/legacy/{digital_agc.c,analog_agc.c} and the wmix
 * handle layer R:src/webrtc.c:694-857.
 *
 * What wmix actually exercises: mode kAgcModeAdaptiveDigital,
 * targetLevelDbfs 0, limiter off, compressionGaindB = `value`, one band, inMicLevel 0,
 * echo 0.  AddMic/VirtualMic/AddFarend are never called, so the far-end VAD stays at its
 * init state (counter 3 -> the far-end mix at digital_agc.c:335-339 never fires) and
 * lowLevelSignal stays 0.  ProcessAnalog only moves mic-volume bookkeeping that
 * R:src/webrtc.c:817 discards, so it is not restated (analog_agc.c:1191-1203).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

/* T:.../digital_agc.c:36-53 kGenFuncTable: the tabulated generating function is
 * round(256*log2(1+e^i)), i = 0..127 (checked entry by entry against the reference's literal
 * table in tests/test_oracle_pin.py via the gain-table comparison over every gain). */
static uint16_t orc_genfunc(int i)
{
    return (uint16_t)floor(256.0 * log2(1.0 + exp((double)i)) + 0.5);
}

static int16_t orc_div16(int32_t num, int16_t den) /* division_operations.c:48-57 */
{
    return den ? (int16_t)(num / den) : (int16_t)0x7FFF;
}

static int32_t orc_shift(int32_t x, int c) /* WEBRTC_SPL_SHIFT_W32 */
{
    return c >= 0 ? (int32_t)((uint32_t)x << c) : (x >> (-c));
}

/* T:.../analog_agc.c:424-448 (UpdateAgcThresholds, the part that feeds the gain table) */
int16_t orc_agc_analog_target(int16_t comp_db)
{
    int16_t t = (int16_t)(5 * comp_db + 5);
    t = orc_div16((int32_t)t, 11);
    t = (int16_t)(4 + t);
    if (t < 4)
        t = 4;
    return t;
}

/* T:.../digital_agc.c:57-257 — 32-entry Q16 compressor curve */
int orc_agc_gain_table(int32_t table[32], int16_t comp_db, int16_t target_dbfs,
                       uint8_t limiter, int16_t analog_target)
{
    const uint16_t kLog10 = 54426, kLog10_2 = 49321, kLogE_1 = 23637;
    const int16_t kRatio = 3;
    int16_t lim_offset = 0, lim_idx, lim_x, max_gain, zero_lvl, diff_gain, t16, i;
    int32_t t32, lim_lvl, den;
    uint16_t max_gain_q8;

    t32 = (comp_db - analog_target) * (kRatio - 1);
    t16 = (int16_t)(analog_target - target_dbfs);
    t16 = (int16_t)(t16 + orc_div16(t32 + (kRatio >> 1), kRatio));
    max_gain = (t16 > (analog_target - target_dbfs)) ? t16 : (int16_t)(analog_target - target_dbfs);
    t32 = max_gain * kRatio;
    zero_lvl = comp_db;
    zero_lvl = (int16_t)(zero_lvl - orc_div16(t32 + ((kRatio - 1) >> 1), kRatio - 1));
    if (comp_db <= analog_target && limiter)
        zero_lvl = (int16_t)(zero_lvl + (analog_target - comp_db + 1));
    (void)zero_lvl;

    t32 = comp_db * (kRatio - 1);
    diff_gain = orc_div16(t32 + (kRatio >> 1), kRatio);
    if (diff_gain < 0 || diff_gain >= 128)
        return -1;

    lim_x = (int16_t)(analog_target - lim_offset);
    lim_idx = (int16_t)(2 + orc_div16((int32_t)lim_x << 13, (int16_t)(kLog10_2 / 2)));
    t16 = orc_div16(lim_offset + (kRatio >> 1), kRatio);
    lim_lvl = target_dbfs + t16;

    max_gain_q8 = orc_genfunc(diff_gain);
    den = 20 * (int32_t)max_gain_q8;

    for (i = 0; i < 32; ++i) {
        int32_t in_lvl, num, y, frac32;
        uint32_t abs_lvl, a, b, log_apx;
        uint16_t ipart, fpart, step;
        int zeros, zscale;

        t16 = (int16_t)((kRatio - 1) * (i - 1));
        t32 = (int32_t)t16 * kLog10_2 + 1;
        in_lvl = orc_div_w32_w16(t32, kRatio);
        in_lvl = ((int32_t)diff_gain << 14) - in_lvl;
        abs_lvl = (uint32_t)(in_lvl >= 0 ? in_lvl : -in_lvl);

        ipart = (uint16_t)(abs_lvl >> 14);
        fpart = (uint16_t)(abs_lvl & 0x3FFF);
        step = (uint16_t)(orc_genfunc(ipart + 1) - orc_genfunc(ipart));
        a = (uint32_t)step * fpart;
        a += (uint32_t)orc_genfunc(ipart) << 14;
        log_apx = a >> 8;
        if (in_lvl < 0) {
            zeros = orc_norm_u32(abs_lvl);
            zscale = 0;
            if (zeros < 15) {
                b = abs_lvl >> (15 - zeros);
                b = b * kLogE_1;
                if (zeros < 9) {
                    zscale = 9 - zeros;
                    a >>= zscale;
                } else {
                    b >>= zeros - 9;
                }
            } else {
                b = abs_lvl * kLogE_1;
                b >>= 6;
            }
            log_apx = 0;
            if (b < a)
                log_apx = (a - b) >> (8 - zscale);
        }
        num = (int32_t)((uint32_t)(max_gain * (int32_t)max_gain_q8) << 6);
        num -= (int32_t)log_apx * diff_gain;
        if (num > (den >> 8))
            zeros = orc_norm_w32(num);
        else
            zeros = orc_norm_w32(den) + 8;
        num = (int32_t)((uint32_t)num << zeros);
        t32 = orc_shift(den, zeros - 8);
        if (num < 0)
            num -= t32 / 2;
        else
            num += t32 / 2;
        y = num / t32;
        if (limiter && i < lim_idx) {
            t32 = (int32_t)(int16_t)(i - 1) * kLog10_2;
            t32 -= (int32_t)((uint32_t)lim_lvl << 14);
            y = orc_div_w32_w16(t32 + 10, 20);
        }
        if (y > 39000) {
            t32 = (y >> 1) * kLog10 + 4096;
            t32 >>= 13;
        } else {
            t32 = y * kLog10 + 8192;
            t32 >>= 14;
        }
        t32 += 16 << 14;
        if (t32 > 0) {
            ipart = (uint16_t)(int16_t)(t32 >> 14);
            fpart = (uint16_t)(t32 & 0x3FFF);
            if ((fpart >> 13) != 0) {
                t16 = (int16_t)((2 << 14) - 22817);
                frac32 = (1 << 14) - fpart;
                frac32 *= t16;
                frac32 >>= 13;
                frac32 = (1 << 14) - frac32;
            } else {
                t16 = (int16_t)(22817 - (1 << 14));
                frac32 = (fpart * t16) >> 13;
            }
            fpart = (uint16_t)frac32;
            table[i] = (int32_t)(1u << ipart) + orc_shift((int32_t)fpart, (int)ipart - 14);
        } else {
            table[i] = 0;
        }
    }
    return 0;
}

static void orc_agc_vad_init(orc_agc_vad *s) /* T:.../digital_agc.c:606-631 */
{
    memset(s, 0, sizeof(*s));
    s->mean_long = 15 << 10;
    s->var_long = 500 << 8;
    s->mean_short = 15 << 10;
    s->var_short = 500 << 8;
    s->counter = 3;
}

/* T:.../digital_agc.c:633-771 — level-based activity estimate on a 4 kHz version of the frame */
int16_t orc_agc_process_vad(orc_agc_vad *s, const int16_t *in, int n)
{
    int32_t nrg = 0, t32, t32b, o;
    int16_t hp = s->hp, zeros, dB, sub, k, t16;
    int16_t b1[8], b2[4];

    for (sub = 0; sub < 10; ++sub) {
        if (n == 160) {
            for (k = 0; k < 8; ++k)
                b1[k] = (int16_t)(((int32_t)in[2 * k] + (int32_t)in[2 * k + 1]) >> 1);
            in += 16;
            orc_downsample_by2(b1, 8, b2, s->down);
        } else {
            orc_downsample_by2(in, 8, b2, s->down);
            in += 8;
        }
        for (k = 0; k < 4; ++k) {
            o = b2[k] + hp;
            t32 = 600 * o;
            hp = (int16_t)((t32 >> 10) - b2[k]);
            nrg = (int32_t)((uint32_t)nrg + (uint32_t)((o * o) >> 6));
        }
    }
    s->hp = hp;

    /* leading zeros of nrg seen as uint32, with 0 -> 31 exactly as the open-coded ladder */
    {
        uint32_t u = (uint32_t)nrg;
        zeros = (u & 0xFFFF0000u) ? 0 : 16;
        if (!((u << zeros) & 0xFF000000u)) zeros += 8;
        if (!((u << zeros) & 0xF0000000u)) zeros += 4;
        if (!((u << zeros) & 0xC0000000u)) zeros += 2;
        if (!((u << zeros) & 0x80000000u)) zeros += 1;
    }
    dB = (int16_t)((15 - zeros) << 11);

    if (s->counter < 250)
        s->counter++;

    t32 = s->mean_short * 15 + dB;
    s->mean_short = (int16_t)(t32 >> 4);
    t32 = (dB * dB) >> 12;
    t32 += s->var_short * 15;
    s->var_short = t32 / 16;
    t32 = s->mean_short * s->mean_short;
    t32 = (int32_t)((uint32_t)s->var_short << 12) - t32;
    s->std_short = (int16_t)orc_sqrt(t32);

    t32 = s->mean_long * s->counter + dB;
    s->mean_long = orc_div16(t32, orc_sat16((int32_t)s->counter + 1));
    t32 = (dB * dB) >> 12;
    t32 += s->var_long * s->counter;
    s->var_long = orc_div_w32_w16(t32, orc_sat16((int32_t)s->counter + 1));
    t32 = s->mean_long * s->mean_long;
    t32 = (int32_t)((uint32_t)s->var_long << 12) - t32;
    s->std_long = (int16_t)orc_sqrt(t32);

    t16 = 3 << 12;
    t32 = t16 * (int16_t)(dB - s->mean_long);
    t32 = orc_div_w32_w16(t32, s->std_long);
    t32b = (int32_t)s->log_ratio * (uint16_t)(13 << 12);
    t32 = (int32_t)((uint32_t)t32 + (uint32_t)(t32b >> 10));
    s->log_ratio = (int16_t)(t32 >> 6);
    if (s->log_ratio > 2048)
        s->log_ratio = 2048;
    if (s->log_ratio < -2048)
        s->log_ratio = -2048;
    return s->log_ratio;
}

static int32_t orc_scale32(int32_t a, int32_t b, int32_t c) /* AGC_SCALEDIFF32, digital_agc.h:23 */
{
    return (int32_t)((uint32_t)c + (uint32_t)((b >> 16) * a) + (uint32_t)(((0xFFFF & b) * a) >> 16));
}

static int32_t orc_mul32(int32_t a, int32_t b) /* AGC_MUL32, digital_agc.h:21 */
{
    return (int32_t)((uint32_t)((b >> 13) * a) + (uint32_t)(((0x1FFF & b) * a) >> 13));
}

/* T:.../analog_agc.c:1361-1533 (Init, digital part) + T:.../digital_agc.c:259-281 */
void orc_agc_core_init(orc_agc_core *a, int fs, int comp_db)
{
    memset(a, 0, sizeof(*a));
    a->fs = fs;
    a->cap_slow = 134217728;
    a->cap_fast = 0;
    a->gain = 65536;
    a->gate_prev = 0;
    orc_agc_vad_init(&a->near_vad);
    a->target_dbfs = 0;
    a->limiter = 0;
    orc_agc_core_set_gain(a, comp_db);
}

/* T:.../analog_agc.c:1231-1287 (set_config): thresholds, then the table */
int orc_agc_core_set_gain(orc_agc_core *a, int comp_db)
{
    a->comp_db = (int16_t)comp_db;
    a->analog_target = orc_agc_analog_target(a->comp_db);
    return orc_agc_gain_table(a->table, a->comp_db, a->target_dbfs, a->limiter, a->analog_target);
}

/* T:.../analog_agc.c:1134-1229 (sample-count validation) + T:.../digital_agc.c:294-604 */
int orc_agc_core_process(orc_agc_core *a, const int16_t *in, int16_t *out, int n)
{
    int32_t gains[11], env[10], t32, cur = 0, gain32, delta;
    int16_t logratio, decay, zeros = 0, zeros_fast, frac = 0, gate, adj, k, i;
    int L, L2;

    if (a->fs == 8000) {
        if (n != 80)
            return -1;
        L = 8;
        L2 = 3;
    } else if (a->fs == 16000 || a->fs == 32000 || a->fs == 48000) {
        if (n != 160)
            return -1;
        L = 16;
        L2 = 4;
    } else {
        return -1;
    }
    if (in != out)
        memcpy(out, in, (size_t)(10 * L) * sizeof(int16_t));

    logratio = orc_agc_process_vad(&a->near_vad, out, (int16_t)(L * 10));
    /* far-end VAD counter stays at 3 (< 10) in wmix: no blending */

    if (logratio > 1024)
        decay = -65;
    else if (logratio < 0)
        decay = 0;
    else
        decay = (int16_t)(((0 - logratio) * 65) >> 10);
    /* adaptive-digital mode */
    if (a->near_vad.std_long < 4000)
        decay = 0;
    else if (a->near_vad.std_long < 8096)
        decay = (int16_t)(((a->near_vad.std_long - 4000) * decay) >> 12);

    for (k = 0; k < 10; ++k) {
        int32_t peak = 0;
        for (i = 0; i < L; ++i) {
            int32_t e = out[k * L + i] * out[k * L + i];
            if (e > peak)
                peak = e;
        }
        env[k] = peak;
    }

    gains[0] = a->gain;
    for (k = 0; k < 10; ++k) {
        a->cap_fast = orc_scale32(-1000, a->cap_fast, a->cap_fast);
        if (env[k] > a->cap_fast)
            a->cap_fast = env[k];
        if (env[k] > a->cap_slow)
            a->cap_slow = orc_scale32(500, (int32_t)((uint32_t)env[k] - (uint32_t)a->cap_slow), a->cap_slow);
        else
            a->cap_slow = orc_scale32(decay, a->cap_slow, a->cap_slow);
        cur = (a->cap_fast > a->cap_slow) ? a->cap_fast : a->cap_slow;
        zeros = orc_norm_u32((uint32_t)cur);
        if (cur == 0)
            zeros = 31;
        t32 = (int32_t)(((uint32_t)cur << zeros) & 0x7FFFFFFF);
        frac = (int16_t)(t32 >> 19);
        t32 = (a->table[zeros - 1] - a->table[zeros]) * frac;
        gains[k + 1] = a->table[zeros] + (t32 >> 12);
    }

    zeros = (int16_t)((zeros << 9) - (frac >> 3));
    zeros_fast = orc_norm_u32((uint32_t)a->cap_fast);
    if (a->cap_fast == 0)
        zeros_fast = 31;
    t32 = (int32_t)(((uint32_t)a->cap_fast << zeros_fast) & 0x7FFFFFFF);
    zeros_fast = (int16_t)(zeros_fast << 9);
    zeros_fast = (int16_t)(zeros_fast - (int16_t)(t32 >> 22));

    gate = (int16_t)(1000 + zeros_fast - zeros - a->near_vad.std_short);
    if (gate < 0) {
        a->gate_prev = 0;
    } else {
        t32 = a->gate_prev * 7;
        gate = (int16_t)((gate + t32) >> 3);
        a->gate_prev = gate;
    }
    if (gate > 0) {
        adj = (gate < 2500) ? (int16_t)((2500 - gate) >> 5) : 0;
        for (k = 0; k < 10; ++k) {
            if ((gains[k + 1] - a->table[0]) > 8388608) {
                t32 = (gains[k + 1] - a->table[0]) >> 8;
                t32 *= 178 + adj;
            } else {
                t32 = (gains[k + 1] - a->table[0]) * (178 + adj);
                t32 >>= 8;
            }
            gains[k + 1] = a->table[0] + t32;
        }
    }

    /* limiter: back the gain off until peak*gain^2 fits */
    for (k = 0; k < 10; ++k) {
        zeros = 10;
        if (gains[k + 1] > 47453132)
            zeros = (int16_t)(16 - orc_norm_w32(gains[k + 1]));
        gain32 = (gains[k + 1] >> zeros) + 1;
        gain32 = (int32_t)((uint32_t)gain32 * (uint32_t)gain32);
        while (orc_mul32((env[k] >> 12) + 1, gain32) > orc_shift((int32_t)32767, 2 * (1 - zeros + 10))) {
            if (gains[k + 1] > 8388607)
                gains[k + 1] = (gains[k + 1] / 256) * 253;
            else
                gains[k + 1] = (gains[k + 1] * 253) / 256;
            gain32 = (gains[k + 1] >> zeros) + 1;
            gain32 = (int32_t)((uint32_t)gain32 * (uint32_t)gain32);
        }
    }
    for (k = 1; k < 10; ++k)
        if (gains[k] > gains[k + 1])
            gains[k] = gains[k + 1];
    a->gain = gains[10];

    /* first sub-frame: apply with an explicit saturation test; the rest: plain ramp */
    delta = (int32_t)((uint32_t)(gains[1] - gains[0]) << (4 - L2));
    gain32 = (int32_t)((uint32_t)gains[0] << 4);
    for (i = 0; i < L; ++i) {
        int32_t o;
        t32 = out[i] * ((gain32 + 127) >> 7);
        o = t32 >> 16;
        if (o > 4095)
            out[i] = 32767;
        else if (o < -4096)
            out[i] = -32768;
        else {
            t32 = (int32_t)((uint32_t)(int32_t)out[i] * (uint32_t)(gain32 >> 4));
            out[i] = (int16_t)(t32 >> 16);
        }
        gain32 += delta;
    }
    for (k = 1; k < 10; ++k) {
        delta = (int32_t)((uint32_t)(gains[k + 1] - gains[k]) << (4 - L2));
        gain32 = (int32_t)((uint32_t)gains[k] << 4);
        for (i = 0; i < L; ++i) {
            t32 = (int32_t)((uint32_t)(int32_t)out[k * L + i] * (uint32_t)(gain32 >> 4));
            out[k * L + i] = (int16_t)(t32 >> 16);
            gain32 += delta;
        }
    }
    return 0;
}

/* ---- wmix handle layer: R:src/webrtc.c:694-857 ---- */

orc_agc *orc_agc_init(int chn, int freq, int interval_ms, int value)
{
    orc_agc *h;
    (void)interval_ms;
    if (freq > 32000 || freq % 8000 != 0)
        return NULL;
    h = (orc_agc *)calloc(1, sizeof(*h));
    orc_agc_core_init(&h->core, freq, (int16_t)value);
    h->chn = chn;
    h->freq = freq;
    h->interval_ms = (freq <= 16000) ? 10 : 5;
    h->pkg = freq / 1000 * h->interval_ms;
    return h;
}

int orc_agc_process(orc_agc *h, int16_t *in, int16_t *out, int frame_num)
{
    int16_t mono[160], res[160];
    int total = frame_num * h->chn, step = h->pkg * h->chn, pos, i, c, r;
    for (pos = 0; pos < total; pos += step) {
        for (i = 0; i < h->pkg; ++i) {
            int32_t s = 0;
            for (c = 0; c < h->chn; ++c)
                s += *in++;
            mono[i] = (int16_t)(s / h->chn);
        }
        r = orc_agc_core_process(&h->core, mono, res, h->pkg);
        if (r != 0)
            return r;
        for (i = 0; i < h->pkg; ++i)
            for (c = 0; c < h->chn; ++c)
                *out++ = res[i];
    }
    return 0;
}

void orc_agc_addition(orc_agc *h, uint8_t value) { orc_agc_core_set_gain(&h->core, (int16_t)value); }
void orc_agc_release(orc_agc *h) { free(h); }
