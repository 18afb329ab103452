/* ORACLE (test infrastructure) — WebRTC FIXED-POINT noise suppressor ("nsx") restated from
 * T:webrtc/modules/audio_processing/ns/nsx_core.c, nsx_core_c.c, noise_suppression_x.c, the SPL
 * complex FFT it runs on (T:webrtc/common_audio/signal_processing/{real_fft,complex_fft,
 * complex_bit_reverse}.c) and the wmix handle layer with its switch thrown
 * (R:src/webrtc.c:511-523: `#define MAKE_WEBRTC_NSX`, then ns_init / ns_process at :563-660).
 *
 * All arithmetic is integer.  Where the reference leans on what the C standard leaves open, this file
 * pins what the reference compiled for x86-64 by gcc does (oracle/_ref is that build):
 *   - signed overflow wraps (two's complement; the Makefile passes -fwrapv);
 *   - a 32-bit shift takes its count modulo 32 (`sh_l` / `sh_r` below) — reached by
 *     `magnEnergy >> (2*normData + stages - 1)` on near-silent frames (nsx_core.c:1140, :1716);
 *   - narrowing to int16_t / uint16_t truncates.
 * A negative index into the 17-entry sigmoid table (nsx_core_c.c:179-181, spectral-difference branch with
 * no normalising energy) reads outside the table in the reference; here it counts as "past the table".
 *
 * Tables: everything that has a closed form is computed at first use (and checked against the
 * reference's literals by tests/test_oracle_pin.py); only the 17-entry sigmoid table is literal.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define ANA_MAX 256
#define BINS_MAX 129
#define NHIST 1000
#define N_EST 3
#define STARTUP_SHORT 50
#define STARTUP_LONG 200
#define START_BAND 5

typedef struct {
    int16_t log_stage[9];        /* WebRtcNsx_kLogTable      nsx_core.c:28  */
    int16_t counter_div[201];    /* WebRtcNsx_kCounterDiv    nsx_core.c:32  */
    int16_t log_frac[256];       /* WebRtcNsx_kLogTableFrac  nsx_core.c:49  */
    int16_t win128[128];         /* kBlocks80w128x           nsx_core.c:74  */
    int16_t win256[256];         /* kBlocks160w256x          nsx_core.c:90  */
    int16_t factor1[257];        /* kFactor1Table            nsx_core.c:135 */
    int16_t factor2[3][257];     /* kFactor2Aggressiveness1..3  nsx_core.c:170-236 */
    int16_t sum_log[66];         /* kSumLogIndex             nsx_core.c:240 */
    int16_t sum_sq_log[66];      /* kSumSquareLogIndex       nsx_core.c:254 */
    int16_t log_index[129];      /* kLogIndex                nsx_core.c:268 */
    int16_t determinant[66];     /* kDeterminantEstMatrix    nsx_core.c:290 */
    int16_t sine[1024];          /* kSinTable1024            complex_fft_tables.h:17 */
    int ready;
} nsx_tables;
static nsx_tables g_t;

/* nsx_core_c.c:17 (kIndicatorTable): a tabulated sigmoid with no closed form that reproduces it */
static const int16_t k_sigmoid[17] = {0, 2017, 3809, 5227, 6258, 6963, 7424, 7718, 7901, 8014, 8084, 8126, 8152, 8168, 8177, 8183, 8187};

static int nearest(double x) { return (int)floor(x + 0.5); }

static void make_window(int16_t *w, int ana, int ramp)
{
    int i;
    for (i = 0; i < ana; ++i) {
        int k = i < ana - i ? i : ana - i;
        w[i] = (int16_t)(k >= ramp ? 16384 : nearest(16384.0 * sin(M_PI * k / (2.0 * ramp))));
    }
}

static void make_tables(void)
{
    int i, m;
    if (g_t.ready)
        return;
    for (i = 0; i < 9; ++i)
        g_t.log_stage[i] = (int16_t)nearest(i * log(2.0) * 256.0);
    for (i = 0; i < 201; ++i) {
        int v = nearest(32768.0 / (i + 1));
        g_t.counter_div[i] = (int16_t)(v > 32767 ? 32767 : v);
    }
    for (i = 0; i < 256; ++i)
        g_t.log_frac[i] = (int16_t)nearest(256.0 * log2(1.0 + i / 256.0));
    make_window(g_t.win128, 128, 48);
    make_window(g_t.win256, 256, 96);
    /* the gain-map tables are indexed by an ENERGY ratio in Q8, the float formulas in the comments of
     * nsx_core.c:125-169 take the amplitude gain: gain = sqrt(index / 256), blim = 0.5, truncated to Q13 */
    for (i = 0; i <= 256; ++i) {
        static const double bound[3] = {0.25, 0.125, 0.09};
        double g = sqrt(i / 256.0), f = 1.0;
        if (g > 0.5) {
            f = 1.0 + 1.3 * (g - 0.5);
            if (g * f > 1.0)
                f = 1.0 / g;
        }
        g_t.factor1[i] = (int16_t)(8192.0 * f);
        for (m = 0; m < 3; ++m) {
            double f2 = 1.0;
            if (g <= 0.5)
                f2 = 1.0 - 0.3 * (0.5 - (g <= bound[m] ? bound[m] : g));
            g_t.factor2[m][i] = (int16_t)(8192.0 * f2);
        }
    }
    for (i = 1; i <= 128; ++i)
        g_t.log_index[i] = (int16_t)nearest(4096.0 * log2((double)i));
    for (i = 1; i < 66; ++i) {
        double s = 0.0, q = 0.0;
        int j;
        for (j = i; j <= 128; ++j) {
            double l = log2((double)j);
            s += l;
            q += l * l;
        }
        g_t.sum_log[i] = (int16_t)nearest(32.0 * s);
        g_t.sum_sq_log[i] = (int16_t)nearest(4.0 * q);
        g_t.determinant[i] = (int16_t)nearest((129 - i) * q - s * s);
    }
    for (i = 0; i < 1024; ++i)
        g_t.sine[i] = (int16_t)(32767.0 * sin(2.0 * M_PI * i / 1024.0));   /* truncated, not rounded */
    g_t.ready = 1;
}

/* table dump for the pin test: every table above in declaration order, then the sigmoid */
int orc_nsx_tables(int16_t *out, int cap)
{
    const int n = (int)((offsetof(nsx_tables, sine) + sizeof g_t.sine) / sizeof(int16_t));   /* no padding before `ready` counted */
    make_tables();
    if (out && cap >= n + 17) {
        memcpy(out, &g_t, (size_t)n * sizeof(int16_t));
        memcpy(out + n, k_sigmoid, sizeof k_sigmoid);
    }
    return n + 17;
}

struct orc_nsx_core {
    int fs, block, ana, half, bins, stages, mode;
    const int16_t *window;
    int16_t ana_buf[ANA_MAX], syn_buf[ANA_MAX];
    uint16_t filter[BINS_MAX];                 /* noiseSupFilter, Q14 */
    uint16_t overdrive, floor_gain;            /* Q8, Q14 (denoiseBound) */
    const int16_t *factor2;
    int gain_map;
    int16_t lquant[N_EST * BINS_MAX], density[N_EST * BINS_MAX], counter[N_EST], quant[BINS_MAX];
    int32_t lrt_max, lrt_min;
    int32_t lrt_avg[BINS_MAX];
    int32_t feat_lrt, thr_lrt;
    int16_t w_lrt, w_flat, w_diff;
    uint32_t feat_diff, thr_diff, feat_flat, thr_flat;
    int32_t pause_avg[BINS_MAX];
    uint32_t magn_energy, sum_magn, cur_avg_energy, time_avg_energy, time_avg_energy_acc;
    uint32_t white_level;
    uint32_t init_magn[BINS_MAX];
    int32_t pink_num, pink_exp;
    int min_norm, zero_input;
    uint32_t noise_prev[BINS_MAX];
    uint16_t magn_prev[BINS_MAX];
    int16_t prior_nonspeech;                   /* Q14 */
    int frame_idx, model_window, model_count;
    int16_t hist_lrt[NHIST], hist_flat[NHIST], hist_diff[NHIST];
    int16_t hb_buf[ANA_MAX];                   /* dataBufHBFX[0] */
    int q_noise, q_noise_prev, q_magn_prev;
    int16_t re[ANA_MAX], im[ANA_MAX];
    int32_t energy_in;
    int scale_energy_in, norm;
    /* per-frame results the high band needs */
    uint16_t nonspeech[BINS_MAX];
};

/* x86 shift semantics for 32-bit operands: the count is taken modulo 32 */
static uint32_t sh_r(uint32_t x, int c) { return x >> (c & 31); }
static uint32_t sh_l(uint32_t x, int c) { return x << (c & 31); }
static int32_t sh_ra(int32_t x, int c) { return x >> (c & 31); }
/* WEBRTC_SPL_SHIFT_W32: left for c >= 0, arithmetic right otherwise */
static int32_t shift_w32(int32_t x, int c) { return c >= 0 ? (int32_t)sh_l((uint32_t)x, c) : sh_ra(x, -c); }
static int norm_w16(int16_t a)
{
    /* spl_inl.h:140-160 */
    int z;
    if (a == 0)
        return 0;
    if (a < 0)
        a = (int16_t)~a;
    z = __builtin_clz((uint32_t)(uint16_t)a | 1u) - 17;
    return a == 0 ? 15 : z;   /* ~(-1) = 0: the reference's bit tests all pass -> 8+4+2+1 */
}
static int32_t mul_round(int a, int b, int c) { return (a * b + (1 << (c - 1))) >> c; }   /* MUL_16_16_RSFT_WITH_ROUND */
/* log2 in Q8 of a non-zero 32-bit value: integer part from the leading zeros, fraction from the table
 * (the idiom at nsx_core.c:361-367, :1045-1051, :1291-1296) */
static int log2_q8(uint32_t v)
{
    const int z = orc_norm_u32(v);
    const int frac = (int)(((v << z) & 0x7FFFFFFFu) >> 23);
    return ((31 - z) << 8) + g_t.log_frac[frac];
}
/* WebRtcSpl_SqrtFloor (spl_sqrt_floor.c:50-76) takes an int32: 2^31 (both parts -32768) arrives negative and yields 0 */
static uint32_t sqrt_floor(uint32_t v)
{
    uint32_t r;
    if (v & 0x80000000u)
        return 0;
    r = (uint32_t)sqrt((double)v);
    while (r * r > v)
        --r;
    while ((r + 1) * (r + 1) <= v)
        ++r;
    return r;
}

/* ---- SPL complex FFT, 16-bit in place (complex_fft.c:27-145 forward mode 1, :147-301 inverse mode 1) ---- */
static void bit_reverse(int16_t *x, int stages)
{
    /* complex_bit_reverse.c:51-108: a permutation by bit-reversed index (the tables there list its swaps) */
    const int n = 1 << stages;
    int i, b;
    for (i = 0; i < n; ++i) {
        int r = 0;
        for (b = 0; b < stages; ++b)
            if (i & (1 << b))
                r |= 1 << (stages - 1 - b);
        if (r > i) {
            int16_t tr = x[2 * i], ti = x[2 * i + 1];
            x[2 * i] = x[2 * r];
            x[2 * i + 1] = x[2 * r + 1];
            x[2 * r] = tr;
            x[2 * r + 1] = ti;
        }
    }
}

static int max_abs16(const int16_t *v, int n)
{
    /* min_max_operations.c:36-57 */
    int i, m = 0;
    for (i = 0; i < n; ++i) {
        const int a = abs((int)v[i]);
        if (a > m)
            m = a;
    }
    return m > 32767 ? 32767 : m;
}

static int complex_fft(int16_t *x, int stages, int inverse)
{
    const int n = 1 << stages;
    int l = 1, k = 9, scale = 0;
    while (l < n) {
        const int step = l << 1;
        int shift = 1, round2 = 16384, m, i;
        if (inverse) {
            /* data-dependent scaling of every pass (complex_fft.c:170-187) */
            const int peak = max_abs16(x, 2 * n);
            shift = 0;
            round2 = 8192;
            if (peak > 13573) {
                ++shift;
                ++scale;
                round2 <<= 1;
            }
            if (peak > 27146) {
                ++shift;
                ++scale;
                round2 <<= 1;
            }
        }
        for (m = 0; m < l; ++m) {
            const int j = m << k;
            const int wr = g_t.sine[j + 256];
            const int wi = inverse ? g_t.sine[j] : -g_t.sine[j];
            for (i = m; i < n; i += step) {
                const int p = i + l;
                const int32_t tr = (wr * x[2 * p] - wi * x[2 * p + 1] + 1) >> 1;
                const int32_t ti = (wr * x[2 * p + 1] + wi * x[2 * p] + 1) >> 1;
                const int32_t qr = (int32_t)x[2 * i] * 16384, qi = (int32_t)x[2 * i + 1] * 16384;
                x[2 * p] = (int16_t)((qr - tr + round2) >> (shift + 14));
                x[2 * p + 1] = (int16_t)((qi - ti + round2) >> (shift + 14));
                x[2 * i] = (int16_t)((qr + tr + round2) >> (shift + 14));
                x[2 * i + 1] = (int16_t)((qi + ti + round2) >> (shift + 14));
            }
        }
        --k;
        l = step;
    }
    return scale;
}

/* real_fft.c:46-73: n real points -> bins 0..n/2 as (re, im) pairs */
static void real_forward(const int16_t *in, int16_t *out, int stages)
{
    int16_t buf[2 * ANA_MAX];
    const int n = 1 << stages;
    int i;
    for (i = 0; i < n; ++i) {
        buf[2 * i] = in[i];
        buf[2 * i + 1] = 0;
    }
    bit_reverse(buf, stages);
    complex_fft(buf, stages, 0);
    memcpy(out, buf, sizeof(int16_t) * (size_t)(n + 2));
}

/* real_fft.c:75-103: the upper half rebuilt by conjugate symmetry, real parts returned; result = scale shifts */
static int real_inverse(const int16_t *in, int16_t *out, int stages)
{
    int16_t buf[2 * ANA_MAX];
    const int n = 1 << stages;
    int i, scale;
    memcpy(buf, in, sizeof(int16_t) * (size_t)(n + 2));
    for (i = n + 2; i < 2 * n; i += 2) {
        buf[i] = in[2 * n - i];
        buf[i + 1] = (int16_t)-in[2 * n - i + 1];
    }
    bit_reverse(buf, stages);
    scale = complex_fft(buf, stages, 1);
    for (i = 0; i < n; ++i)
        out[i] = buf[2 * i];
    return scale;
}

/* ---- init: nsx_core.c:631-784 (InitCore) and :786-814 (set_policy_core) ---- */
static int nsx_set_policy(orc_nsx_core *s, int mode)
{
    static const uint16_t over[4] = {256, 256, 282, 320}, floor_q14[4] = {8192, 4096, 2048, 1475};
    if (mode < 0 || mode > 3)
        return -1;
    s->mode = mode;
    s->overdrive = over[mode];
    s->floor_gain = floor_q14[mode];
    s->gain_map = mode != 0;
    s->factor2 = mode ? g_t.factor2[mode - 1] : NULL;
    return 0;
}

static int nsx_core_init(orc_nsx_core *s, int fs)
{
    int i;
    make_tables();
    memset(s, 0, sizeof *s);
    if (fs != 8000 && fs != 16000 && fs != 32000 && fs != 48000)
        return -1;
    s->fs = fs;
    if (fs == 8000) {
        s->block = 80;
        s->ana = 128;
        s->stages = 7;
        s->window = g_t.win128;
        s->thr_lrt = 131072;
        s->lrt_max = 0x0040000;
        s->lrt_min = 52429;
    } else {
        s->block = 160;
        s->ana = 256;
        s->stages = 8;
        s->window = g_t.win256;
        s->thr_lrt = 212644;
        s->lrt_max = 0x0080000;
        s->lrt_min = 104858;
    }
    s->half = s->ana / 2;
    s->bins = s->half + 1;
    for (i = 0; i < N_EST * BINS_MAX; ++i) {
        s->lquant[i] = 2048;
        s->density[i] = 153;
    }
    for (i = 0; i < N_EST; ++i)
        s->counter[i] = (int16_t)((int16_t)(STARTUP_LONG * (i + 1)) / N_EST);
    for (i = 0; i < BINS_MAX; ++i)
        s->filter[i] = 16384;
    s->prior_nonspeech = 8192;
    s->thr_diff = 50;
    s->thr_flat = 20480;
    s->feat_lrt = s->thr_lrt;
    s->feat_flat = s->thr_flat;
    s->feat_diff = s->thr_diff;
    s->w_lrt = 6;
    s->frame_idx = -1;
    s->model_window = 1 << 9;
    s->min_norm = 15;
    return nsx_set_policy(s, 0);
}

/* ---- analysis: nsx_core.c:1184-1419 (DataAnalysis), :524-552 ---- */
static void nsx_analysis(orc_nsx_core *s, const int16_t *speech, uint16_t *magn)
{
    int16_t win[ANA_MAX], shifted[ANA_MAX], spec[ANA_MAX + 2];
    const int keep = s->ana - s->block;
    int i, peak, net_norm, drop_magn, drop_init;
    memmove(s->ana_buf, s->ana_buf + s->block, sizeof(int16_t) * (size_t)keep);
    memcpy(s->ana_buf + keep, speech, sizeof(int16_t) * (size_t)s->block);
    for (i = 0; i < s->ana; ++i)
        win[i] = (int16_t)mul_round(s->window[i], s->ana_buf[i], 14);
    s->energy_in = orc_energy(win, s->ana, &s->scale_energy_in);
    s->zero_input = 0;
    peak = max_abs16(win, s->ana);
    s->norm = norm_w16((int16_t)peak);
    if (peak == 0) {
        s->zero_input = 1;
        return;
    }
    net_norm = s->stages - s->norm;
    drop_magn = s->norm - s->min_norm;
    drop_init = -drop_magn > 0 ? -drop_magn : 0;
    s->min_norm -= drop_init;
    if (drop_magn < 0)
        drop_magn = 0;
    for (i = 0; i < s->ana; ++i)
        shifted[i] = (int16_t)(win[i] << s->norm);
    real_forward(shifted, spec, s->stages);

    s->im[0] = 0;
    s->im[s->half] = 0;
    s->re[0] = spec[0];
    s->re[s->half] = spec[s->ana];
    s->magn_energy = (uint32_t)(s->re[0] * s->re[0]);
    s->magn_energy += (uint32_t)(s->re[s->half] * s->re[s->half]);
    magn[0] = (uint16_t)abs((int)s->re[0]);
    magn[s->half] = (uint16_t)abs((int)s->re[s->half]);
    s->sum_magn = (uint32_t)magn[0] + (uint32_t)magn[s->half];

    if (s->frame_idx >= STARTUP_SHORT) {
        for (i = 1; i < s->half; ++i) {
            uint32_t e;
            s->re[i] = spec[2 * i];
            s->im[i] = (int16_t)-spec[2 * i + 1];
            e = (uint32_t)(spec[2 * i] * spec[2 * i]) + (uint32_t)(spec[2 * i + 1] * spec[2 * i + 1]);
            s->magn_energy += e;
            magn[i] = (uint16_t)sqrt_floor(e);
            s->sum_magn += magn[i];
        }
        return;
    }
    /* start-up: accumulate the average spectrum and fit white / pink noise models (:1274-1418) */
    {
        int32_t sum_log_magn, sum_log_i_log_magn, t1, t2;
        int16_t lg = 0, sum_log_i, sum_log_i_sq, det;
        uint16_t sum_log_magn_u16, tu16;
        uint32_t tu32;
        int zeros;
        s->init_magn[0] >>= drop_init;
        s->init_magn[s->half] >>= drop_init;
        s->init_magn[0] += (uint32_t)(magn[0] >> drop_magn);
        s->init_magn[s->half] += (uint32_t)(magn[s->half] >> drop_magn);
        if (magn[s->half])
            lg = (int16_t)log2_q8(magn[s->half]);
        sum_log_magn = lg;
        sum_log_i_log_magn = (g_t.log_index[s->half] * lg) >> 3;
        for (i = 1; i < s->half; ++i) {
            uint32_t e;
            s->re[i] = spec[2 * i];
            s->im[i] = (int16_t)-spec[2 * i + 1];
            e = (uint32_t)(spec[2 * i] * spec[2 * i]) + (uint32_t)(spec[2 * i + 1] * spec[2 * i + 1]);
            s->magn_energy += e;
            magn[i] = (uint16_t)sqrt_floor(e);
            s->sum_magn += magn[i];
            s->init_magn[i] >>= drop_init;
            s->init_magn[i] += (uint32_t)(magn[i] >> drop_magn);
            if (i >= START_BAND) {
                lg = magn[i] ? (int16_t)log2_q8(magn[i]) : 0;
                sum_log_magn += lg;
                sum_log_i_log_magn += (g_t.log_index[i] * lg) >> 3;
            }
        }
        s->white_level >>= drop_init;
        tu32 = s->sum_magn * (uint32_t)s->overdrive;
        tu32 >>= s->stages + 8;
        tu32 >>= drop_magn;
        s->white_level += tu32;

        det = g_t.determinant[START_BAND];
        sum_log_i = g_t.sum_log[START_BAND];
        sum_log_i_sq = g_t.sum_sq_log[START_BAND];
        if (s->fs == 8000) {
            t1 = det;
            t1 += (g_t.sum_log[65] * sum_log_i) >> 9;
            t1 -= (g_t.sum_log[65] * g_t.sum_log[65]) >> 10;
            t1 -= (int32_t)sum_log_i_sq << 4;
            t1 -= ((s->bins - START_BAND) * g_t.sum_sq_log[65]) >> 2;
            det = (int16_t)t1;
            sum_log_i = (int16_t)(sum_log_i - g_t.sum_log[65]);
            sum_log_i_sq = (int16_t)(sum_log_i_sq - g_t.sum_sq_log[65]);
        }
        zeros = 16 - orc_norm_w32(sum_log_magn);
        if (zeros < 0)
            zeros = 0;
        t1 = sum_log_magn << 1;
        sum_log_magn_u16 = (uint16_t)(t1 >> zeros);
        t2 = (int32_t)sum_log_i_sq * sum_log_magn_u16;
        tu32 = (uint32_t)(sum_log_i_log_magn >> 12);
        tu16 = (uint16_t)((uint16_t)sum_log_i << 1);
        if ((uint32_t)sum_log_i > tu32)
            tu16 = (uint16_t)(tu16 >> zeros);
        else
            tu32 >>= zeros;
        t2 -= (int32_t)(tu32 * (uint32_t)tu16);
        det = (int16_t)(det >> zeros);
        t2 = orc_div_w32_w16(t2, det);
        t2 += (int32_t)net_norm * 2048;
        if (t2 < 0)
            t2 = 0;
        s->pink_num += t2;
        t2 = (int32_t)sum_log_i * sum_log_magn_u16;
        t1 = sum_log_i_log_magn >> (3 + zeros);
        t1 *= s->bins - START_BAND;
        t2 -= t1;
        if (t2 > 0) {
            t1 = orc_div_w32_w16(t2, det);
            s->pink_exp += t1 > 16384 ? 16384 : (t1 < 0 ? 0 : t1);
        }
    }
}

/* ---- synthesis: nsx_core.c:1421-1499 (DataSynthesis), :456-521 ---- */
static void nsx_read_out(orc_nsx_core *s, int16_t *out)
{
    const int keep = s->ana - s->block;
    memcpy(out, s->syn_buf, sizeof(int16_t) * (size_t)s->block);
    memmove(s->syn_buf, s->syn_buf + s->block, sizeof(int16_t) * (size_t)keep);
    memset(s->syn_buf + keep, 0, sizeof(int16_t) * (size_t)s->block);
}

static void nsx_synthesis(orc_nsx_core *s, int16_t *out)
{
    int16_t freq[ANA_MAX + 2], time[ANA_MAX];
    int i, scale_ifft, gain = 8192;
    if (s->zero_input) {
        nsx_read_out(s, out);
        return;
    }
    for (i = 0; i < s->bins; ++i) {
        s->re[i] = (int16_t)((s->re[i] * (int16_t)s->filter[i]) >> 14);
        s->im[i] = (int16_t)((s->im[i] * (int16_t)s->filter[i]) >> 14);
    }
    for (i = 0; i <= s->half; ++i) {
        freq[2 * i] = s->re[i];
        freq[2 * i + 1] = (int16_t)-s->im[i];
    }
    scale_ifft = real_inverse(freq, time, s->stages);
    for (i = 0; i < s->ana; ++i)
        s->re[i] = orc_sat16(shift_w32((int32_t)time[i], scale_ifft - s->norm));
    if (s->gain_map == 1 && s->frame_idx > STARTUP_LONG && s->energy_in > 0) {
        int scale_out = 0;
        int32_t e_out = orc_energy(s->re, s->ana, &scale_out);
        int16_t ratio, g1, g2;
        if (scale_out == 0 && !(e_out & 0x7f800000))
            e_out = shift_w32(e_out, 8 + scale_out - s->scale_energy_in);
        else
            s->energy_in = sh_ra(s->energy_in, 8 + scale_out - s->scale_energy_in);
        /* the reference asserts energy_in > 0 here (:1478); a zero would be a division fault there */
        ratio = s->energy_in ? (int16_t)((e_out + s->energy_in / 2) / s->energy_in) : 256;
        ratio = (int16_t)(ratio > 256 ? 256 : (ratio < 0 ? 0 : ratio));
        g1 = g_t.factor1[ratio];
        g2 = s->factor2[ratio];
        gain = (int16_t)((int16_t)(((16384 - s->prior_nonspeech) * g1) >> 14) + (int16_t)((s->prior_nonspeech * g2) >> 14));
    }
    for (i = 0; i < s->ana; ++i) {
        const int16_t a = (int16_t)mul_round(s->window[i], s->re[i], 14);
        const int16_t b = orc_sat16(mul_round(a, (int16_t)gain, 13));
        s->syn_buf[i] = orc_sat16((int32_t)s->syn_buf[i] + b);
    }
    nsx_read_out(s, out);
}

/* ---- spectral flatness: nsx_core.c:1022-1084 ---- */
static void nsx_flatness(orc_nsx_core *s, const uint16_t *magn)
{
    uint32_t num = 0;
    const uint32_t den = s->sum_magn - (uint32_t)magn[0];
    int32_t lg, cur, t;
    int i, int_part;
    for (i = 1; i < s->bins; ++i) {
        if (!magn[i]) {
            s->feat_flat -= (s->feat_flat * 4915u) >> 14;
            return;
        }
        num += (uint32_t)log2_q8(magn[i]);
    }
    {
        const int z = orc_norm_u32(den);
        const int frac = (int)(((den << z) & 0x7FFFFFFFu) >> 23);
        t = ((31 - z) << 8) + g_t.log_frac[frac];
    }
    lg = (int32_t)num;
    lg += (int32_t)(s->stages - 1) << (s->stages + 7);
    lg -= t << (s->stages - 1);
    lg = (int32_t)((uint32_t)lg << (10 - s->stages));
    t = (int32_t)(0x00020000 | ((lg >= 0 ? lg : -lg) & 0x0001FFFF));
    int_part = (int16_t)(7 - (lg >> 17));
    cur = int_part > 0 ? sh_ra(t, int_part) : (int32_t)sh_l((uint32_t)t, -int_part);
    t = cur - (int32_t)s->feat_flat;
    t *= 4915;
    s->feat_flat += (uint32_t)(t >> 14);
}

/* ---- quantile noise estimate: nsx_core.c:303-453 ---- */
static void nsx_refresh_quantile(orc_nsx_core *s, int offset)
{
    int i, peak = -32768;
    for (i = 0; i < s->bins; ++i)
        if (s->lquant[offset + i] > peak)
            peak = s->lquant[offset + i];
    s->q_noise = 14 - mul_round(11819, peak, 21);
    for (i = 0; i < s->bins; ++i) {
        const int32_t p = 11819 * s->lquant[offset + i];
        int32_t v = 0x00200000 | (p & 0x001FFFFF);
        int16_t sh = (int16_t)(p >> 21);
        sh = (int16_t)(sh - 21);
        sh = (int16_t)(sh + (int16_t)s->q_noise);
        v = sh < 0 ? sh_ra(v, -sh) : (int32_t)sh_l((uint32_t)v, sh);
        s->quant[i] = orc_sat16(v);
    }
}

static void nsx_noise_estimate(orc_nsx_core *s, const uint16_t *magn, uint32_t *noise, int16_t *q_noise)
{
    int16_t lmagn[BINS_MAX];
    const int tab = s->stages - s->norm;
    const int16_t logval = (int16_t)(tab < 0 ? -g_t.log_stage[-tab] : g_t.log_stage[tab]);
    int i, e, offset = 0;
    for (i = 0; i < s->bins; ++i) {
        if (magn[i]) {
            const int16_t l2 = (int16_t)log2_q8(magn[i]);
            lmagn[i] = (int16_t)((l2 * 22713) >> 15);
            lmagn[i] = (int16_t)(lmagn[i] + logval);
        } else {
            lmagn[i] = logval;
        }
    }
    for (e = 0; e < N_EST; ++e) {
        const int16_t counter = s->counter[e];
        const int16_t cdiv = g_t.counter_div[counter];
        const int16_t cprod = (int16_t)(counter * cdiv);
        offset = e * s->bins;
        for (i = 0; i < s->bins; ++i) {
            int16_t delta, step;
            int16_t *lq = &s->lquant[offset + i], *dn = &s->density[offset + i];
            if (*dn > 512) {
                delta = (int16_t)(2621440 >> (14 - norm_w16(*dn)));
            } else {
                delta = s->frame_idx < STARTUP_LONG ? 1024 : 5120;
            }
            step = (int16_t)((delta * cdiv) >> 14);
            if (lmagn[i] > *lq) {
                step = (int16_t)(step + 2);
                *lq = (int16_t)(*lq + step / 4);
            } else {
                step = (int16_t)(step + 1);
                *lq = (int16_t)(*lq - (int16_t)((step / 2) * 3 / 2));
                if (*lq < logval)
                    *lq = logval;
            }
            if (abs(lmagn[i] - *lq) < 3) {
                const int16_t a = (int16_t)mul_round(*dn, cprod, 15);
                const int16_t b = (int16_t)mul_round(21845, cdiv, 15);
                *dn = (int16_t)(a + b);
            }
        }
        if (counter >= STARTUP_LONG) {
            s->counter[e] = 0;
            if (s->frame_idx >= STARTUP_LONG)
                nsx_refresh_quantile(s, offset);
        }
        s->counter[e]++;
    }
    if (s->frame_idx < STARTUP_LONG)
        nsx_refresh_quantile(s, offset);   /* offset of the LAST estimate, as in the reference (:445-447) */
    for (i = 0; i < s->bins; ++i)
        noise[i] = (uint32_t)s->quant[i];
    *q_noise = (int16_t)s->q_noise;
}

/* ---- spectral difference: nsx_core.c:1091-1181 ---- */
static void nsx_difference(orc_nsx_core *s, const uint16_t *magn)
{
    int32_t avg_pause = 0, max_pause = 0, min_pause = s->pause_avg[0], avg_magn, cov = 0, t;
    uint32_t var_magn = 0, var_pause = 0, diff, u1, u2;
    int i, shifts, norm32;
    for (i = 0; i < s->bins; ++i) {
        avg_pause += s->pause_avg[i];
        if (s->pause_avg[i] > max_pause)
            max_pause = s->pause_avg[i];
        if (s->pause_avg[i] < min_pause)
            min_pause = s->pause_avg[i];
    }
    avg_pause >>= s->stages - 1;
    avg_magn = (int32_t)(s->sum_magn >> (s->stages - 1));
    t = max_pause - avg_pause > avg_pause - min_pause ? max_pause - avg_pause : avg_pause - min_pause;
    shifts = 10 + s->stages - orc_norm_w32(t);
    if (shifts < 0)
        shifts = 0;
    for (i = 0; i < s->bins; ++i) {
        const int16_t dm = (int16_t)((int32_t)magn[i] - avg_magn);
        const int32_t dp = s->pause_avg[i] - avg_pause;
        int32_t q;
        var_magn += (uint32_t)(dm * dm);
        cov += dp * dm;
        q = dp >> shifts;
        var_pause += (uint32_t)(q * q);
    }
    s->cur_avg_energy += sh_r(s->magn_energy, 2 * s->norm + s->stages - 1);
    diff = var_magn;
    if (var_pause && cov) {
        u1 = (uint32_t)(cov >= 0 ? cov : -cov);
        norm32 = orc_norm_u32(u1) - 16;
        u1 = norm32 > 0 ? u1 << norm32 : u1 >> -norm32;
        u2 = u1 * u1;
        shifts += norm32;
        shifts <<= 1;
        if (shifts < 0) {
            var_pause = sh_r(var_pause, -shifts);
            shifts = 0;
        }
        if (var_pause > 0) {
            u1 = u2 / var_pause;
            u1 = sh_r(u1, shifts);
            diff -= diff < u1 ? diff : u1;
        } else {
            diff = 0;
        }
    }
    u1 = diff >> (2 * s->norm);
    if (s->feat_diff > u1)
        s->feat_diff -= ((s->feat_diff - u1) * 77u) >> 8;
    else
        s->feat_diff += ((u1 - s->feat_diff) * 77u) >> 8;
}

/* ---- feature histograms and threshold re-learning: nsx_core.c:821-1016 ---- */
static void two_peaks(const int16_t *hist, uint32_t *pos1, uint32_t *pos2, int *w1, int *w2)
{
    int i, max1 = 0, max2 = 0;
    *pos1 = *pos2 = 0;
    *w1 = *w2 = 0;
    for (i = 0; i < NHIST; ++i) {
        if (hist[i] > max1) {
            max2 = max1;
            *w2 = *w1;
            *pos2 = *pos1;
            max1 = hist[i];
            *w1 = hist[i];
            *pos1 = (uint32_t)(2 * i + 1);
        } else if (hist[i] > max2) {
            max2 = hist[i];
            *w2 = hist[i];
            *pos2 = (uint32_t)(2 * i + 1);
        }
    }
}

static void nsx_feature_update(orc_nsx_core *s, int relearn)
{
    if (!relearn) {
        uint32_t idx = (uint32_t)s->feat_lrt;
        if (idx < NHIST)
            s->hist_lrt[idx]++;
        idx = (s->feat_flat * 5) >> 8;
        if (idx < NHIST)
            s->hist_flat[idx]++;
        idx = NHIST;
        if (s->time_avg_energy > 0)
            idx = ((s->feat_diff * 5) >> s->stages) / s->time_avg_energy;
        if (idx < NHIST)
            s->hist_diff[idx]++;
        return;
    }
    {
        int32_t avg = 0, avg_sq = 0, avg_all, fluct, thr_fluct, t;
        int16_t count = 0;
        uint32_t u, p1, p2;
        int i, w1, w2, use_flat = 1, use_diff = 1, share;
        for (i = 0; i < 10; ++i) {
            const int16_t j = (int16_t)(2 * i + 1);
            t = s->hist_lrt[i] * j;
            avg += t;
            count = (int16_t)(count + s->hist_lrt[i]);
            avg_sq += t * j;
        }
        avg_all = avg;
        for (; i < NHIST; ++i) {
            const int16_t j = (int16_t)(2 * i + 1);
            t = s->hist_lrt[i] * j;
            avg_all += t;
            avg_sq += t * j;
        }
        fluct = avg_sq * count - avg * avg_all;
        thr_fluct = 10240 * count;
        u = 6u * (uint32_t)avg;
        if (fluct < thr_fluct || count == 0 || u > (uint32_t)(100 * count)) {
            s->thr_lrt = s->lrt_max;
        } else {
            t = (int32_t)((u << (9 + s->stages)) / (uint32_t)count / 25);
            s->thr_lrt = t > s->lrt_max ? s->lrt_max : (t < s->lrt_min ? s->lrt_min : t);
        }
        if (fluct < thr_fluct)
            use_diff = 0;
        two_peaks(s->hist_flat, &p1, &p2, &w1, &w2);
        if (p1 - p2 < 4 && w2 * 2 > w1) {
            w1 += w2;
            p1 = (p1 + p2) >> 1;
        }
        if (w1 < 154 || p1 < 24) {
            use_flat = 0;
        } else {
            u = 922u * p1;
            s->thr_flat = u > 38912u ? 38912u : (u < 4096u ? 4096u : u);
        }
        if (use_diff) {
            two_peaks(s->hist_diff, &p1, &p2, &w1, &w2);
            if (p1 - p2 < 4 && w2 * 2 > w1) {
                w1 += w2;
                p1 = (p1 + p2) >> 1;
            }
            u = 6u * p1;
            s->thr_diff = u > 100u ? 100u : (u < 16u ? 16u : u);
            if (w1 < 154)
                use_diff = 0;
        }
        share = 6 / (1 + use_flat + use_diff);
        s->w_lrt = (int16_t)share;
        s->w_flat = (int16_t)(use_flat * share);
        s->w_diff = (int16_t)(use_diff * share);
        memset(s->hist_lrt, 0, sizeof s->hist_lrt);
        memset(s->hist_diff, 0, sizeof s->hist_diff);
        memset(s->hist_flat, 0, sizeof s->hist_flat);
    }
}

/* ---- speech / noise probability: nsx_core_c.c:26-260 ---- */
static int16_t sigmoid_q14(uint32_t x_q14, int upper, int rounded)
{
    /* 8192 +/- interpolated table value; past the table the caller's 16384 / 0 stands */
    const int16_t idx = (int16_t)(x_q14 >> 14);
    int16_t v;
    if (idx < 0 || idx >= 16)
        return (int16_t)(upper ? 16384 : 0);
    {
        const int16_t d = (int16_t)(k_sigmoid[idx + 1] - k_sigmoid[idx]);
        const int16_t frac = (int16_t)(x_q14 & 0x3fff);
        v = (int16_t)(k_sigmoid[idx] + (int16_t)(rounded ? mul_round(d, frac, 14) : (d * frac) >> 14));
    }
    return (int16_t)(upper ? 8192 + v : 8192 - v);
}

static void nsx_speech_prob(orc_nsx_core *s, uint16_t *nonspeech, const uint32_t *prior_snr, const uint32_t *post_snr)
{
    int32_t lrt_sum = 0, ind, t;
    int16_t ind16, d16;
    int i, shifts, upper;
    for (i = 0; i < s->bins; ++i) {
        int32_t bessel = (int32_t)post_snr[i], frac32, lg, half_sum;
        const int n = orc_norm_u32(post_snr[i]);
        const uint32_t num = post_snr[i] << n;
        const uint32_t den = n > 10 ? sh_l(prior_snr[i], n - 11) : sh_r(prior_snr[i], 11 - n);
        int z;
        if (den > 0)
            bessel -= (int32_t)(num / den);
        else
            bessel = 0;
        z = orc_norm_u32(prior_snr[i]);
        frac32 = (int32_t)(((prior_snr[i] << z) & 0x7FFFFFFFu) >> 19);
        t = (frac32 * frac32 * -43) >> 19;
        t += ((int16_t)frac32 * 5412) >> 12;
        frac32 = t + 37;
        t = (int32_t)(((31 - z) << 12) + frac32) - (11 << 12);
        lg = (t * 178) >> 8;
        half_sum = (lg + s->lrt_avg[i]) / 2;
        s->lrt_avg[i] += bessel - half_sum;
        lrt_sum += s->lrt_avg[i];
    }
    s->feat_lrt = (lrt_sum * 10) >> (s->stages + 11);

    /* prior from the three features: sigmoid maps around the learned thresholds */
    t = lrt_sum - s->thr_lrt;
    shifts = 7 - s->stages;
    upper = 1;
    if (t < 0) {
        upper = 0;
        t = -t;
        ++shifts;
    }
    t = shift_w32(t, shifts);
    {
        /* this branch tests 0 <= index < 16 on the SIGNED shifted value (:98-100) */
        const int16_t idx = (int16_t)(t >> 14);
        int16_t v = (int16_t)(upper ? 16384 : 0);
        if (idx < 16 && idx >= 0) {
            const int16_t d = (int16_t)(k_sigmoid[idx + 1] - k_sigmoid[idx]);
            const int16_t frac = (int16_t)(t & 0x3fff);
            const int16_t y = (int16_t)(k_sigmoid[idx] + (int16_t)((d * frac) >> 14));
            v = (int16_t)(upper ? 8192 + y : 8192 - y);
        }
        ind = s->w_lrt * v;
    }
    if (s->w_flat) {
        const uint32_t f = s->feat_flat * 400u;
        uint32_t d = s->thr_flat - f;
        upper = 1;
        shifts = 4;
        if (s->thr_flat < f) {
            upper = 0;
            d = f - s->thr_flat;
            ++shifts;
        }
        ind += s->w_flat * sigmoid_q14((d << shifts) / 25u, upper, 0);
    }
    if (s->w_diff) {
        uint32_t u1 = 0, u2, u3;
        if (s->feat_diff) {
            int n = orc_norm_u32(s->feat_diff);
            if (n > 20 - s->stages)
                n = 20 - s->stages;
            u1 = s->feat_diff << n;
            u2 = s->time_avg_energy >> (20 - s->stages - n);
            u1 = u2 > 0 ? u1 / u2 : 0x7fffffffu;
        }
        u3 = (s->thr_diff << 17) / 25u;
        u2 = u1 - u3;
        shifts = 1;
        upper = 1;
        if (u2 & 0x80000000u) {
            upper = 0;
            u2 = u3 - u1;
            --shifts;
        }
        ind += s->w_diff * sigmoid_q14(u2 >> shifts, upper, 1);
    }
    ind16 = (int16_t)((98307 - ind) / 6);
    d16 = (int16_t)(ind16 - s->prior_nonspeech);
    s->prior_nonspeech = (int16_t)(s->prior_nonspeech + (int16_t)((1638 * d16) >> 14));

    memset(nonspeech, 0, sizeof(uint16_t) * (size_t)s->bins);
    if (s->prior_nonspeech <= 0)
        return;
    for (i = 0; i < s->bins; ++i) {
        int32_t inv_lrt, t2;
        int16_t int_part, frac;
        int n1, n2;
        if (s->lrt_avg[i] >= 65300)
            continue;
        t = (s->lrt_avg[i] * 23637) >> 14;
        int_part = (int16_t)(t >> 12);
        if (int_part < -8)
            int_part = -8;
        frac = (int16_t)(t & 0xfff);
        t2 = (frac * frac * 44) >> 19;
        t2 += (frac * 84) >> 7;
        inv_lrt = (int32_t)sh_l(1u, 8 + int_part) + shift_w32(t2, int_part - 4);
        n1 = orc_norm_w32(inv_lrt);
        n2 = norm_w16((int16_t)(16384 - s->prior_nonspeech));
        if (n1 + n2 < 7)
            continue;
        if (n1 + n2 < 15) {
            inv_lrt = sh_ra(inv_lrt, 15 - n2 - n1);
            t = inv_lrt * (16384 - s->prior_nonspeech);
            inv_lrt = shift_w32(t, 7 - n1 - n2);
        } else {
            t = inv_lrt * (16384 - s->prior_nonspeech);
            inv_lrt = t >> 8;
        }
        t = (int32_t)s->prior_nonspeech << 8;
        nonspeech[i] = (uint16_t)(t / (s->prior_nonspeech + inv_lrt));
    }
}

/* ---- parametric start-up noise: nsx_core.c:586-628 ---- */
static void nsx_pink_estimate(const orc_nsx_core *s, int16_t exp_avg, int32_t num_avg, int bin, uint32_t *est, uint32_t *est_avg)
{
    int32_t t2 = (exp_avg * g_t.log_index[bin]) >> 15;
    int32_t t1 = num_avg - t2;
    t1 += (s->min_norm - s->stages) * 2048;
    if (t1 > 0) {
        const int16_t int_part = (int16_t)(t1 >> 11), frac = (int16_t)(t1 & 0x7ff);
        if (frac >> 10) {
            t2 = (2048 - frac) * 1244;
            t2 = 2048 - (t2 >> 10);
        } else {
            t2 = (frac * 804) >> 10;
        }
        t2 = shift_w32(t2, int_part - 11);
        *est_avg = sh_l(1u, int_part) + (uint32_t)t2;
        *est = *est_avg * (uint32_t)(s->frame_idx + 1);
    }
}

/* ---- one frame: nsx_core.c:1501-2118 (ProcessCore) ---- */
static void nsx_frame(orc_nsx_core *s, const int16_t *in, int16_t *out, const int16_t *in_hb, int16_t *out_hb)
{
    uint16_t magn[BINS_MAX], noise_prev16[BINS_MAX], filter_model[BINS_MAX];
    uint32_t noise[BINS_MAX], post_snr[BINS_MAX], prior_snr[BINS_MAX], near_prev[BINS_MAX];
    const uint32_t sat_max = 1048575u;
    uint32_t max_noise = 0;
    int16_t q_magn, q_noise;
    int i, shifts, post_shifts, relearn, norm_noise;
    const int keep = s->ana - s->block;

    nsx_analysis(s, in, magn);
    if (s->zero_input) {
        nsx_synthesis(s, out);
        if (in_hb) {
            memmove(s->hb_buf, s->hb_buf + s->block, sizeof(int16_t) * (size_t)keep);
            memcpy(s->hb_buf + keep, in_hb, sizeof(int16_t) * (size_t)s->block);
            memcpy(out_hb, s->hb_buf, sizeof(int16_t) * (size_t)s->block);
        }
        return;
    }
    s->frame_idx++;
    q_magn = (int16_t)(s->norm - s->stages);
    nsx_flatness(s, magn);
    nsx_noise_estimate(s, magn, noise, &q_noise);
    for (i = 0; i < s->bins; ++i)
        noise_prev16[i] = (uint16_t)(s->noise_prev[i] >> 11);

    if (s->frame_idx < STARTUP_SHORT) {
        /* blend the quantile estimate with the white / pink model (:1614-1710) */
        const int q_use = q_noise < s->min_norm - s->stages ? q_noise : s->min_norm - s->stages;
        uint32_t est = 0, est_avg = 0;
        int16_t exp_avg = 0;
        int32_t num_avg = 0;
        if (s->pink_exp) {
            exp_avg = (int16_t)orc_div_w32_w16(s->pink_exp, (int16_t)(s->frame_idx + 1));
            num_avg = orc_div_w32_w16(s->pink_num, (int16_t)(s->frame_idx + 1));
            nsx_pink_estimate(s, exp_avg, num_avg, START_BAND, &est, &est_avg);
        } else {
            est = s->white_level;
            est_avg = est / (uint32_t)(s->frame_idx + 1);
        }
        for (i = 0; i < s->bins; ++i) {
            uint32_t a, b;
            if (s->pink_exp && i >= START_BAND) {
                est = 0;
                est_avg = 0;
                nsx_pink_estimate(s, exp_avg, num_avg, i, &est, &est_avg);
            }
            filter_model[i] = s->floor_gain;
            if (s->init_magn[i]) {
                uint32_t numer = s->init_magn[i] << 8;
                a = est * (uint32_t)s->overdrive;
                if (numer > a) {
                    int n;
                    numer -= a;
                    n = orc_norm_u32(numer);
                    if (n > 6)
                        n = 6;
                    numer <<= n;
                    a = s->init_magn[i] >> (6 - n);
                    if (a == 0)
                        a = 1;
                    b = numer / a;
                    filter_model[i] = (uint16_t)(b > 16384u ? 16384u : (b < (uint32_t)s->floor_gain ? (uint32_t)s->floor_gain : b));
                }
            }
            a = sh_r(noise[i], q_noise - q_use);
            b = sh_r(est_avg, s->min_norm - s->stages - q_use);
            shifts = 0;
            if (a & 0xfc000000u) {
                a >>= 6;
                b >>= 6;
                shifts = 6;
            }
            a *= (uint32_t)s->frame_idx;
            b *= (uint32_t)(STARTUP_SHORT - s->frame_idx);
            noise[i] = (a + b) / STARTUP_SHORT;
            noise[i] <<= shifts;
        }
        q_noise = (int16_t)q_use;
    }
    if (s->frame_idx < STARTUP_LONG) {
        s->time_avg_energy_acc += sh_r(s->magn_energy, 2 * s->norm + s->stages - 1);
        s->time_avg_energy = s->time_avg_energy_acc / (uint32_t)(uint16_t)(s->frame_idx + 1);
    }

    /* step 1: prior / posterior SNR against the quantile noise (:1724-1785) */
    post_shifts = 6 + q_magn - q_noise;
    shifts = 5 - s->q_magn_prev + s->q_noise_prev;
    for (i = 0; i < s->bins; ++i) {
        uint32_t a = (uint32_t)magn[i] << 6, b, c;
        post_snr[i] = 2048;
        b = post_shifts < 0 ? sh_r(noise[i], -post_shifts) : sh_l(noise[i], post_shifts);
        if (a > b) {
            a <<= 11;
            if (b > 0) {
                a /= b;
                post_snr[i] = a < sat_max ? a : sat_max;
            } else {
                post_snr[i] = sat_max;
            }
        }
        a = ((uint32_t)s->magn_prev[i] * (uint32_t)s->filter[i]) << 3;
        b = sh_r(s->noise_prev[i], shifts);
        if (b > 0) {
            a /= b;
            if (a > sat_max)
                a = sat_max;
        } else {
            a = sat_max;
        }
        near_prev[i] = a;
        b = near_prev[i] * 2007u;
        c = (post_snr[i] - 2048u) * 41u;
        prior_snr[i] = 2048u + ((b + c + 512u) >> 10);
    }

    /* step 2: features, prior model, per-bin speech probability, noise update (:1788-1946) */
    nsx_difference(s, magn);
    s->model_count++;
    relearn = s->model_count == s->model_window;
    nsx_feature_update(s, relearn);
    if (relearn) {
        uint32_t avg;
        s->model_count = 0;
        s->cur_avg_energy >>= 9;
        avg = (s->cur_avg_energy + s->time_avg_energy + 1) >> 1;
        if (avg != s->time_avg_energy && s->feat_diff && s->time_avg_energy > 0) {
            uint32_t a = avg, b = s->feat_diff;
            int n = 0;
            while (a & 0xFFFF0000u) {
                a >>= 1;
                ++n;
            }
            while (b & 0xFFFF0000u) {
                b >>= 1;
                ++n;
            }
            a = a * b;
            a /= s->time_avg_energy;
            if (orc_norm_u32(a) < n)
                s->feat_diff = 0x007FFFFF;
            else
                s->feat_diff = 0x007FFFFFu < sh_l(a, n) ? 0x007FFFFFu : sh_l(a, n);
        }
        s->time_avg_energy = avg;
        s->cur_avg_energy = 0;
    }
    nsx_speech_prob(s, s->nonspeech, prior_snr, post_snr);

    {
        uint16_t gamma = 26, gamma_prev;
        post_shifts = s->q_noise_prev - q_magn;
        shifts = s->q_magn_prev - q_magn;
        for (i = 0; i < s->bins; ++i) {
            uint32_t m, d, upd = s->noise_prev[i], w = 0, step;
            int32_t pause, dp;
            int up;
            m = post_shifts < 0 ? sh_r((uint32_t)magn[i], -post_shifts) : sh_l((uint32_t)magn[i], post_shifts);
            if (noise_prev16[i] > m) {
                up = 0;
                d = noise_prev16[i] - m;
            } else {
                up = 1;
                d = m - noise_prev16[i];
            }
            if (d && s->nonspeech[i]) {
                w = d * (uint32_t)s->nonspeech[i];
                step = (w & 0x7c000000u) ? (w >> 5) * gamma : (w * gamma) >> 5;
                upd = up ? upd + step : upd - step;
            }
            gamma_prev = gamma;
            gamma = s->nonspeech[i] < 205 ? 3 : 26;
            if (gamma_prev != gamma) {
                uint32_t alt;
                step = (w & 0x7c000000u) ? (w >> 5) * gamma : (w * gamma) >> 5;
                alt = up ? s->noise_prev[i] + step : s->noise_prev[i] - step;
                if (upd > alt)
                    upd = alt;
            }
            noise[i] = upd;
            if (upd > max_noise)
                max_noise = upd;
            pause = shift_w32(s->pause_avg[i], -shifts);
            if (s->nonspeech[i] > 205) {
                if (shifts < 0) {
                    dp = (int32_t)magn[i] - pause;
                    dp *= 13;
                    dp = (dp + 128) >> 8;
                } else {
                    dp = (int32_t)sh_l((uint32_t)magn[i], shifts) - s->pause_avg[i];
                    dp *= 13;
                    dp = sh_ra(dp + (int32_t)sh_l(128u, shifts), 8 + shifts);
                }
                pause += dp;
            }
            s->pause_avg[i] = pause;
        }
    }
    norm_noise = orc_norm_u32(max_noise);
    q_noise = (int16_t)(s->q_noise_prev + norm_noise - 5);

    /* step 3: Wiener gain from the updated noise (:1948-2014) */
    shifts = s->q_noise_prev + 11 - q_magn;
    for (i = 0; i < s->bins; ++i) {
        uint32_t cur = 0, m, nz, a, b, prior;
        uint16_t g;
        if (shifts < 0) {
            m = magn[i];
            nz = sh_l(noise[i], -shifts);
        } else if (shifts > 17) {
            m = (uint32_t)magn[i] << 17;
            nz = sh_r(noise[i], shifts - 17);
        } else {
            m = sh_l((uint32_t)magn[i], shifts);
            nz = noise[i];
        }
        if (m > nz) {
            int n;
            a = m - nz;
            n = orc_norm_u32(a);
            if (n > 11)
                n = 11;
            a <<= n;
            b = nz >> (11 - n);
            if (b > 0)
                a /= b;
            cur = a < sat_max ? a : sat_max;
        }
        prior = near_prev[i] * 2007u + cur * 41u;
        a = (uint32_t)s->overdrive + ((prior + 8192u) >> 14);
        g = (uint16_t)((prior + a / 2) / a);
        s->filter[i] = (uint16_t)(g > 16384 ? 16384 : (g < s->floor_gain ? s->floor_gain : g));
        if (s->frame_idx < STARTUP_SHORT) {
            a = (uint32_t)s->filter[i] * (uint32_t)s->frame_idx;
            a += (uint32_t)filter_model[i] * (uint32_t)(STARTUP_SHORT - s->frame_idx);
            s->filter[i] = (uint16_t)(a / STARTUP_SHORT);
        }
    }
    s->q_noise_prev = q_noise;
    s->q_magn_prev = q_magn;
    for (i = 0; i < s->bins; ++i) {
        s->noise_prev[i] = norm_noise > 5 ? noise[i] << (norm_noise - 5) : noise[i] >> (5 - norm_noise);
        s->magn_prev[i] = magn[i];
    }
    nsx_synthesis(s, out);

    if (in_hb) {
        /* second band: delayed by the analysis overlap, one gain per frame from the upper quarter of the
         * low band's speech probability and filter (:2042-2117) */
        uint32_t sum_filter = 0;
        uint16_t sum_prob = 0;
        int16_t avg_prob, avg_filter, gain_mod, gain;
        memmove(s->hb_buf, s->hb_buf + s->block, sizeof(int16_t) * (size_t)keep);
        memcpy(s->hb_buf + keep, in_hb, sizeof(int16_t) * (size_t)s->block);
        for (i = s->half - (s->half >> 2); i < s->half; ++i) {
            sum_prob = (uint16_t)(sum_prob + s->nonspeech[i]);
            sum_filter += s->filter[i];
        }
        avg_prob = (int16_t)(4096 - (sum_prob >> (s->stages - 7)));
        avg_filter = (int16_t)(sum_filter >> (s->stages - 3));
        gain_mod = avg_prob < 3607 ? avg_prob : 3607;
        if (avg_prob < 2048) {
            gain = (int16_t)((gain_mod << 1) + (avg_filter >> 1));
        } else {
            gain = (int16_t)((3 * avg_filter) >> 2);
            gain = (int16_t)(gain + gain_mod);
        }
        gain = (int16_t)(gain > 16384 ? 16384 : (gain < (int16_t)s->floor_gain ? (int16_t)s->floor_gain : gain));
        for (i = 0; i < s->block; ++i)
            out_hb[i] = (int16_t)((gain * s->hb_buf[i]) >> 14);
    }
}

/* ---- handle layer: R:src/webrtc.c:563-660 with MAKE_WEBRTC_NSX defined (:512) ---- */
orc_nsx *orc_nsx_init(int chn, int freq) { return orc_nsx_init_policy(chn, freq, 2); }

orc_nsx *orc_nsx_init_policy(int chn, int freq, int policy)
{
    orc_nsx *h;
    if (freq > 32000 || freq % 8000 != 0)
        return NULL;
    if (chn != 1 && chn != 2)
        return NULL;
    h = (orc_nsx *)calloc(1, sizeof *h);
    h->core = (orc_nsx_core *)malloc(sizeof(orc_nsx_core));
    if (nsx_core_init(h->core, freq) != 0 || nsx_set_policy(h->core, policy) != 0) {
        free(h->core);
        free(h);
        return NULL;
    }
    h->chn = chn;
    h->freq = freq;
    h->pkg = freq / 1000 * 10;
    return h;
}

void orc_nsx_process(orc_nsx *h, const int16_t *in, int16_t *out, int frame_num)
{
    int16_t lo[160], hi[160], lo_out[160], hi_out[160];
    orc_nsx_core *s = h->core;
    const int chn = h->chn;
    int pos, i;
    /* 32 kHz: the packet is 320 samples, the core reads and writes one 160-sample block of it, the rest of the
     * reference's calloc'ed out[] stays zero (the same quirk as the float path, orc_ns.c) */
    for (pos = 0; pos + h->pkg <= frame_num; pos += h->pkg) {
        for (i = 0; i < s->block; ++i) {
            lo[i] = in[(pos + i) * chn];
            if (chn == 2)
                hi[i] = in[(pos + i) * chn + 1];
        }
        nsx_frame(s, lo, lo_out, chn == 2 ? hi : NULL, chn == 2 ? hi_out : NULL);
        for (i = 0; i < h->pkg; ++i) {
            out[(pos + i) * chn] = i < s->block ? lo_out[i] : 0;
            if (chn == 2)
                out[(pos + i) * chn + 1] = i < s->block ? hi_out[i] : 0;
        }
    }
}

void orc_nsx_release(orc_nsx *h)
{
    if (!h)
        return;
    free(h->core);
    free(h);
}

int orc_nsx_block_index(const orc_nsx *h) { return h->core->frame_idx; }

/* canonical int32 dump of the per-stream state, for state-level comparisons with the CUDA path:
 * [0..15] scalars, then per bin (bins entries each): filter, lquant x3, density x3, quant, lrt_avg, pause_avg,
 * init_magn, noise_prev, magn_prev; then ana history (ana-block), synthesis tail (ana-block) */
int orc_nsx_state(const orc_nsx *h, int32_t *out, int cap)
{
    const orc_nsx_core *s = h->core;
    const int keep = s->ana - s->block, n = 40 + 13 * s->bins + 2 * keep;
    int i, k = 0, e;
    if (!out || cap < n)
        return n;
    out[k++] = s->frame_idx; out[k++] = s->model_count; out[k++] = s->counter[0]; out[k++] = s->counter[1];
    out[k++] = s->counter[2]; out[k++] = s->q_noise; out[k++] = s->q_noise_prev; out[k++] = s->q_magn_prev;
    out[k++] = s->min_norm; out[k++] = s->prior_nonspeech; out[k++] = s->feat_lrt; out[k++] = s->thr_lrt;
    out[k++] = (int32_t)s->feat_flat; out[k++] = (int32_t)s->thr_flat; out[k++] = (int32_t)s->feat_diff; out[k++] = (int32_t)s->thr_diff;
    out[k++] = s->w_lrt; out[k++] = s->w_flat; out[k++] = s->w_diff; out[k++] = (int32_t)s->cur_avg_energy;
    out[k++] = (int32_t)s->time_avg_energy; out[k++] = (int32_t)s->time_avg_energy_acc; out[k++] = (int32_t)s->white_level;
    out[k++] = s->pink_num; out[k++] = s->pink_exp;
    while (k < 40)
        out[k++] = 0;
    for (i = 0; i < s->bins; ++i) out[k++] = s->filter[i];
    for (e = 0; e < N_EST; ++e)
        for (i = 0; i < s->bins; ++i) out[k++] = s->lquant[e * s->bins + i];
    for (e = 0; e < N_EST; ++e)
        for (i = 0; i < s->bins; ++i) out[k++] = s->density[e * s->bins + i];
    for (i = 0; i < s->bins; ++i) out[k++] = s->quant[i];
    for (i = 0; i < s->bins; ++i) out[k++] = s->lrt_avg[i];
    for (i = 0; i < s->bins; ++i) out[k++] = s->pause_avg[i];
    for (i = 0; i < s->bins; ++i) out[k++] = (int32_t)s->init_magn[i];
    for (i = 0; i < s->bins; ++i) out[k++] = (int32_t)s->noise_prev[i];
    for (i = 0; i < s->bins; ++i) out[k++] = s->magn_prev[i];
    for (i = 0; i < keep; ++i) out[k++] = s->ana_buf[s->block + i];
    for (i = 0; i < keep; ++i) out[k++] = s->syn_buf[i];
    return n;
}
