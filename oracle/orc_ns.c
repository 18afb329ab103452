/* ORACLE (test infrastructure) — WebRTC float noise suppressor restated from
 * T:webrtc/modules/audio_processing/ns/ns_core.c (+ noise_suppression.c) and the wmix handle
 * layer R:src/webrtc.c:560-660.  Mono only (wmix's stereo path feeds the right channel in
 * as a "high band"; not restated).
 *
 * Numerics follow the reference literally: float arithmetic with no contraction, the same
 * summation order, and *double* libm log/exp/pow/tanh/sqrt rounded back to float exactly
 * where ns_core.c does so.  Build with -ffp-contract=off.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define ANA_MAX 256
#define BINS_MAX 129
#define NHIST 1000
#define STARTUP_SHORT 50
#define STARTUP_LONG 200

struct orc_ns_core {
    int fs, block, ana, bins;
    float window[ANA_MAX];
    float inbuf[ANA_MAX];      /* analyzeBuf and dataBuf: always fed the same frames */
    float inbuf_hb[ANA_MAX];   /* dataBufHB[0]: wmix's right channel when chn == 2 (R:src/webrtc.c:633) */
    float synth[ANA_MAX];
    float density[3 * BINS_MAX], lquant[3 * BINS_MAX], quant[BINS_MAX];
    int counter[3], updates;
    float smooth[BINS_MAX];
    float overdrive, floor_gain;
    int gainmap;
    int ip[ANA_MAX / 2];
    float wfft[ANA_MAX / 2];
    int frame_idx;             /* blockInd */
    int upd_mode, upd_window, upd_countdown; /* modelUpdatePars[0], [1], [3] */
    float prior_model[7];
    float noise[BINS_MAX], noise_prev[BINS_MAX];
    float magn_prev_an[BINS_MAX], magn_prev_pr[BINS_MAX];
    float lrt_avg[BINS_MAX];
    float prior_prob;
    float feat[7];
    float pause_avg[BINS_MAX];
    float signal_energy, sum_magn;
    float white_level, pink_num, pink_exp;
    float init_magn[BINS_MAX], param_noise[BINS_MAX];
    float speech_prob[BINS_MAX];
    int hist_lrt[NHIST], hist_flat[NHIST], hist_diff[NHIST];
};

/* T:.../ns/windows_private.h:64 (kBlocks80w128), :94 (kBlocks160w256): a flat-top window whose
 * edges are quarter sine waves.  The reference stores it printed to 8 decimals; reproducing
 * that print-and-parse step gives the identical floats (tests compare all 384 entries). */
static void orc_ns_make_window(float *w, int ana, int block)
{
    int ov = ana - block, i;
    for (i = 0; i < ana; ++i) {
        int k = i < ana - i ? i : ana - i;
        char txt[32];
        if (k > ov)
            k = ov;
        snprintf(txt, sizeof txt, "%.8f", sin(3.14159265358979323846 * k / (2.0 * ov)));
        w[i] = strtof(txt, NULL);
    }
}

/* T:.../ns/ns_core.c:72-215 (InitCore) + :1012-1041 (set_policy_core, mode 2) */
static int orc_ns_core_init(orc_ns_core *s, int fs, int mode)
{
    int i;
    memset(s, 0, sizeof(*s));
    if (fs != 8000 && fs != 16000 && fs != 32000 && fs != 48000)
        return -1;
    s->fs = fs;
    if (fs == 8000) {
        s->block = 80;
        s->ana = 128;
    } else {
        s->block = 160;
        s->ana = 256;
    }
    s->bins = s->ana / 2 + 1;
    orc_ns_make_window(s->window, s->ana, s->block);
    s->ip[0] = 0;
    orc_rdft(s->ana, 1, s->inbuf, s->ip, s->wfft);   /* builds the tables */
    memset(s->inbuf, 0, sizeof(s->inbuf));
    for (i = 0; i < 3 * BINS_MAX; ++i) {
        s->lquant[i] = 8.f;
        s->density[i] = 0.3f;
    }
    for (i = 0; i < 3; ++i)
        s->counter[i] = (int)floor((float)(STARTUP_LONG * (i + 1)) / (float)3);
    for (i = 0; i < BINS_MAX; ++i) {
        s->smooth[i] = 1.f;
        s->lrt_avg[i] = 0.5f;
    }
    s->prior_prob = 0.5f;
    s->feat[0] = 0.5f;
    s->feat[3] = 0.5f;
    s->feat[4] = 0.5f;
    s->frame_idx = -1;
    s->prior_model[0] = 0.5f;
    s->prior_model[1] = 0.5f;
    s->prior_model[2] = 1.f;
    s->prior_model[3] = 0.5f;
    s->prior_model[4] = 1.f;
    s->upd_mode = 2;
    s->upd_window = 500;
    s->upd_countdown = 500;
    switch (mode) {
    case 0: s->overdrive = 1.f; s->floor_gain = 0.5f; s->gainmap = 0; break;
    case 1: s->overdrive = 1.f; s->floor_gain = 0.25f; s->gainmap = 1; break;
    case 2: s->overdrive = 1.1f; s->floor_gain = 0.125f; s->gainmap = 1; break;
    case 3: s->overdrive = 1.25f; s->floor_gain = 0.09f; s->gainmap = 1; break;
    default: return -1;
    }
    return 0;
}

static void orc_slide_in(float *buf, int ana, int block, const float *frame)
{
    memmove(buf, buf + block, sizeof(float) * (size_t)(ana - block));
    if (frame)
        memcpy(buf + ana - block, frame, sizeof(float) * (size_t)block);
    else
        memset(buf + ana - block, 0, sizeof(float) * (size_t)block);
}

static float orc_energy_f(const float *x, int n)
{
    float e = 0.f;
    int i;
    for (i = 0; i < n; ++i)
        e += x[i] * x[i];
    return e;
}

/* window -> rdft -> re/im/|X|+1   (T:.../ns/ns_core.c:886-911) */
static void orc_ns_spectrum(orc_ns_core *s, float *t, float *re, float *im, float *mag)
{
    int i, nb = s->bins;
    orc_rdft(s->ana, 1, t, s->ip, s->wfft);
    im[0] = 0;
    re[0] = t[0];
    mag[0] = (float)(fabs(re[0]) + 1.f);
    im[nb - 1] = 0;
    re[nb - 1] = t[1];
    mag[nb - 1] = (float)(fabs(re[nb - 1]) + 1.f);
    for (i = 1; i < nb - 1; ++i) {
        re[i] = t[2 * i];
        im[i] = t[2 * i + 1];
        mag[i] = sqrtf(re[i] * re[i] + im[i] * im[i]) + 1.f;
    }
}

/* three staggered log-quantile trackers  (T:.../ns/ns_core.c:217-285) */
static void orc_ns_quantile_noise(orc_ns_core *s, const float *mag, float *noise)
{
    float lm[BINS_MAX];
    int i, t, off = 0, nb = s->bins;
    if (s->updates < STARTUP_LONG)
        s->updates++;
    for (i = 0; i < nb; ++i)
        lm[i] = (float)log(mag[i]);
    for (t = 0; t < 3; ++t) {
        off = t * nb;
        for (i = 0; i < nb; ++i) {
            float step;
            if (s->density[off + i] > 1.0)
                step = 40.f * 1.f / s->density[off + i];
            else
                step = 40.f;
            if (lm[i] > s->lquant[off + i])
                s->lquant[off + i] += 0.25f * step / (float)(s->counter[t] + 1);
            else
                s->lquant[off + i] -= (1.f - 0.25f) * step / (float)(s->counter[t] + 1);
            if (fabs(lm[i] - s->lquant[off + i]) < 0.01f)
                s->density[off + i] = ((float)s->counter[t] * s->density[off + i] + 1.f / (2.f * 0.01f)) /
                                      (float)(s->counter[t] + 1);
        }
        if (s->counter[t] >= STARTUP_LONG) {
            s->counter[t] = 0;
            if (s->updates >= STARTUP_LONG)
                for (i = 0; i < nb; ++i)
                    s->quant[i] = (float)exp(s->lquant[off + i]);
        }
        s->counter[t]++;
    }
    if (s->updates < STARTUP_LONG)
        for (i = 0; i < nb; ++i)
            s->quant[i] = (float)exp(s->lquant[off + i]);
    for (i = 0; i < nb; ++i)
        noise[i] = s->quant[i];
}

/* histogram bookkeeping and the every-500-frames threshold re-learn
 * (T:.../ns/ns_core.c:293-520) */
static void orc_ns_hist_add(int *h, float v, float bin)
{
    if (v < NHIST * bin && v >= 0.0)
        h[(int)(v / bin)]++;
}

static void orc_two_peaks(const int *h, float bin, float *p1, float *p2, int *w1, int *w2)
{
    int i, m1 = 0, m2 = 0;
    *p1 = *p2 = 0.f;
    *w1 = *w2 = 0;
    for (i = 0; i < NHIST; ++i) {
        float mid = ((float)i + 0.5f) * bin;
        if (h[i] > m1) {
            m2 = m1;
            *w2 = *w1;
            *p2 = *p1;
            m1 = h[i];
            *w1 = h[i];
            *p1 = mid;
        } else if (h[i] > m2) {
            m2 = h[i];
            *w2 = h[i];
            *p2 = mid;
        }
    }
}

static void orc_ns_relearn(orc_ns_core *s)
{
    const float binL = 0.1f, binF = 0.05f, binD = 0.1f;
    const int min_weight = (int)(0.3 * s->upd_window);
    float avg = 0.f, avg_all = 0.f, avg_sq = 0.f, fluct, p1, p2;
    int n = 0, i, w1, w2, use_flat = 1, use_diff = 1;
    float fsum;

    for (i = 0; i < NHIST; ++i) {
        float mid = ((float)i + 0.5f) * binL;
        if (mid <= 1.f) {
            avg += s->hist_lrt[i] * mid;
            n += s->hist_lrt[i];
        }
        avg_sq += s->hist_lrt[i] * mid * mid;
        avg_all += s->hist_lrt[i] * mid;
    }
    if (n > 0)
        avg = avg / ((float)n);
    avg_all = avg_all / ((float)s->upd_window);
    avg_sq = avg_sq / ((float)s->upd_window);
    fluct = avg_sq - avg * avg_all;
    if (fluct < 0.05f) {
        s->prior_model[0] = 1.f;
    } else {
        s->prior_model[0] = 1.2f * avg;
        if (s->prior_model[0] < 0.2f)
            s->prior_model[0] = 0.2f;
        if (s->prior_model[0] > 1.f)
            s->prior_model[0] = 1.f;
    }

    orc_two_peaks(s->hist_flat, binF, &p1, &p2, &w1, &w2);
    if ((fabs(p2 - p1) < 2 * binF) && (w2 > 0.5f * w1)) {
        w1 += w2;
        p1 = 0.5f * (p1 + p2);
    }
    if (w1 < min_weight || p1 < 0.6f)
        use_flat = 0;
    if (use_flat) {
        s->prior_model[1] = 0.9f * p1;
        if (s->prior_model[1] < 0.1f)
            s->prior_model[1] = 0.1f;
        if (s->prior_model[1] > 0.95f)
            s->prior_model[1] = 0.95f;
    }

    orc_two_peaks(s->hist_diff, binD, &p1, &p2, &w1, &w2);
    if ((fabs(p2 - p1) < 2 * binD) && (w2 > 0.5f * w1)) {
        w1 += w2;
        p1 = 0.5f * (p1 + p2);
    }
    s->prior_model[3] = 1.2f * p1;
    if (w1 < min_weight)
        use_diff = 0;
    if (s->prior_model[3] < 0.16f)
        s->prior_model[3] = 0.16f;
    if (s->prior_model[3] > 1.f)
        s->prior_model[3] = 1.f;
    if (fluct < 0.05f)
        use_diff = 0;

    fsum = (float)(1 + use_flat + use_diff);
    s->prior_model[4] = 1.f / fsum;
    s->prior_model[5] = ((float)use_flat) / fsum;
    s->prior_model[6] = ((float)use_diff) / fsum;
    if (s->upd_mode >= 1) {
        memset(s->hist_lrt, 0, sizeof(s->hist_lrt));
        memset(s->hist_flat, 0, sizeof(s->hist_flat));
        memset(s->hist_diff, 0, sizeof(s->hist_diff));
    }
}

/* T:.../ns/ns_core.c:1043-1181 */
static void orc_ns_analyze(orc_ns_core *s, const float *frame)
{
    const int first = 5;
    int i, nb, flag = s->upd_mode;
    float t[ANA_MAX], re[ANA_MAX], im[BINS_MAX], mag[BINS_MAX], noise[BINS_MAX];
    float post[BINS_MAX], prior[BINS_MAX];
    float energy, sig_e = 0.f, sum_mag = 0.f;
    float s_li = 0.f, s_li2 = 0.f, s_lm = 0.f, s_lilm = 0.f;
    float f1, f2, f3, pnum = 0.f, pexp = 0.f;

    nb = s->bins;
    /* the analysis buffer itself is advanced by the caller (shared with the process pass) */
    for (i = 0; i < s->ana; ++i)
        t[i] = s->window[i] * s->inbuf[i];
    energy = orc_energy_f(t, s->ana);
    if (energy == 0.0)
        return;
    (void)frame;
    s->frame_idx++;
    orc_ns_spectrum(s, t, re, im, mag);

    for (i = 0; i < nb; ++i) {
        sig_e += re[i] * re[i] + im[i] * im[i];
        sum_mag += mag[i];
        if (s->frame_idx < STARTUP_SHORT && i >= first) {
            f2 = log((float)i);
            s_li += f2;
            s_li2 += f2 * f2;
            f1 = log(mag[i]);
            s_lm += f1;
            s_lilm += f2 * f1;
        }
    }
    sig_e = sig_e / ((float)nb);
    s->signal_energy = sig_e;
    s->sum_magn = sum_mag;

    orc_ns_quantile_noise(s, mag, noise);

    if (s->frame_idx < STARTUP_SHORT) {
        s->white_level += sum_mag / ((float)nb) * s->overdrive;
        f1 = s_li2 * ((float)(nb - first));
        f1 -= (s_li * s_li);
        f2 = (s_li2 * s_lm - s_li * s_lilm);
        f3 = f2 / f1;
        if (f3 < 0.f)
            f3 = 0.f;
        s->pink_num += f3;
        f2 = (s_li * s_lm);
        f2 -= ((float)(nb - first)) * s_lilm;
        f3 = f2 / f1;
        if (f3 < 0.f)
            f3 = 0.f;
        if (f3 > 1.f)
            f3 = 1.f;
        s->pink_exp += f3;
        if (s->pink_exp > 0.f) {
            pnum = exp(s->pink_num / (float)(s->frame_idx + 1));
            pnum *= (float)(s->frame_idx + 1);
            pexp = s->pink_exp / (float)(s->frame_idx + 1);
        }
        for (i = 0; i < nb; ++i) {
            if (s->pink_exp == 0.f) {
                s->param_noise[i] = s->white_level;
            } else {
                float band = (float)(i < first ? first : i);
                s->param_noise[i] = pnum / pow(band, pexp);
            }
            noise[i] *= (s->frame_idx);
            f2 = s->param_noise[i] * (STARTUP_SHORT - s->frame_idx);
            noise[i] += (f2 / (float)(s->frame_idx + 1));
            noise[i] /= STARTUP_SHORT;
        }
    }
    if (s->frame_idx < STARTUP_LONG) {
        s->feat[5] *= s->frame_idx;
        s->feat[5] += sig_e;
        s->feat[5] /= (s->frame_idx + 1);
    }

    /* decision-directed SNR (T:.../ns/ns_core.c:566-589) */
    for (i = 0; i < nb; ++i) {
        float prev = s->magn_prev_an[i] / (s->noise_prev[i] + 0.0001f) * s->smooth[i];
        post[i] = 0.f;
        if (mag[i] > noise[i])
            post[i] = mag[i] / (noise[i] + 0.0001f) - 1.f;
        prior[i] = 0.98f * prev + (1.f - 0.98f) * post[i];
    }

    /* spectral flatness (T:.../ns/ns_core.c:523-557) */
    {
        float num = 0.f, den = s->sum_magn, v;
        int ok = 1;
        den -= mag[0];
        for (i = 1; i < nb; ++i) {
            if (mag[i] > 0.0) {
                num += (float)log(mag[i]);
            } else {
                s->feat[0] -= 0.3f * s->feat[0];
                ok = 0;
                break;
            }
        }
        if (ok) {
            den = den / nb;
            num = num / nb;
            v = (float)exp(num) / den;
            s->feat[0] += 0.3f * (v - s->feat[0]);
        }
    }
    /* spectral difference against the pause template (T:.../ns/ns_core.c:595-633) */
    {
        float ap = 0.f, am = s->sum_magn, cov = 0.f, vp = 0.f, vm = 0.f, d;
        for (i = 0; i < nb; ++i)
            ap += s->pause_avg[i];
        ap = ap / ((float)nb);
        am = am / ((float)nb);
        for (i = 0; i < nb; ++i) {
            cov += (mag[i] - am) * (s->pause_avg[i] - ap);
            vp += (s->pause_avg[i] - ap) * (s->pause_avg[i] - ap);
            vm += (mag[i] - am) * (mag[i] - am);
        }
        cov = cov / ((float)nb);
        vp = vp / ((float)nb);
        vm = vm / ((float)nb);
        s->feat[6] += s->signal_energy;
        d = vm - (cov * cov) / (vp + 0.0001f);
        d = (float)(d / (s->feat[5] + 0.0001f));
        s->feat[4] += 0.3f * (d - s->feat[4]);
    }
    /* histograms / re-learn (T:.../ns/ns_core.c:755-790) */
    if (flag >= 1) {
        s->upd_countdown--;
        if (s->upd_countdown > 0) {
            orc_ns_hist_add(s->hist_lrt, s->feat[3], 0.1f);
            orc_ns_hist_add(s->hist_flat, s->feat[0], 0.05f);
            orc_ns_hist_add(s->hist_diff, s->feat[4], 0.1f);
        }
        if (s->upd_countdown == 0) {
            orc_ns_relearn(s);
            s->upd_countdown = s->upd_window;
            if (flag == 1) {
                s->upd_mode = 0;
            } else {
                s->feat[6] = s->feat[6] / ((float)s->upd_window);
                s->feat[5] = 0.5f * (s->feat[6] + s->feat[5]);
                s->feat[6] = 0.f;
            }
        }
    }

    /* speech probability (T:.../ns/ns_core.c:642-748) */
    {
        float ksum = 0.f, ind0, ind1, ind2, width, ind, gain_prior, x;
        float thr0 = s->prior_model[0], thr1 = s->prior_model[1], thr2 = s->prior_model[3];
        int sgn = (int)(s->prior_model[2]);
        for (i = 0; i < nb; ++i) {
            float a = 1.f + 2.f * prior[i];
            float b = 2.f * prior[i] / (a + 0.0001f);
            float bessel = (post[i] + 1.f) * b;
            s->lrt_avg[i] += 0.5f * (bessel - (float)log(a) - s->lrt_avg[i]);
            ksum += s->lrt_avg[i];
        }
        ksum = (float)ksum / (nb);
        s->feat[3] = ksum;
        width = 4.f;
        if (ksum < thr0)
            width = 2.f * 4.f;
        ind0 = 0.5f * ((float)tanh(width * (ksum - thr0)) + 1.f);
        x = s->feat[0];
        width = 4.f;
        if (sgn == 1 && (x > thr1))
            width = 2.f * 4.f;
        if (sgn == -1 && (x < thr1))
            width = 2.f * 4.f;
        ind1 = 0.5f * ((float)tanh((float)sgn * width * (thr1 - x)) + 1.f);
        x = s->feat[4];
        width = 4.f;
        if (x < thr2)
            width = 2.f * 4.f;
        ind2 = 0.5f * ((float)tanh(width * (x - thr2)) + 1.f);
        ind = s->prior_model[4] * ind0 + s->prior_model[5] * ind1 + s->prior_model[6] * ind2;
        s->prior_prob += 0.1f * (ind - s->prior_prob);
        if (s->prior_prob > 1.f)
            s->prior_prob = 1.f;
        if (s->prior_prob < 0.01f)
            s->prior_prob = 0.01f;
        gain_prior = (1.f - s->prior_prob) / (s->prior_prob + 0.0001f);
        for (i = 0; i < nb; ++i) {
            float inv = (float)exp(-s->lrt_avg[i]);
            inv = (float)gain_prior * inv;
            s->speech_prob[i] = 1.f / (1.f + inv);
        }
    }

    /* noise update (T:.../ns/ns_core.c:800-846); the smoothing constant of bin i-1 leaks
     * into the provisional estimate of bin i exactly as in the reference loop */
    {
        float gamma = 0.9f;
        for (i = 0; i < nb; ++i) {
            float ps = s->speech_prob[i], pn = 1.f - ps, prov, gamma_old;
            prov = gamma * s->noise_prev[i] + (1.f - gamma) * (pn * mag[i] + ps * s->noise_prev[i]);
            gamma_old = gamma;
            gamma = 0.9f;
            if (ps > 0.2f)
                gamma = 0.99f;
            if (ps < 0.2f)
                s->pause_avg[i] += 0.05f * (mag[i] - s->pause_avg[i]);
            if (gamma == gamma_old) {
                noise[i] = prov;
            } else {
                noise[i] = gamma * s->noise_prev[i] + (1.f - gamma) * (pn * mag[i] + ps * s->noise_prev[i]);
                if (prov < noise[i])
                    noise[i] = prov;
            }
        }
    }
    memcpy(s->noise, noise, sizeof(float) * (size_t)nb);
    memcpy(s->magn_prev_an, mag, sizeof(float) * (size_t)nb);
}

/* T:.../ns/ns_core.c:1183-1415 (single band) */
/* out_hb != NULL: ProcessCore with num_bands = 2 (ns_core.c:1214-1234, :1252-1261, :1361-1414).  The "high band" only
 * gets a time-domain gain derived from the low band's speech probability and filter over the upper quarter of the
 * spectrum; its samples come out of dataBufHB, i.e. delayed by ana - block. */
static void orc_ns_synthesize(orc_ns_core *s, float *out, const float *in_hb, float *out_hb)
{
    int i, nb = s->bins;
    float t[ANA_MAX], re[ANA_MAX], im[BINS_MAX], mag[BINS_MAX], h[BINS_MAX], h0[BINS_MAX];
    float fout[160];
    float e1, e2, gain, factor, f1, f2;

    if (out_hb)
        orc_slide_in(s->inbuf_hb, s->ana, s->block, in_hb);
    for (i = 0; i < s->ana; ++i)
        t[i] = s->window[i] * s->inbuf[i];
    e1 = orc_energy_f(t, s->ana);
    if (e1 == 0.0) {
        for (i = 0; i < s->block; ++i)
            fout[i] = s->synth[i];
        orc_slide_in(s->synth, s->ana, s->block, NULL);
        for (i = 0; i < s->block; ++i)
            out[i] = fout[i] > 32767 ? 32767 : (fout[i] < -32768 ? -32768 : fout[i]);
        if (out_hb)
            for (i = 0; i < s->block; ++i)
                out_hb[i] = s->inbuf_hb[i] > 32767 ? 32767 : (s->inbuf_hb[i] < -32768 ? -32768 : s->inbuf_hb[i]);
        return;
    }
    orc_ns_spectrum(s, t, re, im, mag);
    if (s->frame_idx < STARTUP_SHORT)
        for (i = 0; i < nb; ++i)
            s->init_magn[i] += mag[i];

    for (i = 0; i < nb; ++i) {
        float prev = s->magn_prev_pr[i] / (s->noise_prev[i] + 0.0001f) * s->smooth[i];
        float cur = 0.f, snr;
        if (mag[i] > s->noise[i])
            cur = mag[i] / (s->noise[i] + 0.0001f) - 1.f;
        snr = 0.98f * prev + (1.f - 0.98f) * cur;
        h[i] = snr / (s->overdrive + snr);
    }
    for (i = 0; i < nb; ++i) {
        if (h[i] < s->floor_gain)
            h[i] = s->floor_gain;
        if (h[i] > 1.f)
            h[i] = 1.f;
        if (s->frame_idx < STARTUP_SHORT) {
            h0[i] = (s->init_magn[i] - s->overdrive * s->param_noise[i]);
            h0[i] /= (s->init_magn[i] + 0.0001f);
            if (h0[i] < s->floor_gain)
                h0[i] = s->floor_gain;
            if (h0[i] > 1.f)
                h0[i] = 1.f;
            h[i] *= (s->frame_idx);
            h0[i] *= (STARTUP_SHORT - s->frame_idx);
            h[i] += h0[i];
            h[i] /= (STARTUP_SHORT);
        }
        s->smooth[i] = h[i];
        re[i] *= s->smooth[i];
        im[i] *= s->smooth[i];
    }
    memcpy(s->magn_prev_pr, mag, sizeof(float) * (size_t)nb);
    memcpy(s->noise_prev, s->noise, sizeof(float) * (size_t)nb);

    t[0] = re[0];
    t[1] = re[nb - 1];
    for (i = 1; i < nb - 1; ++i) {
        t[2 * i] = re[i];
        t[2 * i + 1] = im[i];
    }
    orc_rdft(s->ana, -1, t, s->ip, s->wfft);
    for (i = 0; i < s->ana; ++i)
        t[i] *= 2.f / s->ana;

    factor = 1.f;
    if (s->gainmap == 1 && s->frame_idx > STARTUP_LONG) {
        f1 = 1.f;
        f2 = 1.f;
        e2 = orc_energy_f(t, s->ana);
        gain = (float)sqrt(e2 / (e1 + 1.f));
        if (gain > 0.5f) {
            f1 = 1.f + 1.3f * (gain - 0.5f);
            if (gain * f1 > 1.f)
                f1 = 1.f / gain;
        }
        if (gain < 0.5f) {
            if (gain <= s->floor_gain)
                gain = s->floor_gain;
            f2 = 1.f - 0.3f * (0.5f - gain);
        }
        factor = s->prior_prob * f1 + (1.f - s->prior_prob) * f2;
    }
    for (i = 0; i < s->ana; ++i)
        t[i] = s->window[i] * t[i];
    for (i = 0; i < s->ana; ++i)
        s->synth[i] += factor * t[i];
    for (i = 0; i < s->block; ++i)
        fout[i] = s->synth[i];
    orc_slide_in(s->synth, s->ana, s->block, NULL);
    for (i = 0; i < s->block; ++i)
        out[i] = fout[i] > 32767 ? 32767 : (fout[i] < -32768 ? -32768 : fout[i]);
    if (out_hb) {
        const int delta = nb / 4;
        float p_hb = 0.f, g_hb = 0.f, sa = 0.f, sp = 0.f, mod, g;
        for (i = nb - delta - 1; i < nb - 1; ++i)
            p_hb += s->speech_prob[i];
        p_hb = p_hb / ((float)delta);
        for (i = 0; i < nb; ++i) {
            sa += s->magn_prev_an[i];
            sp += s->magn_prev_pr[i];
        }
        p_hb *= sp / sa;
        for (i = nb - delta - 1; i < nb - 1; ++i)
            g_hb += s->smooth[i];
        g_hb = g_hb / ((float)delta);
        mod = 0.5f * (1.f + (float)tanh(1.0f * (2.f * p_hb - 1.f)));
        g = 0.5f * mod + 0.5f * g_hb;
        if (p_hb >= 0.5f)
            g = 0.25f * mod + 0.75f * g_hb;
        g = g * 1.0f;
        if (g < s->floor_gain)
            g = s->floor_gain;
        if (g > 1.f)
            g = 1.f;
        for (i = 0; i < s->block; ++i) {
            float v = g * s->inbuf_hb[i];
            out_hb[i] = v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
        }
    }
}

/* ---- wmix handle layer: R:src/webrtc.c:560-660 ---- */

orc_ns *orc_ns_init(int chn, int freq)
{
    orc_ns *h;
    if (freq > 32000 || freq % 8000 != 0)
        return NULL;
    if (chn != 1 && chn != 2)
        return NULL;
    h = (orc_ns *)calloc(1, sizeof(*h));
    h->core = (orc_ns_core *)malloc(sizeof(orc_ns_core));
    if (orc_ns_core_init(h->core, freq, 2) != 0) {
        free(h->core);
        free(h);
        return NULL;
    }
    h->chn = chn;
    h->freq = freq;
    h->pkg = freq / 1000 * 10;
    return h;
}

void orc_ns_process(orc_ns *h, const int16_t *in, int16_t *out, int frame_num)
{
    float fin[160], fo[160], fin_hb[160], fo_hb[160];
    int pos, i;
    const int chn = h->chn;
    orc_ns_core *s = h->core;
    /* At 32 kHz the packet is 320 samples but the core still works on 160-sample blocks and wmix passes ONE band
     * (R:src/webrtc.c:633), so only the first 160 samples of a packet are analysed and written; the rest of the
     * reference's calloc'ed out[0] is never touched and reads back as zero.  With two channels the interleaved
     * right channel is handed over as a SECOND BAND (num_bands = chn, R:src/webrtc.c:624-636). */
    for (pos = 0; pos + h->pkg <= frame_num; pos += h->pkg) {
        for (i = 0; i < s->block; ++i) {
            fin[i] = (float)in[(pos + i) * chn];
            if (chn == 2)
                fin_hb[i] = (float)in[(pos + i) * chn + 1];
        }
        orc_slide_in(s->inbuf, s->ana, s->block, fin);
        orc_ns_analyze(s, fin);
        orc_ns_synthesize(s, fo, chn == 2 ? fin_hb : NULL, chn == 2 ? fo_hb : NULL);
        for (i = 0; i < h->pkg; ++i) {
            out[(pos + i) * chn] = i < s->block ? (int16_t)fo[i] : 0;
            if (chn == 2)
                out[(pos + i) * chn + 1] = i < s->block ? (int16_t)fo_hb[i] : 0;
        }
    }
}

void orc_ns_release(orc_ns *h)
{
    if (!h)
        return;
    free(h->core);
    free(h);
}

int orc_ns_block_index(const orc_ns *h) { return h->core->frame_idx; }
const float *orc_ns_prior_model(const orc_ns *h) { return h->core->prior_model; }
