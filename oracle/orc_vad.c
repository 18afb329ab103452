/* ORACLE (test infrastructure) — WebRTC VAD restated from T:webrtc/common_audio/vad
 * (vad_sp.c, vad_filterbank.c, vad_gmm.c, vad_core.c, webrtc_vad.c) and the wmix handle
 * layer R:src/webrtc.c:40-167.  Integer only; int16 narrowing wraps, >> on negatives is
 * arithmetic, as in the reference on gcc/x86. */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define NCH 6

/* model tables: T:.../vad/vad_core.c:22-63 (numeric constants of the trained GMM) */
static const int16_t kSpecW[NCH] = {6, 8, 10, 12, 14, 16};
static const int16_t kMinDiff[NCH] = {544, 544, 576, 576, 576, 576};
static const int16_t kMaxSpeech[NCH] = {11392, 11392, 11520, 11520, 11520, 11520};
static const int16_t kMinMean[2] = {640, 768};
static const int16_t kMaxNoise[NCH] = {9216, 9088, 8960, 8832, 8704, 8576};
static const int16_t kNoiseW[12] = {34, 62, 72, 66, 53, 25, 94, 66, 56, 62, 75, 103};
static const int16_t kSpeechW[12] = {48, 82, 45, 87, 50, 47, 80, 46, 83, 41, 78, 81};
static const int16_t kNoiseMean0[12] = {6738, 4892, 7065, 6715, 6771, 3369,
                                        7646, 3863, 7820, 7266, 5020, 4362};
static const int16_t kSpeechMean0[12] = {8306, 10085, 10078, 11823, 11843, 6309,
                                         9473, 9571, 10879, 7581, 8180, 7483};
static const int16_t kNoiseStd0[12] = {378, 1064, 493, 582, 688, 593, 474, 697, 475, 688, 421, 455};
static const int16_t kSpeechStd0[12] = {555, 505, 567, 524, 585, 1231, 509, 828, 492, 1540, 1079, 850};
/* per-mode thresholds for 10/20/30 ms: T:.../vad/vad_core.c:78-100 */
static const int16_t kOh1[4][3] = {{8, 4, 3}, {8, 4, 3}, {6, 3, 2}, {6, 3, 2}};
static const int16_t kOh2[4][3] = {{14, 7, 5}, {14, 7, 5}, {9, 5, 3}, {9, 5, 3}};
static const int16_t kLocal[4][3] = {{24, 21, 24}, {37, 32, 37}, {82, 78, 82}, {94, 94, 94}};
static const int16_t kGlobal[4][3] = {{57, 48, 57}, {100, 80, 100}, {285, 260, 285}, {1100, 1050, 1100}};

/* T:.../vad/vad_core.c:483-530 (InitCore) + :533-582 (set_mode_core) */
void orc_vad_core_init(orc_vad_core *v, int mode)
{
    int i;
    memset(v, 0, sizeof(*v));
    v->vad = 1;
    for (i = 0; i < 12; ++i) {
        v->noise_means[i] = kNoiseMean0[i];
        v->speech_means[i] = kSpeechMean0[i];
        v->noise_stds[i] = kNoiseStd0[i];
        v->speech_stds[i] = kSpeechStd0[i];
    }
    for (i = 0; i < 96; ++i) {
        v->low_value[i] = 10000;
        v->age[i] = 0;
    }
    for (i = 0; i < NCH; ++i)
        v->mean_value[i] = 1600;
    for (i = 0; i < 3; ++i) {
        v->over_hang_max_1[i] = kOh1[mode][i];
        v->over_hang_max_2[i] = kOh2[mode][i];
        v->individual[i] = kLocal[mode][i];
        v->total[i] = kGlobal[mode][i];
    }
}

/* T:.../vad/vad_sp.c:27-54 — two first-order all-pass branches (Q13 5243 / 1392), summed */
void orc_vad_downsample(const int16_t *in, int16_t *out, int32_t st[2], int in_len)
{
    int32_t s0 = st[0], s1 = st[1];
    int n, half = in_len >> 1;
    for (n = 0; n < half; ++n) {
        int16_t a = (int16_t)((s0 >> 1) + ((5243 * in[2 * n]) >> 14));
        int16_t b;
        s0 = (int32_t)in[2 * n] - ((5243 * a) >> 12);
        b = (int16_t)((s1 >> 1) + ((1392 * in[2 * n + 1]) >> 14));
        out[n] = (int16_t)(a + b);
        s1 = (int32_t)in[2 * n + 1] - ((1392 * b) >> 12);
    }
    st[0] = s0;
    st[1] = s1;
}

/* T:.../vad/vad_filterbank.c:83-108 — first-order all-pass on every 2nd sample */
static void orc_allpass(const int16_t *in, int n, int16_t coef, int16_t *state, int16_t *out)
{
    int32_t s = (int32_t)((uint32_t)(int32_t)*state << 16);
    int i;
    for (i = 0; i < n; ++i) {
        int32_t acc = (int32_t)((uint32_t)s + (uint32_t)((int32_t)coef * in[2 * i]));
        int16_t y = (int16_t)(acc >> 16);
        out[i] = y;
        s = (int32_t)((uint32_t)(int32_t)in[2 * i] << 14);
        s = (int32_t)((uint32_t)s - (uint32_t)((int32_t)coef * y));
        s = (int32_t)((uint32_t)s << 1);
    }
    *state = (int16_t)(s >> 16);
}

/* T:.../vad/vad_filterbank.c:121-140 */
static void orc_split(const int16_t *in, int n, int16_t *up_st, int16_t *lo_st,
                      int16_t *hp, int16_t *lp)
{
    int half = n >> 1, i;
    orc_allpass(in, half, 20972, up_st, hp);
    orc_allpass(in + 1, half, 5571, lo_st, lp);
    for (i = 0; i < half; ++i) {
        int16_t u = hp[i];
        hp[i] = (int16_t)(hp[i] - lp[i]);
        lp[i] = (int16_t)(lp[i] + u);
    }
}

/* T:.../vad/vad_filterbank.c:41-72 — 2nd-order HP at ~80 Hz, Q14 coefficients */
static void orc_highpass(const int16_t *in, int n, int16_t st[4], int16_t *out)
{
    int i;
    for (i = 0; i < n; ++i) {
        int32_t acc = 6631 * in[i];
        acc += -13262 * st[0];
        acc += 6631 * st[1];
        st[1] = st[0];
        st[0] = in[i];
        acc -= -7756 * st[2];
        acc -= 5620 * st[3];
        st[3] = st[2];
        st[2] = (int16_t)(acc >> 14);
        out[i] = st[2];
    }
}

/* T:.../vad/vad_filterbank.c:155-236 — 10*log10(energy) in Q4 via norm + linear mantissa */
static void orc_log_energy(const int16_t *in, int n, int16_t offset, int16_t *total, int16_t *out)
{
    int rshifts = 0;
    uint32_t e = (uint32_t)orc_energy(in, n, &rshifts);
    if (e == 0) {
        *out = offset;
        return;
    }
    {
        int norm = 17 - orc_norm_u32(e);
        int16_t log2e = 14336;
        int16_t r;
        rshifts += norm;
        if (norm < 0)
            e <<= -norm;
        else
            e >>= norm;
        log2e = (int16_t)(log2e + (int16_t)((e & 0x3FFF) >> 4));
        r = (int16_t)(((24660 * log2e) >> 19) + ((rshifts * 24660) >> 9));
        if (r < 0)
            r = 0;
        *out = (int16_t)(r + offset);
    }
    if (*total <= 10) {
        if (rshifts >= 0)
            *total = (int16_t)(*total + 11);
        else
            *total = (int16_t)(*total + (int16_t)(e >> -rshifts));
    }
}

/* T:.../vad/vad_filterbank.c:246-333 — 5 split stages -> 6 band log-energies */
int16_t orc_vad_features(orc_vad_core *v, const int16_t *in, int len, int16_t feat[6])
{
    static const int16_t off[6] = {368, 368, 272, 176, 176, 176};
    int16_t total = 0;
    int16_t hpA[120], lpA[120], hpB[60], lpB[60];
    int half = len >> 1, n;

    orc_split(in, len, &v->upper_state[0], &v->lower_state[0], hpA, lpA);     /* 0-2k | 2-4k */
    orc_split(hpA, half, &v->upper_state[1], &v->lower_state[1], hpB, lpB);   /* 2-3k | 3-4k */
    n = half >> 1;
    orc_log_energy(hpB, n, off[5], &total, &feat[5]);
    orc_log_energy(lpB, n, off[4], &total, &feat[4]);
    orc_split(lpA, half, &v->upper_state[2], &v->lower_state[2], hpB, lpB);   /* 0-1k | 1-2k */
    orc_log_energy(hpB, n, off[3], &total, &feat[3]);
    orc_split(lpB, n, &v->upper_state[3], &v->lower_state[3], hpA, lpA);      /* 0-500 | 500-1k */
    n >>= 1;
    orc_log_energy(hpA, n, off[2], &total, &feat[2]);
    orc_split(lpA, n, &v->upper_state[4], &v->lower_state[4], hpB, lpB);      /* 0-250 | 250-500 */
    n >>= 1;
    orc_log_energy(hpB, n, off[1], &total, &feat[1]);
    orc_highpass(lpB, n, v->hp_state, hpA);                                   /* 80-250 */
    orc_log_energy(hpA, n, off[0], &total, &feat[0]);
    return total;
}

/* T:.../vad/vad_gmm.c:30-83 */
int32_t orc_vad_gaussian(int16_t input, int16_t mean, int16_t std, int16_t *delta)
{
    int16_t t16, inv_std, inv_std2, expv = 0;
    int32_t t32;
    t32 = 131072 + (int32_t)(std >> 1);
    inv_std = (int16_t)orc_div_w32_w16(t32, std);
    t16 = (int16_t)(inv_std >> 2);
    inv_std2 = (int16_t)((t16 * t16) >> 2);
    t16 = (int16_t)(input << 3);
    t16 = (int16_t)(t16 - mean);
    *delta = (int16_t)((inv_std2 * t16) >> 10);
    t32 = (*delta * t16) >> 9;
    if (t32 < 22005) {
        t16 = (int16_t)((5909 * t32) >> 12);
        t16 = (int16_t)-t16;
        expv = (int16_t)(0x0400 | (t16 & 0x03FF));
        t16 ^= (int16_t)0xFFFF;
        t16 >>= 10;
        t16 += 1;
        expv >>= t16;
    }
    return inv_std * expv;
}

/* T:.../vad/vad_sp.c:59-177 — 16 smallest of the last 100 frames + smoothed median */
int16_t orc_vad_find_minimum(orc_vad_core *v, int16_t feature, int channel)
{
    int16_t *age = &v->age[channel * 16];
    int16_t *low = &v->low_value[channel * 16];
    int i, j, pos = -1;
    int16_t median = 1600, alpha = 0;
    int32_t acc;

    for (i = 0; i < 16; ++i) {
        if (age[i] != 100) {
            age[i]++;
        } else {
            /* drop entry i; the reference's copy loop reads one past the end for j == 15
             * (T:.../vad/vad_sp.c:83-86) and then overwrites slot 15, so only j < 15 matters */
            for (j = i; j < 15; ++j) {
                low[j] = low[j + 1];
                age[j] = age[j + 1];
            }
            age[15] = 101;
            low[15] = 10000;
        }
    }
    /* the reference walks a fixed comparison tree (T:.../vad/vad_sp.c:93-146); on the sorted
     * list it keeps, that tree returns the first slot whose value exceeds the feature */
    if (feature < low[7]) {
        if (feature < low[3]) {
            if (feature < low[1])
                pos = (feature < low[0]) ? 0 : 1;
            else
                pos = (feature < low[2]) ? 2 : 3;
        } else if (feature < low[5]) {
            pos = (feature < low[4]) ? 4 : 5;
        } else {
            pos = (feature < low[6]) ? 6 : 7;
        }
    } else if (feature < low[15]) {
        if (feature < low[11]) {
            if (feature < low[9])
                pos = (feature < low[8]) ? 8 : 9;
            else
                pos = (feature < low[10]) ? 10 : 11;
        } else if (feature < low[13]) {
            pos = (feature < low[12]) ? 12 : 13;
        } else {
            pos = (feature < low[14]) ? 14 : 15;
        }
    }
    if (pos > -1) {
        for (i = 15; i > pos; --i) {
            low[i] = low[i - 1];
            age[i] = age[i - 1];
        }
        low[pos] = feature;
        age[pos] = 1;
    }
    if (v->frame_counter > 2)
        median = low[2];
    else if (v->frame_counter > 0)
        median = low[0];
    if (v->frame_counter > 0)
        alpha = (median < v->mean_value[channel]) ? 6553 : 32439;
    acc = (alpha + 1) * v->mean_value[channel];
    acc += (32767 - alpha) * median;
    acc += 16384;
    v->mean_value[channel] = (int16_t)(acc >> 15);
    return v->mean_value[channel];
}

/* weighted mean of the two Gaussians of one band, after shifting both means by `shift`
 * (T:.../vad/vad_core.c:110-122) */
static int32_t orc_shift_and_average(int16_t *means, int16_t shift, const int16_t *w)
{
    int32_t acc = 0;
    int k;
    for (k = 0; k < 2; ++k) {
        means[k * NCH] = (int16_t)(means[k * NCH] + shift);
        acc += means[k * NCH] * w[k * NCH];
    }
    return acc;
}

/* T:.../vad/vad_core.c:124-479 */
static int16_t orc_vad_gmm(orc_vad_core *v, const int16_t *feat, int16_t total_power, int len)
{
    int ch, k, col;
    int16_t vadflag = 0;
    int16_t dN[12], dS[12], pN[12], pS[12];
    int32_t sum_llr = 0;
    int16_t oh1, oh2, thr_local, thr_global;

    col = (len == 80) ? 0 : (len == 160 ? 1 : 2);
    oh1 = v->over_hang_max_1[col];
    oh2 = v->over_hang_max_2[col];
    thr_local = v->individual[col];
    thr_global = v->total[col];
    memset(pN, 0, sizeof(pN));
    memset(pS, 0, sizeof(pS));

    if (total_power > 10) {
        int16_t maxspe = 12800;
        for (ch = 0; ch < NCH; ++ch) {
            int32_t h0 = 0, h1 = 0, probN[2], probS[2];
            int16_t sh0, sh1, llr, q;
            for (k = 0; k < 2; ++k) {
                int g = ch + k * NCH;
                probN[k] = kNoiseW[g] * orc_vad_gaussian(feat[ch], v->noise_means[g], v->noise_stds[g], &dN[g]);
                h0 += probN[k];
                probS[k] = kSpeechW[g] * orc_vad_gaussian(feat[ch], v->speech_means[g], v->speech_stds[g], &dS[g]);
                h1 += probS[k];
            }
            sh0 = h0 ? orc_norm_w32(h0) : 31;
            sh1 = h1 ? orc_norm_w32(h1) : 31;
            llr = (int16_t)(sh0 - sh1);
            sum_llr += (int32_t)(llr * kSpecW[ch]);
            if ((llr << 2) > thr_local)
                vadflag = 1;
            q = (int16_t)(h0 >> 12);
            if (q > 0) {
                int32_t t = (int32_t)(((uint32_t)probN[0] & 0xFFFFF000u) << 2);
                pN[ch] = (int16_t)orc_div_w32_w16(t, q);
                pN[ch + NCH] = (int16_t)(16384 - pN[ch]);
            } else {
                pN[ch] = 16384;
            }
            q = (int16_t)(h1 >> 12);
            if (q > 0) {
                int32_t t = (int32_t)(((uint32_t)probS[0] & 0xFFFFF000u) << 2);
                pS[ch] = (int16_t)orc_div_w32_w16(t, q);
                pS[ch + NCH] = (int16_t)(16384 - pS[ch]);
            }
        }
        vadflag |= (sum_llr >= thr_global);

        for (ch = 0; ch < NCH; ++ch) {
            int16_t fmin = orc_vad_find_minimum(v, feat[ch], ch);
            int32_t ngm = orc_shift_and_average(&v->noise_means[ch], 0, &kNoiseW[ch]);
            int32_t sgm;
            int16_t ngm_q8 = (int16_t)(ngm >> 6);
            int16_t diff, t16;
            for (k = 0; k < 2; ++k) {
                int g = ch + k * NCH;
                int16_t nmk = v->noise_means[g], smk = v->speech_means[g];
                int16_t nsk = v->noise_stds[g], ssk = v->speech_stds[g];
                int16_t nmk2 = nmk, nmk3, d, nd, lim;
                int32_t a32, b32;
                if (!vadflag) {
                    d = (int16_t)((pN[g] * dN[g]) >> 11);
                    nmk2 = (int16_t)(nmk + (int16_t)((d * 655) >> 22));
                }
                nd = (int16_t)((fmin << 4) - ngm_q8);
                nmk3 = (int16_t)(nmk2 + (int16_t)((nd * 154) >> 9));
                lim = (int16_t)((k + 5) << 7);
                if (nmk3 < lim)
                    nmk3 = lim;
                lim = (int16_t)((72 + k - ch) << 7);
                if (nmk3 > lim)
                    nmk3 = lim;
                v->noise_means[g] = nmk3;

                if (vadflag) {
                    int16_t smk2, maxmu;
                    d = (int16_t)((pS[g] * dS[g]) >> 11);
                    t16 = (int16_t)((d * 6554) >> 21);
                    smk2 = (int16_t)(smk + ((t16 + 1) >> 1));
                    maxmu = (int16_t)(maxspe + 640);
                    if (smk2 < kMinMean[k])
                        smk2 = kMinMean[k];
                    if (smk2 > maxmu)
                        smk2 = maxmu;
                    v->speech_means[g] = smk2;
                    t16 = (int16_t)((smk + 4) >> 3);
                    t16 = (int16_t)(feat[ch] - t16);
                    a32 = (dS[g] * t16) >> 3;
                    b32 = a32 - 4096;
                    t16 = (int16_t)(pS[g] >> 2);
                    a32 = (int32_t)((uint32_t)(int32_t)t16 * (uint32_t)b32);
                    b32 = a32 >> 4;
                    if (b32 > 0)
                        t16 = (int16_t)orc_div_w32_w16(b32, (int16_t)(ssk * 10));
                    else {
                        t16 = (int16_t)orc_div_w32_w16(-b32, (int16_t)(ssk * 10));
                        t16 = (int16_t)-t16;
                    }
                    t16 = (int16_t)(t16 + 128);
                    ssk = (int16_t)(ssk + (t16 >> 8));
                    if (ssk < 384)
                        ssk = 384;
                    v->speech_stds[g] = ssk;
                } else {
                    t16 = (int16_t)(feat[ch] - (nmk >> 3));
                    a32 = (dN[g] * t16) >> 3;
                    a32 -= 4096;
                    t16 = (int16_t)((pN[g] + 2) >> 2);
                    b32 = (int32_t)((uint32_t)(int32_t)t16 * (uint32_t)a32);
                    a32 = b32 >> 14;
                    if (a32 > 0)
                        t16 = (int16_t)orc_div_w32_w16(a32, nsk);
                    else {
                        t16 = (int16_t)orc_div_w32_w16(-a32, nsk);
                        t16 = (int16_t)-t16;
                    }
                    t16 = (int16_t)(t16 + 32);
                    nsk = (int16_t)(nsk + (t16 >> 6));
                    if (nsk < 384)
                        nsk = 384;
                    v->noise_stds[g] = nsk;
                }
            }
            /* keep the speech and noise models apart (T:.../vad/vad_core.c:406-438) */
            ngm = orc_shift_and_average(&v->noise_means[ch], 0, &kNoiseW[ch]);
            sgm = orc_shift_and_average(&v->speech_means[ch], 0, &kSpeechW[ch]);
            diff = (int16_t)((int16_t)(sgm >> 9) - (int16_t)(ngm >> 9));
            if (diff < kMinDiff[ch]) {
                int16_t gap = (int16_t)(kMinDiff[ch] - diff);
                int16_t up = (int16_t)((13 * gap) >> 2);
                int16_t dn = (int16_t)((3 * gap) >> 2);
                sgm = orc_shift_and_average(&v->speech_means[ch], up, &kSpeechW[ch]);
                ngm = orc_shift_and_average(&v->noise_means[ch], (int16_t)-dn, &kNoiseW[ch]);
            }
            maxspe = kMaxSpeech[ch];
            t16 = (int16_t)(sgm >> 7);
            if (t16 > maxspe) {
                t16 = (int16_t)(t16 - maxspe);
                for (k = 0; k < 2; ++k)
                    v->speech_means[ch + k * NCH] = (int16_t)(v->speech_means[ch + k * NCH] - t16);
            }
            t16 = (int16_t)(ngm >> 7);
            if (t16 > kMaxNoise[ch]) {
                t16 = (int16_t)(t16 - kMaxNoise[ch]);
                for (k = 0; k < 2; ++k)
                    v->noise_means[ch + k * NCH] = (int16_t)(v->noise_means[ch + k * NCH] - t16);
            }
        }
        v->frame_counter++;
    }
    /* hang-over (T:.../vad/vad_core.c:462-477) */
    if (!vadflag) {
        if (v->over_hang > 0) {
            vadflag = (int16_t)(2 + v->over_hang);
            v->over_hang--;
        }
        v->num_of_speech = 0;
    } else {
        v->num_of_speech++;
        if (v->num_of_speech > 6) {
            v->num_of_speech = 6;
            v->over_hang = oh2;
        } else {
            v->over_hang = oh1;
        }
    }
    return vadflag;
}

/* T:.../vad/webrtc_vad.c:71-105 + :107-130, T:.../vad/vad_core.c:612-674.  8/16/32 kHz. */
int orc_vad_core_process(orc_vad_core *v, int fs, const int16_t *frame, int len)
{
    int16_t nb[240], wb[480], feat[6], power;
    const int16_t *x = frame;
    int ms, ok = 0;
    if (!frame)
        return -1;
    if (fs != 8000 && fs != 16000 && fs != 32000)
        return -1;
    for (ms = 10; ms <= 30; ms += 10)
        if (len == fs / 1000 * ms)
            ok = 1;
    if (!ok)
        return -1;
    if (fs == 32000) {
        orc_vad_downsample(frame, wb, &v->ds_state[2], len);
        len /= 2;
        orc_vad_downsample(wb, nb, &v->ds_state[0], len);
        len /= 2;
        x = nb;
    } else if (fs == 16000) {
        orc_vad_downsample(frame, nb, &v->ds_state[0], len);
        len /= 2;
        x = nb;
    }
    power = orc_vad_features(v, x, len, feat);
    v->vad = orc_vad_gmm(v, feat, power, len);
    return v->vad > 0 ? 1 : v->vad;
}

/* ---- wmix handle layer ---- */

/* R:src/webrtc.c:40-81 */
orc_vad *orc_vad_init(int chn, int freq, int interval_ms)
{
    orc_vad *h;
    if (freq > 32000 || freq % 8000 != 0)
        return NULL;
    h = (orc_vad *)calloc(1, sizeof(*h));
    orc_vad_core_init(&h->core, 3);
    h->chn = chn;
    h->freq = freq;
    if (freq <= 16000)
        h->interval_ms = (interval_ms % 20 == 0) ? 20 : 10;
    else
        h->interval_ms = 10;
    h->pkg = freq / 1000 * h->interval_ms;
    h->reduce = 4;
    return h;
}

/* R:src/webrtc.c:91-151.  Note the two quirks SURVEY.md §8 a5 records: the detector is always
 * fed the *start* of the buffer and the attenuation loop runs from cLen to pkg, so only the
 * first packet of a call is ever attenuated (fine for one packet per call). */
void orc_vad_process(orc_vad *h, int16_t *frame, int frame_num)
{
    int total = frame_num * h->chn, mono = total, pos, c, i, r;
    if (h->chn > 1) {
        for (pos = mono = 0; pos < total;) {
            int32_t s = 0;
            for (c = 0; c < h->chn; ++c)
                s += frame[pos++];
            frame[mono++] = (int16_t)(s / h->chn);
        }
    }
    for (pos = 0; pos < mono; pos += h->pkg) {
        r = orc_vad_core_process(&h->core, h->freq, frame, h->pkg);
        if (r < 0)
            return;
        if (r == 0) {
            if (h->reduce < 4)
                h->reduce++;
        } else if (h->reduce > 0) {
            h->reduce--;
        }
        for (i = pos; i < h->pkg; ++i)
            frame[i] = (int16_t)(frame[i] >> h->reduce);
    }
    if (h->chn > 1) {
        int m = mono - 1;
        for (pos = total - 1; pos >= 0; --m)
            for (c = 0; c < h->chn; ++c)
                frame[pos--] = frame[m];
    }
}

void orc_vad_release(orc_vad *h) { free(h); }
