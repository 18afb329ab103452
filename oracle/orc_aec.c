/* ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.h).
 *
 * Plain-C restatement of the WebRTC float echo canceller exactly as wmix configures and drives it:
 *   wrapper      R:src/webrtc.c:217-500           aec_init / aec_setFrameFar / aec_process / aec_process2
 *   API layer    T:webrtc/modules/audio_processing/aec/echo_cancellation.c
 *                  Init :196-276, BufferFarend :278-339, Process :341-409, ProcessNormal :599-747,
 *                  EstBufDelayNormal :821-872
 *   core         T:.../aec/aec_core.c   FilterFar :148, ScaleErrorSignal :172, FilterAdaptation :222,
 *                  OverdriveAndSuppress :272, PartitionDelay :295, SmoothedPSD :333, SubbandCoherence :412,
 *                  ComfortNoise :462, TimeToFrequency :831, NonLinearProcessing :911, ProcessBlock :1143,
 *                  InitAec :1509-1688, BufferFarendPartition :1690, MoveFarReadPtr :1709, ProcessFrames :1719
 *   transform    T:.../aec/aec_rdft.c:126-557 (fixed 128-point Ooura rdft; same butterflies as fft4g.c, own tables)
 *   ring         T:webrtc/common_audio/ring_buffer.c:112-247
 *   noise LCG    T:webrtc/common_audio/signal_processing/randomization_functions.c:98-118
 * Fixed by wmix: nlpMode = aggressive (2), skew / metrics / delay logging off, reported-delay mode on,
 * normal (12-partition) filter, one band, plain-C kernels.  Branches those settings make unreachable
 * (resampler, delay estimator, extended filter, metrics, high bands) are not restated.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

enum { PART = 64, PART1 = 65, PART2 = 128, FRAME = 80, NPART = 12, FAR_RING = 250, PRE_RING = 128 + 4 * FRAME };

/* ---------------------------------------------------------------- tables */
static float g_w[64];          /* rdft_w            aec_rdft.c:32-49   */
static float g_hann[65];       /* sqrt-Hanning      aec_core.c:49-66   */
static float g_weight[65];     /* weightCurve       aec_core.c:71-81   */
static float g_over[65];       /* overDriveCurve    aec_core.c:86-96   */
static int g_ip[34];
static int g_tables_ready;

void orc_aec_tables(float *w64, float *hann65, float *weight65, float *over65)
{
    if (!g_tables_ready) {
        /* makewt/makect of a 128-point transform, correctly rounded ... */
        int ip[34] = {0, 0};
        float a[128] = {0};
        orc_rdft(128, 1, a, ip, g_w);               /* first call fills ip / w exactly as fft4g.c:324-340 */
        /* ... except that the literal table the AEC ships was produced by a libm whose cosf/sinf were
         * off by one ulp in eight places (found by comparing against the reference's exported
         * `rdft_w`; tests/test_oracle_pin.py re-checks it): */
        static const struct { int idx, ulps; } fix[8] = {{4, 1}, {7, 1}, {20, 1}, {27, 1}, {40, 1}, {41, -1}, {42, 1}, {47, 1}};
        for (int k = 0; k < 8; ++k) {
            int32_t bits;
            memcpy(&bits, &g_w[fix[k].idx], 4);
            bits += fix[k].ulps;
            memcpy(&g_w[fix[k].idx], &bits, 4);
        }
        g_ip[0] = 32;
        g_ip[1] = 32;
        memcpy(g_ip + 2, ip + 2, 32 * sizeof(int));
        for (int i = 0; i <= 64; ++i) {
            g_hann[i] = (float)sin(3.14159265358979323846 * i / 128.0);
            /* 4-decimal Matlab prints: 0.3*sqrt(linspace(0,1,64))+0.1 with a leading 0, and sqrt(linspace(0,1,65))+1 */
            g_weight[i] = i == 0 ? 0.f : (float)(floor((0.3 * sqrt((i - 1) / 63.0) + 0.1) * 1e4 + 0.5) / 1e4);
            g_over[i] = (float)(floor((sqrt(i / 64.0) + 1.0) * 1e4 + 0.5) / 1e4);
        }
        g_tables_ready = 1;
    }
    if (w64) memcpy(w64, g_w, sizeof g_w);
    if (hann65) memcpy(hann65, g_hann, sizeof g_hann);
    if (weight65) memcpy(weight65, g_weight, sizeof g_weight);
    if (over65) memcpy(over65, g_over, sizeof g_over);
}

/* aec_rdft_forward_128 / aec_rdft_inverse_128 (aec_rdft.c:539-557) */
void orc_aec_rdft(float *a, int inverse)
{
    orc_aec_tables(NULL, NULL, NULL, NULL);
    orc_rdft(128, inverse ? -1 : 1, a, g_ip, g_w);
}

/* ---------------------------------------------------------------- element ring (ring_buffer.c) */
typedef struct {
    size_t rd, wr, count, esize;
    int diff_wrap;
    char *data;
} ring_t;

static ring_t *ring_new(size_t count, size_t esize)
{
    ring_t *r = calloc(1, sizeof *r);
    r->count = count;
    r->esize = esize;
    r->data = calloc(count, esize);
    return r;
}
static void ring_free(ring_t *r)
{
    if (r) { free(r->data); free(r); }
}
static size_t ring_avail_read(const ring_t *r) { return r->diff_wrap ? r->count - r->rd + r->wr : r->wr - r->rd; }
static size_t ring_avail_write(const ring_t *r) { return r->count - ring_avail_read(r); }
/* WebRtc_MoveReadPtr :192-224 — note the `>` (not `>=`) wrap test */
static int ring_move(ring_t *r, int n)
{
    const int fr = (int)ring_avail_write(r), rd = (int)ring_avail_read(r);
    int pos = (int)r->rd;
    if (n > rd) n = rd;
    if (n < -fr) n = -fr;
    pos += n;
    if (pos > (int)r->count) { pos -= (int)r->count; r->diff_wrap = 0; }
    if (pos < 0) { pos += (int)r->count; r->diff_wrap = 1; }
    r->rd = (size_t)pos;
    return n;
}
/* WebRtc_ReadBuffer :112-156, always copying */
static size_t ring_read(ring_t *r, void *dst, size_t n)
{
    const size_t have = ring_avail_read(r), take = have < n ? have : n, margin = r->count - r->rd;
    if (take > margin) {
        memcpy(dst, r->data + r->rd * r->esize, margin * r->esize);
        memcpy((char *)dst + margin * r->esize, r->data, (take - margin) * r->esize);
    } else {
        memcpy(dst, r->data + r->rd * r->esize, take * r->esize);
    }
    ring_move(r, (int)take);
    return take;
}
/* WebRtc_WriteBuffer :158-190 */
static size_t ring_write(ring_t *r, const void *src, size_t n)
{
    const size_t fr = ring_avail_write(r), put = fr < n ? fr : n, margin = r->count - r->wr;
    size_t left = put;
    if (put > margin) {
        memcpy(r->data + r->wr * r->esize, src, margin * r->esize);
        r->wr = 0;
        left -= margin;
        r->diff_wrap = 1;
    }
    memcpy(r->data + r->wr * r->esize, (const char *)src + (put - left) * r->esize, left * r->esize);
    r->wr += left;
    return put;
}

/* ---------------------------------------------------------------- state */
typedef struct {
    /* core (aec_core_internal.h:52-169, the live subset) */
    int known_delay, delay_est_ctr, delay_idx, xf_pos, system_delay, mult, nlp_mode, noise_ctr;
    ring_t *near_fr, *out_fr, *far, *far_win;
    float d_buf[PART2], e_buf[PART2], out_buf[PART];
    float x_pow[PART1], d_pow[PART1], d_min[PART1], d_init_min[PART1];
    int noise_from_init;                              /* noisePow points at dInitMinPow (1) or dMinPow (0) */
    float xf[2][NPART * PART1], wf[2][NPART * PART1];
    float xfw[NPART][2][PART1];
    float sde[PART1][2], sxd[PART1][2], sx[PART1], sd[PART1], se[PART1];
    float hnl_fb_min, hnl_fb_local_min, hnl_xd_avg_min, over_drive, over_drive_sm, mu, err_thr;
    int hnl_new_min, hnl_min_ctr;
    short st_near_state, echo_state, diverge_state;
    uint32_t seed;
    /* API layer (echo_cancellation_internal.h:17-65, the live subset) */
    int rate_factor, buf_size_start, api_known_delay, time_for_delay_change, startup_phase, check_buff_size, sum;
    short counter, first_val, check_ctr, ms_in_snd, filt_delay, last_delay_diff;
    ring_t *far_pre;
} aec_t;

struct orc_aec {
    aec_t *a;
    int chn, freq, interval_ms, pkg;
    float *in, *out, *far;
};

/* ---------------------------------------------------------------- core pieces */
/* TimeToFrequency :831-856 */
static void time_to_freq(float t[PART2], float f[2][PART1], int window)
{
    if (window)
        for (int i = 0; i < PART; ++i) {
            t[i] *= g_hann[i];
            t[PART + i] *= g_hann[PART - i];
        }
    orc_aec_rdft(t, 0);
    f[1][0] = 0;
    f[1][PART] = 0;
    f[0][0] = t[0];
    f[0][PART] = t[1];
    for (int i = 1; i < PART; ++i) {
        f[0][i] = t[2 * i];
        f[1][i] = t[2 * i + 1];
    }
}

/* WebRtcAec_MoveFarReadPtr :1709-1717 */
static int move_far(aec_t *a, int n)
{
    const int moved = ring_move(a->far_win, n);
    ring_move(a->far, n);
    a->system_delay -= moved * PART;
    return moved;
}

/* WebRtcAec_BufferFarendPartition :1690-1707 */
static void buffer_far_partition(aec_t *a, const float *block)
{
    float t[PART2], xf[2][PART1];
    if (ring_avail_write(a->far) < 1) move_far(a, 1);
    memcpy(t, block, sizeof t);
    time_to_freq(t, xf, 0);
    ring_write(a->far, xf, 1);
    memcpy(t, block, sizeof t);
    time_to_freq(t, xf, 1);
    ring_write(a->far_win, xf, 1);
}

/* PartitionDelay :295-321 */
static int partition_delay(const aec_t *a)
{
    float best = 0;
    int delay = 0;
    for (int i = 0; i < NPART; ++i) {
        const int pos = i * PART1;
        float en = 0;
        for (int j = 0; j < PART1; ++j) en += a->wf[0][pos + j] * a->wf[0][pos + j] + a->wf[1][pos + j] * a->wf[1][pos + j];
        if (en > best) { best = en; delay = i; }
    }
    return delay;
}

static void window_block(float *dst, const float *x)
{
    for (int i = 0; i < PART; ++i) {
        dst[i] = x[i] * g_hann[i];
        dst[PART + i] = x[PART + i] * g_hann[PART - i];
    }
}
static void as_complex(const float *d, float c[2][PART1])
{
    c[0][0] = d[0];
    c[1][0] = 0;
    for (int i = 1; i < PART; ++i) { c[0][i] = d[2 * i]; c[1][i] = d[2 * i + 1]; }
    c[0][PART] = d[1];
    c[1][PART] = 0;
}

static int cmp_float(const void *pa, const void *pb)
{
    const float x = *(const float *)pa, y = *(const float *)pb;
    return (x > y) - (x < y);
}

/* NonLinearProcessing :911-1141 with SubbandCoherence :412-450, SmoothedPSD :333-395,
 * OverdriveAndSuppress :272-293 and ComfortNoise :462-546 folded in (one band) */
static void nlp(aec_t *a, float *output)
{
    static const float k_target_supp[3] = {-6.9f, -11.5f, -18.4f};
    static const float k_min_over[3] = {1.0f, 2.0f, 5.0f};
    static const float k_smooth[2][2] = {{0.9f, 0.1f}, {0.93f, 0.07f}};
    float efw[2][PART1], xfw[2][PART1], dfw[2][PART1], fft[PART2];
    float cohde[PART1], cohxd[PART1], hnl[PART1], pref[24];
    float hnl_de_avg, hnl_xd_avg, hnl_fb = 0, hnl_fb_low = 0;
    const int pref_size = 24 / a->mult, pref_min = 4 / a->mult, interval = 10 * a->mult;
    const float *g = k_smooth[a->mult - 1];
    int i;

    a->delay_est_ctr++;
    if (a->delay_est_ctr == interval) a->delay_est_ctr = 0;

    ring_read(a->far_win, xfw, 1);
    memcpy(a->xfw[0], xfw, sizeof xfw);

    /* SubbandCoherence: note delay_est_ctr was already advanced, so the filter-energy scan runs on
     * the block where the counter wraps to 0 */
    if (a->delay_est_ctr == 0) a->delay_idx = partition_delay(a);
    memcpy(xfw, a->xfw[a->delay_idx], sizeof xfw);
    window_block(fft, a->d_buf);
    orc_aec_rdft(fft, 0);
    as_complex(fft, dfw);
    window_block(fft, a->e_buf);
    orc_aec_rdft(fft, 0);
    as_complex(fft, efw);
    {   /* SmoothedPSD */
        float sd_sum = 0, se_sum = 0;
        for (i = 0; i < PART1; ++i) {
            a->sd[i] = g[0] * a->sd[i] + g[1] * (dfw[0][i] * dfw[0][i] + dfw[1][i] * dfw[1][i]);
            a->se[i] = g[0] * a->se[i] + g[1] * (efw[0][i] * efw[0][i] + efw[1][i] * efw[1][i]);
            {
                const float px = xfw[0][i] * xfw[0][i] + xfw[1][i] * xfw[1][i];
                a->sx[i] = g[0] * a->sx[i] + g[1] * (px > 15.f ? px : 15.f);
            }
            a->sde[i][0] = g[0] * a->sde[i][0] + g[1] * (dfw[0][i] * efw[0][i] + dfw[1][i] * efw[1][i]);
            a->sde[i][1] = g[0] * a->sde[i][1] + g[1] * (dfw[0][i] * efw[1][i] - dfw[1][i] * efw[0][i]);
            a->sxd[i][0] = g[0] * a->sxd[i][0] + g[1] * (dfw[0][i] * xfw[0][i] + dfw[1][i] * xfw[1][i]);
            a->sxd[i][1] = g[0] * a->sxd[i][1] + g[1] * (dfw[0][i] * xfw[1][i] - dfw[1][i] * xfw[0][i]);
            sd_sum += a->sd[i];
            se_sum += a->se[i];
        }
        a->diverge_state = (a->diverge_state ? 1.05f : 1.0f) * se_sum > sd_sum;
        if (a->diverge_state) memcpy(efw, dfw, sizeof efw);
        if (se_sum > (19.95f * sd_sum)) memset(a->wf, 0, sizeof a->wf);
    }
    for (i = 0; i < PART1; ++i) {
        cohde[i] = (a->sde[i][0] * a->sde[i][0] + a->sde[i][1] * a->sde[i][1]) / (a->sd[i] * a->se[i] + 1e-10f);
        cohxd[i] = (a->sxd[i][0] * a->sxd[i][0] + a->sxd[i][1] * a->sxd[i][1]) / (a->sx[i] * a->sd[i] + 1e-10f);
    }

    hnl_xd_avg = 0;
    for (i = pref_min; i < pref_size + pref_min; ++i) hnl_xd_avg += cohxd[i];
    hnl_xd_avg /= pref_size;
    hnl_xd_avg = 1 - hnl_xd_avg;
    hnl_de_avg = 0;
    for (i = pref_min; i < pref_size + pref_min; ++i) hnl_de_avg += cohde[i];
    hnl_de_avg /= pref_size;

    if (hnl_xd_avg < 0.75f && hnl_xd_avg < a->hnl_xd_avg_min) a->hnl_xd_avg_min = hnl_xd_avg;
    if (hnl_de_avg > 0.98f && hnl_xd_avg > 0.9f) a->st_near_state = 1;
    else if (hnl_de_avg < 0.95f || hnl_xd_avg < 0.8f) a->st_near_state = 0;

    if (a->hnl_xd_avg_min == 1) {
        a->echo_state = 0;
        a->over_drive = k_min_over[a->nlp_mode];
        if (a->st_near_state == 1) {
            memcpy(hnl, cohde, sizeof hnl);
            hnl_fb = hnl_de_avg;
            hnl_fb_low = hnl_de_avg;
        } else {
            for (i = 0; i < PART1; ++i) hnl[i] = 1 - cohxd[i];
            hnl_fb = hnl_xd_avg;
            hnl_fb_low = hnl_xd_avg;
        }
    } else if (a->st_near_state == 1) {
        a->echo_state = 0;
        memcpy(hnl, cohde, sizeof hnl);
        hnl_fb = hnl_de_avg;
        hnl_fb_low = hnl_de_avg;
    } else {
        a->echo_state = 1;
        for (i = 0; i < PART1; ++i) {
            const float alt = 1 - cohxd[i];
            hnl[i] = cohde[i] < alt ? cohde[i] : alt;
        }
        memcpy(pref, &hnl[pref_min], sizeof(float) * pref_size);
        qsort(pref, pref_size, sizeof(float), cmp_float);
        hnl_fb = pref[(int)floor(0.75f * (pref_size - 1))];
        hnl_fb_low = pref[(int)floor(0.5f * (pref_size - 1))];
    }

    if (hnl_fb_low < 0.6f && hnl_fb_low < a->hnl_fb_local_min) {
        a->hnl_fb_local_min = hnl_fb_low;
        a->hnl_fb_min = hnl_fb_low;
        a->hnl_new_min = 1;
        a->hnl_min_ctr = 0;
    }
    {
        float v = a->hnl_fb_local_min + 0.0008f / a->mult;
        a->hnl_fb_local_min = v < 1 ? v : 1;
        v = a->hnl_xd_avg_min + 0.0006f / a->mult;
        a->hnl_xd_avg_min = v < 1 ? v : 1;
    }
    if (a->hnl_new_min == 1) a->hnl_min_ctr++;
    if (a->hnl_min_ctr == 2) {
        a->hnl_new_min = 0;
        a->hnl_min_ctr = 0;
        {
            const float cand = k_target_supp[a->nlp_mode] / ((float)log(a->hnl_fb_min + 1e-10f) + 1e-10f);
            a->over_drive = cand > k_min_over[a->nlp_mode] ? cand : k_min_over[a->nlp_mode];
        }
    }
    if (a->over_drive < a->over_drive_sm) a->over_drive_sm = 0.99f * a->over_drive_sm + 0.01f * a->over_drive;
    else a->over_drive_sm = 0.9f * a->over_drive_sm + 0.1f * a->over_drive;

    /* OverdriveAndSuppress */
    for (i = 0; i < PART1; ++i) {
        if (hnl[i] > hnl_fb) hnl[i] = g_weight[i] * hnl_fb + (1 - g_weight[i]) * hnl[i];
        hnl[i] = powf(hnl[i], a->over_drive_sm * g_over[i]);
        efw[0][i] *= hnl[i];
        efw[1][i] *= hnl[i];
        efw[1][i] *= -1;
    }

    {   /* ComfortNoise (one band) */
        const float *noise_pow = a->noise_from_init ? a->d_init_min : a->d_min;
        float u[PART1][2], rnd[PART];
        for (i = 0; i < PART; ++i) {
            a->seed = (a->seed * 69069u + 1u) & 0x7fffffffu;          /* WebRtcSpl_RandU */
            rnd[i] = ((float)(int16_t)(a->seed >> 16)) / 32768;
        }
        u[0][0] = 0;
        u[0][1] = 0;
        for (i = 1; i < PART1; ++i) {
            const float ang = 6.28318530717959f * rnd[i - 1];
            const float amp = sqrtf(noise_pow[i]);
            u[i][0] = amp * cosf(ang);
            u[i][1] = -amp * sinf(ang);
        }
        u[PART][1] = 0;
        for (i = 0; i < PART1; ++i) {
            const float rest = 1 - hnl[i] * hnl[i];
            const float wgt = sqrtf(rest > 0 ? rest : 0);
            efw[0][i] += wgt * u[i][0];
            efw[1][i] += wgt * u[i][1];
        }
    }

    fft[0] = efw[0][0];
    fft[1] = efw[0][PART];
    for (i = 1; i < PART; ++i) {
        fft[2 * i] = efw[0][i];
        fft[2 * i + 1] = -efw[1][i];
    }
    orc_aec_rdft(fft, 1);
    {
        const float scale = 2.0f / PART2;
        for (i = 0; i < PART; ++i) {
            fft[i] *= scale;
            fft[i] = fft[i] * g_hann[i] + a->out_buf[i];
            fft[PART + i] *= scale;
            a->out_buf[i] = fft[PART + i] * g_hann[PART - i];
            output[i] = fft[i] > 32767 ? 32767 : (fft[i] < -32768 ? -32768 : fft[i]);
        }
    }
    memcpy(a->d_buf, a->d_buf + PART, sizeof(float) * PART);
    memcpy(a->e_buf, a->e_buf + PART, sizeof(float) * PART);
    memmove(a->xfw[1], a->xfw[0], sizeof(a->xfw) - sizeof(a->xfw[0]));
}

/* ProcessBlock :1143-1340 */
static void process_block(aec_t *a)
{
    float near[PART], fft[PART2], xf[2][PART1], yf[2][PART1], ef[2][PART1], df[2][PART1], e[PART], out[PART];
    const int noise_init_blocks = 500 * a->mult;
    int i, p;

    ring_read(a->near_fr, near, PART);
    memcpy(a->d_buf + PART, near, sizeof near);
    ring_read(a->far, xf, 1);

    memcpy(fft, a->d_buf, sizeof fft);
    time_to_freq(fft, df, 0);

    for (i = 0; i < PART1; ++i) {
        const float far_spec = (xf[0][i] * xf[0][i]) + (xf[1][i] * xf[1][i]);
        const float near_spec = df[0][i] * df[0][i] + df[1][i] * df[1][i];
        a->x_pow[i] = 0.9f * a->x_pow[i] + 0.1f * NPART * far_spec;
        a->d_pow[i] = 0.9f * a->d_pow[i] + 0.1f * near_spec;
    }
    if (a->noise_ctr > 50)
        for (i = 0; i < PART1; ++i) {
            if (a->d_pow[i] < a->d_min[i]) a->d_min[i] = (a->d_pow[i] + 0.1f * (a->d_min[i] - a->d_pow[i])) * 1.0002f;
            else a->d_min[i] *= 1.0002f;
        }
    if (a->noise_ctr < noise_init_blocks) {
        a->noise_ctr++;
        for (i = 0; i < PART1; ++i) {
            if (a->d_min[i] > a->d_init_min[i]) a->d_init_min[i] = 0.999f * a->d_init_min[i] + 0.001f * a->d_min[i];
            else a->d_init_min[i] = a->d_min[i];
        }
        a->noise_from_init = 1;
    } else {
        a->noise_from_init = 0;
    }

    a->xf_pos--;
    if (a->xf_pos == -1) a->xf_pos = NPART - 1;
    memcpy(a->xf[0] + a->xf_pos * PART1, xf[0], sizeof(float) * PART1);
    memcpy(a->xf[1] + a->xf_pos * PART1, xf[1], sizeof(float) * PART1);

    memset(yf, 0, sizeof yf);
    for (p = 0; p < NPART; ++p) {                                     /* FilterFar */
        int xp = (p + a->xf_pos) * PART1;
        const int pos = p * PART1;
        if (p + a->xf_pos >= NPART) xp -= NPART * PART1;
        for (i = 0; i < PART1; ++i) {
            yf[0][i] += a->xf[0][xp + i] * a->wf[0][pos + i] - a->xf[1][xp + i] * a->wf[1][pos + i];
            yf[1][i] += a->xf[0][xp + i] * a->wf[1][pos + i] + a->xf[1][xp + i] * a->wf[0][pos + i];
        }
    }
    fft[0] = yf[0][0];
    fft[1] = yf[0][PART];
    for (i = 1; i < PART; ++i) {
        fft[2 * i] = yf[0][i];
        fft[2 * i + 1] = yf[1][i];
    }
    orc_aec_rdft(fft, 1);
    for (i = 0; i < PART; ++i) e[i] = near[i] - fft[PART + i] * (2.0f / PART2);

    memcpy(a->e_buf + PART, e, sizeof e);
    memset(fft, 0, sizeof(float) * PART);
    memcpy(fft + PART, e, sizeof e);
    orc_aec_rdft(fft, 0);
    ef[1][0] = 0;
    ef[1][PART] = 0;
    ef[0][0] = fft[0];
    ef[0][PART] = fft[1];
    for (i = 1; i < PART; ++i) {
        ef[0][i] = fft[2 * i];
        ef[1][i] = fft[2 * i + 1];
    }

    for (i = 0; i < PART1; ++i) {                                     /* ScaleErrorSignal */
        float mag;
        ef[0][i] /= (a->x_pow[i] + 1e-10f);
        ef[1][i] /= (a->x_pow[i] + 1e-10f);
        mag = sqrtf(ef[0][i] * ef[0][i] + ef[1][i] * ef[1][i]);
        if (mag > a->err_thr) {
            mag = a->err_thr / (mag + 1e-10f);
            ef[0][i] *= mag;
            ef[1][i] *= mag;
        }
        ef[0][i] *= a->mu;
        ef[1][i] *= a->mu;
    }

    for (p = 0; p < NPART; ++p) {                                     /* FilterAdaptation */
        int xp = (p + a->xf_pos) * PART1;
        const int pos = p * PART1;
        if (p + a->xf_pos >= NPART) xp -= NPART * PART1;
        for (i = 0; i < PART; ++i) {
            const float xr = a->xf[0][xp + i], xi = -a->xf[1][xp + i];
            fft[2 * i] = xr * ef[0][i] - xi * ef[1][i];
            fft[2 * i + 1] = xr * ef[1][i] + xi * ef[0][i];
        }
        fft[1] = a->xf[0][xp + PART] * ef[0][PART] - (-a->xf[1][xp + PART]) * ef[1][PART];
        orc_aec_rdft(fft, 1);
        memset(fft + PART, 0, sizeof(float) * PART);
        for (i = 0; i < PART; ++i) fft[i] *= (2.0f / PART2);
        orc_aec_rdft(fft, 0);
        a->wf[0][pos] += fft[0];
        a->wf[0][pos + PART] += fft[1];
        for (i = 1; i < PART; ++i) {
            a->wf[0][pos + i] += fft[2 * i];
            a->wf[1][pos + i] += fft[2 * i + 1];
        }
    }

    nlp(a, out);
    ring_write(a->out_fr, out, PART);
}

/* WebRtcAec_ProcessFrames :1719-1860 (reported-delay branch) */
static void process_frames(aec_t *a, const float *near, int n, int known_delay, float *out)
{
    for (int j = 0; j < n; j += FRAME) {
        ring_write(a->near_fr, near + j, FRAME);
        if (a->system_delay < FRAME) move_far(a, -(a->mult + 1));
        {
            const int want = (a->known_delay - known_delay - 32) / PART;
            const int moved = ring_move(a->far, want);
            ring_move(a->far_win, want);
            a->known_delay -= moved * PART;
        }
        while (ring_avail_read(a->near_fr) >= PART) process_block(a);
        a->system_delay -= FRAME;
        {
            const int have = (int)ring_avail_read(a->out_fr);
            if (have < FRAME) ring_move(a->out_fr, have - FRAME);
        }
        ring_read(a->out_fr, out + j, FRAME);
    }
}

/* WebRtcAec_InitAec :1509-1688 + WebRtcAec_Init :196-276 + set_config (nlp aggressive) */
static void aec_reset(aec_t *a, int fs)
{
    ring_t *keep[6] = {a->near_fr, a->out_fr, a->far, a->far_win, a->far_pre, NULL};
    memset(a, 0, sizeof *a);
    a->near_fr = keep[0];
    a->out_fr = keep[1];
    a->far = keep[2];
    a->far_win = keep[3];
    a->far_pre = keep[4];
    for (int k = 0; k < 5; ++k) {
        ring_t *r = keep[k];
        r->rd = r->wr = 0;
        r->diff_wrap = 0;
        memset(r->data, 0, r->count * r->esize);
    }
    if (fs == 8000) { a->mu = 0.6f; a->err_thr = 2e-6f; }
    else { a->mu = 0.5f; a->err_thr = 1.5e-6f; }
    a->mult = fs / 8000;
    a->nlp_mode = 2;                                  /* kAecNlpAggressive, R:src/webrtc.c:224 */
    for (int i = 0; i < PART1; ++i) {
        a->d_min[i] = 1.0e6f;
        a->sd[i] = 1;
        a->sx[i] = 1;
    }
    a->noise_from_init = 1;
    a->hnl_fb_min = 1;
    a->hnl_fb_local_min = 1;
    a->hnl_xd_avg_min = 1;
    a->over_drive = 2;
    a->over_drive_sm = 2;
    a->seed = 777;
    ring_move(a->far_pre, -PART);                     /* start overlap, echo_cancellation.c:223-224 */
    a->rate_factor = fs / 8000;
    a->check_buff_size = 1;
    a->startup_phase = 1;                             /* reported_delay_enabled */
    a->filt_delay = -1;
}

/* WebRtcAec_BufferFarend :278-339 (no resampling) */
static int buffer_farend(aec_t *a, const float *far, int n)
{
    if (n != 80 && n != 160) return -1;
    a->system_delay += n;
    ring_write(a->far_pre, far, (size_t)n);
    while (ring_avail_read(a->far_pre) >= PART2) {
        float block[PART2];
        ring_read(a->far_pre, block, PART2);
        buffer_far_partition(a, block);
        ring_move(a->far_pre, -PART);
    }
    return 0;
}

/* EstBufDelayNormal :821-872 */
static void est_buf_delay(aec_t *a)
{
    const int n_snd = a->ms_in_snd * 8 * a->rate_factor;
    int cur = n_snd - a->system_delay, diff;
    cur += FRAME * a->rate_factor;
    if (cur < PART) cur += move_far(a, 1) * PART;
    a->filt_delay = a->filt_delay < 0 ? 0 : a->filt_delay;
    {
        const short f = (short)(0.8 * a->filt_delay + 0.2 * cur);
        a->filt_delay = f > 0 ? f : 0;
    }
    diff = a->filt_delay - a->api_known_delay;
    if (diff > 224) {
        if (a->last_delay_diff < 96) a->time_for_delay_change = 0;
        else a->time_for_delay_change++;
    } else if (diff < 96 && a->api_known_delay > 0) {
        if (a->last_delay_diff > 224) a->time_for_delay_change = 0;
        else a->time_for_delay_change++;
    } else {
        a->time_for_delay_change = 0;
    }
    a->last_delay_diff = (short)diff;
    if (a->time_for_delay_change > 25) {
        const int v = (int)a->filt_delay - 160;
        a->api_known_delay = v > 0 ? v : 0;
    }
}

/* WebRtcAec_Process :341-409 + ProcessNormal :599-747 (skew off) */
static int aec_process_frame(aec_t *a, const float *near, float *out, int n, int ms_in_snd_in)
{
    int ret = 0;
    short ms = (short)ms_in_snd_in, blocks10;
    if (n != 80 && n != 160) return -1;
    if (ms < 0) { ms = 0; ret = -1; }
    else if (ms > 500) ret = -1;
    ms = ms > 500 ? 500 : ms;
    ms += 10;
    a->ms_in_snd = ms;
    blocks10 = (short)(n / (FRAME * a->rate_factor));
    if (a->startup_phase) {
        if (near != out) memcpy(out, near, sizeof(float) * (size_t)n);
        if (a->check_buff_size) {
            a->check_ctr++;
            if (a->counter == 0) {
                a->first_val = a->ms_in_snd;
                a->sum = 0;
            }
            {
                const double lim = 0.2 * a->ms_in_snd > 8 ? 0.2 * a->ms_in_snd : 8;
                if (abs(a->first_val - a->ms_in_snd) < lim) {
                    a->sum += a->ms_in_snd;
                    a->counter++;
                } else {
                    a->counter = 0;
                }
            }
            if (a->counter * blocks10 >= 6) {
                const int v = (3 * a->sum * a->rate_factor * 8) / (4 * a->counter * PART);
                a->buf_size_start = v < 62 ? v : 62;
                a->check_buff_size = 0;
            }
            if (a->check_ctr * blocks10 > 50) {
                const int v = (a->ms_in_snd * a->rate_factor * 3) / 40;
                a->buf_size_start = v < 62 ? v : 62;
                a->check_buff_size = 0;
            }
        }
        if (!a->check_buff_size) {
            const int overhead = a->system_delay / PART - a->buf_size_start;
            if (overhead == 0) {
                a->startup_phase = 0;
            } else if (overhead > 0) {
                move_far(a, overhead);
                a->startup_phase = 0;
            }
        }
    } else {
        est_buf_delay(a);
        process_frames(a, near, n, a->api_known_delay, out);
    }
    return ret;
}

/* ---------------------------------------------------------------- wmix handle layer (R:src/webrtc.c:217-500) */
orc_aec *orc_aec_init(int chn, int freq, int interval_ms)
{
    orc_aec *h;
    if (freq > 16000 || freq % 8000 != 0) return NULL;
    orc_aec_tables(NULL, NULL, NULL, NULL);
    h = calloc(1, sizeof *h);
    h->a = calloc(1, sizeof *h->a);
    h->a->near_fr = ring_new(FRAME + PART, sizeof(float));
    h->a->out_fr = ring_new(FRAME + PART, sizeof(float));
    h->a->far = ring_new(FAR_RING, sizeof(float) * 2 * PART1);
    h->a->far_win = ring_new(FAR_RING, sizeof(float) * 2 * PART1);
    h->a->far_pre = ring_new(PRE_RING, sizeof(float));
    aec_reset(h->a, freq);
    h->chn = chn;
    h->freq = freq;
    if (freq <= 8000) h->interval_ms = (interval_ms % 20 == 0) ? 20 : 10;
    else h->interval_ms = 10;
    h->pkg = freq / 1000 * h->interval_ms;
    h->in = calloc((size_t)h->pkg, sizeof(float));
    h->out = calloc((size_t)h->pkg, sizeof(float));
    h->far = calloc((size_t)h->pkg, sizeof(float));
    return h;
}

int orc_aec_set_frame_far(orc_aec *h, const int16_t *far, int frame_num)
{
    const int total = frame_num * h->chn, step = h->pkg * h->chn;
    for (int done = 0; done < total; done += step) {
        for (int i = 0; i < h->pkg; ++i) {
            h->far[i] = (float)(*far);
            far += h->chn;
        }
        {
            const int rc = buffer_farend(h->a, h->far, (int16_t)h->pkg);
            if (rc) return rc;
        }
    }
    return 0;
}

int orc_aec_process(orc_aec *h, const int16_t *near, int16_t *out, int frame_num, int delay_ms)
{
    const int total = frame_num * h->chn, step = h->pkg * h->chn;
    for (int done = 0; done < total; done += step) {
        for (int i = 0; i < h->pkg; ++i) {
            h->in[i] = (float)(*near);
            near += h->chn;
        }
        {
            const int rc = aec_process_frame(h->a, h->in, h->out, h->pkg, (int16_t)delay_ms);
            if (rc) return rc;
        }
        for (int i = 0; i < h->pkg; ++i)
            for (int c = 0; c < h->chn; ++c) *out++ = (int16_t)h->out[i];
    }
    return 0;
}

int orc_aec_process2(orc_aec *h, const int16_t *far, const int16_t *near, int16_t *out, int frame_num, int delay_ms)
{
    const int total = frame_num * h->chn, step = h->pkg * h->chn;
    for (int done = 0; done < total; done += step) {
        for (int i = 0; i < h->pkg; ++i) {
            h->far[i] = (float)(*far);
            h->in[i] = (float)(*near);
            far += h->chn;
            near += h->chn;
        }
        {
            int rc = buffer_farend(h->a, h->far, (int16_t)h->pkg);
            if (rc) return rc;
            rc = aec_process_frame(h->a, h->in, h->out, h->pkg, (int16_t)delay_ms);
            if (rc) return rc;
        }
        for (int i = 0; i < h->pkg; ++i)
            for (int c = 0; c < h->chn; ++c) *out++ = (int16_t)h->out[i];
    }
    return 0;
}

void orc_aec_release(orc_aec *h)
{
    if (!h) return;
    ring_free(h->a->near_fr);
    ring_free(h->a->out_fr);
    ring_free(h->a->far);
    ring_free(h->a->far_win);
    ring_free(h->a->far_pre);
    free(h->a);
    free(h->in);
    free(h->out);
    free(h->far);
    free(h);
}

/* introspection for tests: far-ring fill and read-pointer excursions */
int orc_aec_far_available(const orc_aec *h) { return (int)ring_avail_read(h->a->far); }
int orc_aec_system_delay(const orc_aec *h) { return h->a->system_delay; }
int orc_aec_startup(const orc_aec *h) { return h->a->startup_phase; }
