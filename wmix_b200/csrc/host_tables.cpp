// Host-side, init-time constant tables for the kernels: NS window / FFT twiddles / log table,
// the AGC compressor curve, and the VAD mode thresholds.  These are identical for every stream
// of an engine, are computed once on the CPU (with the same libm the reference uses, so the
// floats agree bit for bit) and uploaded.  No per-sample work happens here.
#include "host_tables.h"
#include "nsx.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace wmx {
namespace host {

// ---- NS -------------------------------------------------------------------------------------

// Flat-top window with quarter-sine edges (T:.../ns/windows_private.h:64 kBlocks80w128, :94
// kBlocks160w256).  The reference ships it as literals printed with 8 decimals; printing and
// re-parsing the closed form reproduces every float of those tables exactly (tested).
void ns_window(int ana, int block, float* w)
{
    const int ov = ana - block;
    for (int i = 0; i < ana; ++i) {
        int k = i < ana - i ? i : ana - i;
        if (k > ov) k = ov;
        char txt[32];
        snprintf(txt, sizeof txt, "%.8f", sin(3.14159265358979323846 * k / (2.0 * ov)));
        w[i] = strtof(txt, nullptr);
    }
}

static void bitrev_pairs(int n, float* a)
{
    const int nc = n >> 1;
    int bits = 0;
    while ((1 << bits) < nc) ++bits;
    for (int c = 0; c < nc; ++c) {
        int r = 0;
        for (int b = 0; b < bits; ++b) r |= ((c >> b) & 1) << (bits - 1 - b);
        if (r > c) {
            float t0 = a[2 * c], t1 = a[2 * c + 1];
            a[2 * c] = a[2 * r]; a[2 * c + 1] = a[2 * r + 1];
            a[2 * r] = t0; a[2 * r + 1] = t1;
        }
    }
}

// cos/sin table of the complex passes, stored bit-reversed (T:.../fft4g.c:642-668 makewt)
void fft_w_table(int nw, float* w)
{
    const int h = nw >> 1;
    const float d = (float)atan((double)1.0f) / h;   // NB: double libm, as the C reference calls it
    w[0] = 1;
    w[1] = 0;
    w[h] = (float)cos((double)(d * h));
    w[h + 1] = w[h];
    for (int j = 2; j < h; j += 2) {
        const float x = (float)cos((double)(d * j)), y = (float)sin((double)(d * j));
        w[j] = x; w[j + 1] = y;
        w[nw - j] = y; w[nw - j + 1] = x;
    }
    if (h > 2) bitrev_pairs(nw, w);
}

// half-scaled cos/sin table of the real split (T:.../fft4g.c:671-688 makect)
void fft_c_table(int nc, float* c)
{
    const int h = nc >> 1;
    const float d = (float)atan((double)1.0f) / h;
    c[0] = (float)cos((double)(d * h));
    c[h] = 0.5f * c[0];
    for (int j = 1; j < h; ++j) {
        c[j] = 0.5f * (float)cos((double)(d * j));
        c[nc - j] = 0.5f * (float)sin((double)(d * j));
    }
}

// (float)log((float)i) and the two start-up regressor sums the reference re-derives every
// frame (T:.../ns/ns_core.c:1089-1101); accumulated in float, i = 5 .. bins-1, in order
void ns_log_table(int bins, float* log_i, float* sum, float* sum_sq)
{
    float s = 0.f, s2 = 0.f;
    log_i[0] = 0.f;
    for (int i = 1; i < bins; ++i) {
        const float v = (float)log((double)(float)i);
        log_i[i] = v;
        if (i >= 5) { s += v; s2 += v * v; }
    }
    *sum = s;
    *sum_sq = s2;
}

// tables of ns.cuh's log_f / exp_f: sub-interval centres, their reciprocals and logs, and
// 2^(j/128), evaluated in long double and rounded once to double
void dmath_tables(double* invc, double* logc, double* exp2jn)
{
    for (int i = 0; i < 128; ++i) {
        union { uint64_t u; double d; } lo, hi;
        lo.u = 0x3fe6000000000000ull + ((uint64_t)i << 45);
        hi.u = 0x3fe6000000000000ull + ((uint64_t)(i + 1) << 45);
        const long double c = (i == 80) ? 1.0L : 0.5L * ((long double)lo.d + (long double)hi.d);
        invc[i] = (double)(1.0L / c);
        logc[i] = (i == 80) ? 0.0 : (double)(-logl((long double)invc[i]));
        exp2jn[i] = (double)exp2l((long double)i / 128.0L);
    }
}

// T:.../ns/ns_core.c:1012-1041
int ns_policy(int mode, float* overdrive, float* floor_gain, int* gainmap)
{
    switch (mode) {
    case 0: *overdrive = 1.f; *floor_gain = 0.5f; *gainmap = 0; return 0;
    case 1: *overdrive = 1.f; *floor_gain = 0.25f; *gainmap = 1; return 0;
    case 2: *overdrive = 1.1f; *floor_gain = 0.125f; *gainmap = 1; return 0;
    case 3: *overdrive = 1.25f; *floor_gain = 0.09f; *gainmap = 1; return 0;
    }
    return -1;
}

// ---- VAD ------------------------------------------------------------------------------------

// T:.../vad/vad_core.c:78-100: per mode, columns = 10 / 20 / 30 ms
int vad_thresholds(int mode, int frame_ms, int16_t out[4])
{
    static const int16_t oh1[4][3] = {{8, 4, 3}, {8, 4, 3}, {6, 3, 2}, {6, 3, 2}};
    static const int16_t oh2[4][3] = {{14, 7, 5}, {14, 7, 5}, {9, 5, 3}, {9, 5, 3}};
    static const int16_t loc[4][3] = {{24, 21, 24}, {37, 32, 37}, {82, 78, 82}, {94, 94, 94}};
    static const int16_t glo[4][3] = {{57, 48, 57}, {100, 80, 100}, {285, 260, 285}, {1100, 1050, 1100}};
    if (mode < 0 || mode > 3) return -1;
    const int col = frame_ms == 10 ? 0 : (frame_ms == 20 ? 1 : 2);
    out[0] = oh1[mode][col];
    out[1] = oh2[mode][col];
    out[2] = loc[mode][col];
    out[3] = glo[mode][col];
    return 0;
}

// initial GMM (T:.../vad/vad_core.c:40-63) in the kernel's packed word layout
void vad_initial_words(int32_t* words /* vad::N_WORDS */)
{
    static const int16_t nmean[12] = {6738, 4892, 7065, 6715, 6771, 3369, 7646, 3863, 7820, 7266, 5020, 4362};
    static const int16_t smean[12] = {8306, 10085, 10078, 11823, 11843, 6309, 9473, 9571, 10879, 7581, 8180, 7483};
    static const int16_t nstd[12] = {378, 1064, 493, 582, 688, 593, 474, 697, 475, 688, 421, 455};
    static const int16_t sstd[12] = {555, 505, 567, 524, 585, 1231, 509, 828, 492, 1540, 1079, 850};
    auto pk = [](int16_t lo, int16_t hi) { return (int32_t)(((uint32_t)(uint16_t)hi << 16) | (uint16_t)lo); };
    memset(words, 0, sizeof(int32_t) * 138);
    for (int ch = 0; ch < 6; ++ch) {
        words[2 + ch] = pk(nmean[ch], nmean[ch + 6]);
        words[8 + ch] = pk(smean[ch], smean[ch + 6]);
        words[14 + ch] = pk(nstd[ch], nstd[ch + 6]);
        words[20 + ch] = pk(sstd[ch], sstd[ch + 6]);
    }
    for (int i = 0; i < 48; ++i) {
        words[28 + i] = 0;                       // ages
        words[76 + i] = pk(10000, 10000);        // smallest values
    }
    for (int i = 0; i < 3; ++i) words[124 + i] = pk(1600, 1600);
    words[135] = 4;                              // wmix starts muted (R:src/webrtc.c:64)
}

// ---- AGC ------------------------------------------------------------------------------------

static int16_t div16(int32_t num, int16_t den) { return den ? (int16_t)(num / den) : (int16_t)0x7FFF; }
static int clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static int normw32(int32_t a) { if (!a) return 0; if (a < 0) a = ~a; return clz((uint32_t)a) - 1; }
static int32_t shl(int32_t x, int c) { return c >= 0 ? (int32_t)((uint32_t)x << c) : (x >> (-c)); }

// generating function table of the compressor, round(256*log2(1+e^i)) (digital_agc.c:36-53)
static uint16_t genfunc(int i) { return (uint16_t)floor(256.0 * log2(1.0 + exp((double)i)) + 0.5); }

// analog target implied by the compression gain (T:.../agc/legacy/analog_agc.c:424-448)
int16_t agc_analog_target(int16_t comp_db)
{
    int16_t t = div16((int32_t)(int16_t)(5 * comp_db + 5), 11);
    t = (int16_t)(4 + t);
    return t < 4 ? (int16_t)4 : t;
}

// 32-entry Q16 gain curve (T:.../agc/legacy/digital_agc.c:57-257)
int agc_gain_table(int32_t table[32], int16_t comp_db, int16_t target_dbfs, int limiter, int16_t analog_target)
{
    const int32_t kLog10 = 54426, kLog10_2 = 49321, kLogE_1 = 23637, kRatio = 3;
    const int16_t headroom = (int16_t)(analog_target - target_dbfs);
    int16_t t16 = (int16_t)(headroom + div16((comp_db - analog_target) * (kRatio - 1) + (kRatio >> 1), kRatio));
    const int16_t max_gain = t16 > headroom ? t16 : headroom;
    const int16_t diff_gain = div16(comp_db * (kRatio - 1) + (kRatio >> 1), kRatio);
    if (diff_gain < 0 || diff_gain >= 128) return -1;
    const int16_t lim_idx = (int16_t)(2 + div16((int32_t)analog_target << 13, (int16_t)(kLog10_2 / 2)));
    const int32_t lim_lvl = target_dbfs + div16(kRatio >> 1, kRatio);
    const uint16_t gmax = genfunc(diff_gain);
    const int32_t den = 20 * (int32_t)gmax;

    for (int i = 0; i < 32; ++i) {
        // input level of this table slot, Q14
        int32_t lvl = (int32_t)(int16_t)((kRatio - 1) * (i - 1)) * kLog10_2 + 1;
        lvl = ((int32_t)diff_gain << 14) - lvl / kRatio;
        const uint32_t mag = (uint32_t)(lvl >= 0 ? lvl : -lvl);
        const int ip = (int)(mag >> 14);
        const uint32_t fp = mag & 0x3FFF;
        uint32_t a = (uint32_t)(uint16_t)(genfunc(ip + 1) - genfunc(ip)) * fp + ((uint32_t)genfunc(ip) << 14);
        uint32_t approx = a >> 8;
        if (lvl < 0) {
            const int z = mag ? clz(mag) : 0;
            int zs = 0;
            uint32_t b;
            if (z < 15) {
                b = (mag >> (15 - z)) * (uint32_t)kLogE_1;
                if (z < 9) { zs = 9 - z; a >>= zs; }
                else b >>= z - 9;
            } else {
                b = (mag * (uint32_t)kLogE_1) >> 6;
            }
            approx = (b < a) ? (a - b) >> (8 - zs) : 0;
        }
        int32_t num = (int32_t)((uint32_t)(max_gain * (int32_t)gmax) << 6);
        num -= (int32_t)approx * diff_gain;
        const int z = (num > (den >> 8)) ? normw32(num) : normw32(den) + 8;
        num = (int32_t)((uint32_t)num << z);
        const int32_t d = shl(den, z - 8);
        num += (num < 0) ? -(d / 2) : (d / 2);
        int32_t y = num / d;
        if (limiter && i < lim_idx) {
            int32_t t = (int32_t)(int16_t)(i - 1) * kLog10_2;
            t -= (int32_t)((uint32_t)lim_lvl << 14);
            y = (t + 10) / 20;
        }
        int32_t g = (y > 39000) ? (((y >> 1) * kLog10 + 4096) >> 13) : ((y * kLog10 + 8192) >> 14);
        g += 16 << 14;
        if (g <= 0) { table[i] = 0; continue; }
        const int e = (uint16_t)(int16_t)(g >> 14);
        const int32_t f = g & 0x3FFF;
        int32_t frac;
        if (f >> 13) frac = (1 << 14) - ((((1 << 14) - f) * ((2 << 14) - 22817)) >> 13);
        else frac = (f * (22817 - (1 << 14))) >> 13;
        table[i] = (int32_t)(1u << e) + shl((int32_t)(uint16_t)frac, e - 14);
    }
    return 0;
}

// AgcVad / DigitalAgc init (T:.../agc/legacy/digital_agc.c:259-281, :606-631) in word layout
void agc_initial_words(int32_t* words /* agc::N_WORDS */)
{
    auto pk = [](int16_t lo, int16_t hi) { return (int32_t)(((uint32_t)(uint16_t)hi << 16) | (uint16_t)lo); };
    memset(words, 0, sizeof(int32_t) * 18);
    words[0] = 134217728;            // capacitorSlow
    words[2] = 65536;                // gain
    words[12] = pk(0, 3);            // (HPstate, counter)
    words[13] = pk(0, 15 << 10);     // (logRatio, meanLongTerm)
    words[14] = 500 << 8;            // varianceLongTerm
    words[15] = pk(0, 15 << 10);     // (stdLongTerm, meanShortTerm)
    words[16] = 500 << 8;            // varianceShortTerm
}


// ---- AEC (T:webrtc/modules/audio_processing/aec) ------------------------------------------------
// rdft_w (aec_rdft.c:32-49) is Ooura's makewt(32)/makect(32) table for a 128-point transform,
// except that the shipped literal was generated by a cosf/sinf that was one ulp off in eight
// entries.  The table below is therefore the correctly rounded one plus those eight one-ulp
// corrections (pinned against the reference's exported symbol by tests/test_oracle_pin.py).
// sqrtHanning (aec_core.c:49-66) is sin(pi i/128); weightCurve (:71-81) and overDriveCurve
// (:86-96) are the 4-decimal Matlab prints 0.3 sqrt(linspace(0,1,64)) + 0.1 (with a leading 0)
// and sqrt(linspace(0,1,65)) + 1.
void aec_tables(float* w, float* c, float* hann, float* weight, float* over, uint32_t* lcg_mul, uint32_t* lcg_add)
{
    fft_w_table(32, w);
    fft_c_table(32, c);
    static const struct { int idx, ulps; } fix[8] = {{4, 1}, {7, 1}, {20, 1}, {27, 1}, {40, 1}, {41, -1}, {42, 1}, {47, 1}};
    for (int k = 0; k < 8; ++k) {
        float* p = fix[k].idx < 32 ? &w[fix[k].idx] : &c[fix[k].idx - 32];
        int32_t bits;
        memcpy(&bits, p, 4);
        bits += fix[k].ulps;
        memcpy(p, &bits, 4);
    }
    for (int i = 0; i <= 64; ++i) {
        hann[i] = (float)sin(3.14159265358979323846 * i / 128.0);
        weight[i] = i == 0 ? 0.f : (float)(floor((0.3 * sqrt((i - 1) / 63.0) + 0.1) * 1e4 + 0.5) / 1e4);
        over[i] = (float)(floor((sqrt(i / 64.0) + 1.0) * 1e4 + 0.5) / 1e4);
    }
    // x_{k} = mul[k] * x_0 + add[k]  (mod 2^32; the generator keeps the low 31 bits)
    uint32_t m = 1, a = 0;
    for (int k = 0; k <= 64; ++k) {
        lcg_mul[k] = m;
        lcg_add[k] = a;
        a = a * 69069u + 1u;
        m = m * 69069u;
    }
}

// ---- wmix_pcm_zoom / wmix_len_of_* (R:src/wmix.c:49-222) ----
// One phase accumulator drives all three: a float that gains `div` (the smaller rate over the larger) per
// step; when its integer part turns positive the slower side moves and 1.0 is taken off — in double,
// stored back to float, which is what `divStep -= 1.0` does to a float in C.
namespace {
struct ZoomPhase {
    float div, acc = 0.f;
    bool up;
    ZoomPhase(int in_freq, int out_freq) : up(in_freq < out_freq)
    {
        div = up ? (float)in_freq / (float)out_freq : (float)out_freq / (float)in_freq;
    }
    bool step()
    {
        acc += div;
        if ((int)acc > 0) {
            acc = (float)((double)acc - 1.0);
            return true;
        }
        return false;
    }
};
}  // namespace

uint32_t zoom_len_of_out(int in_chn, int in_freq, uint32_t in_len, int out_chn, int out_freq)
{
    if (in_freq == out_freq && in_chn == out_chn) return in_len;
    ZoomPhase ph(in_freq, out_freq);
    uint32_t in_n = 0, out_n = 0;
    while (in_n < in_len) {
        if (ph.up) { out_n += out_chn; if (ph.step()) in_n += in_chn; }
        else { if (ph.step()) out_n += out_chn; in_n += in_chn; }
    }
    return out_n;
}

uint32_t zoom_len_of_in(int in_chn, int in_freq, int out_chn, int out_freq, uint32_t out_len)
{
    if (in_freq == out_freq && in_chn == out_chn) return out_len;
    ZoomPhase ph(in_freq, out_freq);
    uint32_t in_n = 0, out_n = 0;
    while (out_n < out_len) {
        if (ph.up) { out_n += out_chn; if (ph.step()) in_n += in_chn; }
        else { if (ph.step()) out_n += out_chn; in_n += in_chn; }
    }
    return in_n;
}

uint32_t zoom_map(int in_chn, int in_freq, uint32_t in_bytes, int out_chn, int out_freq, int32_t* map)
{
    const uint32_t in_samples = in_bytes / 2;
    if (in_freq == out_freq && in_chn == out_chn) {                     // memcpy branch
        if (map) for (uint32_t i = 0; i < in_samples; ++i) map[i] = (int32_t)i;
        return in_samples;
    }
    // samples one output frame takes from the input frame at `pos`: 1->1 and 2->1 copy the first (left)
    // sample, 1->2 writes it twice, 2->2 is unreachable in the reference and writes nothing
    const int mode = (in_chn << 4) | (out_chn & 0x0F);
    const int emit = mode == 0x11 || mode == 0x21 ? 1 : (mode == 0x12 ? 2 : 0);
    ZoomPhase ph(in_freq, out_freq);
    uint32_t pos = 0, n = 0;
    // the reference compares int16 pointers, so an odd trailing byte still counts as a readable sample
    const uint32_t end = (in_bytes + 1) / 2;
    while (pos < end) {
        bool write = ph.up;
        bool advance = !ph.up;
        if (ph.step()) { if (ph.up) advance = true; else write = true; }
        if (write) {
            for (int k = 0; k < emit; ++k) { if (map) map[n] = (int32_t)pos; ++n; }
        }
        if (advance) pos += in_chn;
    }
    return n;
}


// ---- resampling branches of wmix_load_data for a mono bus (R:src/wmix.c:1704-1939) ----
// The float phase accumulator decides, per bus sample, whether it is a copied source frame or the k-th of n ramp
// values between the frame just copied and the next one.  None of that depends on the audio, so it is walked once
// here, exactly as the reference walks it, and the kernel only evaluates the ramps.
//   map[i]  = int16 index of the (left) source sample feeding bus sample i
//   ramp[i] = 0 for a copied frame, else (n << 8) | (k + 1): value = (int16)(src[map] + (k+1 accumulated steps of
//             (float)(src[map + chn] - src[map]) / n))
// Returns the number of bus samples, or UINT32_MAX where the reference would overrun its 64-entry ramp buffer.
uint32_t mix_plan(int src_chn, int src_freq, uint32_t src_bytes, int mix_freq, int32_t* map, uint16_t* ramp)
{
    const uint32_t frame_bytes = 2u * (uint32_t)src_chn;
    uint32_t used = 0, n_out = 0;
    int32_t pos = 0;
    float acc = 0.f;
    if (src_freq > mix_freq) {
        const float gain = (float)(src_freq - mix_freq) / (float)mix_freq;
        while (used < src_bytes) {
            if (acc >= 1.0) {
                acc = (float)((double)acc - 1.0);
            } else {
                if (map) { map[n_out] = pos; ramp[n_out] = 0; }
                ++n_out;
                acc += gain;
            }
            pos += src_chn;
            used += frame_bytes;
        }
    } else {
        const float gain = (float)(mix_freq - src_freq) / (float)src_freq;
        int n = 0, k = 0;
        while (used < src_bytes) {
            if (acc >= 1.0) {
                if (map) { map[n_out] = pos - src_chn; ramp[n_out] = (uint16_t)((n << 8) | (k + 1)); }
                ++k;
                acc = (float)((double)acc - 1.0);
            } else {
                if (map) { map[n_out] = pos; ramp[n_out] = 0; }
                pos += src_chn;
                used += frame_bytes;
                acc += gain;
                if (acc >= 1.0) {
                    n = (int)acc + 1;
                    k = 0;
                    if (n > 64) return 0xFFFFFFFFu;
                }
            }
            ++n_out;
        }
    }
    return n_out;
}

// ---- playPkgBuff_get's slot choice (R:src/wmix.c:496-509) for a delay of whole packages ----
// count = the ring's next write index.  The clamp-then-subtract of the reference makes the result `delay` while
// count >= delay and `count` (the oldest package) otherwise; reproduced literally.
int play_fifo_slot(int count, int n_pkg, int delay_pkgs)
{
    int k = count - delay_pkgs;
    k = k >= n_pkg ? n_pkg : (k < 0 ? 0 : k);
    k = count - k;
    if (k >= n_pkg) k -= n_pkg;
    else if (k < 0) k += n_pkg;
    return k;
}


// ---- NS, fixed-point core ---------------------------------------------------------------------
// The reference ships these as literal tables (T:.../ns/nsx_core.c:28-300, nsx_core_c.c:17,
// T:.../signal_processing/complex_fft_tables.h:17).  All but the 17-entry sigmoid have closed forms that reproduce the
// literals exactly (tests pin every entry against hashes of the reference's sources): rounded logarithms and reciprocals,
// a quarter-sine-edged window in Q14, the gain-map curves of the comments at nsx_core.c:125-169 evaluated on
// gain = sqrt(energy ratio) and TRUNCATED to Q13, sums of log2(i) for the pink-noise fit, and a truncated Q15 sine.
static int nearest_int(double x) { return (int)floor(x + 0.5); }

int nsx_tables(int freq, int policy, nsx::Tables* T, int32_t* thr_lrt)
{
    if (policy < 0 || policy > 3) return -1;
    if (freq != 8000 && freq != 16000 && freq != 32000) return -1;
    memset(T, 0, sizeof *T);
    const bool nb = freq == 8000;
    const int ana = nb ? 128 : 256, ramp = nb ? 48 : 96, step = 1024 / ana;
    const double pi = 3.14159265358979323846;
    for (int m = 0; m < ana / 2; ++m) {
        const int j = m * step;
        const int16_t sn = (int16_t)(32767.0 * sin(2.0 * pi * j / 1024.0)), cs = (int16_t)(32767.0 * sin(2.0 * pi * (j + 256) / 1024.0));
        T->twiddle[m] = ((uint32_t)(uint16_t)cs << 16) | (uint16_t)sn;
    }
    for (int i = 0; i < ana; ++i) {
        const int k = i < ana - i ? i : ana - i;
        T->window[i] = (int16_t)(k >= ramp ? 16384 : nearest_int(16384.0 * sin(pi * k / (2.0 * ramp))));
    }
    for (int i = 0; i < 256; ++i) T->log_frac[i] = (int16_t)nearest_int(256.0 * log2(1.0 + i / 256.0));
    for (int i = 0; i < 201; ++i) {
        const int v = nearest_int(32768.0 / (i + 1));
        T->counter_div[i] = (int16_t)(v > 32767 ? 32767 : v);
    }
    for (int i = 1; i <= 128; ++i) T->log_index[i] = (int16_t)nearest_int(4096.0 * log2((double)i));
    static const double floor_amp[4] = {0.5, 0.25, 0.125, 0.09};
    for (int i = 0; i <= 256; ++i) {
        const double g = sqrt(i / 256.0);
        double f1 = 1.0, f2 = 1.0;
        if (g > 0.5) {
            f1 = 1.0 + 1.3 * (g - 0.5);
            if (g * f1 > 1.0) f1 = 1.0 / g;
        } else {
            f2 = 1.0 - 0.3 * (0.5 - (g <= floor_amp[policy] ? floor_amp[policy] : g));
        }
        T->factor1[i] = (int16_t)(8192.0 * f1);
        T->factor2[i] = (int16_t)(8192.0 * f2);   // policy 0 never reads it (gain_map = 0)
    }
    static const int16_t sigmoid[17] = {0, 2017, 3809, 5227, 6258, 6963, 7424, 7718, 7901, 8014, 8084, 8126, 8152, 8168, 8177, 8183, 8187};
    memcpy(T->sigmoid, sigmoid, sizeof sigmoid);
    for (int i = 0; i < 9; ++i) T->log_stage[i] = (int16_t)nearest_int(i * log(2.0) * 256.0);
    // pink-noise fit over bins kStartBand .. 128, cut back to .. 64 for the narrow band (nsx_core.c:1364-1377)
    auto sums = [](int from, int16_t* s5, int16_t* q2, int16_t* det) {
        double s = 0.0, q = 0.0;
        for (int j = from; j <= 128; ++j) {
            const double l = log2((double)j);
            s += l;
            q += l * l;
        }
        *s5 = (int16_t)nearest_int(32.0 * s);
        *q2 = (int16_t)nearest_int(4.0 * q);
        *det = (int16_t)nearest_int((129 - from) * q - s * s);
    };
    int16_t s5, q2, det, s65, q65, d65;
    sums(nsx::kStartBand, &s5, &q2, &det);
    sums(65, &s65, &q65, &d65);
    if (nb) {
        int32_t t = det;
        t += (s65 * s5) >> 9;
        t -= (s65 * s65) >> 10;
        t -= (int32_t)q2 << 4;
        t -= ((65 - nsx::kStartBand) * q65) >> 2;
        det = (int16_t)t;
        s5 = (int16_t)(s5 - s65);
        q2 = (int16_t)(q2 - q65);
    }
    T->fit_det = det;
    T->fit_sum_log = s5;
    T->fit_sum_sq = q2;
    static const int over[4] = {256, 256, 282, 320}, bound[4] = {8192, 4096, 2048, 1475};   // nsx_core.c:786-814
    T->overdrive = over[policy];
    T->floor_gain = bound[policy];
    T->gain_map = policy != 0;
    T->lrt_max = nb ? 0x0040000 : 0x0080000;     // nsx_core.c:647-663
    T->lrt_min = nb ? 52429 : 104858;
    *thr_lrt = nb ? 131072 : 212644;
    return 0;
}

}  // namespace host
}  // namespace wmx
