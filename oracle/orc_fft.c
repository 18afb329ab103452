/* ORACLE (test infrastructure) — Ooura's split-radix-4/2 real FFT as WebRTC's NS uses it,
 * restated from T:webrtc/common_audio/fft4g.c (rdft :324-362, makewt :642-668, makect
 * :671-688, bitrv2 :693, cftfsub :902, cftbsub :952, cft1st :1002, cftmdl :1107, rftfsub
 * :1234, rftbsub :1259).
 *
 * The float results must be bit-identical to the reference, so every butterfly keeps the
 * reference's operand order; what differs is the organisation: the index scrambling is an
 * explicit bit-reversal of complex indices (checked equal to bitrv2 for n = 16..256), and
 * the radix-4 passes are one routine parameterised by the span instead of a first-pass /
 * middle-pass pair, with the four twiddle cases (unit, pi/4, even, odd group) selected per
 * group.
 *
 * Packing (same as the reference): forward output a[0]=Re X0, a[1]=Re X(n/2),
 * a[2k],a[2k+1]=Re,Im X(k); inverse expects that and returns n/2 times the signal. */
#include <math.h>
#include "oracle.h"

static void orc_bitrev_pairs(int n, float *a)
{
    int nc = n >> 1, bits = 0, c;
    while ((1 << bits) < nc)
        ++bits;
    for (c = 0; c < nc; ++c) {
        int r = 0, b;
        for (b = 0; b < bits; ++b)
            if (c & (1 << b))
                r |= 1 << (bits - 1 - b);
        if (r > c) {
            float tr = a[2 * c], ti = a[2 * c + 1];
            a[2 * c] = a[2 * r];
            a[2 * c + 1] = a[2 * r + 1];
            a[2 * r] = tr;
            a[2 * r + 1] = ti;
        }
    }
}

/* twiddle table for the complex passes: nw floats (makewt) */
static void orc_make_w(int nw, int *ip, float *w)
{
    int j, h = nw >> 1;
    float d;
    ip[0] = nw;
    ip[1] = 1;
    if (nw <= 2)
        return;
    d = (float)atan(1.0f) / h;
    w[0] = 1;
    w[1] = 0;
    w[h] = (float)cos(d * h);
    w[h + 1] = w[h];
    for (j = 2; j < h; j += 2) {
        float x = (float)cos(d * j), y = (float)sin(d * j);
        w[j] = x;
        w[j + 1] = y;
        w[nw - j] = y;
        w[nw - j + 1] = x;
    }
    if (h > 2)
        orc_bitrev_pairs(nw, w);
}

/* table for the real<->complex split: nc floats (makect) */
static void orc_make_c(int nc, int *ip, float *c)
{
    int j, h = nc >> 1;
    float d;
    ip[1] = nc;
    if (nc <= 1)
        return;
    d = (float)atan(1.0f) / h;
    c[0] = (float)cos(d * h);
    c[h] = 0.5f * c[0];
    for (j = 1; j < h; ++j) {
        c[j] = 0.5f * (float)cos(d * j);
        c[nc - j] = 0.5f * (float)sin(d * j);
    }
}

/* One radix-4 pass with butterflies spanning l floats (l = 2 is the first pass).
 * Group g covers floats [g*4l, (g+1)*4l): g = 0 has unit twiddles, g = 1 the pi/4 ones,
 * g = 2t / 2t+1 (t >= 1) use w[2t], w[4t] / w[4t+2] as in cft1st/cftmdl. */
static void orc_radix4_pass(int n, int l, float *a, const float *w)
{
    int m = l << 2, g, j;
    for (g = 0; g * m < n; ++g) {
        int base = g * m, t = g >> 1;
        float w1r = 0, w1i = 0, w2r = 0, w2i = 0, w3r = 0, w3i = 0;
        if (g >= 2) {
            float ar = w[2 * t], ai = w[2 * t + 1];
            if (g & 1) {
                w1r = w[4 * t + 2];
                w1i = w[4 * t + 3];
                w3r = w1r - 2 * ar * w1i;
                w3i = 2 * ar * w1r - w1i;
                w2r = -ai;
                w2i = ar;
            } else {
                w1r = w[4 * t];
                w1i = w[4 * t + 1];
                w3r = w1r - 2 * ai * w1i;
                w3i = 2 * ai * w1r - w1i;
                w2r = ar;
                w2i = ai;
            }
        }
        for (j = base; j < base + l; j += 2) {
            int j1 = j + l, j2 = j1 + l, j3 = j2 + l;
            float x0r = a[j] + a[j1], x0i = a[j + 1] + a[j1 + 1];
            float x1r = a[j] - a[j1], x1i = a[j + 1] - a[j1 + 1];
            float x2r = a[j2] + a[j3], x2i = a[j2 + 1] + a[j3 + 1];
            float x3r = a[j2] - a[j3], x3i = a[j2 + 1] - a[j3 + 1];
            float pr, pi;
            a[j] = x0r + x2r;
            a[j + 1] = x0i + x2i;
            if (g == 0) {
                a[j2] = x0r - x2r;
                a[j2 + 1] = x0i - x2i;
                a[j1] = x1r - x3i;
                a[j1 + 1] = x1i + x3r;
                a[j3] = x1r + x3i;
                a[j3 + 1] = x1i - x3r;
            } else if (g == 1) {
                float q = w[2];
                a[j2] = x2i - x0i;
                a[j2 + 1] = x0r - x2r;
                pr = x1r - x3i;
                pi = x1i + x3r;
                a[j1] = q * (pr - pi);
                a[j1 + 1] = q * (pr + pi);
                pr = x3i + x1r;
                pi = x3r - x1i;
                a[j3] = q * (pi - pr);
                a[j3 + 1] = q * (pi + pr);
            } else {
                pr = x0r - x2r;
                pi = x0i - x2i;
                a[j2] = w2r * pr - w2i * pi;
                a[j2 + 1] = w2r * pi + w2i * pr;
                pr = x1r - x3i;
                pi = x1i + x3r;
                a[j1] = w1r * pr - w1i * pi;
                a[j1 + 1] = w1r * pi + w1i * pr;
                pr = x1r + x3i;
                pi = x1i - x3r;
                a[j3] = w3r * pr - w3i * pi;
                a[j3 + 1] = w3r * pi + w3i * pr;
            }
        }
    }
}

/* complex passes after the bit reversal; `back` selects cftbsub's conjugating last pass */
static void orc_complex_passes(int n, float *a, const float *w, int back)
{
    int l = 2, j;
    if (n > 8) {
        orc_radix4_pass(n, 2, a, w);
        l = 8;
        while ((l << 2) < n) {
            orc_radix4_pass(n, l, a, w);
            l <<= 2;
        }
    }
    if ((l << 2) == n) {
        for (j = 0; j < l; j += 2) {
            int j1 = j + l, j2 = j1 + l, j3 = j2 + l;
            float x0r = a[j] + a[j1], x1r = a[j] - a[j1];
            float x0i, x1i;
            float x2r = a[j2] + a[j3], x2i = a[j2 + 1] + a[j3 + 1];
            float x3r = a[j2] - a[j3], x3i = a[j2 + 1] - a[j3 + 1];
            if (!back) {
                x0i = a[j + 1] + a[j1 + 1];
                x1i = a[j + 1] - a[j1 + 1];
                a[j] = x0r + x2r;
                a[j + 1] = x0i + x2i;
                a[j2] = x0r - x2r;
                a[j2 + 1] = x0i - x2i;
                a[j1] = x1r - x3i;
                a[j1 + 1] = x1i + x3r;
                a[j3] = x1r + x3i;
                a[j3 + 1] = x1i - x3r;
            } else {
                x0i = -a[j + 1] - a[j1 + 1];
                x1i = -a[j + 1] + a[j1 + 1];
                a[j] = x0r + x2r;
                a[j + 1] = x0i - x2i;
                a[j2] = x0r - x2r;
                a[j2 + 1] = x0i + x2i;
                a[j1] = x1r - x3i;
                a[j1 + 1] = x1i - x3r;
                a[j3] = x1r + x3i;
                a[j3 + 1] = x1i + x3r;
            }
        }
    } else {
        for (j = 0; j < l; j += 2) {
            int j1 = j + l;
            float dr = a[j] - a[j1], di;
            if (!back) {
                di = a[j + 1] - a[j1 + 1];
                a[j] += a[j1];
                a[j + 1] += a[j1 + 1];
            } else {
                di = -a[j + 1] + a[j1 + 1];
                a[j] += a[j1];
                a[j + 1] = -a[j + 1] - a[j1 + 1];
            }
            a[j1] = dr;
            a[j1 + 1] = di;
        }
    }
}

void orc_rdft(int n, int isgn, float *a, int *ip, float *w)
{
    int nw = ip[0], nc, m = n >> 1, j;
    const float *c;
    if (n > (nw << 2)) {
        nw = n >> 2;
        orc_make_w(nw, ip, w);
    }
    nc = ip[1];
    if (n > (nc << 2)) {
        nc = n >> 2;
        orc_make_c(nc, ip, w + nw);
    }
    c = w + nw;
    if (isgn >= 0) {
        int ks = 2 * nc / m;
        orc_bitrev_pairs(n, a);
        orc_complex_passes(n, a, w, 0);
        for (j = 2; j < m; j += 2) {               /* rftfsub */
            int k = n - j, kk = ks * (j >> 1);
            float wkr = 0.5f - c[nc - kk], wki = c[kk];
            float xr = a[j] - a[k], xi = a[j + 1] + a[k + 1];
            float yr = wkr * xr - wki * xi, yi = wkr * xi + wki * xr;
            a[j] -= yr;
            a[j + 1] -= yi;
            a[k] += yr;
            a[k + 1] -= yi;
        }
        {
            float d = a[0] - a[1];
            a[0] += a[1];
            a[1] = d;
        }
    } else {
        int ks = 2 * nc / m;
        a[1] = 0.5f * (a[0] - a[1]);
        a[0] -= a[1];
        a[1] = -a[1];                               /* rftbsub */
        for (j = 2; j < m; j += 2) {
            int k = n - j, kk = ks * (j >> 1);
            float wkr = 0.5f - c[nc - kk], wki = c[kk];
            float xr = a[j] - a[k], xi = a[j + 1] + a[k + 1];
            float yr = wkr * xr + wki * xi, yi = wkr * xi - wki * xr;
            a[j] -= yr;
            a[j + 1] = yi - a[j + 1];
            a[k] += yr;
            a[k + 1] = yi - a[k + 1];
        }
        a[m + 1] = -a[m + 1];
        orc_bitrev_pairs(n, a);
        orc_complex_passes(n, a, w, 1);
    }
}
