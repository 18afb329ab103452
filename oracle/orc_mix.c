/* ORACLE (test infrastructure) — the int16 mix bus, restating R:src/wmix.c:1617-1702. */
#include "oracle.h"
#include <string.h>

/* R:src/wmix.c:1617-1636.  The zero short-circuits are value-neutral (x+0 never clips) but
 * are kept so the statement reads like the reference. */
int16_t orc_volume_add(int16_t a, int16_t b)
{
    int32_t s;
    if (a == 0)
        return b;
    if (b == 0)
        return a;
    s = (int32_t)a + (int32_t)b;
    if (s > 32767)
        s = 32767;
    if (s < -32768)
        s = -32768;
    return (int16_t)s;
}

/* Same-format branch of wmix_load_data (R:src/wmix.c:1678-1702): the source sample is first
 * divided by the background-reduce factor with C integer division (truncation toward zero,
 * so -4786/3 = -1595), then saturating-added into the ring, which wraps at its end. */
uint32_t orc_mix_same_format(int16_t *ring, uint32_t ring_len, uint32_t pos,
                             const int16_t *src, uint32_t n, uint8_t rdce)
{
    uint32_t i;
    int d = rdce ? rdce : 1;
    for (i = 0; i < n; ++i) {
        ring[pos] = orc_volume_add(ring[pos], (int16_t)(src[i] / d));
        if (++pos >= ring_len)
            pos = 0;
    }
    return pos;
}

/* Conference bus as an exact int32 sum (associative, so it can be all-reduced; equals the
 * chained volumeAdd whenever no partial sum leaves the int16 range — SURVEY.md §8e). */
void orc_bus_sum(int32_t *bus, const int16_t *pcm, int n_part, int frame)
{
    int p, i;
    for (i = 0; i < frame; ++i)
        bus[i] = 0;
    for (p = 0; p < n_part; ++p)
        for (i = 0; i < frame; ++i)
            bus[i] += pcm[(size_t)p * frame + i];
}

void orc_bus_nminus1(int16_t *out, const int32_t *bus, const int16_t *own, int frame)
{
    int i;
    for (i = 0; i < frame; ++i) {
        int32_t v = bus[i] - own[i];
        out[i] = (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v));
    }
}

/* ---- nearest-sample rate / channel conversion, restating R:src/wmix.c:49-222 ----
 * All three reference functions walk the same float phase accumulator: the slower side advances when
 * (int)acc > 0, after which 1.0 is subtracted in DOUBLE and stored back to float.  Lengths are whatever
 * unit the caller uses for wmix_len_of_* (the reference adds the channel count per step) and bytes for
 * wmix_pcm_zoom.  Quirk kept: the stereo->stereo copy is dead code there (its test repeats 0x12,
 * R:src/wmix.c:178, :212), so a 2ch->2ch rate change writes nothing and returns 0. */
static float orc_zoom_ratio(uint16_t in_freq, uint16_t out_freq, int *up)
{
    *up = in_freq < out_freq;
    return *up ? (float)in_freq / out_freq : (float)out_freq / in_freq;   /* smaller over larger */
}

static int orc_zoom_tick(float *acc, float div)
{
    *acc += div;
    if ((int)*acc > 0) {
        *acc -= 1.0;
        return 1;
    }
    return 0;
}

/* R:src/wmix.c:49-91 */
uint32_t orc_len_of_out(uint8_t in_chn, uint16_t in_freq, uint32_t in_len, uint8_t out_chn, uint16_t out_freq)
{
    uint32_t in_n = 0, out_n = 0;
    float acc = 0;
    int up;
    float div;
    if (in_freq == out_freq && in_chn == out_chn)
        return in_len;
    div = orc_zoom_ratio(in_freq, out_freq, &up);
    while (in_n < in_len) {
        if (up) {
            out_n += out_chn;
            if (orc_zoom_tick(&acc, div))
                in_n += in_chn;
        } else {
            if (orc_zoom_tick(&acc, div))
                out_n += out_chn;
            in_n += in_chn;
        }
    }
    return out_n;
}

/* R:src/wmix.c:94-136 */
uint32_t orc_len_of_in(uint8_t in_chn, uint16_t in_freq, uint8_t out_chn, uint16_t out_freq, uint32_t out_len)
{
    uint32_t in_n = 0, out_n = 0;
    float acc = 0;
    int up;
    float div;
    if (in_freq == out_freq && in_chn == out_chn)
        return out_len;
    div = orc_zoom_ratio(in_freq, out_freq, &up);
    while (out_n < out_len) {
        if (up) {
            out_n += out_chn;
            if (orc_zoom_tick(&acc, div))
                in_n += in_chn;
        } else {
            if (orc_zoom_tick(&acc, div))
                out_n += out_chn;
            in_n += in_chn;
        }
    }
    return in_n;
}

/* one output "frame" for the channel pairing (R:src/wmix.c:160-176 / :194-210); returns samples written */
static int orc_zoom_emit(int mode, const int16_t *src, int16_t *dst)
{
    switch (mode) {
    case 0x11: dst[0] = src[0]; return 1;
    case 0x12: dst[0] = src[0]; dst[1] = src[0]; return 2;
    case 0x21: dst[0] = src[0]; return 1;           /* left channel only */
    default: return 0;                               /* 0x22: never reached in the reference */
    }
}

/* R:src/wmix.c:139-222.  in_len and the return value are bytes. */
uint32_t orc_pcm_zoom(uint8_t in_chn, uint16_t in_freq, const uint8_t *in, uint32_t in_len, uint8_t out_chn,
                      uint16_t out_freq, uint8_t *out)
{
    const int16_t *p = (const int16_t *)in, *end = (const int16_t *)(in + in_len);
    int16_t *q = (int16_t *)out;
    const int mode = (in_chn << 4) | (out_chn & 0x0F);
    float acc = 0, div;
    int up;
    if (in_freq == out_freq && in_chn == out_chn) {
        memcpy(out, in, in_len);
        return in_len;
    }
    div = orc_zoom_ratio(in_freq, out_freq, &up);
    while (p < end) {
        if (up) {
            q += orc_zoom_emit(mode, p, q);
            if (orc_zoom_tick(&acc, div))
                p += in_chn;
        } else {
            if (orc_zoom_tick(&acc, div))
                q += orc_zoom_emit(mode, p, q);
            p += in_chn;
        }
    }
    return (uint32_t)((uint8_t *)q - out);
}

/* ---- different-format branches of wmix_load_data, MONO mix bus (R:src/wmix.c:1704-1939) ----
 * Only 16-bit sources do anything there (the 8- and 32-bit cases are empty).  One float phase accumulator:
 *   source faster than the bus (R:src/wmix.c:1707-1789): gains (freq - bus)/bus per copied frame; while it is
 *     >= 1.0 the next source frame is skipped and 1.0 is taken off (in double, stored to float);
 *   source slower or only the channel count differs (R:src/wmix.c:1791-1925): gains (bus - freq)/freq per
 *     copied frame; while it is >= 1.0 the bus is fed from a linear ramp between the frame just copied and
 *     the next one: n = (int)acc + 1 steps of (float)(next - last) / n, accumulated in float and added to
 *     the int16 `last` (float arithmetic, truncated on the store to int16).
 * A stereo source contributes its left sample only (bus is mono).  Every written sample goes through
 * volumeAdd(bus, value / rdce) like the same-format branch.  The reference computes a ramp after the very
 * last source frame too (reading one sample past the buffer) but its loop ends before using it; that read is
 * skipped here.  Returns the new position; *written = samples added to the bus. */
uint32_t orc_mix_resample(int16_t *ring, uint32_t ring_len, uint32_t pos, const int16_t *src, uint32_t src_bytes,
                          uint16_t freq, uint8_t channels, uint16_t mix_freq, uint8_t rdce, uint32_t *written)
{
    const int d = rdce ? rdce : 1;
    const uint32_t frame_bytes = 2u * channels;
    uint32_t used = 0, n_out = 0;
    const int16_t *p = src;
    float acc = 0, pow_;
    int16_t ramp[64];
    int ramp_i = 0;
    if (written)
        *written = 0;
    if (channels != 1 && channels != 2)
        return pos;
    if (freq > mix_freq) {
        pow_ = (float)(freq - mix_freq) / mix_freq;
        while (used < src_bytes) {
            if (acc >= 1.0) {
                acc -= 1.0;
            } else {
                ring[pos] = orc_volume_add(ring[pos], (int16_t)(p[0] / d));
                if (++pos >= ring_len)
                    pos = 0;
                ++n_out;
                acc += pow_;
            }
            p += channels;
            used += frame_bytes;
        }
    } else {
        pow_ = (float)(mix_freq - freq) / freq;
        while (used < src_bytes) {
            if (acc >= 1.0) {
                ring[pos] = orc_volume_add(ring[pos], (int16_t)(ramp[ramp_i] / d));
                ++ramp_i;
                acc -= 1.0;
            } else {
                ring[pos] = orc_volume_add(ring[pos], (int16_t)(p[0] / d));
                p += channels;
                used += frame_bytes;
                acc += pow_;
                if (acc >= 1.0 && used < src_bytes) {
                    const int n = (int)acc + 1;
                    const int16_t last = p[-(int)channels];
                    const float step = (float)(p[0] - last) / n;
                    float run = step;
                    int k;
                    for (k = 0; k < n && k < 64; ++k) {
                        ramp[k] = (int16_t)(last + run);
                        run += step;
                    }
                    ramp_i = 0;
                }
            }
            if (++pos >= ring_len)
                pos = 0;
            ++n_out;
        }
    }
    if (written)
        *written = n_out;
    return pos;
}

/* ---- play-package FIFO that feeds the echo canceller its far end (R:src/wmix.c:482-526) ----
 * playPkgBuff_add stores the package the play thread just wrote to the sound card (R:src/wmix.c:1419) in a ring of
 * n_pkg = AEC_INTERVALMS / WMIX_INTERVAL_MS + 2 packages (R:src/wmixConf.h:141); playPkgBuff_get(AEC_INTERVALMS)
 * then hands aec_process2 its far end (R:src/wmix.c:653).  Restated for delays that are whole packages (the only
 * case wmix uses: 400 ms / 20 ms); with a remainder the reference copies from before the ring row (R:src/wmix.c:510-516).
 * The index arithmetic is kept as written, including its effect: `count - clamp(count - d)` is d while count >= d
 * and count itself otherwise, so the slot read is the OLDEST package (the one the next add overwrites) except on
 * the ticks where count has just passed d. */
void orc_play_fifo_init(orc_play_fifo *f, int n_pkg, int pkg_bytes)
{
    memset(f, 0, sizeof(*f));
    f->n_pkg = n_pkg > ORC_FIFO_MAX_PKG ? ORC_FIFO_MAX_PKG : n_pkg;
    f->pkg_bytes = pkg_bytes > ORC_FIFO_MAX_BYTES ? ORC_FIFO_MAX_BYTES : pkg_bytes;
}

void orc_play_fifo_add(orc_play_fifo *f, const uint8_t *pkg)
{
    memcpy(f->buf[f->count++], pkg, (size_t)f->pkg_bytes);
    if (f->count >= f->n_pkg)
        f->count = 0;
}

int orc_play_fifo_slot(int count, int n_pkg, int delay_pkgs)
{
    int k = count - delay_pkgs;
    if (k >= n_pkg)
        k = n_pkg;
    else if (k < 0)
        k = 0;
    k = count - k;
    if (k >= n_pkg)
        k -= n_pkg;
    else if (k < 0)
        k += n_pkg;
    return k;
}

void orc_play_fifo_get(const orc_play_fifo *f, uint8_t *out, int delay_pkgs)
{
    memcpy(out, f->buf[orc_play_fifo_slot(f->count, f->n_pkg, delay_pkgs)], (size_t)f->pkg_bytes);
}

/* ---- the whole of wmix_load_data for a mono 16-bit bus (R:src/wmix.c:1639-1956): bookkeeping around the adds ----
 * The six WMix_Struct fields it reads are passed as a view (byte OFFSETS into the ring instead of pointers; head_off /
 * the return value < 0 stand for a NULL head).  Order of business, as in the reference:
 *   (1) nothing happens without a running mixer, a source or at least one byte (:1664);
 *   (2) a producer with no head yet, or whose tick fell behind the play pointer's, restarts play_correct bytes ahead of
 *       the play pointer (:1667-1674) — and is put at the ring START, not wrapped, if that lands on or past the end;
 *   (3) its samples are divided by reduceMode unless its own `reduce` equals reduceMode (:1676-1677);
 *   (4) same format: plain adds; 16-bit mono / stereo of another rate: the drop / ramp branches; anything else: no adds;
 *   (5) the producer's tick advances by the bytes written (:1942-1953; the "fell behind" branch there cannot fire when
 *       the view is a snapshot, since step (2) just made *tick >= view.tick). */
int32_t orc_wmix_load_data(const orc_mix_view *w, uint8_t *ring, const uint8_t *src, uint32_t src_bytes, uint16_t freq,
                           uint8_t channels, uint8_t sample, int32_t head_off, uint8_t reduce, uint32_t *tick)
{
    uint32_t written = 0, pos;
    uint8_t d;
    if (!w || !w->run || !src || src_bytes < 1)
        return head_off;
    if (head_off < 0 || *tick < w->tick) {
        head_off = (int32_t)(w->head_off + w->play_correct);
        *tick = w->tick + w->play_correct;
        if ((uint32_t)head_off >= w->ring_bytes)
            head_off = 0;
    }
    d = (reduce == w->reduce_mode) ? 1 : w->reduce_mode;
    pos = (uint32_t)head_off / 2;
    if (freq == w->mix_freq && channels == 1 && sample == 16) {
        pos = orc_mix_same_format((int16_t *)ring, w->ring_bytes / 2, pos, (const int16_t *)src, src_bytes / 2, d);
        written = src_bytes / 2;
    } else if (sample == 16 && (channels == 1 || channels == 2)) {
        pos = orc_mix_resample((int16_t *)ring, w->ring_bytes / 2, pos, (const int16_t *)src, src_bytes, freq, channels,
                               w->mix_freq, d, &written);
    }
    *tick += written * 2;
    return (int32_t)(pos * 2);
}
