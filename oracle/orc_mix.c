/* ORACLE (test infrastructure) — the int16 mix bus, restating R:src/wmix.c:1617-1702. */
#include "oracle.h"

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
