/* ORACLE (test infrastructure) — G.711 A-law / mu-law, restating R:src/g711codec.c.
 * Quirks kept on purpose because bit-parity needs them:
 *   - A-law of a negative sample uses (-pcm - 8) and then an *arithmetic* shift, so the
 *     value can stay negative for pcm in [-7,-1]  (R:src/g711codec.c:94-97, :112-115);
 *   - PCM2G711* take a BYTE count and return samples, G711*2PCM take a code count and
 *     return BYTES (R:src/g711codec.c:227-308, :154-190).
 */
#include "oracle.h"

/* segment = index of the first upper bound 0xFF,0x1FF,...,0x7FFF that val fits under
 * (R:src/g711codec.c:9-21 seg_end/search) */
static int orc_segment(int val)
{
    int s;
    for (s = 0; s < 8; ++s)
        if (val <= ((0x100 << s) - 1))
            return s;
    return 8;
}

/* R:src/g711codec.c:82-118 */
uint8_t orc_linear2alaw(int pcm)
{
    int flip = 0xD5, s, q;
    if (pcm < 0) {
        flip = 0x55;
        pcm = -pcm - 8;
    }
    s = orc_segment(pcm);
    if (s >= 8)
        return (uint8_t)(0x7F ^ flip);
    q = (s < 2) ? (pcm >> 4) : (pcm >> (s + 3));
    return (uint8_t)((uint8_t)((s << 4) | (q & 0xF)) ^ flip);
}

/* R:src/g711codec.c:120-152 */
uint8_t orc_linear2ulaw(int pcm)
{
    int flip, s;
    if (pcm < 0) {
        pcm = 0x84 - pcm;
        flip = 0x7F;
    } else {
        pcm += 0x84;
        flip = 0xFF;
    }
    s = orc_segment(pcm);
    if (s >= 8)
        return (uint8_t)(0x7F ^ flip);
    return (uint8_t)((uint8_t)((s << 4) | ((pcm >> (s + 3)) & 0xF)) ^ flip);
}

/* R:src/g711codec.c:28-50 */
int16_t orc_alaw2linear(uint8_t code)
{
    int a = code ^ 0x55;
    int mant = (a & 0x0F) << 4;
    int s = (a & 0x70) >> 4;
    if (s == 0)
        mant += 8;
    else {
        mant += 0x108;
        if (s > 1)
            mant <<= (s - 1);
    }
    return (int16_t)((a & 0x80) ? mant : -mant);
}

/* R:src/g711codec.c:61-76 */
int16_t orc_ulaw2linear(uint8_t code)
{
    int u = (uint8_t)~code;
    int t = (((u & 0x0F) << 3) + 0x84) << ((u & 0x70) >> 4);
    return (int16_t)((u & 0x80) ? (0x84 - t) : (t - 0x84));
}

/* R:src/g711codec.c:227-243 -> g711a_encode :192-202 */
int orc_PCM2G711a(const char *in, char *out, int in_bytes)
{
    const int16_t *p = (const int16_t *)in;
    int n = in_bytes / 2, i;
    if (!in && !out && in_bytes == 0)
        return -1;
    for (i = 0; i < n; ++i)
        out[i] = (char)orc_linear2alaw(p[i]);
    return n;
}

/* R:src/g711codec.c:246-262 */
int orc_PCM2G711u(const char *in, char *out, int in_bytes)
{
    const int16_t *p = (const int16_t *)in;
    int n = in_bytes / 2, i;
    if (!in && !out && in_bytes == 0)
        return -1;
    for (i = 0; i < n; ++i)
        out[i] = (char)orc_linear2ulaw(p[i]);
    return n;
}

/* R:src/g711codec.c:273-289 -> g711a_decode :154-171 (returns samples*2) */
int orc_G711a2PCM(const char *in, char *out, int n_codes)
{
    int16_t *p = (int16_t *)out;
    int i;
    if (!in && !out && n_codes == 0)
        return -1;
    for (i = 0; i < n_codes; ++i)
        p[i] = orc_alaw2linear((uint8_t)in[i]);
    return (n_codes > 0 ? n_codes : 0) * 2;
}

/* R:src/g711codec.c:292-308 */
int orc_G711u2PCM(const char *in, char *out, int n_codes)
{
    int16_t *p = (int16_t *)out;
    int i;
    if (!in && !out && n_codes == 0)
        return -1;
    for (i = 0; i < n_codes; ++i)
        p[i] = orc_ulaw2linear((uint8_t)in[i]);
    return (n_codes > 0 ? n_codes : 0) * 2;
}
