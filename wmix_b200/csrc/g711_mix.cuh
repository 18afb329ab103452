// G.711 A-law / mu-law and the int16 mix primitives, per sample.
// Bit-exact with R:src/g711codec.c and R:src/wmix.c:1617-1702 (see the quirk list in
// SURVEY.md §8 a24): branch-free segment search via clz instead of the reference's table walk.
#pragma once
#include "common.cuh"

namespace wmx {

// segment number = index of the first bound 0xFF,0x1FF,..,0x7FFF that `v` does not exceed
// (R:src/g711codec.c:9-21).  Negative v (A-law of -7..-1) lands in segment 0, > 0x7FFF in 8.
WMX_HD int g711_segment(int v)
{
    if (v <= 0xFF) return 0;
    if (v > 0x7FFF) return 8;
    return 24 - clz32((uint32_t)v);   // 0x100..0x1FF -> 1, ... 0x4000..0x7FFF -> 7
}

// R:src/g711codec.c:82-118
WMX_HD uint8_t linear2alaw(int pcm)
{
    int flip = 0xD5;
    if (pcm < 0) { flip = 0x55; pcm = -pcm - 8; }
    int s = g711_segment(pcm);
    if (s >= 8) return (uint8_t)(0x7F ^ flip);
    int q = (s < 2) ? (pcm >> 4) : (pcm >> (s + 3));      // arithmetic >> keeps -7..-1 negative
    return (uint8_t)(((s << 4) | (q & 0xF)) ^ flip);
}

// R:src/g711codec.c:120-152
WMX_HD uint8_t linear2ulaw(int pcm)
{
    int flip;
    if (pcm < 0) { pcm = 0x84 - pcm; flip = 0x7F; } else { pcm += 0x84; flip = 0xFF; }
    int s = g711_segment(pcm);
    if (s >= 8) return (uint8_t)(0x7F ^ flip);
    return (uint8_t)(((s << 4) | ((pcm >> (s + 3)) & 0xF)) ^ flip);
}

// R:src/g711codec.c:28-50
WMX_HD int16_t alaw2linear(uint8_t code)
{
    int a = code ^ 0x55;
    int t = (a & 0x0F) << 4;
    int s = (a & 0x70) >> 4;
    t += (s == 0) ? 8 : 0x108;
    if (s > 1) t <<= (s - 1);
    return (int16_t)((a & 0x80) ? t : -t);
}

// R:src/g711codec.c:61-76
WMX_HD int16_t ulaw2linear(uint8_t code)
{
    int u = (uint8_t)~code;
    int t = (((u & 0x0F) << 3) + 0x84) << ((u & 0x70) >> 4);
    return (int16_t)((u & 0x80) ? (0x84 - t) : (t - 0x84));
}

// R:src/wmix.c:1617-1636 (the zero short-cuts cannot change the value: x+0 never clips)
WMX_HD int16_t volume_add(int16_t a, int16_t b) { return sat16((int32_t)a + (int32_t)b); }

// one step of the same-format branch of wmix_load_data (R:src/wmix.c:1686): C division
// truncates toward zero, so -4786/3 == -1595
WMX_HD int16_t mix_step(int16_t bus, int16_t src, int rdce) { return volume_add(bus, (int16_t)(src / rdce)); }

}  // namespace wmx
