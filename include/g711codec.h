/* Drop-in for R:src/g711codec.h:24-34 — host-buffer G.711 entry points, computed on the GPU.
 * Argument quirks are the reference's: PCM2G711* take the PCM size in BYTES and return the
 * number of codes written; G711*2PCM take the number of codes and return BYTES written
 * (R:src/g711codec.c:227-308).  `reserve` is ignored.  Returns -1 when all of in/out/len are
 * zero (the reference's only argument check) or when the GPU call fails.
 * For device-resident buffers use wmixb_g711_{encode,decode}_device (wmixb.h). */
#ifndef WMIX_B200_G711CODEC_H
#define WMIX_B200_G711CODEC_H
#ifdef __cplusplus
extern "C" {
#endif
/* encoders: pcm_bytes of int16 PCM in -> pcm_bytes / 2 codes out; the return value is the number of codes */
int PCM2G711a(char *pcm, char *codes, int pcm_bytes, int reserve);
int PCM2G711u(char *pcm, char *codes, int pcm_bytes, int reserve);
/* decoders: n_codes codes in -> n_codes int16 samples out; the return value is the number of BYTES written */
int G711a2PCM(char *codes, char *pcm, int n_codes, int reserve);
int G711u2PCM(char *codes, char *pcm, int n_codes, int reserve);
/* the array forms underneath them (same units: samples in, samples / bytes out as above) */
int g711a_encode(unsigned char codes[], const short pcm[], int n_samples);
int g711u_encode(unsigned char codes[], const short pcm[], int n_samples);
int g711a_decode(short pcm[], const unsigned char codes[], int n_codes);
int g711u_decode(short pcm[], const unsigned char codes[], int n_codes);
#ifdef __cplusplus
}
#endif
#endif
