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
int PCM2G711a(char *InAudioData, char *OutAudioData, int DataLen, int reserve);
int PCM2G711u(char *InAudioData, char *OutAudioData, int DataLen, int reserve);
int G711a2PCM(char *InAudioData, char *OutAudioData, int DataLen, int reserve);
int G711u2PCM(char *InAudioData, char *OutAudioData, int DataLen, int reserve);
int g711a_decode(short amp[], const unsigned char g711a_data[], int g711a_bytes);
int g711u_decode(short amp[], const unsigned char g711u_data[], int g711u_bytes);
int g711a_encode(unsigned char g711_data[], const short amp[], int len);
int g711u_encode(unsigned char g711_data[], const short amp[], int len);
#ifdef __cplusplus
}
#endif
#endif
