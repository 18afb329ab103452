/* Nearest-sample rate / channel conversion of the wmix daemon (R:src/wmix.h:112-128, R:src/wmix.c:49-222):
 * the step between the record chain and the G.711 / RTP legs (R:src/wmix.c:736, R:src/wmixTask.c:1137) and
 * between file/FIFO producers and the mix bus (R:src/wmixTask.c:186-748).
 *
 * Same symbols, same units (bytes for wmix_pcm_zoom, the caller's unit for wmix_len_of_*), same float
 * phase accumulator and the same quirk (a stereo -> stereo rate change writes nothing and returns 0, because
 * the reference's 0x22 case can never match, R:src/wmix.c:178, :212).
 *
 * Which input sample lands at each output position depends only on the formats and the length, never on
 * the audio, so it is worked out once on the host (exactly as the reference walks its accumulator) and
 * kept as a device gather table; moving the samples — for any number of streams at once — is a CUDA
 * kernel.  There is no CPU copy path: without a CUDA device wmix_pcm_zoom returns 0. */
#ifndef WMIX_B200_ZOOM_H
#define WMIX_B200_ZOOM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* drop-in (R:src/wmix.h:112-128) */
uint32_t wmix_len_of_out(uint8_t inChn, uint16_t inFreq, uint32_t inLen, uint8_t outChn, uint16_t outFreq);
uint32_t wmix_len_of_in(uint8_t inChn, uint16_t inFreq, uint8_t outChn, uint16_t outFreq, uint32_t outLen);
uint32_t wmix_pcm_zoom(uint8_t inChn, uint16_t inFreq, uint8_t* in, uint32_t inLen, uint8_t outChn, uint16_t outFreq,
                       uint8_t* out);

/* batched: one plan per (formats, input length), then any number of streams per launch */
typedef struct wmixb_zoom wmixb_zoom;
int wmixb_zoom_create(int in_chn, int in_freq, uint32_t in_bytes, int out_chn, int out_freq, int device, wmixb_zoom** out);
void wmixb_zoom_destroy(wmixb_zoom* z);
uint32_t wmixb_zoom_out_bytes(const wmixb_zoom* z);            /* per stream; what wmix_pcm_zoom returns */
/* d_in: int16 [n_streams][in_bytes/2], d_out: int16 [n_streams][out_bytes/2]; asynchronous on `stream` */
int wmixb_zoom_device(const wmixb_zoom* z, const int16_t* d_in, int16_t* d_out, int n_streams, void* stream);
/* the gather table itself (host copy), for tests: h_map[k] = input sample index of output sample k */
int wmixb_zoom_map(const wmixb_zoom* z, int32_t* h_map);

#ifdef __cplusplus
}
#endif
#endif
