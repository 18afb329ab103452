/* Drop-in for the reference's speech-processing handle API, R:src/webrtc.h:32-61
 * (implemented in R:src/webrtc.c by wrapping the WebRTC C libraries).  Same names, same
 * argument meaning, same error behaviour; the work runs on the GPU through wmixb (see wmixb.h)
 * with one stream per handle.  A daemon that wants the batched throughput calls wmixb directly;
 * this header exists so wmix's own call sites (R:src/wmix.c:565-710, :1372-1385) link unchanged.
 *
 *   n_frames  = frames (samples per channel; the reference's frameNum), a multiple of the handle's packet
 *               (10 ms; VAD may use 20 ms at <= 16 kHz as in R:src/webrtc.c:56-65).
 *   in / out  may alias (wmix always passes the same buffer, R:src/wmix.c:624-625).
 *   *_init    returns NULL on an unsupported rate (R:src/webrtc.c:43, :220, :563, :711), when
 *             no CUDA device is usable, or for the cases listed as not yet covered in
 *             INTEGRATION.md.  `debug` is a borrowed flag read at call time; NULL = quiet.
 */
#ifndef WMIX_B200_WEBRTC_H
#define WMIX_B200_WEBRTC_H

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* handle = what *_init returned; pcm = interleaved int16, n_frames samples per channel */

/* VAD — R:src/webrtc.h:32-36, R:src/webrtc.c:40-167: mutes pcm in place while nobody speaks */
void *vad_init(int channels, int rate_hz, int interval_ms, bool *debug);
void vad_process(void *handle, int16_t *pcm, int n_frames);
void vad_release(void *handle);

/* NS — R:src/webrtc.h:50-54, R:src/webrtc.c:560-660 */
void *ns_init(int channels, int rate_hz, bool *debug);
void ns_process(void *handle, int16_t *pcm_in, int16_t *pcm_out, int n_frames);
void ns_release(void *handle);

/* AGC — R:src/webrtc.h:58-63, R:src/webrtc.c:694-857: gain_db is WebRTC's compressionGaindB */
void *agc_init(int channels, int rate_hz, int interval_ms, int gain_db, bool *debug);
int agc_process(void *handle, int16_t *pcm_in, int16_t *pcm_out, int n_frames);
void agc_addition(void *handle, uint8_t gain_db);
void agc_release(void *handle);

/* AEC — R:src/webrtc.h:40-46, R:src/webrtc.c:217-505: far = what the loudspeaker played, near = the microphone */
void *aec_init(int channels, int rate_hz, int interval_ms, bool *debug);
int aec_setFrameFar(void *handle, int16_t *far, int n_frames);
int aec_process(void *handle, int16_t *near, int16_t *pcm_out, int n_frames, int delay_ms);
int aec_process2(void *handle, int16_t *far, int16_t *near, int16_t *pcm_out, int n_frames, int delay_ms);
void aec_release(void *handle);

#ifdef __cplusplus
}
#endif
#endif
