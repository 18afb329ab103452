/* Drop-in for the reference's speech-processing handle API, R:src/webrtc.h:32-61
 * (implemented in R:src/webrtc.c by wrapping the WebRTC C libraries).  Same names, same
 * argument meaning, same error behaviour; the work runs on the GPU through wmixb (see wmixb.h)
 * with one stream per handle.  A daemon that wants the batched throughput calls wmixb directly;
 * this header exists so wmix's own call sites (R:src/wmix.c:565-710, :1372-1385) link unchanged.
 *
 *   frameNum  = frames (samples per channel) and must be a multiple of the handle's packet
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

/* VAD — R:src/webrtc.h:32-36, R:src/webrtc.c:40-167 */
void *vad_init(int chn, int freq, int intervalMs, bool *debug);
void vad_process(void *fp, int16_t *frame, int frameNum);
void vad_release(void *fp);

/* AEC — R:src/webrtc.h:40-46, R:src/webrtc.c:217-505 */
void *aec_init(int chn, int freq, int intervalMs, bool *debug);
int aec_setFrameFar(void *fp, int16_t *frameFar, int frameNum);
int aec_process(void *fp, int16_t *frameNear, int16_t *frameOut, int frameNum, int delayms);
int aec_process2(void *fp, int16_t *frameFar, int16_t *frameNear, int16_t *frameOut, int frameNum, int delayms);
void aec_release(void *fp);

/* NS — R:src/webrtc.h:50-54, R:src/webrtc.c:560-660 */
void *ns_init(int chn, int freq, bool *debug);
void ns_process(void *fp, int16_t *frame, int16_t *frameOut, int frameNum);
void ns_release(void *fp);

/* AGC — R:src/webrtc.h:58-63, R:src/webrtc.c:694-857 */
void *agc_init(int chn, int freq, int intervalMs, int value, bool *debug);
int agc_process(void *fp, int16_t *frame, int16_t *frameOut, int frameNum);
void agc_addition(void *fp, uint8_t value);
void agc_release(void *fp);

#ifdef __cplusplus
}
#endif
#endif
