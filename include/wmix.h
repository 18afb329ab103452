/* wmix's mix entry point under its own symbol and signature (R:src/wmix.h:40-49, R:src/wmix.c:1639-1956), backed by
 * the GPU: the six call sites of the daemon's producers (R:src/wmixTask.c:85, :973, :1311, :1484, :1704, :1927) link
 * against this library unchanged.
 *
 * WMix_Point and the LEADING fields of WMix_Struct are restated from R:src/wmixConf.h:156-207 — a configuration
 * schema, not code: wmix_load_data reads run, start, end, head, tick and reduceMode (R:src/wmix.c:1664-1677) and nothing
 * behind reduceMode.  The daemon's own definition continues with control-plane fields this library never touches;
 * a daemon built against its own wmixConf.h passes its full struct, this prefix only pins the offsets (checked against
 * the compiled reference in tests/test_oracle_pin.py).  When the daemon's header was included first, its definitions
 * are used and only the prototypes below are added.
 *
 * WMIX_FREQ and VIEW_PLAY_CORRECT are build constants of the daemon (R:src/wmixPlat.h:16, :20); the library defaults
 * to the alsa platform's (8000 Hz, 3200 bytes = 200 ms, R:platform/alsa/plat.h:17-21) and wmix_load_data_config sets
 * them for another build.  The bus is mono 16-bit, like every platform of the reference (PLAT_CHN 1).
 * There is no CPU path: without a CUDA device the ring is left untouched and the incoming head is returned. */
#ifndef WMIX_B200_WMIX_H
#define WMIX_B200_WMIX_H
#include <stdbool.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#ifndef _WMIXCONF_H_
typedef union {
    int8_t *S8;
    uint8_t *U8;
    int16_t *S16;
    uint16_t *U16;
    int32_t *S32;
    uint32_t *U32;
} WMix_Point;

typedef struct {
    void *objAo, *objAi;
    uint8_t *buff;
    WMix_Point start, end;      /* ring bounds                                   */
    WMix_Point head, tail;      /* play pointer (head) of the ring               */
    bool run;
    uint8_t loopWord, loopWordRecord, loopWordFifo, loopWordRtp;
    uint32_t tick;              /* bytes the play pointer has walked since start */
    uint32_t thread_sys, thread_record, thread_play;
    bool playRun, recordRun;
    int shmemRun;
    int msg_key;                /* key_t in the daemon's header (int on Linux; strict C99 does not expose the name) */
    int msg_fd;
    uint8_t reduceMode;         /* background-reduce divisor, 1 = off            */
    /* ... the daemon's struct continues (R:src/wmixConf.h:208-232); not read here */
} WMix_Struct;
#endif

WMix_Point wmix_load_data(WMix_Struct *wmix, WMix_Point src, uint32_t srcU8Len, uint16_t freq, uint8_t channels,
                          uint8_t sample, WMix_Point head, uint8_t reduce, uint32_t *tick);

/* the daemon's build constants WMIX_FREQ / VIEW_PLAY_CORRECT (bytes) */
void wmix_load_data_config(uint16_t mix_freq, uint32_t play_correct_bytes);

#ifdef __cplusplus
}
#endif
#endif
