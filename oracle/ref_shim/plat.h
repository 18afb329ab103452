/* TEST INFRASTRUCTURE (oracle build only).
 * Stand-in for the reference's platform header (R:platform/alsa/plat.h:15-36) so that
 * R:src/wmix.c can be compiled without a sound card.  Only the constants matter to the
 * hot path: they set WMIX_FREQ / WMIX_CHN that wmix_load_data compares against
 * (R:src/wmix.c:1678-1680).  ORACLE_PLAT_FREQ is chosen by oracle/build_ref.sh. */
#ifndef ORACLE_REF_SHIM_PLAT_H
#define ORACLE_REF_SHIM_PLAT_H
#include <stdint.h>
#ifndef ORACLE_PLAT_FREQ
#define ORACLE_PLAT_FREQ 16000
#endif
#define PLAT_CHN 1
#define PLAT_SAMPLE 16
#define PLAT_FREQ ORACLE_PLAT_FREQ
#define PLAT_AEC_INTERVALMS 400
#define PLAT_PLAY_CORRECT (PLAT_CHN * PLAT_FREQ * 16 / 8 / 5)
void *plat_ao_init(int chn, int freq);
void *plat_ai_init(int chn, int freq);
int plat_ao_write(void *ao, uint8_t *data, int len);
int plat_ai_read(void *ai, uint8_t *data, int len);
void plat_ao_vol_set(void *ao, int vol);
void plat_ai_vol_set(void *ai, int vol);
int plat_ao_vol_get(void *ao);
int plat_ai_vol_get(void *ai);
void plat_ao_exit(void *ao);
void plat_ai_exit(void *ai);
#endif
