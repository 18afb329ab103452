/* TEST INFRASTRUCTURE (oracle build only): no-op sound card so R:src/wmix.c links. */
#include <string.h>
#include "plat.h"
static int dummy_ao, dummy_ai;
void *plat_ao_init(int chn, int freq) { (void)chn; (void)freq; return &dummy_ao; }
void *plat_ai_init(int chn, int freq) { (void)chn; (void)freq; return &dummy_ai; }
int plat_ao_write(void *ao, uint8_t *data, int len) { (void)ao; (void)data; return len; }
int plat_ai_read(void *ai, uint8_t *data, int len) { (void)ai; memset(data, 0, (size_t)len); return len; }
void plat_ao_vol_set(void *ao, int vol) { (void)ao; (void)vol; }
void plat_ai_vol_set(void *ai, int vol) { (void)ai; (void)vol; }
int plat_ao_vol_get(void *ao) { (void)ao; return 10; }
int plat_ai_vol_get(void *ai) { (void)ai; return 10; }
void plat_ao_exit(void *ao) { (void)ao; }
void plat_ai_exit(void *ai) { (void)ai; }

/* Tiny helpers so tests can poke the reference without knowing WMix_Struct's layout. */
#include "wmix.h"
size_t oracle_ref_sizeof_wmix(void) { return sizeof(WMix_Struct); }
/* Seat a caller-provided ring in a zeroed WMix_Struct the way wmix_init does
 * (R:src/wmix.c:1547-1560): start/end/head/tail, run=1, tick=0, reduceMode. */
void oracle_ref_wmix_seat(WMix_Struct *w, uint8_t *ring, uint32_t ring_bytes, uint8_t reduceMode,
                          uint32_t head_off, uint32_t tick)
{
    memset(w, 0, sizeof(*w));
    w->buff = ring;
    w->start.U8 = ring;
    w->end.U8 = ring + ring_bytes;
    w->head.U8 = ring + head_off;
    w->tail.U8 = ring + head_off;
    w->run = true;
    w->tick = tick;
    w->reduceMode = reduceMode;
}
int oracle_ref_wmix_freq(void) { return WMIX_FREQ; }
int oracle_ref_wmix_buff_size(void) { return WMIX_BUFF_SIZE; }
int oracle_ref_wmix_play_correct(void) { return VIEW_PLAY_CORRECT; }
int oracle_ref_wmix_pkg_size(void) { return WMIX_PKG_SIZE; }
int oracle_ref_wmix_aec_fifo_pkgs(void) { return AEC_FIFO_PKG_NUM; }
int oracle_ref_wmix_aec_interval_ms(void) { return AEC_INTERVALMS; }
int oracle_ref_wmix_interval_ms(void) { return WMIX_INTERVAL_MS; }
