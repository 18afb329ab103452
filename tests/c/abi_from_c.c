#include "webrtc.h"
#include "g711codec.h"
#include "wmix.h"
#include "wmix_zoom.h"
#include "wmix_rtp.h"
#include "wmixb.h"
#include <stdio.h>
int main(void)
{
    wmixb_config cfg = { .n_streams = 4, .freq = 16000, .stages = WMIXB_NS | WMIXB_AGC | WMIXB_VAD, .ns_policy = 2, .agc_gain_db = 5, .vad_mode = 3 };
    wmixb_engine *e = 0;
    int rc = wmixb_create(&cfg, &e);
    printf("wmixb_create without a GPU -> %d (%s)\n", rc, wmixb_last_error());
    void *h = ns_init(1, 16000, 0);
    printf("ns_init without a GPU -> %p\n", h);
    {
        /* the reference's compile-time suppressor switch (MAKE_WEBRTC_NSX) as a run-time call; a 32 kHz engine with the fixed-point core */
        wmixb_config c32 = { .n_streams = 2, .freq = 32000, .stages = WMIXB_NS, .ns_policy = 2, .ns_core = 1 };
        wmixb_engine *e32 = 0;
        if (wmixb_set_default_ns_core(1) != WMIXB_OK || wmixb_default_ns_core() != 1 || wmixb_set_default_ns_core(7) == WMIXB_OK) return 3;
        {
            void *hx = ns_init(2, 32000, 0);
            printf("ns_init (fixed-point core) without a GPU -> %p\n", hx);
            if (hx) ns_release(hx);
        }
        wmixb_set_default_ns_core(0);
        printf("wmixb_create(32 kHz, ns_core 1) without a GPU -> %d\n", wmixb_create(&c32, &e32));
        if (e32) wmixb_destroy(e32);
    }
    {
        /* the NCCL exchange opens libnccl at run time: a wrong path is an error code and a message, not a link failure */
        wmixb_nccl_bus *nb = 0;
        int nrc = wmixb_nccl_load("/nonexistent/libnccl.so.2");
        printf("wmixb_nccl_load(bad path) -> %d (%s)\n", nrc, wmixb_last_error());
        if (nrc != WMIXB_ENODEV || wmixb_nccl_bus_create(0, 0, 1, "x", &nb) != WMIXB_EINVAL || nb) return 4;
        wmixb_nccl_bus_destroy(0);
    }
    wmixb_mix_view v = {0};
    uint32_t tick = 0;
    printf("load_data on a stopped mixer -> %p\n", (void *)wmixb_load_data_host(&v, (const uint8_t *)"ab", 2, 16000, 1, 16, 0, 0, &tick));
    {
        /* the mix entry point under the reference's own prototype, on a daemon-shaped struct */
        int16_t ring[64] = {0}, pcm[8] = {1, 2, 3, 4, 5, 6, 7, 8};
        WMix_Struct wm = {0};
        WMix_Point src, head, ret;
        wm.start.S16 = ring; wm.end.S16 = ring + 64; wm.head.S16 = ring; wm.run = false; wm.reduceMode = 1;
        src.S16 = pcm; head.S16 = ring + 4;
        ret = wmix_load_data(&wm, src, sizeof pcm, 8000, 1, 16, head, 1, &tick);
        printf("wmix_load_data on a stopped mixer -> head %s\n", ret.U8 == head.U8 ? "unchanged" : "moved");
        if (ret.U8 != head.U8) return 2;
    }
    printf("len_of_out %u\n", wmix_len_of_out(1, 8000, 320, 1, 16000));
    return rc == WMIXB_ENODEV && !h ? 0 : 1;
}
