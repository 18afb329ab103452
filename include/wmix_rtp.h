/* RTP / G.711 wire framing for batches of legs: the step either side of the conference
 * mix.  The reference handles one leg per thread: rtp_recv -> G711a2PCM -> wmix_load_data
 * (R:src/wmixTask.c:1278-1311) and wmix_pcm_zoom -> PCM2G711a -> timestamp += samples -> rtp_send -> seq++
 * (R:src/wmixTask.c:1137-1143, R:src/rtp.c:35-70), with the 12-byte header of R:src/rtp.h:51-70:
 *
 *   byte 0  V(2) P(1) X(1) CC(4)      byte 1  M(1) PT(7)      bytes 2-3 sequence (big endian)
 *   bytes 4-7 timestamp (big endian)  bytes 8-11 SSRC (big endian)            then the payload
 *
 * Sockets stay on the host (out of scope).  What is batched here is the byte work: a host receives N datagrams
 * into one pinned slab (one slot of `stride` bytes per leg), copies the slab to the device once, and
 *   wmixb_rtp_unpack_device  validates and parses every header and gathers the payloads into the dense
 *                            codes matrix [n][payload] that wmixb_g711_bus_sum / wmixb_peer_bus_tick consume;
 *   wmixb_rtp_pack_device    is the reverse for the egress legs: advances each leg's timestamp and sequence
 *                            number exactly as the reference's send loop does and writes header + payload.
 * Like the reference's receiver (R:src/rtp.c:89-91) a PCMA / PCMU packet is taken to carry exactly
 * WMIXB_RTP_PCMA_PAYLOAD (160) payload bytes = 20 ms at 8 kHz mono, whatever its datagram length. */
#ifndef WMIX_B200_RTP_H
#define WMIX_B200_RTP_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define WMIXB_RTP_HEADER 12          /* RTP_HEADER_SIZE, R:src/rtp.h:35 */
#define WMIXB_RTP_PCMA_PAYLOAD 160   /* RTP_PCMA_PKT_SIZE, R:src/rtp.h:33 */
#define WMIXB_RTP_PT_PCMU 0
#define WMIXB_RTP_PT_PCMA 8

/* per-leg header fields in host byte order */
typedef struct wmixb_rtp_meta {
    uint32_t timestamp;
    uint32_t ssrc;
    uint16_t seq;
    uint8_t pt;        /* 7-bit payload type */
    uint8_t marker;
    uint8_t vpxcc;     /* byte 0 as received: V<<6 | P<<5 | X<<4 | CC */
    uint8_t ok;        /* 1: version 2, PT is PCMA/PCMU and the slot held at least header + 160 bytes */
    uint16_t reserved;
} wmixb_rtp_meta;      /* 16 bytes */

/* per-leg sender state (host byte order); what rtp_header(.., cc=0, x=0, p=0, v=2, pt, m=1, seq=0, ts=0, ssrc)
 * initialises in the reference (R:src/wmixTask.c:1058) */
typedef struct wmixb_rtp_state {
    uint32_t timestamp;
    uint32_t ssrc;
    uint16_t seq;
    uint8_t pt;
    uint8_t marker;
} wmixb_rtp_state;     /* 12 bytes */

/* host helpers: one header <-> 12 bytes (what rtp_header + the htons/htonl of rtp_send put on the wire) */
void wmixb_rtp_write_header(uint8_t out[12], uint8_t vpxcc, uint8_t marker, uint8_t pt, uint16_t seq, uint32_t timestamp,
                            uint32_t ssrc);
void wmixb_rtp_read_header(const uint8_t in[12], wmixb_rtp_meta* meta);

/* d_slab: n slots of `stride` bytes (stride >= 172, multiple of 4), each holding one datagram from byte 0;
 * d_sizes (nullable): received length of each datagram.  d_codes: uint8 [n][160]; a leg that is not ok gets
 * the codec's silence (0xD5 A-law / 0xFF mu-law for `law_fill` 0 / 1).  d_meta: [n]. */
int wmixb_rtp_unpack_device(const uint8_t* d_slab, const int32_t* d_sizes, int n, int stride, int law_fill, uint8_t* d_codes,
                            wmixb_rtp_meta* d_meta, void* stream);
/* d_codes: uint8 [n][160]; d_state [n] is advanced in place: timestamp += 160 / chn BEFORE the header is written,
 * seq += 1 AFTER (R:src/wmixTask.c:1141-1143, R:src/rtp.c:68); the marker bit is sent as stored (the reference
 * never clears it).  d_slab: n slots of `stride` bytes receive header + payload (172 bytes each). */
int wmixb_rtp_pack_device(const uint8_t* d_codes, int n, int chn, wmixb_rtp_state* d_state, uint8_t* d_slab, int stride,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif
