/* ORACLE (test infrastructure) — RTP framing of the G.711 legs, restating R:src/rtp.h:51-70 (header bit
 * fields), R:src/rtp.c:20-70 (rtp_header, the byte-order swaps and seq++ of rtp_send), R:src/rtp.c:72-99
 * (rtp_recv's fixed 160-byte PCMA/PCMU payload) and the send loop's timestamp rule R:src/wmixTask.c:1139-1143. */
#include "oracle.h"
#include <string.h>

/* the twelve bytes rtp_send puts on the wire for a header filled by rtp_header() (little-endian bit fields:
 * cc in the low nibble of byte 0, v in its top two bits; pt in the low seven bits of byte 1, m on top) */
void orc_rtp_header_bytes(uint8_t out[12], uint8_t cc, uint8_t x, uint8_t p, uint8_t v, uint8_t pt, uint8_t m,
                          uint16_t seq, uint32_t timestamp, uint32_t ssrc)
{
    out[0] = (uint8_t)((cc & 15) | ((x & 1) << 4) | ((p & 1) << 5) | ((v & 3) << 6));
    out[1] = (uint8_t)((pt & 127) | ((m & 1) << 7));
    out[2] = (uint8_t)(seq >> 8);
    out[3] = (uint8_t)seq;
    out[4] = (uint8_t)(timestamp >> 24);
    out[5] = (uint8_t)(timestamp >> 16);
    out[6] = (uint8_t)(timestamp >> 8);
    out[7] = (uint8_t)timestamp;
    out[8] = (uint8_t)(ssrc >> 24);
    out[9] = (uint8_t)(ssrc >> 16);
    out[10] = (uint8_t)(ssrc >> 8);
    out[11] = (uint8_t)ssrc;
}

/* one iteration of the PCMA send loop for one leg: timestamp advances by samples-per-channel, the packet is
 * emitted, then the sequence number advances.  state = {timestamp, ssrc, seq, pt, marker}. */
void orc_rtp_send_step(uint32_t *timestamp, uint32_t ssrc, uint16_t *seq, uint8_t pt, uint8_t marker, int chn,
                       const uint8_t *codes, int n_codes, uint8_t *packet)
{
    *timestamp += (uint32_t)(n_codes / chn);
    orc_rtp_header_bytes(packet, 0, 0, 0, 2, pt, marker, *seq, *timestamp, ssrc);
    memcpy(packet + 12, codes, (size_t)n_codes);
    *seq = (uint16_t)(*seq + 1);
}

/* receive side: returns the payload size rtp_recv reports (160 for PCMA / PCMU, else 0 — AAC is out of scope)
 * and parses the header fields to host order */
int orc_rtp_parse(const uint8_t *packet, uint16_t *seq, uint32_t *timestamp, uint32_t *ssrc, uint8_t *pt, uint8_t *marker)
{
    *pt = packet[1] & 127;
    *marker = packet[1] >> 7;
    *seq = (uint16_t)((packet[2] << 8) | packet[3]);
    *timestamp = ((uint32_t)packet[4] << 24) | ((uint32_t)packet[5] << 16) | ((uint32_t)packet[6] << 8) | packet[7];
    *ssrc = ((uint32_t)packet[8] << 24) | ((uint32_t)packet[9] << 16) | ((uint32_t)packet[10] << 8) | packet[11];
    return (*pt == 8 || *pt == 0) ? 160 : 0;
}
