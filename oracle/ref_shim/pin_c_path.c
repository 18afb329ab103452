/* TEST INFRASTRUCTURE (oracle build only).
 * Pins the reference AEC to its portable C kernels: WebRtc_GetCPUInfo is a swappable
 * function pointer (T:webrtc/system_wrappers/source/cpu_features.cc:71-72) that
 * aec_core.c/aec_rdft.c consult once at init to install SSE2 variants
 * (T:.../aec/aec_core.c:1451-1455).  The SSE2 path uses a polynomial pow and differs from
 * the C path by a few LSB, so the oracle of record is the C path. */
#include "webrtc/system_wrappers/interface/cpu_features_wrapper.h"
static WebRtc_CPUInfo oracle_saved_probe;
__attribute__((constructor)) static void oracle_pin_plain_c(void)
{
    oracle_saved_probe = WebRtc_GetCPUInfo;
    WebRtc_GetCPUInfo = WebRtc_GetCPUInfoNoASM;
}
/* Tests may flip between the two builds the reference itself ships (C vs SSE2) to measure
 * how far the reference's own variants disagree.  Takes effect at the next aec_init. */
void oracle_ref_use_sse2(int on)
{
    WebRtc_GetCPUInfo = (on && oracle_saved_probe) ? oracle_saved_probe : WebRtc_GetCPUInfoNoASM;
}
