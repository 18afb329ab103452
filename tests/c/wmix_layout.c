/* offsets of the WMix_Struct prefix declared by include/wmix.h, for tests/test_oracle_pin.py */
#include <stddef.h>
#include "wmix.h"
size_t wmixh_off_start(void) { return offsetof(WMix_Struct, start); }
size_t wmixh_off_end(void) { return offsetof(WMix_Struct, end); }
size_t wmixh_off_head(void) { return offsetof(WMix_Struct, head); }
size_t wmixh_off_run(void) { return offsetof(WMix_Struct, run); }
size_t wmixh_off_tick(void) { return offsetof(WMix_Struct, tick); }
size_t wmixh_off_reduce(void) { return offsetof(WMix_Struct, reduceMode); }
size_t wmixh_sizeof(void) { return sizeof(WMix_Struct); }
