#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds the *unmodified* reference (wexiangis/wmix) hot path
# into oracle/_ref/libwmix_ref.so so the C restatement under oracle/ and the CUDA product
# can be checked against the real thing.  Nothing under wmix_b200/ may link or load this.
#
# What goes in (all compiled from where they lie; no reference source is copied into the repo):
#   * /root/reference/pkg/webrtc_cut.tar.gz  -> untarred into a scratch dir (default
#     /tmp/wmix_ref_build); the same *.c sets T:build_{vad,ns,agc,aec}_so.sh glob, minus the
#     mips/neon variants those scripts also exclude.
#   * /root/reference/src/webrtc.c      (the handle API, R:src/webrtc.c:40-860)
#   * /root/reference/src/g711codec.c   (G.711, R:src/g711codec.c)
#   * /root/reference/src/{wmix,wmixTask,wmixMem,wav,rtp,delay}.c  for the real
#     wmix_load_data (R:src/wmix.c:1639); main() renamed, codecs compiled out, platform
#     audio replaced by oracle/ref_shim/plat.h + plat_stub.c (ours).
# -O2 is used: SURVEY.md §8c records that -O0 (as shipped) and -O2 are bit-identical.
# The AEC is pinned to its plain-C kernels by oracle/ref_shim/pin_c_path.c.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${WMIX_REFERENCE:-/root/reference}"
SCRATCH="${WMIX_REF_SCRATCH:-/tmp/wmix_ref_build}"
OUT="$HERE/_ref"
CC="${CC:-gcc}"
OPT="${WMIX_REF_OPT:--O2}"
SUFFIX="${WMIX_REF_SUFFIX:-}"

if [ ! -f "$REF/pkg/webrtc_cut.tar.gz" ]; then
    echo "build_ref: $REF not present (GPU box?) - keeping any prebuilt $OUT/libwmix_ref.so" >&2
    exit 0
fi
mkdir -p "$SCRATCH" "$OUT"
if [ ! -d "$SCRATCH/webrtc_cut/webrtc" ]; then
    tar -xzf "$REF/pkg/webrtc_cut.tar.gz" -C "$SCRATCH"
fi
W="$SCRATCH/webrtc_cut"

srcs=()
add_dir() {
    local f
    for f in "$1"/*.c; do
        case "$(basename "$f")" in
            *_mips.c|*_neon.c) ;;
            *) srcs+=("$f") ;;
        esac
    done
}
add_dir "$W/webrtc/common_audio"
add_dir "$W/webrtc/common_audio/vad"
add_dir "$W/webrtc/common_audio/signal_processing"
add_dir "$W/webrtc/modules/audio_processing/ns"
add_dir "$W/webrtc/modules/audio_processing/agc/legacy"
add_dir "$W/webrtc/modules/audio_processing/aec"
add_dir "$W/webrtc/modules/audio_processing/aecm"
add_dir "$W/webrtc/modules/audio_processing/utility"

INC=(-I"$W" -I"$W/webrtc/common_audio/vad/include"
     -I"$W/webrtc/modules/audio_processing/ns/include"
     -I"$W/webrtc/modules/audio_processing/agc/legacy"
     -I"$W/webrtc/modules/audio_processing/aec/include"
     -I"$W/webrtc/modules/audio_processing/aecm/include")
CFLAGS=($OPT -fPIC -ffp-contract=off -w -DWEBRTC_POSIX)

OBJ="$SCRATCH/obj$SUFFIX"
mkdir -p "$OBJ"
objs=()
i=0
for s in "${srcs[@]}"; do
    o="$OBJ/w$i.o"; i=$((i+1))
    "$CC" "${CFLAGS[@]}" "${INC[@]}" -c "$s" -o "$o" &
    objs+=("$o")
    if (( i % 16 == 0 )); then wait; fi
done
wait
# cpu_features.cc is plain C in a .cc file; the reference script hands it to gcc as-is.
g++ "${CFLAGS[@]}" "${INC[@]}" -c "$W/webrtc/system_wrappers/source/cpu_features.cc" -o "$OBJ/cpu.o"
objs+=("$OBJ/cpu.o")

# wmix's own layer: handle API + G.711 + the mix entry point.
WM=(-DMAKE_MP3=0 -DMAKE_AAC=0 -DMAKE_SPEEX=0 -DMAKE_SPEEX_BETA3=0 -DMAKE_MATH_FFT=0
    -DORACLE_PLAT_FREQ="${ORACLE_PLAT_FREQ:-16000}" -I"$HERE/ref_shim" -I"$REF/src")
"$CC" "${CFLAGS[@]}" "${INC[@]}" "${WM[@]}" -c "$REF/src/webrtc.c" -o "$OBJ/r_webrtc.o"
"$CC" "${CFLAGS[@]}" "${WM[@]}" -c "$REF/src/g711codec.c" -o "$OBJ/r_g711.o"
"$CC" "${CFLAGS[@]}" "${WM[@]}" -Dmain=wmix_main -c "$REF/src/wmix.c" -o "$OBJ/r_wmix.o"
for f in wmixTask wmixMem wav rtp delay; do
    "$CC" "${CFLAGS[@]}" "${WM[@]}" -c "$REF/src/$f.c" -o "$OBJ/r_$f.o"
done
"$CC" "${CFLAGS[@]}" "${INC[@]}" "${WM[@]}" -c "$HERE/ref_shim/plat_stub.c" -o "$OBJ/s_plat.o"
"$CC" "${CFLAGS[@]}" "${INC[@]}" -c "$HERE/ref_shim/pin_c_path.c" -o "$OBJ/s_pin.o"

g++ -shared -o "$OUT/libwmix_ref$SUFFIX.so" "${objs[@]}" "$OBJ"/r_*.o "$OBJ"/s_*.o -lpthread -lm
echo "build_ref: wrote $OUT/libwmix_ref$SUFFIX.so ($OPT)"

# The same library with the reference's own NS switch thrown (R:src/webrtc.c:511-523: `#define MAKE_WEBRTC_NSX`,
# commented out as shipped): ns_init / ns_process then run the fixed-point WebRtcNsx_* core.  Only webrtc.c differs.
"$CC" "${CFLAGS[@]}" "${INC[@]}" "${WM[@]}" -DMAKE_WEBRTC_NSX -c "$REF/src/webrtc.c" -o "$OBJ/x_webrtc_nsx.o"
nsx_objs=()
for o in "$OBJ"/r_*.o; do
    [ "$o" = "$OBJ/r_webrtc.o" ] || nsx_objs+=("$o")
done
g++ -shared -o "$OUT/libwmix_ref_nsx$SUFFIX.so" "${objs[@]}" "${nsx_objs[@]}" "$OBJ/x_webrtc_nsx.o" "$OBJ"/s_*.o -lpthread -lm
echo "build_ref: wrote $OUT/libwmix_ref_nsx$SUFFIX.so (MAKE_WEBRTC_NSX)"
