"""wmix_b200 — B200-native, stream-batched implementation of wmix's speech-processing and mix
hot path (NS / AGC / VAD / G.711 / conference bus).  See DESIGN.md."""
from ._lib import AEC, AGC, NS, VAD, LIB_PATH, WmixError, lib  # noqa: F401
from .engine import Engine, g711_decode, g711_encode, kernel_launches, mix_load  # noqa: F401
