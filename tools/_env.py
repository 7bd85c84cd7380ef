"""Imported first by every bring-up script: selects the tools build of the library (prof hooks, debug switches, probe
kernels: include/m2tts_b200_tools.h) unless M2TTS_B200_LIB already points somewhere."""
import os
from pathlib import Path

_TOOLS_LIB = Path(__file__).resolve().parents[1] / "m2-tts_b200" / "lib" / "libm2tts_b200_tools.so"
os.environ.setdefault("M2TTS_B200_LIB", str(_TOOLS_LIB))
