"""DEVELOPER TOOL: pytest plugin / import hook that points kmerjs_b200._abi at the emulated library
(build/emu/libkmerjs_b200_emu.so).  Use:  tools/cuemu/run.sh -m gpu tests/...   Never used by the
driver's test runs, the bench or the package."""
import os

import kmerjs_b200.build as _b

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_b.LIB = os.path.join(_ROOT, "build", "emu", "libkmerjs_b200_emu.so")
_b.needs_build = lambda: False
os.environ["KMERJS_B200_EMU"] = "1"
