"""Host-side mirror of lib/stats.js: ``etta``, ``zScore``, ``fastp``.  The reference returns
bignumber.js objects; here they are ``decimal.Decimal`` built from the exact decimal text that
libkmerjs_b200.so computes (kj_stats_zscore / kj_stats_fastp_text, the same routine kj_wta_next uses
to finish a row)."""
from __future__ import annotations

import ctypes as C
from decimal import Decimal

from . import _abi

etta = Decimal("1E-8")          # lib/stats.js:6
_rounding_mode = 4              # bignumber.js default ROUND_HALF_UP; lib/kmerFinderServer.js:7 sets 2


def set_rounding_mode(mode: int) -> None:
    """BN.config({ROUNDING_MODE: mode}) for the functions of this module."""
    global _rounding_mode
    if not 0 <= int(mode) <= 6:
        raise ValueError("rounding mode must be 0..6")
    _rounding_mode = int(mode)


def zScore(r1: int, n1: int, r2: int, n2: int) -> Decimal:
    """lib/stats.js:19-45 (exact decimal, 20 places)."""
    z = C.c_double()
    buf = C.create_string_buffer(256)
    _abi.check(_abi.lib().kj_stats_zscore(_rounding_mode, int(r1), int(n1), int(r2), int(n2),
                                          C.byref(z), buf, 256))
    return Decimal(buf.value.decode())


def fastp(z) -> Decimal:
    """lib/stats.js:52-115."""
    p = C.c_double()
    text = format(Decimal(z), "f") if not isinstance(z, str) else z
    _abi.check(_abi.lib().kj_stats_fastp_text(text.encode(), C.byref(p)))
    return Decimal(repr(p.value))
