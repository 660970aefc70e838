"""Rust ``{:?}`` formatting for the values the reference prints (CsrRow / GEMM Display)."""
from __future__ import annotations

import math

import numpy as np


def debug_f64(x: float) -> str:
    """``format!("{:?}", x)`` for f64: shortest round-trip digits; decimal notation with at least
    one fractional digit for 1e-4 <= |x| < 1e16 (and 0), scientific (``1e-5``, ``1.5e16``) outside."""
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "inf" if x > 0 else "-inf"
    if x == 0.0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    r = repr(x)  # shortest round-trip, like Rust's Grisu/Ryu output
    mant, _, exp = r.partition("e")
    ax = abs(x)
    if 1e-4 <= ax < 1e16:
        if exp:  # python switched to exponent form (e.g. 1e-05 never lands here; 1e+16 neither)
            r = format(x, "f").rstrip("0")
            if r.endswith("."):
                r += "0"
            return r
        return r if "." in r else r + ".0"
    # scientific: Rust prints the shortest digits, mantissa d[.ddd], exponent without sign/padding
    if not exp:
        digits = mant.replace("-", "").replace(".", "").lstrip("0")
        e = int(math.floor(math.log10(ax)))
        digits = digits.rstrip("0") or "0"
        m = digits[0] + ("." + digits[1:] if len(digits) > 1 else "")
        return ("-" if x < 0 else "") + f"{m}e{e}"
    if mant.endswith(".0"):
        mant = mant[:-2]
    return f"{mant}e{int(exp)}"


def debug_list(seq) -> str:
    parts = []
    for v in seq:
        if isinstance(v, (float, np.floating)):
            parts.append(debug_f64(v))
        else:
            parts.append(str(int(v)))
    return "[" + ", ".join(parts) + "]"
