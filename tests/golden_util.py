"""Load golden fixtures (written by tests/golden/make_golden.py from the compiled reference)."""
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(tag):
    with open(os.path.join(GOLDEN, f"{tag}.json")) as f:
        fx = json.load(f)
    for q in fx["queries"]:
        q["weights"] = [float.fromhex(w) for w in q["weights"]]
        q["max_possible"] = float.fromhex(q["max_possible"])
        q["max_attained"] = float.fromhex(q["max_attained"])
    return fx


def sortable_key_to_int(hexkey: str) -> int:
    """Invert Xapian::sortable_serialise for the non-negative integers the fixtures use.
    Positive x = m * 2^e is stored as 0b11 [large-exponent bit, 3-bit or 10-bit exponent] mantissa..
    (src/xapian/api/sortable-serialise.cc); decoding is only needed to compare sort order, so we map the
    byte string to an integer that preserves order instead: zero-padded big-endian bytes."""
    if hexkey == "-":
        return 0
    b = bytes.fromhex(hexkey)
    return int.from_bytes(b.ljust(10, b"\0"), "big")
