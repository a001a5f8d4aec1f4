import base64
import json
import os
import zlib

_G = None


def load():
    global _G
    if _G is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.json")) as f:
            _G = json.load(f)
    return _G


def blob(key):
    if key is None:
        return None
    return zlib.decompress(base64.b64decode(load()["blobs"][key]))


def cases(op=None):
    return [c for c in load()["cases"] if op is None or c["op"] == op]


_GN = None


def load_next():
    """tests/golden/golden_next.json: the SURVEY 8(f) operators (tests/golden/make_golden_next.py)."""
    global _GN
    if _GN is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_next.json")) as f:
            _GN = json.load(f)
    return _GN


def next_blob(key):
    if key is None:
        return None
    return zlib.decompress(base64.b64decode(load_next()["blobs"][key]))


def next_cases():
    return load_next()["cases"]
