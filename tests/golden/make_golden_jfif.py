#!/usr/bin/env python3
"""tests/golden/make_golden_jfif.py — regenerate jfif_ref_vectors.json and og_crop_q75.jpg.

Run in the build container (needs oracle/_ref/libref_jfif.so, i.e. /root/reference mounted and ``python
oracle/build.py``).  For every case of tests/cases.py:jfif_cases() the REFERENCE's vendored stb_image_write.h
(stbi_write_jpg_to_func, compiled where it lies) encodes the pixels; size and SHA-256 of the file are committed,
and one small real file (og_crop at quality 75) verbatim.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.pyoracle import Ref  # noqa: E402
import cases  # noqa: E402


def main() -> None:
    ref = Ref("jfif")
    vec = {}
    for name, px, quality, sub in cases.jfif_cases():
        jpg = ref.jfif_encode(px, quality, sub)
        vec[name] = {"shape": list(px.shape), "quality": quality, "subsample": sub, "size": int(jpg.size),
                     "sha256": hashlib.sha256(jpg.tobytes()).hexdigest()}
        if name == "og_crop_q75":
            jpg.tofile(os.path.join(HERE, "og_crop_q75.jpg"))
    with open(os.path.join(HERE, "jfif_ref_vectors.json"), "w") as f:
        json.dump(vec, f, indent=1, sort_keys=True)
    print(f"{len(vec)} cases written")


if __name__ == "__main__":
    main()
