#!/usr/bin/env python3
"""tests/golden/make_golden.py — regenerate the committed golden fixtures.

Run in the build container, where /root/reference is mounted and oracle/_ref has been built
(``python oracle/build.py``).  The GPU box has neither, so everything the tests need at run time is
committed here as small files:

  lz4_input.txt / lz4_compressed.bin   the reference's OWN golden vector, copied verbatim from
                                       Output-Input/input/input.txt and Output-Input/out/compressed.bin
  Metamorphosis.txt                    the text corpus the reference's workload generator samples
                                       (Output-Input/input/Metamorphosis.txt; public-domain text, data not code)
  og_crop.png + og_crop_{lum,cb,cr}.png  a 256x100 crop (x=768, y=200) of Assets/Images/og.png and the
                                       same crop of the reference's committed colour-plane renderings
                                       Output-Input/Images/{luminance,bChrominance,rChrominance}.png
  lz4_ref_vectors.json                 sha256/size/offsets of streams produced by the REFERENCE build
                                       (oracle/_ref/libref_lz4.so, bounded variant) on seeded inputs
  jpeg_ref_vectors.npz                 coefficients / bit streams produced by the REFERENCE build
                                       (oracle/_ref/libref_jpeg.so) on og_crop, seeded noise, the
                                       SURVEY.md Appendix C block and a few degenerate blocks
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.pyoracle import Oracle, Ref  # noqa: E402
import cases  # noqa: E402  (tests/cases.py: the seeded inputs shared with the tests)

REF = "/root/reference"


def lz4_vectors_add_missing() -> None:
    """Adds the reference-build hashes of cases that tests/cases.py gained since the file was written (the existing
    entries stay as they are)."""
    path = f"{HERE}/lz4_ref_vectors.json"
    vec = json.load(open(path))
    ref = Ref("lz4")
    for name, data, block_len in cases.lz4_cases():
        if name in vec:
            continue
        stream, offs, _ = ref.lz4_compress(data, block_len)
        vec[name] = {"n": int(data.size), "block_len": block_len, "size": int(stream.size),
                     "sha256": hashlib.sha256(stream.tobytes()).hexdigest(),
                     "offsets_sha256": hashlib.sha256(offs.tobytes()).hexdigest()}
        print("added", name, vec[name]["size"])
    with open(path, "w") as f:
        json.dump(vec, f, indent=1, sort_keys=True)


def og_full() -> None:
    """BASELINE.json configs[1] on the image the reference really ships (input.jpg does not exist): the whole
    Assets/Images/og.png (1200 x 630 RGBA; 630 is not a multiple of 8) through the REFERENCE build.  The image is copied as a
    fixture; of the results only hashes and sizes are committed."""
    shutil.copy(f"{REF}/Assets/Images/og.png", f"{HERE}/og.png")
    rgba = np.asarray(Image.open(f"{HERE}/og.png").convert("RGBA"))
    rj = Ref("jpeg")
    r = rj.jpeg_encode(rgba)
    h, w, _ = rgba.shape
    rec = rj.jpeg_decode(r["coefs"], w, h, rgba)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    out = {"width": w, "height": h, "groups": int(r["coefs"].shape[0]), "stream_bytes": int(r["stream"].size),
           "coefs_sha256": sha(r["coefs"]), "stream_sha256": sha(r["stream"]), "bits_sha256": sha(r["bits"]),
           "offsets_sha256": sha(r["offsets"]), "reconstructed_sha256": sha(rec), "max_code_len": int(r["max_code_len"])}
    rf = Ref("jfif")
    for name, sub in (("444", 0), ("420", -1)):
        f = rf.jfif_encode(rgba, 75, sub)
        out[f"jfif_q75_{name}_bytes"] = int(f.size)
        out[f"jfif_q75_{name}_sha256"] = sha(f)
    with open(f"{HERE}/og_full_ref.json", "w") as fo:
        json.dump(out, fo, indent=1, sort_keys=True)
    print(out)


def main() -> None:
    shutil.copy(f"{REF}/Output-Input/input/input.txt", f"{HERE}/lz4_input.txt")
    shutil.copy(f"{REF}/Output-Input/out/compressed.bin", f"{HERE}/lz4_compressed.bin")
    shutil.copy(f"{REF}/Output-Input/input/Metamorphosis.txt", f"{HERE}/Metamorphosis.txt")
    shutil.copy(f"{REF}/Output-Input/out/compressed.txt", f"{HERE}/lz4_compressed_hex.txt")    # hex dump of compressed.bin
    shutil.copy(f"{REF}/Output-Input/out/uncompressed.txt", f"{HERE}/lz4_uncompressed.txt")   # the reference decoder's output

    x0, y0, cw, ch = 768, 200, 256, 100
    for src, dst in (("Assets/Images/og.png", "og_crop.png"),
                     ("Output-Input/Images/luminance.png", "og_crop_lum.png"),
                     ("Output-Input/Images/bChrominance.png", "og_crop_cb.png"),
                     ("Output-Input/Images/rChrominance.png", "og_crop_cr.png")):
        im = Image.open(f"{REF}/{src}").convert("RGBA").crop((x0, y0, x0 + cw, y0 + ch))
        im.save(f"{HERE}/{dst}", optimize=True)

    # ---- LZ4: reference (bounded) streams on the shared seeded cases -------------------------
    ref = Ref("lz4")
    vec = {}
    for name, data, block_len in cases.lz4_cases():
        stream, offs, _ = ref.lz4_compress(data, block_len)
        vec[name] = {"n": int(data.size), "block_len": block_len, "size": int(stream.size),
                     "sha256": hashlib.sha256(stream.tobytes()).hexdigest(),
                     "offsets_sha256": hashlib.sha256(offs.tobytes()).hexdigest()}
    with open(f"{HERE}/lz4_ref_vectors.json", "w") as f:
        json.dump(vec, f, indent=1, sort_keys=True)

    # ---- JPEG: reference coefficients + streams ------------------------------------------------
    rj = Ref("jpeg")
    out = {}
    for name, rgba in cases.jpeg_cases():
        r = rj.jpeg_encode(rgba)
        assert r["max_code_len"] <= 31, (name, r["max_code_len"])
        out[f"{name}__coefs"] = r["coefs"]
        out[f"{name}__stream"] = r["stream"]
        out[f"{name}__bits"] = r["bits"]
        out[f"{name}__offsets"] = r["offsets"]
        if name == "og_crop":  # decode half of the reference (Inverse_quantize, IDCT, assemble_image) on a real image crop
            h, w, _ = rgba.shape
            out[f"{name}__reconstructed"] = rj.jpeg_decode(r["coefs"], w, h, rgba)
    np.savez_compressed(f"{HERE}/jpeg_ref_vectors.npz", **out)
    print("golden fixtures written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    if "--lz4-add" in sys.argv:
        lz4_vectors_add_missing()
    elif "--og-full" in sys.argv:
        og_full()
    else:
        main()
