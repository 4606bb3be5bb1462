#!/usr/bin/env python
"""Fixture generator for the OpenEXR reader (needs oracle/_ref/exr_ref, i.e. the reference checkout + oracle/build_ref.sh):
  * zip_half.exr / zip_float.exr are WRITTEN by the reference's tinyexr (SaveMultiChannelEXRToFile);
  * none_*.exr / zips_*.exr come from tests/exr_writer.py;
  * every <name>.npy is what the reference's load_exr path (tinyexr, restated in oracle/exr_ref.cpp) reads back from
    <name>.exr -- the product's reader must reproduce it bit for bit.
  * piz_*.exr come from the writer's PIZ encoder (range compaction + wavelet + canonical Huffman); tinyexr decodes them;
RLE is not supported by the reference's tinyexr: rle_half.exr is pinned by the writer's input array instead."""
import os, subprocess, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, ROOT)
from tests.exr_writer import write_exr
TOOL = os.path.join(ROOT, "oracle", "_ref", "exr_ref")


def ref_load(path):
    out = path[:-4] + ".bin"
    subprocess.run([TOOL, "load", path, out], check=True)
    raw = open(out, "rb").read(); os.remove(out)
    w, h = np.frombuffer(raw[:8], np.int32)
    return np.frombuffer(raw[8:], np.float32).reshape(h, w, 3).copy()


def image(w, h, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([0.3 + 0.3 * np.sin(0.2 * x + c) * np.cos(0.13 * y) + 0.05 * rng.random((h, w)) for c in range(3)], -1)
    img[h // 5, w // 2] = (1500.0, 1200.0, 900.0)
    return np.maximum(img, 0).astype(np.float32)


if __name__ == "__main__":
    subprocess.run([TOOL, "save", os.path.join(HERE, "zip_half.exr"), "40", "37", "half", "3"], check=True)
    subprocess.run([TOOL, "save", os.path.join(HERE, "zip_float.exr"), "33", "18", "float", "4"], check=True)
    a = image(29, 35, 1)
    write_exr(os.path.join(HERE, "none_float.exr"), {"R": a[..., 0], "G": a[..., 1], "B": a[..., 2]}, "none")
    h16 = a.astype(np.float16)
    write_exr(os.path.join(HERE, "zips_half.exr"), {"R": h16[..., 0], "G": h16[..., 1], "B": h16[..., 2]}, "zips")
    write_exr(os.path.join(HERE, "zip_half_decreasing.exr"), {"R": h16[..., 0], "G": h16[..., 1], "B": h16[..., 2]}, "zip", line_order=1)
    write_exr(os.path.join(HERE, "rle_half.exr"), {"R": h16[..., 0], "G": h16[..., 1], "B": h16[..., 2]}, "rle")
    np.save(os.path.join(HERE, "rle_half.npy"), h16.astype(np.float32))
    # PIZ: quantised images (compress, 14-bit wavelet path) and a ramp of distinct codes (>= 2^14 distinct words: 16-bit path);
    # the reference's tinyexr does not implement the format's "stored raw when not smaller" rule for PIZ, so both must shrink
    y, x = np.mgrid[0:45, 0:37]
    q = np.stack([np.round((0.3 + 0.3 * np.sin(0.2 * x + c) * np.cos(0.13 * y)) * 16) / 16 for c in range(3)], -1).astype(np.float32)
    q[7, 20] = (1500, 1200, 900)
    q16 = q.astype(np.float16)
    write_exr(os.path.join(HERE, "piz_half.exr"), {"R": q16[..., 0], "G": q16[..., 1], "B": q16[..., 2]}, "piz")
    # 17 280 distinct half codes in one 32-line block: the range table exceeds 2^14 entries -> 16-bit wavelet path
    codes = (0x0400 + np.arange(32 * 180, dtype=np.uint16).reshape(32, 180))
    ramp = {"R": codes.view(np.float16), "G": (codes + 5760).view(np.float16), "B": (codes + 11520).view(np.float16)}
    write_exr(os.path.join(HERE, "piz_half_wide.exr"), ramp, "piz")
    assert os.path.getsize(os.path.join(HERE, "piz_half_wide.exr")) < 0.9 * 32 * 180 * 6
    f32 = np.stack([np.round(q[..., c] * 4) / 4 for c in range(3)], -1).astype(np.float32)
    write_exr(os.path.join(HERE, "piz_float.exr"), {"R": f32[..., 0], "G": f32[..., 1], "B": f32[..., 2]}, "piz")
    assert os.path.getsize(os.path.join(HERE, "piz_float.exr")) < 0.9 * f32.nbytes
    for n in ("zip_half", "zip_float", "none_float", "zips_half", "zip_half_decreasing", "piz_half", "piz_float", "piz_half_wide"):
        np.save(os.path.join(HERE, n + ".npy"), ref_load(os.path.join(HERE, n + ".exr")))
    print("wrote", sorted(os.listdir(HERE)))
