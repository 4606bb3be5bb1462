"""TEST INFRASTRUCTURE: a minimal scan-line OpenEXR writer (NONE / RLE / ZIPS / ZIP, HALF / FLOAT channels) used to make
fixtures for the product's reader.  Independent fixtures come from the reference's tinyexr (oracle/_ref/exr_ref save)."""
import struct
import zlib

import numpy as np


def _attr(name, typ, payload):
    return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(payload)) + payload


def _predict(raw):
    a = np.frombuffer(raw, np.uint8)
    t = np.concatenate([a[0::2], a[1::2]])                     # even bytes, then odd bytes
    d = t.astype(np.int32)
    out = t.copy()
    out[1:] = ((d[1:] - d[:-1] + 128 + 256) % 256).astype(np.uint8)
    return out.tobytes()


def _rle(data):
    out = bytearray(); i = 0; n = len(data)
    while i < n:
        j = i
        while j + 1 < n and data[j + 1] == data[i] and j - i < 126:
            j += 1
        if j - i >= 2:                                          # run of j-i+1 equal bytes
            out += struct.pack("b", j - i) + bytes([data[i]]); i = j + 1
        else:                                                   # literal stretch until the next run of 3
            k = i
            while k < n and k - i < 127 and not (k + 2 < n and data[k] == data[k + 1] == data[k + 2]):
                k += 1
            out += struct.pack("b", -(k - i)) + data[i:k]; i = k
    return bytes(out)


def write_exr(path, channels, compression="zip", line_order=0, data_window_origin=(0, 0)):
    """channels: dict name -> 2-D array (float16 -> HALF, float32 -> FLOAT, uint32 -> UINT), row 0 = top."""
    names = sorted(channels)
    h, w = channels[names[0]].shape
    comp = {"none": 0, "rle": 1, "zips": 2, "zip": 3}[compression]
    lines = 16 if comp == 3 else 1
    ptype = {np.dtype(np.uint32): 0, np.dtype(np.float16): 1, np.dtype(np.float32): 2}
    chl = b"".join(n.encode() + b"\0" + struct.pack("<iBxxxii", ptype[channels[n].dtype], 0, 1, 1) for n in names) + b"\0"
    x0, y0 = data_window_origin
    box = struct.pack("<iiii", x0, y0, x0 + w - 1, y0 + h - 1)
    hdr = struct.pack("<II", 20000630, 2)
    hdr += _attr("channels", "chlist", chl) + _attr("compression", "compression", bytes([comp]))
    hdr += _attr("dataWindow", "box2i", box) + _attr("displayWindow", "box2i", box)
    hdr += _attr("lineOrder", "lineOrder", bytes([line_order])) + _attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    hdr += _attr("screenWindowCenter", "v2f", struct.pack("<ff", 0, 0)) + _attr("screenWindowWidth", "float", struct.pack("<f", 1.0))
    hdr += b"\0"
    chunks = []
    for b0 in range(0, h, lines):
        raw = b"".join(np.ascontiguousarray(channels[n][y]).tobytes() for y in range(b0, min(h, b0 + lines)) for n in names)
        if comp == 0:
            data = raw
        else:
            p = _predict(raw)
            data = _rle(p) if comp == 1 else zlib.compress(p)
            if len(data) >= len(raw):
                data = raw                                      # the format stores a chunk raw when compression does not help
        chunks.append((y0 + b0, data))
    order = range(len(chunks)) if line_order == 0 else range(len(chunks) - 1, -1, -1)   # file order of the chunks
    pos = len(hdr) + 8 * len(chunks)
    offs = [0] * len(chunks); body = b""
    for i in order:
        offs[i] = pos + len(body)
        body += struct.pack("<ii", chunks[i][0], len(chunks[i][1])) + chunks[i][1]
    with open(path, "wb") as f:
        f.write(hdr + b"".join(struct.pack("<Q", o) for o in offs) + body)
