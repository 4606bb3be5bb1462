"""TEST INFRASTRUCTURE: a minimal scan-line OpenEXR writer (NONE / RLE / ZIPS / ZIP / PIZ, HALF / FLOAT channels) used to make
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


# ---- PIZ (encoder side: range compaction, forward wavelet, canonical Huffman with run-length symbols) --------------------
def _wenc14(a, b):
    a = a - 65536 if a >= 32768 else a; b = b - 65536 if b >= 32768 else b
    return ((a + b) >> 1) & 0xffff, (a - b) & 0xffff


def _wenc16(a, b):
    ao = (a + 0x8000) & 0xffff
    m = (ao + b) >> 1; d = ao - b
    if d < 0:
        m = (m + 0x8000) & 0xffff
    return m, d & 0xffff


def _wav2_encode(buf, base, nx, ox, ny, oy, mx):
    enc = _wenc14 if mx < (1 << 14) else _wenc16
    n = min(nx, ny); p = 1; p2 = 2
    while p2 <= n:
        oy1, oy2, ox1, ox2 = oy * p, oy * p2, ox * p, ox * p2
        py = base; ey = base + oy * (ny - p2)
        while py <= ey:
            px = py; ex = py + ox * (nx - p2)
            while px <= ex:
                p01, p10 = px + ox1, px + oy1; p11 = p10 + ox1
                i00, i01 = enc(buf[px], buf[p01]); i10, i11 = enc(buf[p10], buf[p11])
                buf[px], buf[p10] = enc(i00, i10); buf[p01], buf[p11] = enc(i01, i11)
                px += ox2
            if nx & p:
                p10 = px + oy1
                buf[px], buf[p10] = enc(buf[px], buf[p10])
            py += oy2
        if ny & p:
            px = py; ex = py + ox * (nx - p2)
            while px <= ex:
                p01 = px + ox1
                buf[px], buf[p01] = enc(buf[px], buf[p01])
                px += ox2
        p = p2; p2 <<= 1


class _Bits:
    def __init__(self):
        self.out = bytearray(); self.c = 0; self.lc = 0; self.n = 0

    def put(self, nbits, v):
        self.c = (self.c << nbits) | v; self.lc += nbits; self.n += nbits
        while self.lc >= 8:
            self.out.append((self.c >> (self.lc - 8)) & 0xff); self.lc -= 8
        self.c &= (1 << self.lc) - 1

    def flush(self):
        if self.lc:
            self.out.append((self.c << (8 - self.lc)) & 0xff); self.lc = 0; self.c = 0
        return bytes(self.out)


def _huf_compress(symbols):
    import heapq
    freq = {}
    for v in symbols:
        freq[v] = freq.get(v, 0) + 1
    im = min(freq); iM = max(freq) + 1
    freq[iM] = 1                                               # the run-length pseudo symbol
    heap = [(f, i, (sym,)) for i, (sym, f) in enumerate(sorted(freq.items()))]
    heapq.heapify(heap); length = {sym: 0 for sym in freq}; uid = len(heap)
    while len(heap) > 1:
        f1, _, s1 = heapq.heappop(heap); f2, _, s2 = heapq.heappop(heap)
        for sym in s1 + s2:
            length[sym] += 1
        heapq.heappush(heap, (f1 + f2, uid, s1 + s2)); uid += 1
    assert max(length.values()) <= 58
    n = [0] * 59
    for l in length.values():
        n[l] += 1
    n[0] = 0; c = 0
    for i in range(58, 0, -1):
        nc = (c + n[i]) >> 1; n[i] = c; c = nc
    code = {}
    for sym in sorted(length):
        code[sym] = n[length[sym]]; n[length[sym]] += 1
    tb = _Bits(); i = im
    while i <= iM:                                             # code lengths, 6 bits each, zero runs packed
        l = length.get(i, 0)
        if l == 0:
            run = 1
            while i < iM and run < 261 and length.get(i + 1, 0) == 0:
                i += 1; run += 1
            if run >= 2:
                if run >= 6:
                    tb.put(6, 63); tb.put(8, run - 6)
                else:
                    tb.put(6, 59 + run - 2)
                i += 1
                continue
        tb.put(6, l); i += 1
    table = tb.flush()
    db = _Bits()

    def send(sym, run):                                        # `run` repeats after the first occurrence
        if length[sym] + length[iM] + 8 < length[sym] * run:
            db.put(length[sym], code[sym]); db.put(length[iM], code[iM]); db.put(8, run)
        else:
            for _ in range(run + 1):
                db.put(length[sym], code[sym])
    cur = symbols[0]; run = 0
    for v in symbols[1:]:
        if v == cur and run < 255:
            run += 1
        else:
            send(cur, run); run = 0
        cur = v
    send(cur, run)
    nbits = db.n; data = db.flush()
    return struct.pack("<IIIII", im, iM, len(table), nbits, 0) + table + data


def _piz_block(rows_per_channel, types):
    """rows_per_channel: list over channels of 2-D arrays [lines, w] (uint16 view for HALF, uint32 view otherwise)."""
    buf = []; layout = []
    for a, t in zip(rows_per_channel, types):
        words = 1 if t == 1 else 2
        u16 = np.ascontiguousarray(a).view(np.uint16).reshape(a.shape[0], a.shape[1] * words)
        layout.append((len(buf), a.shape[1], words, a.shape[0])); buf.extend(int(x) for x in u16.reshape(-1))
    present = np.zeros(65536, bool); present[np.array(buf, np.int64)] = True; present[0] = False
    nz = np.flatnonzero(present)
    bitmap = np.packbits(present.reshape(-1, 8)[:, ::-1], axis=1).reshape(-1)      # bit i & 7 of byte i >> 3
    if len(nz):
        lo, hi = int(nz[0]) >> 3, int(nz[-1]) >> 3; head = struct.pack("<HH", lo, hi) + bitmap[lo:hi + 1].tobytes()
    else:
        head = struct.pack("<HH", 8191, 0)
    lut = np.zeros(65536, np.int64); k = 0
    for i in range(65536):
        if i == 0 or present[i]:
            lut[i] = k; k += 1
    mx = k - 1
    buf = [int(lut[v]) for v in buf]
    for start, nx, words, ny in layout:
        for j in range(words):
            _wav2_encode(buf, start + j, nx, words, ny, nx * words, mx)
    comp = _huf_compress(buf)
    return head + struct.pack("<i", len(comp)) + comp


def write_exr(path, channels, compression="zip", line_order=0, data_window_origin=(0, 0)):
    """channels: dict name -> 2-D array (float16 -> HALF, float32 -> FLOAT, uint32 -> UINT), row 0 = top."""
    names = sorted(channels)
    h, w = channels[names[0]].shape
    comp = {"none": 0, "rle": 1, "zips": 2, "zip": 3, "piz": 4}[compression]
    lines = {3: 16, 4: 32}.get(comp, 1)
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
        elif comp == 4:
            data = _piz_block([channels[n][b0:min(h, b0 + lines)] for n in names], [ptype[channels[n].dtype] for n in names])
            if len(data) >= len(raw):
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
