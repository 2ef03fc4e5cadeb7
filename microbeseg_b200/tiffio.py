"""TIFF stack I/O for the CLI (`infer_script_local.py:82,165` use tifffile.imread / imwrite).

tifffile is used when it is installed.  It is absent from this image, so a minimal codec covers what the
path needs: uncompressed, strip-based, little-endian grayscale (uint8 / uint16 / float32) or chunky RGB
pages, multi-page = stack, classic TIFF and BigTIFF on read, classic TIFF on write (< 4 GiB).
"""
import struct

import numpy as np

try:  # pragma: no cover - not available in the build image
    import tifffile as _tifffile
except Exception:  # noqa: BLE001
    _tifffile = None

_TYPES = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 16: "Q"}


def imread(path):
    if _tifffile is not None:
        return _tifffile.imread(str(path))
    with open(str(path), "rb") as f:
        data = f.read()
    if data[:2] != b"II":
        raise ValueError("only little-endian TIFF files are supported by the built-in reader")
    magic = struct.unpack_from("<H", data, 2)[0]
    big = magic == 43
    if magic not in (42, 43):
        raise ValueError("not a TIFF file")
    off = struct.unpack_from("<Q", data, 8)[0] if big else struct.unpack_from("<I", data, 4)[0]
    pages = []
    while off:
        if big:
            n = struct.unpack_from("<Q", data, off)[0]
            base, esz = off + 8, 20
        else:
            n = struct.unpack_from("<H", data, off)[0]
            base, esz = off + 2, 12
        tags = {}
        for i in range(n):
            e = base + i * esz
            tag, typ = struct.unpack_from("<HH", data, e)
            cnt = struct.unpack_from("<Q", data, e + 4)[0] if big else struct.unpack_from("<I", data, e + 4)[0]
            vsz = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 16: 8}.get(typ)
            if vsz is None:
                continue
            voff = e + (12 if big else 8)
            inline = 8 if big else 4
            if vsz * cnt > inline:
                voff = struct.unpack_from("<Q" if big else "<I", data, voff)[0]
            if typ in (1, 3, 4, 16):
                tags[tag] = list(struct.unpack_from("<%d%s" % (cnt, _TYPES[typ]), data, voff))
        nxt = base + n * esz
        off = struct.unpack_from("<Q", data, nxt)[0] if big else struct.unpack_from("<I", data, nxt)[0]
        if tags.get(259, [1])[0] != 1:
            raise ValueError("compressed TIFF pages are not supported by the built-in reader (install tifffile)")
        if 322 in tags:
            raise ValueError("tiled TIFF pages are not supported by the built-in reader (install tifffile)")
        w, h = tags[256][0], tags[257][0]
        spp = tags.get(277, [1])[0]
        bits = tags.get(258, [8])[0]
        fmt = tags.get(339, [1])[0]
        dt = {(8, 1): np.uint8, (16, 1): np.uint16, (32, 3): np.float32, (32, 1): np.uint32, (16, 2): np.int16}.get((bits, fmt))
        if dt is None:
            raise ValueError(f"unsupported TIFF sample type: {bits} bits, format {fmt}")
        buf = b"".join(data[o:o + c] for o, c in zip(tags[273], tags[279]))
        arr = np.frombuffer(buf, dtype=np.dtype(dt).newbyteorder("<"), count=h * w * spp)
        pages.append(arr.reshape((h, w, spp) if spp > 1 else (h, w)).astype(dt))
    out = np.stack(pages) if len(pages) > 1 else pages[0]
    return out


def imwrite(path, array):
    if _tifffile is not None:
        return _tifffile.imwrite(str(path), array)
    a = np.ascontiguousarray(array)
    if a.dtype not in (np.uint8, np.uint16, np.float32):
        raise ValueError(f"unsupported dtype {a.dtype}")
    if a.ndim == 2:
        a = a[None]
    if a.ndim != 3:
        raise ValueError("expected a 2-D image or a [T,H,W] stack")
    T, H, W = a.shape
    page_bytes = H * W * a.itemsize
    ifd_bytes = 2 + 10 * 12 + 4
    if 8 + T * (page_bytes + ifd_bytes) >= (1 << 32):
        raise ValueError("stack too large for classic TIFF (install tifffile for BigTIFF output)")
    fmt = 3 if a.dtype == np.float32 else 1
    with open(str(path), "wb") as f:
        f.write(b"II" + struct.pack("<HI", 42, 8))
        pos = 8
        for t in range(T):
            data_off = pos + ifd_bytes
            nxt = data_off + page_bytes if t + 1 < T else 0
            tags = [(256, 4, W), (257, 4, H), (258, 3, a.itemsize * 8), (259, 3, 1), (262, 3, 1), (273, 4, data_off),
                    (277, 3, 1), (278, 4, H), (279, 4, page_bytes), (339, 3, fmt)]
            f.write(struct.pack("<H", len(tags)))
            for tag, typ, val in tags:
                f.write(struct.pack("<HHI", tag, typ, 1) + (struct.pack("<HH", val, 0) if typ == 3 else struct.pack("<I", val)))
            f.write(struct.pack("<I", nxt))
            f.write(a[t].astype(a.dtype.newbyteorder("<"), copy=False).tobytes())
            pos = data_off + page_bytes
