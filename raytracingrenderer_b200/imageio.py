"""Radiance .hdr (RGBE) reader/writer for Film::save (RTBase/Imaging.h:262-271, which calls
stbi_write_hdr) and for reading the reference's committed result_*.hdr renders.

RGBE conversion follows the public Radiance format (Ward, Graphics Gems II): a pixel is
(r, g, b) mantissas sharing one exponent e: value = mantissa * 2^(e - 136).
"""
import numpy as np


def float_to_rgbe(img):
    img = np.asarray(img, np.float32)
    m = np.max(img, axis=-1)
    out = np.zeros(img.shape[:-1] + (4,), np.uint8)
    ok = m >= 1e-32
    mant, exp = np.frexp(m[ok])
    scale = (mant * 256.0 / m[ok]).astype(np.float32)
    rgb = img[ok] * scale[:, None]
    out[ok, :3] = np.clip(rgb, 0, 255).astype(np.uint8)
    out[ok, 3] = (exp + 128).astype(np.uint8)
    return out


def rgbe_to_float(rgbe):
    rgbe = np.asarray(rgbe, np.uint8)
    e = rgbe[..., 3].astype(np.int32)
    scale = np.where(e > 0, np.ldexp(np.float32(1.0), e - 136), np.float32(0.0)).astype(np.float32)
    return rgbe[..., :3].astype(np.float32) * scale[..., None]


def write_hdr(path, img):
    """img: float32 [H, W, 3] (r, g, b).  Flat (non-RLE) scanlines."""
    img = np.asarray(img, np.float32)
    h, w, _ = img.shape
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n")
        f.write(("-Y %d +X %d\n" % (h, w)).encode())
        f.write(float_to_rgbe(img).tobytes())


def read_hdr(path):
    """-> float32 [H, W, 3] (r, g, b).  Handles flat and new-style RLE scanlines."""
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    first = True
    while True:
        end = data.index(b"\n", pos)
        line = data[pos:end]
        pos = end + 1
        if first:
            if not (line.startswith(b"#?RADIANCE") or line.startswith(b"#?RGBE")):
                raise ValueError("%s: not a Radiance file" % path)
            first = False
        if line == b"":
            break
    end = data.index(b"\n", pos)
    dims = data[pos:end].split()
    pos = end + 1
    if len(dims) != 4 or dims[0] != b"-Y" or dims[2] != b"+X":
        raise ValueError("%s: unsupported orientation %r" % (path, dims))
    h, w = int(dims[1]), int(dims[3])
    buf = np.frombuffer(data, np.uint8, offset=pos)
    out = np.zeros((h, w, 4), np.uint8)
    if w < 8 or w > 32767 or not (buf[0] == 2 and buf[1] == 2 and (buf[2] & 0x80) == 0):
        out[:] = buf[:h * w * 4].reshape(h, w, 4)
        return rgbe_to_float(out)
    p = 0
    b = buf
    for y in range(h):
        if not (b[p] == 2 and b[p + 1] == 2 and ((int(b[p + 2]) << 8) | int(b[p + 3])) == w):
            raise ValueError("%s: bad RLE scanline header at row %d" % (path, y))
        p += 4
        for c in range(4):
            x = 0
            row = out[y, :, c]
            while x < w:
                n = int(b[p])
                p += 1
                if n > 128:
                    n -= 128
                    row[x:x + n] = b[p]
                    p += 1
                else:
                    row[x:x + n] = b[p:p + n]
                    p += n
                x += n
    return rgbe_to_float(out)
