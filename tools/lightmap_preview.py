"""Writes the lightmap of one face of a lit .bsp as a PNG (nearest-neighbour upscaled, gamma 2.2), for eyeballing a bake.
    python tools/lightmap_preview.py lit.bsp FACE out.png
No dependencies beyond numpy and zlib; reads the file through the library's container (vrad_b200.bspfile)."""
import os
import struct
import sys
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def write_png(path, rgb8):
    h, w, _ = rgb8.shape
    raw = b"".join(b"\0" + rgb8[y].tobytes() for y in range(h))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    open(path, "wb").write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def face_lightmap(bsp_path, face):
    from vrad_b200 import bspfile as B
    f = B.BspFile(bsp_path)
    L = f.lumps()
    lump, _ = f.get(B.LUMP["LIGHTING"])
    f.close()
    fc = L.faces[face]
    w, h = int(fc["lm_size"][0]) + 1, int(fc["lm_size"][1]) + 1
    ofs = int(fc["lightofs"])
    if ofs < 0:
        raise SystemExit(f"face {face} has no lightmap")
    colors = np.frombuffer(lump[ofs:ofs + 4 * w * h], B.RGBEXP32)
    return B.color_from_rgbexp32(colors).reshape(h, w, 3)


def main():
    bsp_path, face, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    rgb = face_lightmap(bsp_path, face)
    scale = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0 / max(float(np.percentile(rgb, 99)), 1e-6)
    img = np.clip(rgb * scale, 0, 1) ** (1 / 2.2)
    img8 = (img * 255 + 0.5).astype(np.uint8)
    zoom = max(1, 512 // max(img8.shape[0], img8.shape[1]))
    img8 = np.repeat(np.repeat(img8, zoom, axis=0), zoom, axis=1)[::-1]      # t grows upwards
    write_png(out, img8)
    print(f"face {face}: {rgb.shape[1]} x {rgb.shape[0]} luxels, max {rgb.max():.1f}, wrote {out} ({img8.shape[1]} x {img8.shape[0]})")


if __name__ == "__main__":
    main()
