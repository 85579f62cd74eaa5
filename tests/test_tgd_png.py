"""The PNG codec of the libtgd stand-in (include/tgd/png.hpp): TGD::load() must decode PNG files made elsewhere
(here: with Python's zlib, all five row filters, 8 and 16 bit, grey / grey+alpha / rgb / rgba / palette, split IDAT
chunks, dynamic Huffman blocks), TGD::save() must write files other decoders accept, and both must flip the rows
(PNG: top row first, TGD arrays as WurblPT uses them: bottom row first)."""
import os
import struct
import subprocess
import zlib

import numpy as np

from util import read_png, read_wptarray

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROBE = r'''
#include <cstdio>
#include <cstring>
#include <tgd/array.hpp>
#include <tgd/io.hpp>
int main(int argc, char** argv) {
    TGD::Array<uint8_t> img({ 5, 3 }, 3);
    for (size_t y = 0; y < 3; y++)
        for (size_t x = 0; x < 5; x++) {
            uint8_t* p = img.get<uint8_t>(y * 5 + x);
            p[0] = 10 * x; p[1] = 50 * y; p[2] = 200;
        }
    if (TGD::save(img, "written.png") != TGD::ErrorNone) return 1;
    TGD::Error err;
    TGD::ArrayContainer back = TGD::load("written.png", TGD::TagList(), &err);
    if (err != TGD::ErrorNone || back.dimension(0) != 5 || back.dimension(1) != 3 || back.componentCount() != 3) return 2;
    if (std::memcmp(back.data(), img.data(), img.dataSize()) != 0) return 3;
    for (int i = 1; i < argc; i++) {
        TGD::ArrayContainer a = TGD::load(argv[i], TGD::TagList(), &err);
        if (err != TGD::ErrorNone) { std::printf("%s: %s\n", argv[i], TGD::strerror(err)); return 4; }
        TGD::save(a, std::string(argv[i]) + ".raw");
    }
    return 0;
}
'''


def chunk(kind, payload):
    body = kind + payload
    return struct.pack(">I", len(payload)) + body + struct.pack(">I", zlib.crc32(body) & 0xFFFFFFFF)


def write_png(path, img, depth=8, palette=None):
    """A PNG encoder for the test: filter type y % 5 on row y, zlib level 9, the stream split into two IDAT chunks."""
    h, w = img.shape[:2]
    channels = 1 if img.ndim == 2 else img.shape[2]
    colour = 3 if palette is not None else {1: 0, 2: 4, 3: 2, 4: 6}[channels]
    bpp = channels * depth // 8
    raw = bytearray()
    previous = bytearray(w * bpp)
    for y in range(h):
        row = bytearray(img[y].astype(np.uint8 if depth == 8 else ">u2").tobytes())
        kind = y % 5
        out = bytearray(len(row))
        for i in range(len(row)):
            a = row[i - bpp] if i >= bpp else 0
            b = previous[i]
            c = previous[i - bpp] if i >= bpp else 0
            if kind == 0:
                p = 0
            elif kind == 1:
                p = a
            elif kind == 2:
                p = b
            elif kind == 3:
                p = (a + b) // 2
            else:
                e = a + b - c
                pa, pb, pc = abs(e - a), abs(e - b), abs(e - c)
                p = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
            out[i] = (row[i] - p) & 255
        raw += bytes([kind]) + out
        previous = row
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, colour, 0, 0, 0))
    if palette is not None:
        data += chunk(b"PLTE", palette.astype(np.uint8).tobytes())
    stream = zlib.compress(bytes(raw), 9)
    half = len(stream) // 2
    data += chunk(b"IDAT", stream[:half]) + chunk(b"IDAT", stream[half:]) + chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(data)


def test_png_files_round_trip_through_the_stand_in(tmp_path):
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, size=(40, 61, 3)).astype(np.uint8)
    rgb[10:30, 5:50] = (12, 200, 99)
    palette = rng.integers(0, 256, size=(16, 3)).astype(np.uint8)
    cases = {
        "grey.png": ((np.add.outer(np.arange(37), np.arange(53)) * 3 % 256).astype(np.uint8), 8, None),
        "rgb.png": (rgb, 8, None),
        "rgba.png": (rng.integers(0, 256, size=(19, 23, 4)).astype(np.uint8), 8, None),
        "greyalpha.png": (rng.integers(0, 256, size=(8, 9, 2)).astype(np.uint8), 8, None),
        "deep.png": (rng.integers(0, 65536, size=(11, 13, 3)).astype(np.uint16), 16, None),
        "palette.png": (rng.integers(0, 16, size=(14, 15)).astype(np.uint8), 8, palette),
        "big.png": (np.tile(rgb, (6, 6, 1)), 8, None),      # long matches, several deflate blocks
    }
    for name, (img, depth, pal) in cases.items():
        write_png(str(tmp_path / name), img, depth, pal)
    (tmp_path / "probe.cpp").write_text(PROBE)
    r = subprocess.run(["/usr/bin/g++", "-std=c++20", "-O1", "-I" + os.path.join(ROOT, "include"), "probe.cpp", "-o", "probe"],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run(["./probe"] + list(cases), cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout)
    for name, (img, depth, pal) in cases.items():
        got = read_wptarray(str(tmp_path / (name + ".raw")))[::-1]      # row 0 of the array is the bottom row
        want = img if pal is None else pal[img]
        assert np.array_equal(got, want.reshape(want.shape[0], want.shape[1], -1)), name
    # what the stand-in wrote: CRCs, zlib stream and pixels, read with an independent decoder
    written = read_png(str(tmp_path / "written.png"))
    want = np.zeros((3, 5, 3), np.uint8)
    for y in range(3):
        for x in range(5):
            want[y, x] = (10 * x, 50 * y, 200)
    assert np.array_equal(written, want)
    data = (tmp_path / "written.png").read_bytes()
    pos = 8
    while pos < len(data):
        n = struct.unpack(">I", data[pos:pos + 4])[0]
        assert zlib.crc32(data[pos + 4:pos + 8 + n]) & 0xFFFFFFFF == struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0]
        pos += 12 + n


MALFORMED_PROBE = r'''
#include <cstdio>
#include <tgd/array.hpp>
#include <tgd/io.hpp>
int main(int argc, char** argv) {
    for (int i = 1; i < argc; i++) {
        TGD::Error err = TGD::ErrorNone;
        TGD::ArrayContainer a = TGD::load(argv[i], TGD::TagList(), &err);
        std::printf("%d %zu\n", int(err), a.elementCount());
    }
    return 0;
}
'''


def test_png_reader_refuses_malformed_files(tmp_path):
    """Crafted headers (dimensions whose products wrap), a decompression bomb, a wrong Adler-32, a damaged chunk and
    an interlaced file must come back as load errors, not crashes or out-of-bounds accesses; a good file still loads."""
    sig = b"\x89PNG\r\n\x1a\n"
    good_raw = b"".join(b"\x00" + bytes([1, 2, 3] * 4) for _ in range(2))      # 4x2 rgb, filter 0
    def ihdr(w, h, depth=8, colour=2, interlace=0):
        return chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, colour, 0, 0, interlace))
    bad_adler = bytearray(zlib.compress(good_raw))
    bad_adler[-1] ^= 0xFF
    damaged = bytearray(sig + ihdr(4, 2) + chunk(b"IDAT", zlib.compress(good_raw)) + chunk(b"IEND", b""))
    damaged[45] ^= 0x01
    files = {
        "huge.png": sig + ihdr(0x7FFFFFFF, 0x7FFFFFFF, 16, 6) + chunk(b"IDAT", zlib.compress(good_raw)) + chunk(b"IEND", b""),
        "bomb.png": sig + ihdr(4, 2) + chunk(b"IDAT", zlib.compress(b"\x00" * 5_000_000)) + chunk(b"IEND", b""),
        "adler.png": sig + ihdr(4, 2) + chunk(b"IDAT", bytes(bad_adler)) + chunk(b"IEND", b""),
        "damaged.png": bytes(damaged),
        "interlaced.png": sig + ihdr(4, 2, interlace=1) + chunk(b"IDAT", zlib.compress(good_raw)) + chunk(b"IEND", b""),
        "short.png": sig + ihdr(4, 2) + chunk(b"IDAT", zlib.compress(good_raw[:-5])) + chunk(b"IEND", b""),
        "ok.png": sig + ihdr(4, 2) + chunk(b"IDAT", zlib.compress(good_raw)) + chunk(b"IEND", b""),
    }
    for name, data in files.items():
        (tmp_path / name).write_bytes(data)
    (tmp_path / "probe.cpp").write_text(MALFORMED_PROBE)
    subprocess.run(["g++", "-std=c++20", "-O1", "-fsanitize=address", "-I" + os.path.join(ROOT, "include"),
                    str(tmp_path / "probe.cpp"), "-o", str(tmp_path / "probe")], check=True)
    r = subprocess.run([str(tmp_path / "probe")] + [str(tmp_path / n) for n in files], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [line.split() for line in r.stdout.strip().splitlines()]
    assert len(lines) == len(files)
    for (name, _), (err, count) in zip(files.items(), lines):
        if name == "ok.png":
            assert err == "0" and count == "8"
        else:
            assert err != "0" and count == "0", name
