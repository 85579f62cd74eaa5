"""The Radiance RGBE codec of the libtgd stand-in (include/tgd/hdr.hpp) and the size checks of TGD::load():
TGD::load() must decode .hdr files made elsewhere (here: by an encoder of this test; flat pixels, the by-component
run-length form, the old repeat markers, both row orientations), TGD::save() must write files this test's decoder
accepts, both must keep WurblPT's row order (row 0 = bottom), and no header may make load() allocate more than the
file could fill."""
import os
import subprocess

import numpy as np

from util import read_wptarray

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROBE = r'''
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <tgd/array.hpp>
#include <tgd/io.hpp>
int main(int argc, char** argv) {
    /* round trip through save() and load(): 40 x 7 pixels with runs, a black pixel and a large dynamic range */
    const size_t w = 40, h = 7;
    TGD::Array<float> img({ w, h }, 3);
    for (size_t y = 0; y < h; y++)
        for (size_t x = 0; x < w; x++) {
            float* p = img.get<float>(y * w + x);
            float level = std::ldexp(1.0f, int(y) * 5 - 12);
            p[0] = level * (x < 20 ? 1.0f : 0.25f + 0.01f * x);
            p[1] = level * 0.5f;
            p[2] = (x == 3 && y == 2) ? 0.0f : level * (0.1f + 0.02f * (x % 5));
            if (x == 7 && y == 4) { p[0] = p[1] = p[2] = 0.0f; }
        }
    if (TGD::save(img, "written.hdr") != TGD::ErrorNone) return 1;
    TGD::Error err;
    TGD::ArrayContainer back = TGD::load("written.hdr", TGD::TagList(), &err);
    if (err != TGD::ErrorNone || back.dimension(0) != w || back.dimension(1) != h || back.componentCount() != 3
            || back.componentType() != TGD::float32) return 2;
    const float* a = static_cast<const float*>(img.data());
    const float* b = static_cast<const float*>(back.data());
    for (size_t i = 0; i < w * h; i++) {
        float largest = std::fmax(a[3 * i], std::fmax(a[3 * i + 1], a[3 * i + 2]));
        for (int c = 0; c < 3; c++)
            if (std::fabs(a[3 * i + c] - b[3 * i + c]) > largest / 128.0f) {
                std::printf("pixel %zu component %d: %g -> %g\n", i, c, a[3 * i + c], b[3 * i + c]);
                return 3;
            }
    }
    /* a narrow picture is written flat */
    TGD::Array<float> narrow({ 3, 2 }, 3);
    for (size_t i = 0; i < 18; i++) static_cast<float*>(narrow.data())[i] = 0.125f * float(i + 1);
    if (TGD::save(narrow, "narrow.hdr") != TGD::ErrorNone) return 4;
    for (int i = 1; i < argc; i++) {
        TGD::ArrayContainer x = TGD::load(argv[i], TGD::TagList(), &err);
        std::printf("%s: %s\n", argv[i], TGD::strerror(err));
        if (err == TGD::ErrorNone)
            TGD::save(x, std::string(argv[i]) + ".raw");
    }
    return 0;
}
'''


def decode_rgbe(px):
    px = np.asarray(px, dtype=np.uint8).reshape(-1, 4)
    scale = np.where(px[:, 3] == 0, 0.0, np.ldexp(1.0, px[:, 3].astype(np.int32) - 136))
    rgb = (px[:, :3].astype(np.float64) + 0.5) * scale[:, None]
    rgb[px[:, 3] == 0] = 0.0
    return rgb.astype(np.float32)


def header(w, h, y_sign="-", fmt="32-bit_rle_rgbe", magic="#?RADIANCE"):
    return ("%s\n# made by the test\nFORMAT=%s\nEXPOSURE=1.0\n\n%sY %d +X %d\n" % (magic, fmt, y_sign, h, w)).encode()


def rle_scanline(row):
    """The by-component form: runs of three or more equal bytes as (128 + n, value), the rest as literals."""
    w = len(row)
    out = bytearray([2, 2, w >> 8, w & 255])
    for c in range(4):
        plane = [int(p[c]) for p in row]
        x = 0
        while x < w:
            run = 1
            while x + run < w and run < 127 and plane[x + run] == plane[x]:
                run += 1
            if run >= 3:
                out += bytes([128 + run, plane[x]])
                x += run
            else:
                n = min(run, w - x)
                out += bytes([n] + plane[x:x + n])
                x += n
    return bytes(out)


def make(tmp, name, data):
    path = os.path.join(tmp, name)
    with open(path, "wb") as f:
        f.write(data)
    return path


def read_hdr(path):
    """Decoder of this test for what TGD::save() writes: returns rows from the top."""
    data = open(path, "rb").read()
    end = data.index(b"\n\n") + 2
    assert data.startswith(b"#?RADIANCE\n") and b"FORMAT=32-bit_rle_rgbe\n" in data[:end]
    nl = data.index(b"\n", end)
    parts = data[end:nl].split()
    assert parts[0] == b"-Y" and parts[2] == b"+X"
    h, w = int(parts[1]), int(parts[3])
    pos = nl + 1
    rows = []
    for _ in range(h):
        if 8 <= w <= 32767:
            assert data[pos] == 2 and data[pos + 1] == 2 and data[pos + 2] * 256 + data[pos + 3] == w
            pos += 4
            line = np.zeros((w, 4), dtype=np.uint8)
            for c in range(4):
                x = 0
                while x < w:
                    n = data[pos]
                    pos += 1
                    if n > 128:
                        line[x:x + n - 128, c] = data[pos]
                        pos += 1
                        x += n - 128
                    else:
                        assert n > 0
                        line[x:x + n, c] = list(data[pos:pos + n])
                        pos += n
                        x += n
                assert x == w
        else:
            line = np.frombuffer(data[pos:pos + 4 * w], dtype=np.uint8).reshape(w, 4)
            pos += 4 * w
        rows.append(decode_rgbe(line))
    assert pos == len(data)
    return np.array(rows)


def test_radiance_hdr_codec(tmp_path):
    tmp = str(tmp_path)
    rng = np.random.default_rng(3)
    w, h = 24, 5
    pixels = rng.integers(0, 256, (h, w, 4)).astype(np.uint8)
    pixels[:, :, 3] = rng.integers(120, 136, (h, w))
    pixels[1, 4:17] = pixels[1, 4]              # a long run in every component
    pixels[2, :, 3] = 128                       # constant exponent plane
    pixels[3, 5] = 0                            # black
    pixels[pixels[:, :, :3].sum(2) == 3] += 1   # (never the old repeat marker 1 1 1 n by accident)
    expected_top_first = decode_rgbe(pixels).reshape(h, w, 3)
    files = {
        "flat_top.hdr": header(w, h) + pixels.tobytes(),
        "flat_bottom.hdr": header(w, h, "+") + pixels.tobytes(),
        "rle.hdr": header(w, h, magic="#?RGBE") + b"".join(rle_scanline(pixels[y]) for y in range(h)),
    }
    # the old run-length form: 1 1 1 n repeats the previous pixel n times
    old = bytearray()
    old_expected = []
    for y in range(h):
        first, second = pixels[y, 0], pixels[y, 1]
        old += bytes(first) + bytes([1, 1, 1, 10]) + bytes(second) + bytes([1, 1, 1, 12])
        old_expected.append([first] * 11 + [second] * 13)
    files["old_rle.hdr"] = header(w, h) + bytes(old)
    old_expected = decode_rgbe(np.array(old_expected, dtype=np.uint8)).reshape(h, w, 3)
    # what must be refused without a crash
    bad = {
        "truncated.hdr": files["rle.hdr"][:-7],
        "xyze.hdr": header(w, h, fmt="32-bit_rle_xyze") + pixels.tobytes(),
        "rotated.hdr": header(w, h).replace(b"-Y %d +X %d" % (h, w), b"+X %d -Y %d" % (w, h)) + pixels.tobytes(),
        "huge.hdr": header(1 << 25, 1 << 25) + pixels.tobytes(),
        "huge_rows.hdr": header(w, 1 << 30) + pixels.tobytes(),
        "wrong_width.hdr": header(w, h) + rle_scanline(pixels[0][:20]) * h,
        "huge.pfm": b"PF\n3000000000 3000000000\n-1.0\n" + b"\0" * 64,
        "huge.ppm": b"P6\n4000000000 4000000000\n255\n" + b"\0" * 64,
        "huge.raw": b"WPTARRAY float32 3 2 4000000000 4000000000\n" + b"\0" * 64,
    }
    paths = [make(tmp, k, v) for k, v in list(files.items()) + list(bad.items())]
    src = os.path.join(tmp, "probe.cpp")
    with open(src, "w") as f:
        f.write(PROBE)
    exe = os.path.join(tmp, "probe")
    subprocess.run(["g++", "-std=c++20", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
    r = subprocess.run([exe] + paths, cwd=tmp, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    verdict = dict(line.rsplit(": ", 1) for line in r.stdout.strip().splitlines())
    for name in files:
        assert verdict[os.path.join(tmp, name)] == "success", (name, verdict)
    for name in bad:
        assert verdict[os.path.join(tmp, name)] != "success", (name, verdict)
        assert not os.path.exists(os.path.join(tmp, name + ".raw"))
    assert verdict[os.path.join(tmp, "xyze.hdr")] == "unsupported file features"
    assert verdict[os.path.join(tmp, "rotated.hdr")] == "unsupported file features"
    assert verdict[os.path.join(tmp, "truncated.hdr")] == "invalid data"
    # decoded values: exact (same formula), rows bottom first in the array
    for name, top_first in (("flat_top.hdr", expected_top_first), ("rle.hdr", expected_top_first),
                            ("old_rle.hdr", old_expected)):
        got = read_wptarray(os.path.join(tmp, name + ".raw"))
        assert got.shape == (h, w, 3) and got.dtype == np.float32
        assert np.array_equal(got, top_first[::-1]), name
    got = read_wptarray(os.path.join(tmp, "flat_bottom.hdr.raw"))
    assert np.array_equal(got, expected_top_first)
    # what save() wrote: by component for the wide picture, flat for the narrow one; decodable, right way up
    written = read_hdr(os.path.join(tmp, "written.hdr"))
    assert written.shape == (7, 40, 3)
    level_top, level_bottom = 2.0 ** (6 * 5 - 12), 2.0 ** (-12)
    assert abs(written[0, 0, 0] - level_top) <= level_top / 128 and abs(written[-1, 0, 0] - level_bottom) <= level_bottom / 128
    assert (written[7 - 1 - 4, 7] == 0).all()          # the black pixel of array row 4
    narrow = read_hdr(os.path.join(tmp, "narrow.hdr"))
    assert narrow.shape == (2, 3, 3)
    want = (0.125 * (np.arange(18) + 1)).reshape(2, 3, 3)[::-1]
    assert np.abs(narrow - want).max() <= want.max() / 128


def test_hdr_file_becomes_an_environment_map_texture(tmp_path):
    """createTextureImage(filename) (texture_image.hpp:343-366) on a Radiance file: a linear float texture that an
    EnvironmentMapEquiRect accepts, and Scene::updateBVH flattens it (host side only here: WURBLPT_DEVICE=none)."""
    w, h = 16, 8
    rows = []
    for y in range(h):
        row = np.zeros((w, 4), dtype=np.uint8)
        row[:, 0] = 10 + 10 * np.arange(w)
        row[:, 1] = 100
        row[:, 2] = 20 * y + 5
        row[:, 3] = 129 + (y % 3)
        rows.append(row)
    data = header(w, h) + b"".join(rle_scanline(r) for r in rows)
    make(str(tmp_path), "sky.hdr", data)
    src = tmp_path / "t.cpp"
    src.write_text('''
#include <cstdio>
#include <wurblpt/wurblpt.hpp>
using namespace WurblPT;
int main() {
    Scene scene;
    Texture* sky = createTextureImage("sky.hdr");
    if (!sky) return 1;
    scene.take(sky);
    scene.take(new EnvironmentMapEquiRect(sky));
    scene.take(new Sphere(vec3(0.0f), 1.0f, scene.take(new MaterialLambertian(vec3(0.5f)))));
    scene.updateBVH();
    if (createTextureImage("missing.hdr")) return 2;
    std::printf("ok\\n");
    return 0;
}
''')
    lib = os.path.join(ROOT, "wurblpt_b200", "lib")
    exe = str(tmp_path / "t")
    r = subprocess.run(["/usr/bin/g++", "-std=c++20", "-O1", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe,
                        "-L" + lib, "-lwurblpt_b200", "-Wl,-rpath," + lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    env = dict(os.environ, WURBLPT_DEVICE="none")
    r = subprocess.run([exe], cwd=str(tmp_path), capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
