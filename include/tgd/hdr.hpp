/*
 * hdr.hpp -- Radiance RGBE pictures (.hdr / .pic) for the libtgd stand-in: the usual container of HDR environment maps
 * (createTextureImage(filename), reference libwurblpt/texture_image.hpp:343-366, and the environment maps built on it).
 *
 * Format (Ward, "Real Pixels", Graphics Gems II; Radiance's color.c): a text header that starts with "#?RADIANCE" (or
 * "#?RGBE") and ends with an empty line, a resolution line ("-Y h +X w": rows from the top, or "+Y h +X w": rows from
 * the bottom), then h scanlines of 4-byte pixels (red, green, blue mantissas and a shared exponent),
 * value = (mantissa + 0.5) * 2^(exponent - 136), all zero = black. A scanline of 8..32767 pixels may be stored by
 * component (four runs-length coded byte planes, announced by the bytes 2 2 hi lo); the older form repeats the previous
 * pixel with pixels of the form 1 1 1 n. Both are read; written files use the by-component form where it applies.
 *
 * TGD arrays as WurblPT uses them hold the bottom row first. Reader and writer bound their work by the size of the
 * input: a header cannot make them allocate more than the scanline data that follows could fill.
 */
#pragma once

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "array.hpp"

namespace TGD {
namespace hdr {

inline void decodePixel(const unsigned char* rgbe, float* out)
{
    if (rgbe[3] == 0) {
        out[0] = out[1] = out[2] = 0.0f;
    } else {
        const float scale = std::ldexp(1.0f, int(rgbe[3]) - 136);
        out[0] = (float(rgbe[0]) + 0.5f) * scale;
        out[1] = (float(rgbe[1]) + 0.5f) * scale;
        out[2] = (float(rgbe[2]) + 0.5f) * scale;
    }
}

inline void encodePixel(const float* in, unsigned char* rgbe)
{
    float largest = in[0] > in[1] ? in[0] : in[1];
    if (in[2] > largest)
        largest = in[2];
    if (!(largest > 1e-32f) || !std::isfinite(largest)) {
        rgbe[0] = rgbe[1] = rgbe[2] = rgbe[3] = 0;      /* black (also what negative / non-finite input becomes) */
        return;
    }
    int exponent;
    const float mantissa = std::frexp(largest, &exponent);      /* largest = mantissa * 2^exponent, mantissa in [0.5, 1) */
    const float scale = mantissa * 256.0f / largest;
    for (int c = 0; c < 3; c++) {
        const float v = in[c] > 0.0f ? in[c] * scale : 0.0f;
        rgbe[c] = (unsigned char)(v >= 255.0f ? 255 : int(v));
    }
    rgbe[3] = (unsigned char)(exponent + 128);
}

/* One scanline of `width` pixels from data[pos...]; false if the data ends early or is inconsistent. */
inline bool readScanline(const std::vector<unsigned char>& data, size_t& pos, size_t width, std::vector<unsigned char>& line)
{
    line.resize(width * 4);
    const size_t n = data.size();
    if (width >= 8 && width <= 32767 && pos + 4 <= n && data[pos] == 2 && data[pos + 1] == 2 && (data[pos + 2] & 0x80) == 0) {
        if (size_t(data[pos + 2]) * 256 + data[pos + 3] != width)
            return false;
        pos += 4;
        for (int c = 0; c < 4; c++) {
            size_t x = 0;
            while (x < width) {
                if (pos >= n)
                    return false;
                unsigned int count = data[pos++];
                if (count > 128) {                      /* a run */
                    count -= 128;
                    if (x + count > width || pos >= n)
                        return false;
                    const unsigned char v = data[pos++];
                    for (unsigned int k = 0; k < count; k++)
                        line[(x++) * 4 + c] = v;
                } else {                                /* literal bytes */
                    if (count == 0 || x + count > width || pos + count > n)
                        return false;
                    for (unsigned int k = 0; k < count; k++)
                        line[(x++) * 4 + c] = data[pos++];
                }
            }
        }
        return true;
    }
    /* flat pixels, with the old repeat markers */
    size_t x = 0;
    int shift = 0;
    while (x < width) {
        if (pos + 4 > n)
            return false;
        const unsigned char* p = data.data() + pos;
        pos += 4;
        if (p[0] == 1 && p[1] == 1 && p[2] == 1) {
            if (x == 0 || shift > 24)
                return false;
            size_t count = size_t(p[3]) << shift;
            if (x + count > width)
                return false;
            for (size_t k = 0; k < count; k++, x++)
                std::memcpy(&line[x * 4], &line[(x - 1) * 4], 4);
            shift += 8;
        } else {
            std::memcpy(&line[x * 4], p, 4);
            x++;
            shift = 0;
        }
    }
    return true;
}

/* 0 = fine, 1 = not an RGBE picture this reader understands (XYZE, rotated orientations), 2 = damaged */
inline int read(const std::vector<unsigned char>& data, ArrayContainer& result)
{
    size_t pos = 0;
    auto nextLine = [&](std::string& s) {
        s.clear();
        while (pos < data.size() && data[pos] != '\n') {
            if (s.size() > 4096)
                return false;
            s.push_back(char(data[pos++]));
        }
        if (pos >= data.size())
            return false;
        pos++;
        if (!s.empty() && s.back() == '\r')
            s.pop_back();
        return true;
    };
    std::string line;
    if (!nextLine(line) || (line.rfind("#?RADIANCE", 0) != 0 && line.rfind("#?RGBE", 0) != 0))
        return 1;
    bool rgbe = true;
    for (;;) {
        if (!nextLine(line))
            return 2;
        if (line.empty())
            break;
        if (line.rfind("FORMAT=", 0) == 0)
            rgbe = (line == "FORMAT=32-bit_rle_rgbe");
    }
    if (!rgbe)
        return 1;
    if (!nextLine(line))
        return 2;
    char ySign = 0, xSign = 0;
    unsigned long h = 0, w = 0;
    if (std::sscanf(line.c_str(), "%cY %lu %cX %lu", &ySign, &h, &xSign, &w) != 4)
        return 1;       /* "+X w +Y h" and friends: rotated pictures */
    if ((ySign != '-' && ySign != '+') || xSign != '+')
        return 1;
    /* every scanline takes at least 4 bytes (by component: 4 + 4 * 2), so the rest of the file bounds the height; a
     * scanline of w pixels takes at least w / 127 bytes per component */
    const size_t rest = data.size() - pos;
    if (w == 0 || h == 0 || h > rest / 4 || w > (1ul << 26) || w / 127 > rest)
        return 2;
    if (double(w) * double(h) > 64.0 * double(rest) + 1024.0)       /* best case of the run-length code: 128 pixels in 2 bytes per component */
        return 2;
    result = ArrayContainer({ size_t(w), size_t(h) }, 3, float32);
    float* pixels = static_cast<float*>(result.data());
    std::vector<unsigned char> scan;
    for (unsigned long row = 0; row < h; row++) {
        if (!readScanline(data, pos, w, scan))
            return 2;
        const unsigned long y = (ySign == '-' ? h - 1 - row : row);    /* array row 0 = bottom */
        float* out = pixels + size_t(y) * w * 3;
        for (unsigned long x = 0; x < w; x++)
            decodePixel(&scan[x * 4], out + x * 3);
    }
    return 0;
}

inline bool write(FILE* f, const ArrayContainer& array)
{
    const size_t w = array.dimension(0), h = array.dimension(1);
    const float* pixels = static_cast<const float*>(array.data());
    if (std::fprintf(f, "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\nSOFTWARE=wurblpt-b200 tgd stand-in\n\n-Y %zu +X %zu\n", h, w) < 0)
        return false;
    std::vector<unsigned char> scan(w * 4), out;
    for (size_t row = 0; row < h; row++) {
        const float* in = pixels + (h - 1 - row) * w * 3;      /* top row first */
        for (size_t x = 0; x < w; x++)
            encodePixel(in + x * 3, &scan[x * 4]);
        out.clear();
        if (w >= 8 && w <= 32767) {
            out.push_back(2); out.push_back(2); out.push_back((unsigned char)(w >> 8)); out.push_back((unsigned char)(w & 255));
            for (int c = 0; c < 4; c++) {
                size_t x = 0;
                while (x < w) {
                    /* length of the run that starts here */
                    size_t run = 1;
                    while (x + run < w && run < 127 && scan[(x + run) * 4 + c] == scan[x * 4 + c])
                        run++;
                    if (run >= 3) {
                        out.push_back((unsigned char)(128 + run));
                        out.push_back(scan[x * 4 + c]);
                        x += run;
                    } else {
                        /* literal bytes up to the next run of three or more */
                        size_t end = x;
                        while (end < w && end - x < 128) {
                            size_t r = 1;
                            while (end + r < w && r < 3 && scan[(end + r) * 4 + c] == scan[end * 4 + c])
                                r++;
                            if (r >= 3)
                                break;
                            end++;
                        }
                        if (end == x)
                            end = x + 1;
                        out.push_back((unsigned char)(end - x));
                        for (size_t k = x; k < end; k++)
                            out.push_back(scan[k * 4 + c]);
                        x = end;
                    }
                }
            }
        } else {
            /* flat; a pixel that would read as a repeat marker (1 1 1 n) gets its blue mantissa nudged */
            out = scan;
            for (size_t x = 0; x < w; x++)
                if (out[x * 4] == 1 && out[x * 4 + 1] == 1 && out[x * 4 + 2] == 1)
                    out[x * 4 + 2] = 2;
        }
        if (std::fwrite(out.data(), 1, out.size(), f) != out.size())
            return false;
    }
    return true;
}

}
}
