/*
 * PNG reading and writing for the libtgd stand-in (include/tgd/io.hpp), without any library.
 *
 * Writer: 8-bit grey / grey+alpha / rgb / rgba, filter type 0 on every row, the image data in stored (not
 * compressed) deflate blocks. Valid PNG that every viewer opens; files are as large as the raw data.
 * Reader: non-interlaced PNG with 8 or 16 bits per channel, colour types 0 (grey), 2 (rgb), 4 (grey+alpha),
 * 6 (rgba) and 3 (palette, expanded to rgb), all five row filters, and a complete inflate (stored, fixed and
 * dynamic Huffman blocks, RFC 1951). 16-bit files come back as uint16.
 * Row order: PNG stores the top row first, TGD arrays (as WurblPT uses them) the bottom row; both directions flip.
 */
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "array.hpp"

namespace TGD {
namespace png {

inline uint32_t crc32(const unsigned char* data, size_t n, uint32_t crc = 0)
{
    static uint32_t table[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++)
                c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; i++)
        crc = table[(crc ^ data[i]) & 0xFFu] ^ (crc >> 8);
    return ~crc;
}

inline uint32_t adler32(const unsigned char* data, size_t n)
{
    uint32_t a = 1, b = 0;
    for (size_t i = 0; i < n; i++) {
        a = (a + data[i]) % 65521u;
        b = (b + a) % 65521u;
    }
    return (b << 16) | a;
}

inline void putBigEndian(std::vector<unsigned char>& out, uint32_t v)
{
    for (int shift = 24; shift >= 0; shift -= 8)
        out.push_back((unsigned char)(v >> shift));
}

inline bool writeChunk(FILE* f, const char type[4], const std::vector<unsigned char>& payload)
{
    std::vector<unsigned char> chunk;
    putBigEndian(chunk, uint32_t(payload.size()));
    chunk.insert(chunk.end(), type, type + 4);
    chunk.insert(chunk.end(), payload.begin(), payload.end());
    putBigEndian(chunk, crc32(chunk.data() + 4, chunk.size() - 4));
    return std::fwrite(chunk.data(), 1, chunk.size(), f) == chunk.size();
}

/* `array`: two-dimensional, uint8, 1 to 4 components */
inline bool write(FILE* f, const ArrayContainer& array)
{
    const size_t w = array.dimension(0), h = array.dimension(1), comps = array.componentCount();
    static const unsigned char signature[8] = { 0x89, 'P', 'N', 'G', '\r', '\n', 0x1A, '\n' };
    static const unsigned char colourTypes[5] = { 0, 0, 4, 2, 6 };
    if (std::fwrite(signature, 1, 8, f) != 8)
        return false;
    std::vector<unsigned char> header;
    putBigEndian(header, uint32_t(w));
    putBigEndian(header, uint32_t(h));
    header.push_back(8);
    header.push_back(colourTypes[comps]);
    header.push_back(0);
    header.push_back(0);
    header.push_back(0);
    if (!writeChunk(f, "IHDR", header))
        return false;
    /* filter byte 0 + the row, top row first */
    const size_t rowBytes = w * comps;
    std::vector<unsigned char> raw;
    raw.reserve(h * (rowBytes + 1));
    const unsigned char* base = static_cast<const unsigned char*>(array.data());
    for (size_t y = h; y > 0; y--) {
        raw.push_back(0);
        raw.insert(raw.end(), base + (y - 1) * rowBytes, base + y * rowBytes);
    }
    /* zlib stream: header, stored deflate blocks of at most 65535 bytes, Adler-32 of the raw data */
    std::vector<unsigned char> z;
    z.push_back(0x78);
    z.push_back(0x01);
    size_t pos = 0;
    do {
        const size_t n = (raw.size() - pos > 65535 ? 65535 : raw.size() - pos);
        const bool last = (pos + n == raw.size());
        z.push_back(last ? 1 : 0);
        z.push_back((unsigned char)(n & 0xFF));
        z.push_back((unsigned char)(n >> 8));
        z.push_back((unsigned char)(~n & 0xFF));
        z.push_back((unsigned char)((~n >> 8) & 0xFF));
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
        pos += n;
    } while (pos < raw.size());
    putBigEndian(z, adler32(raw.data(), raw.size()));
    return writeChunk(f, "IDAT", z) && writeChunk(f, "IEND", std::vector<unsigned char>());
}

/* ---- inflate (RFC 1951) ---- */

class BitReader
{
public:
    const unsigned char* data;
    size_t size, pos;
    uint32_t hold;
    int count;
    bool overrun;

    BitReader(const unsigned char* d, size_t n) : data(d), size(n), pos(0), hold(0), count(0), overrun(false) {}

    uint32_t bits(int n)
    {
        while (count < n) {
            if (pos >= size) {
                overrun = true;
                return 0;
            }
            hold |= uint32_t(data[pos++]) << count;
            count += 8;
        }
        const uint32_t v = hold & ((n == 32) ? 0xFFFFFFFFu : ((1u << n) - 1u));
        hold >>= n;
        count -= n;
        return v;
    }

    void alignToByte()
    {
        hold = 0;
        count = 0;
    }
};

/* canonical Huffman code: counts per length and the symbols in code order */
class Huffman
{
public:
    uint16_t counts[16];
    std::vector<uint16_t> symbols;

    bool build(const unsigned char* lengths, int n)
    {
        std::memset(counts, 0, sizeof(counts));
        for (int i = 0; i < n; i++)
            counts[lengths[i]]++;
        counts[0] = 0;
        int left = 1;
        for (int len = 1; len < 16; len++) {
            left = left * 2 - counts[len];
            if (left < 0)
                return false;       /* over-subscribed */
        }
        uint16_t offsets[16];
        offsets[1] = 0;
        for (int len = 1; len < 15; len++)
            offsets[len + 1] = offsets[len] + counts[len];
        symbols.assign(n, 0);
        for (int i = 0; i < n; i++)
            if (lengths[i] != 0)
                symbols[offsets[lengths[i]]++] = uint16_t(i);
        return true;
    }

    int decode(BitReader& in) const
    {
        int code = 0, first = 0, index = 0;
        for (int len = 1; len < 16; len++) {
            code |= int(in.bits(1));
            if (in.overrun)
                return -1;
            const int count = counts[len];
            if (code - count < first)
                return symbols[index + (code - first)];
            index += count;
            first += count;
            first <<= 1;
            code <<= 1;
        }
        return -1;
    }
};

inline bool inflateBlockData(BitReader& in, std::vector<unsigned char>& out, const Huffman& literals, const Huffman& distances, size_t limit)
{
    static const uint16_t lengthBase[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
    static const uint16_t lengthExtra[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
    static const uint16_t distBase[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
    static const uint16_t distExtra[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
    for (;;) {
        const int symbol = literals.decode(in);
        if (symbol < 0)
            return false;
        if (out.size() >= limit && symbol != 256)
            return false;       /* the image is full: anything but end-of-block is too much (decompression bomb) */
        if (symbol < 256) {
            out.push_back((unsigned char)symbol);
        } else if (symbol == 256) {
            return true;
        } else {
            const int li = symbol - 257;
            if (li >= 29)
                return false;
            const int length = lengthBase[li] + int(in.bits(lengthExtra[li]));
            const int di = distances.decode(in);
            if (di < 0 || di >= 30)
                return false;
            const size_t distance = distBase[di] + in.bits(distExtra[di]);
            if (in.overrun || distance > out.size() || out.size() + size_t(length) > limit)
                return false;
            for (int k = 0; k < length; k++)
                out.push_back(out[out.size() - distance]);
        }
    }
}

/* raw deflate stream -> bytes */
inline bool inflate(const unsigned char* data, size_t size, std::vector<unsigned char>& out, size_t limit = ~size_t(0))
{
    BitReader in(data, size);
    bool last = false;
    while (!last) {
        last = in.bits(1) != 0;
        const uint32_t type = in.bits(2);
        if (in.overrun)
            return false;
        if (type == 0) {
            in.alignToByte();
            if (in.pos + 4 > in.size)
                return false;
            const size_t n = in.data[in.pos] | (size_t(in.data[in.pos + 1]) << 8);
            const size_t inverted = in.data[in.pos + 2] | (size_t(in.data[in.pos + 3]) << 8);
            in.pos += 4;
            if ((n ^ inverted) != 0xFFFFu || in.pos + n > in.size)
                return false;
            if (out.size() + n > limit)
                return false;       /* more data than the image can hold: refuse (decompression bomb) */
            out.insert(out.end(), in.data + in.pos, in.data + in.pos + n);
            in.pos += n;
        } else if (type == 1) {
            unsigned char lengths[320];
            for (int i = 0; i < 144; i++) lengths[i] = 8;
            for (int i = 144; i < 256; i++) lengths[i] = 9;
            for (int i = 256; i < 280; i++) lengths[i] = 7;
            for (int i = 280; i < 288; i++) lengths[i] = 8;
            Huffman literals, distances;
            literals.build(lengths, 288);
            for (int i = 0; i < 30; i++) lengths[i] = 5;
            distances.build(lengths, 30);
            if (!inflateBlockData(in, out, literals, distances, limit))
                return false;
        } else if (type == 2) {
            static const unsigned char order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
            const int nLiterals = int(in.bits(5)) + 257;
            const int nDistances = int(in.bits(5)) + 1;
            const int nCodes = int(in.bits(4)) + 4;
            if (in.overrun || nLiterals > 286 || nDistances > 30)
                return false;
            unsigned char lengths[320];
            std::memset(lengths, 0, sizeof(lengths));
            for (int i = 0; i < nCodes; i++)
                lengths[order[i]] = (unsigned char)in.bits(3);
            Huffman codeLengths;
            if (!codeLengths.build(lengths, 19))
                return false;
            std::memset(lengths, 0, sizeof(lengths));
            int index = 0;
            while (index < nLiterals + nDistances) {
                const int symbol = codeLengths.decode(in);
                if (symbol < 0)
                    return false;
                if (symbol < 16) {
                    lengths[index++] = (unsigned char)symbol;
                } else {
                    int previous = 0, repeat = 0;
                    if (symbol == 16) {
                        if (index == 0)
                            return false;
                        previous = lengths[index - 1];
                        repeat = 3 + int(in.bits(2));
                    } else if (symbol == 17) {
                        repeat = 3 + int(in.bits(3));
                    } else {
                        repeat = 11 + int(in.bits(7));
                    }
                    if (index + repeat > nLiterals + nDistances)
                        return false;
                    while (repeat-- > 0)
                        lengths[index++] = (unsigned char)previous;
                }
            }
            if (lengths[256] == 0)
                return false;
            Huffman literals, distances;
            if (!literals.build(lengths, nLiterals) || !distances.build(lengths + nLiterals, nDistances))
                return false;
            if (!inflateBlockData(in, out, literals, distances, limit))
                return false;
        } else {
            return false;
        }
    }
    return true;
}

inline int paeth(int a, int b, int c)
{
    const int p = a + b - c;
    const int pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

/* whole file in `file`; returns an empty container when the file is not a PNG this reader handles */
inline ArrayContainer read(const std::vector<unsigned char>& file)
{
    static const unsigned char signature[8] = { 0x89, 'P', 'N', 'G', '\r', '\n', 0x1A, '\n' };
    if (file.size() < 8 || std::memcmp(file.data(), signature, 8) != 0)
        return ArrayContainer();
    auto bigEndian = [&](size_t at) {
        return (uint32_t(file[at]) << 24) | (uint32_t(file[at + 1]) << 16) | (uint32_t(file[at + 2]) << 8) | uint32_t(file[at + 3]);
    };
    size_t w = 0, h = 0;
    int depth = 0, colourType = -1, interlace = 0;
    std::vector<unsigned char> compressed, palette;
    size_t pos = 8;
    while (pos + 12 <= file.size()) {
        const size_t length = bigEndian(pos);
        const std::string type(reinterpret_cast<const char*>(file.data() + pos + 4), 4);
        const size_t payload = pos + 8;
        if (payload + length + 4 > file.size())
            return ArrayContainer();
        if (crc32(file.data() + pos + 4, length + 4) != bigEndian(payload + length))
            return ArrayContainer();        /* damaged chunk */
        if (type == "IHDR" && length >= 13) {
            w = bigEndian(payload);
            h = bigEndian(payload + 4);
            depth = file[payload + 8];
            colourType = file[payload + 9];
            interlace = file[payload + 12];
        } else if (type == "PLTE") {
            palette.assign(file.begin() + payload, file.begin() + payload + length);
        } else if (type == "IDAT") {
            compressed.insert(compressed.end(), file.begin() + payload, file.begin() + payload + length);
        } else if (type == "IEND") {
            break;
        }
        pos = payload + length + 4;
    }
    int channels = 0;
    switch (colourType) {
    case 0: channels = 1; break;
    case 2: channels = 3; break;
    case 3: channels = 1; break;
    case 4: channels = 2; break;
    case 6: channels = 4; break;
    }
    const bool depthOk = (colourType == 3 ? depth == 8 : (depth == 8 || depth == 16));
    if (w == 0 || h == 0 || channels == 0 || compressed.size() < 6)
        return ArrayContainer();
    /* Images this reader does not decode are refused loudly: the caller would otherwise get an empty array where
     * the reference's libpng-based loader returns the image. */
    if (!depthOk || interlace != 0) {
        std::fprintf(stderr, "tgd png reader: unsupported PNG variant (bit depth %d, colour type %d, %s): "
                "convert the file to 8 or 16 bit, non-interlaced\n", depth, colourType, interlace ? "interlaced" : "non-interlaced");
        return ArrayContainer();
    }
    /* sizes: bounded so that the products below cannot wrap and a crafted header cannot ask for terabytes */
    if (w > (size_t(1) << 20) || h > (size_t(1) << 20) || w * h > (size_t(1) << 32))
        return ArrayContainer();
    const size_t expected = h * (w * size_t(channels) * size_t(depth / 8) + 1);
    /* zlib header (RFC 1950): compression method 8, window <= 32 KiB, no preset dictionary, check bits */
    const unsigned cmf = compressed[0], flg = compressed[1];
    if ((cmf & 0x0F) != 8 || (cmf >> 4) > 7 || (flg & 0x20) != 0 || ((cmf << 8) | flg) % 31 != 0)
        return ArrayContainer();
    std::vector<unsigned char> raw;
    raw.reserve(expected);
    if (!inflate(compressed.data() + 2, compressed.size() - 6, raw, expected))   /* between the zlib header and the Adler-32 */
        return ArrayContainer();
    if (raw.size() != expected)
        return ArrayContainer();
    {
        uint32_t a = 1, b = 0;
        for (size_t i = 0; i < raw.size(); i++) {
            a = (a + raw[i]) % 65521u;
            b = (b + a) % 65521u;
        }
        const size_t at = compressed.size() - 4;
        const uint32_t stored = (uint32_t(compressed[at]) << 24) | (uint32_t(compressed[at + 1]) << 16)
            | (uint32_t(compressed[at + 2]) << 8) | uint32_t(compressed[at + 3]);
        if (stored != ((b << 16) | a))
            return ArrayContainer();
    }
    const size_t bpp = size_t(channels) * (depth / 8);
    const size_t rowBytes = w * bpp;
    /* undo the row filters in place */
    std::vector<unsigned char> previous(rowBytes, 0);
    for (size_t y = 0; y < h; y++) {
        unsigned char* row = raw.data() + y * (rowBytes + 1) + 1;
        const int filter = row[-1];
        for (size_t i = 0; i < rowBytes; i++) {
            const int left = (i >= bpp ? row[i - bpp] : 0);
            const int up = previous[i];
            const int upLeft = (i >= bpp ? previous[i - bpp] : 0);
            int predictor = 0;
            switch (filter) {
            case 0: predictor = 0; break;
            case 1: predictor = left; break;
            case 2: predictor = up; break;
            case 3: predictor = (left + up) / 2; break;
            case 4: predictor = paeth(left, up, upLeft); break;
            default: return ArrayContainer();
            }
            row[i] = (unsigned char)(row[i] + predictor);
        }
        std::memcpy(previous.data(), row, rowBytes);
    }
    const size_t outComps = (colourType == 3 ? 3 : size_t(channels));
    ArrayContainer result({ w, h }, outComps, depth == 16 ? uint16 : uint8);
    unsigned char* dst = static_cast<unsigned char*>(result.data());
    for (size_t y = 0; y < h; y++) {
        const unsigned char* row = raw.data() + y * (rowBytes + 1) + 1;
        unsigned char* out = dst + (h - 1 - y) * w * result.elementSize();
        if (colourType == 3) {
            for (size_t x = 0; x < w; x++)
                for (int c = 0; c < 3; c++)
                    out[x * 3 + c] = (size_t(row[x]) * 3 + c < palette.size() ? palette[size_t(row[x]) * 3 + c] : 0);
        } else if (depth == 8) {
            std::memcpy(out, row, rowBytes);
        } else {
            /* big-endian samples -> host order */
            uint16_t* out16 = reinterpret_cast<uint16_t*>(out);
            for (size_t i = 0; i < w * size_t(channels); i++)
                out16[i] = uint16_t((row[2 * i] << 8) | row[2 * i + 1]);
        }
    }
    return result;
}

}
}
