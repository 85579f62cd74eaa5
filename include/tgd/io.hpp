/*
 * Minimal stand-in for <tgd/io.hpp> of libtgd (marlam/tgd): TGD::save / TGD::load /
 * TGD::Error / TGD::strerror as called by the WurblPT example apps
 * (e.g. reference wurblpt-cornellbox/wurblpt-cornellbox.cpp:272-275) and by
 * createTextureImage(filename) (reference libwurblpt/texture_image.hpp:343-366).
 *
 * No image codec libraries exist in this environment, so files are written as
 *   *.pfm            portable float map (1 or 3 float32 components)
 *   *.ppm / *.pgm / *.pnm   binary netpbm (3 or 1 uint8 components)
 *   *.png            PNG, 8 bit, 1 to 4 components (png.hpp: valid but uncompressed)
 *   *.hdr / *.pic    Radiance RGBE picture (3 float32 components; hdr.hpp)
 *   anything else    "WPTARRAY" container: one text header line
 *                    "WPTARRAY <type> <components> <ndims> <dim0> <dim1> ...\n"
 *                    followed by the raw little-endian data (readable with numpy).
 * load() understands these five; PNG and Radiance files may come from anywhere (PNG: non-interlaced, 8 or 16 bit;
 * Radiance: RGBE in the usual orientations, run-length coded or flat). A header can never make load() allocate more
 * than the file could fill.
 */
#pragma once

#include <cstdio>
#include <string>

#include "array.hpp"
#include "png.hpp"
#include "hdr.hpp"

namespace TGD {

enum Error {
    ErrorNone = 0,
    ErrorFormatUnsupported = 1,
    ErrorFeaturesUnsupported = 2,
    ErrorMissingHints = 3,
    ErrorInvalidData = 4,
    ErrorSeekingNotSupported = 5,
    ErrorAppendingNotSupported = 6,
    ErrorLibrary = 7,
    ErrorSysErrno = 8
};

inline const char* strerror(Error e)
{
    switch (e) {
    case ErrorNone: return "success";
    case ErrorFormatUnsupported: return "unsupported file format";
    case ErrorFeaturesUnsupported: return "unsupported file features";
    case ErrorMissingHints: return "required hints are missing";
    case ErrorInvalidData: return "invalid data";
    case ErrorSeekingNotSupported: return "seeking not supported";
    case ErrorAppendingNotSupported: return "appending not supported";
    case ErrorLibrary: return "a library function failed";
    case ErrorSysErrno: return "system error";
    }
    return "unknown error";
}

namespace detail {

inline bool endsWith(const std::string& s, const char* suffix)
{
    std::string x(suffix);
    return s.size() >= x.size() && s.compare(s.size() - x.size(), x.size(), x) == 0;
}

inline const char* typeName(Type t)
{
    static const char* names[] = { "int8", "uint8", "int16", "uint16", "int32", "uint32",
        "int64", "uint64", "float32", "float64" };
    return names[int(t)];
}

/* true if `count` elements of `size` bytes fit into `available` bytes (no overflow) */
inline bool fits(size_t available, size_t count, size_t size)
{
    return size != 0 && count <= available / size;
}

inline size_t product(const std::vector<size_t>& dims, size_t components, bool& overflow)
{
    size_t n = components;
    overflow = false;
    for (size_t d : dims) {
        if (d != 0 && n > size_t(-1) / d)
            overflow = true;
        n *= d;
    }
    return n;
}

inline bool typeFromName(const std::string& n, Type& t)
{
    for (int i = 0; i <= int(float64); i++) {
        if (n == typeName(Type(i))) {
            t = Type(i);
            return true;
        }
    }
    return false;
}

}

inline Error save(const ArrayContainer& array, const std::string& fileName, bool /* append */ = false,
        const TagList& /* hints */ = TagList())
{
    FILE* f = std::fopen(fileName.c_str(), "wb");
    if (!f)
        return ErrorSysErrno;
    bool ok = true;
    const bool twoD = (array.dimensionCount() == 2);
    if (detail::endsWith(fileName, ".pfm") && twoD && array.componentType() == float32
            && (array.componentCount() == 1 || array.componentCount() == 3)) {
        // PFM stores the bottom row first, which is also WurblPT's row order.
        std::fprintf(f, "%s\n%zu %zu\n-1.0\n", array.componentCount() == 3 ? "PF" : "Pf",
                array.dimension(0), array.dimension(1));
        ok = std::fwrite(array.data(), 1, array.dataSize(), f) == array.dataSize();
    } else if ((detail::endsWith(fileName, ".ppm") || detail::endsWith(fileName, ".pgm") || detail::endsWith(fileName, ".pnm")) && twoD
            && array.componentType() == uint8
            && (array.componentCount() == 1 || array.componentCount() == 3)) {
        // netpbm stores the top row first: flip.
        std::fprintf(f, "%s\n%zu %zu\n255\n", array.componentCount() == 3 ? "P6" : "P5",
                array.dimension(0), array.dimension(1));
        size_t rowSize = array.dimension(0) * array.elementSize();
        const unsigned char* base = static_cast<const unsigned char*>(array.data());
        for (size_t y = array.dimension(1); y > 0 && ok; y--)
            ok = std::fwrite(base + (y - 1) * rowSize, 1, rowSize, f) == rowSize;
    } else if (detail::endsWith(fileName, ".png") && twoD && array.componentType() == uint8
            && array.componentCount() >= 1 && array.componentCount() <= 4) {
        ok = png::write(f, array);
    } else if ((detail::endsWith(fileName, ".hdr") || detail::endsWith(fileName, ".pic")) && twoD
            && array.componentType() == float32 && array.componentCount() == 3) {
        ok = hdr::write(f, array);
    } else {
        std::fprintf(f, "WPTARRAY %s %zu %zu", detail::typeName(array.componentType()),
                array.componentCount(), array.dimensionCount());
        for (size_t d = 0; d < array.dimensionCount(); d++)
            std::fprintf(f, " %zu", array.dimension(d));
        std::fprintf(f, "\n");
        ok = std::fwrite(array.data(), 1, array.dataSize(), f) == array.dataSize();
    }
    ok = (std::fclose(f) == 0) && ok;
    return ok ? ErrorNone : ErrorSysErrno;
}

inline ArrayContainer load(const std::string& fileName, const TagList& /* hints */ = TagList(), Error* error = nullptr)
{
    Error e = ErrorNone;
    ArrayContainer result;
    FILE* f = std::fopen(fileName.c_str(), "rb");
    if (!f) {
        e = ErrorSysErrno;
    } else {
        char magic[16] = { 0 };
        unsigned char first = 0;
        size_t fileSize = 0;
        if (std::fseek(f, 0, SEEK_END) == 0) {
            long end = std::ftell(f);
            fileSize = (end > 0 ? size_t(end) : 0);
        }
        std::rewind(f);
        if (std::fread(&first, 1, 1, f) == 1 && first == 0x89) {
            /* PNG: read the whole file */
            std::vector<unsigned char> bytes(1, first);
            unsigned char buffer[65536];
            size_t n;
            while ((n = std::fread(buffer, 1, sizeof(buffer), f)) > 0)
                bytes.insert(bytes.end(), buffer, buffer + n);
            result = png::read(bytes);
            if (result.elementCount() == 0)
                e = ErrorFeaturesUnsupported;
        } else if (std::fseek(f, 0, SEEK_SET) != 0 || std::fscanf(f, "%15s", magic) != 1) {
            e = ErrorInvalidData;
        } else if (std::string(magic).rfind("#?RADIANCE", 0) == 0 || std::string(magic).rfind("#?RGBE", 0) == 0) {
            std::rewind(f);
            std::vector<unsigned char> bytes(fileSize);
            if (std::fread(bytes.data(), 1, fileSize, f) != fileSize) {
                e = ErrorSysErrno;
            } else {
                int rc = hdr::read(bytes, result);
                e = (rc == 0 ? ErrorNone : rc == 1 ? ErrorFeaturesUnsupported : ErrorInvalidData);
            }
        } else if (std::string(magic) == "WPTARRAY") {
            char tname[16] = { 0 };
            size_t comps = 0, ndims = 0;
            Type type = uint8;
            if (std::fscanf(f, "%15s %zu %zu", tname, &comps, &ndims) != 3
                    || !detail::typeFromName(tname, type) || ndims > 8) {
                e = ErrorInvalidData;
            } else {
                std::vector<size_t> dims(ndims);
                for (size_t d = 0; d < ndims && e == ErrorNone; d++)
                    if (std::fscanf(f, "%zu", &dims[d]) != 1)
                        e = ErrorInvalidData;
                bool overflow = false;
                const size_t elements = detail::product(dims, comps, overflow);
                static const size_t typeSizes[] = { 1, 1, 2, 2, 4, 4, 8, 8, 4, 8 };
                if (e == ErrorNone && (overflow || comps == 0 || !detail::fits(fileSize, elements, typeSizes[int(type)])))
                    e = ErrorInvalidData;       /* more data announced than the file holds */
                if (e == ErrorNone && std::fgetc(f) == '\n') {
                    result = ArrayContainer(dims, comps, type);
                    if (std::fread(result.data(), 1, result.dataSize(), f) != result.dataSize())
                        e = ErrorInvalidData;
                } else {
                    e = ErrorInvalidData;
                }
            }
        } else if (std::string(magic) == "PF" || std::string(magic) == "Pf") {
            size_t w = 0, h = 0;
            float scale = 0.0f;
            if (std::fscanf(f, "%zu %zu %f", &w, &h, &scale) != 3 || scale >= 0.0f || std::fgetc(f) != '\n') {
                e = (scale >= 0.0f ? ErrorFeaturesUnsupported : ErrorInvalidData);
            } else if (w == 0 || h == 0 || !detail::fits(fileSize, w, 4 * (magic[1] == 'F' ? 3 : 1))
                    || !detail::fits(fileSize / (w * 4 * (magic[1] == 'F' ? 3 : 1)), h, 1)) {
                e = ErrorInvalidData;
            } else {
                result = ArrayContainer({ w, h }, magic[1] == 'F' ? 3 : 1, float32);
                if (std::fread(result.data(), 1, result.dataSize(), f) != result.dataSize())
                    e = ErrorInvalidData;
            }
        } else if (std::string(magic) == "P6" || std::string(magic) == "P5") {
            size_t w = 0, h = 0, maxval = 0;
            if (std::fscanf(f, "%zu %zu %zu", &w, &h, &maxval) != 3 || maxval != 255 || std::fgetc(f) != '\n') {
                e = ErrorFeaturesUnsupported;
            } else if (w == 0 || h == 0 || !detail::fits(fileSize, w, magic[1] == '6' ? 3 : 1)
                    || !detail::fits(fileSize / (w * (magic[1] == '6' ? 3 : 1)), h, 1)) {
                e = ErrorInvalidData;
            } else {
                result = ArrayContainer({ w, h }, magic[1] == '6' ? 3 : 1, uint8);
                size_t rowSize = w * result.elementSize();
                unsigned char* base = static_cast<unsigned char*>(result.data());
                for (size_t y = h; y > 0 && e == ErrorNone; y--)
                    if (std::fread(base + (y - 1) * rowSize, 1, rowSize, f) != rowSize)
                        e = ErrorInvalidData;
            }
        } else {
            e = ErrorFormatUnsupported;
        }
        std::fclose(f);
    }
    if (e != ErrorNone)
        result = ArrayContainer();
    if (error)
        *error = e;
    return result;
}

}
