/* import.hpp -- Wavefront OBJ/MTL import (mirror of the reference's libwurblpt/import.hpp:64-590 API:
 * toNormalMap, importTexture, the ImportBit* flags, importIntoScene, importGeometry).
 *
 * The reference parses files with the third-party tiny_obj_loader; this header carries a small reader of its
 * own (detail::ObjFile) that fills the same kind of tables - positions / normals / texture coordinates, shapes
 * with triangulated (vertex, normal, texcoord) index triples and one material number per triangle, MTL
 * materials with their texture statements - and follows tiny_obj_loader's conventions where they decide the
 * result: quads are split along the shorter diagonal, `g` / `o` start a new shape, `usemtl` switches the
 * material of the following faces, material defaults (colours 0, d 1, Ns 1, Ni 1), Tr = 1 - d, texture options
 * -s, -o, -bm. Polygons with more than four corners are split as a fan (tiny_obj_loader clips ears).
 * Everything here is host code; the imported objects are ordinary Mesh / Material / Texture objects. */
#pragma once

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#include <tgd/array.hpp>
#include <tgd/io.hpp>

#include "scene.hpp"
#include "transformation.hpp"
#include "texture_image.hpp"
#include "material_lambertian.hpp"
#include "material_glass.hpp"
#include "material_modphong.hpp"
#include "geometryproc.hpp"
#include "mesh.hpp"
#include "color.hpp"

namespace WurblPT {

namespace detail {

struct ObjTextureStatement
{
    std::string name;
    float scale[3] = { 1.0f, 1.0f, 1.0f };      /* -s */
    float offset[3] = { 0.0f, 0.0f, 0.0f };     /* -o */
    float bumpMultiplier = 1.0f;                /* -bm */
};

struct ObjMaterial
{
    std::string name;
    float diffuse[3] = { 0.0f, 0.0f, 0.0f };
    float specular[3] = { 0.0f, 0.0f, 0.0f };
    float emission[3] = { 0.0f, 0.0f, 0.0f };
    float transmittance[3] = { 0.0f, 0.0f, 0.0f };
    float shininess = 1.0f;
    float dissolve = 1.0f;
    float ior = 1.0f;
    ObjTextureStatement diffuseTex, specularTex, shininessTex, alphaTex, emissiveTex, bumpTex, normalTex;
};

struct ObjCorner { int v, n, t; };      /* zero-based, -1 = absent */

struct ObjShape
{
    std::string name;
    std::vector<ObjCorner> corners;     /* three per triangle */
    std::vector<int> materialIds;       /* one per triangle, -1 = none */
};

class ObjFile
{
public:
    std::vector<float> positions, normals, texcoords;
    std::vector<ObjShape> shapes;
    std::vector<ObjMaterial> materials;
    std::string warning, error;

private:
    std::map<std::string, int> _materialByName;

    static float number(std::istringstream& in, float fallback = 0.0f)
    {
        std::string token;
        if (!(in >> token))
            return fallback;
        return float(std::strtod(token.c_str(), nullptr));
    }

    static std::string restOfLine(std::istringstream& in)
    {
        std::string rest;
        std::getline(in, rest);
        size_t first = rest.find_first_not_of(" \t");
        size_t last = rest.find_last_not_of(" \t\r");
        return first == std::string::npos ? std::string() : rest.substr(first, last - first + 1);
    }

    /* "name" preceded by options such as -s 2 2 1, -o 0.5 0 0, -bm 0.2, -clamp on */
    static void textureStatement(std::istringstream& in, ObjTextureStatement& out)
    {
        std::string token;
        while (in >> token) {
            if (token == "-s" || token == "-o" || token == "-t") {
                /* one to three numbers; unread ones keep their defaults */
                float values[3] = { token == "-s" ? 1.0f : 0.0f, token == "-s" ? 1.0f : 0.0f, token == "-s" ? 1.0f : 0.0f };
                for (int i = 0; i < 3; i++) {
                    std::streampos mark = in.tellg();
                    std::string v;
                    if (!(in >> v))
                        break;
                    char* end = nullptr;
                    float f = float(std::strtod(v.c_str(), &end));
                    if (end == v.c_str() || *end != '\0') {
                        in.clear();
                        in.seekg(mark);
                        break;
                    }
                    values[i] = f;
                }
                if (token == "-s")
                    std::memcpy(out.scale, values, sizeof(values));
                else if (token == "-o")
                    std::memcpy(out.offset, values, sizeof(values));
            } else if (token == "-bm") {
                out.bumpMultiplier = number(in, 1.0f);
            } else if (token == "-mm") {
                number(in); number(in);
            } else if (token == "-blendu" || token == "-blendv" || token == "-clamp" || token == "-boost"
                    || token == "-imfchan" || token == "-texres" || token == "-type" || token == "-colorspace") {
                std::string ignored;
                in >> ignored;
            } else {
                std::string tail = restOfLine(in);
                out.name = tail.empty() ? token : token + " " + tail;
                return;
            }
        }
    }

    void readMaterials(const std::string& path)
    {
        std::ifstream in(path);
        if (!in) {
            warning += "Material file [ " + path + " ] not found.\n";
            return;
        }
        std::string line;
        ObjMaterial current;
        bool open = false;
        auto close = [&]() {
            if (open) {
                _materialByName[current.name] = int(materials.size());
                materials.push_back(current);
            }
        };
        while (std::getline(in, line)) {
            std::istringstream ls(line);
            std::string key;
            if (!(ls >> key) || key[0] == '#')
                continue;
            if (key == "newmtl") {
                close();
                current = ObjMaterial();
                current.name = restOfLine(ls);
                open = true;
            } else if (key == "Kd") {
                for (float& c : current.diffuse) c = number(ls);
            } else if (key == "Ks") {
                for (float& c : current.specular) c = number(ls);
            } else if (key == "Ke") {
                for (float& c : current.emission) c = number(ls);
            } else if (key == "Kt" || key == "Tf") {
                for (float& c : current.transmittance) c = number(ls);
            } else if (key == "Ns") {
                current.shininess = number(ls);
            } else if (key == "Ni") {
                current.ior = number(ls);
            } else if (key == "d") {
                current.dissolve = number(ls, 1.0f);
            } else if (key == "Tr") {
                current.dissolve = 1.0f - number(ls);
            } else if (key == "map_Kd") {
                textureStatement(ls, current.diffuseTex);
            } else if (key == "map_Ks") {
                textureStatement(ls, current.specularTex);
            } else if (key == "map_Ns") {
                textureStatement(ls, current.shininessTex);
            } else if (key == "map_d") {
                textureStatement(ls, current.alphaTex);
            } else if (key == "map_Ke") {
                textureStatement(ls, current.emissiveTex);
            } else if (key == "map_bump" || key == "map_Bump" || key == "bump") {
                textureStatement(ls, current.bumpTex);
            } else if (key == "norm") {
                textureStatement(ls, current.normalTex);
            }
        }
        close();
    }

    /* "v", "v/t", "v//n" or "v/t/n"; one-based, negative = relative to the end */
    bool corner(const std::string& token, ObjCorner& c) const
    {
        int values[3] = { 0, 0, 0 };
        int field = 0;
        size_t pos = 0;
        while (field < 3 && pos <= token.size()) {
            size_t slash = token.find('/', pos);
            std::string part = token.substr(pos, slash == std::string::npos ? std::string::npos : slash - pos);
            if (!part.empty())
                values[field] = std::atoi(part.c_str());
            field++;
            if (slash == std::string::npos)
                break;
            pos = slash + 1;
        }
        auto resolve = [](int index, size_t count) { return index > 0 ? index - 1 : index < 0 ? int(count) + index : -1; };
        c.v = resolve(values[0], positions.size() / 3);
        c.t = resolve(values[1], texcoords.size() / 2);
        c.n = resolve(values[2], normals.size() / 3);
        /* a texture coordinate or normal index that points past what has been read so far is treated as absent
         * (the collector would read outside the arrays); the position index must be valid */
        if (c.t >= 0 && size_t(c.t) >= texcoords.size() / 2) {
            std::fprintf(stderr, "OBJ import: texture coordinate index %d out of range, ignored\n", values[1]);
            c.t = -1;
        }
        if (c.n >= 0 && size_t(c.n) >= normals.size() / 3) {
            std::fprintf(stderr, "OBJ import: normal index %d out of range, ignored\n", values[2]);
            c.n = -1;
        }
        return c.v >= 0 && size_t(c.v) < positions.size() / 3;
    }

    float squaredDistance(int a, int b) const
    {
        float sum = 0.0f;
        for (int k = 0; k < 3; k++) {
            float d = positions[size_t(b) * 3 + k] - positions[size_t(a) * 3 + k];
            sum = (k == 0 ? d * d : sum + d * d);
        }
        return sum;
    }

    void face(ObjShape& shape, const std::vector<ObjCorner>& poly, int material)
    {
        auto triangle = [&](int a, int b, int c) {
            shape.corners.push_back(poly[a]);
            shape.corners.push_back(poly[b]);
            shape.corners.push_back(poly[c]);
            shape.materialIds.push_back(material);
        };
        if (poly.size() == 3) {
            triangle(0, 1, 2);
        } else if (poly.size() == 4) {
            /* the shorter diagonal splits the quad */
            if (squaredDistance(poly[0].v, poly[2].v) < squaredDistance(poly[1].v, poly[3].v)) {
                triangle(0, 1, 2);
                triangle(0, 2, 3);
            } else {
                triangle(0, 1, 3);
                triangle(1, 2, 3);
            }
        } else {
            for (size_t k = 1; k + 1 < poly.size(); k++)
                triangle(0, int(k), int(k + 1));
        }
    }

public:
    bool read(const std::string& fileName, const std::string& baseDir)
    {
        std::ifstream in(fileName);
        if (!in) {
            error = "Cannot open file [" + fileName + "]\n";
            return false;
        }
        ObjShape shape;
        int material = -1;
        auto flush = [&]() {
            if (!shape.corners.empty())
                shapes.push_back(shape);
            shape.corners.clear();
            shape.materialIds.clear();
        };
        std::string line;
        while (std::getline(in, line)) {
            std::istringstream ls(line);
            std::string key;
            if (!(ls >> key) || key[0] == '#')
                continue;
            if (key == "v") {
                for (int k = 0; k < 3; k++)
                    positions.push_back(number(ls));
            } else if (key == "vn") {
                for (int k = 0; k < 3; k++)
                    normals.push_back(number(ls));
            } else if (key == "vt") {
                for (int k = 0; k < 2; k++)
                    texcoords.push_back(number(ls));
            } else if (key == "f") {
                std::vector<ObjCorner> poly;
                std::string token;
                bool valid = true;
                while (ls >> token) {
                    ObjCorner c;
                    valid = corner(token, c) && valid;
                    poly.push_back(c);
                }
                if (!valid)
                    warning += "Face with invalid vertex index found.\n";
                else if (poly.size() >= 3)
                    face(shape, poly, material);
            } else if (key == "usemtl") {
                const std::string wanted = restOfLine(ls);
                auto it = _materialByName.find(wanted);
                material = (it == _materialByName.end() ? -1 : it->second);
                if (material < 0)
                    warning += "material [ '" + wanted + "' ] not found in .mtl\n";
            } else if (key == "mtllib") {
                readMaterials(baseDir + "/" + restOfLine(ls));
            } else if (key == "g" || key == "o") {
                flush();
                shape.name = restOfLine(ls);
            }
        }
        flush();
        return true;
    }
};

inline std::string directoryOf(const std::string& fileName)
{
    size_t sep = fileName.find_last_of("/\\");
    return sep == std::string::npos ? std::string(".") : fileName.substr(0, sep);
}

inline void printLines(const char* label, const std::string& text)
{
    std::istringstream in(text);
    std::string line;
    while (std::getline(in, line))
        if (!line.empty())
            fprintf(stderr, "  %s: %s\n", label, line.c_str());
}

/* Gathers the corners of one shape that use material `materialId` (or of all shapes / all materials) into
 * indexed vertex arrays: one vertex per distinct (position, normal, texcoord) triple, in order of first use. */
class ObjGeometryCollector
{
public:
    std::map<std::tuple<int, int, int>, unsigned int> seen;
    std::vector<vec3> positions, normals;
    std::vector<vec2> texcoords;
    std::vector<unsigned int> indices;
    bool haveNormals = true, haveTexCoords = true;

    /* returns false when the shape's normals turn out to be unusable (reported by the caller) */
    bool add(const ObjFile& obj, const ObjCorner& c, bool checkNormals)
    {
        bool normalsFine = true;
        const std::tuple<int, int, int> key = std::make_tuple(c.v, c.n, c.t);
        auto it = seen.find(key);
        if (it != seen.end()) {
            indices.push_back(it->second);
            return true;
        }
        const unsigned int fresh = seen.size();
        positions.push_back(vec3(obj.positions.data() + 3 * c.v));
        if (c.n < 0)
            haveNormals = false;
        if (c.t < 0)
            haveTexCoords = false;
        if (haveNormals) {
            const vec3 n = vec3(obj.normals.data() + 3 * c.n);
            if (checkNormals && (!all(isfinite(n)) || dot(n, n) < epsilon)) {
                haveNormals = false;
                normalsFine = false;
            } else {
                normals.push_back(normalize(n));        /* do not trust normals to be of unit length */
            }
        }
        if (haveTexCoords)
            texcoords.push_back(vec2(obj.texcoords.data() + 2 * c.t));
        indices.push_back(fresh);
        seen.insert(std::make_pair(key, fresh));
        return normalsFine;
    }

    void finish()
    {
        if (!haveNormals)
            normals = computeNormals(positions, indices);
        if (!haveTexCoords)
            texcoords.clear();
    }
};

}

/* a normal map (uint8 rgb) from a bump map (first component, uint8): central differences, clamped at the edges */
inline TGD::ArrayContainer toNormalMap(const TGD::ArrayContainer& bumpMap, float bumpScaling = 8.0f)
{
    assert(bumpMap.componentType() == TGD::uint8);
    TGD::Array<uint8_t> normalMap(bumpMap.dimensions(), 3);
    const int w = normalMap.dimension(0);
    const int h = normalMap.dimension(1);
    auto heightAt = [&](int x, int y) {
        x = std::min(std::max(x, 0), w - 1);
        y = std::min(std::max(y, 0), h - 1);
        return byte_to_float(static_cast<const uint8_t*>(bumpMap.get(size_t(y) * w + x))[0]);
    };
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            const vec3 alongX = vec3(2.0, 0.0, bumpScaling * (heightAt(x + 1, y) - heightAt(x - 1, y)));
            const vec3 alongY = vec3(0.0, 2.0, bumpScaling * (heightAt(x, y + 1) - heightAt(x, y - 1)));
            const vec3 n = normalize(cross(alongX, alongY));
            uint8_t* rgb = normalMap.get<uint8_t>(size_t(y) * w + x);
            for (int c = 0; c < 3; c++)
                rgb[c] = float_to_byte(0.5f * (n[c] + 1.0f));
        }
    }
    return normalMap;
}

/* One Texture per (file, factor, offset, sRGB handling, bump scaling); `textureFileMap` holds the loaded
 * images (empty container = file could not be loaded: a grey dummy texture takes its place). */
inline Texture* importTexture(
        std::map<std::string, Texture*>& textureMap,
        const std::map<std::string, TGD::ArrayContainer>& textureFileMap,
        const std::string& name,
        const float factor[3],
        const float offset[3],
        int* componentCount = nullptr,
        LinearizeSRGBType linearizeSRGBType = LinearizeSRGB_Auto,
        float bumpScaling = -1.0f /* < 0 means this is no bump map */)
{
    static const char* srgbNames[3] = { "on", "off", "auto" };
    const std::string cacheName = name
        + "_factor=" + std::to_string(factor[0]) + ',' + std::to_string(factor[1])
        + "_offset=" + std::to_string(offset[0]) + ',' + std::to_string(offset[1])
        + "_linsrgb=" + srgbNames[linearizeSRGBType == LinearizeSRGB_On ? 0 : linearizeSRGBType == LinearizeSRGB_Off ? 1 : 2]
        + "_bumpscal=" + std::to_string(bumpScaling);
    auto cached = textureMap.find(cacheName);
    fprintf(stderr, "    texture %s: %s\n", cacheName.c_str(), cached != textureMap.end() ? "found in cache" : "creating");
    if (cached != textureMap.end())
        return cached->second;
    auto file = textureFileMap.find(name);
    assert(file != textureFileMap.end());
    TGD::ArrayContainer img = file->second;
    Texture* tex = nullptr;
    const vec2 coordFactor(factor[0], factor[1]), coordOffset(offset[0], offset[1]);
    if (img.elementCount() == 0) {
        tex = new TextureConstant(vec4(0.5f));
    } else {
        if (componentCount)
            *componentCount = img.componentCount();
        if (bumpScaling > 0.0f && img.componentType() != TGD::uint8) {
            fprintf(stderr, "    texture %s: cannot handle bump/normal that is not uint8\n", cacheName.c_str());
            tex = new TextureConstant(vec4(0.5f, 0.5f, 1.0f, 0.0f));
        } else {
            if (bumpScaling > 0.0f) {
                /* grey images (also when stored as rgb) are bump maps and become normal maps first */
                bool grey = true;
                if (img.componentCount() >= 3) {
                    for (size_t e = 0; e < img.elementCount() && grey; e++) {
                        const uint8_t* px = img.get<uint8_t>(e);
                        grey = (px[0] == px[1] && px[0] == px[2]);
                    }
                }
                if (grey)
                    img = toNormalMap(img, bumpScaling);
            }
            tex = createTextureImage(img, linearizeSRGBType, coordFactor, coordOffset);
        }
    }
    if (img.elementCount() == 0 || !tex)
        fprintf(stderr, "    texture %s: replaced with dummy texture\n", cacheName.c_str());
    textureMap.insert(std::make_pair(cacheName, tex));
    return tex;
}

/*! \brief Flag: Disable light sources. */
constexpr unsigned int ImportBitDisableLightSources = (1 << 0);
/*! \brief Flag: Do not mark light emitting materials as hot spots. */
constexpr unsigned int ImportBitDisableHotSpots     = (1 << 1);
/*! \brief Flag: Import all materials as two-sided, e.g. for San Miguel. */
constexpr unsigned int ImportBitTwoSidedMaterials   = (1 << 2);
/*! \brief Flag: Tf=0 (or d=1) means opaque, e.g. for San Miguel. */
constexpr unsigned int ImportBitInvertedTf          = (1 << 3);
/*! \brief Flag: Use glass material for transparency instead of ModPhong. */
constexpr unsigned int ImportBitWithGlass           = (1 << 4);

/* Imports a triangle scene: one material per MTL entry (Lambertian where that suffices, glass on request,
 * ModPhong otherwise), one Mesh + MeshInstance per (material, shape) pair, emitters as hot spots. */
inline bool importIntoScene(Scene& scene, const std::string& filename,
        const Transformation& transformation = Transformation(),
        unsigned int importBits = 0)
{
    fprintf(stderr, "%s: importing...\n", filename.c_str());
    const std::string basedir = detail::directoryOf(filename);
    detail::ObjFile obj;
    const bool parsed = obj.read(filename, basedir);
    detail::printLines("warning", obj.warning);
    if (!parsed) {
        if (obj.error.empty())
            fprintf(stderr, "  unknown error\n");
        else
            detail::printLines("error", obj.error);
        fprintf(stderr, "%s: import failure\n", filename.c_str());
        return false;
    }

    /* texture files, each loaded once */
    fprintf(stderr, "  loading texture files...\n");
    std::set<std::string> textureFiles;
    for (const detail::ObjMaterial& M : obj.materials)
        for (const detail::ObjTextureStatement* t : { &M.normalTex, &M.bumpTex, &M.diffuseTex, &M.specularTex,
                &M.shininessTex, &M.alphaTex, &M.emissiveTex })
            if (!t->name.empty())
                textureFiles.insert(t->name);
    std::map<std::string, TGD::ArrayContainer> textureFileMap;
    for (const std::string& name : textureFiles) {
        std::string path = basedir + "/" + name;
        for (char& ch : path)
            if (ch == '\\')
                ch = '/';
        TGD::Error err;
        TGD::ArrayContainer img = TGD::load(path, TGD::TagList(), &err);
        fprintf(stderr, "    %s: %s\n", name.c_str(), (err != TGD::ErrorNone ? TGD::strerror(err) : "ok"));
        textureFileMap.insert(std::make_pair(name, img));
    }

    /* materials */
    std::vector<Material*> materials;
    std::vector<std::string> materialNames;
    std::vector<bool> materialIsLight, materialWantsTangents;
    std::map<std::string, Texture*> textureMap;
    const bool lightsOff = (importBits & ImportBitDisableLightSources) != 0;
    for (const detail::ObjMaterial& M : obj.materials) {
        fprintf(stderr, "  material '%s'...\n", M.name.c_str());
        vec3 dif = vec3(M.diffuse), spc = vec3(M.specular), emi = vec3(M.emission), tra = vec3(M.transmittance);
        float opa = M.dissolve;
        if (importBits & ImportBitInvertedTf)
            tra = vec3(1.0f) - tra;
        if (opa >= 1.0f && max(tra) < 1.0f) {
            opa = average(tra);
            tra = vec3(1.0f) - tra;
        }
        if (opa < 1.0f && max(tra) <= 0.0f)
            tra = (1.0f - opa) * dif;
        if (any(greaterThan(dif + spc, vec3(1.0f))))
            fprintf(stderr, "    warning: not energy conserving\n");
        auto fetch = [&](const detail::ObjTextureStatement& t, int* comps = nullptr,
                LinearizeSRGBType srgb = LinearizeSRGB_Auto, float bumpScaling = -1.0f) -> Texture* {
            return t.name.empty() ? nullptr
                : importTexture(textureMap, textureFileMap, t.name, t.scale, t.offset, comps, srgb, bumpScaling);
        };
        Texture* nrmTex = !M.normalTex.name.empty() ? fetch(M.normalTex, nullptr, LinearizeSRGB_Off)
            : fetch(M.bumpTex, nullptr, LinearizeSRGB_Off, M.bumpTex.bumpMultiplier);
        int difComps = 0;
        Texture* difTex = fetch(M.diffuseTex, &difComps);
        const bool difTexHasAlpha = (difTex && (difComps == 2 || difComps == 4));
        const bool emits = max(emi) > 0.0f || !M.emissiveTex.name.empty();
        Material* made = nullptr;
        bool isLight = false;
        if (!difTexHasAlpha && max(spc) <= 0.0f && M.specularTex.name.empty() && (lightsOff || !emits)
                && opa >= 1.0f && M.alphaTex.name.empty()) {
            fprintf(stderr, "    using Lambertian material model\n");
            made = new MaterialLambertian(dif, difTex);
        } else if ((importBits & ImportBitWithGlass) && !difTex && M.specularTex.name.empty() && !emits
                && opa < 1.0f && M.alphaTex.name.empty()) {
            fprintf(stderr, "    using Glass material model\n");
            made = new MaterialGlass(MaterialGlass::transparentColorToAbsorption(dif), M.ior);
        } else {
            fprintf(stderr, "    using ModPhong material model\n");
            MaterialModPhong* mat = new MaterialModPhong;
            int comps = 0;
            mat->haveNIR = false;
            mat->diffuse = vec4(dif, 0.0f);
            mat->diffuseTex = difTex;
            mat->diffuseTexHasAlpha = difTexHasAlpha;
            mat->specular = vec4(spc, 0.0f);
            mat->specularTex = fetch(M.specularTex, &comps);
            mat->specularTexHasAlpha = (mat->specularTex && (comps == 2 || comps == 4));
            mat->shininess = M.shininess;
            mat->shininessTex = fetch(M.shininessTex, nullptr, LinearizeSRGB_Off);
            mat->opacity = opa;
            mat->opacityTex = fetch(M.alphaTex, nullptr, LinearizeSRGB_Off);
            mat->indexOfRefraction = M.ior;
            mat->transmissive = vec4(tra, 0.0f);
            if (!lightsOff) {
                mat->emissive = vec4(emi, 0.0f);
                mat->emissiveTex = fetch(M.emissiveTex);
            }
            isLight = dot(mat->emissive, mat->emissive) > 0.0f || mat->emissiveTex;
            made = mat;
        }
        made->normalTex = nrmTex;
        materials.push_back(made);
        materialNames.push_back(M.name);
        materialIsLight.push_back(isLight);
        materialWantsTangents.push_back(nrmTex != nullptr);
    }
    for (const auto& entry : textureMap)
        scene.take(entry.second);
    for (size_t i = 0; i < materials.size(); i++)
        scene.take(materials[i], materialNames[i]);
    if (importBits & ImportBitTwoSidedMaterials) {
        for (size_t i = 0; i < materials.size(); i++) {
            materials[i] = new MaterialTwoSided(materials[i], materials[i]);
            scene.take(materials[i], materialNames[i]);
        }
    }

    /* geometry: everything of one shape that uses one material becomes one MeshInstance */
    fprintf(stderr, "  importing %zu shapes\n", obj.shapes.size());
    Material* nullMaterial = scene.take(new MaterialLambertian(vec4(0.5f)));
    for (int matId = -1; matId < int(materials.size()); matId++) {
        if (matId < 0)
            fprintf(stderr, "  importing all geometry that uses no material\n");
        else
            fprintf(stderr, "  importing all geometry that uses material '%s'\n", obj.materials[matId].name.c_str());
        for (size_t s = 0; s < obj.shapes.size(); s++) {
            const detail::ObjShape& shape = obj.shapes[s];
            detail::ObjGeometryCollector geometry;
            for (size_t i = 0; i < shape.corners.size(); i++) {
                if (shape.materialIds[i / 3] != matId)
                    continue;
                if (!geometry.add(obj, shape.corners[i], true))
                    fprintf(stderr, "      warning: invalid normals in shape %zu '%s'\n", s, shape.name.c_str());
            }
            if (geometry.indices.empty())
                continue;
            if (!geometry.haveNormals)
                fprintf(stderr, "      computing normals for shape %zu '%s'\n", s, shape.name.c_str());
            geometry.finish();
            Material* mat = (matId < 0 ? nullMaterial : materials[matId]);
            const bool isLight = (matId >= 0 && materialIsLight[matId]);
            const bool wantsTangents = (matId >= 0 && materialWantsTangents[matId]);
            Mesh* mesh = new Mesh(geometry.positions, geometry.normals, geometry.texcoords, geometry.indices,
                    transformation, wantsTangents);
            scene.take(mesh);
            scene.take(new MeshInstance(mesh, mat),
                    (isLight && !(importBits & ImportBitDisableHotSpots)) ? HotSpot : ColdSpot);
        }
    }
    fprintf(stderr, "%s: import done\n", filename.c_str());
    return true;
}

/* The triangle geometry of a file (or of its shape `shapeName`) as one Mesh; materials are ignored. */
inline Mesh* importGeometry(const std::string& filename,
        const std::string& shapeName = std::string(),
        const Transformation& T = Transformation())
{
    detail::ObjFile obj;
    const bool parsed = obj.read(filename, detail::directoryOf(filename));
    detail::printLines("warning", obj.warning);
    if (!parsed) {
        if (obj.error.empty())
            fprintf(stderr, "  unknown error\n");
        else
            detail::printLines("error", obj.error);
        fprintf(stderr, "%s: import failure\n", filename.c_str());
        return nullptr;
    }
    detail::ObjGeometryCollector geometry;
    for (const detail::ObjShape& shape : obj.shapes) {
        if (!shapeName.empty() && shapeName != shape.name)
            continue;
        for (const detail::ObjCorner& c : shape.corners)
            geometry.add(obj, c, false);
    }
    geometry.finish();
    return new Mesh(geometry.positions, geometry.normals, geometry.texcoords, geometry.indices, T);
}

}
