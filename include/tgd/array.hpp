/*
 * Minimal stand-in for the subset of libtgd (marlam/tgd, <tgd/array.hpp>) that the
 * WurblPT API surface touches: an N-dimensional multi-component array with tag lists.
 *
 * libtgd is an external dependency of the reference (find_package(libtgd), reference
 * libwurblpt/CMakeLists.txt:23) that is neither vendored nor installed here. It holds
 * no path-tracing arithmetic; only storage. This file is written from the call sites
 * in the reference (sensor_rgb.hpp:36-96, sensor_tof_amcw.hpp:73-213,
 * texture_image.hpp:45-180, wurblpt.hpp:301,425-435, postproc.hpp:44-121), not from
 * libtgd's sources.
 */
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace TGD {

enum Type {
    int8 = 0, uint8 = 1, int16 = 2, uint16 = 3, int32 = 4, uint32 = 5,
    int64 = 6, uint64 = 7, float32 = 8, float64 = 9
};

inline size_t typeSize(Type t)
{
    switch (t) {
    case int8: case uint8: return 1;
    case int16: case uint16: return 2;
    case int32: case uint32: case float32: return 4;
    default: return 8;
    }
}

template<typename T> constexpr Type typeFromTemplate();
template<> constexpr Type typeFromTemplate<int8_t>() { return int8; }
template<> constexpr Type typeFromTemplate<uint8_t>() { return uint8; }
template<> constexpr Type typeFromTemplate<int16_t>() { return int16; }
template<> constexpr Type typeFromTemplate<uint16_t>() { return uint16; }
template<> constexpr Type typeFromTemplate<int32_t>() { return int32; }
template<> constexpr Type typeFromTemplate<uint32_t>() { return uint32; }
template<> constexpr Type typeFromTemplate<int64_t>() { return int64; }
template<> constexpr Type typeFromTemplate<uint64_t>() { return uint64; }
template<> constexpr Type typeFromTemplate<float>() { return float32; }
template<> constexpr Type typeFromTemplate<double>() { return float64; }

class TagList
{
private:
    std::map<std::string, std::string> _tags;

public:
    TagList() {}

    size_t size() const { return _tags.size(); }
    void clear() { _tags.clear(); }
    void set(const std::string& key, const std::string& value) { _tags[key] = value; }
    void unset(const std::string& key) { _tags.erase(key); }
    bool contains(const std::string& key) const { return _tags.find(key) != _tags.end(); }
    const std::string& value(const std::string& key, const std::string& fallback = std::string()) const
    {
        auto it = _tags.find(key);
        return it == _tags.end() ? fallback : it->second;
    }
    std::map<std::string, std::string>::const_iterator cbegin() const { return _tags.cbegin(); }
    std::map<std::string, std::string>::const_iterator cend() const { return _tags.cend(); }
    bool operator==(const TagList& o) const { return _tags == o._tags; }
};

class ArrayDescription
{
protected:
    std::vector<size_t> _dims;
    size_t _components;
    Type _type;
    TagList _globalTags;
    std::vector<TagList> _dimTags;
    std::vector<TagList> _compTags;

public:
    ArrayDescription() : _components(0), _type(uint8) {}

    ArrayDescription(const std::vector<size_t>& dims, size_t components, Type type) :
        _dims(dims), _components(components), _type(type),
        _dimTags(dims.size()), _compTags(components)
    {
    }

    size_t dimensionCount() const { return _dims.size(); }
    size_t dimension(size_t d) const { return _dims[d]; }
    const std::vector<size_t>& dimensions() const { return _dims; }
    size_t componentCount() const { return _components; }
    Type componentType() const { return _type; }
    size_t componentSize() const { return typeSize(_type); }
    size_t elementSize() const { return _components * componentSize(); }
    size_t elementCount() const
    {
        if (_dims.size() == 0)
            return 0;
        size_t n = 1;
        for (size_t d : _dims)
            n *= d;
        return n;
    }
    size_t dataSize() const { return elementCount() * elementSize(); }

    const TagList& globalTagList() const { return _globalTags; }
    TagList& globalTagList() { return _globalTags; }
    const TagList& dimensionTagList(size_t d) const { return _dimTags[d]; }
    TagList& dimensionTagList(size_t d) { return _dimTags[d]; }
    const TagList& componentTagList(size_t c) const { return _compTags[c]; }
    TagList& componentTagList(size_t c) { return _compTags[c]; }

    size_t toLinearIndex(const std::vector<size_t>& index) const
    {
        size_t lin = 0, stride = 1;
        for (size_t d = 0; d < _dims.size(); d++) {
            lin += index[d] * stride;
            stride *= _dims[d];
        }
        return lin;
    }
    size_t toLinearIndex(std::initializer_list<size_t> index) const
    {
        size_t lin = 0, stride = 1, d = 0;
        for (size_t i : index) {
            lin += i * stride;
            stride *= _dims[d++];
        }
        return lin;
    }
};

/* Storage is shared between copies (like libtgd's containers), so passing a
 * container by value into a texture does not duplicate the texels. */
class ArrayContainer : public ArrayDescription
{
protected:
    std::shared_ptr<unsigned char[]> _data;

public:
    ArrayContainer() {}

    explicit ArrayContainer(const ArrayDescription& desc) :
        ArrayDescription(desc), _data(new unsigned char[desc.dataSize() > 0 ? desc.dataSize() : 1]())
    {
    }

    ArrayContainer(const std::vector<size_t>& dims, size_t components, Type type) :
        ArrayContainer(ArrayDescription(dims, components, type))
    {
    }

    const ArrayDescription& description() const { return *this; }

    const void* data() const { return _data.get(); }
    void* data() { return _data.get(); }

    const void* get(size_t elementIndex) const { return _data.get() + elementIndex * elementSize(); }
    void* get(size_t elementIndex) { return _data.get() + elementIndex * elementSize(); }

    template<typename T> const T* get(size_t elementIndex) const { return static_cast<const T*>(get(elementIndex)); }
    template<typename T> T* get(size_t elementIndex) { return static_cast<T*>(get(elementIndex)); }
    template<typename T> const T* get(std::initializer_list<size_t> index) const { return get<T>(toLinearIndex(index)); }
    template<typename T> T* get(std::initializer_list<size_t> index) { return get<T>(toLinearIndex(index)); }
    template<typename T> const T* get(const std::vector<size_t>& index) const { return get<T>(toLinearIndex(index)); }
    template<typename T> T* get(const std::vector<size_t>& index) { return get<T>(toLinearIndex(index)); }
    template<typename T> T get(size_t elementIndex, size_t componentIndex) const { return get<T>(elementIndex)[componentIndex]; }
    template<typename T> T get(std::initializer_list<size_t> index, size_t componentIndex) const { return get<T>(index)[componentIndex]; }

    template<typename T> void set(size_t elementIndex, std::initializer_list<T> values)
    {
        T* p = get<T>(elementIndex);
        size_t c = 0;
        for (T v : values)
            p[c++] = v;
    }
    template<typename T> void set(std::initializer_list<size_t> index, std::initializer_list<T> values)
    {
        set<T>(toLinearIndex(index), values);
    }
    template<typename T> void set(size_t elementIndex, size_t componentIndex, T value)
    {
        get<T>(elementIndex)[componentIndex] = value;
    }

    ArrayContainer deepCopy() const
    {
        ArrayContainer r(description());
        if (dataSize() > 0)
            std::memcpy(r.data(), data(), dataSize());
        return r;
    }
};

template<typename T>
class Array : public ArrayContainer
{
public:
    Array() {}

    explicit Array(const ArrayDescription& desc) :
        ArrayContainer(ArrayDescription(desc.dimensions(), desc.componentCount(), typeFromTemplate<T>()))
    {
        _globalTags = desc.globalTagList();
        for (size_t d = 0; d < dimensionCount(); d++)
            _dimTags[d] = desc.dimensionTagList(d);
        for (size_t c = 0; c < componentCount(); c++)
            _compTags[c] = desc.componentTagList(c);
    }

    Array(const std::vector<size_t>& dims, size_t components) :
        ArrayContainer(dims, components, typeFromTemplate<T>())
    {
    }

    Array(std::initializer_list<size_t> dims, size_t components) :
        ArrayContainer(std::vector<size_t>(dims), components, typeFromTemplate<T>())
    {
    }

    /* View an untyped container as a typed one (shares storage). */
    explicit Array(const ArrayContainer& container) : ArrayContainer(container)
    {
    }

    const T* operator[](size_t elementIndex) const { return ArrayContainer::get<T>(elementIndex); }
    T* operator[](size_t elementIndex) { return ArrayContainer::get<T>(elementIndex); }
    const T* operator[](std::initializer_list<size_t> index) const { return ArrayContainer::get<T>(index); }
    T* operator[](std::initializer_list<size_t> index) { return ArrayContainer::get<T>(index); }
    const T* operator[](const std::vector<size_t>& index) const { return ArrayContainer::get<T>(index); }
    T* operator[](const std::vector<size_t>& index) { return ArrayContainer::get<T>(index); }

    const T* cbegin() const { return static_cast<const T*>(data()); }
    const T* cend() const { return cbegin() + elementCount() * componentCount(); }
    T* begin() { return static_cast<T*>(data()); }
    T* end() { return begin() + elementCount() * componentCount(); }
};

}
