// vsg_stub.hpp -- TEST INFRASTRUCTURE: just enough of vsg's container surface for the reference's HOST conversion code
// (source/io/RenderIO.cpp: GBufferIO::convert_normal_to_spherical, GBufferIO::compress_albedo, the position -> depth
// block of import_g_buffer_position) to compile unchanged.  The arithmetic types -- vsg::vec2/3/4, ubvec4, mat4, length(),
// the vec4 * float and float-vector -> byte-vector conversions the code relies on -- are vsg's OWN headers
// (external/vsg/include/vsg/maths, header-only), included from the reference tree; only the array / ref_ptr plumbing
// below is this repository's.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <iostream>
#include <memory>
#include <optional>
#include <vector>

#include <vsg/maths/mat4.h>
#include <vsg/maths/vec2.h>
#include <vsg/maths/vec3.h>
#include <vsg/maths/vec4.h>

enum { VK_FORMAT_R32_SFLOAT = 100, VK_FORMAT_R32G32_SFLOAT = 103, VK_FORMAT_R32G32B32A32_SFLOAT = 109, VK_FORMAT_R8G8B8A8_UNORM = 37 };

namespace vsg {

template <class T>
class ref_ptr {
public:
    ref_ptr() = default;
    ref_ptr(std::shared_ptr<T> p) : _p(std::move(p)) {}
    template <class U> ref_ptr(const ref_ptr<U>& o) : _p(o.shared()) {}
    bool valid() const { return (bool)_p; }
    explicit operator bool() const { return (bool)_p; }
    bool operator!() const { return !_p; }
    T* operator->() const { return _p.get(); }
    T* get() const { return _p.get(); }
    template <class U> ref_ptr<U> cast() const { return ref_ptr<U>(std::dynamic_pointer_cast<U>(_p)); }
    const std::shared_ptr<T>& shared() const { return _p; }
private:
    std::shared_ptr<T> _p;
};

class Data {
public:
    struct Layout { int format; };
    virtual ~Data() = default;
    virtual std::size_t valueCount() const = 0;
    virtual uint32_t width() const = 0;
    virtual uint32_t height() const = 0;
};

template <class T>
class Array2D : public Data {
public:
    Array2D(uint32_t w, uint32_t h, T* data, Layout layout) : _w(w), _h(h), _data(data), layout(layout) {}
    ~Array2D() override { delete[] _data; }                     // vsg's Array2D takes ownership of the new[]'d block too
    static ref_ptr<Array2D> create(uint32_t w, uint32_t h, T* data, Layout layout) { return ref_ptr<Array2D>(std::make_shared<Array2D>(w, h, data, layout)); }
    std::size_t valueCount() const override { return (std::size_t)_w * _h; }
    uint32_t width() const override { return _w; }
    uint32_t height() const override { return _h; }
    T* data() { return _data; }
    Layout layout;
private:
    uint32_t _w, _h;
    T* _data;
};
using floatArray2D = Array2D<float>;
using vec2Array2D = Array2D<vec2>;
using vec4Array2D = Array2D<vec4>;
using ubvec4Array2D = Array2D<ubvec4>;
using usvec4Array2D = Array2D<usvec4>;
using uivec4Array2D = Array2D<uivec4>;

}  // namespace vsg
