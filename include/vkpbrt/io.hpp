// io.hpp -- offline sequence I/O of the denoising path in the C++ layer: the reference's way
// of feeding pre-rendered sequences to the modules, with the reference's class and function names
// (source/io/RenderIO.hpp:28-121):
//
//   MatrixIO::import_matrices / export_matrices        RenderIO.cpp:593-711  (JSON layout of the reference, BMFR-dataset text)
//   GBufferIO::import_g_buffer_depth / _position       RenderIO.cpp:6-158    (printf-style per-frame EXR file names)
//   IlluminationBufferIO::import_illumination          RenderIO.cpp:501-548
//   OfflineGBuffer::upload_to_g_buffer                 RenderIO.cpp:718-746  (upload_to_g_buffer_command)
//   OfflineIllumination::upload_to_illumination_buffer RenderIO.cpp:384-399
//   GBufferIO::export_g_buffer + its conversions       RenderIO.cpp:213-382  (host loops, as in the reference)
//   OfflineGBuffer::download_from_g_buffer             RenderIO.cpp:748-928  (download command + transfer_staging_data_to)
//   OfflineIllumination::download_from_illumination_buffer RenderIO.cpp:401-467
//
// What differs from the reference: the conversions between what the files hold and what the G-buffer holds --
// world position -> Euclidean depth (:101-120), cartesian -> spherical normals (:160-178), float albedo -> rgba8
// (:180-211) -- do not run in host loops at import time but as ONE device launch at upload time
// (GBuffer::import_planes -> vkpbrt_gbuffer_import_record); an OfflineGBuffer therefore keeps the decoded rgba32f planes.
// The EXR codec is a small reader / writer of its own (the reference uses vsgXchange::openexr): single-part scanline
// files, HALF / FLOAT channels, compression NONE / RLE / ZIPS / ZIP -- what OpenEXR writers produce by default
// (OpenCV, which the Python layer and the tests use, writes ZIP); tiled, deep, multi-part, PIZ / PXR24 / B44 / DWA files
// are refused with a message.  Needs zlib (-lz).  Header-only; link libvkpbrt_b200.so.
#pragma once

#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "vkpbrt.hpp"

namespace vkpbrt {

// ---- a JSON reader just large enough for the matrix files -----------------------------------------------------------
namespace json {
struct Value {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    bool b = false;
    double number = 0;
    std::string string;
    std::vector<Value> array;
    std::map<std::string, Value> object;
    const Value& at(const std::string& key) const
    {
        auto it = object.find(key);
        if (type != Object || it == object.end()) throw std::runtime_error("json: missing key '" + key + "'");
        return it->second;
    }
    bool has(const std::string& key) const { return type == Object && object.count(key); }
};
class Parser {
public:
    explicit Parser(const std::string& text) : s(text) {}
    Value parse()
    {
        Value v = value();
        ws();
        if (i != s.size()) fail("trailing characters");
        return v;
    }
private:
    const std::string& s;
    size_t i = 0;
    [[noreturn]] void fail(const char* what) const { throw std::runtime_error(std::string("json: ") + what + " at offset " + std::to_string(i)); }
    void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) ++i; }
    Value value()
    {
        ws();
        if (i >= s.size()) fail("unexpected end");
        Value v;
        const char c = s[i];
        if (c == '{') {
            v.type = Value::Object;
            ++i; ws();
            if (i < s.size() && s[i] == '}') { ++i; return v; }
            for (;;) {
                ws();
                if (i >= s.size() || s[i] != '"') fail("expected a key");
                const std::string key = str();
                ws();
                if (i >= s.size() || s[i] != ':') fail("expected ':'");
                ++i;
                v.object[key] = value();
                ws();
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == '}') { ++i; return v; }
                fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            v.type = Value::Array;
            ++i; ws();
            if (i < s.size() && s[i] == ']') { ++i; return v; }
            for (;;) {
                v.array.push_back(value());
                ws();
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == ']') { ++i; return v; }
                fail("expected ',' or ']'");
            }
        }
        if (c == '"') { v.type = Value::String; v.string = str(); return v; }
        if (s.compare(i, 4, "true") == 0) { v.type = Value::Bool; v.b = true; i += 4; return v; }
        if (s.compare(i, 5, "false") == 0) { v.type = Value::Bool; i += 5; return v; }
        if (s.compare(i, 4, "null") == 0) { i += 4; return v; }
        char* end = nullptr;
        v.number = std::strtod(s.c_str() + i, &end);
        if (end == s.c_str() + i) fail("unexpected character");
        v.type = Value::Number;
        i = (size_t)(end - s.c_str());
        return v;
    }
    std::string str()
    {
        std::string out;
        for (++i; i < s.size() && s[i] != '"'; ++i) {
            if (s[i] == '\\' && i + 1 < s.size()) {
                const char e = s[++i];
                out += e == 'n' ? '\n' : e == 't' ? '\t' : e;
            } else out += s[i];
        }
        if (i >= s.size()) fail("unterminated string");
        ++i;
        return out;
    }
};
}  // namespace json

// ---- matrices ---------------------------------------------------------------------------------------------------------
using CameraMatricesVec = std::vector<CameraMatrices>;

class MatrixIO {   // source/io/RenderIO.hpp:37-42
public:
    // RenderIO.cpp:593-666.  A file that cannot be opened yields {} (the reference prints a message and returns {});
    // malformed JSON throws, as nlohmann::json does.
    static std::vector<CameraMatrices> import_matrices(const std::string& matrix_path)
    {
        std::ifstream f(matrix_path);
        if (!f) { std::cout << "Matrix file " << matrix_path << " unable to open." << std::endl; return {}; }
        std::stringstream ss;
        ss << f.rdbuf();
        const std::string text = ss.str();
        std::vector<CameraMatrices> out;
        if (matrix_path.size() >= 5 && matrix_path.compare(matrix_path.size() - 5, 5, ".json") == 0) {
            const json::Value doc = json::Parser(text).parse();
            const int n = (int)doc.at("amtOfFrames").number;
            const json::Value& mats = doc.at("matrices");
            if (mats.type != json::Value::Array || (int)mats.array.size() < n) throw std::runtime_error("json: 'matrices' is shorter than amtOfFrames");
            for (int i = 0; i < n; ++i) {
                const json::Value& m = mats.array[i];
                CameraMatrices cm;
                cm.view = to_mat(m.at("view"));
                cm.inv_view = to_mat(m.at("invView"));
                if (m.has("type") && m.at("type").string == "ModelView+Projection") { cm.proj = to_mat(m.at("proj")); cm.inv_proj = to_mat(m.at("invProj")); }
                out.push_back(cm);
            }
            return out;
        }
        // BMFR-dataset text (:639-664): tokens separated by whitespace, optional trailing ',' and leading '{'; every token
        // that starts with a digit or '-' is a number (std::stof: longest valid prefix), 16 numbers are one combined
        // view-projection matrix, whose inverse is computed on load
        std::istringstream in(text);
        std::string tok;
        float cur[16];
        int count = 0;
        while (in >> tok) {
            if (!tok.empty() && tok.back() == ',') tok.pop_back();
            if (!tok.empty() && tok.front() == '{') tok.erase(0, 1);
            if (tok.empty() || !(std::isdigit((unsigned char)tok[0]) || tok[0] == '-')) continue;
            cur[count++] = std::stof(tok);
            if (count == 16) {
                CameraMatrices cm;
                std::memcpy(cm.view.m, cur, sizeof cur);
                inverse(cm.view.m, cm.inv_view.m);
                out.push_back(cm);
                count = 0;
            }
        }
        return out;
    }
    // RenderIO.cpp:668-711 (same keys and values; a matrix without projection is typed "ModelViewProjection")
    static bool export_matrices(const std::string& matrix_path, const CameraMatricesVec& matrices)
    {
        std::ofstream f(matrix_path);
        if (!f) { std::cout << "Matrix file " << matrix_path << " unable to open." << std::endl; return false; }
        f << "{\n    \"amtOfFrames\": " << matrices.size() << ",\n    \"matrices\": [";
        for (size_t i = 0; i < matrices.size(); ++i) {
            const CameraMatrices& m = matrices[i];
            const bool has_proj = m.proj && m.inv_proj;
            f << (i ? "," : "") << "\n        {\n            \"type\": \"" << (has_proj ? "ModelView+Projection" : "ModelViewProjection")
              << "\",\n            \"storageType\": \"ColumnMajor\",\n            \"view\": " << arr(m.view) << ",\n            \"invView\": " << arr(m.inv_view);
            if (has_proj) f << ",\n            \"proj\": " << arr(*m.proj) << ",\n            \"invProj\": " << arr(*m.inv_proj);
            f << "\n        }";
        }
        f << "\n    ]\n}\n";
        return (bool)f;
    }
    // vsg::inverse(mat4), what the reference's import calls (RenderIO.cpp:659): the library's restatement of it
    // (vkpbrt_mat4_inverse; external/vsg/src/vsg/maths/maths_transform.cpp:36-156)
    static void inverse(const float* m, float* inv) { check(vkpbrt_mat4_inverse(m, inv)); }
private:
    static mat4 to_mat(const json::Value& v)
    {
        if (v.type != json::Value::Array || v.array.size() != 16) throw std::runtime_error("json: a matrix needs 16 numbers");
        mat4 m;
        for (int i = 0; i < 16; ++i) m.m[i] = (float)v.array[i].number;
        return m;
    }
    static std::string arr(const mat4& m)
    {
        std::string out = "[";
        for (int i = 0; i < 16; ++i) {
            std::ostringstream o;
            o.precision(9);                  // round-trips binary32
            o << m.m[i];
            std::string t = o.str();
            // keep it a JSON *number with a fraction* (nlohmann writes floats that way): "-0" would be read back as the
            // integer 0 and lose its sign
            if (t.find_first_of(".en") == std::string::npos) t += ".0";
            out += (i ? ", " : "") + t;
        }
        return out + "]";
    }
};

// ---- OpenEXR planes ---------------------------------------------------------------------------------------------------
namespace exr {

// decoded image: float32, row 0 at the top, channels interleaved in R, G, B, A order (1 = luminance / single channel "Y" or
// "R", 3 = RGB, 4 = RGBA)
struct Image {
    int width = 0, height = 0, channels = 0;
    std::vector<float> data;
};

inline float half_to_float(uint16_t h)
{
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {
            int e = -1;
            do { ++e; man <<= 1; } while (!(man & 0x400u));
            bits = sign | (uint32_t)(127 - 15 - e) << 23 | (man & 0x3ffu) << 13;
        }
    } else if (exp == 31) bits = sign | 0x7f800000u | man << 13;
    else bits = sign | (exp + 127 - 15) << 23 | man << 13;
    float f;
    std::memcpy(&f, &bits, 4);
    return f;
}

namespace detail {
struct Reader {
    const std::vector<unsigned char>& b;
    size_t i = 0;
    template <typename T> T get()
    {
        if (sizeof(T) > b.size() || i > b.size() - sizeof(T)) throw std::runtime_error("truncated file");      // (i may come from the file: no i + n that can wrap)
        T v;
        std::memcpy(&v, b.data() + i, sizeof(T));
        i += sizeof(T);
        return v;
    }
    std::string cstr()
    {
        std::string s;
        while (i < b.size() && b[i]) s += (char)b[i++];
        if (i >= b.size()) throw std::runtime_error("truncated file");
        ++i;
        return s;
    }
};
// OpenEXR's byte predictor + interleave of the zip / rle codecs, undone
inline void unpredict(std::vector<unsigned char>& t, std::vector<unsigned char>& out)
{
    for (size_t k = 1; k < t.size(); ++k) t[k] = (unsigned char)(t[k - 1] + t[k] - 128);
    out.resize(t.size());
    const size_t half = (t.size() + 1) / 2;
    for (size_t k = 0, a = 0, c = half; k < t.size();) {
        out[k++] = t[a++];
        if (k < t.size()) out[k++] = t[c++];
    }
}
}  // namespace detail

// returns false (with a message on stderr) when the file cannot be read or uses a feature that is not supported
inline bool read(const std::string& path, Image& out)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    const std::vector<unsigned char> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    try {
        detail::Reader r{bytes};
        if (r.get<uint32_t>() != 20000630u) throw std::runtime_error("not an OpenEXR file");
        const uint32_t version = r.get<uint32_t>();
        if ((version & 0xffu) != 2 || (version & 0x1a00u)) throw std::runtime_error("tiled / deep / multi-part files are not supported");
        struct Channel { std::string name; int type; };
        std::vector<Channel> channels;
        int compression = -1, x0 = 0, y0 = 0, x1 = -1, y1 = -1, line_order = 0;
        for (;;) {
            const std::string name = r.cstr();
            if (name.empty()) break;
            const std::string type = r.cstr();
            const uint32_t size = r.get<uint32_t>();
            const size_t end = r.i + size;
            if (end > bytes.size()) throw std::runtime_error("truncated header");
            if (name == "channels") {
                for (;;) {
                    const std::string cn = r.cstr();
                    if (cn.empty()) break;
                    const int pt = r.get<int32_t>();
                    r.i += 4;                                   // pLinear + reserved
                    const int xs = r.get<int32_t>(), ys = r.get<int32_t>();
                    if (xs != 1 || ys != 1) throw std::runtime_error("subsampled channels are not supported");
                    if (pt != 1 && pt != 2) throw std::runtime_error("only HALF and FLOAT channels are supported");
                    channels.push_back({cn, pt});
                }
            } else if (name == "compression") compression = r.get<uint8_t>();
            else if (name == "dataWindow") { x0 = r.get<int32_t>(); y0 = r.get<int32_t>(); x1 = r.get<int32_t>(); y1 = r.get<int32_t>(); }
            else if (name == "lineOrder") line_order = r.get<uint8_t>();
            r.i = end;
        }
        if (channels.empty() || x1 < x0 || y1 < y0) throw std::runtime_error("incomplete header");
        if (compression < 0 || compression > 3) throw std::runtime_error("compression " + std::to_string(compression) + " is not supported (NONE, RLE, ZIPS, ZIP are)");
        (void)line_order;                                       // every block carries its y: any order is fine
        // a corrupt data window must not turn into a giant allocation: every scan line block takes 8 bytes of the offset table
        // and 8 of its own header, and no codec here expands more than deflate's 1032 : 1
        const int64_t W64 = (int64_t)x1 - x0 + 1, H64 = (int64_t)y1 - y0 + 1;
        const int lines_per_block = compression == 3 ? 16 : 1;
        if (W64 > (1 << 24) || H64 > (1 << 24) || (H64 + lines_per_block - 1) / lines_per_block * 16 > (int64_t)bytes.size() ||
            W64 * H64 * (int64_t)channels.size() * 2 > (int64_t)bytes.size() * 1100 + (1 << 20))
            throw std::runtime_error("data window does not fit the file");
        const int W = (int)W64, H = (int)H64, nblocks = (H + lines_per_block - 1) / lines_per_block;
        // where each file channel goes: R, G, B, A -> 0..3; a lone Y (or any single channel) -> 0
        std::vector<int> slot(channels.size(), -1);
        int nout = 0;
        for (size_t c = 0; c < channels.size(); ++c) {
            const std::string& n = channels[c].name;
            slot[c] = n == "R" ? 0 : n == "G" ? 1 : n == "B" ? 2 : n == "A" ? 3 : -1;
            if (slot[c] >= 0 && slot[c] + 1 > nout) nout = slot[c] + 1;
        }
        if (nout == 0) { slot[0] = 0; nout = 1; }               // "Y", "Z", ...: the first channel
        if (nout == 2) nout = 3;
        size_t bytes_per_line = 0;
        for (const Channel& c : channels) bytes_per_line += (size_t)W * (c.type == 1 ? 2 : 4);
        out.width = W; out.height = H; out.channels = nout;
        out.data.assign((size_t)W * H * nout, nout == 4 ? 1.0f : 0.0f);
        std::vector<uint64_t> offsets(nblocks);
        for (int k = 0; k < nblocks; ++k) offsets[k] = r.get<uint64_t>();
        std::vector<unsigned char> raw, tmp;
        for (int k = 0; k < nblocks; ++k) {
            r.i = (size_t)offsets[k];
            const int y = r.get<int32_t>();
            const uint32_t size = r.get<uint32_t>();
            if (r.i + size > bytes.size() || y < y0 || y > y1) throw std::runtime_error("bad scan line block");
            const int lines = std::min(lines_per_block, y1 - y + 1);
            const size_t expect = bytes_per_line * lines;
            const unsigned char* src = bytes.data() + r.i;
            if (compression == 0 || size == expect) raw.assign(src, src + size);          // a block that did not shrink is stored raw
            else if (compression == 1) {
                tmp.clear();
                for (size_t p = 0; p < size;) {
                    const int count = (signed char)src[p++];
                    if (count < 0) { if (p + (size_t)-count > size) throw std::runtime_error("bad RLE data"); tmp.insert(tmp.end(), src + p, src + p - count); p += (size_t)-count; }
                    else { if (p >= size) throw std::runtime_error("bad RLE data"); tmp.insert(tmp.end(), (size_t)count + 1, src[p++]); }
                }
                detail::unpredict(tmp, raw);
            } else {
                tmp.resize(expect);
                uLongf n = (uLongf)expect;
                if (uncompress(tmp.data(), &n, src, size) != Z_OK) throw std::runtime_error("zlib: bad ZIP block");
                tmp.resize(n);
                detail::unpredict(tmp, raw);
            }
            if (raw.size() != expect) throw std::runtime_error("scan line block has the wrong size");
            const unsigned char* p = raw.data();
            for (int l = 0; l < lines; ++l) {
                float* row = out.data.data() + (size_t)(y - y0 + l) * W * nout;
                for (size_t c = 0; c < channels.size(); ++c) {
                    const int bpp = channels[c].type == 1 ? 2 : 4;
                    if (slot[c] >= 0 && slot[c] < nout)
                        for (int x = 0; x < W; ++x) {
                            if (bpp == 2) { uint16_t h; std::memcpy(&h, p + (size_t)x * 2, 2); row[(size_t)x * nout + slot[c]] = half_to_float(h); }
                            else std::memcpy(&row[(size_t)x * nout + slot[c]], p + (size_t)x * 4, 4);
                        }
                    p += (size_t)W * bpp;
                }
            }
        }
        return true;
    } catch (const std::exception& e) {
        std::cerr << "exr: " << path << ": " << e.what() << std::endl;
        return false;
    }
}

// FLOAT channels, no compression, increasing y.  channels = 1 ("Y"), 3 (RGB) or 4 (RGBA), interleaved input.
inline bool write(const std::string& path, const float* data, int width, int height, int channels)
{
    if (width <= 0 || height <= 0 || (channels != 1 && channels != 3 && channels != 4)) return false;
    std::vector<unsigned char> b;
    auto put = [&](const void* p, size_t n) { b.insert(b.end(), (const unsigned char*)p, (const unsigned char*)p + n); };
    auto u32 = [&](uint32_t v) { put(&v, 4); };
    auto str = [&](const char* s) { put(s, std::strlen(s) + 1); };
    u32(20000630u); u32(2u);
    const char* names1[] = {"Y"}; const int order1[] = {0};
    const char* names3[] = {"B", "G", "R"}; const int order3[] = {2, 1, 0};
    const char* names4[] = {"A", "B", "G", "R"}; const int order4[] = {3, 2, 1, 0};            // chlist entries are sorted by name
    const char** names = channels == 1 ? names1 : channels == 3 ? names3 : names4;
    const int* order = channels == 1 ? order1 : channels == 3 ? order3 : order4;
    str("channels"); str("chlist");
    uint32_t chsize = 1;
    for (int c = 0; c < channels; ++c) chsize += (uint32_t)std::strlen(names[c]) + 1 + 16;
    u32(chsize);
    for (int c = 0; c < channels; ++c) { str(names[c]); u32(2); u32(0); u32(1); u32(1); }
    b.push_back(0);
    str("compression"); str("compression"); u32(1); b.push_back(0);
    const int32_t box[4] = {0, 0, width - 1, height - 1};
    str("dataWindow"); str("box2i"); u32(16); put(box, 16);
    str("displayWindow"); str("box2i"); u32(16); put(box, 16);
    str("lineOrder"); str("lineOrder"); u32(1); b.push_back(0);
    const float one = 1.f, zero2[2] = {0.f, 0.f};
    str("pixelAspectRatio"); str("float"); u32(4); put(&one, 4);
    str("screenWindowCenter"); str("v2f"); u32(8); put(zero2, 8);
    str("screenWindowWidth"); str("float"); u32(4); put(&one, 4);
    b.push_back(0);
    const size_t line_bytes = (size_t)width * channels * 4, table = b.size();
    b.resize(table + (size_t)height * 8);
    std::vector<float> line((size_t)width);
    for (int y = 0; y < height; ++y) {
        const uint64_t off = b.size();
        std::memcpy(b.data() + table + (size_t)y * 8, &off, 8);
        u32((uint32_t)y); u32((uint32_t)line_bytes);
        for (int c = 0; c < channels; ++c) {
            for (int x = 0; x < width; ++x) line[x] = data[((size_t)y * width + x) * channels + order[c]];
            put(line.data(), (size_t)width * 4);
        }
    }
    std::ofstream f(path, std::ios::binary);
    f.write((const char*)b.data(), (std::streamsize)b.size());
    return (bool)f;
}
}  // namespace exr

// ---- G-buffer -----------------------------------------------------------------------------------------------------------
namespace detail {
inline std::string frame_name(const std::string& format, int f)
{
    char buff[512];
    snprintf(buff, sizeof(buff), format.c_str(), f);     // printf-style per-frame names, RenderIO.cpp:23-24
    return buff;
}
// vec4Array2D view of whatever the file held (3-channel files get w = 1), RenderIO.cpp reads through vsg::read_cast<vec4Array2D>
inline std::vector<float> to_rgba(const exr::Image& im)
{
    std::vector<float> out((size_t)im.width * im.height * 4);
    for (size_t p = 0; p < (size_t)im.width * im.height; ++p)
        for (int c = 0; c < 4; ++c) out[p * 4 + c] = c < im.channels ? im.data[p * im.channels + c] : (c == 3 ? 1.0f : 0.0f);
    return out;
}
}  // namespace detail

// source/io/RenderIO.hpp:45-64.  Holds one frame's planes as decoded (rgba32f; depth r32f when the sequence stores depth);
// the reference's host-side conversions happen on the device in upload_to_g_buffer.
class OfflineGBuffer : public Inherit<OfflineGBuffer> {
public:
    uint32_t width = 0, height = 0;
    std::vector<float> depth;                       // [H][W]     import_g_buffer_depth
    std::vector<float> position, normal, albedo;    // [H][W][4]  world position (import_g_buffer_position), cartesian normal, float albedo
    std::optional<mat4> inv_view;                   // of the frame's (combined) matrices: the eye point for position -> depth
    bool valid() const { return width && !normal.empty() && !albedo.empty() && (!depth.empty() || !position.empty()); }

    // upload_to_g_buffer_command (RenderIO.cpp:718-746) + the import conversions, enqueued on the context's stream
    void upload_to_g_buffer(ref_ptr<GBuffer>& g_buffer, Context& context)
    {
        if (!valid()) throw std::runtime_error("OfflineGBuffer: frame was not loaded");
        if (g_buffer->width != width || g_buffer->height != height) throw std::runtime_error("OfflineGBuffer: extent differs from the GBuffer's");
        auto stage = [&](ref_ptr<DescriptorImage>& im, const std::vector<float>& plane) {
            if (plane.empty()) return ref_ptr<DescriptorImage>();
            if (!im) { im = DescriptorImage::create(context, VKPBRT_FORMAT_R32G32B32A32_SFLOAT, width, height); im->compile(context); }
            check(vkpbrt_image_upload(im->handle, plane.data(), plane.size() * sizeof(float)));
            return im;
        };
        auto p = stage(_position, position), n = stage(_normal, normal), a = stage(_albedo, albedo);
        if (!depth.empty()) check(vkpbrt_image_upload(g_buffer->depth->handle, depth.data(), depth.size() * sizeof(float)));
        if (p && !inv_view) throw std::runtime_error("OfflineGBuffer: positions need the frame's camera matrices");
        g_buffer->import_planes(p, p ? &*inv_view : nullptr, n, a);
    }
    // the export side (download_from_g_buffer_command + transfer_staging_data_to, RenderIO.cpp:748-928): the GBuffer's own
    // planes -- depth above, normals as (theta, phi), material / albedo as rgba8 -- read back once the frame's work is done
    std::vector<float> normal_spherical;            // [H][W][2]
    std::vector<uint8_t> material_unorm, albedo_unorm;   // [H][W][4]
    void download_from_g_buffer(ref_ptr<GBuffer>& g_buffer, Context& context)
    {
        width = g_buffer->width; height = g_buffer->height;
        const size_t n = (size_t)width * height;
        depth.resize(n); normal_spherical.resize(n * 2); albedo_unorm.resize(n * 4);
        check(vkpbrt_image_download(g_buffer->depth->handle, depth.data(), n * sizeof(float)));
        check(vkpbrt_image_download(g_buffer->normal->handle, normal_spherical.data(), n * 2 * sizeof(float)));
        check(vkpbrt_image_download(g_buffer->albedo->handle, albedo_unorm.data(), n * 4));
        if (g_buffer->material) { material_unorm.resize(n * 4); check(vkpbrt_image_download(g_buffer->material->handle, material_unorm.data(), n * 4)); }
        context.waitForCompletion();
    }
private:
    ref_ptr<DescriptorImage> _position, _normal, _albedo;      // staging images, reused from frame to frame
};
using OfflineGBuffers = std::vector<ref_ptr<OfflineGBuffer>>;

class GBufferIO {   // source/io/RenderIO.hpp:67-87
public:
    // RenderIO.cpp:6-69 (the reference loads no material plane either)
    static OfflineGBuffers import_g_buffer_depth(const std::string& depth_format, const std::string& normal_format, const std::string& material_format,
                                                 const std::string& albedo_format, int num_frames, int verbosity = 1)
    {
        (void)material_format;
        if (verbosity > 0) std::cout << "Start loading GBuffer" << std::endl;
        OfflineGBuffers out(num_frames);
        for (int f = 0; f < num_frames; ++f) {
            out[f] = OfflineGBuffer::create();
            exr::Image d;
            const std::string name = detail::frame_name(depth_format, f);
            if (!exr::read(name, d)) { std::cerr << "Failed to load image: " << name << std::endl; continue; }
            out[f]->width = d.width; out[f]->height = d.height;
            out[f]->depth.resize((size_t)d.width * d.height);
            for (size_t p = 0; p < out[f]->depth.size(); ++p) out[f]->depth[p] = d.data[p * d.channels];
            load_normal_albedo(*out[f], detail::frame_name(normal_format, f), detail::frame_name(albedo_format, f));
        }
        if (verbosity > 0) std::cout << "Done loading GBuffer" << std::endl;
        return out;
    }
    // RenderIO.cpp:71-158
    static OfflineGBuffers import_g_buffer_position(const std::string& position_format, const std::string& normal_format, const std::string& material_format,
                                                    const std::string& albedo_format, const std::vector<CameraMatrices>& matrices, int num_frames, int verbosity = 1)
    {
        (void)material_format;
        if (verbosity > 0) std::cout << "Start loading GBuffer" << std::endl;
        OfflineGBuffers out(num_frames);
        for (int f = 0; f < num_frames; ++f) {
            out[f] = OfflineGBuffer::create();
            exr::Image pos;
            const std::string name = detail::frame_name(position_format, f);
            if (!exr::read(name, pos)) { std::cerr << "Failed to load image: " << name << std::endl; continue; }
            if (pos.channels < 3) { std::cerr << "Unexpected position format" << std::endl; continue; }
            if ((size_t)f >= matrices.size()) { std::cerr << "No camera matrices for frame " << f << std::endl; continue; }
            out[f]->width = pos.width; out[f]->height = pos.height;
            out[f]->position = detail::to_rgba(pos);
            out[f]->inv_view = matrices[f].inv_view;
            load_normal_albedo(*out[f], detail::frame_name(normal_format, f), detail::frame_name(albedo_format, f));
        }
        if (verbosity > 0) std::cout << "Done loading GBuffer" << std::endl;
        return out;
    }
    // ---- the export side, RenderIO.cpp:213-382: host loops like the reference's (an offline tool, not on the frame path) ----
    // :312-329.  cos / sin are called unqualified on floats there: with <cmath> alone in scope those are the C library's double
    // routines and each product is rounded once -- written out here so that it does not depend on the headers in scope
    static std::vector<float> spherical_to_cartesian(const std::vector<float>& normals)
    {
        std::vector<float> out(normals.size() * 2);
        for (size_t i = 0; i < normals.size() / 2; ++i) {
            const double theta = normals[2 * i], phi = normals[2 * i + 1];
            out[4 * i] = (float)(std::cos(phi) * std::sin(theta));
            out[4 * i + 1] = (float)(std::sin(phi) * std::sin(theta));
            out[4 * i + 2] = (float)std::cos(theta);
            out[4 * i + 3] = 1.0f;
        }
        return out;
    }
    // :331-346
    static std::vector<float> unorm_to_float(const std::vector<uint8_t>& array)
    {
        std::vector<float> out(array.size());
        for (size_t i = 0; i < array.size(); ++i) out[i] = (float)array[i] / 255.0f;
        return out;
    }
    // :348-382: world position = inv_view[3] + depth * (inv_view * normalize((inv_proj * (clip, 1, 1)) with w = 0)); empty
    // without a separate projection matrix.  mat4 * vec4 and normalize as vsg writes them (sums left to right, v * (1 / length))
    static std::vector<float> depth_to_position(const std::vector<float>& depths, uint32_t w, uint32_t h, const CameraMatrices& matrix)
    {
        if (depths.empty()) return {};
        if (!matrix.proj || !matrix.inv_proj) {
            std::cout << "GBufferIO::depthToPosition: Camera matrix in wrong layout. Expected camera matrix with separate projection matrix" << std::endl;
            return {};
        }
        auto mat_vec = [](const float* m, const float* v, float* o) {
            for (int r = 0; r < 4; ++r) {
                volatile float s = m[r] * v[0];         // one rounding per operation whatever the compiler's contraction mode
                s = s + m[4 + r] * v[1]; s = s + m[8 + r] * v[2]; s = s + m[12 + r] * v[3];
                o[r] = s;
            }
        };
        const float* ip = matrix.inv_proj->m; const float* iv = matrix.inv_view.m;
        std::vector<float> out((size_t)w * h * 4);
        for (uint32_t i = 0; i < w * h; ++i) {
            const uint32_t x = i % w, y = i / w;
            const float clip[4] = {((float)x + .5f) / (float)w * 2.0f - 1.0f, ((float)y + .5f) / (float)h * 2.0f - 1.0f, 1.0f, 1.0f};
            float dir[4], world[4];
            mat_vec(ip, clip, dir);
            dir[3] = 0.0f;
            volatile float l2 = dir[0] * dir[0];
            l2 = l2 + dir[1] * dir[1]; l2 = l2 + dir[2] * dir[2]; l2 = l2 + dir[3] * dir[3];
            const float inv_len = 1.0f / std::sqrt((float)l2);
            for (float& c : dir) c *= inv_len;
            mat_vec(iv, dir, world);
            for (int c = 0; c < 3; ++c) { volatile float t = world[c] * depths[i]; out[4 * (size_t)i + c] = iv[12 + c] + t; }
            out[4 * (size_t)i + 3] = 1.0f;
        }
        return out;
    }
    // :213-310: empty format strings skip a plane; g_buffers hold what download_from_g_buffer read back
    static bool export_g_buffer(const std::string& position_format, const std::string& depth_format, const std::string& normal_format,
                                const std::string& material_format, const std::string& albedo_format, int num_frames, const OfflineGBuffers& g_buffers,
                                const std::vector<CameraMatrices>& matrices, int verbosity = 1)
    {
        if (verbosity > 0) std::cout << "Start exporting GBuffer" << std::endl;
        bool fine = true;
        for (int f = 0; f < num_frames; ++f) {
            const OfflineGBuffer* g = (size_t)f < g_buffers.size() ? g_buffers[f].get() : nullptr;
            auto store = [&](const std::string& format, const std::vector<float>& plane, int channels) {
                if (format.empty()) return true;
                const std::string name = detail::frame_name(format, f);
                if (g && !plane.empty() && plane.size() == (size_t)g->width * g->height * channels &&
                    exr::write(name, plane.data(), (int)g->width, (int)g->height, channels)) return true;
                std::cerr << "Failed to store image: " << name << std::endl;
                return false;
            };
            static const std::vector<float> none;
            const bool ok = store(depth_format, g ? g->depth : none, 1) &&
                            store(position_format, g && !position_format.empty() && (size_t)f < matrices.size() ? depth_to_position(g->depth, g->width, g->height, matrices[f]) : none, 4) &&
                            store(normal_format, g && !normal_format.empty() ? spherical_to_cartesian(g->normal_spherical) : none, 4) &&
                            store(material_format, g && !material_format.empty() ? unorm_to_float(g->material_unorm) : none, 4) &&
                            store(albedo_format, g && !albedo_format.empty() ? unorm_to_float(g->albedo_unorm) : none, 4);
            fine = fine && ok;       // like the reference, a frame stops at its first failed plane and the other frames go on
        }
        if (verbosity > 0) std::cout << "Done exporting GBuffer" << std::endl;
        return fine;
    }
private:
    static bool load_normal_albedo(OfflineGBuffer& g, const std::string& normal_path, const std::string& albedo_path)
    {
        exr::Image n, a;
        if (!exr::read(normal_path, n)) { std::cerr << "Failed to load image: " << normal_path << std::endl; return false; }
        if (!exr::read(albedo_path, a)) { std::cerr << "Failed to load image: " << albedo_path << std::endl; return false; }
        if (n.width != (int)g.width || n.height != (int)g.height || a.width != (int)g.width || a.height != (int)g.height) {
            std::cerr << "GBuffer planes of different extents: " << normal_path << std::endl;
            return false;
        }
        g.normal = detail::to_rgba(n);
        g.albedo = detail::to_rgba(a);
        return true;
    }
};

// ---- illumination -------------------------------------------------------------------------------------------------------
class OfflineIllumination : public Inherit<OfflineIllumination> {   // source/io/RenderIO.hpp:90-108
public:
    uint32_t width = 0, height = 0;
    std::vector<float> noisy;                       // [H][W][4]
    // upload_to_illumination_buffer_command (RenderIO.cpp:384-399): image 0 of an IlluminationBufferDemodulatedFloat
    void upload_to_illumination_buffer(ref_ptr<IlluminationBuffer>& illu_buffer, Context&)
    {
        if (noisy.empty()) throw std::runtime_error("OfflineIllumination: frame was not loaded");
        const vkpbrt_image_info i = illu_buffer->illumination_images.at(0)->info();
        if (i.format != VKPBRT_FORMAT_R32G32B32A32_SFLOAT || i.width != width || i.height != height)
            throw std::runtime_error("OfflineIllumination: the illumination buffer must be rgba32f of the sequence's extent");
        check(vkpbrt_image_upload(illu_buffer->illumination_images[0]->handle, noisy.data(), noisy.size() * sizeof(float)));
    }
    // download_from_illumination_buffer_command + transfer_staging_data_to (RenderIO.cpp:401-467): image 0, rgba32f
    void download_from_illumination_buffer(ref_ptr<IlluminationBuffer>& illu_buffer, Context& context)
    {
        const vkpbrt_image_info i = illu_buffer->illumination_images.at(0)->info();
        if (i.format != VKPBRT_FORMAT_R32G32B32A32_SFLOAT) throw std::runtime_error("OfflineIllumination: the illumination buffer must be rgba32f");
        width = i.width; height = i.height;
        noisy.resize((size_t)width * height * 4);
        check(vkpbrt_image_download(illu_buffer->illumination_images[0]->handle, noisy.data(), noisy.size() * sizeof(float)));
        context.waitForCompletion();
    }
};
using OfflineIlluminations = std::vector<ref_ptr<OfflineIllumination>>;

class IlluminationBufferIO {   // source/io/RenderIO.hpp:111-118
public:
    // RenderIO.cpp:502-547
    static OfflineIlluminations import_illumination(const std::string& illumination_format, int num_frames, int verbosity = 1)
    {
        if (verbosity > 0) std::cout << "Start loading Illumination" << std::endl;
        OfflineIlluminations out(num_frames);
        for (int f = 0; f < num_frames; ++f) {
            out[f] = OfflineIllumination::create();
            exr::Image im;
            const std::string name = detail::frame_name(illumination_format, f);
            if (!exr::read(name, im)) { std::cerr << "Failed to load image: " << name << std::endl; continue; }
            out[f]->width = im.width; out[f]->height = im.height;
            out[f]->noisy = detail::to_rgba(im);
        }
        if (verbosity > 0) std::cout << "Done loading Illumination" << std::endl;
        return out;
    }
    // RenderIO.cpp:549-591
    static bool export_illumination(const std::string& illumination_format, int num_frames, const OfflineIlluminations& illus, int verbosity = 1)
    {
        (void)verbosity;
        bool fine = true;
        for (int f = 0; f < num_frames; ++f) {
            const std::string name = detail::frame_name(illumination_format, f);
            if (!illus[f] || illus[f]->noisy.empty() || !exr::write(name, illus[f]->noisy.data(), (int)illus[f]->width, (int)illus[f]->height, 4)) {
                std::cerr << "Faled to store image: " << name << std::endl;
                fine = false;
            }
        }
        return fine;
    }
};

}  // namespace vkpbrt
