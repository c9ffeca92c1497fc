// prv_io.hpp -- the file formats of the PRV_simulation output contract, written without OpenCV / JsonCpp / PCL:
//   * RGBA 8-bit PNG (zlib deflate + CRC), the `rgbaClip_<i>.png` / `rgb_<i>.png` files (main.cpp:87,1617 of the reference)
//   * the instant-ngp transforms JSON `<N>.json` (main.cpp:1584-1651), laid out like Json::StyledWriter
//     (keys in alphabetical order as Json::Value stores them, 3-space indent, %.17g doubles)
//   * vertex-only PLY reader (x y z red green blue; ascii or binary_little_endian), the ShapeNet clouds of main.cpp:647
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace prv {

// ---------------------------------------------------------------- PNG
inline void png_chunk(std::vector<uint8_t>& out, const char type[4], const uint8_t* data, size_t len) {
    auto be32 = [&](uint32_t v) {
        out.push_back((uint8_t)(v >> 24));
        out.push_back((uint8_t)(v >> 16));
        out.push_back((uint8_t)(v >> 8));
        out.push_back((uint8_t)v);
    };
    be32((uint32_t)len);
    const size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    if (len) out.insert(out.end(), data, data + len);
    be32((uint32_t)crc32(0L, out.data() + start, (uInt)(out.size() - start)));
}

// channels: 3 (RGB) or 4 (RGBA); rows top to bottom
inline bool write_png(const std::string& path, const uint8_t* pixels, int width, int height, int channels) {
    if (width <= 0 || height <= 0 || (channels != 3 && channels != 4)) return false;
    std::vector<uint8_t> raw((size_t)height * ((size_t)width * channels + 1));
    for (int y = 0; y < height; y++) {
        uint8_t* row = raw.data() + (size_t)y * ((size_t)width * channels + 1);
        row[0] = 0;  // filter: none
        std::memcpy(row + 1, pixels + (size_t)y * width * channels, (size_t)width * channels);
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 3) != Z_OK) return false;
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    uint8_t ihdr[13] = {(uint8_t)(width >> 24), (uint8_t)(width >> 16), (uint8_t)(width >> 8), (uint8_t)width,
                        (uint8_t)(height >> 24), (uint8_t)(height >> 16), (uint8_t)(height >> 8), (uint8_t)height,
                        8, (uint8_t)(channels == 4 ? 6 : 2), 0, 0, 0};
    png_chunk(out, "IHDR", ihdr, 13);
    png_chunk(out, "IDAT", z.data(), zlen);
    png_chunk(out, "IEND", nullptr, 0);
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok;
}

// ---------------------------------------------------------------- JSON (StyledWriter look-alike, just what <N>.json needs)
struct JsonFrame {
    std::string file_path;
    double transform[4][4];
};

inline std::string json_number(double v) {
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.17g", v);
    std::string s(buf);
    if (s.find_first_of(".eEn") == std::string::npos) s += ".0";  // JsonCpp keeps reals looking like reals
    return s;
}

inline bool write_transforms_json(const std::string& path, const std::map<std::string, double>& reals, const std::map<std::string, long long>& ints,
                                  const double offset[3], const std::vector<JsonFrame>& frames) {
    // Json::Value is a std::map: members come out in alphabetical order
    std::map<std::string, std::string> members;
    for (const auto& kv : reals) members[kv.first] = json_number(kv.second);
    for (const auto& kv : ints) members[kv.first] = std::to_string(kv.second);
    {
        std::ostringstream o;
        o << "[ " << json_number(offset[0]) << ", " << json_number(offset[1]) << ", " << json_number(offset[2]) << " ]";
        members["offset"] = o.str();
    }
    {
        std::ostringstream o;
        o << "[\n";
        for (size_t i = 0; i < frames.size(); i++) {
            o << "      {\n         \"file_path\" : \"" << frames[i].file_path << "\",\n         \"transform_matrix\" : [\n";
            for (int r = 0; r < 4; r++) {
                o << "            [ ";
                for (int c = 0; c < 4; c++) o << json_number(frames[i].transform[r][c]) << (c < 3 ? ", " : " ]");
                o << (r < 3 ? ",\n" : "\n");
            }
            o << "         ]\n      }" << (i + 1 < frames.size() ? ",\n" : "\n");
        }
        o << "   ]";
        members["frames"] = o.str();
    }
    std::ofstream f(path);
    if (!f.is_open()) return false;
    f << "{\n";
    size_t k = 0;
    for (const auto& kv : members) f << "   \"" << kv.first << "\" : " << kv.second << (++k < members.size() ? ",\n" : "\n");
    f << "}\n";
    return f.good();
}

// ---------------------------------------------------------------- PLY (vertex element only)
inline bool read_ply_xyzrgb(const std::string& path, std::vector<float>& xyz, std::vector<uint8_t>& rgb) {
    std::ifstream f(path, std::ios::binary);
    if (!f.is_open()) return false;
    std::string line;
    bool binary = false, in_vertex = false;
    size_t nv = 0;
    struct Prop {
        std::string type, name;
    };
    std::vector<Prop> props;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ls(line);
        std::string w;
        ls >> w;
        if (w == "format") {
            std::string fmt;
            ls >> fmt;
            binary = fmt == "binary_little_endian";
            if (fmt == "binary_big_endian") return false;
        } else if (w == "element") {
            std::string name;
            size_t n;
            ls >> name >> n;
            in_vertex = name == "vertex";
            if (in_vertex) nv = n;
        } else if (w == "property" && in_vertex) {
            Prop p;
            ls >> p.type;
            if (p.type == "list") return false;
            ls >> p.name;
            props.push_back(p);
        } else if (w == "end_header") {
            break;
        }
    }
    auto size_of = [](const std::string& t) -> int {
        if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
        if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
        if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
        if (t == "double" || t == "float64") return 8;
        return 0;
    };
    xyz.assign(nv * 3, 0.0f);
    rgb.assign(nv * 3, 0);
    auto store = [&](size_t i, const std::string& name, double v) {
        if (name == "x") xyz[3 * i] = (float)v;
        else if (name == "y") xyz[3 * i + 1] = (float)v;
        else if (name == "z") xyz[3 * i + 2] = (float)v;
        else if (name == "red" || name == "r") rgb[3 * i] = (uint8_t)v;
        else if (name == "green" || name == "g") rgb[3 * i + 1] = (uint8_t)v;
        else if (name == "blue" || name == "b") rgb[3 * i + 2] = (uint8_t)v;
    };
    for (size_t i = 0; i < nv; i++) {
        if (!binary) {
            if (!std::getline(f, line)) return false;
            std::istringstream ls(line);
            for (const Prop& p : props) {
                double v = 0;
                ls >> v;
                store(i, p.name, v);
            }
        } else {
            for (const Prop& p : props) {
                unsigned char buf[8] = {0};
                const int n = size_of(p.type);
                if (n == 0 || !f.read(reinterpret_cast<char*>(buf), n)) return false;
                double v = 0;
                if (p.type == "float" || p.type == "float32") { float t; std::memcpy(&t, buf, 4); v = t; }
                else if (p.type == "double" || p.type == "float64") { std::memcpy(&v, buf, 8); }
                else if (n == 1) v = (p.type == "char" || p.type == "int8") ? (double)(signed char)buf[0] : (double)buf[0];
                else if (n == 2) { uint16_t t; std::memcpy(&t, buf, 2); v = (p.type == "short" || p.type == "int16") ? (double)(int16_t)t : (double)t; }
                else { uint32_t t; std::memcpy(&t, buf, 4); v = (p.type == "int" || p.type == "int32") ? (double)(int32_t)t : (double)t; }
                store(i, p.name, v);
            }
        }
    }
    return true;
}

}  // namespace prv
