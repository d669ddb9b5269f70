#include "image_io.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

namespace gb {
namespace {

uint16_t floatToHalf(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    int32_t exp = (int32_t)((x >> 23) & 0xFF) - 127 + 15;
    uint32_t man = x & 0x7FFFFFu;
    if (((x >> 23) & 0xFF) == 0xFF) { // inf / nan
        return (uint16_t)(sign | 0x7C00u | (man ? 0x200u : 0u));
    }
    if (exp >= 31) return (uint16_t)(sign | 0x7C00u); // overflow -> inf
    if (exp <= 0) {                                    // subnormal half or zero
        if (exp < -10) return (uint16_t)sign;
        man |= 0x800000u;
        uint32_t shift = (uint32_t)(14 - exp);
        uint32_t half = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1), mid = 1u << (shift - 1);
        if (rem > mid || (rem == mid && (half & 1))) ++half;
        return (uint16_t)(sign | half);
    }
    uint32_t half = ((uint32_t)exp << 10) | (man >> 13);
    uint32_t rem = man & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) ++half; // may carry into the exponent: correct
    return (uint16_t)(sign | half);
}

void put(std::vector<uint8_t>& b, const void* p, size_t n) {
    const uint8_t* c = (const uint8_t*)p;
    b.insert(b.end(), c, c + n);
}
void putStr(std::vector<uint8_t>& b, const char* s) { put(b, s, strlen(s) + 1); }
void putI32(std::vector<uint8_t>& b, int32_t v) { put(b, &v, 4); }
void putF32(std::vector<uint8_t>& b, float v) { put(b, &v, 4); }
void attr(std::vector<uint8_t>& b, const char* name, const char* type, const std::vector<uint8_t>& val) {
    putStr(b, name);
    putStr(b, type);
    putI32(b, (int32_t)val.size());
    put(b, val.data(), val.size());
}

bool writeEXR(const std::string& path, const std::vector<float>& rgb, int w, int h, std::string* error) {
    std::vector<uint8_t> hd;
    const uint32_t magic = 20000630u, version = 2u;
    put(hd, &magic, 4);
    put(hd, &version, 4);
    std::vector<uint8_t> ch;
    for (const char* name : {"B", "G", "R"}) { // alphabetical, as the format requires
        putStr(ch, name);
        putI32(ch, 1); // HALF
        ch.push_back(0); ch.push_back(0); ch.push_back(0); ch.push_back(0); // pLinear + reserved
        putI32(ch, 1);
        putI32(ch, 1);
    }
    ch.push_back(0);
    attr(hd, "channels", "chlist", ch);
    attr(hd, "compression", "compression", {0});
    std::vector<uint8_t> win;
    putI32(win, 0); putI32(win, 0); putI32(win, w - 1); putI32(win, h - 1);
    attr(hd, "dataWindow", "box2i", win);
    attr(hd, "displayWindow", "box2i", win);
    attr(hd, "lineOrder", "lineOrder", {0});
    std::vector<uint8_t> one; putF32(one, 1.0f);
    attr(hd, "pixelAspectRatio", "float", one);
    std::vector<uint8_t> ctr; putF32(ctr, 0.0f); putF32(ctr, 0.0f);
    attr(hd, "screenWindowCenter", "v2f", ctr);
    attr(hd, "screenWindowWidth", "float", one);
    hd.push_back(0);

    const size_t rowBytes = (size_t)w * 3 * 2;
    uint64_t offset = hd.size() + (uint64_t)h * 8;
    std::vector<uint8_t> table;
    for (int y = 0; y < h; ++y) {
        put(table, &offset, 8);
        offset += 8 + rowBytes;
    }
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) {
        if (error) *error = "can not open file " + path;
        return false;
    }
    fwrite(hd.data(), 1, hd.size(), f);
    fwrite(table.data(), 1, table.size(), f);
    std::vector<uint16_t> row((size_t)w * 3);
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            const float* p = &rgb[3 * ((size_t)y * w + x)];
            row[x] = floatToHalf(p[2]);                 // B
            row[(size_t)w + x] = floatToHalf(p[1]);     // G
            row[2 * (size_t)w + x] = floatToHalf(p[0]); // R
        }
        int32_t yy = y, sz = (int32_t)rowBytes;
        fwrite(&yy, 4, 1, f);
        fwrite(&sz, 4, 1, f);
        fwrite(row.data(), 1, rowBytes, f);
    }
    fclose(f);
    return true;
}

bool writePFM(const std::string& path, const std::vector<float>& rgb, int w, int h, std::string* error) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) {
        if (error) *error = "can not open file " + path;
        return false;
    }
    fprintf(f, "PF\n%d %d\n-1.0\n", w, h);
    for (int y = h - 1; y >= 0; --y) fwrite(&rgb[3 * (size_t)y * w], 4, (size_t)w * 3, f);
    fclose(f);
    return true;
}

bool writePPM(const std::string& path, const std::vector<float>& rgb, int w, int h, std::string* error) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) {
        if (error) *error = "can not open file " + path;
        return false;
    }
    fprintf(f, "P3\n%d %d\n%d\n", w, h, 255);
    const float invGamma = 1.0f / 2.2f;
    auto q = [&](float c) {
        c = (float)pow(c, invGamma);
        c = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
        return static_cast<int>(c * 255.0f);
    };
    for (size_t i = 0; i < (size_t)w * h; ++i) fprintf(f, "%d %d %d ", q(rgb[3 * i]), q(rgb[3 * i + 1]), q(rgb[3 * i + 2]));
    fclose(f);
    return true;
}


// Goblin::toneMapping (src/GoblinImageIO.cpp:220-237): Reinhard's global operator around the
// log-average "world adaptation" luminance.  Serial float sums and the host libm, like the
// reference, so the result is the same bit for bit.
void toneMap(std::vector<float>& rgb, int w, int h) {
    auto lum = [&](size_t i) { return 0.212671f * rgb[3 * i] + 0.715160f * rgb[3 * i + 1] + 0.072169f * rgb[3 * i + 2]; };
    float Ywa = 0.0f;
    for (size_t i = 0; i < (size_t)w * h; ++i) Ywa += logf(1e4f + lum(i));
    Ywa = expf(Ywa / (w * h));
    const float invy2 = 1.0f / (Ywa * Ywa);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        const float y = lum(i);
        const float s = (1.0f + y * invy2) / (1.0f + y);
        rgb[3 * i] *= s;
        rgb[3 * i + 1] *= s;
        rgb[3 * i + 2] *= s;
    }
}

} // namespace

int bloomFilterWidth(float bloomRadius, int xres, int yres) {
    // ceilInt(bloomRadius * std::max(width, height)) / 2, integer division
    return (int)ceilf(bloomRadius * (float)(xres > yres ? xres : yres)) / 2;
}

void bloomFilterTable(int filterWidth, float* table) {
    for (int y = 0; y < filterWidth; ++y) {
        for (int x = 0; x < filterWidth; ++x) {
            float d = sqrtf((float)(x * x + y * y)) / (float)filterWidth;
            table[y * filterWidth + x] = powf(std::max(0.0f, 1.0f - d), 4.0f);
        }
    }
}

bool writeRgb(const std::string& path, const float* rgbIn, int xres, int yres, bool toneMapping, std::string* error) {
    std::vector<float> rgb(rgbIn, rgbIn + (size_t)xres * yres * 3);
    size_t dot = path.rfind('.');
    std::string ext = dot == std::string::npos ? "" : path.substr(dot);
    if (ext == ".exr" || ext == ".EXR") return writeEXR(path, rgb, xres, yres, error);
    if (ext == ".pfm" || ext == ".PFM") return writePFM(path, rgb, xres, yres, error);
    if (ext == ".ppm" || ext == ".PPM") {
        if (toneMapping) toneMap(rgb, xres, yres);
        return writePPM(path, rgb, xres, yres, error);
    }
    return writePPM(path + ".ppm", rgb, xres, yres, error); // the reference's fallback (no tone mapping there)
}

bool writeFilm(const std::string& path, const float* rgbw, int xres, int yres, std::string* error) {
    std::vector<float> rgb((size_t)xres * yres * 3);
    for (size_t i = 0; i < (size_t)xres * yres; ++i) {
        float inv = 1.0f / rgbw[4 * i + 3]; // Color::operator/(float); 0 * inf = NaN outside the crop window
        rgb[3 * i] = rgbw[4 * i] * inv;
        rgb[3 * i + 1] = rgbw[4 * i + 1] * inv;
        rgb[3 * i + 2] = rgbw[4 * i + 2] * inv;
    }
    return writeRgb(path, rgb.data(), xres, yres, false, error);
}

} // namespace gb
