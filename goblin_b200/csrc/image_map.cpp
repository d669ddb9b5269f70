#include "image_map.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include <zlib.h>

namespace gb {
namespace {

float halfToFloat(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu;
    uint32_t man = h & 0x3FFu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else { // subnormal half -> normal float
            int e = -1;
            do { ++e; man <<= 1; } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
    else bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    float f;
    std::memcpy(&f, &bits, 4);
    return f;
}

struct Reader {
    const std::vector<uint8_t>& b;
    size_t pos = 0;
    bool ok = true;
    bool need(size_t n) { if (pos + n > b.size()) ok = false; return ok; }
    uint8_t u8() { if (!need(1)) return 0; return b[pos++]; }
    int32_t i32() { int32_t v = 0; if (need(4)) { std::memcpy(&v, &b[pos], 4); pos += 4; } return v; }
    uint64_t u64() { uint64_t v = 0; if (need(8)) { std::memcpy(&v, &b[pos], 8); pos += 8; } return v; }
    std::string str() {
        std::string s;
        while (need(1) && b[pos] != 0) s.push_back((char)b[pos++]);
        if (ok) ++pos;
        return s;
    }
};

struct Channel { std::string name; int type = 0; };

// the byte shuffling OpenEXR applies around ZIP / RLE: delta predictor, then the two interleaved halves
void unpredict(std::vector<uint8_t>& t, std::vector<uint8_t>& out) {
    for (size_t i = 1; i < t.size(); ++i) t[i] = (uint8_t)(t[i - 1] + t[i] - 128);
    out.resize(t.size());
    const size_t half = (t.size() + 1) / 2;
    size_t a = 0, b2 = half;
    for (size_t i = 0; i < t.size();) {
        out[i++] = t[a++];
        if (i < t.size()) out[i++] = t[b2++];
    }
}

bool rleDecode(const uint8_t* in, size_t n, std::vector<uint8_t>& out, size_t expected) {
    out.clear();
    size_t i = 0;
    while (i < n) {
        int8_t c = (int8_t)in[i++];
        if (c < 0) {
            size_t cnt = (size_t)(-(int)c);
            if (i + cnt > n) return false;
            out.insert(out.end(), in + i, in + i + cnt);
            i += cnt;
        } else {
            if (i >= n) return false;
            out.insert(out.end(), (size_t)c + 1, in[i++]);
        }
        if (out.size() > expected) return false;
    }
    return out.size() == expected;
}

} // namespace

bool loadEXR(const std::string& path, int* width, int* height, std::vector<float>* rgba, std::string* error) {
    auto fail = [&](const std::string& m) { if (error) *error = "unable to read image " + path + " : " + m; return false; };
    std::vector<uint8_t> bytes;
    {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) return fail("cannot open file");
        fseek(f, 0, SEEK_END);
        long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        bytes.resize(n > 0 ? (size_t)n : 0);
        size_t got = bytes.empty() ? 0 : fread(bytes.data(), 1, bytes.size(), f);
        fclose(f);
        if (got != bytes.size()) return fail("short read");
    }
    Reader r{bytes};
    if (bytes.size() < 8 || r.i32() != 20000630) return fail("not an OpenEXR file");
    const int32_t version = r.i32();
    if ((version & 0xFF) != 2 || (version & 0x1A00)) return fail("tiled, deep and multi-part files are not supported");
    std::vector<Channel> channels;
    int compression = -1, x0 = 0, y0 = 0, x1 = -1, y1 = -1;
    while (r.ok) {
        std::string name = r.str();
        if (name.empty()) break;
        std::string type = r.str();
        int32_t size = r.i32();
        if (size < 0 || !r.need((size_t)size)) return fail("truncated header");
        const size_t end = r.pos + (size_t)size;
        if (name == "channels") {
            while (r.pos < end && bytes[r.pos] != 0) {
                Channel c;
                c.name = r.str();
                c.type = r.i32();
                r.pos += 4; // pLinear + reserved
                int32_t xs = r.i32(), ys = r.i32();
                if (xs != 1 || ys != 1) return fail("subsampled channels are not supported");
                if (c.type < 0 || c.type > 2) return fail("unknown pixel type");
                channels.push_back(c);
            }
        } else if (name == "compression") compression = r.u8();
        else if (name == "dataWindow") { x0 = r.i32(); y0 = r.i32(); x1 = r.i32(); y1 = r.i32(); }
        r.pos = end;
    }
    if (!r.ok || channels.empty() || compression < 0 || x1 < x0 || y1 < y0) return fail("malformed header");
    if (compression > 3) return fail("only none / RLE / ZIPS / ZIP compression is supported");
    const int w = x1 - x0 + 1, h = y1 - y0 + 1;
    if ((int64_t)w * h > (1ll << 28)) return fail("image too large");
    const int linesPerBlock = compression == 3 ? 16 : 1;
    const int nBlocks = (h + linesPerBlock - 1) / linesPerBlock;
    std::vector<uint64_t> offsets(nBlocks);
    for (int i = 0; i < nBlocks; ++i) offsets[i] = r.u64();
    if (!r.ok) return fail("truncated offset table");
    size_t lineBytes = 0;
    for (const Channel& c : channels) lineBytes += (size_t)w * (c.type == 1 ? 2 : 4);
    // channel -> rgba slot, like LoadEXR: by name, one channel broadcast
    int slot[4] = {-1, -1, -1, -1};
    for (size_t c = 0; c < channels.size(); ++c) {
        if (channels[c].name == "R") slot[0] = (int)c;
        else if (channels[c].name == "G") slot[1] = (int)c;
        else if (channels[c].name == "B") slot[2] = (int)c;
        else if (channels[c].name == "A") slot[3] = (int)c;
    }
    const bool single = channels.size() == 1;
    if (!single && (slot[0] < 0 || slot[1] < 0 || slot[2] < 0)) return fail("R, G or B channel not found");
    rgba->assign((size_t)w * h * 4, 1.0f);
    std::vector<uint8_t> raw, tmp;
    for (int blk = 0; blk < nBlocks; ++blk) {
        Reader c{bytes};
        c.pos = (size_t)offsets[blk];
        const int32_t y = c.i32();
        const int32_t dataSize = c.i32();
        if (!c.ok || dataSize < 0 || !c.need((size_t)dataSize)) return fail("truncated pixel data");
        const int lines = std::min(linesPerBlock, y1 - y + 1);
        if (y < y0 || lines <= 0) return fail("scanline outside the data window");
        const size_t expected = lineBytes * (size_t)lines;
        const uint8_t* src = &bytes[c.pos];
        if (compression == 0 || (size_t)dataSize == expected) {
            if ((size_t)dataSize < expected) return fail("short scanline block");
            raw.assign(src, src + expected);
        } else if (compression == 1) {
            if (!rleDecode(src, (size_t)dataSize, tmp, expected)) return fail("bad RLE data");
            unpredict(tmp, raw);
        } else {
            tmp.resize(expected);
            uLongf n = (uLongf)expected;
            if (uncompress(tmp.data(), &n, src, (uLong)dataSize) != Z_OK || n != expected) return fail("bad ZIP data");
            unpredict(tmp, raw);
        }
        for (int l = 0; l < lines; ++l) {
            const uint8_t* line = raw.data() + lineBytes * (size_t)l;
            float* outRow = rgba->data() + 4 * (size_t)w * (size_t)(y - y0 + l);
            size_t off = 0;
            for (size_t ch = 0; ch < channels.size(); ++ch) {
                const int type = channels[ch].type;
                const size_t esz = type == 1 ? 2 : 4;
                for (int k = 0; k < 4; ++k) {
                    if (!(single || slot[k] == (int)ch)) continue;
                    for (int x = 0; x < w; ++x) {
                        float v;
                        if (type == 1) { uint16_t hv; std::memcpy(&hv, line + off + 2 * (size_t)x, 2); v = halfToFloat(hv); }
                        else if (type == 2) std::memcpy(&v, line + off + 4 * (size_t)x, 4);
                        else { uint32_t uv; std::memcpy(&uv, line + off + 4 * (size_t)x, 4); v = (float)uv; }
                        outRow[4 * x + k] = v;
                    }
                }
                off += esz * (size_t)w;
            }
        }
    }
    *width = w;
    *height = h;
    return true;
}

namespace {
// gaussian(x, w, falloff = 2), src/GoblinTexture.cpp:516-518
float gaussianWeight(float x, float w) {
    const float falloff = 2.0f;
    return std::max(0.0f, expf(-falloff * x * x) - expf(-falloff * w * w));
}
int floorInt(float f) { return (int)floor(f); }
int clampInt(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

void axisWeights(int src, int dst, float filterWidth, int nSamples, std::vector<float>* weight, std::vector<int>* index) {
    weight->assign((size_t)dst * nSamples, 0.0f);
    index->assign(dst, 0);
    for (int s = 0; s < dst; ++s) {
        float center = ((float)s + 0.5f) / dst * src;
        (*index)[s] = floorInt(center - filterWidth + 0.5f);
        float weightSum = 0.0f;
        const int wOffset = s * nSamples;
        for (int i = 0; i < nSamples; ++i) {
            float p = (*index)[s] + 0.5f + i;
            (*weight)[wOffset + i] = gaussianWeight(p - center, filterWidth);
            weightSum += (*weight)[wOffset + i];
        }
        float invW = 1.0f / weightSum;
        for (int i = 0; i < nSamples; ++i) (*weight)[wOffset + i] *= invW;
    }
}
} // namespace

void resizeImage(const float* src, int srcWidth, int srcHeight, int dstWidth, int dstHeight, std::vector<float>* dst) {
    float filterWidth = floor(std::max(2.0f, std::max((float)srcWidth / (float)dstWidth, (float)srcHeight / (float)dstHeight)));
    int nSamples = floorInt(filterWidth) * 2;
    std::vector<float> sW, tW;
    std::vector<int> sI, tI;
    axisWeights(srcWidth, dstWidth, filterWidth, nSamples, &sW, &sI);
    axisWeights(srcHeight, dstHeight, filterWidth, nSamples, &tW, &tI);
    dst->assign((size_t)dstWidth * dstHeight * 4, 0.0f);
    for (int t = 0; t < dstHeight; ++t) {
        for (int s = 0; s < dstWidth; ++s) {
            float* o = dst->data() + 4 * ((size_t)t * dstWidth + s);
            o[0] = o[1] = o[2] = 0.0f;
            o[3] = 1.0f; // Color(0.0f): alpha 1, and += leaves it alone
            for (int i = 0; i < nSamples; ++i) {
                int srcT = clampInt(tI[t] + i, 0, srcHeight - 1);
                for (int j = 0; j < nSamples; ++j) {
                    int srcS = clampInt(sI[s] + j, 0, srcWidth - 1);
                    float w = tW[t * nSamples + i] * sW[s * nSamples + j];
                    const float* c = src + 4 * ((size_t)srcT * srcWidth + srcS);
                    o[0] += c[0] * w;
                    o[1] += c[1] * w;
                    o[2] += c[2] * w;
                }
            }
        }
    }
}

void buildMipmap(std::vector<float> rgba, int width, int height, std::vector<MipLevel>* levels) {
    auto isPow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    auto roundUpPow2 = [](int v) { int p = 1; while (p < v) p <<= 1; return p; };
    levels->clear();
    if (!isPow2(width) || !isPow2(height)) {
        const int wP = roundUpPow2(width), hP = roundUpPow2(height);
        std::vector<float> resized;
        resizeImage(rgba.data(), width, height, wP, hP, &resized);
        rgba.swap(resized);
        width = wP;
        height = hP;
    }
    MipLevel l0;
    l0.width = width; l0.height = height; l0.rgba = std::move(rgba);
    levels->push_back(std::move(l0));
    const int levelsNum = floorInt(std::max(log2((float)width), log2((float)height))) + 1;
    for (int i = 1; i < levelsNum; ++i) {
        const MipLevel& prev = (*levels)[i - 1];
        MipLevel l;
        l.width = std::max(1, prev.width >> 1);
        l.height = std::max(1, prev.height >> 1);
        resizeImage(prev.rgba.data(), prev.width, prev.height, l.width, l.height, &l.rgba);
        levels->push_back(std::move(l));
    }
}

void mipLookup(const std::vector<MipLevel>& levels, int level, float s, float t, float out[4]) {
    level = clampInt(level, 0, (int)levels.size() - 1);
    const MipLevel& im = levels[level];
    float sRes = s * im.width - 0.5f;
    float tRes = t * im.height - 0.5f;
    int s0 = floorInt(sRes);
    float ds = sRes - (float)s0;
    int t0 = floorInt(tRes);
    float dt = tRes - (float)t0;
    auto texel = [&](int ss, int tt) { // AddressRepeat
        ss = ss % im.width;
        tt = tt % im.height;
        if (ss < 0) ss += im.width;
        if (tt < 0) tt += im.height;
        return im.rgba.data() + 4 * ((size_t)tt * im.width + ss);
    };
    const float* a = texel(s0, t0);
    const float* b = texel(s0 + 1, t0);
    const float* c = texel(s0, t0 + 1);
    const float* d = texel(s0 + 1, t0 + 1);
    const float wa = (1.0f - ds) * (1.0f - dt), wb = ds * (1.0f - dt), wc = (1.0f - ds) * dt, wd = ds * dt;
    for (int k = 0; k < 3; ++k) out[k] = ((a[k] * wa + b[k] * wb) + c[k] * wc) + d[k] * wd;
    out[3] = a[3]; // Color::operator+ keeps the left operand's alpha
}

namespace {
// CDF1D::init, src/GoblinSampler.cpp:317-331
float cdf1D(const float* f, int n, float* cdf) {
    const float dx = 1.0f / n;
    cdf[0] = 0.0f;
    for (int i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + (f[i - 1] * dx);
    const float integral = cdf[n];
    for (int i = 1; i < n + 1; ++i) cdf[i] /= integral;
    return integral;
}
} // namespace

void buildDistribution2D(const float* f2D, int width, int height, std::vector<float>* out) {
    const size_t rowsF = 0, rowsC = (size_t)height * width, margF = rowsC + (size_t)height * (width + 1);
    const size_t margC = margF + height, margI = margC + height + 1;
    out->assign(margI + 1, 0.0f);
    float* o = out->data();
    std::memcpy(o + rowsF, f2D, sizeof(float) * (size_t)width * height);
    for (int i = 0; i < height; ++i) {
        o[margF + i] = cdf1D(f2D + (size_t)i * width, width, o + rowsC + (size_t)i * (width + 1));
    }
    o[margI] = cdf1D(o + margF, height, o + margC);
}

} // namespace gb
