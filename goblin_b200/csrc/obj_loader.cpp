#include "obj_loader.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <string>
#include <thread>

namespace gb {
namespace {

struct Corner { int v, n, t; };

inline bool isWs(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// Parses `count` floats from a NUL-terminated line the way `stream >> float`
// would: leading whitespace skipped, decimal literal, failure on anything else.
bool parseFloats(const char* s, int count, float* out) {
    for (int i = 0; i < count; ++i) {
        while (isWs(*s)) ++s;
        if (*s == '\0') return false;
        // iostreams do not accept inf / nan / hex floats
        const char* c = s;
        if (*c == '+' || *c == '-') ++c;
        if (!((*c >= '0' && *c <= '9') || *c == '.')) return false;
        char* e = nullptr;
        out[i] = strtof(s, &e);
        if (e == s) return false;
        s = e;
    }
    return true;
}

// open-addressing map (v, n, t) -> output vertex index, first insertion wins
struct CornerMap {
    std::vector<Corner> keys;
    std::vector<uint32_t> vals;
    size_t mask = 0, used = 0;
    explicit CornerMap(size_t expected) {
        size_t cap = 64;
        while (cap < expected * 2) cap <<= 1;
        keys.assign(cap, Corner{-2, -2, -2});
        vals.assign(cap, 0);
        mask = cap - 1;
    }
    static uint64_t hash(const Corner& c) {
        uint64_t h = (uint64_t)(uint32_t)c.v * 0x9E3779B97F4A7C15ull;
        h ^= ((uint64_t)(uint32_t)c.n + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
        h ^= ((uint64_t)(uint32_t)c.t + 0x165667B1ull) * 0x85EBCA77C2B2AE63ull;
        return h ^ (h >> 29);
    }
    void grow() {
        std::vector<Corner> ok;
        std::vector<uint32_t> ov;
        ok.swap(keys);
        ov.swap(vals);
        size_t cap = ok.size() * 2;
        keys.assign(cap, Corner{-2, -2, -2});
        vals.assign(cap, 0);
        mask = cap - 1;
        for (size_t i = 0; i < ok.size(); ++i) {
            if (ok[i].v == -2) continue;
            size_t h = hash(ok[i]) & mask;
            while (keys[h].v != -2) h = (h + 1) & mask;
            keys[h] = ok[i];
            vals[h] = ov[i];
        }
    }
    // returns true when newly inserted; *index receives the stored value
    bool insert(const Corner& c, uint32_t value, uint32_t* index) {
        if ((used + 1) * 2 > keys.size()) grow();
        size_t h = hash(c) & mask;
        while (keys[h].v != -2) {
            if (keys[h].v == c.v && keys[h].n == c.n && keys[h].t == c.t) {
                *index = vals[h];
                return false;
            }
            h = (h + 1) & mask;
        }
        keys[h] = c;
        vals[h] = value;
        ++used;
        *index = value;
        return true;
    }
};

enum Format { kV, kVT, kVN, kVTN };

// Runs body(t) for t in [0, n) on up to n threads.
template <typename F>
void runParallel(unsigned n, F body) {
    if (n <= 1) { if (n) body(0u); return; }
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < n; ++t) pool.emplace_back([=]() { body(t); });
    for (auto& th : pool) th.join();
}

struct Chunk {
    char* begin = nullptr;
    char* end = nullptr;
    std::vector<float> v, n, t;
    std::vector<char*> faceLines;
    std::vector<int> faceLineNums; // local, 1-based
    std::vector<Corner> faces;
    int lines = 0;
    int errLine = 0;               // local line of the first error, 0 = none
    const char* errMsg = nullptr;
    void fail(int line, const char* msg) {
        if (errLine == 0 || line < errLine) { errLine = line; errMsg = msg; }
    }
};

// pass A: positions / normals / uvs, and where the face lines are
void parseVertexLines(Chunk& c) {
    char* p = c.begin;
    int lineNum = 0;
    while (p < c.end) {
        char* eol = (char*)memchr(p, '\n', (size_t)(c.end - p));
        char* next = eol ? eol + 1 : c.end;
        if (eol) *eol = '\0';
        ++lineNum;
        char* s = p;
        p = next;
        while (isWs(*s)) ++s;
        if (s[0] == 'v' && (isWs(s[1]) || s[1] == '\0')) {
            float v[3];
            if (!parseFloats(s + 1, 3, v)) { c.fail(lineNum, "position syntax error"); break; }
            c.v.insert(c.v.end(), v, v + 3);
        } else if (s[0] == 'v' && s[1] == 'n' && (isWs(s[2]) || s[2] == '\0')) {
            float v[3];
            if (!parseFloats(s + 2, 3, v)) { c.fail(lineNum, "normal syntax error"); break; }
            c.n.insert(c.n.end(), v, v + 3);
        } else if (s[0] == 'v' && s[1] == 't' && (isWs(s[2]) || s[2] == '\0')) {
            float v[2];
            if (!parseFloats(s + 2, 2, v)) { c.fail(lineNum, "uv syntax error"); break; }
            c.t.insert(c.t.end(), v, v + 2);
        } else if (s[0] == 'f' && (isWs(s[1]) || s[1] == '\0')) {
            c.faceLines.push_back(s + 1);
            c.faceLineNums.push_back(lineNum);
        }
    }
    c.lines = lineNum;
}

// pass B: the face lines, scanned with the format the file's first face fixed
void parseFaceLines(Chunk& c, Format format) {
    c.faces.reserve(3 * c.faceLines.size());
    for (size_t k = 0; k < c.faceLines.size(); ++k) {
        const char* tok[5];
        int ntok = 0;
        char* q = c.faceLines[k];
        while (true) {
            while (isWs(*q)) ++q;
            if (*q == '\0') break;
            if (ntok < 5) tok[ntok] = q;
            ++ntok;
            while (*q != '\0' && !isWs(*q)) ++q;
            if (*q != '\0') *q++ = '\0';
        }
        if (ntok > 4 || ntok < 3) { c.fail(c.faceLineNums[k], "incorrect face vertices number"); return; }
        Corner cs[4];
        for (int i = 0; i < ntok; ++i) {
            const char* t = tok[i];
            cs[i].v = atoi(t);
            cs[i].n = 0;
            cs[i].t = 0;
            // the reference advances past the separators with strcspn + 1 regardless of what is
            // there; clamp at the terminator
            auto advance = [](const char* z, int extra) {
                z += strcspn(z, "/");
                for (int e = 0; e < extra && *z != '\0'; ++e) ++z;
                return z;
            };
            switch (format) {
            case kV: break;
            case kVT: t = advance(t, 1); cs[i].t = atoi(t); break;
            case kVN: t = advance(t, 2); cs[i].n = atoi(t); break;
            case kVTN:
                t = advance(t, 1); cs[i].t = atoi(t);
                t = advance(t, 1); cs[i].n = atoi(t);
                break;
            }
        }
        c.faces.push_back(cs[0]); c.faces.push_back(cs[1]); c.faces.push_back(cs[2]);
        if (ntok == 4) { // quad -> (0,1,2) + (0,2,3)
            c.faces.push_back(cs[0]); c.faces.push_back(cs[2]); c.faces.push_back(cs[3]);
        }
    }
}

template <typename T>
void concat(const std::vector<Chunk>& chunks, std::vector<T> Chunk::*member, std::vector<T>& out) {
    std::vector<size_t> off(chunks.size() + 1, 0);
    for (size_t k = 0; k < chunks.size(); ++k) off[k + 1] = off[k] + (chunks[k].*member).size();
    out.resize(off.back());
    runParallel((unsigned)chunks.size(), [&](unsigned k) {
        const std::vector<T>& src = chunks[k].*member;
        if (!src.empty()) std::memcpy(out.data() + off[k], src.data(), src.size() * sizeof(T));
    });
}

} // namespace

// The file is cut into one chunk per core at line boundaries; vertex data and faces are parsed
// per chunk and concatenated in file order, and the vertex de-duplication (first occurrence of a
// v/vt/vn corner in face order defines the output vertex index,
// src/GoblinPolygonMesh.cpp:66,234-258) runs as a bucketed parallel pass.  The result is
// identical to a sequential parse; only the 10 M-triangle meshes care (14 s -> ~3 s on 8 cores).
bool loadObjMesh(const std::string& path, MeshData* mesh, std::string* error) {
    *mesh = MeshData();
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) {
        if (error) *error = "can't open obj file: " + path;
        return false;
    }
    fseek(f, 0, SEEK_END);
    long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)size + 1);
    size_t got = fread(buf.data(), 1, (size_t)size, f);
    fclose(f);
    buf[got] = '\0';

    // ---- chunks
    unsigned nThreads = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    if (got < (1u << 20)) nThreads = 1;
    std::vector<Chunk> chunks(nThreads);
    {
        char* base = buf.data();
        char* end = base + got;
        char* cur = base;
        for (unsigned k = 0; k < nThreads; ++k) {
            chunks[k].begin = cur;
            char* target = k + 1 == nThreads ? end : base + (size_t)((double)got * (k + 1) / nThreads);
            if (target < cur) target = cur;
            if (target < end) {
                char* nl = (char*)memchr(target, '\n', (size_t)(end - target));
                target = nl ? nl + 1 : end;
            }
            chunks[k].end = target;
            cur = target;
        }
    }
    runParallel(nThreads, [&](unsigned k) { parseVertexLines(chunks[k]); });
    // the first face of the file fixes the scan format
    Format format = kV;
    for (const Chunk& c : chunks) {
        if (c.faceLines.empty()) continue;
        const char* t0 = c.faceLines[0];
        while (isWs(*t0)) ++t0;
        std::string first(t0, strcspn(t0, " \t\r\v\f"));
        const char* s0 = first.c_str();
        const char* s1 = strchr(s0, '/');
        if (strstr(s0, "//")) { format = kVN; mesh->hasNormal = true; }
        else if (!s1) { format = kV; }
        else if (s1 == strrchr(s0, '/')) { format = kVT; mesh->hasUv = true; }
        else { format = kVTN; mesh->hasNormal = true; mesh->hasUv = true; }
        break;
    }
    runParallel(nThreads, [&](unsigned k) { parseFaceLines(chunks[k], format); });
    {
        int lineBase = 0;
        for (const Chunk& c : chunks) { // the first error in file order, as a sequential parse reports it
            if (c.errLine) {
                if (error) *error = std::string(c.errMsg) + " on line " + std::to_string(lineBase + c.errLine) + " of " + path;
                *mesh = MeshData();
                return false;
            }
            lineBase += c.lines;
        }
    }
    std::vector<float> vlist, nlist, tlist;
    std::vector<Corner> faces; // 3 corners per triangle
    concat(chunks, &Chunk::v, vlist);
    concat(chunks, &Chunk::n, nlist);
    concat(chunks, &Chunk::t, tlist);
    concat(chunks, &Chunk::faces, faces);
    std::vector<Chunk>().swap(chunks);
    std::vector<char>().swap(buf);

    const int nv = (int)(vlist.size() / 3), nn = (int)(nlist.size() / 3), nt = (int)(tlist.size() / 2);
    const size_t nC = faces.size();
    const unsigned P = nC < (1u << 16) ? 1u : nThreads;
    const size_t per = (nC + P - 1) / P;
    {
        std::vector<size_t> bad(P, (size_t)-1);
        runParallel(P, [&](unsigned k) {
            for (size_t i = k * per, e = std::min(nC, (k + 1) * per); i < e; ++i) {
                Corner& c = faces[i];
                if (c.v < 0) c.v += nv + 1; // python-style backward indices
                if (c.n < 0) c.n += nn + 1;
                if (c.t < 0) c.t += nt + 1;
                --c.v; --c.n; --c.t;        // OBJ is 1-based; 0 (absent) becomes -1
                if (c.v < 0 || c.v >= nv || c.n < -1 || c.n >= nn || c.t < -1 || c.t >= nt) {
                    if (bad[k] == (size_t)-1) bad[k] = i;
                }
            }
        });
        for (unsigned k = 0; k < P; ++k) {
            if (bad[k] != (size_t)-1) {
                if (error) *error = "invalid index in face " + std::to_string(bad[k] / 3) + " of " + path;
                *mesh = MeshData();
                return false;
            }
        }
    }

    // ---- de-duplicate corners: output vertex k is the k-th distinct corner in face order
    std::vector<uint32_t> firstPos(nC); // position of the first occurrence of corner i's key
    if (P == 1) {
        CornerMap map(nC / 2 + 16);
        for (size_t i = 0; i < nC; ++i) {
            uint32_t first;
            map.insert(faces[i], (uint32_t)i, &first);
            firstPos[i] = first;
        }
    } else {
        // bucket by key hash: a key's occurrences all land in one bucket, in ascending position
        std::vector<std::vector<uint32_t>> cell((size_t)P * P); // [chunk][bucket] -> positions
        runParallel(P, [&](unsigned k) {
            for (size_t i = k * per, e = std::min(nC, (k + 1) * per); i < e; ++i) {
                cell[(size_t)k * P + (CornerMap::hash(faces[i]) >> 40) % P].push_back((uint32_t)i);
            }
        });
        runParallel(P, [&](unsigned b) {
            size_t total = 0;
            for (unsigned k = 0; k < P; ++k) total += cell[(size_t)k * P + b].size();
            CornerMap map(total / 2 + 16);
            for (unsigned k = 0; k < P; ++k) {
                for (uint32_t i : cell[(size_t)k * P + b]) {
                    uint32_t first;
                    map.insert(faces[i], i, &first);
                    firstPos[i] = first;
                }
            }
        });
    }
    std::vector<uint32_t> newIndex(nC);
    uint32_t counter = 0;
    for (size_t i = 0; i < nC; ++i) {
        newIndex[i] = counter;
        if (firstPos[i] == i) ++counter;
    }
    mesh->idx.resize(nC);
    mesh->pos.resize(3 * (size_t)counter);
    mesh->nrm.resize(3 * (size_t)counter);
    mesh->uv.resize(2 * (size_t)counter);
    runParallel(P, [&](unsigned k) {
        for (size_t i = k * per, e = std::min(nC, (k + 1) * per); i < e; ++i) {
            const uint32_t out = newIndex[firstPos[i]];
            mesh->idx[i] = out;
            if (firstPos[i] != i) continue;
            const Corner& c = faces[i];
            float* pp = &mesh->pos[3 * (size_t)out];
            pp[0] = vlist[3 * c.v]; pp[1] = vlist[3 * c.v + 1]; pp[2] = vlist[3 * c.v + 2];
            float* pn = &mesh->nrm[3 * (size_t)out];
            if (c.n == -1) { pn[0] = pn[1] = pn[2] = 0.0f; }
            else { pn[0] = nlist[3 * c.n]; pn[1] = nlist[3 * c.n + 1]; pn[2] = nlist[3 * c.n + 2]; }
            float* pt = &mesh->uv[2 * (size_t)out];
            if (c.t == -1) { pt[0] = pt[1] = 0.0f; }
            else { pt[0] = tlist[2 * c.t]; pt[1] = tlist[2 * c.t + 1]; }
        }
    });
    for (size_t v = 0; v < counter; ++v) { // min / max: order does not matter
        mesh->bound.expand(Vec3(mesh->pos[3 * v], mesh->pos[3 * v + 1], mesh->pos[3 * v + 2]));
    }

    // PolygonMesh::recalculateArea, src/GoblinPolygonMesh.cpp:347-359 (a sequential float sum)
    mesh->area = 0.0f;
    for (size_t t = 0; t + 2 < mesh->idx.size(); t += 3) {
        const float* a = &mesh->pos[3 * mesh->idx[t]];
        const float* b = &mesh->pos[3 * mesh->idx[t + 1]];
        const float* c = &mesh->pos[3 * mesh->idx[t + 2]];
        Vec3 v0(a[0], a[1], a[2]), v1(b[0], b[1], b[2]), v2(c[0], c[1], c[2]);
        mesh->area += 0.5f * length(cross(v1 - v0, v2 - v0));
    }
    return true;
}

} // namespace gb
