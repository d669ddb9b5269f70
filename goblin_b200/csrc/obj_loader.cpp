#include "obj_loader.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

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

} // namespace

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

    std::vector<float> vlist, nlist, tlist;
    std::vector<Corner> faces; // 3 corners per triangle
    Format format = kV;
    bool haveFormat = false;
    int lineNum = 0;
    char* p = buf.data();
    char* end = buf.data() + got;
    auto fail = [&](const char* what) {
        if (error) *error = std::string(what) + " on line " + std::to_string(lineNum) + " of " + path;
        *mesh = MeshData();
        return false;
    };
    while (p < end) {
        char* eol = (char*)memchr(p, '\n', (size_t)(end - p));
        char* next = eol ? eol + 1 : end;
        if (eol) *eol = '\0';
        ++lineNum;
        char* s = p;
        p = next;
        while (isWs(*s)) ++s;
        if (s[0] == 'v' && (isWs(s[1]) || s[1] == '\0')) {
            float v[3];
            if (!parseFloats(s + 1, 3, v)) return fail("position syntax error");
            vlist.insert(vlist.end(), v, v + 3);
        } else if (s[0] == 'v' && s[1] == 'n' && (isWs(s[2]) || s[2] == '\0')) {
            float v[3];
            if (!parseFloats(s + 2, 3, v)) return fail("normal syntax error");
            nlist.insert(nlist.end(), v, v + 3);
        } else if (s[0] == 'v' && s[1] == 't' && (isWs(s[2]) || s[2] == '\0')) {
            float v[2];
            if (!parseFloats(s + 2, 2, v)) return fail("uv syntax error");
            tlist.insert(tlist.end(), v, v + 2);
        } else if (s[0] == 'f' && (isWs(s[1]) || s[1] == '\0')) {
            const char* tok[5];
            int ntok = 0;
            char* c = s + 1;
            while (true) {
                while (isWs(*c)) ++c;
                if (*c == '\0') break;
                if (ntok < 5) tok[ntok] = c;
                ++ntok;
                while (*c != '\0' && !isWs(*c)) ++c;
                if (*c != '\0') *c++ = '\0';
            }
            if (ntok > 4 || ntok < 3) return fail("incorrect face vertices number");
            if (!haveFormat) { // the first face fixes the scan format for the file
                haveFormat = true;
                const char* t0 = tok[0];
                const char* s1 = strchr(t0, '/');
                if (strstr(t0, "//")) { format = kVN; mesh->hasNormal = true; }
                else if (!s1) { format = kV; }
                else if (s1 == strrchr(t0, '/')) { format = kVT; mesh->hasUv = true; }
                else { format = kVTN; mesh->hasNormal = true; mesh->hasUv = true; }
            }
            Corner cs[4];
            for (int i = 0; i < ntok; ++i) {
                const char* t = tok[i];
                cs[i].v = atoi(t);
                cs[i].n = 0;
                cs[i].t = 0;
                // the reference advances past the separators with strcspn + 1
                // regardless of what is there; clamp at the terminator
                auto advance = [](const char* q, int extra) {
                    q += strcspn(q, "/");
                    for (int k = 0; k < extra && *q != '\0'; ++k) ++q;
                    return q;
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
            faces.push_back(cs[0]); faces.push_back(cs[1]); faces.push_back(cs[2]);
            if (ntok == 4) { // quad -> (0,1,2) + (0,2,3)
                faces.push_back(cs[0]); faces.push_back(cs[2]); faces.push_back(cs[3]);
            }
        }
    }

    const int nv = (int)(vlist.size() / 3), nn = (int)(nlist.size() / 3), nt = (int)(tlist.size() / 2);
    for (size_t i = 0; i < faces.size(); ++i) {
        Corner& c = faces[i];
        if (c.v < 0) c.v += nv + 1; // python-style backward indices
        if (c.n < 0) c.n += nn + 1;
        if (c.t < 0) c.t += nt + 1;
        --c.v; --c.n; --c.t;        // OBJ is 1-based; 0 (absent) becomes -1
        if (c.v < 0 || c.v >= nv || c.n < -1 || c.n >= nn || c.t < -1 || c.t >= nt) {
            lineNum = 0;
            if (error) *error = "invalid index in face " + std::to_string(i / 3) + " of " + path;
            *mesh = MeshData();
            return false;
        }
    }

    CornerMap map(faces.size() / 2 + 16);
    mesh->idx.resize(faces.size());
    uint32_t counter = 0;
    for (size_t i = 0; i < faces.size(); ++i) {
        const Corner& c = faces[i];
        uint32_t index;
        if (map.insert(c, counter, &index)) {
            Vec3 pos(vlist[3 * c.v], vlist[3 * c.v + 1], vlist[3 * c.v + 2]);
            mesh->pos.insert(mesh->pos.end(), {pos.x, pos.y, pos.z});
            if (c.n == -1) mesh->nrm.insert(mesh->nrm.end(), {0.0f, 0.0f, 0.0f});
            else mesh->nrm.insert(mesh->nrm.end(), {nlist[3 * c.n], nlist[3 * c.n + 1], nlist[3 * c.n + 2]});
            if (c.t == -1) mesh->uv.insert(mesh->uv.end(), {0.0f, 0.0f});
            else mesh->uv.insert(mesh->uv.end(), {tlist[2 * c.t], tlist[2 * c.t + 1]});
            mesh->bound.expand(pos);
            ++counter;
        }
        mesh->idx[i] = index;
    }

    // PolygonMesh::recalculateArea, src/GoblinPolygonMesh.cpp:347-359
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
