// scene_gen: deterministic synthetic scenes in Goblin's own JSON + OBJ formats
// (scenes S1..S5 plus a tiny mixed scene for golden
// vectors).  The same files feed the reference oracle and the GPU renderer.
//
// All randomness is an integer hash; floats are written with a decimal point
// because the reference's ParamSet does not convert JSON ints to floats
// (GoblinParamSet.cpp:104-120).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <sys/stat.h>

struct V3 { double x, y, z; };
static V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
static V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static double len(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
static V3 norm(V3 a) { double l = len(a); return l > 0 ? a * (1.0 / l) : a; }

static uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
static uint32_t hash3(int32_t x, int32_t y, int32_t z, uint32_t seed) {
    uint32_t h = hash32((uint32_t)x * 0x9e3779b1U + seed);
    h = hash32(h ^ ((uint32_t)y * 0x85ebca77U));
    h = hash32(h ^ ((uint32_t)z * 0xc2b2ae3dU));
    return h;
}
static double u01(uint32_t h) { return (h >> 8) * (1.0 / 16777216.0); }
static double rnd(uint32_t seed, uint32_t i, uint32_t k) { return u01(hash3((int)i, (int)k, 17, seed)); }

static double valueNoise(V3 p, uint32_t seed) {
    double fx = std::floor(p.x), fy = std::floor(p.y), fz = std::floor(p.z);
    int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    double tx = p.x - fx, ty = p.y - fy, tz = p.z - fz;
    tx = tx * tx * (3 - 2 * tx); ty = ty * ty * (3 - 2 * ty); tz = tz * tz * (3 - 2 * tz);
    double c[2][2][2];
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int d = 0; d < 2; ++d)
        c[a][b][d] = u01(hash3(ix + a, iy + b, iz + d, seed));
    auto lerp = [](double a, double b, double t) { return a + (b - a) * t; };
    double x00 = lerp(c[0][0][0], c[1][0][0], tx), x10 = lerp(c[0][1][0], c[1][1][0], tx);
    double x01 = lerp(c[0][0][1], c[1][0][1], tx), x11 = lerp(c[0][1][1], c[1][1][1], tx);
    return lerp(lerp(x00, x10, ty), lerp(x01, x11, ty), tz);
}
static double fbm(V3 p, uint32_t seed, int octaves) {
    double a = 0.5, s = 0.0;
    for (int o = 0; o < octaves; ++o) { s += a * (valueNoise(p, seed + o * 101) - 0.5); p = p * 2.0; a *= 0.5; }
    return s;
}

static std::string ff(double v) {
    char b[64];
    snprintf(b, sizeof b, "%.9g", (double)(float)v);
    if (!strpbrk(b, ".eE")) strcat(b, ".0");
    return b;
}
static std::string v3(double a, double b, double c) { return "[" + ff(a) + ", " + ff(b) + ", " + ff(c) + "]"; }
static std::string v4(double a, double b, double c, double d) {
    return "[" + ff(a) + ", " + ff(b) + ", " + ff(c) + ", " + ff(d) + "]";
}

struct Mesh {
    std::vector<V3> pos, nrm;
    std::vector<double> uv; // 2 per vertex, optional
    std::vector<uint32_t> idx;
};

static void computeNormals(Mesh& m) {
    m.nrm.assign(m.pos.size(), V3{0, 0, 0});
    for (size_t t = 0; t + 2 < m.idx.size(); t += 3) {
        uint32_t a = m.idx[t], b = m.idx[t + 1], c = m.idx[t + 2];
        V3 n = cross(m.pos[b] - m.pos[a], m.pos[c] - m.pos[a]);
        m.nrm[a] = m.nrm[a] + n; m.nrm[b] = m.nrm[b] + n; m.nrm[c] = m.nrm[c] + n;
    }
    for (V3& n : m.nrm) n = norm(n);
}

// format: 0 = "f v", 1 = "f v//vn", 2 = "f v/vt/vn", 3 = "f v/vt"
static void writeObj(const std::string& path, const Mesh& m, int format, bool quads = false) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
    std::vector<char> buf(1 << 22);
    setvbuf(f, buf.data(), _IOFBF, buf.size());
    for (const V3& p : m.pos) fprintf(f, "v %.9g %.9g %.9g\n", (double)(float)p.x, (double)(float)p.y, (double)(float)p.z);
    if (format == 2 || format == 3)
        for (size_t i = 0; i < m.pos.size(); ++i) fprintf(f, "vt %.9g %.9g\n", (double)(float)m.uv[2 * i], (double)(float)m.uv[2 * i + 1]);
    if (format == 1 || format == 2)
        for (const V3& n : m.nrm) fprintf(f, "vn %.9g %.9g %.9g\n", (double)(float)n.x, (double)(float)n.y, (double)(float)n.z);
    auto tok = [&](uint32_t i) {
        unsigned v = i + 1;
        switch (format) {
        case 0: fprintf(f, " %u", v); break;
        case 1: fprintf(f, " %u//%u", v, v); break;
        case 2: fprintf(f, " %u/%u/%u", v, v, v); break;
        default: fprintf(f, " %u/%u", v, v); break;
        }
    };
    if (quads) { // idx holds 4 per face
        for (size_t t = 0; t + 3 < m.idx.size(); t += 4) {
            fputc('f', f); tok(m.idx[t]); tok(m.idx[t + 1]); tok(m.idx[t + 2]); tok(m.idx[t + 3]); fputc('\n', f);
        }
    } else {
        for (size_t t = 0; t + 2 < m.idx.size(); t += 3) {
            fputc('f', f); tok(m.idx[t]); tok(m.idx[t + 1]); tok(m.idx[t + 2]); fputc('\n', f);
        }
    }
    fclose(f);
}

static Mesh icosphere(int levels) {
    Mesh m;
    const double t = (1.0 + std::sqrt(5.0)) / 2.0;
    const double P[12][3] = {{-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0}, {0, -1, t}, {0, 1, t},
        {0, -1, -t}, {0, 1, -t}, {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1}};
    const int F[20][3] = {{0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11}, {1, 5, 9}, {5, 11, 4},
        {11, 10, 2}, {10, 7, 6}, {7, 1, 8}, {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8}, {3, 8, 9},
        {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};
    for (auto& p : P) m.pos.push_back(norm(V3{p[0], p[1], p[2]}));
    for (auto& f : F) { m.idx.push_back(f[0]); m.idx.push_back(f[1]); m.idx.push_back(f[2]); }
    for (int l = 0; l < levels; ++l) {
        std::map<uint64_t, uint32_t> mid;
        std::vector<uint32_t> out;
        auto midpoint = [&](uint32_t a, uint32_t b) {
            uint64_t key = a < b ? ((uint64_t)a << 32) | b : ((uint64_t)b << 32) | a;
            auto it = mid.find(key);
            if (it != mid.end()) return it->second;
            uint32_t id = (uint32_t)m.pos.size();
            m.pos.push_back(norm((m.pos[a] + m.pos[b]) * 0.5));
            mid[key] = id;
            return id;
        };
        for (size_t i = 0; i + 2 < m.idx.size(); i += 3) {
            uint32_t a = m.idx[i], b = m.idx[i + 1], c = m.idx[i + 2];
            uint32_t ab = midpoint(a, b), bc = midpoint(b, c), ca = midpoint(c, a);
            uint32_t tri[12] = {a, ab, ca, b, bc, ab, c, ca, bc, ab, bc, ca};
            out.insert(out.end(), tri, tri + 12);
        }
        m.idx.swap(out);
    }
    return m;
}

// Stand-in for the absent examples/models/bunny.obj (.MISSING_LARGE_BLOBS):
// a noise-displaced icosphere fitted to the Stanford bunny's approximate bounds
// so examples/bunny.json's instance transform rests it on the floor.
static Mesh bunnyStandin(int levels, uint32_t seed) {
    Mesh m = icosphere(levels);
    for (V3& p : m.pos) {
        V3 d = p;
        double disp = 1.0 + 0.55 * fbm(d * 2.3 + V3{3.1, 1.7, 0.4}, seed, 3)
            + 0.25 * std::max(0.0, d.y) * (valueNoise(d * 4.0, seed + 7) - 0.3);
        V3 q = d * disp;
        p = V3{-0.0168 + 0.072 * q.x, 0.110 + 0.070 * q.y, -0.0015 + 0.056 * q.z};
    }
    double ymin = 1e30;
    for (V3& p : m.pos) ymin = std::min(ymin, p.y);
    for (V3& p : m.pos) p.y += 0.0330 - ymin;
    computeNormals(m);
    return m;
}

static Mesh gridMesh(int n, double half, double amp, uint32_t seed, bool withUv) {
    Mesh m; // (n+1)^2 vertices over [-half, half]^2 in xz, height from noise
    for (int j = 0; j <= n; ++j) for (int i = 0; i <= n; ++i) {
        double x = -half + 2 * half * i / n, z = -half + 2 * half * j / n;
        double y = amp == 0.0 ? 0.0 : amp * fbm(V3{x * 0.35, 0.0, z * 0.35}, seed, 5);
        m.pos.push_back(V3{x, y, z});
        if (withUv) { m.uv.push_back((double)i / n); m.uv.push_back((double)j / n); }
    }
    for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
        uint32_t a = j * (n + 1) + i, b = a + 1, c = a + (n + 1), d = c + 1;
        uint32_t t[6] = {a, c, b, b, c, d}; // +y facing
        m.idx.insert(m.idx.end(), t, t + 6);
    }
    computeNormals(m);
    return m;
}

static void quatYaw(double deg, double q[4]) {
    double h = deg * M_PI / 360.0;
    q[0] = std::cos(h); q[1] = 0; q[2] = std::sin(h); q[3] = 0;
}
static void quatHashed(uint32_t seed, uint32_t i, double tilt, double q[4]) {
    // small random tilt around a hashed horizontal axis
    double ax = rnd(seed, i, 40) * 2 - 1, az = rnd(seed, i, 41) * 2 - 1;
    double l = std::sqrt(ax * ax + az * az) + 1e-9;
    double h = tilt * (rnd(seed, i, 42) * 2 - 1) * 0.5;
    q[0] = std::cos(h); q[1] = std::sin(h) * ax / l; q[2] = 0; q[3] = std::sin(h) * az / l;
}

struct Json {
    FILE* f;
    explicit Json(const std::string& path) { f = fopen(path.c_str(), "w"); if (!f) { perror(path.c_str()); exit(2); } }
    ~Json() { fclose(f); }
    void raw(const std::string& s) { fputs(s.c_str(), f); }
};

static void renderSetting(Json& j, const char* method, int spp, int depth, int ao = 0) {
    fprintf(j.f, "  \"render_setting\": {\"render_method\": \"%s\", \"sample_per_pixel\": %d, \"max_ray_depth\": %d",
        method, spp, depth);
    if (ao) fprintf(j.f, ", \"ao_sample_num\": %d", ao);
    fprintf(j.f, "},\n");
}
static void camera(Json& j, const std::string& pos, const std::string& orient, double fov, int w, int h,
    const char* file = nullptr) {
    fprintf(j.f, "  \"camera\": {\"position\": %s, \"orientation\": %s, \"fov\": %s, \"near_plane\": 0.1, "
        "\"far_plane\": 5000.0,\n    \"film\": {\"resolution\": [%d, %d]%s%s%s},\n"
        "    \"filter\": {\"type\": \"gaussian\", \"width\": [2.0, 2.0], \"falloff\": 2.0}},\n",
        pos.c_str(), orient.c_str(), ff(fov).c_str(), w, h, file ? ", \"file\": \"" : "", file ? file : "",
        file ? "\"" : "");
}

// orientation quaternion (w,x,y,z) that looks from eye toward target, D3D LH, +Z forward, +Y up
static std::string lookAt(V3 eye, V3 target) {
    V3 fwd = norm(target - eye);
    V3 right = norm(cross(V3{0, 1, 0}, fwd));
    V3 up = cross(fwd, right);
    double R[3][3] = {{right.x, up.x, fwd.x}, {right.y, up.y, fwd.y}, {right.z, up.z, fwd.z}};
    double tr = R[0][0] + R[1][1] + R[2][2], q[4];
    if (tr > 0) {
        double s = std::sqrt(tr + 1.0) * 2; q[0] = 0.25 * s;
        q[1] = (R[2][1] - R[1][2]) / s; q[2] = (R[0][2] - R[2][0]) / s; q[3] = (R[1][0] - R[0][1]) / s;
    } else if (R[0][0] > R[1][1] && R[0][0] > R[2][2]) {
        double s = std::sqrt(1.0 + R[0][0] - R[1][1] - R[2][2]) * 2; q[0] = (R[2][1] - R[1][2]) / s;
        q[1] = 0.25 * s; q[2] = (R[0][1] + R[1][0]) / s; q[3] = (R[0][2] + R[2][0]) / s;
    } else if (R[1][1] > R[2][2]) {
        double s = std::sqrt(1.0 + R[1][1] - R[0][0] - R[2][2]) * 2; q[0] = (R[0][2] - R[2][0]) / s;
        q[1] = (R[0][1] + R[1][0]) / s; q[2] = 0.25 * s; q[3] = (R[1][2] + R[2][1]) / s;
    } else {
        double s = std::sqrt(1.0 + R[2][2] - R[0][0] - R[1][1]) * 2; q[0] = (R[1][0] - R[0][1]) / s;
        q[1] = (R[0][2] + R[2][0]) / s; q[2] = (R[1][2] + R[2][1]) / s; q[3] = 0.25 * s;
    }
    return v4(q[0], q[1], q[2], q[3]);
}

static void mkdirs(const std::string& d) { mkdir(d.c_str(), 0755); mkdir((d + "/models").c_str(), 0755); }

// ------------------------------------------------------------------ S1 / S2
// examples/bunny.json verbatim except render_method / max_ray_depth (and the
// unused "sphere" geometry kept); mesh = stand-in.
static void genBunny(const std::string& dir, int levels, const char* name, const char* method, int w, int h,
    int spp, int depth, int ao) {
    Json j(dir + "/" + name);
    j.raw("{\n");
    renderSetting(j, method, spp, depth, ao);
    camera(j, "[1.77271, 0.774149, 0.830583]", "[0.352428, 0.114509, -0.883349, 0.287014]", 45.0, w, h);
    j.raw("  \"geometries\": [\n"
          "    {\"name\": \"bunny\", \"type\": \"mesh\", \"file\": \"models/bunny.obj\"},\n"
          "    {\"name\": \"plane\", \"type\": \"mesh\", \"file\": \"models/plane.obj\"},\n"
          "    {\"name\": \"sphere\", \"type\": \"sphere\", \"radius\": 0.05}\n  ],\n"
          "  \"lights\": [\n"
          "    {\"name\": \"spot light\", \"type\": \"spot\", \"intensity\": [200, 200, 200], \"position\": [-10, 5, 15],\n"
          "     \"target\": [0.323236, -0.44923, 0.0354459], \"theta_max\": 10.0, \"falloff_start\": 5.0}\n  ],\n"
          "  \"textures\": [\n"
          "    {\"format\": \"color\", \"name\": \"purple\", \"type\": \"constant\", \"color\": [0.7, 0.7, 1]},\n"
          "    {\"format\": \"color\", \"name\": \"white\", \"type\": \"constant\", \"color\": [1, 1, 1]}\n  ],\n"
          "  \"materials\": [\n"
          "    {\"name\": \"white\", \"type\": \"lambert\", \"Kd\": \"white\"},\n"
          "    {\"name\": \"glass\", \"type\": \"transparent\", \"Kr\": \"purple\", \"Kt\": \"purple\", \"index\": 1.5}\n  ],\n"
          "  \"primitives\": [\n"
          "    {\"type\": \"model\", \"name\": \"bunny\", \"geometry\": \"bunny\", \"material\": \"glass\"},\n"
          "    {\"type\": \"model\", \"name\": \"plane\", \"geometry\": \"plane\", \"material\": \"white\"},\n"
          "    {\"type\": \"instance\", \"name\": \"bunny\", \"model\": \"bunny\", \"position\": [0.4, -1, 0],\n"
          "     \"orientation\": [0.96592582628, 0, 0.2588190451, 0], \"scale\": [5, 5, 5]},\n"
          "    {\"type\": \"instance\", \"name\": \"floor\", \"model\": \"plane\", \"position\": [0, -0.835065, 0],\n"
          "     \"orientation\": [1, 0, 0, 0], \"scale\": [200, 200, 200]}\n  ]\n}\n");
    (void)levels;
}

static void writePlane(const std::string& path) {
    // unit quad in xz with uv + normals, two triangles (same shape as the
    // reference's examples/models/plane.obj; written from scratch)
    Mesh m = gridMesh(1, 1.0, 0.0, 0, true);
    // gridMesh's z runs -half..half with +y winding; fine for a floor
    writeObj(path, m, 2);
}

// --------------------------------------------------------------------- tiny
static void genTiny(const std::string& dir) {
    mkdirs(dir);
    writeObj(dir + "/models/floor.obj", gridMesh(16, 1.0, 0.04, 5, true), 2);
    Mesh ico = icosphere(2);
    writeObj(dir + "/models/ico_flat.obj", ico, 0);
    Mesh blob = bunnyStandin(2, 99);
    writeObj(dir + "/models/blob.obj", blob, 1);
    // a quad-faced box (tests the quad split) with v/vt only.  uv = (x - z, y - z) / 2 + 1/2: the
    // map's null direction (1, 1, 1) lies in no face plane, so no triangle has a zero uv
    // determinant (the reference reads a stale Fragment for those, GoblinTriangle.cpp:113-117).
    Mesh box;
    const double c[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
    for (auto& p : c) {
        box.pos.push_back(V3{p[0] * 0.5, p[1] * 0.5, p[2] * 0.5});
        box.uv.push_back((p[0] - p[2]) * 0.25 + 0.5);
        box.uv.push_back((p[1] - p[2]) * 0.25 + 0.5);
    }
    const uint32_t q[24] = {0, 3, 2, 1, 4, 5, 6, 7, 0, 1, 5, 4, 2, 3, 7, 6, 1, 2, 6, 5, 0, 4, 7, 3};
    box.idx.assign(q, q + 24);
    writeObj(dir + "/models/box.obj", box, 3, true);
    // a small bumpy panel with vn: the mesh emitter (GeometrySet over 8 triangles)
    writeObj(dir + "/models/panel.obj", gridMesh(2, 0.5, 0.06, 9, true), 2);

    const char* methods[2] = {"path_tracing", "ao"};
    const char* names[2] = {"tiny_pt.json", "tiny_ao.json"};
    for (int k = 0; k < 2; ++k) {
        Json j(dir + "/" + names[k]);
        j.raw("{\n");
        renderSetting(j, methods[k], 16, 6, k == 1 ? 9 : 0);
        camera(j, v3(0.3, 2.2, -6.5), lookAt(V3{0.3, 2.2, -6.5}, V3{0, 0.6, 0}), 40.0, 96, 64);
        j.raw("  \"geometries\": [\n"
              "    {\"name\": \"floor\", \"type\": \"mesh\", \"file\": \"models/floor.obj\"},\n"
              "    {\"name\": \"ico\", \"type\": \"mesh\", \"file\": \"models/ico_flat.obj\"},\n"
              "    {\"name\": \"blob\", \"type\": \"mesh\", \"file\": \"models/blob.obj\"},\n"
              "    {\"name\": \"box\", \"type\": \"mesh\", \"file\": \"models/box.obj\"},\n"
              "    {\"name\": \"ball\", \"type\": \"sphere\", \"radius\": 0.6},\n"
              "    {\"name\": \"unitball\", \"type\": \"sphere\"},\n"
              "    {\"name\": \"plate\", \"type\": \"disk\", \"radius\": 0.8},\n"
              "    {\"name\": \"lamp\", \"type\": \"disk\", \"radius\": 0.7},\n"
              "    {\"name\": \"bulb\", \"type\": \"sphere\", \"radius\": 0.25},\n"
              "    {\"name\": \"panel\", \"type\": \"mesh\", \"file\": \"models/panel.obj\"}\n  ],\n"
              "  \"textures\": [\n"
              "    {\"format\": \"color\", \"name\": \"grey\", \"type\": \"constant\", \"color\": [0.6, 0.6, 0.6]},\n"
              "    {\"format\": \"color\", \"name\": \"red\", \"type\": \"constant\", \"color\": [0.8, 0.25, 0.2]},\n"
              "    {\"format\": \"color\", \"name\": \"green\", \"type\": \"constant\", \"color\": [0.2, 0.7, 0.3]},\n"
              "    {\"format\": \"color\", \"name\": \"tint\", \"type\": \"constant\", \"color\": [0.9, 0.95, 1.0]},\n"
              "    {\"format\": \"color\", \"name\": \"white\", \"type\": \"constant\", \"color\": [1, 1, 1]},\n"
              "    {\"format\": \"float\", \"name\": \"shiny\", \"type\": \"constant\", \"float\": 40.0},\n"
              "    {\"format\": \"float\", \"name\": \"satin\", \"type\": \"constant\", \"float\": 6.0}\n  ],\n"
              "  \"materials\": [\n"
              "    {\"name\": \"gloss\", \"type\": \"blinn\", \"Kg\": \"red\", \"exponent\": \"shiny\", \"index\": 1.5},\n"
              "    {\"name\": \"metal\", \"type\": \"blinn\", \"Kg\": \"tint\", \"exponent\": \"satin\", \"index\": 0.8, \"k\": 3.0},\n"
              "    {\"name\": \"grey\", \"type\": \"lambert\", \"Kd\": \"grey\"},\n"
              "    {\"name\": \"red\", \"type\": \"lambert\", \"Kd\": \"red\"},\n"
              "    {\"name\": \"green\", \"type\": \"lambert\", \"Kd\": \"green\"},\n"
              "    {\"name\": \"mirror\", \"type\": \"mirror\", \"Kr\": \"white\"},\n"
              "    {\"name\": \"glass\", \"type\": \"transparent\", \"Kr\": \"tint\", \"Kt\": \"tint\", \"index\": 1.5}\n  ],\n"
              "  \"primitives\": [\n"
              "    {\"type\": \"model\", \"name\": \"floor\", \"geometry\": \"floor\", \"material\": \"grey\"},\n"
              "    {\"type\": \"model\", \"name\": \"ico\", \"geometry\": \"ico\", \"material\": \"metal\"},\n"
              "    {\"type\": \"model\", \"name\": \"blob\", \"geometry\": \"blob\", \"material\": \"glass\"},\n"
              "    {\"type\": \"model\", \"name\": \"box\", \"geometry\": \"box\", \"material\": \"green\"},\n"
              "    {\"type\": \"model\", \"name\": \"ball\", \"geometry\": \"ball\", \"material\": \"mirror\"},\n"
              "    {\"type\": \"model\", \"name\": \"gball\", \"geometry\": \"unitball\", \"material\": \"glass\"},\n"
              "    {\"type\": \"model\", \"name\": \"plate\", \"geometry\": \"plate\", \"material\": \"gloss\"},\n"
              "    {\"type\": \"instance\", \"name\": \"floor\", \"model\": \"floor\", \"position\": [0.0, 0.0, 0.0], \"scale\": [6.0, 6.0, 6.0]},\n"
              "    {\"type\": \"instance\", \"name\": \"ico\", \"model\": \"ico\", \"position\": [-1.6, 0.75, 0.4], \"scale\": [0.7, 0.7, 0.7]},\n"
              "    {\"type\": \"instance\", \"name\": \"blob\", \"model\": \"blob\", \"position\": [0.3, -0.15, -1.2],\n"
              "     \"orientation\": [0.96592582628, 0, 0.2588190451, 0], \"scale\": [7.0, 7.0, 7.0]},\n"
              "    {\"type\": \"instance\", \"name\": \"blob2\", \"model\": \"blob\", \"position\": [2.4, -0.1, 1.8],\n"
              "     \"euler\": [0.0, 75.0, 10.0], \"scale\": [5.0, 6.0, 5.0]},\n"
              "    {\"type\": \"instance\", \"name\": \"box\", \"model\": \"box\", \"position\": [1.7, 0.55, 0.2],\n"
              "     \"euler\": [0.0, 30.0, 0.0], \"scale\": [1.0, 1.0, 1.0]},\n"
              "    {\"type\": \"instance\", \"name\": \"ball\", \"model\": \"ball\", \"position\": [0.0, 0.85, 1.6]},\n"
              "    {\"type\": \"instance\", \"name\": \"gball\", \"model\": \"gball\", \"position\": [-0.9, 0.5, -1.6], \"scale\": [0.45, 0.45, 0.45]},\n"
              "    {\"type\": \"instance\", \"name\": \"plate\", \"model\": \"plate\", \"position\": [-2.6, 1.0, 2.0],\n"
              "     \"euler\": [-35.0, 40.0, 0.0]}\n  ],\n"
              "  \"lights\": [\n"
              "    {\"name\": \"lamp\", \"type\": \"area\", \"geometry\": \"lamp\", \"radiance\": [14.0, 13.0, 12.0],\n"
              "     \"position\": [0.5, 4.0, 0.0], \"euler\": [90.0, 0.0, 0.0]},\n"
              "    {\"name\": \"bulb\", \"type\": \"area\", \"geometry\": \"bulb\", \"radiance\": [9.0, 9.0, 12.0],\n"
              "     \"position\": [-2.2, 1.4, -1.4]},\n"
              "    {\"name\": \"panel\", \"type\": \"area\", \"geometry\": \"panel\", \"radiance\": [5.0, 6.0, 8.0],\n"
              "     \"position\": [2.2, 2.8, 1.2], \"euler\": [150.0, 20.0, 0.0], \"scale\": [1.5, 1.5, 1.5]},\n"
              "    {\"name\": \"pt\", \"type\": \"point\", \"intensity\": [6.0, 5.0, 4.0], \"position\": [3.0, 2.5, -2.5]},\n"
              "    {\"name\": \"spot\", \"type\": \"spot\", \"intensity\": [30.0, 30.0, 24.0], \"position\": [-3.0, 4.0, -3.0],\n"
              "     \"target\": [0.0, 0.0, 0.0], \"theta_max\": 25.0, \"falloff_start\": 15.0},\n"
              "    {\"name\": \"sun\", \"type\": \"directional\", \"radiance\": [0.4, 0.4, 0.35], \"direction\": [0.3, -1.0, 0.4]}\n  ]\n}\n");
    }
}

// ----------------------------------------------------------------------- S3
static void genSpheres(const std::string& dir, int w, int h, int spp) {
    mkdirs(dir);
    writeObj(dir + "/models/plane.obj", gridMesh(1, 1.0, 0.0, 0, true), 2);
    const uint32_t seed = 42;
    Json j(dir + "/spheres_pt.json");
    j.raw("{\n");
    renderSetting(j, "path_tracing", spp, 8);
    camera(j, v3(0.0, 14.0, -30.0), lookAt(V3{0, 14, -30}, V3{0, 0, -2}), 45.0, w, h);
    std::string geos, tex, mats, prims;
    geos += "    {\"name\": \"plane\", \"type\": \"mesh\", \"file\": \"models/plane.obj\"},\n";
    geos += "    {\"name\": \"lampdisk\", \"type\": \"disk\", \"radius\": 2.0},\n";
    tex += "    {\"format\": \"color\", \"name\": \"white\", \"type\": \"constant\", \"color\": [1.0, 1.0, 1.0]},\n";
    tex += "    {\"format\": \"color\", \"name\": \"floor\", \"type\": \"constant\", \"color\": [0.55, 0.55, 0.55]},\n";
    mats += "    {\"name\": \"floor\", \"type\": \"lambert\", \"Kd\": \"floor\"},\n";
    mats += "    {\"name\": \"mirror\", \"type\": \"mirror\", \"Kr\": \"white\"},\n";
    mats += "    {\"name\": \"glass\", \"type\": \"transparent\", \"Kr\": \"white\", \"Kt\": \"white\", \"index\": 1.5},\n";
    prims += "    {\"type\": \"model\", \"name\": \"floor\", \"geometry\": \"plane\", \"material\": \"floor\"},\n";
    prims += "    {\"type\": \"instance\", \"name\": \"floor\", \"model\": \"floor\", \"position\": [0.0, 0.0, 0.0], \"scale\": [20.0, 20.0, 20.0]},\n";
    char b[1024];
    for (uint32_t i = 0; i < 1024; ++i) {
        int gx = i % 32, gz = i / 32;
        bool disk = (i % 4) == 3;
        double r = disk ? 0.2 + 0.3 * rnd(seed, i, 0) : 0.15 + 0.30 * rnd(seed, i, 0);
        double x = -20 + 40.0 * (gx + 0.2 + 0.6 * rnd(seed, i, 1)) / 32.0;
        double z = -20 + 40.0 * (gz + 0.2 + 0.6 * rnd(seed, i, 2)) / 32.0;
        double y = disk ? 0.3 + 0.5 * rnd(seed, i, 3) : r + 0.02 + 0.3 * rnd(seed, i, 3);
        int mt = i % 3;
        snprintf(b, sizeof b, "    {\"name\": \"g%u\", \"type\": \"%s\", \"radius\": %s},\n", i, disk ? "disk" : "sphere", ff(r).c_str());
        geos += b;
        std::string matName;
        if (mt == 0) {
            snprintf(b, sizeof b, "    {\"format\": \"color\", \"name\": \"c%u\", \"type\": \"constant\", \"color\": %s},\n", i,
                v3(0.2 + 0.7 * rnd(seed, i, 4), 0.2 + 0.7 * rnd(seed, i, 5), 0.2 + 0.7 * rnd(seed, i, 6)).c_str());
            tex += b;
            snprintf(b, sizeof b, "    {\"name\": \"m%u\", \"type\": \"lambert\", \"Kd\": \"c%u\"},\n", i, i);
            mats += b;
            matName = "m" + std::to_string(i);
        } else {
            matName = mt == 1 ? "mirror" : "glass";
        }
        snprintf(b, sizeof b, "    {\"type\": \"model\", \"name\": \"p%u\", \"geometry\": \"g%u\", \"material\": \"%s\"},\n", i, i, matName.c_str());
        prims += b;
        if (disk) {
            double q[4];
            quatHashed(seed, i, 1.2, q);
            // tilt around a horizontal axis from facing +y: first rotate +z to +y
            snprintf(b, sizeof b, "    {\"type\": \"instance\", \"name\": \"i%u\", \"model\": \"p%u\", \"position\": %s, \"euler\": %s},\n",
                i, i, v3(x, y, z).c_str(), v3(-90.0 + 40.0 * (rnd(seed, i, 7) - 0.5), 360.0 * rnd(seed, i, 8), 0.0).c_str());
            (void)q;
        } else {
            snprintf(b, sizeof b, "    {\"type\": \"instance\", \"name\": \"i%u\", \"model\": \"p%u\", \"position\": %s},\n",
                i, i, v3(x, y, z).c_str());
        }
        prims += b;
    }
    auto chop = [](std::string& s) { size_t p = s.rfind(','); if (p != std::string::npos) s.erase(p, 1); };
    chop(geos); chop(tex); chop(mats); chop(prims);
    j.raw("  \"geometries\": [\n" + geos + "  ],\n  \"textures\": [\n" + tex + "  ],\n  \"materials\": [\n" + mats +
        "  ],\n  \"primitives\": [\n" + prims + "  ],\n  \"lights\": [\n");
    for (int l = 0; l < 4; ++l) {
        double lx = (l % 2 ? 9.0 : -9.0), lz = (l / 2 ? 9.0 : -9.0);
        fprintf(j.f, "    {\"name\": \"lamp%d\", \"type\": \"area\", \"geometry\": \"lampdisk\", \"radiance\": [15.0, 15.0, 15.0],\n"
            "     \"position\": %s, \"euler\": [90.0, 0.0, 0.0]}%s\n", l, v3(lx, 8.0, lz).c_str(), l < 3 ? "," : "");
    }
    j.raw("  ]\n}\n");
}

// ----------------------------------------------------------------------- S4
static void genGrid(const std::string& dir, int n, int w, int h, int spp) {
    mkdirs(dir);
    writeObj(dir + "/models/terrain.obj", gridMesh(n, 10.0, 3.0, 7, false), 1);
    Json j(dir + "/grid_pt.json");
    j.raw("{\n");
    renderSetting(j, "path_tracing", spp, 8);
    camera(j, v3(0.0, 7.0, -16.0), lookAt(V3{0, 7, -16}, V3{0, 0, 0}), 45.0, w, h);
    j.raw("  \"geometries\": [\n"
          "    {\"name\": \"terrain\", \"type\": \"mesh\", \"file\": \"models/terrain.obj\"},\n"
          "    {\"name\": \"sun\", \"type\": \"sphere\", \"radius\": 1.5}\n  ],\n"
          "  \"textures\": [\n"
          "    {\"format\": \"color\", \"name\": \"sand\", \"type\": \"constant\", \"color\": [0.75, 0.65, 0.5]}\n  ],\n"
          "  \"materials\": [\n"
          "    {\"name\": \"sand\", \"type\": \"lambert\", \"Kd\": \"sand\"}\n  ],\n"
          "  \"primitives\": [\n"
          "    {\"type\": \"model\", \"name\": \"terrain\", \"geometry\": \"terrain\", \"material\": \"sand\"},\n"
          "    {\"type\": \"instance\", \"name\": \"terrain\", \"model\": \"terrain\", \"position\": [0.0, 0.0, 0.0]}\n  ],\n"
          "  \"lights\": [\n"
          "    {\"name\": \"sun\", \"type\": \"area\", \"geometry\": \"sun\", \"radiance\": [40.0, 38.0, 34.0], \"position\": [6.0, 9.0, 4.0]},\n"
          "    {\"name\": \"fill\", \"type\": \"point\", \"intensity\": [60.0, 70.0, 90.0], \"position\": [-8.0, 6.0, -6.0]}\n  ]\n}\n");
}

// ----------------------------------------------------------------------- S5
static void genField(const std::string& dir, int levels, int side, int w, int h, int spp) {
    mkdirs(dir);
    writeObj(dir + "/models/bunny.obj", bunnyStandin(levels, 1234), 1);
    writeObj(dir + "/models/plane.obj", gridMesh(1, 1.0, 0.0, 0, true), 2);
    const uint32_t seed = 99;
    Json j(dir + "/field_pt.json");
    j.raw("{\n");
    renderSetting(j, "path_tracing", spp, 8);
    camera(j, v3(0.0, 9.0, -22.0), lookAt(V3{0, 9, -22}, V3{0, 0, -2}), 45.0, w, h);
    j.raw("  \"geometries\": [\n"
          "    {\"name\": \"bunny\", \"type\": \"mesh\", \"file\": \"models/bunny.obj\"},\n"
          "    {\"name\": \"plane\", \"type\": \"mesh\", \"file\": \"models/plane.obj\"},\n"
          "    {\"name\": \"lamp\", \"type\": \"disk\", \"radius\": 4.0}\n  ],\n"
          "  \"textures\": [\n"
          "    {\"format\": \"color\", \"name\": \"white\", \"type\": \"constant\", \"color\": [0.8, 0.8, 0.8]},\n"
          "    {\"format\": \"color\", \"name\": \"clay\", \"type\": \"constant\", \"color\": [0.7, 0.45, 0.35]},\n"
          "    {\"format\": \"color\", \"name\": \"tint\", \"type\": \"constant\", \"color\": [0.8, 0.9, 1.0]}\n  ],\n"
          "  \"materials\": [\n"
          "    {\"name\": \"white\", \"type\": \"lambert\", \"Kd\": \"white\"},\n"
          "    {\"name\": \"clay\", \"type\": \"lambert\", \"Kd\": \"clay\"},\n"
          "    {\"name\": \"glass\", \"type\": \"transparent\", \"Kr\": \"tint\", \"Kt\": \"tint\", \"index\": 1.5}\n  ],\n"
          "  \"primitives\": [\n"
          "    {\"type\": \"model\", \"name\": \"bunny_clay\", \"geometry\": \"bunny\", \"material\": \"clay\"},\n"
          "    {\"type\": \"model\", \"name\": \"bunny_glass\", \"geometry\": \"bunny\", \"material\": \"glass\"},\n"
          "    {\"type\": \"model\", \"name\": \"floor\", \"geometry\": \"plane\", \"material\": \"white\"},\n"
          "    {\"type\": \"instance\", \"name\": \"floor\", \"model\": \"floor\", \"position\": [0.0, 0.0, 0.0], \"scale\": [40.0, 40.0, 40.0]},\n");
    for (int i = 0; i < side * side; ++i) {
        int gx = i % side, gz = i / side;
        double s = 3.0 + 3.0 * rnd(seed, i, 0);
        double x = -18 + 36.0 * (gx + 0.25 + 0.5 * rnd(seed, i, 1)) / side;
        double z = -18 + 36.0 * (gz + 0.25 + 0.5 * rnd(seed, i, 2)) / side;
        double q[4];
        quatYaw(360.0 * rnd(seed, i, 3), q);
        fprintf(j.f, "    {\"type\": \"instance\", \"name\": \"b%d\", \"model\": \"%s\", \"position\": %s, \"orientation\": %s, \"scale\": %s}%s\n",
            i, (i & 1) ? "bunny_glass" : "bunny_clay", v3(x, -0.033 * s, z).c_str(), v4(q[0], q[1], q[2], q[3]).c_str(),
            v3(s, s, s).c_str(), i + 1 < side * side ? "," : "");
    }
    j.raw("  ],\n  \"lights\": [\n"
          "    {\"name\": \"lamp\", \"type\": \"area\", \"geometry\": \"lamp\", \"radiance\": [20.0, 19.0, 17.0],\n"
          "     \"position\": [0.0, 14.0, 0.0], \"euler\": [90.0, 0.0, 0.0]},\n"
          "    {\"name\": \"spot\", \"type\": \"spot\", \"intensity\": [900.0, 900.0, 800.0], \"position\": [-14.0, 12.0, -14.0],\n"
          "     \"target\": [0.0, 0.0, 0.0], \"theta_max\": 40.0, \"falloff_start\": 25.0}\n  ]\n}\n");
}

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: scene_gen <tiny|bunny|spheres|grid|field> <outdir> [size]\n"
            "  bunny   [levels=6]   S1 bunny_pt.json (512x384, 100 spp, depth 8) + S2 bunny_ao.json (1920x1080, 16 spp, ao 25)\n"
            "  spheres              S3 spheres_pt.json (2048x2048, 64 spp)\n"
            "  grid    [n=2236]     S4 grid_pt.json (2*n*n triangles, 1920x1080, 64 spp)\n"
            "  field   [side=27]    S5 field_pt.json (side^2 bunny instances, 3840x2160, 64 spp)\n");
        return 1;
    }
    std::string what = argv[1], dir = argv[2];
    int size = argc > 3 ? atoi(argv[3]) : 0;
    if (what == "tiny") genTiny(dir);
    else if (what == "bunny") {
        int levels = size ? size : 6;
        mkdirs(dir);
        writeObj(dir + "/models/bunny.obj", bunnyStandin(levels, 1234), 1);
        writePlane(dir + "/models/plane.obj");
        genBunny(dir, levels, "bunny_pt.json", "path_tracing", 512, 384, 100, 8, 0);
        genBunny(dir, levels, "bunny_ao.json", "ao", 1920, 1080, 16, 8, 25);
        genBunny(dir, levels, "bunny_pt_small.json", "path_tracing", 128, 96, 16, 8, 0);
    } else if (what == "spheres") genSpheres(dir, 2048, 2048, 64);
    else if (what == "grid") genGrid(dir, size ? size : 2236, 1920, 1080, 64);
    else if (what == "field") genField(dir, 6, size ? size : 27, 3840, 2160, 64);
    else { fprintf(stderr, "unknown scene %s\n", what.c_str()); return 1; }
    return 0;
}
