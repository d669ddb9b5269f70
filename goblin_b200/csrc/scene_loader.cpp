// Scene loading and flattening: the host half of the drop-in.
//
// Mirrors ContextLoader::load (src/GoblinContextLoader.cpp:447-503) step by
// step -- renderer, camera/film/filter, geometries, textures, materials,
// primitives, lights, scene -- including its soft fallbacks (unknown names map
// to a magenta "error" object, src/GoblinScene.cpp:112-128) and the int/float
// ParamSet quirk, but writes plain arrays instead of an object graph.
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "bvh_builder.h"
#include "image_map.h"
#include "host_scene.h"
#include "obj_loader.h"
#include "param_set.h"

namespace gb {
namespace {

struct Geometry {
    int kind = GB_GEOM_SPHERE;
    float radius = 1.0f;
    MeshData mesh;
    BBox bound;
    float area = 0.0f;
    // filled when the geometry is first referenced by an instanced model
    bool flattened = false;
    uint32_t nodeOffset = 0, nodeCount = 0, triOffset = 0, triCount = 0, vertOffset = 0, vertCount = 0;
};

struct ModelDef {
    int geometry = 0;
    int material = 0;
    int areaLight = -1;
    bool isCameraLens = false;
    int flatIndex = -1; // index in gb_scene::models once referenced
};

struct PrimitiveDef { // SceneCache::mPrimitiveMap holds models and instances alike
    bool isInstance = false;
    int model = -1;    // ModelDef index when !isInstance
};

struct Color3 { float r = 0, g = 0, b = 0; };

struct Loader {
    std::string sceneDir;
    std::string err;
    gb_scene* out = nullptr;

    std::vector<Geometry> geometries;
    std::map<std::string, int> geometryByName;
    std::map<std::string, int> colorTextures; // name -> index into out->textures
    std::map<std::string, int> floatTextures;
    std::map<std::string, int> materialByName;
    std::vector<ModelDef> models;
    std::map<std::string, PrimitiveDef> primitiveByName;

    struct InstanceDef { Transform toWorld; int model; };
    std::vector<InstanceDef> instances;

    // insert-if-absent, the std::map::insert semantics of SceneCache::addX
    template <typename M, typename V>
    static void addFirst(M& m, const std::string& k, const V& v) { m.insert(std::make_pair(k, v)); }

    std::string resolvePath(const std::string& f) const { // SceneCache::resolvePath
        if (!f.empty() && (f[0] == '/' || (f.size() > 1 && f[1] == ':'))) return f;
        return sceneDir + "/" + f;
    }

    int addMaterial(const std::string& name, const gb_material& m) {
        out->materials.push_back(m);
        int id = (int)out->materials.size() - 1;
        addFirst(materialByName, name, id);
        return id;
    }
    static gb_material lambert(const Color3& kd) {
        gb_material m{};
        m.type = GB_MAT_LAMBERT;
        m.kd[0] = kd.r; m.kd[1] = kd.g; m.kd[2] = kd.b;
        return m;
    }
    // SceneCache::getFloatTexture / getColorTexture (src/GoblinScene.cpp:197-216): index of the named
    // texture, the "error" entry (0.5 / magenta) when it was never defined
    int floatTextureIndex(const std::string& name) const {
        auto it = floatTextures.find(name);
        if (it == floatTextures.end()) {
            std::cerr << "Texture " << name << " not defined!\n";
            return floatTextures.find("error")->second;
        }
        return it->second;
    }
    int colorTextureIndex(const std::string& name) const {
        auto it = colorTextures.find(name);
        if (it == colorTextures.end()) {
            std::cerr << "Texture " << name << " not defined!\n";
            return colorTextures.find("error")->second;
        }
        return it->second;
    }
    int addTexture(const gb_texture& t) {
        out->textures.push_back(t);
        return (int)out->textures.size() - 1;
    }
    static gb_texture constantTexture(bool isFloat, float a, float b = 0.0f, float c = 0.0f) {
        gb_texture t{};
        t.type = GB_TEX_CONSTANT;
        t.is_float = isFloat ? 1 : 0;
        t.value[0] = a; t.value[1] = b; t.value[2] = c;
        t.child[0] = t.child[1] = -1;
        return t;
    }
    // A material slot: the constant folded in, or a reference (1 + index) to a procedural texture.
    struct ColorSlot { Color3 c; int tex = 0; };
    ColorSlot getColorTexture(const std::string& name) const {
        const int i = colorTextureIndex(name);
        const gb_texture& t = out->textures[i];
        ColorSlot s;
        if (t.type == GB_TEX_CONSTANT) s.c = Color3{t.value[0], t.value[1], t.value[2]};
        else s.tex = i + 1;
        return s;
    }
    struct FloatSlot { float v = 0.0f; int tex = 0; };
    FloatSlot getFloatTexture(const std::string& name) const {
        const int i = floatTextureIndex(name);
        const gb_texture& t = out->textures[i];
        FloatSlot s;
        if (t.type == GB_TEX_CONSTANT) s.v = t.value[0];
        else s.tex = i + 1;
        return s;
    }
    int getMaterial(const std::string& name) const {
        auto it = materialByName.find(name);
        if (it == materialByName.end()) {
            std::cerr << "Material " << name << " not defined!\n";
            return materialByName.find("error")->second;
        }
        return it->second;
    }
    int getGeometry(const std::string& name) const {
        auto it = geometryByName.find(name);
        if (it == geometryByName.end()) {
            std::cerr << "Geometry " << name << " not defined!\n";
            return geometryByName.find("error")->second;
        }
        return it->second;
    }
    PrimitiveDef getPrimitive(const std::string& name) const {
        auto it = primitiveByName.find(name);
        if (it == primitiveByName.end()) {
            std::cerr << "Primitive " << name << " not defined!\n";
            return primitiveByName.find("error")->second;
        }
        return it->second;
    }

    int addSphere(float r) {
        Geometry g;
        g.kind = GB_GEOM_SPHERE;
        g.radius = r;
        g.bound = makeBBox(Vec3(r, r, r), Vec3(-r, -r, -r)); // Sphere::getObjectBound
        g.area = 4.0f * kPi * r * r;
        geometries.push_back(std::move(g));
        return (int)geometries.size() - 1;
    }
    int addDisk(float r) {
        Geometry g;
        g.kind = GB_GEOM_DISK;
        g.radius = r;
        g.bound = makeBBox(Vec3(r, r, 0.0f), Vec3(-r, -r, 0.0f)); // Disk::getObjectBound
        g.area = kPi * r * r;
        geometries.push_back(std::move(g));
        return (int)geometries.size() - 1;
    }
    int addMesh(const std::string& file) {
        Geometry g;
        g.kind = GB_GEOM_MESH;
        g.radius = 0.0f;
        std::string e;
        if (!loadObjMesh(resolvePath(file), &g.mesh, &e)) {
            // the reference prints and carries on with an empty mesh
            std::cerr << "Error loading mesh: " << e << std::endl;
        }
        g.bound = g.mesh.bound;
        g.area = g.mesh.area;
        geometries.push_back(std::move(g));
        return (int)geometries.size() - 1;
    }

    int addModel(int geometry, int material, int areaLight, bool lens) {
        ModelDef m;
        m.geometry = geometry;
        m.material = material;
        m.areaLight = areaLight;
        m.isCameraLens = lens;
        models.push_back(m);
        return (int)models.size() - 1;
    }

    void initDefault() { // SceneCache::initDefault
        Color3 magenta{1.0f, 0.0f, 1.0f};
        colorTextures["error"] = addTexture(constantTexture(false, 1.0f, 0.0f, 1.0f));
        floatTextures["error"] = addTexture(constantTexture(true, 0.5f));
        int errMat = addMaterial("error", lambert(magenta));
        int errGeo = addSphere(1.0f);
        geometryByName["error"] = errGeo;
        int errModel = addModel(errGeo, errMat, -1, false);
        PrimitiveDef pd;
        pd.isInstance = false;
        pd.model = errModel;
        primitiveByName["error"] = pd;
    }

    // --- flattening -----------------------------------------------------
    bool flattenGeometry(Geometry& g) {
        if (g.flattened) return true;
        g.flattened = true;
        if (g.kind != GB_GEOM_MESH) return true;
        const MeshData& m = g.mesh;
        g.vertOffset = (uint32_t)(out->vertPos.size() / 3);
        g.vertCount = (uint32_t)m.numVerts();
        g.triOffset = (uint32_t)(out->triIndex.size() / 3);
        g.triCount = (uint32_t)m.numTris();
        out->vertPos.insert(out->vertPos.end(), m.pos.begin(), m.pos.end());
        out->vertNrm.insert(out->vertNrm.end(), m.nrm.begin(), m.nrm.end());
        out->vertUv.insert(out->vertUv.end(), m.uv.begin(), m.uv.end());
        out->triIndex.insert(out->triIndex.end(), m.idx.begin(), m.idx.end());
        // per-triangle object bounds: Triangle::getObjectBound
        std::vector<BBox> boxes(m.numTris());
        for (size_t t = 0; t < m.numTris(); ++t) {
            BBox b;
            for (int k = 0; k < 3; ++k) {
                const float* p = &m.pos[3 * m.idx[3 * t + k]];
                b.expand(Vec3(p[0], p[1], p[2]));
            }
            boxes[t] = b;
        }
        BuiltBVH bvh;
        buildBVH(boxes, &bvh, (BvhMethod)out->bvhMethod);
        g.nodeOffset = (uint32_t)out->modelNodes.size();
        g.nodeCount = (uint32_t)bvh.nodes.size();
        out->modelNodes.insert(out->modelNodes.end(), bvh.nodes.begin(), bvh.nodes.end());
        out->modelOrder.insert(out->modelOrder.end(), bvh.order.begin(), bvh.order.end());
        out->modelDepth = std::max(out->modelDepth, bvh.maxDepth);
        return true;
    }

    int flattenModel(int modelIndex) {
        ModelDef& md = models[modelIndex];
        if (md.flatIndex >= 0) return md.flatIndex;
        Geometry& g = geometries[md.geometry];
        flattenGeometry(g);
        gb_model m{};
        m.kind = g.kind;
        m.radius = g.radius;
        m.material = md.material;
        m.area_light = md.areaLight;
        m.node_offset = g.nodeOffset; m.node_count = g.nodeCount;
        m.tri_offset = g.triOffset; m.tri_count = g.triCount;
        m.vert_offset = g.vertOffset; m.vert_count = g.vertCount;
        m.has_normal = g.mesh.hasNormal ? 1 : 0;
        m.has_uv = g.mesh.hasUv ? 1 : 0;
        m.is_camera_lens = md.isCameraLens ? 1 : 0;
        m.bound[0] = g.bound.pMin.x; m.bound[1] = g.bound.pMin.y; m.bound[2] = g.bound.pMin.z;
        m.bound[3] = g.bound.pMax.x; m.bound[4] = g.bound.pMax.y; m.bound[5] = g.bound.pMax.z;
        out->models.push_back(m);
        md.flatIndex = (int)out->models.size() - 1;
        return md.flatIndex;
    }

    static void store3x4(const Mat4& m, float* dst) {
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) dst[4 * r + c] = m.m[r][c];
    }

    // --- loader steps ---------------------------------------------------
    bool createRenderer(const JsonValue& root) {
        ParamSet s;
        if (const JsonValue* v = root.find("render_setting")) s.parse(*v);
        std::string method = s.getString("render_method", "path_tracing");
        out->methodName = method;
        gb_render_setting& rs = out->setting;
        rs.spp = s.getInt("sample_per_pixel", 1);
        out->threadNum = s.getInt("thread_num", 0);
        rs.max_ray_depth = std::max(1, s.getInt("max_ray_depth", 5));
        rs.ao_sample_num = s.getInt("ao_sample_num", 25);
        rs.gpu_num = std::max(0, s.getInt("gpu_num", 0));
        rs.seed = std::max(0, s.getInt("seed", 0));
        {
            // AORenderer::querySampleQuota -> SampleQuota::requestTwoDQuota rounds the ray count
            // up to a perfect square (src/GoblinSampler.cpp:29-33, src/GoblinUtils.h:126-132)
            int root = (int)std::ceil(std::sqrt((float)std::max(rs.ao_sample_num, 0)));
            rs.ao_sample_num = root * root;
        }
        if (method == "ao") rs.method = GB_METHOD_AO;
        else if (method == "whitted" || method == "light_tracing" || method == "bdpt" || method == "sppm") {
            // other integrators are outside the accelerated path; the caller
            // (g_ray --method) may override, otherwise rendering refuses
            rs.method = -1;
        } else rs.method = GB_METHOD_PATH_TRACING;
        return true;
    }

    void buildFilter(const JsonValue& cameraCtx) {
        ParamSet p;
        if (const JsonValue* v = cameraCtx.find("filter")) p.parse(*v);
        std::string type = p.getString("type");
        Vec2 one; one.x = 1.0f; one.y = 1.0f;
        Vec2 w = p.getVector2("width", one);
        gb_film_desc& film = out->film;
        film.filter_width[0] = w.x;
        film.filter_width[1] = w.y;
        const int N = 16; // FILTER_TABLE_WIDTH
        float deltaX = w.x / N, deltaY = w.y / N;
        auto fill = [&](auto evaluate, float normalizeTerm) { // FilterTable ctor, GoblinFilm.cpp:10-27
            size_t index = 0;
            for (int y = 0; y < N; ++y) {
                float fy = y * deltaY;
                for (int x = 0; x < N; ++x) {
                    float fx = x * deltaX;
                    film.filter_table[index++] = evaluate(fx, fy) / normalizeTerm;
                }
            }
        };
        if (type == "box") {
            fill([](float, float) { return 1.0f; }, 4.0f * w.x * w.y);
        } else if (type == "triangle") {
            fill([&](float x, float y) { return std::max(0.0f, w.x - fabsf(x)) * std::max(0.0f, w.y - fabsf(y)); },
                w.x * w.x * w.y * w.y);
        } else if (type == "mitchell") {
            float B = p.getFloat("b", 2.0f), C = p.getFloat("c", 2.0f);
            float invX = 1.0f / w.x, invY = 1.0f / w.y;
            auto mitchell = [&](float x) { // GoblinFilter.cpp:80-92
                x = (float)fabs(2.0f * x);
                float m;
                if (x > 1.0f) {
                    m = ((-B - 6 * C) * x * x * x + (6 * B + 30 * C) * x * x + (-12 * B - 48 * C) * x +
                        (8 * B + 24 * C)) / 6.0f;
                } else {
                    m = ((12 - 9 * B - 6 * C) * x * x * x + (-18 + 12 * B + 6 * C) * x * x + (6 - 2 * B)) / 6.0f;
                }
                return m;
            };
            float norm = 4.0f * ((12 - 9 * B - 6 * C) / 4 + (-18 + 12 * B + 6 * C) / 3 + (6 - 2 * B) +
                15 * (-B - 6 * B) / 4 + 7 * (6 * B + 30 * C) / 3 + 3 * (-12 * B - 48 * C) / 2 +
                (8 * B + 24 * C)) / 6.0f;
            fill([&](float x, float y) { return mitchell(x * invX) * mitchell(y * invY); }, norm);
        } else { // gaussian, also the fallback for unknown types
            float alpha = p.getFloat("falloff", 2.0f);
            float expX = expf(-alpha * w.x * w.x), expY = expf(-alpha * w.y * w.y);
            auto gaussian = [&](float v, float base) { return std::max(0.0f, expf(-alpha * v * v) - base); };
            // GaussianFilter::getNormalizeTerm: 20 x 20 Riemann sum
            size_t step = 20;
            float dX = w.x / static_cast<float>(step), dY = w.y / static_cast<float>(step);
            float norm = 0.0f;
            for (size_t i = 0; i < step; ++i) {
                for (size_t j = 0; j < step; ++j) {
                    norm += 4.0f * dX * dY * gaussian(i * dX, expX) * gaussian(j * dY, expY);
                }
            }
            fill([&](float x, float y) { return gaussian(x, expX) * gaussian(y, expY); }, norm);
        }
    }

    bool createCamera(const JsonValue& root, const std::string& defaultOutput) {
        ParamSet cp;
        JsonValue empty;
        empty.type = JsonValue::Object;
        const JsonValue* cc = root.find("camera");
        if (!cc) cc = &empty;
        cp.parse(*cc);
        std::string type = cp.getString("type");
        float lensRadius = cp.getFloat("lens_radius");
        if (lensRadius != 0.0f) {
            // the lens disk is an intersectable black-Lambert instance, first
            // in the instance list (GoblinContextLoader.cpp:141-164)
            int geo = addDisk(lensRadius);
            addFirst(geometryByName, type + "_lens_geom", geo);
            int mat = addMaterial(type + "_lens_material", lambert(Color3{0, 0, 0}));
            int model = addModel(geo, mat, -1, true);
            PrimitiveDef pd;
            pd.model = model;
            addFirst(primitiveByName, type + "_lens_model", pd);
            InstanceDef inst;
            inst.toWorld = getTransform(cp);
            inst.model = model;
            instances.push_back(inst);
        }
        // film
        ParamSet fp;
        if (const JsonValue* v = cc->find("film")) fp.parse(*v);
        out->outputPath = fp.hasString("file") ? fp.getString("file") : defaultOutput;
        Vec2 defRes; defRes.x = 512; defRes.y = 512;
        Vec2 res = fp.getVector2("resolution", defRes);
        gb_film_desc& film = out->film;
        film.xres = static_cast<int>(res.x);
        film.yres = static_cast<int>(res.y);
        if (film.xres <= 0 || film.yres <= 0) { err = "film resolution must be positive"; return false; }
        Vec4 defCrop; defCrop.x = 0; defCrop.y = 1; defCrop.z = 0; defCrop.w = 1;
        Vec4 crop = fp.getVector4("crop", defCrop);
        film.xstart = (int)ceil(film.xres * crop.x); // Film ctor, GoblinFilm.cpp:103-106
        film.xcount = std::max(1, (int)ceil(film.xres * crop.y) - film.xstart);
        film.ystart = (int)ceil(film.yres * crop.z);
        film.ycount = std::max(1, (int)ceil(film.yres * crop.w) - film.ystart);
        film.tone_mapping = fp.getBool("tone_mapping") ? 1 : 0; // createImageFilm, GoblinFilm.cpp:207-209
        film.bloom_radius = fp.getFloat("bloom_radius");
        film.bloom_weight = fp.getFloat("bloom_weight");
        buildFilter(*cc);
        float xw = film.filter_width[0], yw = film.filter_width[1];
        film.sx0 = (int)floor(film.xstart + 0.5f - xw); // Film::getSampleRange
        film.sx1 = (int)floor(film.xstart + 0.5f + film.xcount + xw);
        film.sy0 = (int)floor(film.ystart + 0.5f - yw);
        film.sy1 = (int)floor(film.ystart + 0.5f + film.ycount + yw);
        // camera (createPerspectiveCamera + PerspectiveCamera ctor)
        gb_camera& cam = out->camera;
        Vec3 pos = cp.getVector3("position");
        Quat q = getQuaternion(cp);
        float fov = radians(cp.getFloat("fov", 60.0f));
        cam.position[0] = pos.x; cam.position[1] = pos.y; cam.position[2] = pos.z;
        cam.orientation[0] = q.w; cam.orientation[1] = q.x; cam.orientation[2] = q.y; cam.orientation[3] = q.z;
        float aspect = static_cast<float>(film.xres) / static_cast<float>(film.yres);
        // matrixPerspectiveLHD3D: tan() binds to the double overload
        float yScale = (float)(1.0f / ::tan((double)(fov / 2.0f)));
        float xScale = yScale / aspect;
        cam.proj00 = xScale;
        cam.proj11 = yScale;
        cam.lens_radius = lensRadius;
        cam.focal_distance = cp.getFloat("focal_distance", 1.0f);
        if (type == "orthographic") { // createOrthographicCamera + ctor, src/GoblinCamera.cpp:290-299,390-398
            cam.orthographic = 1;
            cam.lens_radius = 0.0f; // the orthographic camera has no lens model (a lens_radius still adds the lens disk)
            cam.focal_distance = 0.0f;
            cam.film_width = cp.getFloat("film_width", 35.0f);
            cam.film_height = cam.film_width / aspect;
            cam.proj00 = 2.0f / cam.film_width; // matrixOrthoLHD3D
            cam.proj11 = 2.0f / cam.film_height;
        }
        return true;
    }

    bool createGeometries(const JsonValue& root) {
        const JsonValue* list = root.find("geometries");
        if (!list || !list->isArray()) return true;
        for (const JsonValue& g : list->arr) {
            ParamSet p(g);
            std::string type = p.getString("type"), name = p.getString("name");
            int id;
            if (type == "mesh") id = addMesh(p.getString("file"));
            else if (type == "disk") id = addDisk(p.getFloat("radius", 1.0f));
            else id = addSphere(p.getFloat("radius", 1.0f)); // "sphere" and the fallback
            addFirst(geometryByName, name, id);
        }
        return true;
    }

    // getTextureMapping, src/GoblinTexture.cpp:600-615
    void readMapping(const ParamSet& p, gb_texture* t) {
        std::string type = p.getString("mapping", "uv");
        t->mapping = GB_MAPPING_UV;
        t->map_scale[0] = t->map_scale[1] = 1.0f;
        t->map_offset[0] = t->map_offset[1] = 0.0f;
        if (type == "uv") {
            Vec2 one; one.x = 1.0f; one.y = 1.0f;
            Vec2 zero; zero.x = 0.0f; zero.y = 0.0f;
            Vec2 sc = p.getVector2("scale", one), of = p.getVector2("offset", zero);
            t->map_scale[0] = sc.x; t->map_scale[1] = sc.y;
            t->map_offset[0] = of.x; t->map_offset[1] = of.y;
        } else if (type == "spherical") {
            t->mapping = GB_MAPPING_SPHERICAL;
            Transform toTex = getTransform(p);
            store3x4(toTex.matrix, t->to_tex); // SphericalMapping::pointToST applies mToTex.onPoint
        } else {
            std::cerr << "undefined mapping type " << type << std::endl;
        }
    }

    // ImageTexture<T>::getMIPMap + convertTexel (src/GoblinTexture.cpp:454-503): one pyramid per
    // (format, file, gamma, channel); max_anisotropy only matters to the lookup.
    std::map<std::string, std::pair<int, int>> imagePyramids; // key -> (first level, level count)
    void loadImagePyramid(const std::string& filePath, bool isFloat, float gamma, int channel, gb_texture* t) {
        std::string key = (isFloat ? "f|" : "c|") + filePath + "|" + std::to_string(channel) + "|";
        key.append(reinterpret_cast<const char*>(&gamma), 4);
        auto it = imagePyramids.find(key);
        if (it == imagePyramids.end()) {
            int w = 0, h = 0;
            std::vector<float> rgba;
            std::string ierr;
            if (!loadEXR(filePath, &w, &h, &rgba, &ierr)) {
                std::cerr << ierr << std::endl << "error loading image file " << filePath << std::endl;
                w = h = 1;
                rgba = {1.0f, 0.0f, 1.0f, 1.0f}; // Color::Magenta
            }
            auto powG = [&](float v) { return (float)::pow((double)v, (double)gamma); }; // pow(float, float) binds to the double overload there
            for (size_t i = 0; i < (size_t)w * h; ++i) {
                float* c = &rgba[4 * i];
                if (isFloat) { // ImageTexture<float>::convertTexel: always through pow
                    float v = channel == 4 ? 0.212671f * c[0] + 0.715160f * c[1] + 0.072169f * c[2] : c[channel];
                    v = powG(v);
                    c[0] = c[1] = c[2] = v;
                    c[3] = 1.0f;
                } else { // ImageTexture<Color>::convertTexel
                    float r = c[0], g = c[1], b = c[2], a = c[3];
                    if (channel != 4) { r = g = b = c[channel]; a = 1.0f; } // Color(float): alpha 1
                    c[0] = gamma == 1.0f ? r : powG(r);
                    c[1] = gamma == 1.0f ? g : powG(g);
                    c[2] = gamma == 1.0f ? b : powG(b);
                    c[3] = a;
                }
            }
            std::vector<MipLevel> pyramid;
            buildMipmap(std::move(rgba), w, h, &pyramid);
            const int first = (int)out->imageLevels.size();
            for (const MipLevel& l : pyramid) {
                gb_image_level gl{};
                gl.width = l.width; gl.height = l.height;
                gl.texel_offset = out->imageTexels.size() / 4;
                out->imageLevels.push_back(gl);
                out->imageTexels.insert(out->imageTexels.end(), l.rgba.begin(), l.rgba.end());
            }
            it = imagePyramids.emplace(key, std::make_pair(first, (int)pyramid.size())).first;
        }
        t->first_level = it->second.first;
        t->n_levels = it->second.second;
    }

    // createTextures, src/GoblinContextLoader.cpp:246-303: file order, children resolved by name at
    // creation (so only earlier textures are visible), first definition of a name wins
    bool createTextures(const JsonValue& root) {
        const JsonValue* list = root.find("textures");
        if (!list || !list->isArray()) return true;
        for (const JsonValue& tj : list->arr) {
            ParamSet p(tj);
            std::string type = p.getString("type"), name = p.getString("name");
            std::string format = p.getString("format", "color");
            if (format != "float" && format != "color") {
                std::cerr << "unrecognize texture format" << format << std::endl;
                continue;
            }
            const bool isFloat = format == "float";
            gb_texture t{};
            if (type == "image") { // createFloat/ColorImageTexture + ImageTexture::getMIPMap
                t.type = GB_TEX_IMAGE;
                t.is_float = isFloat ? 1 : 0;
                t.child[0] = t.child[1] = -1;
                readMapping(p, &t);
                std::string filePath = resolvePath(p.getString("file"));
                std::string filterStr = p.getString("filter", "nearest");
                if (filterStr == "nearest") t.image_filter = GB_FILTER_NEAREST;
                else if (filterStr == "bilinear") t.image_filter = GB_FILTER_BILINEAR;
                else if (filterStr == "trilinear") t.image_filter = GB_FILTER_TRILINEAR;
                else if (filterStr == "EWA") t.image_filter = GB_FILTER_EWA;
                else { std::cerr << "unrecognize filter: " << filterStr << std::endl; t.image_filter = GB_FILTER_NEAREST; }
                std::string addressStr = p.getString("address", "repeat");
                if (addressStr == "repeat") t.address_mode = GB_ADDRESS_REPEAT;
                else if (addressStr == "clamp") t.address_mode = GB_ADDRESS_CLAMP;
                else if (addressStr == "border") t.address_mode = GB_ADDRESS_BORDER;
                else { std::cerr << "unrecognize address mode: " << addressStr << std::endl; t.address_mode = GB_ADDRESS_REPEAT; }
                float gamma = p.getFloat("gamma", 1.0f);
                std::string channelStr = p.getString("channel", "All");
                int channel = 4; // ChannelAll
                if (channelStr == "R") channel = 0;
                else if (channelStr == "G") channel = 1;
                else if (channelStr == "B") channel = 2;
                else if (channelStr == "A") channel = 3;
                else if (channelStr != "All") std::cerr << "unrecognize channel: " << channelStr << std::endl;
                // the colour variant never forwards max_anisotropy (src/GoblinTexture.cpp:739-744): default 10
                t.max_anisotropy = isFloat ? p.getFloat("max_anisotropy", 10.0f) : 10.0f;
                loadImagePyramid(filePath, isFloat, gamma, channel, &t);
            } else if (type == "checkerboard") { // createFloat/ColorCheckerboardTexture
                t.type = GB_TEX_CHECKERBOARD;
                t.is_float = isFloat ? 1 : 0;
                readMapping(p, &t);
                t.child[0] = isFloat ? floatTextureIndex(p.getString("texture1")) : colorTextureIndex(p.getString("texture1"));
                t.child[1] = isFloat ? floatTextureIndex(p.getString("texture2")) : colorTextureIndex(p.getString("texture2"));
                t.filter = p.getBool("filter", false) ? 1 : 0;
            } else if (type == "scale") { // createFloat/ColorScaleTexture: the scale is looked up first
                t.type = GB_TEX_SCALE;
                t.is_float = isFloat ? 1 : 0;
                std::string textureName = p.getString("texture"), scaleName = p.getString("scale");
                t.child[1] = floatTextureIndex(scaleName);
                t.child[0] = isFloat ? floatTextureIndex(textureName) : colorTextureIndex(textureName);
            } else if (isFloat) { // "constant" and the fallback
                t = constantTexture(true, p.getFloat("float", 0.5f));
            } else {
                Vec3 c = p.getVector3("color");
                t = constantTexture(false, c.x, c.y, c.z);
            }
            const int id = addTexture(t);
            addFirst(isFloat ? floatTextures : colorTextures, name, id);
        }
        return true;
    }

    bool createMaterials(const JsonValue& root) {
        const JsonValue* list = root.find("materials");
        if (!list || !list->isArray()) return true;
        for (const JsonValue& mj : list->arr) {
            ParamSet p(mj);
            std::string type = p.getString("type"), name = p.getString("name");
            gb_material m{};
            if (type == "blinn") { // createBlinnMaterial, src/GoblinMaterial.cpp:834-854
                m.type = GB_MAT_BLINN;
                ColorSlot kg = getColorTexture(p.getString("Kg"));
                m.kd[0] = kg.c.r; m.kd[1] = kg.c.g; m.kd[2] = kg.c.b;
                m.kd_tex = kg.tex;
                FloatSlot ex = getFloatTexture(p.getString("exponent"));
                m.exponent = ex.v;
                m.exponent_tex = ex.tex;
                m.eta = p.getFloat("index", 1.5f);
                m.k = p.getFloat("k", -1.0f);
                m.fresnel = m.k > 0.0f ? GB_FRESNEL_CONDUCTOR : GB_FRESNEL_DIELECTRIC;
            } else if (type == "subsurface") {
                err = "material '" + name + "' of type '" + type + "' is outside the accelerated path";
                return false;
            } else if (type == "mask") { // createMaskMaterial, src/GoblinMaterial.cpp:929-950
                FloatSlot alpha;
                alpha.v = 1.0f;
                if (p.hasString("alpha")) alpha = getFloatTexture(p.getString("alpha"));
                else std::cout << "no feed in alpha" << std::endl;
                ColorSlot tc;
                tc.c = Color3{1.0f, 1.0f, 1.0f};
                if (p.hasString("transparent_color")) tc = getColorTexture(p.getString("transparent_color"));
                else std::cout << "no feed in tr color" << std::endl;
                m = out->materials[getMaterial(p.getString("material"))]; // the masked material's record
                if (m.mask) { err = "material '" + name + "': a mask around a mask is outside the accelerated path"; return false; }
                m.mask = 1;
                m.alpha = alpha.v;
                m.alpha_tex = alpha.tex;
                m.transparent_color[0] = tc.c.r; m.transparent_color[1] = tc.c.g; m.transparent_color[2] = tc.c.b;
                m.transparent_tex = tc.tex;
            } else if (type == "transparent") {
                m.type = GB_MAT_TRANSPARENT;
                ColorSlot kr = getColorTexture(p.getString("Kr")), kt = getColorTexture(p.getString("Kt"));
                m.kd[0] = kr.c.r; m.kd[1] = kr.c.g; m.kd[2] = kr.c.b;
                m.kt[0] = kt.c.r; m.kt[1] = kt.c.g; m.kt[2] = kt.c.b;
                m.kd_tex = kr.tex;
                m.kt_tex = kt.tex;
                m.eta = p.getFloat("index", 1.5f);
            } else if (type == "mirror") {
                m.type = GB_MAT_MIRROR;
                ColorSlot kr = getColorTexture(p.getString("Kr"));
                m.kd[0] = kr.c.r; m.kd[1] = kr.c.g; m.kd[2] = kr.c.b;
                m.kd_tex = kr.tex;
                m.eta = p.getFloat("index", 0.8f);
                m.k = p.getFloat("k", 6.0f);
            } else { // "lambert" and the fallback
                ColorSlot kd = getColorTexture(p.getString("Kd"));
                m = lambert(kd.c);
                m.kd_tex = kd.tex;
            }
            if (type != "mask") { // getBumpShaders, src/GoblinMaterial.cpp:813-824 (a mask defers to the masked material)
                if (p.hasString("bumpmap")) m.bump_tex = floatTextureIndex(p.getString("bumpmap")) + 1;
                if (p.hasString("normalmap")) m.normal_tex = colorTextureIndex(p.getString("normalmap")) + 1;
            }
            addMaterial(name, m);
        }
        return true;
    }

    bool createPrimitives(const JsonValue& root) {
        const JsonValue* list = root.find("primitives");
        if (!list || !list->isArray()) return true;
        for (const JsonValue& pj : list->arr) {
            ParamSet p(pj);
            std::string type = p.getString("type"), name = p.getString("name");
            PrimitiveDef pd;
            if (type == "instance") {
                PrimitiveDef target = getPrimitive(p.getString("model"));
                if (target.isInstance) {
                    err = "instance '" + name + "' refers to another instance; nested instancing is "
                        "outside the accelerated path";
                    return false;
                }
                InstanceDef inst;
                inst.toWorld = getTransform(p);
                inst.model = target.model;
                instances.push_back(inst);
                pd.isInstance = true;
            } else { // "model" and the fallback
                int geo = getGeometry(p.getString("geometry"));
                int mat = getMaterial(p.getString("material"));
                // "area_light" can never resolve here: area lights are created
                // after primitives, so the lookup yields the null error entry
                if (p.hasString("area_light")) {
                    std::cerr << "Area Light " << p.getString("area_light") << " not defined!\n";
                }
                pd.model = addModel(geo, mat, -1, p.getBool("is_camera_lens"));
            }
            addFirst(primitiveByName, name, pd);
        }
        return true;
    }

    struct PendingLight { gb_light l; Transform xf; int geometry = -1; int model = -1; Color3 averageRadiance; };
    std::vector<PendingLight> lights;

    static Vec3 lightAxis(const Vec3& dir, Transform* xf) { // Light::setOrientation + onVector(UnitZ)
        Vec3 xAxis, yAxis;
        coordinateAxises(dir, &xAxis, &yAxis);
        float R[3][3] = {{xAxis.x, yAxis.x, dir.x}, {xAxis.y, yAxis.y, dir.y}, {xAxis.z, yAxis.z, dir.z}};
        xf->orientation = quatFromMatrix3(R);
        xf->update();
        return xf->onVector(Vec3(0.0f, 0.0f, 1.0f));
    }

    bool createLights(const JsonValue& root) {
        const JsonValue* list = root.find("lights");
        if (!list || !list->isArray()) return true;
        for (const JsonValue& lj : list->arr) {
            ParamSet p(lj);
            std::string type = p.getString("type"), name = p.getString("name");
            PendingLight pl;
            gb_light& l = pl.l;
            std::memset(&l, 0, sizeof l);
            l.instance = -1;
            l.model = -1;
            if (type == "ibl") { // createImageBasedLight + ImageBasedLight ctor, src/GoblinLight.cpp:464-506,681-691
                l.type = GB_LIGHT_IBL;
                std::string filePath = resolvePath(p.getString("file"));
                Vec3 filter = p.getVector3("filter");
                // default orientation faces the centre of the map (spherical coordinates are z-up):
                // rotateX(-PI/2), rotateY(-PI/2), then the light's own orientation in front
                Quat q = quatNormalize(quatMul(quatFromAxisAngle(Vec3(1, 0, 0), -0.5f * kPi), Quat()));
                q = quatNormalize(quatMul(quatFromAxisAngle(Vec3(0, 1, 0), -0.5f * kPi), q));
                pl.xf.orientation = quatMul(getQuaternion(p), q);
                pl.xf.update();
                store3x4(pl.xf.matrix, l.to_world);
                store3x4(pl.xf.inv, l.to_object);
                int w = 0, h = 0;
                std::vector<float> rgba;
                std::string ierr;
                if (!loadEXR(filePath, &w, &h, &rgba, &ierr)) {
                    std::cerr << ierr << std::endl << "errror loading image " << filePath << std::endl;
                    w = h = 1;
                    rgba = {1.0f, 0.0f, 1.0f, 1.0f}; // Color::Magenta
                }
                for (size_t i = 0; i < (size_t)w * h; ++i) { // buffer[i] *= filter
                    rgba[4 * i] *= filter.x; rgba[4 * i + 1] *= filter.y; rgba[4 * i + 2] *= filter.z;
                }
                std::vector<MipLevel> pyramid;
                buildMipmap(std::move(rgba), w, h, &pyramid);
                const int maxLevel = (int)pyramid.size() - 1;
                float avg[4];
                mipLookup(pyramid, maxLevel, 0.0f, 0.0f, avg); // mAverageRadiance
                pl.averageRadiance = Color3{avg[0], avg[1], avg[2]};
                const MipLevel& dl = pyramid[std::max(0, maxLevel - 8)];
                std::vector<float> dist((size_t)dl.width * dl.height);
                for (int i = 0; i < dl.height; ++i) {
                    float sinTheta = (float)::sin((double)(((float)i + 0.5f) / (float)dl.height * kPi));
                    for (int j = 0; j < dl.width; ++j) {
                        const float* c = &dl.rgba[4 * ((size_t)i * dl.width + j)];
                        dist[(size_t)i * dl.width + j] = (0.212671f * c[0] + 0.715160f * c[1] + 0.072169f * c[2]) * sinTheta;
                    }
                }
                std::vector<float> table;
                buildDistribution2D(dist.data(), dl.width, dl.height, &table);
                l.dist_width = dl.width;
                l.dist_height = dl.height;
                l.dist_offset = out->lightDist.size();
                out->lightDist.insert(out->lightDist.end(), table.begin(), table.end());
                l.image_width = pyramid[0].width;
                l.image_height = pyramid[0].height;
                l.image_offset = out->imageTexels.size() / 4;
                out->imageTexels.insert(out->imageTexels.end(), pyramid[0].rgba.begin(), pyramid[0].rgba.end());
            } else if (type == "directional") {
                l.type = GB_LIGHT_DIRECTIONAL;
                Vec3 c = p.getVector3("radiance"), d = p.getVector3("direction");
                l.color[0] = c.x; l.color[1] = c.y; l.color[2] = c.z;
                Vec3 axis = lightAxis(d, &pl.xf);
                l.direction[0] = axis.x; l.direction[1] = axis.y; l.direction[2] = axis.z;
            } else if (type == "spot") {
                l.type = GB_LIGHT_SPOT;
                Vec3 c = p.getVector3("intensity"), pos = p.getVector3("position");
                Vec3 dir = p.hasVector3("target") ? normalize(p.getVector3("target") - pos) : p.getVector3("direction");
                l.cos_theta_max = (float)::cos((double)radians(p.getFloat("theta_max")));
                l.cos_falloff_start = (float)::cos((double)radians(p.getFloat("falloff_start")));
                l.color[0] = c.x; l.color[1] = c.y; l.color[2] = c.z;
                l.position[0] = pos.x; l.position[1] = pos.y; l.position[2] = pos.z;
                pl.xf.position = pos;
                Vec3 axis = lightAxis(normalize(dir), &pl.xf);
                l.direction[0] = axis.x; l.direction[1] = axis.y; l.direction[2] = axis.z;
            } else if (type == "area") {
                l.type = GB_LIGHT_AREA;
                Vec3 c = p.getVector3("radiance");
                l.color[0] = c.x; l.color[1] = c.y; l.color[2] = c.z;
                int geo = getGeometry(p.getString("geometry"));
                const Geometry& g = geometries[geo];
                pl.geometry = geo;
                pl.xf = getTransform(p);
                l.geom_kind = g.kind;
                l.radius = g.radius;
                l.area = 0.0f + g.area; // GeometrySet::mSumArea accumulates from 0
                store3x4(pl.xf.matrix, l.to_world);
                store3x4(pl.xf.inv, l.to_object);
                l.position[0] = pl.xf.position.x; l.position[1] = pl.xf.position.y; l.position[2] = pl.xf.position.z;
                // hidden black-Lambert model + instance carrying the emitter
                // (GoblinContextLoader.cpp:419-441)
                int mat = addMaterial(type + "_" + name + "_material", lambert(Color3{0, 0, 0}));
                int model = addModel(geo, mat, (int)lights.size(), false);
                pl.model = model;
                if (g.kind == GB_GEOM_MESH) {
                    // GeometrySet: one Triangle per face, areas in face order, CDF1D over them
                    // (src/GoblinLight.cpp:289-306, src/GoblinSampler.cpp:312-330)
                    const MeshData& md = g.mesh;
                    const size_t nt = md.numTris();
                    l.area_offset = (uint32_t)out->lightTriArea.size();
                    l.cdf_offset = (uint32_t)out->lightTriCdf.size();
                    float sum = 0.0f;
                    for (size_t t = 0; t < nt; ++t) {
                        const float* a = &md.pos[3 * md.idx[3 * t]];
                        const float* b = &md.pos[3 * md.idx[3 * t + 1]];
                        const float* c = &md.pos[3 * md.idx[3 * t + 2]];
                        Vec3 p0(a[0], a[1], a[2]), p1(b[0], b[1], b[2]), p2(c[0], c[1], c[2]);
                        float area = 0.5f * length(cross(p1 - p0, p2 - p0)); // Triangle::area
                        out->lightTriArea.push_back(area);
                        sum += area;
                    }
                    l.area = sum;
                    std::vector<float> cdf(nt + 1, 0.0f);
                    if (nt) {
                        float dx = 1.0f / nt;
                        for (size_t i = 1; i < nt + 1; ++i) cdf[i] = cdf[i - 1] + out->lightTriArea[l.area_offset + i - 1] * dx;
                        float integral = cdf[nt];
                        for (size_t i = 1; i < nt + 1; ++i) cdf[i] /= integral;
                    }
                    out->lightTriCdf.insert(out->lightTriCdf.end(), cdf.begin(), cdf.end());
                }
                PrimitiveDef pd;
                pd.model = model;
                addFirst(primitiveByName, type + "_" + name + "_model", pd);
                // the instance is created from the light's own params with
                // "model" appended; getString finds the first "model" entry
                PrimitiveDef target = p.hasString("model") ? getPrimitive(p.getString("model")) : pd;
                if (target.isInstance) { err = "area light '" + name + "': nested instancing"; return false; }
                InstanceDef inst;
                inst.toWorld = getTransform(p);
                inst.model = target.model;
                l.instance = (int)instances.size();
                instances.push_back(inst);
            } else { // "point" and the fallback
                l.type = GB_LIGHT_POINT;
                Vec3 c = p.getVector3("intensity"), pos = p.getVector3("position");
                l.color[0] = c.x; l.color[1] = c.y; l.color[2] = c.z;
                l.position[0] = pos.x; l.position[1] = pos.y; l.position[2] = pos.z;
            }
            lights.push_back(pl);
        }
        return true;
    }

    bool buildScene() {
        // instances -> flattened models, matrices, world boxes
        std::vector<BBox> boxes;
        for (const InstanceDef& id : instances) {
            int flat = flattenModel(id.model);
            gb_instance gi{};
            store3x4(id.toWorld.matrix, gi.to_world);
            store3x4(id.toWorld.inv, gi.to_object);
            const gb_model& m = out->models[flat];
            BBox ob;
            ob.pMin = Vec3(m.bound[0], m.bound[1], m.bound[2]);
            ob.pMax = Vec3(m.bound[3], m.bound[4], m.bound[5]);
            BBox wb = id.toWorld.onBBox(ob); // InstancedPrimitive::getAABB
            gi.aabb[0] = wb.pMin.x; gi.aabb[1] = wb.pMin.y; gi.aabb[2] = wb.pMin.z;
            gi.aabb[3] = wb.pMax.x; gi.aabb[4] = wb.pMax.y; gi.aabb[5] = wb.pMax.z;
            gi.model = flat;
            out->instances.push_back(gi);
            boxes.push_back(wb);
        }
        BuiltBVH top;
        buildBVH(boxes, &top, (BvhMethod)out->bvhMethod); // Scene::mBVH
        out->topNodes = top.nodes;
        out->topOrder = top.order;
        out->topDepth = top.maxDepth;
        out->worldBound[0] = top.bound.pMin.x; out->worldBound[1] = top.bound.pMin.y; out->worldBound[2] = top.bound.pMin.z;
        out->worldBound[3] = top.bound.pMax.x; out->worldBound[4] = top.bound.pMax.y; out->worldBound[5] = top.bound.pMax.z;

        // lights: power-proportional pick distribution (Scene ctor + CDF1D::init)
        float worldRadius = length(top.bound.pMax - top.bound.pMin); // BBox::getBoundingSphere
        for (const PendingLight& pl : lights) {
            const gb_light& l = pl.l;
            float pr = 0, pg = 0, pb = 0;
            auto scale3 = [&](float s) { pr = l.color[0] * s; pg = l.color[1] * s; pb = l.color[2] * s; };
            switch (l.type) {
            case GB_LIGHT_POINT: { // 4.0f * PI * mIntensity
                float s = 4.0f * kPi;
                scale3(s);
                break;
            }
            case GB_LIGHT_DIRECTIONAL: { // radius * radius * PI * mRadiance
                float s = worldRadius * worldRadius * kPi;
                scale3(s);
                break;
            }
            case GB_LIGHT_SPOT: { // mIntensity * TWO_PI * (1 - 0.5 * (cosMax + cosFalloff))
                scale3(kTwoPi);
                float s = 1.0f - 0.5f * (l.cos_theta_max + l.cos_falloff_start);
                pr *= s; pg *= s; pb *= s;
                break;
            }
            case GB_LIGHT_IBL: { // mAverageRadiance * PI * (4.0f * PI * radius * radius)
                float s = 4.0f * kPi * worldRadius * worldRadius;
                pr = pl.averageRadiance.r * kPi * s; pg = pl.averageRadiance.g * kPi * s; pb = pl.averageRadiance.b * kPi * s;
                break;
            }
            default: { // area: mLe * PI * worldArea
                const Vec3& sc = pl.xf.scale;
                float worldArea = l.area * (sc.x * sc.y);
                scale3(kPi);
                pr *= worldArea; pg *= worldArea; pb *= worldArea;
                break;
            }
            }
            out->lightPower.push_back(0.212671f * pr + 0.715160f * pg + 0.072169f * pb);
            out->lights.push_back(l);
            if (pl.model >= 0 && l.geom_kind == GB_GEOM_MESH) out->lights.back().model = flattenModel(pl.model);
        }
        size_t n = out->lightPower.size();
        out->lightCdf.assign(n + 1, 0.0f);
        if (n > 0) {
            float dx = 1.0f / n;
            for (size_t i = 1; i < n + 1; ++i) out->lightCdf[i] = out->lightCdf[i - 1] + out->lightPower[i - 1] * dx;
            float integral = out->lightCdf[n];
            for (size_t i = 1; i < n + 1; ++i) out->lightCdf[i] /= integral;
        }
        return true;
    }
};

} // namespace

int loadSceneString(const std::string& json, const std::string& sceneDir,
    const std::string& defaultOutput, gb_scene* out, std::string* error, int bvhMethod) {
    JsonValue root;
    std::string perr;
    if (!parseJson(json, &root, &perr)) {
        if (error) *error = "ill-formed scene JSON: " + perr;
        return GB_ERR_INVALID;
    }
    if (!root.isObject()) {
        if (error) *error = "scene JSON root must be an object";
        return GB_ERR_INVALID;
    }
    *out = gb_scene();
    out->bvhMethod = bvhMethod;
    Loader L;
    L.sceneDir = sceneDir;
    L.out = out;
    L.initDefault();
    bool ok = L.createRenderer(root) && L.createCamera(root, defaultOutput) && L.createGeometries(root) &&
        L.createTextures(root) && L.createMaterials(root) && L.createPrimitives(root) && L.createLights(root) &&
        L.buildScene();
    if (!ok) {
        if (error) *error = L.err;
        return GB_ERR_INVALID;
    }
    return GB_OK;
}

int loadSceneFile(const std::string& filename, gb_scene* out, std::string* error, int bvhMethod) {
    std::ifstream in(filename, std::ios::binary);
    if (!in.is_open()) {
        if (error) *error = "error reading scene file: " + filename;
        return GB_ERR_IO;
    }
    std::stringstream ss;
    ss << in.rdbuf();
    // scene directory and default output path, GoblinContextLoader.cpp:461-484
    std::string sceneDir = ".";
    size_t slash = filename.find_last_of('/');
    if (slash != std::string::npos) sceneDir = filename.substr(0, slash);
    else {
        slash = filename.find_last_of('\\');
        if (slash != std::string::npos) sceneDir = filename.substr(0, slash);
    }
    size_t ext = filename.find_last_of('.');
    std::string defaultOutput;
    if (ext != std::string::npos && ext < filename.length() - 1 && filename[ext + 1] != '/' &&
        filename[ext + 1] != '\\') {
        defaultOutput = filename.substr(0, ext) + ".exr";
    } else {
        defaultOutput = filename + ".exr";
    }
    return loadSceneString(ss.str(), sceneDir, defaultOutput, out, error, bvhMethod);
}

} // namespace gb

void gb_scene::fillDesc(gb_scene_desc* d) const {
    std::memset(d, 0, sizeof *d);
    d->top_nodes = topNodes.data();
    d->n_top_nodes = (uint32_t)topNodes.size();
    d->top_order = topOrder.data();
    d->instances = instances.data();
    d->n_instances = (uint32_t)instances.size();
    d->models = models.data();
    d->n_models = (uint32_t)models.size();
    d->model_nodes = modelNodes.data();
    d->n_model_nodes = modelNodes.size();
    d->model_order = modelOrder.data();
    d->tri_index = triIndex.data();
    d->n_tris = triIndex.size() / 3;
    d->vert_pos = vertPos.data();
    d->vert_nrm = vertNrm.data();
    d->vert_uv = vertUv.data();
    d->n_verts = vertPos.size() / 3;
    d->materials = materials.data();
    d->n_materials = (uint32_t)materials.size();
    d->lights = lights.data();
    d->n_lights = (uint32_t)lights.size();
    d->light_power = lightPower.data();
    d->light_cdf = lightCdf.data();
    d->light_tri_area = lightTriArea.data();
    d->light_tri_cdf = lightTriCdf.data();
    d->n_light_tri_area = (uint32_t)lightTriArea.size();
    d->n_light_tri_cdf = (uint32_t)lightTriCdf.size();
    std::memcpy(d->world_bound, worldBound, sizeof worldBound);
    d->camera = camera;
    d->film = film;
    d->setting = setting;
    d->image_levels = imageLevels.data();
    d->n_image_levels = (uint32_t)imageLevels.size();
    d->image_texels = imageTexels.data();
    d->n_image_texels = imageTexels.size() / 4;
    d->light_dist = lightDist.data();
    d->n_light_dist = lightDist.size();
    d->textures = textures.data();
    d->n_textures = (uint32_t)textures.size();
}
