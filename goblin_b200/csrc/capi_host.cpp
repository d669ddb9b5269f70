// extern "C" entry points that need no GPU: scene loading / flattening, the
// raw BVH builder, image output, error reporting.
#include <cstring>
#include <string>

#include "bvh_builder.h"
#include "capi_util.h"
#include "host_scene.h"
#include "image_io.h"

namespace gb {
static thread_local std::string g_lastError;
void setLastError(const std::string& e) { g_lastError = e; }
int failWith(int code, const std::string& e) {
    g_lastError = e;
    return code;
}
} // namespace gb

extern "C" {

const char* gb_last_error(void) { return gb::g_lastError.c_str(); }
const char* gb_version(void) { return "goblin_b200 0.1 (sm_100a)"; }

static int bvhMethodOf(const gb_load_options* o, int* method) {
    *method = o ? o->bvh_method : GB_BVH_EQUAL_COUNT;
    if (*method < GB_BVH_EQUAL_COUNT || *method > GB_BVH_SAH) return gb::failWith(GB_ERR_INVALID, "unknown bvh_method");
    return GB_OK;
}

int gb_scene_load_json(const char* path, gb_scene** out) { return gb_scene_load_json_ex(path, nullptr, out); }

int gb_scene_load_json_string(const char* json, const char* scene_dir, gb_scene** out) {
    return gb_scene_load_json_string_ex(json, scene_dir, nullptr, out);
}

int gb_scene_load_json_ex(const char* path, const gb_load_options* options, gb_scene** out) {
    if (!path || !out) return gb::failWith(GB_ERR_INVALID, "null argument");
    *out = nullptr;
    int method = 0;
    if (int rc = bvhMethodOf(options, &method)) return rc;
    try {
        gb_scene* s = new gb_scene();
        std::string err;
        int rc = gb::loadSceneFile(path, s, &err, method);
        if (rc != GB_OK) {
            delete s;
            return gb::failWith(rc, err);
        }
        *out = s;
        return GB_OK;
    } catch (const std::exception& e) {
        return gb::failWith(GB_ERR_INVALID, std::string("exception while loading scene: ") + e.what());
    }
}

int gb_scene_load_json_string_ex(const char* json, const char* scene_dir, const gb_load_options* options,
    gb_scene** out) {
    if (!json || !out) return gb::failWith(GB_ERR_INVALID, "null argument");
    *out = nullptr;
    int method = 0;
    if (int rc = bvhMethodOf(options, &method)) return rc;
    try {
        gb_scene* s = new gb_scene();
        std::string err;
        std::string dir = scene_dir ? scene_dir : ".";
        int rc = gb::loadSceneString(json, dir, dir + "/goblin.exr", s, &err, method);
        if (rc != GB_OK) {
            delete s;
            return gb::failWith(rc, err);
        }
        *out = s;
        return GB_OK;
    } catch (const std::exception& e) {
        return gb::failWith(GB_ERR_INVALID, std::string("exception while loading scene: ") + e.what());
    }
}

void gb_scene_destroy(gb_scene* scene) { delete scene; }

int gb_scene_get_desc(const gb_scene* scene, gb_scene_desc* out) {
    if (!scene || !out) return gb::failWith(GB_ERR_INVALID, "null argument");
    scene->fillDesc(out);
    return GB_OK;
}

const char* gb_scene_output_path(const gb_scene* scene) { return scene ? scene->outputPath.c_str() : ""; }

int gb_bvh_build(const float* aabbs, uint32_t n, gb_bvh_node* nodes, uint32_t* n_nodes, uint32_t* order) {
    return gb_bvh_build_method(aabbs, n, GB_BVH_EQUAL_COUNT, nodes, n_nodes, order);
}

int gb_bvh_build_method(const float* aabbs, uint32_t n, int method, gb_bvh_node* nodes, uint32_t* n_nodes,
    uint32_t* order) {
    if ((n && (!aabbs || !nodes || !order)) || !n_nodes) return gb::failWith(GB_ERR_INVALID, "null argument");
    if (method < GB_BVH_EQUAL_COUNT || method > GB_BVH_SAH) return gb::failWith(GB_ERR_INVALID, "unknown bvh method");
    try {
        std::vector<gb::BBox> boxes(n);
        for (uint32_t i = 0; i < n; ++i) {
            boxes[i].pMin = gb::Vec3(aabbs[6 * i], aabbs[6 * i + 1], aabbs[6 * i + 2]);
            boxes[i].pMax = gb::Vec3(aabbs[6 * i + 3], aabbs[6 * i + 4], aabbs[6 * i + 5]);
        }
        gb::BuiltBVH bvh;
        gb::buildBVH(boxes, &bvh, (gb::BvhMethod)method);
        *n_nodes = (uint32_t)bvh.nodes.size();
        if (!bvh.nodes.empty()) std::memcpy(nodes, bvh.nodes.data(), bvh.nodes.size() * sizeof(gb_bvh_node));
        if (n) std::memcpy(order, bvh.order.data(), n * sizeof(uint32_t));
        return GB_OK;
    } catch (const std::exception& e) {
        return gb::failWith(GB_ERR_INVALID, e.what());
    }
}

int gb_write_image(const char* path, const float* rgbw, int xres, int yres) {
    if (!path || !rgbw || xres <= 0 || yres <= 0) return gb::failWith(GB_ERR_INVALID, "bad argument");
    std::string err;
    if (!gb::writeFilm(path, rgbw, xres, yres, &err)) return gb::failWith(GB_ERR_IO, err);
    return GB_OK;
}

int gb_write_rgb(const char* path, const float* rgb, int xres, int yres, int tone_mapping) {
    if (!path || !rgb || xres <= 0 || yres <= 0) return gb::failWith(GB_ERR_INVALID, "bad argument");
    std::string err;
    if (!gb::writeRgb(path, rgb, xres, yres, tone_mapping != 0, &err)) return gb::failWith(GB_ERR_IO, err);
    return GB_OK;
}

} // extern "C"
