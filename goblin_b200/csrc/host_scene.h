// HostScene: the flattened, device-ready description of a Goblin scene, owned
// on the host.  It is what ContextLoader::load + Scene::Scene + Model::Model
// produce in the reference (src/GoblinContextLoader.cpp:447-503,
// src/GoblinScene.cpp:11-27, src/GoblinModel.cpp:10-26), laid out as the
// plain arrays of gb_scene_desc.
#pragma once
#include <string>
#include <vector>

#include "goblin_b200.h"

struct gb_scene {
    std::vector<gb_bvh_node> topNodes;
    std::vector<uint32_t> topOrder;
    std::vector<gb_instance> instances;
    std::vector<gb_model> models;
    std::vector<gb_bvh_node> modelNodes;
    std::vector<uint32_t> modelOrder;
    std::vector<uint32_t> triIndex;
    std::vector<float> vertPos, vertNrm, vertUv;
    std::vector<gb_material> materials;
    std::vector<gb_texture> textures;
    std::vector<gb_light> lights;
    std::vector<float> lightPower, lightCdf;
    std::vector<float> lightTriArea, lightTriCdf; // mesh emitters: per-face areas and their CDF
    std::vector<float> imageTexels, lightDist;    // image based lights: level-0 radiance, CDF2D tables
    std::vector<gb_image_level> imageLevels;      // image textures: MIPMap pyramids in imageTexels
    float worldBound[6] = {0, 0, 0, 0, 0, 0};
    gb_camera camera{};
    gb_film_desc film{};
    gb_render_setting setting{};
    int threadNum = 0;
    int bvhMethod = 0;      // GB_BVH_* every BVH of this scene was built with
    int topDepth = 0;       // deepest level of the top-level BVH
    int modelDepth = 0;     // deepest level over all per-model BVHs
    std::string outputPath; // film "file" or <scene>.exr
    std::string methodName; // render_method as written in the scene

    void fillDesc(gb_scene_desc* d) const;
};

namespace gb {

// Loads and flattens a scene; returns a GB_* status and sets *error.
int loadSceneFile(const std::string& path, gb_scene* out, std::string* error, int bvhMethod = 0);
int loadSceneString(const std::string& json, const std::string& sceneDir,
    const std::string& defaultOutput, gb_scene* out, std::string* error, int bvhMethod = 0);

} // namespace gb
