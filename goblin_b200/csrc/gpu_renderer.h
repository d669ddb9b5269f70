// Host-side mirror of the reference's Renderer / Film / RenderContext surface
// for the accelerated path, written on top of the C ABI only.
//
//   reference                                   here
//   ContextLoader::load(file) -> RenderContext* gb::ContextLoader::load(file)
//   RenderContext::render()                     gb::RenderContext::render()
//     Renderer::preprocess(scene)                 GpuRenderer::preprocess(scene)  (upload)
//     Renderer::render(scene)                     GpuRenderer::render(scene)      (waves + film write)
//   Film::getSampleRange / writeImage           gb::Film::getSampleRange / writeImage
// (src/GoblinRenderContext.h:19-22, src/GoblinRenderer.h:55-57,
//  src/GoblinFilm.cpp:131-138,164-192)
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "goblin_b200.h"

namespace gb {

struct SampleRange { int xStart, xEnd, yStart, yEnd; };

inline void checkRc(int rc, const char* what) {
    if (rc != GB_OK) throw std::runtime_error(std::string(what) + ": " + gb_last_error());
}

class Scene {
public:
    explicit Scene(gb_scene* s) : mScene(s) { gb_scene_get_desc(s, &mDesc); }
    ~Scene() { gb_scene_destroy(mScene); }
    Scene(const Scene&) = delete;
    Scene& operator=(const Scene&) = delete;
    const gb_scene_desc& desc() const { return mDesc; }
    gb_scene_desc& desc() { return mDesc; }
    std::string outputPath() const { return gb_scene_output_path(mScene); }
private:
    gb_scene* mScene;
    gb_scene_desc mDesc;
};
typedef std::shared_ptr<Scene> ScenePtr;

class Film {
public:
    Film(const gb_film_desc& d, const std::string& file) : mDesc(d), mFilename(file),
        mPixels((size_t)d.xres * d.yres * 4, 0.0f) {}
    int getXResolution() const { return mDesc.xres; }
    int getYResolution() const { return mDesc.yres; }
    void getSampleRange(SampleRange& r) const {
        r.xStart = mDesc.sx0; r.xEnd = mDesc.sx1; r.yStart = mDesc.sy0; r.yEnd = mDesc.sy1;
    }
    void clear() { std::fill(mPixels.begin(), mPixels.end(), 0.0f); }
    // Film::mergeTile: add another (r, g, b, weight) buffer into the film
    void merge(const std::vector<float>& rgbw) {
        for (size_t i = 0; i < mPixels.size(); ++i) mPixels[i] += rgbw[i];
    }
    // Film::writeImage (src/GoblinFilm.cpp:164-192): colour / weight, bloom, then the writer
    // (tone mapping for .ppm).  The resolve and the bloom run on `ctx`'s GPU over the merged film.
    // The device film of `ctx` already holds the merged frame (it is where mPixels was downloaded from).
    void writeImage(gb_context* ctx) {
        std::printf("write image to : %s\n", mFilename.c_str());
        checkRc(gb_film_write(ctx, mFilename.c_str()), "writeImage");
    }
    const std::vector<float>& pixels() const { return mPixels; }
    void setFilename(const std::string& f) { mFilename = f; }
private:
    gb_film_desc mDesc;
    std::string mFilename;
    std::vector<float> mPixels;
};

struct RenderStats {
    double seconds = 0.0;
    unsigned long long cameraSamples = 0, raysClosest = 0, raysAny = 0, launches = 0;
};

// The device dispatch that replaces GoblinThreadPool: one context (and one
// host thread) per GPU, each rendering a slice of the per-pixel sample indices
// of a full scene replica; the films are summed afterwards, which is what
// Film::mergeTile does for the reference's per-thread tiles.
class GpuRenderer {
public:
    GpuRenderer(int gpuNum, unsigned long long seed) : mGpuNum(gpuNum), mSeed(seed) {}
    ~GpuRenderer() { for (gb_context* c : mContexts) gb_destroy(c); }

    void preprocess(const ScenePtr& scene) {
        int available = 0;
        checkRc(gb_device_count(&available), "gb_device_count");
        if (available < 1) throw std::runtime_error("no CUDA device: this renderer has no CPU path");
        if (mGpuNum <= 0 || mGpuNum > available) mGpuNum = mGpuNum <= 0 ? 1 : available;
        for (int g = 0; g < mGpuNum; ++g) {
            gb_context* c = nullptr;
            checkRc(gb_create(g, &c), "gb_create");
            mContexts.push_back(c);
            checkRc(gb_upload_scene(c, &scene->desc()), "gb_upload_scene");
        }
        // Film::mergeTile across GPUs is an NCCL all-reduce of the device films (src/GoblinFilm.cpp:140-153)
        if (mGpuNum > 1) checkRc(gb_comm_init_all(mContexts.data(), mGpuNum), "gb_comm_init_all");
    }

    void render(const ScenePtr& scene, Film* film) {
        const gb_render_setting& rs = scene->desc().setting;
        int root = (int)std::ceil(std::sqrt((float)rs.spp)); // roundToSquare
        if (root < 1) root = 1;
        const int sppTotal = root * root;
        const int G = (int)mContexts.size();
        // a second RenderContext::render() starts from an empty film, on the devices and on the host
        film->clear();
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> workers;
        std::vector<std::string> errors(G);
        for (int g = 0; g < G; ++g) {
            workers.emplace_back([&, g]() {
                if (gb_film_clear(mContexts[g]) != GB_OK) { errors[g] = gb_last_error(); return; }
                gb_render_params p{};
                p.seed = mSeed;
                p.spp_total = sppTotal;
                p.spp_begin = (int)((long long)sppTotal * g / G);
                p.spp_end = (int)((long long)sppTotal * (g + 1) / G);
                p.max_ray_depth = 0;
                p.method = -1;
                p.ao_sample_num = 0;
                // with several GPUs the render stays queued: the all-reduce below goes behind it on the same stream
                if (gb_render(mContexts[g], &p) != GB_OK || (G == 1 && gb_synchronize(mContexts[g]) != GB_OK)) {
                    errors[g] = gb_last_error();
                }
            });
        }
        for (auto& w : workers) w.join();
        for (const std::string& e : errors) if (!e.empty()) throw std::runtime_error("gb_render: " + e);
        // Film::mergeTile: the G device films are summed over NVLink; every GPU then holds the frame, GPU 0's is
        // the one that is normalised and written
        if (G > 1) {
            checkRc(gb_film_allreduce_all(mContexts.data(), G), "gb_film_allreduce_all");
            for (int g = 0; g < G; ++g) checkRc(gb_synchronize(mContexts[g]), "gb_synchronize");
        }
        mStats.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::vector<float> tmp(film->pixels().size());
        checkRc(gb_film_download(mContexts[0], tmp.data()), "gb_film_download");
        film->merge(tmp);
        for (int g = 0; g < G; ++g) {
            gb_counters c{};
            gb_get_counters(mContexts[g], &c);
            mStats.cameraSamples += c.camera_samples;
            mStats.raysClosest += c.rays_closest;
            mStats.raysAny += c.rays_any;
            mStats.launches += c.kernel_launches;
        }
        film->writeImage(mContexts[0]);
    }
    const RenderStats& stats() const { return mStats; }
    int gpuNum() const { return mGpuNum; }
private:
    int mGpuNum;
    unsigned long long mSeed;
    std::vector<gb_context*> mContexts;
    RenderStats mStats;
};

class RenderContext {
public:
    RenderContext(std::shared_ptr<GpuRenderer> r, ScenePtr s, std::shared_ptr<Film> f)
        : mRenderer(r), mScene(s), mFilm(f) {}
    void render() {
        mRenderer->preprocess(mScene);
        mRenderer->render(mScene, mFilm.get());
    }
    std::shared_ptr<GpuRenderer> mRenderer;
    ScenePtr mScene;
    std::shared_ptr<Film> mFilm;
};

class ContextLoader {
public:
    // returns nullptr on an unreadable / ill-formed scene, like the reference
    // bvhMethod: GB_BVH_EQUAL_COUNT = the reference's tree; GB_BVH_SAH = the non-parity fast tree
    static RenderContext* load(const std::string& filename, int gpuNum = 1, unsigned long long seed = 1,
        int bvhMethod = GB_BVH_EQUAL_COUNT) {
        gb_scene* s = nullptr;
        gb_load_options opt{};
        opt.bvh_method = bvhMethod;
        if (gb_scene_load_json_ex(filename.c_str(), &opt, &s) != GB_OK) {
            std::fprintf(stderr, "%s\n", gb_last_error());
            return nullptr;
        }
        ScenePtr scene(new Scene(s));
        std::shared_ptr<Film> film(new Film(scene->desc().film, scene->outputPath()));
        // 0 = not given by the caller: the scene's optional render_setting.gpu_num / seed, else 1
        const gb_render_setting& rs = scene->desc().setting;
        if (gpuNum <= 0) gpuNum = rs.gpu_num > 0 ? rs.gpu_num : 1;
        if (seed == 0) seed = rs.seed > 0 ? (unsigned long long)rs.seed : 1ull;
        return new RenderContext(std::make_shared<GpuRenderer>(gpuNum, seed), scene, film);
    }
};

} // namespace gb
