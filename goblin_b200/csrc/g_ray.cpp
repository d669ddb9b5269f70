// g_ray: the reference's CLI (src/g_ray.cpp:7-27), same usage and same
// "render complete in N seconds" line, with the path_tracing / ao integrators
// running on B200 GPUs.  Extra, optional flags after the scene:
//   --method path_tracing|ao   override render_setting.render_method
//                              (examples/bunny.json ships "sppm")
//   --spp N  --depth N         override sample_per_pixel / max_ray_depth
//   --gpus N                   shard the per-pixel samples over N GPUs
//   --seed S  --out FILE       Philox seed, output image (.exr .pfm .ppm)
//   --stats                    print Msamples/s and Mrays/s as a JSON line
//   --accel equal_count|middle|sah
//                              BVH split method: equal_count (default) is the reference's tree,
//                              sah the non-parity fast tree (same image in distribution)
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <memory>

#include "gpu_renderer.h"

int main(int argc, char** argv) {
    if (argc < 2) {
        std::cout << "Usage: g_ray scene.json" << std::endl;
        return 0;
    }
    int gpus = 0, spp = 0, depth = 0;        // 0: take render_setting.gpu_num / seed, else 1
    unsigned long long seed = 0;
    std::string method, outFile, accel = "equal_count";
    bool stats = false;
    for (int i = 2; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--gpus") gpus = atoi(next());
        else if (a == "--spp") spp = atoi(next());
        else if (a == "--depth") depth = atoi(next());
        else if (a == "--seed") seed = strtoull(next(), nullptr, 10);
        else if (a == "--method") method = next();
        else if (a == "--out") outFile = next();
        else if (a == "--stats") stats = true;
        else if (a == "--accel") accel = next();
        else {
            std::cout << "Usage: g_ray scene.json [--method m] [--spp n] [--depth n] [--gpus n] [--seed s] [--out file] [--stats] [--accel equal_count|middle|sah]" << std::endl;
            return 0;
        }
    }
    int bvhMethod = GB_BVH_EQUAL_COUNT;
    if (accel == "middle") bvhMethod = GB_BVH_MIDDLE;
    else if (accel == "sah") bvhMethod = GB_BVH_SAH;
    else if (accel != "equal_count") {
        std::cerr << "g_ray: --accel must be equal_count, middle or sah" << std::endl;
        return 1;
    }
    std::unique_ptr<gb::RenderContext> renderContext(gb::ContextLoader::load(argv[1], gpus, seed, bvhMethod));
    if (renderContext) {
        gb_render_setting& rs = renderContext->mScene->desc().setting;
        if (method == "path_tracing") rs.method = GB_METHOD_PATH_TRACING;
        else if (method == "ao") rs.method = GB_METHOD_AO;
        else if (!method.empty()) {
            std::cerr << "g_ray: --method must be path_tracing or ao" << std::endl;
            return 1;
        }
        if (spp > 0) rs.spp = spp;
        if (depth > 0) rs.max_ray_depth = depth;
        if (!outFile.empty()) renderContext->mFilm->setFilename(outFile);
        if (rs.method != GB_METHOD_PATH_TRACING && rs.method != GB_METHOD_AO) {
            std::cerr << "g_ray: this scene selects a render_method outside the accelerated path; "
                         "pass --method path_tracing (or ao)" << std::endl;
            return 1;
        }
        std::cout << "\nsuccessfully loaded scene, start rendering..." << std::endl;
        time_t beforeRender;
        time(&beforeRender);
        try {
            renderContext->render();
        } catch (const std::exception& e) {
            std::cerr << "g_ray: " << e.what() << std::endl;
            return 1;
        }
        time_t afterRender;
        time(&afterRender);
        double seconds = difftime(afterRender, beforeRender);
        std::cout << "render complete in " << seconds << " seconds!" << std::endl;
        if (stats) {
            const gb::RenderStats& s = renderContext->mRenderer->stats();
            std::printf("{\"seconds\": %.6f, \"gpus\": %d, \"camera_samples\": %llu, \"rays_closest\": %llu, "
                        "\"rays_any\": %llu, \"msamples_per_s\": %.3f, \"mrays_per_s\": %.3f, \"kernel_launches\": %llu}\n",
                s.seconds, renderContext->mRenderer->gpuNum(), s.cameraSamples, s.raysClosest, s.raysAny,
                s.cameraSamples / s.seconds * 1e-6, (s.raysClosest + s.raysAny) / s.seconds * 1e-6, s.launches);
        }
    }
    return 0;
}
