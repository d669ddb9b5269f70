/*
 * goblin_b200.h -- C ABI of the B200-native Goblin path-tracing hot path.
 *
 * The reference (bachi95/Goblin) has no FFI; its only seam is the virtual
 * Renderer chosen by JSON (src/GoblinRenderer.h:50-61, selected in
 * src/GoblinContextLoader.cpp:67-92) and driven from
 * RenderContext::render() (src/GoblinRenderContext.h:19-22).  This header is
 * the thin extern "C" layer that a Goblin maintainer would bind behind that
 * seam: plain structs and pointers, every call returns an
 * int status (0 = ok) and never throws, the caller owns all input arrays and
 * may free them when the call returns, outputs go to caller-allocated buffers,
 * one context per GPU driven by one host thread.  There is no CPU fallback:
 * every compute entry point fails with GB_ERR_CUDA when no device is usable.
 */
#ifndef GOBLIN_B200_H
#define GOBLIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GB_OK 0
#define GB_ERR_INVALID 1  /* bad argument / malformed scene description      */
#define GB_ERR_IO 2       /* unreadable scene / mesh / output file           */
#define GB_ERR_CUDA 3     /* no device, launch failure, out of device memory */
#define GB_ERR_STATE 4    /* call made before its prerequisite (no scene ..) */
#define GB_ERR_LIMIT 5    /* scene exceeds a compiled-in limit (stack depth) */

/* ------------------------------------------------------------------ records */

/* Bit-for-bit the reference's CompactBVHNode (src/GoblinBVH.h:8-30): 24-byte
 * box, first-primitive index (leaf) or second-child index (interior; the first
 * child is node + 1), primitive count (0 = interior), split axis, 2 pad bytes
 * that are zero.  One node is one 32-byte sector. */
typedef struct gb_bvh_node {
    float bmin[3];
    float bmax[3];
    uint32_t offset;
    uint8_t nprims;
    uint8_t axis;
    uint8_t pad[2];
} gb_bvh_node;

/* Ray as the reference carries it (src/GoblinRay.h:9-19). */
typedef struct gb_ray {
    float o[3];
    float d[3];
    float mint;
    float maxt;
} gb_ray;

/* Closest hit.  inst = index of the hit instance in scene order
 * ([camera lens] + JSON "instance" primitives + area-light instances,
 * src/GoblinContextLoader.cpp:161-163,381-383,437-439), prim = triangle index
 * in the mesh's face order (0 for sphere / disk); inst = -1 on a miss.
 * t is the shrunk ray.maxt, eps the reference's 1e-3 * t
 * (src/GoblinTriangle.cpp:80-81). */
typedef struct gb_hit {
    float t;
    float eps;
    int32_t inst;
    int32_t prim;
} gb_hit;

enum { GB_GEOM_MESH = 0, GB_GEOM_SPHERE = 1, GB_GEOM_DISK = 2 };
enum { GB_MAT_LAMBERT = 0, GB_MAT_MIRROR = 1, GB_MAT_TRANSPARENT = 2, GB_MAT_BLINN = 3, GB_MAT_COUNT = 4 };
enum { GB_FRESNEL_DIELECTRIC = 0, GB_FRESNEL_CONDUCTOR = 1 };
enum { GB_LIGHT_POINT = 0, GB_LIGHT_DIRECTIONAL = 1, GB_LIGHT_SPOT = 2, GB_LIGHT_AREA = 3, GB_LIGHT_IBL = 4 };
enum { GB_METHOD_PATH_TRACING = 0, GB_METHOD_AO = 1 };
/* BVH split methods.  EQUAL_COUNT is what every BVH of the reference is built
 * with (src/GoblinModel.cpp:24, src/GoblinScene.cpp:15) and the only parity
 * mode; MIDDLE is the reference's other, unused method
 * (src/GoblinBVH.cpp:124-134); SAH is this library's non-parity "fast" tree
 * (binned surface-area heuristic, same node format, same kernels): closest-hit
 * distances are unchanged, hit ids can differ where two primitives tie. */
enum { GB_BVH_EQUAL_COUNT = 0, GB_BVH_MIDDLE = 1, GB_BVH_SAH = 2 };

/* Model = geometry + material (+ area light) (src/GoblinModel.cpp:10-26).
 * Mesh models own a private BVH over their triangles; the node / order /
 * triangle / vertex ranges index the concatenated arrays of gb_scene_desc. */
typedef struct gb_model {
    int32_t kind;          /* GB_GEOM_*                                        */
    float radius;          /* sphere / disk                                    */
    int32_t material;      /* index into materials                             */
    int32_t area_light;    /* index into lights, -1 if none                    */
    uint32_t node_offset;  /* first BVH node of this model in model_nodes      */
    uint32_t node_count;
    uint32_t tri_offset;   /* first triangle in tri_index / model_order        */
    uint32_t tri_count;
    uint32_t vert_offset;  /* first vertex in vert_pos / vert_nrm / vert_uv    */
    uint32_t vert_count;
    int32_t has_normal;    /* mesh carries vn (src/GoblinPolygonMesh.cpp:130-147) */
    int32_t has_uv;
    int32_t is_camera_lens;
    float bound[6];        /* object-space AABB (Model::getAABB)               */
} gb_model;

/* Instance = transform + model (src/GoblinPrimitive.cpp:99-122).  Row-major
 * 3x4 matrices: the 4x4 of Transform::update (src/GoblinTransform.cpp:182-193)
 * without its constant last row. */
typedef struct gb_instance {
    float to_world[12];
    float to_object[12];
    float aabb[6];         /* world AABB, Transform::onBBox                    */
    int32_t model;
    int32_t pad;
} gb_instance;

typedef struct gb_material {
    int32_t type;          /* GB_MAT_*                                         */
    float kd[3];           /* lambert Kd | mirror / transparent Kr | blinn Kg  */
    float kt[3];           /* transparent Kt                                   */
    float eta;             /* transparent / mirror / blinn index               */
    float k;               /* mirror / blinn-conductor absorption              */
    float exponent;        /* blinn: the (constant) exponent texture's value   */
    int32_t fresnel;       /* blinn: GB_FRESNEL_*                              */
    /* Non-constant textures: 1 + index into gb_scene_desc.textures, 0 = the
     * constant above.  kd_tex replaces kd (Kd / Kr / Kg), kt_tex replaces kt,
     * exponent_tex (a float texture) replaces exponent. */
    int32_t kd_tex;
    int32_t kt_tex;
    int32_t exponent_tex;
    /* Mask material (src/GoblinMaterial.cpp:747-811): the fields above are the MASKED material's;
     * alpha blends it with an index-matched pass-through of colour transparent_color.  A mask makes
     * its primitives "not opaque" for the path tracer's filtered traces (src/GoblinPathtracer.cpp:5-48). */
    int32_t mask;              /* 1 = this record is a Mask around the material described above */
    float alpha;
    float transparent_color[3];
    int32_t alpha_tex;         /* 1 + texture index (a float texture), 0 = the constant alpha    */
    int32_t transparent_tex;   /* 1 + texture index, 0 = the constant transparent_color          */
    /* BumpShaders (src/GoblinMaterial.cpp:221-281), applied to every hit of the material before it is
     * shaded (Scene::intersect -> Material::perturb): 1 + texture index, 0 = none.  Constant textures
     * are referenced too: even a flat bump map replaces the normal by that of the dpdu x dpdv frame. */
    int32_t bump_tex;          /* float texture: height                                          */
    int32_t normal_tex;        /* colour texture: tangent-space normal, 2 c - 1                  */
} gb_material;

/* Procedural textures (src/GoblinTexture.cpp:292-427): constant, checkerboard
 * over two child textures (point-sampled or box-filtered with the primary
 * ray's uv differentials), scale = float texture x texture.  Children are
 * indices of EARLIER entries of the table (textures are created in file
 * order and look their children up by name at creation,
 * src/GoblinContextLoader.cpp:246-303).  Image textures (:431-503) read an
 * OpenEXR file into a MIPMap pyramid. */
enum { GB_TEX_CONSTANT = 0, GB_TEX_CHECKERBOARD = 1, GB_TEX_SCALE = 2, GB_TEX_IMAGE = 3 };
enum { GB_MAPPING_UV = 0, GB_MAPPING_SPHERICAL = 1 };
/* image textures (src/GoblinTexture.h:12-31): "filter" nearest / bilinear / trilinear / EWA,
 * "address" repeat / clamp / border */
enum { GB_FILTER_NEAREST = 0, GB_FILTER_BILINEAR = 1, GB_FILTER_TRILINEAR = 2, GB_FILTER_EWA = 3 };
enum { GB_ADDRESS_REPEAT = 0, GB_ADDRESS_CLAMP = 1, GB_ADDRESS_BORDER = 2 };
/* One level of a MIPMap pyramid (src/GoblinTexture.cpp:39-71) in image_texels. */
typedef struct gb_image_level {
    int32_t width, height;
    uint64_t texel_offset; /* first RGBA texel (float4 units) in image_texels */
} gb_image_level;
typedef struct gb_texture {
    int32_t type;          /* GB_TEX_*                                         */
    int32_t is_float;      /* float texture: value[0] only                     */
    float value[3];        /* constant                                         */
    int32_t child[2];      /* checkerboard: texture1, texture2; scale: texture, scale */
    int32_t filter;        /* checkerboard "filter"                            */
    int32_t mapping;       /* GB_MAPPING_* (checkerboard)                      */
    float map_scale[2];    /* UVMapping                                        */
    float map_offset[2];
    float to_tex[12];      /* SphericalMapping: world -> texture space, 3x4    */
    /* image texture: the pyramid of its (file, gamma, channel) after convertTexel;
     * float images keep their value in every colour channel */
    int32_t image_filter;  /* GB_FILTER_*                                      */
    int32_t address_mode;  /* GB_ADDRESS_*                                     */
    float max_anisotropy;  /* EWA                                              */
    int32_t first_level;   /* index of level 0 in gb_scene_desc.image_levels   */
    int32_t n_levels;
} gb_texture;

typedef struct gb_light {
    int32_t type;          /* GB_LIGHT_*                                       */
    float color[3];        /* intensity (point, spot) | radiance (dir, area)   */
    float position[3];
    float direction[3];    /* directional: getDirection(); spot: cone axis     */
    float cos_theta_max;   /* spot                                             */
    float cos_falloff_start;
    int32_t geom_kind;     /* area: GB_GEOM_SPHERE / GB_GEOM_DISK / GB_GEOM_MESH */
    float radius;          /* area (sphere / disk)                             */
    float area;            /* area: GeometrySet::mSumArea                      */
    float to_world[12];    /* area: light's own Transform                      */
    float to_object[12];
    int32_t instance;      /* area: scene instance that carries the geometry   */
    /* mesh emitters (GeometrySet over the mesh's triangles, src/GoblinLight.cpp:289-343): */
    int32_t model;         /* the emitting mesh model, -1 otherwise            */
    uint32_t area_offset;  /* first per-face area in light_tri_area            */
    uint32_t cdf_offset;   /* first of tri_count + 1 entries in light_tri_cdf  */
    /* image based light (src/GoblinLight.cpp:464-629): to_world / to_object hold
     * its orientation; level 0 of the radiance MIPMap (already multiplied by the
     * light's filter colour) and the CDF2D over luminance x sin(theta) */
    int32_t image_width, image_height;
    uint64_t image_offset; /* first RGBA texel (float4 units) in image_texels  */
    int32_t dist_width, dist_height;
    uint64_t dist_offset;  /* first float of this light's table in light_dist:
                            * dist_height x dist_width function values, the rows'
                            * dist_width + 1 CDF entries each, then the marginal:
                            * dist_height row integrals, dist_height + 1 CDF
                            * entries, 1 integral                               */
} gb_light;

typedef struct gb_camera {
    float position[3];
    float orientation[4];  /* w, x, y, z (src/GoblinUtils.cpp:78-79)           */
    float proj00;          /* mProj[0][0], mProj[1][1] of matrixPerspectiveLHD3D */
    float proj11;
    float lens_radius;
    float focal_distance;
    /* OrthographicCamera (src/GoblinCamera.cpp:290-329): parallel rays from a film_width x
     * film_height window, mint 0; proj00 / proj11 then hold matrixOrthoLHD3D's 2 / w, 2 / h */
    int32_t orthographic;
    float film_width;
    float film_height;
} gb_camera;

typedef struct gb_film_desc {
    int32_t xres, yres;
    int32_t xstart, xcount, ystart, ycount;   /* crop window (Film ctor)      */
    int32_t sx0, sx1, sy0, sy1;               /* Film::getSampleRange          */
    float filter_width[2];
    float filter_table[256];                  /* FilterTable, 16 x 16          */
    /* Film::writeImage post-processing (src/GoblinFilm.cpp:164-192,203-212) */
    int32_t tone_mapping;                     /* Reinhard, applied to .ppm only */
    float bloom_radius;                       /* fraction of max(xres, yres)   */
    float bloom_weight;
} gb_film_desc;

typedef struct gb_render_setting {
    int32_t method;        /* GB_METHOD_*                                      */
    int32_t spp;           /* sample_per_pixel as written in the scene         */
    int32_t max_ray_depth;
    int32_t ao_sample_num; /* rounded up to a perfect square, as the reference's sample quota does */
    /* optional keys this implementation adds to "render_setting" (ignored by the reference):
     * how many GPUs to shard the samples over and the Philox seed; 0 = not given */
    int32_t gpu_num;
    int32_t seed;
} gb_render_setting;

/* Flattened scene.  All pointers are host pointers owned by whoever filled the
 * struct (gb_scene owns them when it came from gb_scene_get_desc). */
typedef struct gb_scene_desc {
    /* top level: BVH over instances (Scene::mBVH, src/GoblinScene.cpp:15) */
    const gb_bvh_node* top_nodes;
    uint32_t n_top_nodes;
    const uint32_t* top_order;      /* BVH leaf slot -> instance index       */
    const gb_instance* instances;
    uint32_t n_instances;
    /* models */
    const gb_model* models;
    uint32_t n_models;
    const gb_bvh_node* model_nodes; /* per-model BVHs, concatenated           */
    uint64_t n_model_nodes;
    const uint32_t* model_order;    /* BVH leaf slot -> face index (per model) */
    const uint32_t* tri_index;      /* 3 vertex indices per face (model local) */
    uint64_t n_tris;
    const float* vert_pos;          /* 3 per vertex                           */
    const float* vert_nrm;          /* 3 per vertex                           */
    const float* vert_uv;           /* 2 per vertex                           */
    uint64_t n_verts;
    /* shading */
    const gb_material* materials;
    uint32_t n_materials;
    const gb_light* lights;
    uint32_t n_lights;
    const float* light_power;       /* CDF1D::mFunction, n_lights             */
    const float* light_cdf;         /* CDF1D::mCDF, n_lights + 1              */
    const float* light_tri_area;    /* mesh emitters: GeometrySet::mGeometriesArea, face order */
    const float* light_tri_cdf;     /* mesh emitters: mAreaDistribution->mCDF  */
    uint32_t n_light_tri_area, n_light_tri_cdf;
    float world_bound[6];           /* Scene BVH AABB (getBoundingSphere)     */
    gb_camera camera;
    gb_film_desc film;
    gb_render_setting setting;
    const gb_texture* textures;     /* only entries materials reach are read  */
    uint32_t n_textures;
    const gb_image_level* image_levels; /* image textures: pyramid levels     */
    uint32_t n_image_levels;
    const float* image_texels;      /* image based lights / textures: RGBA float texels */
    uint64_t n_image_texels;        /* in texels (4 floats each)              */
    const float* light_dist;        /* image based lights: sampling tables    */
    uint64_t n_light_dist;
} gb_scene_desc;

typedef struct gb_render_params {
    uint64_t seed;          /* Philox key                                      */
    int32_t spp_total;      /* samples per pixel of the whole job (squared up) */
    int32_t spp_begin;      /* this call renders sample indices [begin, end)   */
    int32_t spp_end;
    int32_t max_ray_depth;  /* <= 0: take the scene's                          */
    int32_t method;         /* < 0: take the scene's                           */
    int32_t ao_sample_num;  /* <= 0: take the scene's                          */
} gb_render_params;

typedef struct gb_counters {
    uint64_t camera_samples;
    uint64_t rays_closest;      /* closest-hit traversals executed           */
    uint64_t rays_any;          /* any-hit (shadow / AO) traversals executed */
    uint64_t nodes_visited;     /* BVH nodes whose box test was evaluated    */
    uint64_t prims_tested;      /* triangle / sphere / disk tests            */
    uint64_t instances_entered; /* world->object ray transforms              */
    uint64_t kernel_launches;   /* launches of this library's kernels        */
    uint64_t nodes_visited_any;     /* the any-hit traversals' share of the  */
    uint64_t prims_tested_any;      /* three totals above                    */
    uint64_t instances_entered_any;
} gb_counters;

/* Kernel classes of the wavefront integrator, for gb_get_kernel_times. */
enum { GB_K_RAYGEN = 0, GB_K_EXTEND = 1, GB_K_SHADE = 2, GB_K_SHADOW = 3, GB_K_AO = 4, GB_K_FILM = 5,
       GB_K_TRACE = 6, GB_K_OTHER = 7, GB_K_COUNT = 8 };
typedef struct gb_kernel_times {
    double ms[GB_K_COUNT];          /* device time per class, CUDA events       */
    uint64_t launches[GB_K_COUNT];  /* timed launch groups per class            */
} gb_kernel_times;

typedef struct gb_scene gb_scene;     /* host-side flattened scene            */
typedef struct gb_context gb_context; /* one GPU                              */

/* ------------------------------------------------------------- host scene */

/* ContextLoader::load (src/GoblinContextLoader.cpp:447-503): parse the JSON
 * scene, load OBJ meshes, build the reference's equal_count BVHs bit-exactly,
 * flatten.  Pure host code; needs no GPU. */
int gb_scene_load_json(const char* path, gb_scene** out);
/* The same from a JSON string; mesh paths resolve against scene_dir. */
int gb_scene_load_json_string(const char* json, const char* scene_dir, gb_scene** out);
/* Loader options; zero-initialise for the reference's behaviour. */
typedef struct gb_load_options {
    int32_t bvh_method;    /* GB_BVH_* for the top-level and every per-model BVH */
    int32_t reserved[7];
} gb_load_options;
int gb_scene_load_json_ex(const char* path, const gb_load_options* options, gb_scene** out);
int gb_scene_load_json_string_ex(const char* json, const char* scene_dir, const gb_load_options* options,
                                 gb_scene** out);
void gb_scene_destroy(gb_scene* scene);
int gb_scene_get_desc(const gb_scene* scene, gb_scene_desc* out);
/* film output path chosen by the loader (film "file" or <scene>.exr) */
const char* gb_scene_output_path(const gb_scene* scene);

/* BVH::BVH (src/GoblinBVH.cpp:34-151) on raw boxes: nodes must hold
 * 2 * n entries, order n entries.  Returns the node count in *n_nodes. */
int gb_bvh_build(const float* aabbs /* n x 6 */, uint32_t n, gb_bvh_node* nodes,
                 uint32_t* n_nodes, uint32_t* order);
/* The same with an explicit GB_BVH_* split method. */
int gb_bvh_build_method(const float* aabbs /* n x 6 */, uint32_t n, int method, gb_bvh_node* nodes,
                        uint32_t* n_nodes, uint32_t* order);

/* --------------------------------------------------------------- device */

int gb_device_count(int* count);
int gb_create(int device, gb_context** out);
int gb_destroy(gb_context* ctx);
int gb_upload_scene(gb_context* ctx, const gb_scene_desc* desc);
/* The same without waiting for the device, for callers that stream scenes (an animation, one upload per frame):
 * the scene is staged into a SECOND pinned buffer / device arena and copied on the library's copy stream while the
 * kernels already queued on the context's stream keep reading the current scene; everything queued after the call
 * uses the new scene (the context's stream waits for the copy on the device, not the host).  The caller's arrays
 * may be freed when the call returns, as with gb_upload_scene.  Differences: the film is NOT cleared (the previous
 * frame may still be waiting to be downloaded; call gb_film_clear), unless its resolution changes; a bad vertex /
 * order index in the mesh arrays, which only the device-side derivation sees, is reported by the next
 * gb_synchronize / gb_film_download instead of by this call. */
int gb_upload_scene_async(gb_context* ctx, const gb_scene_desc* desc);
/* bytes the last gb_upload_scene moved host -> device (one copy of the staging arena) */
int gb_upload_bytes(gb_context* ctx, size_t* bytes);

/* Scene::intersect / Scene::occluded (src/GoblinScene.cpp:75-87) over a ray
 * batch.  Host buffers; copies are part of the call. */
int gb_trace_closest(gb_context* ctx, const gb_ray* rays, size_t n, gb_hit* hits);
int gb_trace_any(gb_context* ctx, const gb_ray* rays, size_t n, uint8_t* occluded);
/* The same on device-resident buffers (no copies), asynchronous on the
 * context's stream; gb_synchronize() to wait. */
int gb_trace_closest_device(gb_context* ctx, const gb_ray* d_rays, size_t n, gb_hit* d_hits);
int gb_trace_any_device(gb_context* ctx, const gb_ray* d_rays, size_t n, uint8_t* d_occluded);

/* PerspectiveCamera::generateRay (src/GoblinCamera.cpp:97-148) on explicit
 * samples: n x 4 floats (imageX, imageY, lensU1, lensU2) -> n rays. */
int gb_camera_rays(gb_context* ctx, const float* samples, size_t n, gb_ray* rays);

/* Renderer::Li (src/GoblinPathtracer.cpp:50-179, src/GoblinAO.cpp:12-37) on
 * explicit sample values; row layout as oracle/ref/ref_tool.cpp "li":
 * imageX, imageY, lensU1, lensU2, then 7 floats per bounce (path tracing) or
 * 2 per AO ray.  out: n x 3 radiance. */
int gb_li(gb_context* ctx, const float* samples, size_t n, size_t row_floats, float* out_rgb);

/* Renderer::render (src/GoblinRenderer.cpp:99-126): accumulate the sample
 * indices [spp_begin, spp_end) of every pixel of the sample range into the
 * device film.  Asynchronous on the context's stream. */
int gb_render(gb_context* ctx, const gb_render_params* params);
int gb_film_clear(gb_context* ctx);
/* Film::mPixels as 4 planes-interleaved floats per pixel (r, g, b, weight),
 * yres x xres x 4. Synchronises. */
int gb_film_download(gb_context* ctx, float* rgbw);
int gb_film_upload(gb_context* ctx, const float* rgbw);
/* Device address + float count of the film, for an external all-reduce
 * (torch.distributed / NCCL): the analogue of Film::mergeTile
 * (src/GoblinFilm.cpp:140-153). */
int gb_film_device_ptr(gb_context* ctx, void** ptr, size_t* n_floats);
/* Film::mergeTile across GPUs (src/GoblinFilm.cpp:140-153: every worker's full-frame tile is summed
 * into the film): one NCCL all-reduce(sum) of the device film over NVLink, issued by the library on the
 * context's stream.  NCCL is bound at run time (libnccl.so.2); GB_ERR_STATE when it is not installed.
 *   one process, one context per GPU:  gb_comm_init_all(ctxs, n), then gb_film_allreduce_all(ctxs, n)
 *   one process per GPU:               rank 0: gb_comm_unique_id; every rank: gb_comm_init_rank, then
 *                                      gb_film_allreduce(ctx) after each gb_render
 *   a communicator the caller owns:    gb_comm_attach(ctx, ncclComm_t, nranks)
 * After the all-reduce every rank's film holds the sum; rank 0 normalises and writes. */
#define GB_COMM_ID_BYTES 128
int gb_comm_init_all(gb_context** ctxs, int n);
int gb_comm_unique_id(void* id, size_t bytes);
int gb_comm_init_rank(gb_context* ctx, const void* id, size_t bytes, int nranks, int rank);
int gb_comm_attach(gb_context* ctx, void* nccl_comm, int nranks);
int gb_comm_destroy(gb_context* ctx);
int gb_comm_size(gb_context* ctx, int* nranks);   /* 0 = no communicator */
int gb_film_allreduce(gb_context* ctx);
int gb_film_allreduce_all(gb_context** ctxs, int n);
int gb_nccl_version(int* version);
/* Film::writeImage (src/GoblinFilm.cpp:164-192): colour / weight, written as
 * .exr (half, like src/GoblinImageIO.cpp:35-98), .pfm or .ppm by extension. */
int gb_film_write(gb_context* ctx, const char* path);
/* The image Film::writeImage hands to Goblin::writeImage: colour / weight, then
 * Goblin::bloom (src/GoblinImageIO.cpp:169-218) when the film's bloom_radius
 * and bloom_weight are positive -- both on the device.  rgb: yres x xres x 3. */
int gb_film_resolve(gb_context* ctx, float* rgb);
/* Host-only writers.  gb_write_image takes a raw film (colour / weight, no
 * bloom); gb_write_rgb takes a resolved image and applies the reference's
 * Reinhard tone mapping (src/GoblinImageIO.cpp:220-237) when asked and the
 * target is a .ppm, like Goblin::writeImage (:146-167). */
int gb_write_image(const char* path, const float* rgbw, int xres, int yres);
int gb_write_rgb(const char* path, const float* rgb, int xres, int yres, int tone_mapping);

int gb_synchronize(gb_context* ctx);
int gb_stream(gb_context* ctx, void** cuda_stream);
int gb_enable_counters(gb_context* ctx, int on); /* traversal statistics (slower) */
int gb_get_counters(gb_context* ctx, gb_counters* out);
int gb_reset_counters(gb_context* ctx);
/* milliseconds spent in the last gb_render / gb_trace_*_device call's kernels,
 * from CUDA events on the context's stream. Synchronises. */
int gb_last_kernel_ms(gb_context* ctx, float* ms);

/* Per-kernel-class device time: when enabled, every launch of this library is
 * bracketed by a CUDA event pair on the context's stream; gb_get_kernel_times
 * synchronises and returns the sums since the last reset. */
int gb_enable_kernel_timing(gb_context* ctx, int on);
int gb_get_kernel_times(gb_context* ctx, gb_kernel_times* out);
int gb_reset_kernel_times(gb_context* ctx);
/* Upper bound on the camera samples in flight per wave of the wavefront
 * integrator (path-state memory ~ 200 B per path). */
int gb_set_wave_paths(gb_context* ctx, size_t max_paths);
/* Execution knobs, for experiments: values[0..3] = warp scheduling of the traversal kernels
 * (refill-below, leaf batch, level batch, move floor; lanes, 0..33); values[4] = resident CTAs
 * per SM (0 = as many as fit); values[5] = wave lanes of a render (2 = two waves side by side
 * on their own streams, the default; 1 = one wave at a time); values[6] = run the shadow kernel
 * of a bounce beside the next extend kernel (1, default) or in line (0); values[7] = 1 keeps the
 * general film kernel also for filter windows of at most 32 pixels, where a one-warp form exists
 * (0, default: use it).  Results never depend on them beyond the order of float additions into
 * the film. */
int gb_set_tuning(gb_context* ctx, const int* values, int n);
/* How the traversal kernels walk the reference's tree (BVH::intersect / occluded,
 * src/GoblinBVH.cpp:189-280).  GB_TRACE_PAIR (default): pair nodes -- both children of an interior
 * node in one 64-byte record -- with every box test the reference evaluates, in its order; it is also
 * what runs while the counters are on.  GB_TRACE_WIDE: 4-wide nodes collapsed from the same tree on
 * the device, children still visited in the reference's order: same hit ids and distances (the one
 * exception: a ray with a zero direction component whose origin lies exactly on a plane of an
 * INTERMEDIATE node's box, where the reference's NaN comparison prunes a subtree the wide walk still
 * enters).  Measured 8 - 18 % slower than the pair walk on B200; kept as an option.  A scene whose
 * trees are too deep for the wide walk's shared-memory stack is always walked pair-wise;
 * gb_get_trace_mode reports what the kernels run for the uploaded scene. */
enum { GB_TRACE_WIDE = 0, GB_TRACE_PAIR = 1, GB_TRACE_EXACT = 1 /* older name */ };
int gb_set_trace_mode(gb_context* ctx, int mode);
int gb_get_trace_mode(gb_context* ctx, int* mode);

/* Debug builds (GB_DEBUG_STACK=1): the first out-of-range access of a traversal stack column, as
 * (kind 1 = push / 2 = pop, index, column height, block); all zero = none.  Other builds: all -1. */
int gb_debug_stack_violation(gb_context* ctx, int* out4);

const char* gb_last_error(void);
const char* gb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GOBLIN_B200_H */
