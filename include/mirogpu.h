/* mirogpu.h -- C ABI of the B200-native ray-intersection engine for the Miro ray tracer.
 *
 * This is the drop-in boundary for the reference's hot path.  The reference has no FFI: the path sits
 * behind C++ member functions.  Each entry point below names the reference interface it replaces; the
 * host-side C++ mirror in cse168-raytracer_b200/csrc/miro/ keeps those signatures and forwards here.
 *
 *   mirogpu_scene_create        <- BVH::build(Objects*, int)                         BVH.h:33, BVH.cpp:60-339
 *                                  (called from Scene::preCalc, Scene.cpp:72)
 *   mirogpu_intersect_batch     <- BVH::intersect / BVH::intersectChildren            BVH.h:35-38, BVH.cpp:438-658
 *                                  + Triangle::intersect                              Triangle.cpp:136-169
 *                                  (= Scene::trace with no unbounded objects,         Scene.cpp:214-268)
 *                                  + Sphere::intersect / Plane::intersect              Sphere.cpp:28-69, Plane.cpp:33-48
 *   mirogpu_generate_primary    <- Camera::eyeRay                                     Camera.cpp:104-161
 *   mirogpu_generate_bounce     <- Ray::diffuse / Ray::random                         Ray.h:109-140, Utility.h:34-50
 *   mirogpu_render(_rgb8)       <- Scene::raytraceImage + Scene::traceScene           Scene.cpp:93-212, 270-346
 *                                  + Phong::shade's light loop and shadow query       Phong.cpp:44-161
 *   mirogpu_photon_upload       <- the balanced Photon array Photon_map::balance leaves   PhotonMap.cpp:314-359
 *   mirogpu_photon_gather       <- Photon_map::irradiance_estimate / locate_photons   PhotonMap.cpp:81-243
 *   mirogpu_photon_trace        <- Scene::tracePhoton over the emissions of Scene::tracePhotons /
 *                                  traceCausticPhotons                                     Scene.cpp:351-472, 526-641
 *   mirogpu_photon_pass         <- Scene::tracePhotons / traceCausticPhotons as a whole: emission loop with its stop
 *                                  rule, Photon_map::store, scale_photon_power, balance      Scene.cpp:351-472, PhotonMap.cpp:246-466
 *   mirogpu_photon_balance      <- Photon_map::balance on a caller-filled array              PhotonMap.cpp:314-466
 *   mirogpu_photon_download     <- the Photon array a balanced Photon_map holds              PhotonMap.h:16-22, 81-83
 *   mirogpu_texture_lookup/_bump <- Texture::lookup2D / lookup3D / bumpHeight2D                Texture.h:62-72, Texture.cpp:358-510
 *
 * Conventions: every function returns an int status (MIROGPU_OK = 0), never throws, keeps no global
 * state besides the per-thread last-error string, and works on an opaque scene handle.  The caller owns
 * all host buffers.  Queries on one handle may be issued from several host threads.  Functions named
 * *_device take DEVICE pointers and a cudaStream_t (passed as void*) and are asynchronous on that stream;
 * the others take HOST pointers, copy in and out, and return when the result is in the caller's buffer.
 * There is no CPU fallback: without a CUDA device every call fails with MIROGPU_ERR_NO_DEVICE.
 *
 * Primitive identity: prim_id is the index of the triangle in the arrays given to mirogpu_scene_create,
 * which the host mirror fills in Scene::objects() insertion order (Scene.h:20-26).
 */
#ifndef MIROGPU_H_INCLUDED
#define MIROGPU_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIROGPU_VERSION 1

enum {
    MIROGPU_OK = 0,
    MIROGPU_ERR_INVALID_ARG = 1,
    MIROGPU_ERR_NO_DEVICE = 2,
    MIROGPU_ERR_CUDA = 3,
    MIROGPU_ERR_OOM = 4,
    MIROGPU_ERR_UNSUPPORTED = 5
};

#define MIROGPU_MISS 0xFFFFFFFFu
#define MIROGPU_TMAX 1e12f /* MIRO_TMAX, Miro.h:8 */

/* Ray {o, d} of Ray.h:40-46 plus the [tMin, tMax] arguments of BVH::intersect.  32 bytes. */
typedef struct mirogpu_ray {
    float ox, oy, oz, tmin;
    float dx, dy, dz, tmax;
} mirogpu_ray;

/* What Triangle::intersect computes before it derives P and N (Triangle.cpp:154-156).  16 bytes.
 * Miss: prim_id = MIROGPU_MISS and t = the ray's tmax (the reference leaves minHit.t = tMax, BVH.cpp:444).
 * The host reconstructs P = A + beta(B-A) + gamma(C-A), N = (1-beta-gamma)nA + beta nB + gamma nC. */
typedef struct mirogpu_hit {
    float t;
    uint32_t prim_id;
    float beta, gamma;
} mirogpu_hit;

/* Procedural textures behind TexturedPhong (Texture.h:108-277, Texture.cpp:358-510): what Phong::shade and Scene::tracePhoton
 * look up as the diffuse colour (Phong.cpp:50-55, Scene.cpp:546-551) and Scene::trace as the bump height (Scene.cpp:232-262).
 * tex[] holds the constructor arguments:
 *   CHECKER        color1[3], color2[3], scale          CheckerBoardTexture (2-D lookup: Object::toUVCoordinates)
 *   STONE          scale                                StoneTexture        (2-D lookup, the one texture with a bump height)
 *   STEM           scale                                StemTexture         (2-D lookup)
 *   PETAL          pivot[3], radius                     PetalTexture        (3-D lookup at the hit point)
 *   LEAF           scale                                LeafTexture         (3-D lookup; pivot and direction are not used by it)
 *   FLOWER_CENTER  pivot[3], radius                     FlowerCenterTexture (3-D lookup)
 * A material with a 3-D texture reports UVW lookup coordinates, so Scene::trace leaves its hit normal as the primitive computed
 * it (a triangle's interpolated normal stays un-normalised, Scene.cpp:237-262) -- reproduced. */
enum {
    MIROGPU_TEX_NONE = 0,
    MIROGPU_TEX_CHECKER = 1,
    MIROGPU_TEX_STONE = 2,
    MIROGPU_TEX_STEM = 3,
    MIROGPU_TEX_PETAL = 4,
    MIROGPU_TEX_LEAF = 5,
    MIROGPU_TEX_FLOWER_CENTER = 6
};

/* Phong(kd, ks, kt, shininess, refractIndex) AFTER the constructor's energy clamp (Phong.cpp:13-32); texture != NONE:
 * TexturedPhong(texture, ks, kt, shininess, refractIndex), whose kd is Vector3(1) before the clamp (Texture.cpp:513-517). */
typedef struct mirogpu_material {
    float kd[3];
    float ks[3];
    float kt[3];
    float shininess; /* +inf disables the highlight (Phong.cpp:149) */
    float refract_index;
    int32_t texture; /* MIROGPU_TEX_* */
    float tex[12];
} mirogpu_material;

/* PointLight (kind 0, PointLight.h) or DirectionalAreaLight (kind 1, DirectionalAreaLight.h). */
typedef struct mirogpu_light {
    int32_t kind;
    float position[3];
    float color[3];
    float wattage;
    float normal[3]; /* kind 1 only */
    float radius;    /* kind 1 only */
} mirogpu_light;

/* Camera state of Camera.h:56-61 (eye, up, viewDir already normalised as the setters do; fov in degrees). */
typedef struct mirogpu_camera {
    float eye[3];
    float up[3];
    float view_dir[3];
    float fov_degrees;
} mirogpu_camera;

enum { MIROGPU_LAYOUT_BVH2 = 0, MIROGPU_LAYOUT_CWBVH8 = 1, MIROGPU_LAYOUT_BVH4 = 2, MIROGPU_LAYOUT_QBVH4 = 3 };

/* Who builds the tree (BVH::build, BVH.cpp:60-339).  SAH_HOST: binned SAH on the host cores (0.5 s for 1.4 M triangles).
 * LBVH_DEVICE: Morton-order linear BVH (Karras) built, collapsed four-wide and quantised entirely on the GPU (milliseconds;
 * trees trace much slower where meshes overlap).  PLOC_DEVICE: the same pipeline with the binary tree made by parallel
 * locally-ordered clustering -- bottom-up merges of Morton neighbours by smallest union area (tens of milliseconds; tree
 * quality close to SAH).  The device builders emit QBVH4 only and fall back to SAH_HOST if their tree is too deep for the
 * kernels' stacks.  All give identical hits. */
enum { MIROGPU_BUILDER_SAH_HOST = 0, MIROGPU_BUILDER_LBVH_DEVICE = 1, MIROGPU_BUILDER_PLOC_DEVICE = 2 };

/* The device builders keep their scratch allocation (400-500 bytes per triangle of the largest scene built so far) for the
 * next build in this process; this returns it to the driver. */
void mirogpu_release_build_scratch(void);

typedef struct mirogpu_build_options {
    int32_t layout;       /* MIROGPU_LAYOUT_*; default QBVH4 (the fastest measured) */
    int32_t max_leaf;     /* triangles per leaf (BVH2 / BVH4 / QBVH4: <= 4 like the reference's OBJECTS_PER_LEAF; CWBVH8: <= 3) */
    int32_t sah_bins;     /* binned SAH resolution, default 32 */
    int32_t device;       /* CUDA device ordinal, -1 = current device */
    int32_t builder;      /* MIROGPU_BUILDER_*; default SAH_HOST */
} mirogpu_build_options;

typedef struct mirogpu_scene_info {
    uint32_t num_triangles;
    uint32_t num_nodes;        /* nodes of the flat layout in HBM */
    uint32_t num_binary_nodes; /* nodes of the intermediate binary SAH tree */
    uint32_t num_binary_leaves;
    uint32_t max_depth;
    int32_t layout;
    int32_t builder;           /* MIROGPU_BUILDER_* that produced the tree in HBM */
    int32_t _pad;
    uint64_t node_bytes;
    uint64_t triangle_bytes;
    uint64_t shading_bytes;
    double build_seconds;   /* host SAH build */
    double flatten_seconds; /* wide collapse + encode */
    double upload_seconds;
    float bounds_min[3];
    float bounds_max[3];
} mirogpu_scene_info;

enum { MIROGPU_CLOSEST_HIT = 0, MIROGPU_ANY_HIT = 1 };
/* Optional hint or-ed into a query mode: the batch is coherent (neighbouring rays start close together and point
 * the same way, e.g. camera rays in pixel order).  Coherent batches go to the packet kernel (32 consecutive rays per
 * warp, while-while), everything else to the hybrid-scheduled kernel with ray replacement.  Results are identical. */
enum { MIROGPU_HINT_COHERENT = 0x100 };

/* mode bits for mirogpu_render */
enum {
    MIROGPU_RENDER_WHITTED = 0,        /* the reference's default build: direct light + shadow + reflect/refract recursion */
    MIROGPU_RENDER_DIFFUSE_BOUNCE = 1, /* BASELINE config 3: jittered primary + one cosine-weighted Ray::diffuse bounce */
    MIROGPU_RENDER_PRIMARY_ONLY = 2    /* direct light, no shadow rays (reference -DDISABLE_SHADOWS) */
};

typedef struct mirogpu_render_params {
    int32_t width, height;
    int32_t spp;       /* samples per pixel; 1 with jitter = 0 reproduces the reference's pixel-centre ray */
    int32_t jitter;    /* 1: Camera::eyeRay(randomize = true) with the counter-based RNG */
    int32_t max_depth; /* TRACE_DEPTH, Miro.h:13 */
    int32_t mode;      /* MIROGPU_RENDER_* */
    uint32_t seed;
    int32_t tonemap;   /* 1: apply Scene.cpp:177-202 (NaN -> max, sigmoid(6v-3)); output still float */
    /* Screen-space shard for multi-GPU: this call renders rows [row_begin, row_end) interleaved
     * as row % row_stride == row_phase.  Full frame: 0, height, 1, 0. */
    int32_t row_begin, row_end, row_stride, row_phase;
    float bg_color[3];
    int32_t use_photon_maps; /* 1: add irradiance_estimate of both maps at diffuse hits (Scene.cpp:286-299) */
    int32_t shadows;         /* 0: the reference's -DDISABLE_SHADOWS build (Phong.cpp:91): no shadow rays */
} mirogpu_render_params;

/* Per-call work counters of the instrumented kernels (reference: -DSTATS, Stats.h). */
typedef struct mirogpu_counters {
    uint64_t rays;
    uint64_t node_visits;    /* flat-layout nodes fetched */
    uint64_t box_tests;      /* child boxes tested */
    uint64_t triangle_tests; /* reference: Stats::Ray_Tri_Intersect, BVH.cpp:496 */
    uint64_t hits;
    uint64_t bytes_fetched;  /* node + triangle bytes requested by the traversal */
} mirogpu_counters;

/* Sphere (Sphere.h:7-38, Sphere.cpp:28-69): a bounded object, a leaf primitive of the tree next to the triangles. */
typedef struct mirogpu_sphere {
    float center[3];
    float radius;
    uint32_t material_id;
    uint32_t _pad;
} mirogpu_sphere;

/* Plane (Plane.h:12-36, Plane.cpp:33-48): an unbounded object -- tested after the tree walk like Scene::trace's
 * m_unboundedObjects loop (Scene.cpp:219-230).  normal is used as given (Scene::trace normalises N afterwards). */
typedef struct mirogpu_plane {
    float normal[3];
    float origin[3];
    uint32_t material_id;
    uint32_t _pad;
} mirogpu_plane;

/* Everything mirogpu_scene_create_ex builds a scene from.  Primitive ids: triangle i -> i, sphere j -> ntris + j,
 * plane k -> ntris + nspheres + k.  devices / ndevices: the CUDA devices the scene is REPLICATED on (SURVEY 8b "device_mask",
 * 8e: image rows shard over them, the tree is replicated); NULL / 0 = the one device of mirogpu_build_options.device. */
typedef struct mirogpu_scene_desc {
    const float* tri_vertices;         /* ntris x 9 floats (A, B, C) */
    const float* tri_normals;          /* ntris x 9 floats (nA, nB, nC) or NULL */
    const uint32_t* tri_material_ids;  /* ntris entries or NULL (all 0) */
    uint32_t ntris;
    uint32_t nspheres;
    const mirogpu_sphere* spheres;
    const mirogpu_plane* planes;
    uint32_t nplanes;
    uint32_t nmaterials;
    const mirogpu_material* materials; /* NULL: one white Lambert */
    const int32_t* devices;
    uint32_t ndevices;
    uint32_t _pad;
    const float* tri_texcoords;        /* ntris x 6 floats (tA, tB, tC: TriangleMesh::texCoords of the triangle's corners) or NULL =
                                          meshes without texture coordinates (Triangle::toUVCoordinates then answers (0, 0)) */
} mirogpu_scene_desc;

typedef struct mirogpu_scene* mirogpu_handle;

/* ---- lifetime ------------------------------------------------------------------------------------------ */
int mirogpu_version(void);
const char* mirogpu_last_error(void);
int mirogpu_device_count(int* count);

/* Builds the acceleration structure on the host from per-triangle arrays, flattens it to the GPU layout
 * and uploads it to the selected device.  tri_vertices: ntris x 9 floats (A, B, C); tri_normals: ntris x 9
 * floats (nA, nB, nC) or NULL; material_ids: ntris entries or NULL (all 0); materials may be NULL (one
 * white Lambert).  opt may be NULL. */
int mirogpu_scene_create(const float* tri_vertices, const float* tri_normals, const uint32_t* material_ids,
                         uint32_t ntris, const mirogpu_material* materials, uint32_t nmaterials,
                         const mirogpu_build_options* opt, mirogpu_handle* out);
/* The same with spheres, planes and a device list (see mirogpu_scene_desc).  On a multi-device handle the batch queries run on
 * the first device; mirogpu_render / mirogpu_render_rgb8 shard the image rows over all devices (row % ndevices) and gather the
 * framebuffer on the first one over NVLink peer copies; *_device entry points address the first device's replica. */
int mirogpu_scene_create_ex(const mirogpu_scene_desc* desc, const mirogpu_build_options* opt, mirogpu_handle* out);
/* Number of devices the scene is replicated on and their ordinals (devices may be NULL; at most `capacity` entries written). */
int mirogpu_scene_devices(mirogpu_handle h, int32_t* devices, uint32_t capacity, uint32_t* ndevices);
int mirogpu_scene_destroy(mirogpu_handle h);
int mirogpu_scene_info_get(mirogpu_handle h, mirogpu_scene_info* info);
int mirogpu_scene_set_lights(mirogpu_handle h, const mirogpu_light* lights, uint32_t nlights);

/* Copies the flat node array / the reordered triangle array back to the host for inspection (tests check
 * every flat box against the triangles below it).  *bytes in: capacity, out: size needed/written. */
int mirogpu_debug_copy_nodes(mirogpu_handle h, void* out, uint64_t* bytes);
int mirogpu_debug_copy_triangles(mirogpu_handle h, void* out, uint64_t* bytes);

/* ---- batched intersection (BVH::intersect over n rays) ---------------------------------------------- */
int mirogpu_intersect_batch(mirogpu_handle h, const mirogpu_ray* rays, size_t n, mirogpu_hit* hits, int mode);
int mirogpu_intersect_batch_device(mirogpu_handle h, const mirogpu_ray* d_rays, size_t n, mirogpu_hit* d_hits,
                                   int mode, void* cuda_stream);
/* Same query through the instrumented kernel; counters are accumulated into *c (host struct). */
int mirogpu_intersect_batch_counted(mirogpu_handle h, const mirogpu_ray* rays, size_t n, mirogpu_hit* hits,
                                    int mode, mirogpu_counters* c);
/* Kernel variant selection for measurement: -1 = automatic (default: the hybrid kernel, except BVH2 batches carrying
 * MIROGPU_HINT_COHERENT and all CWBVH8 batches, which go to the packet kernel), 0 = persistent warps taking 32-ray
 * packets (while-while), 1 = one thread per ray, 2 = persistent warps with hybrid step scheduling + ray replacement
 * (BVH2, BVH4, QBVH4). */
int mirogpu_set_kernel_variant(mirogpu_handle h, int variant);

/* Reconstructs P, N (normalised as Scene::trace does for UV materials, Scene.cpp:262) and material id
 * from hits.  Outputs are n x 3 floats / n uint32; any may be NULL. */
int mirogpu_resolve_hits_device(mirogpu_handle h, const mirogpu_hit* d_hits, size_t n, float* d_P, float* d_N,
                                uint32_t* d_material, void* cuda_stream);
/* The same for scenes with spheres or planes, whose hit point is o + t d (Sphere.cpp:62, Plane.cpp:42): needs the rays.
 * (mirogpu_resolve_hits_device fails with MIROGPU_ERR_INVALID_ARG on such a scene.) */
int mirogpu_resolve_hits_rays_device(mirogpu_handle h, const mirogpu_ray* d_rays, const mirogpu_hit* d_hits, size_t n, float* d_P,
                                     float* d_N, uint32_t* d_material, void* cuda_stream);

/* ---- device-side ray generation ------------------------------------------------------------------------ */
/* Camera::eyeRay for the pixels of this call's rows, for `sample_count` samples starting at `sample_begin`.
 * Rows: row_begin + row_phase + j*row_stride < row_end (row 0 = bottom) -- (0, height, 1, 0) is the full frame,
 * (0, height, N, r) is rank r's interleaved share on N GPUs.  Output order: sample-major, then row-major over
 * the local rows: sample_count * local_rows * width rays.  jitter = 0: pixel centre (dx = dy = 0.5). */
int mirogpu_generate_primary_device(mirogpu_handle h, const mirogpu_camera* cam, int width, int height,
                                    int row_begin, int row_end, int row_stride, int row_phase, int jitter, uint32_t seed,
                                    uint32_t sample_begin, uint32_t sample_count, mirogpu_ray* d_rays, void* cuda_stream);
/* Ray::diffuse at every hit of (d_rays, d_hits): phi = asin(sqrt(u1)), theta = 2 pi u2, direction aligned to
 * the normalised shading normal, origin P + eps*dir.  Misses yield a ray with tmax < tmin (never hits).
 * Ray i draws its uniforms at RNG index index_base + i.  d_live_count (device uint64, may be NULL) is
 * incremented by the number of live (non-miss) rays generated. */
int mirogpu_generate_bounce_device(mirogpu_handle h, const mirogpu_ray* d_rays, const mirogpu_hit* d_hits, size_t n,
                                   uint32_t seed, uint32_t sample, uint32_t index_base, mirogpu_ray* d_out,
                                   unsigned long long* d_live_count, void* cuda_stream);
/* The uniform numbers the generators draw, for feeding the oracle the same samples: out = n x 2 floats. */
int mirogpu_rng_uniforms(uint32_t seed, uint32_t sample, uint32_t dimension, size_t first, size_t n, float* out);

/* ---- whole-frame render (Scene::raytraceImage) -------------------------------------------------------- */
/* rgb_out: HOST buffer, width*height*3 floats, row 0 = bottom.  Only the rows of this call's shard are written. */
int mirogpu_render(mirogpu_handle h, const mirogpu_camera* cam, const mirogpu_render_params* p, float* rgb_out);
/* Same frame as the reference's Image holds it (Image.h: 3 bytes per pixel): tone-mapped on the device
 * (Scene.cpp:177-202) and truncated to 8 bits like Image::setPixel's Map() (Image.cpp:47-52).
 * rgb8_out: HOST buffer, width*height*3 bytes, row 0 = bottom. */
int mirogpu_render_rgb8(mirogpu_handle h, const mirogpu_camera* cam, const mirogpu_render_params* p, uint8_t* rgb8_out);
/* Tone map + 8-bit conversion of a complete DEVICE float frame (after a multi-GPU gather). */
int mirogpu_tonemap_rgb8_device(mirogpu_handle h, const float* d_rgb, int width, int height, uint8_t* d_rgb8, void* cuda_stream);
/* The same tone map for a frame whose rows are sharded over several GPUs, in two steps around the one exchange it needs: the
 * largest non-NaN value of the frame replaces NaN pixels (Scene.cpp:157-164), so each rank reduces its own rows
 * (mirogpu_frame_max_device -> *d_max, -inf if it has none), the ranks combine the values (one float, max), and each rank
 * maps its own rows to 8 bits with the combined value (mirogpu_tonemap_rows_rgb8_device).  Rows are those of a render call:
 * row_begin + row_phase + j * row_stride < row_end.  d_rgb and d_rgb8 are full-frame DEVICE buffers; only those rows are
 * read / written. */
int mirogpu_frame_max_device(mirogpu_handle h, const float* d_rgb, int width, int height, int row_begin, int row_end, int row_stride,
                             int row_phase, float* d_max, void* cuda_stream);
int mirogpu_tonemap_rows_rgb8_device(mirogpu_handle h, const float* d_rgb, int width, int height, int row_begin, int row_end,
                                     int row_stride, int row_phase, const float* d_max, uint8_t* d_rgb8, void* cuda_stream);
/* d_rgb: DEVICE buffer of the same shape.  rays_traced (host, may be NULL) receives the ray count after
 * the stream is synchronised by the caller only if sync != 0. */
int mirogpu_render_device(mirogpu_handle h, const mirogpu_camera* cam, const mirogpu_render_params* p, float* d_rgb,
                          void* cuda_stream);
/* Page-locked host memory for buffers that cross the boundary every frame (the reference's Image::m_pixels, Image.cpp:22,
 * ray / hit arrays of a batch): copies to and from it run as one DMA instead of being staged through the driver's bounce
 * buffer.  Plain malloc'd memory is accepted everywhere as well.  mirogpu_host_alloc returns NULL on failure. */
void* mirogpu_host_alloc(size_t bytes);
void mirogpu_host_free(void* p);
/* Rays traced (all kinds) and kernels launched by the last render / intersect call on this handle. */
int mirogpu_last_call_stats(mirogpu_handle h, uint64_t* rays_traced, uint64_t* kernel_launches);

/* ---- photon map ------------------------------------------------------------------------------------------ */
/* which: 0 = global map, 1 = caustic map (Scene.h:58-59).  photons: (stored+1) records of the reference's
 * 28-byte Photon (PhotonMap.h:16-22) in the heap order Photon_map::balance produces; entry 0 is unused. */
int mirogpu_photon_upload(mirogpu_handle h, int which, const void* photons, int stored);
/* Scene::tracePhoton (Scene.cpp:526-641) for emissions [first_emission, first_emission + count) of light
 * `light_index` -- a DirectionalAreaLight, the only kind the reference emits from (Scene.cpp:368): origin =
 * samplePhotonOrigin (disc rejection sampling, DirectionalAreaLight.h:20-24), direction = the light normal, power =
 * color * wattage * PI r^2, divided by 10 when caustic != 0 (Scene.cpp:379-385, 431-434).  Every emission is walked for
 * up to TRACE_DEPTH_PHOTONS + 1 segments with the reference's roulette; a photon is recorded at each diffuse choice
 * after the first hit (caustic pass: only once a specular surface has been left).  counts[i] (HOST, count bytes) = photons
 * recorded by emission first_emission + i (0..5); records (HOST, count * 5 * 9 floats) holds for emission i up to five
 * {power[3], pos[3], incoming dir[3]} triples at records + i*45.  Powers are unscaled: the caller stores them
 * (Photon_map::store) in emission order until its target is met and scales by 1 / emissions (Scene.cpp:400).
 * The walk is a pure function of (seed, emission index): any split of the range into calls / GPUs gives the same map. */
int mirogpu_photon_trace(mirogpu_handle h, int light_index, int caustic, uint32_t seed, uint64_t first_emission, uint32_t count,
                         uint8_t* counts, float* records);
/* Scene::tracePhotons (caustic = 0) / traceCausticPhotons (caustic = 1) as one call that never leaves the device
 * (Scene.cpp:351-472): for every DirectionalAreaLight, in order, emissions are walked in batches (mirogpu_photon_trace's
 * kernel, emission index = random stream, first index (light << 40)), their records stored in emission order while fewer
 * than `target` photons are stored -- the reference's sequential "if (photonsAdded < PhotonsPerLightSource)" rule,
 * evaluated with a prefix sum over the per-emission record counts -- with Photon_map::store's direction quantisation and
 * bounding box (PhotonMap.cpp:246-288); powers are scaled by 1 / emissions (Scene.cpp:400, PhotonMap.cpp:297-306); the
 * map is balanced into the reference's left-balanced heap order (PhotonMap.cpp:314-466: Jensen's median_split is
 * evaluated round for round, each Hoare partition in parallel, so the array equals Photon_map::balance's also among equal
 * keys) and becomes map `which` of the handle (all devices of a multi-device handle).  max_emissions <= 0: 2^28 (the
 * reference loops forever when nothing can be stored).  emissions_out / stored_out may be NULL. */
int mirogpu_photon_pass(mirogpu_handle h, int which, int caustic, uint32_t seed, int target, long long max_emissions,
                        long long* emissions_out, int* stored_out);
/* Map `which` as (stored + 1) records of the reference's 28-byte Photon in heap order (entry 0 unused); photons == NULL
 * only reports the count.  capacity = records the buffer holds beyond entry 0. */
int mirogpu_photon_download(mirogpu_handle h, int which, void* photons, int capacity, int* stored);
/* Photon_map::balance (PhotonMap.cpp:314-466) on device `device` for a caller-filled array of (stored + 1) 28-byte Photon
 * records (entry 0 unused), in place: same heap order, `plane` set on the inner nodes.  bbox_min / bbox_max: the box
 * Photon_map::store accumulated (the split axes derive from it, PhotonMap.cpp:431-436).  Needs no scene handle. */
int mirogpu_photon_balance(int device, void* photons, int stored, const float bbox_min[3], const float bbox_max[3]);
/* ---- procedural textures, single queries on the host ------------------------------------------------------------- */
/* Texture::lookup2D / lookup3D (Texture.h:66-72) and bumpHeight2D for the kinds above, evaluated by the same code the device
 * shading kernels run (csrc/texture.cuh, compiled for the host): what the host layer's Texture classes answer with.  kind:
 * MIROGPU_TEX_*; tex: the 12 parameters of mirogpu_material::tex; (u, v) for the 2-D kinds, (u, v, w) = the point for the 3-D
 * kinds.  Not a rendering path -- frames evaluate textures on the device. */
int mirogpu_texture_lookup(int kind, const float* tex, float u, float v, float w, float* rgb3);
int mirogpu_texture_bump(int kind, const float* tex, float u, float v, float* height);
/* Gather search of map `which`.  exact = 0 (default): one query per warp -- a shared stack of kd nodes and a shared
 * candidate buffer, 32 nodes tested per step, the k-th distance found by bisection when the buffer fills; it ends with the
 * same k nearest photons as the reference's search, summed in another order (estimates agree to ~1e-6 relative; k <= 512).
 * exact = 1: the reference's search verbatim, one query per thread (same visiting order, heap and summation order:
 * bit-identical estimates), several times slower. */
int mirogpu_photon_set_exact(mirogpu_handle h, int which, int exact);
int mirogpu_photon_gather(mirogpu_handle h, int which, const float* pos3, const float* normal3, size_t n,
                          float max_dist, int k, float* irrad3);
int mirogpu_photon_gather_device(mirogpu_handle h, int which, const float* d_pos3, const float* d_normal3, size_t n,
                                 float max_dist, int k, float* d_irrad3, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* MIROGPU_H_INCLUDED */
