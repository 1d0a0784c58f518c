/* igb200.h -- C ABI of the B200-native render device for the Ignis `path` hot path.
 *
 * This is the drop-in boundary: the reference's device is a C++ plugin
 * (`ig_get_interface()` -> IDeviceInterface -> IRenderDevice, src/device/Interface.cpp:70-76,
 * src/runtime/device/IRenderDevice.h:14-81). A plugin shim (`B200Device : IRenderDevice`,
 * ignis_b200/csrc/b200_device.h) forwards each IRenderDevice method to one entry point below;
 * everything else in this repository (Python harness, tests, bench) calls the same entry points.
 *
 * Conventions: every function returns 0 on success and a negative code on failure, with a message available from
 * igb200_last_error() (the reference logs and returns false/nullptr, src/device/Device.cpp:303-306). Calls are
 * synchronous and not thread-safe per context (all IRenderDevice calls come from the runtime's caller thread,
 * src/device/Device.cpp:1632). The caller owns every input buffer (they may be freed after the call returns);
 * the context owns every output buffer. Plain pointers and sizes only.
 */
#ifndef IGB200_H
#define IGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IGB200_VERSION_MAJOR 0
#define IGB200_VERSION_MINOR 3 /* must match the runtime's Build::getVersion(): DeviceManager.cpp:180-194 */

typedef struct igb200_ctx igb200_ctx;

/* ---- binary tables, exactly as the reference runtime produces them --------------------------------------- */

/* DynTable lookup entry: src/runtime/table/DynTable.h:6-10, src/artic/driver/data.art:1-5 */
typedef struct igb200_lookup_entry {
    uint32_t type_id; /* shape provider: IGB200_SHAPE_* */
    uint32_t flags;
    uint64_t offset;  /* byte offset of the entry inside the dyn-table data blob */
} igb200_lookup_entry;

/* Scene-BVH leaf `EntityLeaf1`: src/artic/traversal/bvh.art:52-61, writer src/runtime/bvh/SceneBVHAdapter.h:68-100.
 * The device rebuilds its own BVH8, so user1/user2 (offset of the reference's pre-baked prim BVH) are ignored. */
typedef struct igb200_entity_leaf {
    float    min[3];
    int32_t  entity_id; /* bit 31 = last-in-leaf marker in the reference; masked off here */
    float    max[3];
    int32_t  shape_id;
    float    local[12]; /* world -> local 3x4, column major */
    uint32_t flags;     /* visibility: 1 camera, 2 light, 4 bounce, 8 shadow (LoaderEntity.cpp:122-131) */
    int32_t  mat_id;
    int32_t  user1, user2;
} igb200_entity_leaf;

enum { IGB200_SHAPE_TRIMESH = 0, IGB200_SHAPE_SPHERE = 1 };

/* ---- descriptors: the arguments of the constructor calls in the reference's generated stage shaders ------- */

enum { IGB200_BSDF_DIFFUSE = 0,    /* make_diffuse_bsdf(surf, 0, kd) -> make_lambertian_bsdf  (bsdf/diffuse.art:2-12,55-61) */
       IGB200_BSDF_DIELECTRIC = 1, /* make_dielectric_bsdf(..., delta, thin=false) -> make_pure_dielectric_bsdf (bsdf/dielectric.art:15-37) */
       IGB200_BSDF_CONDUCTOR = 2   /* make_conductor_bsdf(..., delta) -> make_mirror_bsdf / make_pure_conductor_bsdf (bsdf/conductor.art:2-27,131-141) */ };

/* Microfacet distribution of a CONDUCTOR (BSDF::setupRoughness, src/runtime/bsdf/BSDF.cpp:53-98; src/artic/core/microfacet.art):
 * no roughness property -> make_delta_distribution; else make_vndf_ggx_distribution(face_normal, local, alpha_u, alpha_v), which is
 * itself the delta distribution when alpha_u or alpha_v <= 1e-4 (microfacet.art:297,403-425). alpha_u / alpha_v are the results of
 * microfacet::compute_explicit(roughness, anisotropic) (microfacet.art:427-432), evaluated by the host. */
enum { IGB200_MICROFACET_DELTA = 0, IGB200_MICROFACET_VNDF_GGX = 1 };
/* Normal-modifying wrappers around the material's BSDF (src/artic/bsdf/map.art:56-68, generator src/runtime/bsdf/MapBSDF.cpp) */
enum { IGB200_MAP_NONE = 0,
       IGB200_MAP_BUMP = 1,    /* make_bumpmap(ctx, inner, texture_dx(map, ctx).r, texture_dy(map, ctx).r, strength) */
       IGB200_MAP_NORMAL = 2   /* make_normalmap(ctx, inner, map(ctx), strength) */ };

typedef struct igb200_material {
    int32_t bsdf;     /* IGB200_BSDF_* */
    int32_t light_id; /* finite-light index if the entity is an area emitter (make_emissive_material), else -1 */
    float   p[14];    /* DIFFUSE: kd rgb | DIELECTRIC: ext_ior, int_ior, ks rgb, kt rgb |
                         CONDUCTOR: eta rgb, k rgb, ks rgb, mirror flag (eta, k were compile-time constants ~ (0, 1)) */
    int32_t tex[2];   /* >= 0: index of the texture that replaces a colour parameter (ShadingTree::addColor with a texture name):
                         [0] DIFFUSE reflectance, CONDUCTOR / DIELECTRIC specular_reflectance; [1] DIELECTRIC specular_transmittance. -1: the constant in p */
    int32_t distribution;  /* CONDUCTOR: IGB200_MICROFACET_* */
    float   alpha_u, alpha_v;
    int32_t map_kind;      /* IGB200_MAP_* */
    int32_t map_tex;       /* texture of the bump / normal map */
    float   map_strength;
    int32_t reserved[8];   /* 0 */
} igb200_material;     /* 128 bytes */

/* ---- textures: src/artic/texture/{common,checkerboard,image}.art; generators src/runtime/pattern/{CheckerBoard,Image}Pattern.cpp -- */
enum { IGB200_TEX_CHECKERBOARD = 0, /* make_checkerboard_texture(scale, color0, color1, transform), texture/checkerboard.art:4-13 */
       IGB200_TEX_IMAGE = 1         /* make_image_texture(border, filter, image, transform), texture/image.art:148-153 */ };
enum { IGB200_FILTER_NEAREST = 0, IGB200_FILTER_BILINEAR = 1, IGB200_FILTER_BICUBIC = 2 };   /* texture/image.art:78-146 */
enum { IGB200_BORDER_REPEAT = 0, IGB200_BORDER_CLAMP = 1, IGB200_BORDER_MIRROR = 2 };        /* texture/image.art:9-44 */
typedef struct igb200_texture {
    int32_t type;               /* IGB200_TEX_* */
    int32_t image;              /* IMAGE: index into igb200_scene_desc.images */
    int32_t filter;             /* IMAGE: IGB200_FILTER_* */
    int32_t border_u, border_v; /* IMAGE: IGB200_BORDER_* (make_split_border when they differ) */
    int32_t reserved[3];
    float   transform[6];       /* rows 0 and 1 of the 3x3 the generator inlines (LoaderUtils::inlineTransformAs2d): uv' = (row0 . (u,v,1), row1 . (u,v,1)) */
    float   p[10];              /* CHECKERBOARD: scale_x, scale_y, color0 rgb, color1 rgb */
} igb200_texture;               /* 96 bytes */

/* Pixel data exactly as the reference's device keeps it after loading the file (src/device/Device.cpp:735-799, src/runtime/Image.cpp:500-810):
 * rows bottom-up (stbi_set_flip_vertically_on_load), 8-bit images packed and -- unless the texture says `linear` -- already mapped from sRGB
 * to linear BYTES (byte_color_to_linear, Image.cpp:48-51); float images RGBA. Decoding files is the caller's business. */
enum { IGB200_IMAGE_RGBA8 = 0,    /* device.load_packed_image(.., 4, ..): channel k of a pixel = byte k / 255 (driver/image.art:9-14) */
       IGB200_IMAGE_MONO8 = 1,    /* device.load_packed_image(.., 1, ..): grey = byte / 255 (driver/image.art:16,27-34) */
       IGB200_IMAGE_RGBA32F = 2   /* device.load_image(.., 4): four floats per pixel */ };
typedef struct igb200_image {
    int32_t     format, width, height, reserved;
    const void* pixels;
} igb200_image;

enum { IGB200_LIGHT_ENV_CONST = 0,  /* make_environment_light (constant radiance), light/env.art:75-100,161-164 */
       IGB200_LIGHT_POINT = 1,      /* make_point_light, light/point.art:1-18 */
       IGB200_LIGHT_PLANE_AREA = 2, /* make_area_light(make_plane_area_emitter), light/area.art:10-43,124-258 */
       IGB200_LIGHT_SHAPE_AREA = 3, /* make_area_light(make_shape_area_emitter), light/area.art:62-107 */
       IGB200_LIGHT_SPHERE_AREA = 4, /* make_area_light(make_sphere_area_emitter), light/area.art:260-316 */
       IGB200_LIGHT_SPOT = 5,       /* make_spot_light, light/spot.art:8-44 */
       IGB200_LIGHT_SUN = 6,        /* make_sun_light (not handled as delta), light/sun.art:10-48: an INFINITE light */
       IGB200_LIGHT_DIRECTIONAL = 7, /* make_directional_light, light/directional.art:1-17: an INFINITE delta light */
       IGB200_LIGHT_ENV_TEXTURED = 8, /* make_environment_light_textured(.., scale, tex, cdf::make_cdf_2d_from_buffer(..), transform), light/env.art:112-160:
                                         environment map (EnvironmentLight.cpp:60-91 with cdf = conditional) and the sky model (SkyLight.cpp:42-80) */
       IGB200_LIGHT_ENV_TEX = 9       /* make_environment_light(.., scale, tex, transform), light/env.art:161-167: textured, sampled uniformly (cdf = none) */ };

typedef struct igb200_light {
    int32_t type;      /* IGB200_LIGHT_* */
    int32_t entity_id; /* area lights: the emissive entity */
    float   p[30];     /* ENV_CONST: radiance rgb | POINT: position xyz, intensity rgb |
                          PLANE_AREA: origin, x_axis, y_axis, normal (3 each), area, t0..t3 (2 each), radiance rgb |
                          SHAPE_AREA: radiance rgb |
                          SPOT: position xyz, direction xyz, cos(cutoff), cos(falloff), intensity rgb |
                          SPHERE_AREA: radiance rgb, sphere origin xyz (local), radius, area (compute_ellipsoid_area, shapes/sphere.art:21-27) |
                          SUN: direction towards the sun xyz (unit), cos(angle / 2), radiance rgb |
                          DIRECTIONAL: direction the light travels xyz (unit), irradiance rgb |
                          ENV_TEXTURED: scale rgb, transform 3x3 column major (9), then as int32 bits: texture, first word of the cdf inside
                                        aux_data, size_x (conditional), size_y (marginal); the buffer is [marginal size_y | size_y rows of size_x],
                                        each 1-D cdf without its leading 0 (CDF::computeForImage, src/runtime/CDF.cpp:70-150) |
                          ENV_TEX: scale rgb, transform 3x3 column major (9), texture (int32 bits) */
} igb200_light;

/* make_perspective_camera(eye, dir, up, compute_scale_from_{h,v}fov(fov, aspect), w, h, tmin, tmax):
 * src/runtime/camera/PerspectiveCamera.cpp:26-67, src/artic/camera/perspective.art:2-42 */
typedef struct igb200_camera {
    float   eye[3], dir[3], up[3];
    float   fov;          /* radians, as printed into the shader text */
    int32_t fov_vertical; /* 0: horizontal fov */
    float   aspect;       /* <= 0: settings.width / settings.height */
    float   tmin, tmax;   /* near / far clip */
} igb200_camera;

/* Light selectors (src/artic/light/light_selector.art; chosen by technique.light_selector, LoaderLight.cpp:423-452) */
#define IGB200_SELECTOR_UNIFORM   0  /* make_uniform_light_selector :26-44 */
#define IGB200_SELECTOR_CDF       1  /* "simple": make_cdf_light_selector :46-77 over light_cdf.bin (flux CDF, CDF.cpp:14-44) */
#define IGB200_SELECTOR_HIERARCHY 2  /* make_hierarchy_light_selector :79-110 over light_hierarchy.bin (LightHierarchy.cpp:47-129) */

/* make_path_renderer(max_depth, min_depth, light_selector, aovs, clamp, nee): PathTechnique.cpp:35-79 */
typedef struct igb200_technique {
    int32_t max_depth, min_depth;
    float   clamp;
    int32_t nee;
    int32_t light_selector;              /* IGB200_SELECTOR_* */
} igb200_technique;

/* What IRenderDevice::assignScene receives (SceneDatabase + entity_per_material, IRenderDevice.h:22-28,
 * src/runtime/table/SceneDatabase.h:8-21) plus the per-stage descriptors. */
typedef struct igb200_scene_desc {
    const float*               entities;      /* FixTables["entities"]: n_entities x 36 f32 (LoaderEntity.cpp:150-162) */
    int32_t                    n_entities;
    const igb200_lookup_entry* shape_lookups; /* DynTables["shapes"] lookups */
    int32_t                    n_shapes;
    const uint8_t*             shape_data;    /* DynTables["shapes"] data (TriMeshProvider.cpp:575-596, SphereProvider.cpp:42-47) */
    uint64_t                   shape_data_bytes;
    const igb200_entity_leaf*  leaves;        /* SceneBVHs[*].Leaves of all providers, concatenated */
    int32_t                    n_leaves;
    const int32_t*             entity_per_material; /* Runtime.cpp:306-310 */
    int32_t                    n_materials;
    const igb200_material*     materials;     /* n_materials */
    const igb200_light*        infinite_lights;
    int32_t                    n_infinite;
    const igb200_light*        finite_lights;
    int32_t                    n_finite;
    igb200_camera              camera;
    igb200_technique           technique;
    float                      bbox_min[3], bbox_max[3]; /* __scene_bbox_lower / upper */
    /* The buffer the selector's constructor receives (`device.load_buffer(".../light_cdf.bin" | ".../light_hierarchy.bin")`), as
     * 32-bit words, NULL / 0 for the uniform selector. cdf: n_finite floats [x1 .. x(n-1), 1]. hierarchy: u32 codes[round_up(n_finite, 4)],
     * then 8 words per tree node {pos xyz, flux (negative: no direction), dir xyz, id (>= 0 light, < 0: -(left child + 1))}. */
    const float*               selector_data;
    int32_t                    n_selector_data;
    /* textures, images and further read-only buffers the descriptors above refer to by index / offset */
    const igb200_texture*      textures;
    int32_t                    n_textures;
    const igb200_image*        images;
    int32_t                    n_images;
    const float*               aux_data;      /* 32-bit words: the 2-D cdfs of ENV_TEXTURED lights */
    int32_t                    n_aux_data;
} igb200_scene_desc;

/* `Settings`, src/artic/driver/settings.art:2-11, filled by the device at src/device/Device.cpp:384-397 */
typedef struct igb200_settings {
    int32_t device, thread_count, spi, frame, iter, width, height, seed;
} igb200_settings;

/* `StreamRay`, src/artic/traversal/ray.art:2-7 */
typedef struct igb200_ray {
    float org[3], dir[3], tmin, tmax;
} igb200_ray;

typedef struct igb200_hit {
    int32_t ent_id, prim_id; /* -1: no hit */
    float   t, u, v;
} igb200_hit;

/* ---- entry points (each names the IRenderDevice / Device.cpp member it stands for) ------------------------ */

const char* igb200_last_error(void);
int igb200_version(int* major, int* minor);                       /* IDeviceInterface::getVersion, Interface.cpp:24-27 */

int igb200_device_count(int* count);                                  /* sm_100 devices visible to this process (0 is not an error); Device.cpp:1632 knows exactly one */
int igb200_create(int cuda_device, igb200_ctx** out);             /* IDeviceInterface::createRenderDevice, Interface.cpp:34-57 */
int igb200_destroy(igb200_ctx* ctx);                              /* IRenderDevice::~IRenderDevice */

int igb200_set_scene(igb200_ctx* ctx, const igb200_scene_desc* scene); /* IRenderDevice::assignScene, Device.cpp:1667-1670 */
int igb200_resize(igb200_ctx* ctx, int width, int height);        /* IRenderDevice::resize, Device.cpp:1692-1695 */

/* Multi-GPU: this context renders only its share of the tile_size x tile_size framebuffer tiles: the tile in tile column tx and
 * tile row ty belongs to rank (tx + ty) mod world. No reference counterpart (the reference is single-device, Device.cpp:1632). */
int igb200_set_partition(igb200_ctx* ctx, int rank, int world, int tile_size);

/* ---- multi-GPU exchange inside the device (the reference is single-device, Device.cpp:1632) --------------------------------
 * One context per GPU, one process (or thread) per context. The ranks' contexts form an NCCL communicator: rank 0 makes a unique id
 * (ncclGetUniqueId), the caller carries the 128 bytes to every rank by whatever means it has (file, socket, MPI, torch.distributed),
 * every rank calls igb200_comm_init, which also sets the tile partition (as igb200_set_partition). igb200_comm_gather_framebuffer is the
 * path's ONE exchange step: every rank packs the pixels of its own tiles (a kernel on the context's stream), sends them to rank 0
 * (ncclSend / ncclRecv over NVLink, enqueued on the same stream right behind the render kernels and the drain of their deferred tail), and
 * rank 0 unpacks them into a complete frame that it owns separately from its own accumulation buffer. Collective: every rank must call it.
 * NCCL is loaded at run time (libnccl.so.2: the copy already in the process if there is one, else $IGB200_NCCL_LIB, else the system's). */
int igb200_comm_unique_id(uint8_t id[128]);
int igb200_comm_init(igb200_ctx* ctx, int rank, int world, int tile_size, const uint8_t id[128]);
/* Asynchronous on the context's stream. Rank 0: *device_frame = the gathered frame (W*H*3 floats, valid until the next gather or resize);
 * if host_frame is non-NULL the frame is also copied to context-owned pinned memory and the call waits for it. Other ranks: both NULL. */
int igb200_comm_gather_framebuffer(igb200_ctx* ctx, const char* aov, float** device_frame, float** host_frame);
int igb200_comm_destroy(igb200_ctx* ctx);

/* ---- frame streaming: every iteration's accumulated frame delivered to the host while later iterations render ---------------------------
 * igb200_framebuffer is synchronous, as IRenderDevice::getFramebufferForHost is: it finishes the deepest paths of the last iteration (a
 * latency-bound tail of ~50 tiny wavefront turns) and copies 12 bytes per pixel while the GPU idles. A caller that wants EVERY iteration's
 * frame (a progressive viewer, a denoiser feeding on intermediate frames, bench.py's end-to-end leg) streams them instead:
 * after igb200_frame_stream_begin each iteration accumulates into its own slot of a ring of framebuffers; once an iteration is known to be
 * finished (from the launches issued since -- no read-back) its slot is folded into the accumulated frame, in order, and a snapshot of the
 * frame travels to pinned host memory on a copy stream (with a communicator: after the tile gather, on rank 0) while the next iterations
 * render. igb200_frame_stream_next hands the frames out in iteration order: frame k = the sum of iterations 0..k, exactly what
 * igb200_framebuffer would have returned after render(k). wait: 0 = only if one is ready (a poll: render() calls still queued for a fused
 * launch stay queued unless the device has nothing else in the works), 1 = block until the oldest outstanding frame
 * arrives (returns 0 if none is outstanding), 2 = first finish everything rendered so far, then as 1. Returns 1 with a frame (valid until the
 * next call), 0 without. Frames exist on rank 0 only; the other ranks call it all the same (the gather is collective) and get 0. */
int igb200_frame_stream_begin(igb200_ctx* ctx, int slots /* iterations in flight, rounded up to a power of two; 0 = 16 */);
/* Several ranks on ONE host: the streamed frames live in a System V shared-memory segment (key > 0, the same on every rank; rank 0 creates it in
 * igb200_frame_stream_begin, the others attach) that every rank pins and maps. Each rank then writes the pixels of its own tiles straight into
 * the host frame over its own PCIe link -- no exchange between the GPUs, no full frame through one link -- and rank 0's
 * igb200_frame_stream_next hands out a frame once every rank has flagged it. Call on every rank before igb200_frame_stream_begin; 0 = off
 * (the frames are gathered onto rank 0's GPU and copied from there). */
int igb200_frame_stream_share(igb200_ctx* ctx, int key);
int igb200_frame_stream_next(igb200_ctx* ctx, int wait, int* iteration, float** host_rgb);
int igb200_frame_stream_end(igb200_ctx* ctx);

/* One iteration: IRenderDevice::render, Device.cpp:1672-1682. `rays` non-null selects the list emitter of igtrace
 * (Runtime::trace, Runtime.cpp:389-446): width = n_rays, height = 1. Accumulates into the device framebuffer.
 * The call is ASYNCHRONOUS on the context's stream (the reference's GPU device ends every iteration with acc.sync(),
 * driver/mapping_gpu.art:865): it returns once the iteration's kernel is enqueued, and the kernel itself may leave the
 * last few deep paths of the iteration to be finished together with the next one ("deferred tail"). When one
 * iteration is too small to fill the GPU (a rank's share of the frame in a multi-GPU run), consecutive calls with the same
 * settings and iter = previous + 1 are collected and their camera rays generated by ONE launch ("fused iterations").
 * Every entry point that observes results (igb200_framebuffer*, igb200_stats, igb200_sync, ...) or changes what
 * in-flight paths refer to (scene, size, partition, spi, seed) first finishes all outstanding paths, so the
 * observable behaviour is that of a synchronous render. */
int igb200_render(igb200_ctx* ctx, const igb200_settings* settings, const igb200_ray* rays, size_t n_rays);
/* Finishes every outstanding path of earlier igb200_render calls and waits for the device. */
int igb200_sync(igb200_ctx* ctx);

/* IRenderDevice::getFramebufferForHost, Device.cpp:1419-1451: copies the device framebuffer into a context-owned
 * pinned host buffer (RGB f32, width*height*3, row-major, sum over iterations). aov NULL/""/"Color" = main image;
 * "Normals" / "Albedo" with the option "std_aovs". */
int igb200_framebuffer(igb200_ctx* ctx, const char* aov, float** host_ptr);
int igb200_framebuffer_device(igb200_ctx* ctx, const char* aov, float** device_ptr); /* getFramebufferForDevice */
int igb200_clear(igb200_ctx* ctx, const char* aov_or_null);       /* clearFramebuffer / clearAllFramebuffer */
int igb200_upload_framebuffer(igb200_ctx* ctx, const char* aov, const float* host_rgb); /* syncFramebufferHostToDevice */

/* IRenderDevice::getStatistics (ray counters of src/runtime/Statistics.h:57-64): out = {camera rays, shadow rays,
 * bounce rays, framebuffer splats, kernel launches} since the last reset; render_ms = device time spent in the
 * kernels of igb200_render (sum of their durations, measured on the device with %globaltimer) since the last reset. */
int igb200_stats(igb200_ctx* ctx, uint64_t out[5], double* render_ms);
int igb200_reset_stats(igb200_ctx* ctx);

/* Time spent in the two phases of the wavefront turns since the last reset, measured on the device (%globaltimer at the
 * grid barriers of the persistent kernel and at the start of the split-turn kernels): out_ms = {0, trace phases (closest +
 * any hit), shade + generate phases, 0}, out_launches = number of phases. */
int igb200_kernel_times(igb200_ctx* ctx, double out_ms[4], uint64_t out_launches[4]);
/* With the option "profile_kernels" = 1 every kernel launch of igb200_render is bracketed by CUDA events on the context's
 * stream; this returns, since the last reset, the summed durations and launch counts per kernel (0 k_wavefront, 1 k_turn_trace,
 * 2 k_turn_shade, 3 k_turn_end) and the work the k_turn_trace launches did: {primary rays, shadow rays, framebuffer splats}. */
int igb200_launch_profile(igb200_ctx* ctx, double ms[4], uint64_t launches[4], uint64_t split_work[3]);
/* Diagnostics of the LAST igb200_render: for each loop turn of the persistent kernel (at most max_turns, at most 128)
 * the number of rays traced and the duration of its trace phase and of the shade + generate phase before it. */
int igb200_turn_log(igb200_ctx* ctx, uint32_t* items, uint32_t* trace_ns, uint32_t* shade_ns, int max_turns, int* n_turns);
/* Diagnostics build only (-DIGB_STEP_STATS; zeros otherwise), LAST igb200_render: for loop turns < 16 (out[0..7]) and
 * >= 16 (out[8..15]): inner-node visits, triangle-leaf visits, entity visits, max visits of one ray, rays traced. */
int igb200_step_stats(igb200_ctx* ctx, uint64_t out[16]);
/* Tunables: "capacity" (records per ray queue), "refill" (lanes), "stage_budget" (bytes of shared memory for the staged
 * scene copy; used only if the whole scene fits, unless "stage_partial" = 1; "specialise_where" = 0 disables the trace kernel built for a fully staged scene; "carveout" = preferred shared-memory carve-out of the trace kernels in percent, -1 = the driver's choice), "min_blocks" (2|3 CTAs per SM), "vote" (0|2), "defer_permille" (deferred tail threshold, 0 = off), "split_turns"
 * (leading turns run as separate shade / trace launches), "turn_trace_blocks" (2|3), "wide_rays_per_group", "fuse" (iterations per launch, 0 = automatic), "profile_kernels", "std_aovs" (1: the "Normals" and
 * "Albedo" AOVs of the reference's infobuffer wrapper, technique/internal/infobuffer.art, exist and are written at iteration 0),
 * "deterministic" (1: radiance is accumulated per SAMPLE and folded into the pixel in sample order, so the same inputs give the same bits
 * whatever the spi and the scheduling -- the reference's CPU device adds without atomics, driver/accumulator.art:4-21, and its
 * src/tests/integrator/test_reproducibility.py:5-11 expects identical arrays; costs 12 B x W x H x spi of slots, the deferred tail and
 * fused iterations; excludes frame streaming), "flat" (0: never walk the merged single-level tree of small scenes), "flat_block"
 * (256 | 384 | 768: merged-tree trace kernel without / with its ray records staged through shared memory by TMA), "bin_materials"
 * (-1 automatic | 0 | 1: shade through per-material-class index lists), "shade_sync" (1: the split-turn shade kernel walks
 * the queue CTA by CTA with a block barrier per record, so that the warps of a CTA share the instruction lines they fetch), "wave_skip"
 * (1: a render() whose split turns generated every camera ray and left at most the deferred-tail threshold of paths alive does not run
 * a turn of the persistent kernel; the paths ride along with the next render() or are finished by whatever observes results). */
int igb200_set_option(igb200_ctx* ctx, const char* name, int64_t value);
/* BVH construction. Every trimesh shape gets a BVH8 inside igb200_set_scene: by the host builder (binned SAH,
 * csrc/bvh8.h) or ON THE GPU (Morton codes -> sort -> Karras radix tree -> bottom-up boxes -> collapse to BVH8, csrc/bvh_build.cu)
 * -- option "gpu_bvh": 1 = every shape of more than 4 faces, 0 = never, -1 (default) = shapes of at least "gpu_bvh_min_faces" faces
 * (default 2^20); replaces the reference's CPU build through madmann91/bvh, src/runtime/shape/TriMeshProvider.cpp:255-298,
 * src/runtime/bvh/NArityBvh.h:94-143. Renders do not depend on which builder made a tree (the closest hit is a pure function of the ray).
 * igb200_set_cache_dir names a directory in which the trees of shapes with more than "bvh_cache_min_faces" faces (default 500 000, the
 * reference's MinFaceCountForCache, TriMeshProvider.cpp:326-351) are kept as bvh8_<hash of the builder's input>.bin and loaded instead of
 * being rebuilt; NULL / "" = no cache (the default). The directory must exist. Corrupt or foreign files are ignored and rebuilt. */
int igb200_set_cache_dir(igb200_ctx* ctx, const char* dir);
/* What the LAST igb200_set_scene did: out = {shapes built on the host, built on the GPU, loaded from the cache, stored into the cache,
 * microseconds spent on the shapes' trees, BVH8 nodes of all shapes}. */
int igb200_scene_build_info(igb200_ctx* ctx, int64_t out[6]);
/* The CUDA stream (cudaStream_t) every kernel and copy of this context is issued on, so that a caller can record
 * its own events on it or order a collective after a render (the reference has one implicit device queue). */
int igb200_stream(igb200_ctx* ctx, void** cuda_stream);

/* Parity / micro-benchmark hooks on the trace phase of the pipeline: closest hit for a host ray list (ray type
 * flags may be NULL = camera rays) and any hit (always shadow rays, through the shadow queue and the fused splat). */
int igb200_trace_closest(igb200_ctx* ctx, const igb200_ray* rays, const uint32_t* flags, size_t n, igb200_hit* out);
int igb200_trace_any(igb200_ctx* ctx, const igb200_ray* rays, size_t n, int32_t* occluded);
/* Same, device-resident rays (n repeated `repeat` times), returns average milliseconds per pass. */
int igb200_bench_trace(igb200_ctx* ctx, const igb200_ray* rays, size_t n, int any_hit, int repeat, double* ms_per_pass);

/* Test hook: builds a BVH8 over n boxes (lo xyz, hi xyz) with the host builder (builder 0; ctx may be NULL, no GPU needed) or the GPU
 * builder (1), round-trips it through the cache directory when dir is given, validates the tree (every primitive in exactly one leaf,
 * every box contains its subtree) and returns out = {nodes, levels, leaves, 1000 x sum of child half-areas / root half-area}. */
int igb200_test_bvh_build(igb200_ctx* ctx, const float* boxes6, size_t n, int builder, const char* dir, int64_t out[4]);
/* Test hook: evaluates the device's deterministic transcendental (0 sin, 1 cos, 2 acos, 3 atan2(a,b)) on the GPU. */
int igb200_test_detmath(igb200_ctx* ctx, int fn, const float* a, const float* b, float* out, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* IGB200_H */
