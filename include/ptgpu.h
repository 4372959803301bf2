/* ptgpu.h — C ABI of libptgpu, the sm_100a CUDA implementation of PTSharp's Renderer.Render /
 * RenderParallel / IterativeRender hot path.
 *
 * The reference has no plugin or FFI seam on the render path.  The only FFI in the repository is
 * the OpenImageDenoise P/Invoke block (PTSharpCore/OIDN.cs:43-95); this header follows its conventions: cdecl,
 * opaque handle, create / commit(upload) / execute / release verbs, errors polled as a string
 * (oidnGetDeviceError, OIDN.cs:75-76), caller-owned host arrays borrowed for the duration of one call.
 *
 * What each entry point replaces in the reference (paths relative to PTSharpCore/):
 *   ptgpu_upload_scene      the object graph Scene.Compile() leaves behind (Scene.cs:48-68: Shapes[], Lights[],
 *                           Tree built by Tree.cs:201-265, per-mesh trees Mesh.cs:45-57), flattened by the host
 *   ptgpu_render_pass       one call of Renderer.RenderParallel / Render (Renderer.cs:199-338 / :80-198): per pixel
 *                           spp x (Camera.CastRay + DefaultSampler.Sample) then Buffer.AddSample (Buffer.cs:94-97)
 *   ptgpu_accumulate_device the same pass without the Buffer update — the multi-GPU building block (sum buffers of
 *                           several ranks are reduced, then one rank calls ptgpu_add_sample_device)
 *   ptgpu_read_buffer       Buffer.Color / Variance / StandardDeviation / Samples (Buffer.cs:126-132)
 *   ptgpu_intersect_batch   Scene.Intersect (Scene.cs:75-79) + Hit.Info (Hit.cs:26-55) on caller-supplied rays
 *   ptgpu_cast_rays         Camera.CastRay (Camera.cs:98-119) on caller-supplied pixel/sample indices
 *   ptgpu_get_counters      Scene.rays (Scene.cs:15,77) split into path segments and shadow rays
 *
 * All functions return 0 on success, a negative PTGPU_E_* code otherwise; ptgpu_last_error gives the text.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with PTGPU_E_CUDA.
 */
#ifndef PTGPU_H
#define PTGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTGPU_ABI_VERSION 2
#define PTGPU_MAX_DEVICES 8

enum {
    PTGPU_OK = 0,
    PTGPU_E_ARG = -1,     /* bad argument / unsupported scene feature */
    PTGPU_E_CUDA = -2,    /* CUDA runtime error or no device */
    PTGPU_E_STATE = -3,   /* call order (e.g. render before upload) */
    PTGPU_E_LIMIT = -4    /* a compiled-in limit was exceeded (tree depth, SDF program size, ...) */
};

/* Shape type codes (IShape implementations on the path). */
enum {
    PTGPU_SPHERE = 1, PTGPU_CUBE = 2, PTGPU_PLANE = 3, PTGPU_CYLINDER = 4, PTGPU_TRIANGLE = 5, PTGPU_MESH = 6,
    PTGPU_TRANSFORMED = 7, PTGPU_SDF = 8, PTGPU_VOLUME = 9, PTGPU_SH = 10
};
/* LightMode.cs, SpecularMode.cs, BounceType.cs, Axis.cs — the integer codes are part of the ABI. */
enum { PTGPU_LIGHT_RANDOM = 0, PTGPU_LIGHT_ALL = 1 };
enum { PTGPU_SPECULAR_NAIVE = 0, PTGPU_SPECULAR_FIRST = 1, PTGPU_SPECULAR_ALL = 2 };
enum { PTGPU_AXIS_NONE = 0, PTGPU_AXIS_X = 1, PTGPU_AXIS_Y = 2, PTGPU_AXIS_Z = 3 };

/* ---- flat scene -------------------------------------------------------------------------------------------- */

/* kd-tree node (Tree.cs:44-65), 16 bytes, one 128-bit load.
 * interior: split = Node.Point, a = (left << 2) | axis(1..3), b = right   (indices into nodes[])
 * leaf:     a = (first << 2) | 0, b = count                                (range of leafItems[])            */
typedef struct ptgpu_node { double split; uint32_t a; uint32_t b; } ptgpu_node;

/* Tree (Tree.cs:8-42): Box = Box.BoxForShapes, Root.  maxDepth lets the device size-check its stack. */
typedef struct ptgpu_tree { float bmin[3]; uint32_t root; float bmax[3]; uint32_t maxDepth; } ptgpu_tree;

/* Entry of Scene.Shapes / inner shape of a TransformedShape.  data indexes the per-type array.
 * flags bit0: the C# type is a class, so `hit.Shape != light` (Sampler.cs:264) can be false.     */
typedef struct ptgpu_shape { uint32_t type; uint32_t data; int32_t material; uint32_t flags; } ptgpu_shape;

typedef struct ptgpu_sphere { float center[3]; float pad; double radius; double pad2; } ptgpu_sphere;  /* Sphere.cs:7-8 */
typedef struct ptgpu_cube { float min[3]; float max[3]; } ptgpu_cube;                                  /* Cube.cs:7-8 */
typedef struct ptgpu_plane { float point[3]; float normal[3]; } ptgpu_plane;                           /* Plane.cs:7-8 */
typedef struct ptgpu_cylinder { double radius, z0, z1; } ptgpu_cylinder;                               /* Cylinder.cs:7-8 */
/* Mesh.cs:8-10: triangles [triFirst, triFirst+triCount) of the global triangle arrays, own tree. */
typedef struct ptgpu_mesh { uint32_t tree; uint32_t triFirst; uint32_t triCount; uint32_t pad; } ptgpu_mesh;
/* Triangle.cs:11-13 split into what Intersect needs (V1, e1 = fl(V2-V1), e2 = fl(V3-V1): Triangle.cs:97-98) ... */
typedef struct ptgpu_tri_geom { float v1[3]; float pad0; float e1[3]; float pad1; float e2[3]; float pad2; } ptgpu_tri_geom;
/* ... and what NormalAt/UVector/MaterialAt need (Triangle.cs:128-189). */
typedef struct ptgpu_tri_shade { float n1[3], n2[3], n3[3]; float t1[2], t2[2], t3[2]; int32_t material; } ptgpu_tri_shade;
/* TransformedShape.cs:11-13: Matrix, Matrix.Inverse() (Matrix.cs:196-217, evaluated once on the host in the same
 * operation order), inner shape (index into shapes[]).  The inner shape may itself be a TransformedShape (up to four levels): pad[0]
 * is then 1 and the device re-measures Hit.T level by level as the nested Intersect calls do.                  */
typedef struct ptgpu_instance { double m[16]; double inv[16]; uint32_t shape; uint32_t pad[3]; } ptgpu_instance;

/* SDF.cs node types compiled by the host into a linear program (pre-order, explicit point stack). */
enum {
    PTGPU_SDF_SPHERE = 1,     /* p: radius, exponent                          SDF.cs:112-139 */
    PTGPU_SDF_CUBE = 2,       /* p: size.xyz                                  SDF.cs:141-195 */
    PTGPU_SDF_CYLINDER = 3,   /* p: radius, height                            SDF.cs:197-252 */
    PTGPU_SDF_CAPSULE = 4,    /* p: a.xyz, b.xyz, radius, exponent            SDF.cs:254-285 */
    PTGPU_SDF_TORUS = 5,      /* p: major, minor, majorExp, minorExp          SDF.cs:287-319 */
    PTGPU_SDF_PUSH_TRANSFORM = 6, /* p: Inverse[16]; point <- Inverse.MulPosition(point)   SDF.cs:340-344 */
    PTGPU_SDF_PUSH_SCALE = 7,     /* p: factor; point <- point.DivScalar(factor)           SDF.cs:371-374 */
    PTGPU_SDF_PUSH_REPEAT = 8,    /* p: step.xyz; point <- point.Mod(step)-step/2          SDF.cs:549-553 */
    PTGPU_SDF_POP = 9,            /* restore point; n = 1: multiply top value by p[0] (ScaleSDF) */
    PTGPU_SDF_UNION = 10, PTGPU_SDF_DIFFERENCE = 11, PTGPU_SDF_INTERSECTION = 12 /* n = item count  SDF.cs:398-509 */
};
typedef struct ptgpu_sdf_op { uint32_t op; uint32_t n; double p[16]; } ptgpu_sdf_op;
typedef struct ptgpu_sdf_shape { uint32_t progFirst; uint32_t progCount; float bmin[3]; float bmax[3]; } ptgpu_sdf_shape;

/* Volume.cs:22-26 */
typedef struct ptgpu_volume_window { double lo, hi; int32_t material; int32_t pad; } ptgpu_volume_window;
typedef struct ptgpu_volume {
    int32_t w, h, d; uint32_t windowFirst; uint32_t windowCount; uint32_t pad;
    double zscale; uint64_t dataOffset; float bmin[3]; float bmax[3];
} ptgpu_volume;

/* SphericalHarmonic (SH.cs:7-104): Intersect walks `mesh` (the marching-cubes mesh of |p| - |Y_l^m(p/|p|)|, built on the host:
 * MC.NewSDFMesh, MC.cs:9-66) but the Hit names the SphericalHarmonic itself: NormalAt is the gradient of that function (SH.cs:74-86),
 * MaterialAt picks by the sign of Y_l^m (SH.cs:62-72). */
typedef struct ptgpu_sh { int32_t l, m; uint32_t mesh; int32_t positiveMaterial, negativeMaterial; int32_t pad[3]; } ptgpu_sh;

/* Material.cs:11-45.  Texture ids index textures[], -1 = null. */
typedef struct ptgpu_material {
    double color[3]; double bumpMultiplier, emittance, index, gloss, tint, reflectivity;
    int32_t transparent, texture, normalTexture, bumpTexture, glossTexture, pad;
} ptgpu_material;

/* ColorTexture (Texture.cs:96-100): Width, Height, Data (already Pow(2.2)'d, Texture.cs:163) as 4 doubles per texel (r, g, b, 1): the
 * reference's Colour is 3 doubles, and FP32 texels perturb a normal-mapped / bump-mapped normal in its last bits, which is enough to
 * flip the self-intersection of the next ray: texels keep the reference's precision. */
typedef struct ptgpu_texture { int32_t width, height; uint64_t texelOffset; } ptgpu_texture;

typedef struct ptgpu_flat_scene {
    uint32_t abiVersion;
    uint32_t sceneTree;                       /* Scene.tree: leaf items are indices into shapes[0..numSceneShapes) */
    uint32_t numSceneShapes;                  /* Scene.Shapes.Length; shapes[] continues with nested inner shapes   */
    uint32_t numShapes;        const ptgpu_shape* shapes;
    uint32_t numLights;        const uint32_t* lights;       /* Scene.Lights as indices into shapes[] (Scene.cs:33) */
    uint32_t numTrees;         const ptgpu_tree* trees;
    uint64_t numNodes;         const ptgpu_node* nodes;
    uint64_t numLeafItems;     const uint32_t* leafItems;    /* scene tree: shape index; mesh tree: global triangle index */
    uint32_t numSpheres;       const ptgpu_sphere* spheres;
    uint32_t numCubes;         const ptgpu_cube* cubes;
    uint32_t numPlanes;        const ptgpu_plane* planes;
    uint32_t numCylinders;     const ptgpu_cylinder* cylinders;
    uint32_t numMeshes;        const ptgpu_mesh* meshes;
    uint64_t numTriangles;     const ptgpu_tri_geom* triGeom; const ptgpu_tri_shade* triShade;
    uint32_t numInstances;     const ptgpu_instance* instances;
    uint32_t numSdfShapes;     const ptgpu_sdf_shape* sdfShapes;
    uint32_t numSdfOps;        const ptgpu_sdf_op* sdfOps;
    uint32_t numVolumes;       const ptgpu_volume* volumes;
    uint32_t numVolumeWindows; const ptgpu_volume_window* volumeWindows;
    uint64_t numVolumeData;    const double* volumeData;
    uint32_t numMaterials;     const ptgpu_material* materials;
    uint32_t numTextures;      const ptgpu_texture* textures;
    uint64_t numTexels;        const double* texels;         /* 4 doubles per texel */
    double envColor[3];                                      /* Scene.Color (Scene.cs:11) */
    int32_t envTexture; int32_t pad0;                        /* Scene.Texture (Scene.cs:12) */
    double envTextureAngle;                                  /* Scene.TextureAngle (Scene.cs:13) */
    uint32_t numShs;           const ptgpu_sh* shs;          /* SphericalHarmonic shapes */
} ptgpu_flat_scene;

/* Camera.cs:11-17 after LookAt/SetFocus. */
typedef struct ptgpu_camera { float p[3], u[3], v[3], w[3]; double m, focalDistance, apertureRadius; } ptgpu_camera;

/* One pass = one call of RenderParallel with these Renderer / DefaultSampler fields (Renderer.cs:21-31,
 * Sampler.cs:13-18).  Sample k of a pixel (k in [0, spp)) is global sample sampleBase + k*sampleStride: ranks of a
 * multi-GPU job pass (rank, world) so that the union of their samples is the same set a single GPU would draw. */
typedef struct ptgpu_pass {
    int32_t width, height;
    int32_t spp;                 /* Renderer.SamplesPerPixel */
    int32_t stratified;          /* Renderer.StratifiedSampling: spp is floored to a square, strata centres, no jitter */
    int32_t sampleBase, sampleStride;
    int32_t firstHitSamples, maxBounces, directLighting, softShadows, lightMode, specularMode;  /* DefaultSampler */
    uint32_t seed, passIndex;    /* Philox key */
    ptgpu_camera camera;
    /* Extra passes of RenderParallel (Renderer.cs:340-468), run by ptgpu_render_pass after the main pass; 0 = off.
     * adaptiveSamples: that many more samples for EVERY pixel, uniform sub-pixel jitter, each its own Buffer.AddSample
     *   (the parallel path is not adaptive, Renderer.cs:349-364; its second, variance-only loop :376-388 changes nothing
     *   and is not rendered).
     * fireflySamples: for pixels whose StandardDeviation().MaxComponent() > fireflyThreshold (Renderer.cs:426), up to
     *   that many more samples, stopping at the first one IsFirefly() rejects (Renderer.cs:430-441, 474-497). */
    int32_t adaptiveSamples, fireflySamples;
    double fireflyThreshold;     /* Renderer.FireflyThreshold, 1 in NewRenderer (Renderer.cs:47) */
    /* serialRules != 0: the extra passes follow the serial Render() (Renderer.cs:150-191, what IterativeRender runs when
     * NumCPU == 1) instead of RenderParallel:
     *   adaptive: a pixel gets AdaptiveSamples * (int)pow(clamp(StandardDeviation().MaxComponent() / adaptiveThreshold, 0, 1),
     *     adaptiveExponent) more samples (fu, fv = xi: uniform sub-pixel jitter), i.e. AdaptiveSamples of them iff its deviation
     *     reaches the threshold (every pixel when the exponent is 0; a negative exponent is rejected with PTGPU_E_ARG);
     *   firefly: pixels above fireflyThreshold (evaluated after the adaptive samples) get fireflySamples more samples with
     *     fu = (x + xi) * (1.0f / w) - no IsFirefly() rejection in the serial path.
     * Each extra sample is its own Buffer.AddSample. */
    int32_t serialRules;
    /* PTGPU_PASS_* bits.  RUSSIAN_ROULETTE: the `russianRoulette` branch of DefaultSampler.sample (Sampler.cs:133-142), which the
     * reference never enables (its only caller passes the default `false`): at depth >= 4 a vertex survives with
     * probability p = min(max colour component of the material, 0.95) and its children carry 1/p.  Unbiased, NOT part of parity mode. */
    int32_t flags;
    double adaptiveThreshold;    /* Renderer.AdaptiveThreshold, 1 in NewRenderer (Renderer.cs:44) */
    double adaptiveExponent;     /* Renderer.AdaptiveExponent, 1 in NewRenderer (Renderer.cs:45) */
} ptgpu_pass;

enum { PTGPU_PASS_RUSSIAN_ROULETTE = 1 };

typedef struct ptgpu_params {
    int32_t device;              /* CUDA ordinal (used when numDevices == 0) */
    int32_t flags;               /* bits 0-3: number of lanes (independent streams + queues the batches of a pass go round), 0 = default (1) */
    uint64_t queueCapacity;      /* path records in flight over all lanes; 0 = default (2^27 ~ 64 spp of a 1920x1080 frame per batch; queues are allocated on demand, ~50 GB when full) */
    /* Multi-GPU inside the handle (SURVEY 8b "Threading", 8e): numDevices > 1 replicates the scene on devices[0..numDevices) and
     * ptgpu_render_pass splits the samples of every pixel over them (device k draws global samples sampleBase + (k + j*numDevices)
     * * sampleStride); devices[0] owns the Buffer and folds the peers' float sum buffers over NVLink peer access inside the
     * Buffer.AddSample kernel (one fused reduce + Welford per pass, no host copy).  A single-process host (the C# Renderer) thus
     * drives every GPU of the box through the same seven calls. */
    int32_t numDevices;          /* 0 or 1: single device */
    int32_t devices[PTGPU_MAX_DEVICES];
    int32_t reserved;
} ptgpu_params;

typedef struct ptgpu_counters {
    uint64_t cameraSamples;      /* Camera.CastRay + Sampler.Sample calls */
    uint64_t segments;           /* scene.Intersect calls from DefaultSampler.sample (Sampler.cs:62) */
    uint64_t shadowRays;         /* scene.Intersect calls from sampleLight (Sampler.cs:262) */
    uint64_t nanSamples;         /* contributions dropped because a component was NaN/Inf */
    uint64_t kernelLaunches;     /* this library's kernels launched since create/reset */
    double lastPassMs;           /* device time of the last render_pass / accumulate_device (CUDA events) */
    double traceMs, shadeMs, shadowMs, raygenMs;  /* per-stage device time of the last pass when profiling is on */
    double meshMs;               /* of traceMs + shadowMs: time in the mesh-walk kernel (k_mesh), 0 when the scene has no meshes */
    uint64_t meshItems;          /* Mesh.Intersect calls (work items of k_mesh) of the last profiled pass */
    uint64_t meshLaunches;       /* k_mesh launches of the last profiled pass */
    uint64_t traceLaunches;      /* k_scene_trace<START> launches of the last profiled pass (= depths traced) */
    uint64_t queueOverflows;     /* passes since create/reset in which a ray or shadow queue overflowed (records dropped: that pass is not added to the Buffer) */
    uint64_t devices;            /* GPUs behind this handle */
    double sdfMs;                /* like meshMs / meshItems / meshLaunches for k_march<SDF> (SDFShape.Intersect loops) ... */
    uint64_t sdfItems, sdfLaunches;
    double volumeMs;             /* ... and k_march<VOLUME> (Volume.Intersect loops) */
    uint64_t volumeItems, volumeLaunches;
} ptgpu_counters;

typedef struct ptgpu_ctx ptgpu_ctx;

int ptgpu_abi_version(void);
/* sizeof() of the ABI structs as this library was compiled, for bindings to check their marshalling against:
 * 0 ptgpu_params, 1 ptgpu_pass, 2 ptgpu_camera, 3 ptgpu_counters, 4 ptgpu_flat_scene, 5 ptgpu_node, 6 ptgpu_tree, 7 ptgpu_shape,
 * 8 ptgpu_sphere, 9 ptgpu_cube, 10 ptgpu_plane, 11 ptgpu_cylinder, 12 ptgpu_mesh, 13 ptgpu_tri_geom, 14 ptgpu_tri_shade,
 * 15 ptgpu_instance, 16 ptgpu_sdf_op, 17 ptgpu_sdf_shape, 18 ptgpu_volume_window, 19 ptgpu_volume, 20 ptgpu_material,
 * 21 ptgpu_texture, 22 ptgpu_sh; anything else: -1. */
int ptgpu_abi_sizeof(int which);
int ptgpu_create(const ptgpu_params* params, ptgpu_ctx** out);
void ptgpu_destroy(ptgpu_ctx* ctx);
const char* ptgpu_last_error(ptgpu_ctx* ctx);   /* ctx may be NULL: error of the last failed ptgpu_create */

int ptgpu_upload_scene(ptgpu_ctx* ctx, const ptgpu_flat_scene* scene);
uint64_t ptgpu_scene_bytes(ptgpu_ctx* ctx);     /* bytes of scene data resident on the device */

/* Render one pass and add it to the device-resident Buffer (Welford, Buffer.cs:33-44).  out_mean_rgb (host,
 * width*height*3 floats, may be NULL) receives this pass's per-pixel mean, i.e. the value handed to AddSample. */
int ptgpu_render_pass(ptgpu_ctx* ctx, const ptgpu_pass* pass, float* out_mean_rgb);

/* Multi-process multi-GPU building blocks (one rank per GPU, e.g. torchrun; a single-process host uses ptgpu_params.devices
 * instead).  d_sum_rgb is a DEVICE pointer (width*height*3 floats) the pass's radiance is ADDED to; stream is a cudaStream_t the
 * work is ordered on: it starts after everything already queued on that stream and the stream continues after it.  NULL = the
 * legacy default stream (cudaStreamLegacy, what torch's default stream is).  No Buffer update.  A queue overflow is reported by
 * ptgpu_get_counters().queueOverflows. */
int ptgpu_accumulate_device(ptgpu_ctx* ctx, const ptgpu_pass* pass, float* d_sum_rgb, void* stream);
/* Buffer.AddSample(x, y, sum/divisor) for every pixel; d_sum_rgb is a DEVICE pointer. */
int ptgpu_add_sample_device(ptgpu_ctx* ctx, int32_t width, int32_t height, const float* d_sum_rgb, double divisor, void* stream);

/* channel: 0 Color (Pixel.M), 1 Variance, 2 StandardDeviation, 3 Samples, 4 Albedo (Buffer.cs:240-282), 5 Normal
 * (Buffer.cs:99-124, 222-233) — the Channel enum of Buffer.cs:8-16.  out: w*h*3 floats. */
int ptgpu_read_buffer(ptgpu_ctx* ctx, int32_t channel, float* out_rgb);
/* Exact FP64 copy of the Welford state (Pixel.M, Pixel.V, Pixel.Samples; Buffer.cs:18-58) out of / into the device: a
 * checkpoint of a long IterativeRender loop (Example.cs:1696 runs 1000 iterations) and its resume.  Host pointers; any
 * of M_rgb / V_rgb / samples may be NULL on export. */
int ptgpu_export_buffer(ptgpu_ctx* ctx, int32_t* width, int32_t* height, double* M_rgb, double* V_rgb, int32_t* samples);
int ptgpu_import_buffer(ptgpu_ctx* ctx, int32_t width, int32_t height, const double* M_rgb, const double* V_rgb, const int32_t* samples);
int ptgpu_reset_buffer(ptgpu_ctx* ctx);

/* Test hooks (same device functions the pipeline uses). */
int ptgpu_intersect_batch(ptgpu_ctx* ctx, int32_t n, const float* o3, const float* d3, int32_t* shape, int32_t* prim,
                          double* t, float* normal3, float* position3, int32_t* inside, int32_t* material);
int ptgpu_cast_rays(ptgpu_ctx* ctx, const ptgpu_pass* pass, int32_t n, const int32_t* x, const int32_t* y,
                    const double* fu, const double* fv, const int32_t* sample, float* o3, float* d3);
/* One draw of the keyed Philox stream (for cross-checking stream addressing). */
int ptgpu_keyed_draw(ptgpu_ctx* ctx, uint32_t seed, uint32_t pass, uint32_t pixel, uint32_t sample, uint32_t bits,
                     uint32_t first, uint32_t depth, uint32_t sub, uint32_t drawIndex, double* out);
int ptgpu_get_counters(ptgpu_ctx* ctx, ptgpu_counters* out);
int ptgpu_reset_counters(ptgpu_ctx* ctx);
int ptgpu_set_profiling(ptgpu_ctx* ctx, int32_t on);   /* per-stage CUDA-event timing (adds syncs; off by default) */

#ifdef __cplusplus
}
#endif
#endif /* PTGPU_H */
