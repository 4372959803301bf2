// capi.cpp — C entry points of libpthost.so (prefix pth_) so that Python (tests, bench.py) can author scenes through
// the same C++ host classes a C++ application would use (ptsharp.hpp), obtain the flat scene, and drive the
// Renderer.  The authoring verbs take the arguments of the reference's factory methods (Sphere.NewSphere, ...).
#include <cstdio>
#include <stdexcept>

#include "ptsharp.hpp"

using namespace ptsharp;

struct pth_world {
    Scene scene;
    Camera camera;
    DefaultSampler sampler;
    std::vector<ITexture> textures;
    std::vector<Material> materials;
    std::vector<ShapePtr> shapes;
    std::vector<SDFPtr> sdfs;
    std::unique_ptr<FlatScene> flat;
    std::unique_ptr<Renderer> renderer;
    std::string error;
};

static Vector V3(const double* v) { return Vector(v[0], v[1], v[2]); }
static Matrix M16(const double* m) { Matrix r; std::memcpy(r.m, m, sizeof(r.m)); return r; }
static int push(pth_world* w, ShapePtr s) { w->shapes.push_back(std::move(s)); return (int)w->shapes.size() - 1; }
static int pushSdf(pth_world* w, SDFPtr s) { w->sdfs.push_back(std::move(s)); return (int)w->sdfs.size() - 1; }

extern "C" {

pth_world* pth_world_new() { return new pth_world(); }
void pth_world_free(pth_world* w) { delete w; }
const char* pth_last_error(pth_world* w) { return w->error.c_str(); }

int pth_texture(pth_world* w, int width, int height, const double* rgb) {
    auto t = std::make_shared<ColorTexture>();
    t->Width = width; t->Height = height; t->Data.resize((size_t)width * height);
    for (size_t i = 0; i < t->Data.size(); i++) t->Data[i] = Colour(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
    w->textures.push_back(t);
    return (int)w->textures.size() - 1;
}
int pth_material(pth_world* w, const double* color, int tex, int normalTex, int bumpTex, int glossTex, double bumpMultiplier,
                 double emittance, double index, double gloss, double tint, double reflectivity, int transparent) {
    auto T = [&](int id) -> ITexture { return id >= 0 ? w->textures[(size_t)id] : ITexture(); };
    w->materials.emplace_back(Colour(color[0], color[1], color[2]), T(tex), T(normalTex), T(bumpTex), T(glossTex), bumpMultiplier,
                              emittance, index, gloss, tint, reflectivity, transparent != 0);
    return (int)w->materials.size() - 1;
}
int pth_sphere(pth_world* w, const double* c, double r, int mat) { return push(w, Sphere::NewSphere(V3(c), r, w->materials[(size_t)mat])); }
int pth_cube(pth_world* w, const double* mn, const double* mx, int mat) { return push(w, Cube::NewCube(V3(mn), V3(mx), w->materials[(size_t)mat])); }
int pth_plane(pth_world* w, const double* p, const double* n, int mat) { return push(w, Plane::NewPlane(V3(p), V3(n), w->materials[(size_t)mat])); }
int pth_cylinder(pth_world* w, double r, double z0, double z1, int mat) { return push(w, Cylinder::NewCylinder(r, z0, z1, w->materials[(size_t)mat])); }
int pth_mesh(pth_world* w, int ntri, const float* V, const float* N, const float* T, const int* mats, int mat) {
    std::vector<Triangle> tris((size_t)ntri);
    for (int i = 0; i < ntri; i++) {
        Triangle& t = tris[(size_t)i];
        const float* v = V + (size_t)i * 9;
        t.V1 = Vector(v[0], v[1], v[2]); t.V2 = Vector(v[3], v[4], v[5]); t.V3 = Vector(v[6], v[7], v[8]);
        if (N) { const float* n = N + (size_t)i * 9; t.N1 = Vector(n[0], n[1], n[2]); t.N2 = Vector(n[3], n[4], n[5]); t.N3 = Vector(n[6], n[7], n[8]); }
        if (T) { const float* q = T + (size_t)i * 9; t.T1 = Vector(q[0], q[1], q[2]); t.T2 = Vector(q[3], q[4], q[5]); t.T3 = Vector(q[6], q[7], q[8]); }
        t.Mat = w->materials[(size_t)(mats ? mats[i] : mat)];
        t.FixNormals();
    }
    return push(w, Mesh::NewMesh(std::move(tris)));
}
// OBJ.Load / STL.Load (host/loaders.cpp): kind 0 = OBJ, 1 = STL.  Returns the shape id, or -1 with pth_last_error set.
int pth_load_model(pth_world* w, int kind, const char* path, int mat) {
    try {
        const Material& m = w->materials[(size_t)mat];
        return push(w, kind == 0 ? OBJ::Load(path, m) : STL::Load(path, m));
    } catch (const std::exception& e) { w->error = e.what(); return -1; }
}
// Mesh utilities.  op 0 SmoothNormals, 1 SmoothNormalsThreshold(a[0] radians), 2 MoveTo(position a[0..2], anchor a[3..5]),
// 3 FitInside(box min a[0..2], max a[3..5], anchor a[6..8]), 4 Transform(matrix a[0..15]), 5 SetMaterial(material (int)a[0])
int pth_mesh_op(pth_world* w, int shape, int op, const double* a) {
    Mesh* m = dynamic_cast<Mesh*>(w->shapes[(size_t)shape].get());
    if (!m) { w->error = "pth_mesh_op: shape is not a Mesh"; return -1; }
    try {
        switch (op) {
            case 0: m->SmoothNormals(); break;
            case 1: m->SmoothNormalsThreshold(a[0]); break;
            case 2: m->MoveTo(V3(a), V3(a + 3)); break;
            case 3: m->FitInside(Box(V3(a), V3(a + 3)), V3(a + 6)); break;
            case 4: m->Transform(M16(a)); break;
            case 5: m->SetMaterial(w->materials[(size_t)a[0]]); break;
            default: w->error = "pth_mesh_op: unknown op"; return -1;
        }
    } catch (const std::exception& e) { w->error = e.what(); return -1; }
    return 0;
}
// Triangles of a Mesh (V, N, T: 9 floats per triangle each; null = only count them).
int pth_mesh_get(pth_world* w, int shape, float* V, float* N, float* T) {
    Mesh* m = dynamic_cast<Mesh*>(w->shapes[(size_t)shape].get());
    if (!m) if (auto* sh = dynamic_cast<SphericalHarmonic*>(w->shapes[(size_t)shape].get())) m = sh->mesh.get();  // its marching-cubes mesh
    if (!m) { w->error = "pth_mesh_get: shape is not a Mesh"; return -1; }
    for (size_t i = 0; i < m->Triangles.size(); i++) {
        const Triangle& t = m->Triangles[i];
        const Vector* vs[3][3] = {{&t.V1, &t.V2, &t.V3}, {&t.N1, &t.N2, &t.N3}, {&t.T1, &t.T2, &t.T3}};
        float* outs[3] = {V, N, T};
        for (int a = 0; a < 3; a++) if (outs[a]) for (int k = 0; k < 3; k++) { outs[a][i * 9 + k * 3] = vs[a][k]->x; outs[a][i * 9 + k * 3 + 1] = vs[a][k]->y; outs[a][i * 9 + k * 3 + 2] = vs[a][k]->z; }
    }
    return (int)m->Triangles.size();
}
int pth_transformed(pth_world* w, int shape, const double* m16) { return push(w, TransformedShape::NewTransformedShape(w->shapes[(size_t)shape], M16(m16))); }
int pth_sdf_sphere(pth_world* w, double r) { return pushSdf(w, NewSphereSDF(r)); }
int pth_sdf_cube(pth_world* w, const double* size) { return pushSdf(w, NewCubeSDF(V3(size))); }
int pth_sdf_cylinder(pth_world* w, double r, double h) { return pushSdf(w, NewCylinderSDF(r, h)); }
int pth_sdf_capsule(pth_world* w, const double* a, const double* b, double r) { return pushSdf(w, NewCapsuleSDF(V3(a), V3(b), r)); }
int pth_sdf_torus(pth_world* w, double major, double minor) { return pushSdf(w, NewTorusSDF(major, minor)); }
int pth_sdf_transform(pth_world* w, int sdf, const double* m16) { return pushSdf(w, NewTransformSDF(w->sdfs[(size_t)sdf], M16(m16))); }
int pth_sdf_scale(pth_world* w, int sdf, double f) { return pushSdf(w, NewScaleSDF(w->sdfs[(size_t)sdf], f)); }
int pth_sdf_repeat(pth_world* w, int sdf, const double* step) { return pushSdf(w, NewRepeaterSDF(w->sdfs[(size_t)sdf], V3(step))); }
int pth_sdf_combine(pth_world* w, int op, int n, const int* items) {
    std::vector<SDFPtr> v;
    for (int i = 0; i < n; i++) v.push_back(w->sdfs[(size_t)items[i]]);
    return pushSdf(w, op == 0 ? NewUnionSDF(v) : op == 1 ? NewDifferenceSDF(v) : NewIntersectionSDF(v));
}
// MC.NewSDFMesh (MC.cs:9-66): the marching-cubes mesh of an SDF over a box; mat >= 0: Mesh.SetMaterial afterwards (the reference leaves
// `new Material()` on every triangle).  SphericalHarmonic.NewSphericalHarmonic (SH.cs:14-22); step <= 0: the reference's 0.01F.
int pth_mc_mesh(pth_world* w, int sdf, const double* bmin, const double* bmax, double step, int mat) {
    try {
        auto m = MC::NewSDFMesh(w->sdfs[(size_t)sdf], Box(V3(bmin), V3(bmax)), step);
        if (mat >= 0) m->SetMaterial(w->materials[(size_t)mat]);
        return push(w, m);
    } catch (const std::exception& e) { w->error = e.what(); return -1; }
}
int pth_spherical_harmonic(pth_world* w, int l, int m, int pm, int nm, double step) {
    try {
        return push(w, SphericalHarmonic::NewSphericalHarmonic(l, m, w->materials[(size_t)pm], w->materials[(size_t)nm], step > 0 ? step : (double)0.01f));
    } catch (const std::exception& e) { w->error = e.what(); return -1; }
}
int pth_mc_case(int index, int* out15) { return MC::CaseTriangles(index, out15); }
int pth_mc_edges(int index) { return MC::CaseEdges(index); }
int pth_sdf_shape(pth_world* w, int sdf, int mat) { return push(w, SDFShape::NewSDFShape(w->sdfs[(size_t)sdf], w->materials[(size_t)mat])); }
int pth_volume(pth_world* w, const double* bmin, const double* bmax, int W, int H, int D, double zscale, const double* data, int nwin,
               const double* lo, const double* hi, const int* mats) {
    std::vector<Volume::VolumeWindow> wins;
    for (int i = 0; i < nwin; i++) wins.push_back(Volume::VolumeWindow{lo[i], hi[i], w->materials[(size_t)mats[i]]});
    return push(w, Volume::NewVolume(Box(V3(bmin), V3(bmax)), W, H, D, zscale, std::vector<double>(data, data + (size_t)W * H * D), wins));
}
void pth_scene_add(pth_world* w, int shape) { w->scene.Add(w->shapes[(size_t)shape]); }
void pth_scene_env(pth_world* w, const double* color, int tex, double angle) {
    w->scene.Color = Colour(color[0], color[1], color[2]);
    w->scene.Texture = tex >= 0 ? w->textures[(size_t)tex] : ITexture();
    w->scene.TextureAngle = angle;
}
void pth_camera_lookat(pth_world* w, const double* eye, const double* center, const double* up, double fovy) { w->camera = Camera::LookAt(V3(eye), V3(center), V3(up), fovy); }
void pth_camera_focus(pth_world* w, const double* focalPoint, double aperture) { w->camera.SetFocus(V3(focalPoint), aperture); }
void pth_sampler(pth_world* w, int firstHit, int maxBounces, int directLighting, int softShadows, int lightMode, int specularMode) {
    w->sampler = DefaultSampler::NewSampler(firstHit, maxBounces);
    w->sampler.DirectLighting = directLighting != 0; w->sampler.SoftShadows = softShadows != 0;
    w->sampler.LightMode = (LightMode)lightMode; w->sampler.SpecularMode = (SpecularMode)specularMode;
}
void pth_compile(pth_world* w) { w->scene.Compile(); }

// Scene.Compile() + flatten.  The returned view (and everything it points to) lives until the world is freed or
// flattened again.
const ptgpu_flat_scene* pth_flatten(pth_world* w) {
    try {
        w->scene.Compile();
        w->flat = Flatten(w->scene);
        return &w->flat->view;
    } catch (const std::exception& e) { w->error = e.what(); return nullptr; }
}
uint64_t pth_flat_bytes(pth_world* w) { return w->flat ? w->flat->Bytes() : 0; }
// Flat-scene file I/O (SaveFlatScene / LoadFlatScene).  Load replaces the world's flat scene and returns its view.
int pth_save_flat(pth_world* w, const char* path) {
    try {
        if (!w->flat) throw std::runtime_error("flatten the scene first");
        SaveFlatScene(*w->flat, path);
        return 0;
    } catch (const std::exception& e) { w->error = e.what(); return -1; }
}
const ptgpu_flat_scene* pth_load_flat(pth_world* w, const char* path) {
    try {
        w->flat = LoadFlatScene(path);
        return &w->flat->view;
    } catch (const std::exception& e) { w->error = e.what(); return nullptr; }
}
// Fill a ptgpu_pass from the world's camera and sampler.
void pth_make_pass(pth_world* w, int width, int height, int spp, int stratified, unsigned seed, unsigned passIndex, int sampleBase,
                   int sampleStride, ptgpu_pass* out) {
    std::memset(out, 0, sizeof(*out));
    out->width = width; out->height = height; out->spp = spp; out->stratified = stratified;
    out->sampleBase = sampleBase; out->sampleStride = sampleStride;
    out->firstHitSamples = w->sampler.FirstHitSamples; out->maxBounces = w->sampler.MaxBounces;
    out->directLighting = w->sampler.DirectLighting; out->softShadows = w->sampler.SoftShadows;
    out->lightMode = w->sampler.LightMode; out->specularMode = w->sampler.SpecularMode;
    out->seed = seed; out->passIndex = passIndex;
    out->camera = FlattenCamera(w->camera);
    out->adaptiveSamples = 0; out->fireflySamples = 0; out->fireflyThreshold = 1.0;
    out->serialRules = 0; out->flags = w->sampler.RussianRoulette ? PTGPU_PASS_RUSSIAN_ROULETTE : 0; out->adaptiveThreshold = 1.0; out->adaptiveExponent = 1.0;
}

// kd-tree dump in the oracle's canonical pre-order form (builder parity tests).  which = -1: scene tree, else the
// tree of mesh shape `which` (authoring id).  Leaf items are local (triangle index in the mesh / index in Scene.Shapes).
static const Tree* pickTree(pth_world* w, int which) {
    w->scene.Compile();
    if (which < 0) return w->scene.tree.get();
    IShape* s = w->shapes[(size_t)which].get();
    if (s->Type() != PTGPU_MESH) return nullptr;
    s->Compile();
    return static_cast<Mesh*>(s)->tree.get();
}
int pth_tree_stats(pth_world* w, int which, long long* out4, float* box6) {
    const Tree* t = pickTree(w, which);
    if (!t) return -1;
    long long maxLeaf = 0;
    for (const ptgpu_node& n : t->nodes) if ((n.a & 3u) == 0 && (long long)n.b > maxLeaf) maxLeaf = n.b;
    out4[0] = (long long)t->nodes.size(); out4[1] = (long long)t->leafItems.size(); out4[2] = maxLeaf; out4[3] = t->maxDepth;
    if (box6) { box6[0] = t->box.Min.x; box6[1] = t->box.Min.y; box6[2] = t->box.Min.z; box6[3] = t->box.Max.x; box6[4] = t->box.Max.y; box6[5] = t->box.Max.z; }
    return 0;
}
int pth_tree_dump(pth_world* w, int which, int* axis, double* point, int* a, int* b, int* items) {
    const Tree* t = pickTree(w, which);
    if (!t) return -1;
    // the builder already emits nodes and leaf items in pre-order, with local indices
    for (size_t i = 0; i < t->nodes.size(); i++) {
        const ptgpu_node& n = t->nodes[i];
        axis[i] = (int)(n.a & 3u); point[i] = n.split; a[i] = (int)(n.a >> 2); b[i] = (int)n.b;
    }
    for (size_t i = 0; i < t->leafItems.size(); i++) items[i] = (int)t->leafItems[i];
    return 0;
}

// Triangle order for which the reference builder gives a balanced tree (BuilderFriendlyOrder, host.cpp).
// V: ntri*9 floats; perm_out[i] = index of the triangle to place at slot i.
void pth_builder_friendly_order(int ntri, const float* V, double balance, int minRepair, int verbose, int* perm_out) {
    std::vector<Box> boxes((size_t)ntri);
    for (int i = 0; i < ntri; i++) {
        Triangle t;
        const float* v = V + (size_t)i * 9;
        t.V1 = Vector(v[0], v[1], v[2]); t.V2 = Vector(v[3], v[4], v[5]); t.V3 = Vector(v[6], v[7], v[8]);
        boxes[(size_t)i] = t.BoundingBox();
    }
    std::vector<uint32_t> order = BuilderFriendlyOrder(boxes, balance, minRepair, verbose);
    for (int i = 0; i < ntri; i++) perm_out[i] = (int)order[(size_t)i];
}

// ---- Renderer (Renderer.cs): NewRenderer / SamplesPerPixel / StratifiedSampling / RenderParallel / IterativeRender
int pth_renderer_new(pth_world* w, int width, int height, int device) {
    try {
        w->renderer.reset(new Renderer(Renderer::NewRenderer(w->scene, w->camera, w->sampler, width, height, true)));
        w->renderer->Device = device;
        return 0;
    } catch (const std::exception& e) { w->error = e.what(); return -1; }
}
// Renderer.Devices: the GPUs one Renderer drives (multi-GPU inside the ptgpu handle); n <= 1 = the single device of pth_renderer_new.
int pth_renderer_devices(pth_world* w, int n, const int* devices) {
    if (!w->renderer) { w->error = "no renderer"; return -1; }
    w->renderer->Devices.assign(devices, devices + (n > 0 ? n : 0));
    return 0;
}
int pth_renderer_set(pth_world* w, int samplesPerPixel, int stratified, unsigned seed) {
    if (!w->renderer) { w->error = "no renderer"; return -1; }
    w->renderer->SamplesPerPixel = samplesPerPixel; w->renderer->StratifiedSampling = stratified != 0; w->renderer->Seed = seed;
    return 0;
}
int pth_renderer_set_extra(pth_world* w, int adaptiveSamples, int fireflySamples) {
    if (!w->renderer) { w->error = "no renderer"; return -1; }
    w->renderer->AdaptiveSamples = adaptiveSamples; w->renderer->FireflySamples = fireflySamples;
    return 0;
}
int pth_renderer_render(pth_world* w, float* outMeanRgb) {
    try { w->renderer->RenderParallel(outMeanRgb); return 0; } catch (const std::exception& e) { w->error = e.what(); return -1; }
}
int pth_renderer_iterative(pth_world* w, const char* pathTemplate, int iter) {
    try { w->renderer->IterativeRender(pathTemplate, iter); return 0; } catch (const std::exception& e) { w->error = e.what(); return -1; }
}
int pth_renderer_image(pth_world* w, int channel, float* out) {
    try {
        std::vector<float> img = w->renderer->Image((Channel)channel);
        std::memcpy(out, img.data(), img.size() * sizeof(float));
        return 0;
    } catch (const std::exception& e) { w->error = e.what(); return -1; }
}
int pth_renderer_counters(pth_world* w, ptgpu_counters* out) {
    if (!w->renderer) { w->error = "no renderer"; return -1; }
    *out = w->renderer->Counters();
    return 0;
}
void* pth_renderer_ctx(pth_world* w) { return w->renderer ? (void*)w->renderer->Context() : nullptr; }

}  // extern "C"
