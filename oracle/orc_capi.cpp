// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity unpinned (see orc_math.hpp header and DESIGN.md).
// C entry points so tests/ and bench.py's cpu_baseline leg can drive the CPU restatement through ctypes.
// The authoring verbs (orc_sphere, orc_mesh, orc_transformed, ...) take the same arguments as the reference's
// factory methods (Sphere.NewSphere, Mesh.NewMesh, TransformedShape.NewTransformedShape, ...).
#include <cstdlib>
#include <memory>
#include <vector>

#include "orc_render.hpp"

using namespace orc;

struct orc_world {
    Scene scene;
    Camera camera;
    DefaultSampler sampler;
    std::vector<std::unique_ptr<ColorTexture>> textures;
    std::vector<Material> materials;
    std::vector<std::unique_ptr<IShape>> shapes;
    std::vector<std::unique_ptr<SDF>> sdfs;
    std::unique_ptr<Buffer> buffer;
    bool compiled = false;
    int adaptiveSamples = 0, fireflySamples = 0;
    double fireflyThreshold = 1;
    int serialRules = 0;
    int taskStride = 1, taskOffset = 0;  // orc_set_task_sample
    double adaptiveThreshold = 1, adaptiveExponent = 1;
    std::vector<int> lastSamples;  // Pixel.Samples of the last orc_render
};

static Vector V3(const double* v) { return Vector(v[0], v[1], v[2]); }

extern "C" {

orc_world* orc_world_new() { return new orc_world(); }
void orc_world_free(orc_world* w) { delete w; }

int orc_texture(orc_world* w, int width, int height, const double* rgb) {
    auto t = std::make_unique<ColorTexture>();
    t->Width = width;
    t->Height = height;
    t->Data.resize((size_t)width * height);
    for (size_t i = 0; i < t->Data.size(); i++) t->Data[i] = Colour(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
    w->textures.push_back(std::move(t));
    return (int)w->textures.size() - 1;
}

// Material.cs:48-62 constructor argument order (color, textures, b, e, i, g, tint, r, t).
int orc_material(orc_world* w, const double* color, int tex, int normalTex, int bumpTex, int glossTex, double bumpMultiplier,
                 double emittance, double index, double gloss, double tint, double reflectivity, int transparent) {
    Material m;
    m.Color = Colour(color[0], color[1], color[2]);
    auto T = [&](int id) -> const ColorTexture* { return id >= 0 ? w->textures[(size_t)id].get() : nullptr; };
    m.Texture = T(tex);
    m.NormalTexture = T(normalTex);
    m.BumpTexture = T(bumpTex);
    m.GlossTexture = T(glossTex);
    m.BumpMultiplier = bumpMultiplier;
    m.Emittance = emittance;
    m.Index = index;
    m.Gloss = gloss;
    m.Tint = tint;
    m.Reflectivity = reflectivity;
    m.Transparent = transparent != 0;
    m.id = (int)w->materials.size();
    w->materials.push_back(m);
    return m.id;
}

static int push_shape(orc_world* w, IShape* s) {
    w->shapes.emplace_back(s);
    return (int)w->shapes.size() - 1;
}

int orc_sphere(orc_world* w, const double* c, double r, int mat) { return push_shape(w, new Sphere(V3(c), r, w->materials[(size_t)mat])); }
int orc_cube(orc_world* w, const double* mn, const double* mx, int mat) { return push_shape(w, new Cube(V3(mn), V3(mx), w->materials[(size_t)mat])); }
int orc_plane(orc_world* w, const double* p, const double* n, int mat) { return push_shape(w, new Plane(V3(p), V3(n), w->materials[(size_t)mat])); }
int orc_cylinder(orc_world* w, double r, double z0, double z1, int mat) { return push_shape(w, new Cylinder(r, z0, z1, w->materials[(size_t)mat])); }

// V/N/T: ntri*9 floats (three xyz triples per triangle).  N == NULL -> FixNormals() (Triangle.cs:224-237);
// mats == NULL -> every triangle gets `mat`.
int orc_mesh(orc_world* w, int ntri, const float* V, const float* N, const float* T, const int* mats, int mat) {
    Mesh* m = new Mesh();
    m->Triangles.resize((size_t)ntri);
    for (int i = 0; i < ntri; i++) {
        Triangle& t = m->Triangles[(size_t)i];
        const float* v = V + (size_t)i * 9;
        t.V1 = Vector(v[0], v[1], v[2]); t.V2 = Vector(v[3], v[4], v[5]); t.V3 = Vector(v[6], v[7], v[8]);
        if (N) {
            const float* n = N + (size_t)i * 9;
            t.N1 = Vector(n[0], n[1], n[2]); t.N2 = Vector(n[3], n[4], n[5]); t.N3 = Vector(n[6], n[7], n[8]);
        }
        if (T) {
            const float* q = T + (size_t)i * 9;
            t.T1 = Vector(q[0], q[1], q[2]); t.T2 = Vector(q[3], q[4], q[5]); t.T3 = Vector(q[6], q[7], q[8]);
        }
        t.Mat = w->materials[(size_t)(mats ? mats[i] : mat)];
        t.index = i;
        t.FixNormals();
    }
    return push_shape(w, m);
}

// SphericalHarmonic.NewSphericalHarmonic(l, m, pm, nm) with its marching-cubes mesh supplied by the caller (ntri x 9 floats).
int orc_spherical_harmonic(orc_world* w, int l, int m, int pm, int nm, int ntri, const float* V) {
    SphericalHarmonic* sh = new SphericalHarmonic();
    sh->L = l; sh->M = m; sh->PositiveMaterial = w->materials[(size_t)pm]; sh->NegativeMaterial = w->materials[(size_t)nm];
    sh->mesh.Triangles.resize((size_t)ntri);
    for (int i = 0; i < ntri; i++) {
        Triangle& t = sh->mesh.Triangles[(size_t)i];
        const float* v = V + (size_t)i * 9;
        t.V1 = Vector(v[0], v[1], v[2]); t.V2 = Vector(v[3], v[4], v[5]); t.V3 = Vector(v[6], v[7], v[8]);
        t.index = i;
        t.FixNormals();  // MC.cs:107
    }
    return push_shape(w, sh);
}

int orc_transformed(orc_world* w, int shape, const double* m16) {
    return push_shape(w, new TransformedShape(w->shapes[(size_t)shape].get(), Matrix::FromRows(m16)));
}

static int push_sdf(orc_world* w, SDF* s) {
    w->sdfs.emplace_back(s);
    return (int)w->sdfs.size() - 1;
}
int orc_sdf_sphere(orc_world* w, double r) { return push_sdf(w, new SphereSDF(r)); }
int orc_sdf_cube(orc_world* w, const double* size) { return push_sdf(w, new CubeSDF(V3(size))); }
int orc_sdf_cylinder(orc_world* w, double r, double h) { return push_sdf(w, new CylinderSDF(r, h)); }
int orc_sdf_capsule(orc_world* w, const double* a, const double* b, double r) { return push_sdf(w, new CapsuleSDF(V3(a), V3(b), r)); }
int orc_sdf_torus(orc_world* w, double major, double minor) { return push_sdf(w, new TorusSDF(major, minor)); }
int orc_sdf_transform(orc_world* w, int sdf, const double* m16) { return push_sdf(w, new TransformSDF(w->sdfs[(size_t)sdf].get(), Matrix::FromRows(m16))); }
int orc_sdf_scale(orc_world* w, int sdf, double f) { return push_sdf(w, new ScaleSDF(w->sdfs[(size_t)sdf].get(), f)); }
int orc_sdf_repeat(orc_world* w, int sdf, const double* step) { return push_sdf(w, new RepeatSDF(w->sdfs[(size_t)sdf].get(), V3(step))); }
// op: 0 union, 1 difference, 2 intersection
int orc_sdf_combine(orc_world* w, int op, int n, const int* items) {
    std::vector<const SDF*> v;
    for (int i = 0; i < n; i++) v.push_back(w->sdfs[(size_t)items[i]].get());
    if (op == 0) { auto* s = new UnionSDF(); s->Items = v; return push_sdf(w, s); }
    if (op == 1) { auto* s = new DifferenceSDF(); s->Items = v; return push_sdf(w, s); }
    auto* s = new IntersectionSDF(); s->Items = v; return push_sdf(w, s);
}
int orc_sdf_shape(orc_world* w, int sdf, int mat) { return push_shape(w, new SDFShape(w->sdfs[(size_t)sdf].get(), w->materials[(size_t)mat])); }

int orc_volume(orc_world* w, const double* bmin, const double* bmax, int W, int H, int D, double zscale, const double* data,
               int nwin, const double* lo, const double* hi, const int* mats) {
    Volume* v = new Volume();
    v->W = W; v->H = H; v->D = D; v->ZScale = zscale;
    v->Data.assign(data, data + (size_t)W * H * D);
    for (int i = 0; i < nwin; i++) v->Windows.push_back(VolumeWindow{lo[i], hi[i], w->materials[(size_t)mats[i]]});
    v->box = Box(V3(bmin), V3(bmax));
    return push_shape(w, v);
}

void orc_scene_add(orc_world* w, int shape) { w->scene.Add(w->shapes[(size_t)shape].get()); }
void orc_scene_env(orc_world* w, const double* color, int tex, double angle) {
    w->scene.Color = Colour(color[0], color[1], color[2]);
    w->scene.Texture = tex >= 0 ? w->textures[(size_t)tex].get() : nullptr;
    w->scene.TextureAngle = angle;
}
void orc_camera_lookat(orc_world* w, const double* eye, const double* center, const double* up, double fovy) {
    w->camera = Camera::LookAt(V3(eye), V3(center), V3(up), fovy);
}
void orc_camera_focus(orc_world* w, const double* focalPoint, double aperture) { w->camera.SetFocus(V3(focalPoint), aperture); }
void orc_sampler(orc_world* w, int firstHit, int maxBounces, int directLighting, int softShadows, int lightMode, int specularMode) {
    w->sampler.FirstHitSamples = firstHit;
    w->sampler.MaxBounces = maxBounces;
    w->sampler.DirectLighting = directLighting != 0;
    w->sampler.SoftShadows = softShadows != 0;
    w->sampler.lightMode = lightMode;
    w->sampler.specularMode = specularMode;
}

// Renderer.AdaptiveSamples / FireflySamples / FireflyThreshold for subsequent orc_render calls.
void orc_set_extra(orc_world* w, int adaptiveSamples, int fireflySamples, double fireflyThreshold) {
    w->adaptiveSamples = adaptiveSamples; w->fireflySamples = fireflySamples; w->fireflyThreshold = fireflyThreshold;
}
// serial != 0: the extra samples follow the serial Render() (Renderer.cs:150-191) with Renderer.AdaptiveThreshold / AdaptiveExponent.
void orc_set_serial(orc_world* w, int serial, double adaptiveThreshold, double adaptiveExponent) {
    w->serialRules = serial; w->adaptiveThreshold = adaptiveThreshold; w->adaptiveExponent = adaptiveExponent;
}

int orc_last_samples(orc_world* w, int n, int* out) {
    if ((size_t)n != w->lastSamples.size()) return -1;
    std::memcpy(out, w->lastSamples.data(), (size_t)n * sizeof(int));
    return 0;
}

void orc_compile(orc_world* w) {
    w->scene.Compile();
    w->compiled = true;
}
int orc_num_lights(orc_world* w) { return (int)w->scene.Lights.size(); }
void orc_camera_get(orc_world* w, float* puvw12, double* mfa3) {
    const Camera& c = w->camera;
    const Vector* vs[4] = {&c.p, &c.u, &c.v, &c.w};
    for (int i = 0; i < 4; i++) { puvw12[3 * i] = vs[i]->x; puvw12[3 * i + 1] = vs[i]->y; puvw12[3 * i + 2] = vs[i]->z; }
    mfa3[0] = c.m; mfa3[1] = c.focalDistance; mfa3[2] = c.apertureRadius;
}

// Scene.Intersect + Hit.Info on caller-supplied rays.  shape = index in Scene.Shapes of the top-level shape whose
// Intersect produced the hit (-1 = miss); prim = triangle index inside its mesh (-1 if not a triangle).
void orc_intersect_batch(orc_world* w, int n, const float* o, const float* d, int* shape, int* prim, double* t,
                         float* normal, float* position, int* inside, int* material) {
    if (!w->compiled) orc_compile(w);
    for (int i = 0; i < n; i++) {
        Ray r(Vector(o[3 * i], o[3 * i + 1], o[3 * i + 2]), Vector(d[3 * i], d[3 * i + 1], d[3 * i + 2]));
        Hit hit = w->scene.Intersect(r);
        if (!hit.Ok()) {
            shape[i] = -1; prim[i] = -1; t[i] = hit.T;
            if (normal) normal[3 * i] = normal[3 * i + 1] = normal[3 * i + 2] = 0;
            if (position) position[3 * i] = position[3 * i + 1] = position[3 * i + 2] = 0;
            if (inside) inside[i] = 0;
            if (material) material[i] = -1;
            continue;
        }
        int owner = hit.top;
        shape[i] = owner;
        prim[i] = hit.prim;
        t[i] = hit.T;
        HitInfo info = hit.Info(r);
        if (normal) { normal[3 * i] = info.Normal.x; normal[3 * i + 1] = info.Normal.y; normal[3 * i + 2] = info.Normal.z; }
        if (position) { position[3 * i] = info.Position.x; position[3 * i + 1] = info.Position.y; position[3 * i + 2] = info.Position.z; }
        if (inside) inside[i] = info.Inside ? 1 : 0;
        if (material) material[i] = info.material.id;
    }
}

// Camera.CastRay on caller-supplied (x, y, u, v); lens draws (if any) come from the keyed stream of (pixel, sample).
void orc_cast_rays(orc_world* w, int W, int H, int n, const int* x, const int* y, const double* fu, const double* fv,
                   const int* sample, unsigned seed, unsigned pass, float* o, float* d) {
    Rng rng;
    rng.mode = RNG_KEYED;
    for (int i = 0; i < n; i++) {
        rng.SetSample(seed, pass, (uint32_t)(y[i] * W + x[i]), (uint32_t)sample[i]);
        rng.Enter(0, 0, 0, 0);
        rng.NextDouble(); rng.NextDouble();  // the two jitter draws precede the lens draws (Renderer.cs:297-298)
        Ray r = w->camera.CastRay(x[i], y[i], W, H, fu[i], fv[i], rng);
        o[3 * i] = r.Origin.x; o[3 * i + 1] = r.Origin.y; o[3 * i + 2] = r.Origin.z;
        d[3 * i] = r.Direction.x; d[3 * i + 1] = r.Direction.y; d[3 * i + 2] = r.Direction.z;
    }
}

// Bounded samples for timing: render only every stride-th non-empty 32x32 task of the frame (stride <= 1: all of them).
void orc_set_task_sample(orc_world* w, int stride, int offset) { w->taskStride = stride; w->taskOffset = stride > 1 ? ((offset % stride) + stride) % stride : 0; }

// `passes` calls of RenderParallel on a fresh Buffer; mean = Pixel.M, var = Pixel.Variance() (Buffer.cs:46-55).
// window = {x0,y0,x1,y1} or NULL.  counters = {cameraSamples, segments, shadowRays}.
void orc_render(orc_world* w, int W, int H, int spp, int passes, int stratified, int threads, int rngMode, unsigned seed,
                int sampleBase, int sampleStride, const int* window, double* mean, double* var, long long* counters) {
    if (!w->compiled) orc_compile(w);
    Buffer buf(W, H);
    Counters total;
    for (int p = 0; p < passes; p++) {
        RenderOptions opt;
        opt.SamplesPerPixel = spp;
        opt.StratifiedSampling = stratified != 0;
        opt.threads = threads;
        opt.rngMode = rngMode;
        opt.seed = seed;
        opt.pass = (uint32_t)p;
        opt.sampleBase = sampleBase;
        opt.sampleStride = sampleStride > 0 ? sampleStride : 1;
        opt.AdaptiveSamples = w->adaptiveSamples; opt.FireflySamples = w->fireflySamples; opt.FireflyThreshold = w->fireflyThreshold;
        opt.SerialRules = w->serialRules != 0; opt.AdaptiveThreshold = w->adaptiveThreshold; opt.AdaptiveExponent = w->adaptiveExponent;
        if (window) { opt.x0 = window[0]; opt.y0 = window[1]; opt.x1 = window[2]; opt.y1 = window[3]; }
        opt.taskStride = w->taskStride > 1 ? w->taskStride : 1; opt.taskOffset = w->taskOffset;
        Counters c = RenderPass(w->scene, w->camera, w->sampler, buf, opt);
        total.cameraSamples += c.cameraSamples;
        total.segments += c.segments;
        total.shadowRays += c.shadowRays;
    }
    w->lastSamples.resize(buf.Pixels.size());
    for (size_t i = 0; i < buf.Pixels.size(); i++) {
        const Pixel& px = buf.Pixels[i];
        w->lastSamples[i] = px.Samples;
        if (mean) { mean[3 * i] = px.M.r; mean[3 * i + 1] = px.M.g; mean[3 * i + 2] = px.M.b; }
        if (var) { Colour v = px.Variance(); var[3 * i] = v.r; var[3 * i + 1] = v.g; var[3 * i + 2] = v.b; }
    }
    if (counters) { counters[0] = total.cameraSamples; counters[1] = total.segments; counters[2] = total.shadowRays; }
}

// ---- kd-tree dump (pre-order) for builder parity.  which = -1: scene tree; otherwise the tree of mesh shape `which`.
static const Tree* pick_tree(orc_world* w, int which) {
    if (!w->compiled) orc_compile(w);
    if (which < 0) return w->scene.tree.get();
    IShape* s = w->shapes[(size_t)which].get();
    if (s->Kind() != K_MESH) return nullptr;
    Mesh* m = static_cast<Mesh*>(s);
    m->Compile();
    return m->tree.get();
}
static void count_nodes(const Node* n, long long& nodes, long long& items, long long& maxLeaf, int depth, int& maxDepth) {
    nodes++;
    if (depth > maxDepth) maxDepth = depth;
    if (n->Axis == 0) {
        items += (long long)n->Shapes.size();
        if ((long long)n->Shapes.size() > maxLeaf) maxLeaf = (long long)n->Shapes.size();
        return;
    }
    count_nodes(n->Left.get(), nodes, items, maxLeaf, depth + 1, maxDepth);
    count_nodes(n->Right.get(), nodes, items, maxLeaf, depth + 1, maxDepth);
}
// out4 = {nodes, leafItems, maxLeafSize, maxDepth}; box6 = tree box min/max
int orc_tree_stats(orc_world* w, int which, long long* out4, float* box6) {
    const Tree* t = pick_tree(w, which);
    if (!t) return -1;
    long long nodes = 0, items = 0, maxLeaf = 0;
    int maxDepth = 0;
    count_nodes(t->Root.get(), nodes, items, maxLeaf, 0, maxDepth);
    out4[0] = nodes; out4[1] = items; out4[2] = maxLeaf; out4[3] = maxDepth;
    if (box6) {
        box6[0] = t->box.Min.x; box6[1] = t->box.Min.y; box6[2] = t->box.Min.z;
        box6[3] = t->box.Max.x; box6[4] = t->box.Max.y; box6[5] = t->box.Max.z;
    }
    return 0;
}
struct DumpCtx { int* axis; double* point; int* a; int* b; int* items; long long nn = 0, ni = 0; };
static long long dump_node(const Node* n, DumpCtx& c) {
    long long me = c.nn++;
    c.axis[me] = n->Axis;
    c.point[me] = n->Point;
    if (n->Axis == 0) {
        c.a[me] = (int)c.ni;
        c.b[me] = (int)n->Shapes.size();
        for (const IShape* s : n->Shapes) {
            int id = s->Kind() == K_TRIANGLE ? static_cast<const Triangle*>(s)->index : s->sceneIndex;
            c.items[c.ni++] = id;
        }
        return me;
    }
    c.a[me] = (int)dump_node(n->Left.get(), c);
    c.b[me] = (int)dump_node(n->Right.get(), c);
    return me;
}
// Pre-order arrays: axis[n], point[n], a[n] (left child | first item), b[n] (right child | item count), items[leafItems].
int orc_tree_dump(orc_world* w, int which, int* axis, double* point, int* a, int* b, int* items) {
    const Tree* t = pick_tree(w, which);
    if (!t) return -1;
    DumpCtx c{axis, point, a, b, items};
    dump_node(t->Root.get(), c);
    return 0;
}

// Average kd-tree work per ray (nodes visited, leaf shapes tested) over caller-supplied rays; single-threaded.
void orc_traversal_cost(orc_world* w, int n, const float* o, const float* d, double* out2) {
    if (!w->compiled) orc_compile(w);
    probe().on = true; probe().nodes = 0; probe().prims = 0;
    for (int i = 0; i < n; i++) {
        Ray r(Vector(o[3 * i], o[3 * i + 1], o[3 * i + 2]), Vector(d[3 * i], d[3 * i + 1], d[3 * i + 2]));
        w->scene.tree->Intersect(r);
    }
    probe().on = false;
    out2[0] = (double)probe().nodes / (n > 0 ? n : 1);
    out2[1] = (double)probe().prims / (n > 0 ? n : 1);
}

// Known-answer probes for the stateless pieces (tests/test_oracle_kat.py).
void orc_philox(const unsigned* ctr4, const unsigned* key2, unsigned* out4) { philox4x32_10(ctr4, key2, out4); }
double orc_reflectance(const double* n, const double* i, double n1, double n2) { return V3(n).Reflectance(V3(i), n1, n2); }
void orc_refract(const double* n, const double* i, double n1, double n2, float* out) {
    Vector r = V3(n).Refract(V3(i), n1, n2);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void orc_matrix_inverse(const double* m16, double* out16) {
    Matrix r = Matrix::FromRows(m16).Inverse();
    std::memcpy(out16, r.m, sizeof(r.m));
}
void orc_matrix_rotate(const double* axis, double angle, double* out16) {
    Matrix r = Matrix::Rotate(V3(axis), angle);
    std::memcpy(out16, r.m, sizeof(r.m));
}
void orc_hexcolor(int x, double* out3) {
    Colour c = Colour::HexColor(x);
    out3[0] = c.r; out3[1] = c.g; out3[2] = c.b;
}
// Welford accumulator (Buffer.cs:33-55): feed n samples, return mean and variance.
void orc_welford(int n, const double* samples3, double* mean3, double* var3) {
    Pixel p;
    for (int i = 0; i < n; i++) p.AddSample(Colour(samples3[3 * i], samples3[3 * i + 1], samples3[3 * i + 2]));
    mean3[0] = p.M.r; mean3[1] = p.M.g; mean3[2] = p.M.b;
    Colour v = p.Variance();
    var3[0] = v.r; var3[1] = v.g; var3[2] = v.b;
}
// A keyed-stream draw, for cross-checking the GPU's Philox addressing.
double orc_keyed_draw(unsigned seed, unsigned pass, unsigned pixel, unsigned sample, unsigned bits, unsigned first,
                      unsigned depth, unsigned sub, unsigned drawIndex) {
    Rng rng;
    rng.mode = RNG_KEYED;
    rng.SetSample(seed, pass, pixel, sample);
    rng.Enter(bits, first, depth, sub);
    double v = 0;
    for (unsigned i = 0; i <= drawIndex; i++) v = rng.NextDouble();
    return v;
}

}  // extern "C"
