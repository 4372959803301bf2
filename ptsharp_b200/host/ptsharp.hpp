// ptsharp.hpp — C++ host side of the B200 renderer: the reference's authoring API (Scene / Camera / DefaultSampler /
// Material / IShape factories, PTSharpCore/*.cs) restated as data-holding C++ classes, the reference's kd-tree
// builder (Tree.cs:201-265, same split axes and positions), and the flattener that turns the object graph into the
// SoA buffers of include/ptgpu.h.  No ray is ever traced on the host: every Intersect/Sample/Bounce lives in
// csrc/ptgpu.cu.  The reference's toolchain (.NET 9) is absent from this image, so this mirror is C++; the C#
// flattener + P/Invoke stub a PTSharp maintainer would add is in INTEGRATION.md.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../../include/ptgpu.h"

namespace ptsharp {

// ---- value types ---------------------------------------------------------------------------------------------
// Vector.cs:193-234: float32 storage, double-typed accessors.  Only the operations a scene author or the
// tree builder needs exist here.
struct Vector {
    float x = 0, y = 0, z = 0;
    Vector() = default;
    Vector(double X, double Y, double Z) : x((float)X), y((float)Y), z((float)Z) {}
    double X() const { return x; }
    double Y() const { return y; }
    double Z() const { return z; }
    double axis(int a) const { return a == 1 ? x : a == 2 ? y : z; }
};
double NetMin(double a, double b);  // System.Math.Min/Max (NaN-propagating, signed zeros ordered)
double NetMax(double a, double b);
Vector Add(const Vector& a, const Vector& b);
Vector Sub(const Vector& a, const Vector& b);
Vector Mul(const Vector& a, const Vector& b);
Vector Div(const Vector& a, const Vector& b);
Vector MulScalar(const Vector& a, double s);
Vector Min(const Vector& a, const Vector& b);
Vector Max(const Vector& a, const Vector& b);
Vector Cross(const Vector& a, const Vector& b);
float Dot(const Vector& a, const Vector& b);
float Length(const Vector& a);
Vector Normalize(const Vector& a);

struct Colour {  // Colour.cs:8-33
    double r = 0, g = 0, b = 0;
    Colour() = default;
    Colour(double R, double G, double B) : r(R), g(G), b(B) {}
    static Colour HexColor(int x);  // Colour.cs:125-132
    static const Colour Black, White;
};

struct Box {  // Box.cs:5-58
    Vector Min, Max;
    Box() = default;
    Box(const Vector& mn, const Vector& mx) : Min(mn), Max(mx) {}
    Box Extend(const Box& b) const;
    Vector Size() const;
    Vector Center() const;
    Vector Anchor(const Vector& anchor) const;  // Box.cs:48
    double OuterRadius() const;
};

struct Matrix {  // Matrix.cs:8-231 (row-major)
    double m[16] = {0};
    static Matrix Identity();
    // NB: in the reference these three are instance methods that ignore `this` (Matrix.cs:33-54).
    static Matrix Translate(const Vector& v);
    static Matrix Scale(const Vector& v);
    static Matrix Rotate(const Vector& axis, double angle);
    Matrix Mul(const Matrix& b) const;
    Matrix Inverse() const;
    Vector MulPosition(const Vector& b) const;
    Vector MulDirection(const Vector& b) const;  // Matrix.cs:144-150 (normalised)
    Box MulBox(const Box& box) const;
};

// ---- textures / materials ------------------------------------------------------------------------------------
struct ColorTexture {  // Texture.cs:96-100
    int Width = 0, Height = 0;
    std::vector<Colour> Data;
};
using ITexture = std::shared_ptr<ColorTexture>;

struct Material {  // Material.cs:8-100
    Colour Color;
    ITexture Texture, NormalTexture, BumpTexture, GlossTexture;
    double BumpMultiplier = 0, Emittance = 0, Index = 0, Gloss = 0, Tint = 0, Reflectivity = 0;
    bool Transparent = false;
    Material() = default;
    Material(const Colour& color, ITexture tex, ITexture normal, ITexture bump, ITexture glossTex, double b, double e,
             double i, double g, double tint, double r, bool t)
        : Color(color), Texture(tex), NormalTexture(normal), BumpTexture(bump), GlossTexture(glossTex), BumpMultiplier(b),
          Emittance(e), Index(i), Gloss(g), Tint(tint), Reflectivity(r), Transparent(t) {}
    static Material DiffuseMaterial(const Colour& c) { return Material(c, 0, 0, 0, 0, 1, 0, 1, 0, 0, -1, false); }
    static Material SpecularMaterial(const Colour& c, double index) { return Material(c, 0, 0, 0, 0, 1, 0, index, 0, 0, -1, false); }
    static Material GlossyMaterial(const Colour& c, double index, double gloss) { return Material(c, 0, 0, 0, 0, 1, 0, index, gloss, 0, -1, false); }
    static Material ClearMaterial(double index, double gloss) { return Material(Colour(0, 0, 0), 0, 0, 0, 0, 1, 0, index, gloss, 0, -1, true); }
    static Material TransparentMaterial(const Colour& c, double index, double gloss, double tint) { return Material(c, 0, 0, 0, 0, 1, 0, index, gloss, tint, -1, true); }
    static Material MetallicMaterial(const Colour& c, double gloss, double tint) { return Material(c, 0, 0, 0, 0, 1, 0, 1, gloss, tint, 1, false); }
    static Material LightMaterial(const Colour& c, double emittance) { return Material(c, 0, 0, 0, 0, 1, emittance, 1, 0, 0, -1, false); }
    bool SameAs(const Material& o) const;
};

// ---- shapes ----------------------------------------------------------------------------------------------------
struct Tree;
struct IShape {  // IShape.cs:3-11 — host keeps only what authoring, tree building and flattening need
    virtual ~IShape() {}
    virtual int Type() const = 0;
    virtual bool IsClass() const = 0;  // C# class vs struct
    virtual void Compile() {}
    virtual Box BoundingBox() const = 0;
    virtual Material MaterialAt(const Vector&) const = 0;
};
using ShapePtr = std::shared_ptr<IShape>;

struct Sphere : IShape {  // Sphere.cs
    Vector Center; double Radius; Material Mat; Box box;
    static ShapePtr NewSphere(const Vector& center, double radius, const Material& material);
    int Type() const override { return PTGPU_SPHERE; }
    bool IsClass() const override { return true; }
    Box BoundingBox() const override { return box; }
    Material MaterialAt(const Vector&) const override { return Mat; }
};
struct Cube : IShape {  // Cube.cs
    Vector Min, Max; Material Mat;
    static ShapePtr NewCube(const Vector& mn, const Vector& mx, const Material& material);
    int Type() const override { return PTGPU_CUBE; }
    bool IsClass() const override { return true; }
    Box BoundingBox() const override { return Box(Min, Max); }
    Material MaterialAt(const Vector&) const override { return Mat; }
};
struct Plane : IShape {  // Plane.cs
    Vector Point, Normal; Material Mat;
    static ShapePtr NewPlane(const Vector& point, const Vector& normal, const Material& material);
    int Type() const override { return PTGPU_PLANE; }
    bool IsClass() const override { return true; }
    Box BoundingBox() const override;
    Material MaterialAt(const Vector&) const override { return Mat; }
};
struct Cylinder : IShape {  // Cylinder.cs
    double Radius, Z0, Z1; Material Mat;
    static ShapePtr NewCylinder(double radius, double z0, double z1, const Material& material);
    static ShapePtr NewTransformedCylinder(const Vector& v0, const Vector& v1, double radius, const Material& material);
    int Type() const override { return PTGPU_CYLINDER; }
    bool IsClass() const override { return false; }
    Box BoundingBox() const override;
    Material MaterialAt(const Vector&) const override { return Mat; }
};
struct Triangle {  // Triangle.cs:8-79 (never a top-level shape here: it lives in a Mesh)
    Material Mat;
    Vector V1, V2, V3, N1, N2, N3, T1, T2, T3;
    Box BoundingBox() const;
    void FixNormals();
};
struct Mesh : IShape {  // Mesh.cs
    std::vector<Triangle> Triangles;
    std::shared_ptr<Tree> tree;
    mutable bool haveBox = false; mutable Box box;
    static std::shared_ptr<Mesh> NewMesh(std::vector<Triangle> triangles);
    int Type() const override { return PTGPU_MESH; }
    bool IsClass() const override { return false; }
    void Compile() override;
    Box BoundingBox() const override;
    Material MaterialAt(const Vector&) const override { return Material(); }  // Mesh.cs:132-135
    // authoring utilities the example scenes call on loaded meshes (host/loaders.cpp); each drops the cached box and tree
    void SmoothNormals();                              // Mesh.cs:191-229
    void SmoothNormalsThreshold(double radians);       // Mesh.cs:141-189 (incl. its prefix-list quirk)
    void MoveTo(const Vector& position, const Vector& anchor);  // Mesh.cs:237-241
    void FitInside(const Box& box, const Vector& anchor);       // Mesh.cs:243-252
    void Transform(const Matrix& matrix);              // Mesh.cs:254-274
    void SetMaterial(const Material& material);        // Mesh.cs:276-289
};
// Model loaders (host/loaders.cpp).  Both return a Mesh whose triangles went through Triangle.FixNormals.
struct OBJ { static std::shared_ptr<Mesh> Load(const std::string& path, const Material& parent); };   // OBJ.cs:11-163
struct STL { static std::shared_ptr<Mesh> Load(const std::string& path, const Material& material); }; // STL.cs:37-223
struct TransformedShape : IShape {  // TransformedShape.cs
    ShapePtr Shape; Matrix M, Inv;
    static ShapePtr NewTransformedShape(ShapePtr s, const Matrix& m);
    int Type() const override { return PTGPU_TRANSFORMED; }
    bool IsClass() const override { return false; }
    void Compile() override { Shape->Compile(); }
    Box BoundingBox() const override { return M.MulBox(Shape->BoundingBox()); }
    Material MaterialAt(const Vector& p) const override { return Shape->MaterialAt(p); }
};

// SDF.cs node types: bounding boxes on the host, evaluation as a linear program on the device.
struct SDF {
    virtual ~SDF() {}
    virtual Box BoundingBox() const = 0;
    virtual void Emit(std::vector<ptgpu_sdf_op>& prog) const = 0;
};
using SDFPtr = std::shared_ptr<SDF>;
SDFPtr NewSphereSDF(double radius);
SDFPtr NewCubeSDF(const Vector& size);
SDFPtr NewCylinderSDF(double radius, double height);
SDFPtr NewCapsuleSDF(const Vector& a, const Vector& b, double radius);
SDFPtr NewTorusSDF(double major, double minor);
SDFPtr NewTransformSDF(SDFPtr sdf, const Matrix& m);
SDFPtr NewScaleSDF(SDFPtr sdf, double factor);
SDFPtr NewRepeaterSDF(SDFPtr sdf, const Vector& step);
SDFPtr NewUnionSDF(std::vector<SDFPtr> items);
SDFPtr NewDifferenceSDF(std::vector<SDFPtr> items);
SDFPtr NewIntersectionSDF(std::vector<SDFPtr> items);

struct SDFShape : IShape {  // SDF.cs:12-110
    SDFPtr Sdf; Material Mat;
    static ShapePtr NewSDFShape(SDFPtr sdf, const Material& material);
    int Type() const override { return PTGPU_SDF; }
    bool IsClass() const override { return true; }
    Box BoundingBox() const override { return Sdf->BoundingBox(); }
    Material MaterialAt(const Vector&) const override { return Mat; }
};

// Marching cubes (MC.cs) - host/mc.cpp.  The triangles come out in the reference's order.
struct MC {
    static std::shared_ptr<Mesh> NewSDFMesh(const SDFPtr& sdf, const Box& box, double step);                                              // MC.cs:9-66
    static std::shared_ptr<Mesh> NewFieldMesh(const std::function<double(const Vector&)>& evaluate, const Box& box, double step);        // the same loop over any field
    static int CaseTriangles(int index, int out15[15]);  // triangleTable[index] (MC.cs:168-429): returns the triangle count
    static int CaseEdges(int index);                     // edgetable[index] (MC.cs:135-166)
};
struct SphericalHarmonic : IShape {  // SH.cs:7-104
    int L = 0, M = 0;
    Material PositiveMaterial, NegativeMaterial;
    std::shared_ptr<Mesh> mesh;
    static ShapePtr NewSphericalHarmonic(int l, int m, const Material& pm, const Material& nm, double step = (double)0.01f);
    int Type() const override { return PTGPU_SH; }
    bool IsClass() const override { return true; }
    void Compile() override { mesh->Compile(); }
    Box BoundingBox() const override { return Box(Vector(-1, -1, -1), Vector(1, 1, 1)); }
    Material MaterialAt(const Vector& p) const override;
    double EvaluateHarmonic(const Vector& p) const;
    double Evaluate(const Vector& p) const;
};

struct Volume : IShape {  // Volume.cs
    struct VolumeWindow { double Lo, Hi; Material VolumeWindowMaterial; };
    int W = 0, H = 0, D = 0; double ZScale = 1;
    std::vector<double> Data; std::vector<VolumeWindow> Windows; Box box;
    static ShapePtr NewVolume(const Box& box, int w, int h, int d, double zscale, std::vector<double> data, std::vector<VolumeWindow> windows);
    int Type() const override { return PTGPU_VOLUME; }
    bool IsClass() const override { return true; }
    Box BoundingBox() const override { return box; }
    Material MaterialAt(const Vector& p) const override;  // Volume.cs:147-167 (host use: light registration only)
    double Sample(double x, double y, double z) const;    // Volume.cs:73-104
};

// ---- kd-tree (builder only) -------------------------------------------------------------------------------------
struct Tree {  // Tree.cs:8-29, flattened as it is built
    Box box;
    std::vector<ptgpu_node> nodes;    // node 0 is the root
    std::vector<uint32_t> leafItems;  // indices into the shape array handed to NewTree
    uint32_t maxDepth = 0;
    static std::shared_ptr<Tree> NewTree(const std::vector<Box>& shapeBoxes);
};

// Authoring-side utility (not in the reference): a permutation of `shapeBoxes` for which the unmodified reference
// builder (Tree::NewTree above) yields a balanced tree.  See host.cpp.  order[i] = index of the shape to put at slot i.
std::vector<uint32_t> BuilderFriendlyOrder(const std::vector<Box>& shapeBoxes, double balance = 0.70, int minRepair = 16, int verbose = 0);

// ---- scene / camera / sampler -------------------------------------------------------------------------------------
struct Scene {  // Scene.cs
    Colour Color;
    ITexture Texture;
    double TextureAngle = 0;
    std::shared_ptr<Tree> tree;
    std::vector<ShapePtr> Shapes, Lights;
    void Add(ShapePtr p);
    void Compile();
};

struct Camera {  // Camera.cs
    Vector p, u, v, w;
    double m = 0, focalDistance = 0, apertureRadius = 0;
    static Camera LookAt(const Vector& eye, const Vector& center, const Vector& up, double fovy);
    void SetFocus(const Vector& focalPoint, double apertureRadius);
};

enum LightMode { LightModeRandom = 0, LightModeAll = 1 };
enum SpecularMode { SpecularModeNaive = 0, SpecularModeFirst = 1, SpecularModeAll = 2 };

struct DefaultSampler {  // Sampler.cs:10-53
    int FirstHitSamples = 1, MaxBounces = 4;
    bool DirectLighting = true, SoftShadows = true;
    bool RussianRoulette = false;  // opt-in extension (ptgpu_pass.flags): dead code in the reference (Sampler.cs:55, 133-142), off in parity mode
    ptsharp::LightMode LightMode = LightModeRandom;
    ptsharp::SpecularMode SpecularMode = SpecularModeNaive;
    static DefaultSampler NewSampler(int firstHitSamples, int maxBounces) {
        DefaultSampler s; s.FirstHitSamples = firstHitSamples; s.MaxBounces = maxBounces; return s;
    }
    void SetSpecularMode(ptsharp::SpecularMode s) { SpecularMode = s; }
    void SetLightMode(ptsharp::LightMode l) { LightMode = l; }
};

// ---- flattening -----------------------------------------------------------------------------------------------------
// Owns every array a ptgpu_flat_scene points into.
struct FlatScene {
    std::vector<ptgpu_shape> shapes;
    std::vector<uint32_t> lights;
    std::vector<ptgpu_tree> trees;
    std::vector<ptgpu_node> nodes;
    std::vector<uint32_t> leafItems;
    std::vector<ptgpu_sphere> spheres;
    std::vector<ptgpu_cube> cubes;
    std::vector<ptgpu_plane> planes;
    std::vector<ptgpu_cylinder> cylinders;
    std::vector<ptgpu_mesh> meshes;
    std::vector<ptgpu_tri_geom> triGeom;
    std::vector<ptgpu_tri_shade> triShade;
    std::vector<ptgpu_instance> instances;
    std::vector<ptgpu_sdf_shape> sdfShapes;
    std::vector<ptgpu_sdf_op> sdfOps;
    std::vector<ptgpu_volume> volumes;
    std::vector<ptgpu_volume_window> volumeWindows;
    std::vector<ptgpu_sh> shs;
    std::vector<double> volumeData;
    std::vector<ptgpu_material> materials;
    std::vector<ptgpu_texture> textures;
    std::vector<double> texels;
    ptgpu_flat_scene view{};
    uint64_t Bytes() const;
    void Bind(uint32_t sceneTree, uint32_t numSceneShapes);  // point `view` at the vectors (keeps the env fields)
};
// Flat-scene file: what a .NET box running the real PTSharp would dump for this GPU-only harness, and the reverse.
void SaveFlatScene(const FlatScene& scene, const std::string& path);
std::unique_ptr<FlatScene> LoadFlatScene(const std::string& path);
// Scene must be Compile()d.  Throws std::runtime_error on unsupported graphs (nested TransformedShape, ...).
std::unique_ptr<FlatScene> Flatten(const Scene& scene);
ptgpu_camera FlattenCamera(const Camera& c);

// ---- render driver ---------------------------------------------------------------------------------------------------
// Buffer.cs channels, read back from the device-resident buffer.
enum Channel { ColorChannel = 0, VarianceChannel = 1, StandardDeviationChannel = 2, SamplesChannel = 3, AlbedoChannel = 4, NormalChannel = 5 };

class Renderer {  // Renderer.cs:15-56, 199-338, 702-765
public:
    int SamplesPerPixel = 2;
    bool StratifiedSampling = false;
    int AdaptiveSamples = 0, FireflySamples = 0;  // Renderer.cs:340-468, run on the device after the main pass
    double FireflyThreshold = 1;                  // Renderer.cs:47
    double AdaptiveThreshold = 1, AdaptiveExponent = 1;  // Renderer.cs:44-45 (read by the serial Render() only, :153-158)
    int NumCPU = 0;                               // Renderer.cs:38: 1 when NewRenderer(..., multithreaded = false) -> IterativeRender runs the serial Render()
    int Device = 0;
    std::vector<int> Devices;                     // more than one entry: the pass is split over these GPUs inside the handle (ptgpu_params.devices)
    uint32_t Seed = 0x50545348u;
    static Renderer NewRenderer(Scene& scene, Camera& camera, DefaultSampler& sampler, int w, int h, bool multithreaded);
    ~Renderer();
    Renderer(Renderer&&) noexcept;
    Renderer(const Renderer&) = delete;
    // One pass (Renderer.RenderParallel): returns this pass's mean image (w*h*3) if out != nullptr.
    void RenderParallel(float* outMeanRgb = nullptr);
    // One pass with the serial Render()'s rules for the extra samples (Renderer.cs:80-198); the main pass is the same.
    void Render(float* outMeanRgb = nullptr);
    // `iter` passes; after each the Color channel is written as binary PPM to pathTemplate ("{0}" -> pass number).
    // (The reference writes PNG through SkiaSharp, Renderer.cs:721-729; image encoding is out of scope.)
    void IterativeRender(const std::string& pathTemplate, int iter);
    std::vector<float> Image(Channel channel);
    ptgpu_counters Counters();
    ptgpu_pass MakePass() const;
    ptgpu_ctx* Context() { return ctx_; }
    int Width() const { return w_; }
    int Height() const { return h_; }
private:
    Renderer() = default;
    void EnsureUploaded();
    Scene* scene_ = nullptr; Camera* camera_ = nullptr; DefaultSampler* sampler_ = nullptr;
    int w_ = 0, h_ = 0;
    uint32_t passIndex_ = 0;
    ptgpu_ctx* ctx_ = nullptr;
    std::unique_ptr<FlatScene> flat_;
};

}  // namespace ptsharp
