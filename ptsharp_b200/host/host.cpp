// host.cpp — value types, shape factories, the reference's kd-tree builder and the scene flattener.
// See ptsharp.hpp.  Compiled with -ffp-contract=off: the .NET JIT never fuses a*b+c, and the builder's split
// positions / bounding boxes must come out bit-identical to what the C# host would hand over.
#include "ptsharp.hpp"

#include <algorithm>
#include <cstdio>
#include <map>
#include <stdexcept>

namespace ptsharp {

// ---------------------------------------------------------------------------------------------------- values
double NetMax(double a, double b) {
    if (a != b) return std::isnan(a) ? a : (b < a ? a : b);
    return std::signbit(b) ? a : b;
}
double NetMin(double a, double b) {
    if (a != b) return std::isnan(a) ? a : (a < b ? a : b);
    return std::signbit(a) ? a : b;
}
// Vector.cs:408-444 — double arithmetic on widened floats, rounded once by the Vector constructor.
Vector Add(const Vector& a, const Vector& b) { return Vector(a.X() + b.X(), a.Y() + b.Y(), a.Z() + b.Z()); }
Vector Sub(const Vector& a, const Vector& b) { return Vector(a.X() - b.X(), a.Y() - b.Y(), a.Z() - b.Z()); }
Vector Mul(const Vector& a, const Vector& b) { return Vector(a.X() * b.X(), a.Y() * b.Y(), a.Z() * b.Z()); }
Vector Div(const Vector& a, const Vector& b) { return Vector(a.X() / b.X(), a.Y() / b.Y(), a.Z() / b.Z()); }
Vector MulScalar(const Vector& a, double s) { return Vector(a.X() * s, a.Y() * s, a.Z() * s); }
Vector Min(const Vector& a, const Vector& b) { return Vector(NetMin(a.X(), b.X()), NetMin(a.Y(), b.Y()), NetMin(a.Z(), b.Z())); }
Vector Max(const Vector& a, const Vector& b) { return Vector(NetMax(a.X(), b.X()), NetMax(a.Y(), b.Y()), NetMax(a.Z(), b.Z())); }
// System.Numerics.Vector3 float ops (Vector.cs:356-393).
float Dot(const Vector& a, const Vector& b) {
    float px = a.x * b.x, py = a.y * b.y, pz = a.z * b.z;
    float s = px + py;
    return s + pz;
}
Vector Cross(const Vector& a, const Vector& b) {
    Vector r;
    float yz = a.y * b.z, zy = a.z * b.y, zx = a.z * b.x, xz = a.x * b.z, xy = a.x * b.y, yx = a.y * b.x;
    r.x = yz - zy; r.y = zx - xz; r.z = xy - yx;
    return r;
}
float Length(const Vector& a) { return std::sqrt(Dot(a, a)); }
Vector Normalize(const Vector& a) {
    float len = Length(a);
    Vector r; r.x = a.x / len; r.y = a.y / len; r.z = a.z / len;
    return r;
}

const Colour Colour::Black(0, 0, 0);
const Colour Colour::White(1, 1, 1);
Colour Colour::HexColor(int x) {
    float red = (float)((x >> 16) & 0xff) / 255.0f, green = (float)((x >> 8) & 0xff) / 255.0f, blue = (float)(x & 0xff) / 255.0f;
    double e = (double)2.2f;
    return Colour(std::pow((double)red, e), std::pow((double)green, e), std::pow((double)blue, e));
}

Box Box::Extend(const Box& b) const { return Box(ptsharp::Min(Min, b.Min), ptsharp::Max(Max, b.Max)); }
Vector Box::Size() const { return Sub(Max, Min); }
Vector Box::Center() const { return Add(Min, Mul(Size(), Vector(0.5, 0.5, 0.5))); }  // Box.cs:46-48
double Box::OuterRadius() const { return Length(Sub(Min, Center())); }               // Box.cs:50

Matrix Matrix::Identity() { Matrix r; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1; return r; }
Matrix Matrix::Translate(const Vector& v) { Matrix r = Identity(); r.m[3] = v.X(); r.m[7] = v.Y(); r.m[11] = v.Z(); return r; }
Matrix Matrix::Scale(const Vector& v) { Matrix r = Identity(); r.m[0] = v.X(); r.m[5] = v.Y(); r.m[10] = v.Z(); return r; }
Matrix Matrix::Rotate(const Vector& axis, double a) {  // Matrix.cs:44-54
    Vector v = Normalize(axis);
    double s = std::sin(a), c = std::cos(a), k = 1 - c, x = v.X(), y = v.Y(), z = v.Z();
    Matrix r;
    double vals[16] = {k * x * x + c,     k * x * y + z * s, k * z * x - y * s, 0,
                       k * x * y - z * s, k * y * y + c,     k * y * z + x * s, 0,
                       k * z * x + y * s, k * y * z - x * s, k * z * z + c,     0,
                       0, 0, 0, 1};
    std::memcpy(r.m, vals, sizeof(vals));
    return r;
}
Matrix Matrix::Mul(const Matrix& b) const {  // Matrix.cs:111-131
    Matrix r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            r.m[i * 4 + j] = m[i * 4 + 0] * b.m[0 * 4 + j] + m[i * 4 + 1] * b.m[1 * 4 + j] + m[i * 4 + 2] * b.m[2 * 4 + j] + m[i * 4 + 3] * b.m[3 * 4 + j];
    return r;
}
Vector Matrix::MulPosition(const Vector& b) const {  // Matrix.cs:134-141
    return Vector(m[0] * b.X() + m[1] * b.Y() + m[2] * b.Z() + m[3], m[4] * b.X() + m[5] * b.Y() + m[6] * b.Z() + m[7],
                  m[8] * b.X() + m[9] * b.Y() + m[10] * b.Z() + m[11]);
}
Box Matrix::MulBox(const Box& box) const {  // Matrix.cs:157-173
    Vector r(m[0], m[4], m[8]), u(m[1], m[5], m[9]), b(m[2], m[6], m[10]), t(m[3], m[7], m[11]);
    Vector xa = MulScalar(r, box.Min.X()), xb = MulScalar(r, box.Max.X());
    Vector ya = MulScalar(u, box.Min.Y()), yb = MulScalar(u, box.Max.Y());
    Vector za = MulScalar(b, box.Min.Z()), zb = MulScalar(b, box.Max.Z());
    Vector xlo = ptsharp::Min(xa, xb), xhi = ptsharp::Max(xa, xb);
    Vector ylo = ptsharp::Min(ya, yb), yhi = ptsharp::Max(ya, yb);
    Vector zlo = ptsharp::Min(za, zb), zhi = ptsharp::Max(za, zb);
    return Box(Add(Add(Add(xlo, ylo), zlo), t), Add(Add(Add(xhi, yhi), zhi), t));
}
// Matrix.cs:179-217: cofactor expansion, terms in the reference's order (the sums are not reassociated).
Matrix Matrix::Inverse() const {
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7];
    const double i = m[8], j = m[9], k = m[10], l = m[11], mm = m[12], n = m[13], o = m[14], p = m[15];
    double det = (a * f * k * p - a * f * l * o + a * g * l * n - a * g * j * p + a * h * j * o - a * h * k * n -
                  b * g * l * mm + b * g * i * p - b * h * i * o + b * h * k * mm - b * e * k * p + b * e * l * o +
                  c * h * i * n - c * h * j * mm + c * e * j * p - c * e * l * n + c * f * l * mm - c * f * i * p -
                  d * e * j * o + d * e * k * n - d * f * k * mm + d * f * i * o - d * g * i * n + d * g * j * mm);
    Matrix r;
    r.m[0] = (g * l * n - h * k * n + h * j * o - f * l * o - g * j * p + f * k * p) / det;
    r.m[1] = (d * k * n - c * l * n - d * j * o + b * l * o + c * j * p - b * k * p) / det;
    r.m[2] = (c * h * n - d * g * n + d * f * o - b * h * o - c * f * p + b * g * p) / det;
    r.m[3] = (d * g * j - c * h * j - d * f * k + b * h * k + c * f * l - b * g * l) / det;
    r.m[4] = (h * k * mm - g * l * mm - h * i * o + e * l * o + g * i * p - e * k * p) / det;
    r.m[5] = (c * l * mm - d * k * mm + d * i * o - a * l * o - c * i * p + a * k * p) / det;
    r.m[6] = (d * g * mm - c * h * mm - d * e * o + a * h * o + c * e * p - a * g * p) / det;
    r.m[7] = (c * h * i - d * g * i + d * e * k - a * h * k - c * e * l + a * g * l) / det;
    r.m[8] = (f * l * mm - h * j * mm + h * i * n - e * l * n - f * i * p + e * j * p) / det;
    r.m[9] = (d * j * mm - b * l * mm - d * i * n + a * l * n + b * i * p - a * j * p) / det;
    r.m[10] = (b * h * mm - d * f * mm + d * e * n - a * h * n - b * e * p + a * f * p) / det;
    r.m[11] = (d * f * i - b * h * i - d * e * j + a * h * j + b * e * l - a * f * l) / det;
    r.m[12] = (g * j * mm - f * k * mm - g * i * n + e * k * n + f * i * o - e * j * o) / det;
    r.m[13] = (b * k * mm - c * j * mm + c * i * n - a * k * n - b * i * o + a * j * o) / det;
    r.m[14] = (c * f * mm - b * g * mm - c * e * n + a * g * n + b * e * o - a * f * o) / det;
    r.m[15] = (b * g * i - c * f * i + c * e * j - a * g * j - b * e * k + a * f * k) / det;
    return r;
}

bool Material::SameAs(const Material& o) const {
    return Color.r == o.Color.r && Color.g == o.Color.g && Color.b == o.Color.b && Texture == o.Texture &&
           NormalTexture == o.NormalTexture && BumpTexture == o.BumpTexture && GlossTexture == o.GlossTexture &&
           BumpMultiplier == o.BumpMultiplier && Emittance == o.Emittance && Index == o.Index && Gloss == o.Gloss &&
           Tint == o.Tint && Reflectivity == o.Reflectivity && Transparent == o.Transparent;
}

// ---------------------------------------------------------------------------------------------------- shapes
ShapePtr Sphere::NewSphere(const Vector& c, double r, const Material& material) {  // Sphere.cs:27-33
    auto s = std::make_shared<Sphere>();
    s->Center = c; s->Radius = r; s->Mat = material;
    s->box = Box(Vector(c.X() - r, c.Y() - r, c.Z() - r), Vector(c.X() + r, c.Y() + r, c.Z() + r));
    return s;
}
ShapePtr Cube::NewCube(const Vector& mn, const Vector& mx, const Material& material) {  // Cube.cs:25-29
    auto s = std::make_shared<Cube>();
    s->Min = mn; s->Max = mx; s->Mat = material;
    return s;
}
ShapePtr Plane::NewPlane(const Vector& point, const Vector& normal, const Material& material) {  // Plane.cs:26-29
    auto s = std::make_shared<Plane>();
    s->Point = point; s->Normal = Normalize(normal); s->Mat = material;
    return s;
}
Box Plane::BoundingBox() const { return Box(Vector(-1e9, -1e9, -1e9), Vector(1e9, 1e9, 1e9)); }  // Plane.cs:33-36, Util.INF
ShapePtr Cylinder::NewCylinder(double radius, double z0, double z1, const Material& material) {
    auto s = std::make_shared<Cylinder>();
    s->Radius = radius; s->Z0 = z0; s->Z1 = z1; s->Mat = material;
    return s;
}
// Cylinder.cs:22-35.  `new Matrix().Rotate(u, a).Translate(v0)` evaluates to Translate(v0): Matrix.Translate
// ignores its receiver (Matrix.cs:33-36), so the rotation is computed and dropped.
ShapePtr Cylinder::NewTransformedCylinder(const Vector& v0, const Vector& v1, double radius, const Material& material) {
    Vector d = Sub(v1, v0);
    double z = Length(d);
    Matrix m = Matrix::Translate(v0);
    return TransformedShape::NewTransformedShape(NewCylinder(radius, 0, z, material), m);
}
Box Cylinder::BoundingBox() const { return Box(Vector(-Radius, -Radius, Z0), Vector(Radius, Radius, Z1)); }  // Cylinder.cs:37-41

Box Triangle::BoundingBox() const {  // Triangle.cs:80-85
    return Box(ptsharp::Min(ptsharp::Min(V1, V2), V3), ptsharp::Max(ptsharp::Max(V1, V2), V3));
}
void Triangle::FixNormals() {  // Triangle.cs:198-203, 224-237
    Vector n = Normalize(Cross(Sub(V2, V1), Sub(V3, V1)));
    auto isZero = [](const Vector& v) { return v.x == 0 && v.y == 0 && v.z == 0; };
    if (isZero(N1)) N1 = n;
    if (isZero(N2)) N2 = n;
    if (isZero(N3)) N3 = n;
}
std::shared_ptr<Mesh> Mesh::NewMesh(std::vector<Triangle> triangles) {
    auto m = std::make_shared<Mesh>();
    m->Triangles = std::move(triangles);
    return m;
}
void Mesh::Compile() {  // Mesh.cs:45-57
    if (tree) return;
    std::vector<Box> boxes(Triangles.size());
    for (size_t i = 0; i < Triangles.size(); i++) boxes[i] = Triangles[i].BoundingBox();
    tree = Tree::NewTree(boxes);
}
Box Mesh::BoundingBox() const {  // Mesh.cs:88-103
    if (!haveBox) {
        Vector mn = Triangles[0].V1, mx = Triangles[0].V1;
        for (const Triangle& t : Triangles) {
            mn = ptsharp::Min(ptsharp::Min(ptsharp::Min(mn, t.V1), t.V2), t.V3);
            mx = ptsharp::Max(ptsharp::Max(ptsharp::Max(mx, t.V1), t.V2), t.V3);
        }
        box = Box(mn, mx);
        haveBox = true;
    }
    return box;
}
ShapePtr TransformedShape::NewTransformedShape(ShapePtr s, const Matrix& m) {  // TransformedShape.cs:31-34
    auto t = std::make_shared<TransformedShape>();
    t->Shape = std::move(s); t->M = m; t->Inv = m.Inverse();
    return t;
}

// ---- SDF nodes: host keeps parameters + bounding boxes and emits the device program ------------------------------
namespace {
ptgpu_sdf_op Op(uint32_t op, uint32_t n = 0) { ptgpu_sdf_op o; std::memset(&o, 0, sizeof(o)); o.op = op; o.n = n; return o; }
struct SphereSDF : SDF {
    double Radius, Exponent = 2;
    Box BoundingBox() const override { double r = Radius; return Box(Vector(-r, -r, -r), Vector(r, r, r)); }  // SDF.cs:133-137
    void Emit(std::vector<ptgpu_sdf_op>& p) const override { auto o = Op(PTGPU_SDF_SPHERE); o.p[0] = Radius; o.p[1] = Exponent; p.push_back(o); }
};
struct CubeSDF : SDF {
    Vector Size;
    Box BoundingBox() const override { double x = Size.X() / 2, y = Size.Y() / 2, z = Size.Z() / 2; return Box(Vector(-x, -y, -z), Vector(x, y, z)); }  // SDF.cs:190-194
    void Emit(std::vector<ptgpu_sdf_op>& p) const override { auto o = Op(PTGPU_SDF_CUBE); o.p[0] = Size.X(); o.p[1] = Size.Y(); o.p[2] = Size.Z(); p.push_back(o); }
};
struct CylinderSDF : SDF {
    double Radius, Height;
    Box BoundingBox() const override { double r = Radius, h = Height / 2; return Box(Vector(-r, -h, -r), Vector(r, h, r)); }  // SDF.cs:213-225
    void Emit(std::vector<ptgpu_sdf_op>& p) const override { auto o = Op(PTGPU_SDF_CYLINDER); o.p[0] = Radius; o.p[1] = Height; p.push_back(o); }
};
struct CapsuleSDF : SDF {
    Vector A, B; double Radius, Exponent = 2;
    Box BoundingBox() const override {  // SDF.cs:280-284
        Vector a = Min(A, B), b = Max(A, B);
        return Box(Vector(a.X() - Radius, a.Y() - Radius, a.Z() - Radius), Vector(b.X() + Radius, b.Y() + Radius, b.Z() + Radius));
    }
    void Emit(std::vector<ptgpu_sdf_op>& p) const override {
        auto o = Op(PTGPU_SDF_CAPSULE);
        o.p[0] = A.X(); o.p[1] = A.Y(); o.p[2] = A.Z(); o.p[3] = B.X(); o.p[4] = B.Y(); o.p[5] = B.Z(); o.p[6] = Radius; o.p[7] = Exponent;
        p.push_back(o);
    }
};
struct TorusSDF : SDF {
    double Major, Minor, MajorExp = 2, MinorExp = 2;
    Box BoundingBox() const override { double a = Minor, b = Minor + Major; return Box(Vector(-b, -b, a), Vector(b, b, a)); }  // SDF.cs:313-318 (sic)
    void Emit(std::vector<ptgpu_sdf_op>& p) const override { auto o = Op(PTGPU_SDF_TORUS); o.p[0] = Major; o.p[1] = Minor; o.p[2] = MajorExp; o.p[3] = MinorExp; p.push_back(o); }
};
struct TransformSDF : SDF {
    SDFPtr Inner; Matrix M, Inv;
    Box BoundingBox() const override { return M.MulBox(Inner->BoundingBox()); }  // SDF.cs:345-353
    void Emit(std::vector<ptgpu_sdf_op>& p) const override {
        auto o = Op(PTGPU_SDF_PUSH_TRANSFORM);
        std::memcpy(o.p, Inv.m, sizeof(Inv.m));
        p.push_back(o);
        Inner->Emit(p);
        p.push_back(Op(PTGPU_SDF_POP, 0));
    }
};
struct ScaleSDF : SDF {
    SDFPtr Inner; double Factor;
    Box BoundingBox() const override { double f = Factor; return Matrix::Scale(Vector(f, f, f)).MulBox(Inner->BoundingBox()); }  // SDF.cs:376-381
    void Emit(std::vector<ptgpu_sdf_op>& p) const override {
        auto o = Op(PTGPU_SDF_PUSH_SCALE); o.p[0] = Factor; p.push_back(o);
        Inner->Emit(p);
        auto q = Op(PTGPU_SDF_POP, 1); q.p[0] = Factor; p.push_back(q);
    }
};
struct RepeatSDF : SDF {
    SDFPtr Inner; Vector Step;
    Box BoundingBox() const override { return Box(); }  // SDF.cs:555-558 (sic)
    void Emit(std::vector<ptgpu_sdf_op>& p) const override {
        auto o = Op(PTGPU_SDF_PUSH_REPEAT); o.p[0] = Step.X(); o.p[1] = Step.Y(); o.p[2] = Step.Z(); p.push_back(o);
        Inner->Emit(p);
        p.push_back(Op(PTGPU_SDF_POP, 0));
    }
};
struct CombineSDF : SDF {
    uint32_t op; std::vector<SDFPtr> Items;
    Box BoundingBox() const override {
        if (op == PTGPU_SDF_DIFFERENCE) return Items[0]->BoundingBox();  // SDF.cs:478-481
        Box result; int i = 0;                                             // SDF.cs:414-436, 511-532
        for (auto& it : Items) { Box b = it->BoundingBox(); result = i == 0 ? b : result.Extend(b); i++; }
        return result;
    }
    void Emit(std::vector<ptgpu_sdf_op>& p) const override {
        for (auto& it : Items) it->Emit(p);
        p.push_back(Op(op, (uint32_t)Items.size()));
    }
};
}  // namespace
SDFPtr NewSphereSDF(double radius) { auto s = std::make_shared<SphereSDF>(); s->Radius = radius; return s; }
SDFPtr NewCubeSDF(const Vector& size) { auto s = std::make_shared<CubeSDF>(); s->Size = size; return s; }
SDFPtr NewCylinderSDF(double radius, double height) { auto s = std::make_shared<CylinderSDF>(); s->Radius = radius; s->Height = height; return s; }
SDFPtr NewCapsuleSDF(const Vector& a, const Vector& b, double radius) { auto s = std::make_shared<CapsuleSDF>(); s->A = a; s->B = b; s->Radius = radius; return s; }
SDFPtr NewTorusSDF(double major, double minor) { auto s = std::make_shared<TorusSDF>(); s->Major = major; s->Minor = minor; return s; }
SDFPtr NewTransformSDF(SDFPtr sdf, const Matrix& m) { auto s = std::make_shared<TransformSDF>(); s->Inner = sdf; s->M = m; s->Inv = m.Inverse(); return s; }
SDFPtr NewScaleSDF(SDFPtr sdf, double factor) { auto s = std::make_shared<ScaleSDF>(); s->Inner = sdf; s->Factor = factor; return s; }
SDFPtr NewRepeaterSDF(SDFPtr sdf, const Vector& step) { auto s = std::make_shared<RepeatSDF>(); s->Inner = sdf; s->Step = step; return s; }
static SDFPtr Combine(uint32_t op, std::vector<SDFPtr> items) { auto s = std::make_shared<CombineSDF>(); s->op = op; s->Items = std::move(items); return s; }
SDFPtr NewUnionSDF(std::vector<SDFPtr> items) { return Combine(PTGPU_SDF_UNION, std::move(items)); }
SDFPtr NewDifferenceSDF(std::vector<SDFPtr> items) { return Combine(PTGPU_SDF_DIFFERENCE, std::move(items)); }
SDFPtr NewIntersectionSDF(std::vector<SDFPtr> items) { return Combine(PTGPU_SDF_INTERSECTION, std::move(items)); }
ShapePtr SDFShape::NewSDFShape(SDFPtr sdf, const Material& material) {
    auto s = std::make_shared<SDFShape>(); s->Sdf = std::move(sdf); s->Mat = material; return s;
}

ShapePtr Volume::NewVolume(const Box& box, int w, int h, int d, double zscale, std::vector<double> data, std::vector<VolumeWindow> windows) {
    auto v = std::make_shared<Volume>();
    v->box = box; v->W = w; v->H = h; v->D = d; v->ZScale = zscale; v->Data = std::move(data); v->Windows = std::move(windows);
    return v;
}
double Volume::Sample(double x, double y, double z) const {  // Volume.cs:73-104 (index quirks kept)
    auto get = [&](int xi, int yi, int zi) -> double {
        if (xi < 0 || yi < 0 || zi < 0 || xi >= W || yi >= H || zi >= D) return 0;
        return Data[(size_t)xi + (size_t)yi * W + (size_t)zi * W * H];
    };
    z /= ZScale;
    x = ((x + 1) / 2) * (double)W;
    y = ((z + 1) / 2) * (double)H;
    z = ((z + 2) / 2) * (double)D;
    int x0 = (int)std::floor(x), y0 = (int)std::floor(y), z0 = (int)std::floor(z), x1 = x0 + 1, y1 = y0 + 1, z1 = z0 + 1;
    double v000 = get(x0, y0, z0), v001 = get(x0, y0, z1), v010 = get(x0, y1, z0), v011 = get(x0, y1, z1);
    double v100 = get(x1, y0, z0), v101 = get(x1, y0, z1), v110 = get(x1, y1, z0), v111 = get(x1, y1, z1);
    x -= x0; y -= y0; z -= z0;
    double c00 = v000 * (1 - x) + v100 * x, c01 = v001 * (1 - x) + v101 * x, c10 = v010 * (1 - x) + v110 * x, c11 = v011 * (1 - x) + v111 * x;
    double c0 = c00 * (1 - y) + c10 * y, c1 = c01 * (1 - y) + c11 * y;
    return c0 * (1 - z) + c1 * z;
}
Material Volume::MaterialAt(const Vector& p) const {  // Volume.cs:147-167
    double be = (double)1e9f;
    Material bm;
    double s = Sample(p.X(), p.Y(), p.Z());
    for (const VolumeWindow& w : Windows) {
        if (s >= w.Lo && s <= w.Hi) return w.VolumeWindowMaterial;
        double e = NetMin(std::fabs(s - w.Lo), std::fabs(s - w.Hi));
        if (e < be) { be = e; bm = w.VolumeWindowMaterial; }
    }
    return bm;
}

// ---------------------------------------------------------------------------------------------------- kd-tree builder
// Tree.cs:201-265 on an array of bounding boxes.  `items` is the node's Shapes array (order matters).
namespace {
struct Builder {
    const std::vector<Box>& boxes;
    Tree& t;
    // Median() of the ConcurrentBag holding min0,max0,min1,max1,... (Tree.cs:130-148, 212-220).  The bag enumerates
    // LIFO, so its elements N-1 and N are insertion slots N and N-1: (slot[N] + slot[N-1]) / 2.
    double Median(const std::vector<uint32_t>& items, int axis) const {
        size_t n = items.size();
        auto slot = [&](size_t k) { const Box& b = boxes[items[k / 2]]; return (k & 1) ? b.Max.axis(axis) : b.Min.axis(axis); };
        return (slot(n) + slot(n - 1)) / 2;
    }
    int Score(const std::vector<uint32_t>& items, int axis, double point) const {  // Tree.cs:150-175
        int left = 0, right = 0;
        for (uint32_t i : items) {
            const Box& b = boxes[i];
            if (b.Min.axis(axis) <= point) left++;
            if (b.Max.axis(axis) >= point) right++;
        }
        return left >= right ? left : right;
    }
    uint32_t Build(std::vector<uint32_t>& items, uint32_t depth) {
        uint32_t me = (uint32_t)t.nodes.size();
        t.nodes.push_back(ptgpu_node{0.0, 0u, 0u});
        if (depth > t.maxDepth) t.maxDepth = depth;
        int bestAxis = 0;
        double bestPoint = 0;
        size_t n = items.size();
        if (n >= 8) {
            double mx = Median(items, 1), my = Median(items, 2), mz = Median(items, 3);
            int best = (int)((double)n * 0.85);
            int sx = Score(items, 1, mx);
            if (sx < best) { best = sx; bestAxis = 1; bestPoint = mx; }
            int sy = Score(items, 2, my);
            if (sy < best) { best = sy; bestAxis = 2; bestPoint = my; }
            int sz = Score(items, 3, mz);
            if (sz < best) { best = sz; bestAxis = 3; bestPoint = mz; }
        }
        if (bestAxis == 0) {  // leaf keeps its Shapes in array order
            t.nodes[me].a = ((uint32_t)t.leafItems.size() << 2);
            t.nodes[me].b = (uint32_t)n;
            t.leafItems.insert(t.leafItems.end(), items.begin(), items.end());
            return me;
        }
        // Partition (Tree.cs:177-199): each side is a ConcurrentBag.ToArray() = reverse of insertion order.
        std::vector<uint32_t> l, r;
        for (size_t k = n; k-- > 0;) {
            const Box& b = boxes[items[k]];
            if (b.Min.axis(bestAxis) <= bestPoint) l.push_back(items[k]);
            if (b.Max.axis(bestAxis) >= bestPoint) r.push_back(items[k]);
        }
        std::vector<uint32_t>().swap(items);  // Shapes = null (Tree.cs:264), before recursing to bound memory
        uint32_t li = Build(l, depth + 1);
        uint32_t ri = Build(r, depth + 1);
        t.nodes[me].split = bestPoint;
        t.nodes[me].a = (li << 2) | (uint32_t)bestAxis;
        t.nodes[me].b = ri;
        return me;
    }
};
}  // namespace

std::shared_ptr<Tree> Tree::NewTree(const std::vector<Box>& shapeBoxes) {  // Tree.cs:22-29, Box.cs:20-32
    auto t = std::make_shared<Tree>();
    if (!shapeBoxes.empty()) {
        Box box = shapeBoxes[0];
        for (const Box& b : shapeBoxes) box = box.Extend(b);
        t->box = box;
    }
    std::vector<uint32_t> items(shapeBoxes.size());
    for (size_t i = 0; i < items.size(); i++) items[i] = (uint32_t)i;
    Builder b{shapeBoxes, *t};
    b.Build(items, 0);
    return t;
}

// ---------------------------------------------------------------------------------------------------- builder-friendly order
// The reference builder takes each node's split position from the shapes that happen to sit in the middle of the
// node's array (Median() of an unsorted bag, Tree.cs:130-148), so the tree it builds is a function of input order:
// an arbitrary order leaves leaves of thousands of triangles.  This authoring-side utility permutes a
// mesh so that the UNMODIFIED builder produces a balanced tree.
//
// One depth-first pass runs the builder's own recursion while synthesising the root order.  A node's array is, by
// construction of the builder, the root order restricted to the node's set (reversed once per level), so it is derived
// from the current root order when the node is visited.  When the middle slot(s) of a node would give no accepted
// split, or a lopsided one, a triangle of the node lying on its median plane is swapped into the slot.  A swap of
// root positions g1 <-> g2 changes another node's middle element only if that node holds exactly one of the two
// triangles and its middle lies between g1 and g2.  A triangle that never straddled a split plane lives only in the
// current node's ancestor chain (which holds both), so it can always move; for a triangle that did straddle, the
// middle positions of every decided node holding it are remembered and checked.  Decisions of nodes with at least
// `protectMin` shapes therefore stay valid, and above that size the tree simulated here is exactly the tree
// Tree.NewTree builds from the returned order (smaller subtrees may differ; they are leaves-in-waiting anyway).
namespace {
struct Orderer {
    const std::vector<Box>& boxes;
    std::vector<uint32_t> order;   // root order being synthesised
    std::vector<uint32_t> pos;     // pos[e] = index of e in `order`
    std::vector<uint8_t> locked;   // e is the middle element of a decided node
    std::vector<uint8_t> shared;   // e straddled a decided split plane (lives in more than one branch)
    std::vector<uint8_t> onPath;   // per decided node: it is an ancestor of the node being visited
    struct Entry { uint32_t p0, p1; uint32_t node; int32_t next; };  // a decided node holding the element: its middle elements
    std::vector<Entry> entries;
    std::vector<int32_t> head;     // per element: list of decided nodes that held it while it was shared
    double balance;
    size_t minRepair, protectMin;
    long long repairs = 0, blocked = 0, blockedLocked = 0, failed = 0, failedItems = 0, maxFailed = 0;

    bool sideOk(uint32_t e, size_t g1, size_t g2) const {  // moving e from g1 to g2 keeps every decided middle in place
        for (int32_t k = head[e]; k >= 0; k = entries[k].next) {
            const Entry& en = entries[k];
            if (onPath[en.node]) continue;  // an ancestor: holds both swapped elements
            size_t a = pos[en.p0], b = pos[en.p1], lo = std::min(a, b), hi = std::max(a, b);
            bool below = g1 < lo && g2 < lo, above = g1 > hi && g2 > hi;
            if (!(below || above)) return false;
        }
        return true;
    }
    bool swapAllowed(uint32_t ea, uint32_t eb) const {
        size_t ga = pos[ea], gb = pos[eb];
        return sideOk(ea, ga, gb) && sideOk(eb, gb, ga);
    }
    double slotValue(const std::vector<uint32_t>& items, size_t k, int axis) const {
        const Box& b = boxes[items[k / 2]];
        return (k & 1) ? b.Max.axis(axis) : b.Min.axis(axis);
    }
    int score(const std::vector<uint32_t>& items, int axis, double point) const {
        int left = 0, right = 0;
        for (uint32_t i : items) { const Box& b = boxes[i]; if (b.Min.axis(axis) <= point) left++; if (b.Max.axis(axis) >= point) right++; }
        return left >= right ? left : right;
    }
    int decide(const std::vector<uint32_t>& items, double& point, int& bestScore) const {  // Tree.cs:222-255
        size_t n = items.size();
        int best = (int)((double)n * 0.85), bestAxis = 0;
        for (int axis = 1; axis <= 3; axis++) {
            double m = (slotValue(items, n, axis) + slotValue(items, n - 1, axis)) / 2;
            int s = score(items, axis, m);
            if (s < best) { best = s; bestAxis = axis; point = m; }
        }
        bestScore = best;
        return bestAxis;
    }
    void swapGlobal(std::vector<uint32_t>& items, size_t i, size_t j) {
        if (i == j) return;
        uint32_t a = items[i], b = items[j];
        std::swap(items[i], items[j]);
        std::swap(order[pos[a]], order[pos[b]]);
        std::swap(pos[a], pos[b]);
    }
    bool good(const std::vector<uint32_t>& items) const {
        double point; int sc;
        int ax = decide(items, point, sc);
        return ax != 0 && sc < (int)((double)items.size() * balance);
    }
    // Put into `slot` the movable triangle of the node whose key (Min, Max or centre along `axis`) is nearest `target`.
    // role: 0 = Min, 1 = Max, 2 = centre.  Unshared triangles are preferred among the nearest few.
    bool place(std::vector<uint32_t>& items, size_t slot, int axis, int role, double target, uint32_t avoid) {
        const size_t n = items.size();
        const uint32_t cur = items[slot];
        auto key = [&](uint32_t e) {
            const Box& b = boxes[e];
            return role == 0 ? b.Min.axis(axis) : role == 1 ? b.Max.axis(axis) : 0.5 * (b.Min.axis(axis) + b.Max.axis(axis));
        };
        std::vector<std::pair<double, uint32_t>> cand;
        cand.reserve(n);
        for (size_t i = 0; i < n; i++) {
            uint32_t e = items[i];
            if (locked[e] || e == avoid) continue;
            cand.emplace_back(std::fabs(key(e) - target), e);
        }
        if (cand.empty()) return false;
        const size_t K = std::min<size_t>(cand.size(), 64);
        std::partial_sort(cand.begin(), cand.begin() + K, cand.end());
        // among the K nearest, try unshared ones that are almost as near as the best first
        const double tol = cand[0].first * 2 + 1e-12;
        for (int phase = 0; phase < 2; phase++) {
            for (size_t k = 0; k < K; k++) {
                uint32_t e = cand[k].second;
                if (phase == 0 && (shared[e] || cand[k].first > tol)) continue;
                if (e == cur) return true;
                if (!swapAllowed(cur, e)) continue;
                size_t j = 0;
                for (j = 0; j < n; j++) if (items[j] == e) break;
                swapGlobal(items, slot, j);
                return true;
            }
        }
        return false;
    }
    void repair(std::vector<uint32_t>& items) {
        const size_t n = items.size();
        const size_t s0 = (n % 2 == 0) ? n / 2 - 1 : (n - 1) / 2, s1 = (n % 2 == 0) ? n / 2 : (n - 1) / 2;
        const bool lock0 = locked[items[s0]], lock1 = locked[items[s1]];
        if (lock0 && lock1) { blocked++; blockedLocked++; return; }
        double lo[4], hi[4];
        for (int a = 1; a <= 3; a++) { lo[a] = 1e300; hi[a] = -1e300; }
        for (uint32_t e : items)
            for (int a = 1; a <= 3; a++) { double c = 0.5 * (boxes[e].Min.axis(a) + boxes[e].Max.axis(a)); lo[a] = std::min(lo[a], c); hi[a] = std::max(hi[a], c); }
        int axes[3] = {1, 2, 3};
        std::sort(axes, axes + 3, [&](int x, int y) { return (hi[x] - lo[x]) > (hi[y] - lo[y]); });
        std::vector<double> tmp(n);
        bool any = false;
        for (int t = 0; t < 3; t++) {
            const int axis = axes[t];
            for (size_t i = 0; i < n; i++) tmp[i] = 0.5 * (boxes[items[i]].Min.axis(axis) + boxes[items[i]].Max.axis(axis));
            std::nth_element(tmp.begin(), tmp.begin() + n / 2, tmp.end());
            const double plane = tmp[n / 2];
            if (s0 == s1) {  // odd N: the plane is the centre of the middle triangle's box
                any |= place(items, s0, axis, 2, plane, 0xFFFFFFFFu);
            } else {  // even N: plane = (Min of slot N/2 + Max of slot N/2-1) / 2
                if (!lock0 && !lock1) {
                    any |= place(items, s0, axis, 1, plane, 0xFFFFFFFFu);
                    any |= place(items, s1, axis, 0, 2 * plane - boxes[items[s0]].Max.axis(axis), items[s0]);
                } else if (lock0) {
                    any |= place(items, s1, axis, 0, 2 * plane - boxes[items[s0]].Max.axis(axis), items[s0]);
                } else {
                    any |= place(items, s0, axis, 1, 2 * plane - boxes[items[s1]].Min.axis(axis), items[s1]);
                }
            }
            if (good(items)) break;
        }
        if (any) repairs++; else blocked++;
    }
    void run(std::vector<uint32_t>& items, int depth) {
        const size_t n = items.size();
        if (n < 8) return;
        if (n >= minRepair && !good(items)) repair(items);
        double point = 0; int sc = 0;
        const int axis = decide(items, point, sc);
        const size_t s0 = (n % 2 == 0) ? n / 2 - 1 : (n - 1) / 2, s1 = (n % 2 == 0) ? n / 2 : (n - 1) / 2;
        const uint32_t p0 = items[s0], p1 = items[s1];
        locked[p0] = locked[p1] = 1;
        const uint32_t me = (uint32_t)onPath.size();
        onPath.push_back(1);
        if (n >= protectMin) {
            for (uint32_t e : items)
                if (shared[e]) { entries.push_back(Entry{p0, p1, me, head[e]}); head[e] = (int32_t)entries.size() - 1; }
        }
        if (axis == 0) {
            failed++; failedItems += (long long)n; maxFailed = std::max<long long>(maxFailed, (long long)n);
        } else {
            std::vector<uint32_t> l, r;
            for (size_t k = n; k-- > 0;) {
                const Box& b = boxes[items[k]];
                const bool bl = b.Min.axis(axis) <= point, br = b.Max.axis(axis) >= point;
                if (bl && br) shared[items[k]] = 1;
                if (bl) l.push_back(items[k]);
                if (br) r.push_back(items[k]);
            }
            std::vector<uint32_t>().swap(items);
            run(l, depth + 1);  // nothing moved since the partition: l is still the restriction of the root order
            // swaps inside the left subtree may have moved shared triangles: re-derive r from the root order
            if ((depth + 1) % 2 == 1) std::sort(r.begin(), r.end(), [&](uint32_t x, uint32_t y) { return pos[x] > pos[y]; });
            else std::sort(r.begin(), r.end(), [&](uint32_t x, uint32_t y) { return pos[x] < pos[y]; });
            run(r, depth + 1);
        }
        onPath[me] = 0;
    }
};
}  // namespace

std::vector<uint32_t> BuilderFriendlyOrder(const std::vector<Box>& boxes, double balance, int minRepair, int verbose) {
    Orderer o{boxes, {}, {}, {}, {}, {}, {}, {}, balance, (size_t)std::max(8, minRepair), 48};
    size_t n = boxes.size();
    o.order.resize(n); o.pos.resize(n); o.locked.assign(n, 0); o.shared.assign(n, 0); o.head.assign(n, -1);
    for (size_t i = 0; i < n; i++) { o.order[i] = (uint32_t)i; o.pos[i] = (uint32_t)i; }
    std::vector<uint32_t> items(o.order);
    o.run(items, 0);
    if (verbose)
        std::fprintf(stderr, "[BuilderFriendlyOrder] %zu shapes: %lld nodes repaired, %lld could not be (%lld: middle held by another node); %lld oversized leaves holding %lld items (largest %lld)\n",
                     n, o.repairs, o.blocked, o.blockedLocked, o.failed, o.failedItems, o.maxFailed);
    return o.order;
}

// ---------------------------------------------------------------------------------------------------- scene
void Scene::Add(ShapePtr p) {  // Scene.cs:29-38
    Shapes.push_back(p);
    if (p->MaterialAt(Vector()).Emittance > 0) Lights.push_back(p);
}
void Scene::Compile() {  // Scene.cs:48-68
    for (auto& s : Shapes) s->Compile();
    if (!tree) {
        std::vector<Box> boxes(Shapes.size());
        for (size_t i = 0; i < Shapes.size(); i++) boxes[i] = Shapes[i]->BoundingBox();
        tree = Tree::NewTree(boxes);
    }
}
Camera Camera::LookAt(const Vector& eye, const Vector& center, const Vector& up, double fovy) {  // Camera.cs:23-35
    Camera c;
    c.p = eye;
    c.w = Normalize(Sub(center, eye));
    c.u = Normalize(Cross(up, c.w));
    c.v = Normalize(Cross(c.w, c.u));
    c.m = 1 / std::tan(fovy * M_PI / 360);
    return c;
}
void Camera::SetFocus(const Vector& focalPoint, double aperture) {  // Camera.cs:39-43
    focalDistance = Length(Sub(focalPoint, p));
    apertureRadius = aperture;
}

// ---------------------------------------------------------------------------------------------------- flattener
void FlatScene::Bind(uint32_t sceneTree, uint32_t numSceneShapes) {
    ptgpu_flat_scene& v = view;
    const double env[3] = {v.envColor[0], v.envColor[1], v.envColor[2]};
    const int32_t envTex = v.envTexture;
    const double envAngle = v.envTextureAngle;
    std::memset(&v, 0, sizeof(v));
    v.abiVersion = PTGPU_ABI_VERSION;
    v.sceneTree = sceneTree;
    v.numSceneShapes = numSceneShapes;
    v.numShapes = (uint32_t)shapes.size(); v.shapes = shapes.data();
    v.numLights = (uint32_t)lights.size(); v.lights = lights.data();
    v.numTrees = (uint32_t)trees.size(); v.trees = trees.data();
    v.numNodes = nodes.size(); v.nodes = nodes.data();
    v.numLeafItems = leafItems.size(); v.leafItems = leafItems.data();
    v.numSpheres = (uint32_t)spheres.size(); v.spheres = spheres.data();
    v.numCubes = (uint32_t)cubes.size(); v.cubes = cubes.data();
    v.numPlanes = (uint32_t)planes.size(); v.planes = planes.data();
    v.numCylinders = (uint32_t)cylinders.size(); v.cylinders = cylinders.data();
    v.numMeshes = (uint32_t)meshes.size(); v.meshes = meshes.data();
    v.numTriangles = triGeom.size(); v.triGeom = triGeom.data(); v.triShade = triShade.data();
    v.numInstances = (uint32_t)instances.size(); v.instances = instances.data();
    v.numSdfShapes = (uint32_t)sdfShapes.size(); v.sdfShapes = sdfShapes.data();
    v.numSdfOps = (uint32_t)sdfOps.size(); v.sdfOps = sdfOps.data();
    v.numVolumes = (uint32_t)volumes.size(); v.volumes = volumes.data();
    v.numVolumeWindows = (uint32_t)volumeWindows.size(); v.volumeWindows = volumeWindows.data();
    v.numVolumeData = volumeData.size(); v.volumeData = volumeData.data();
    v.numMaterials = (uint32_t)materials.size(); v.materials = materials.data();
    v.numTextures = (uint32_t)textures.size(); v.textures = textures.data();
    v.numTexels = texels.size() / 4; v.texels = texels.data();
    v.numShs = (uint32_t)shs.size(); v.shs = shs.data();
    v.envColor[0] = env[0]; v.envColor[1] = env[1]; v.envColor[2] = env[2];
    v.envTexture = envTex; v.envTextureAngle = envAngle;
}

// Flat-scene file: "PTFS", format version, ABI version, the scalar header, then the 22 arrays in header
// order as (u64 count, u32 element size, bytes).  Little-endian, the in-memory layout of include/ptgpu.h.
namespace {
constexpr uint32_t kFlatMagic = 0x53465450u, kFlatFormat = 2;  // 2: texels as doubles, SphericalHarmonic shapes
template <class T> void put_vec(std::FILE* f, const std::vector<T>& v) {
    const uint64_t n = v.size(); const uint32_t es = (uint32_t)sizeof(T);
    if (std::fwrite(&n, 8, 1, f) != 1 || std::fwrite(&es, 4, 1, f) != 1 || (n && std::fwrite(v.data(), sizeof(T), n, f) != n)) throw std::runtime_error("flat scene: write failed");
}
template <class T> void get_vec(std::FILE* f, std::vector<T>& v) {
    uint64_t n = 0; uint32_t es = 0;
    if (std::fread(&n, 8, 1, f) != 1 || std::fread(&es, 4, 1, f) != 1) throw std::runtime_error("flat scene: truncated file");
    if (es != sizeof(T)) throw std::runtime_error("flat scene: element size mismatch (written by another ABI)");
    v.resize(n);
    if (n && std::fread(v.data(), sizeof(T), n, f) != n) throw std::runtime_error("flat scene: truncated file");
}
template <class F> void each_array(FlatScene& s, F&& fn) {
    fn(s.shapes); fn(s.lights); fn(s.trees); fn(s.nodes); fn(s.leafItems); fn(s.spheres); fn(s.cubes); fn(s.planes); fn(s.cylinders); fn(s.meshes);
    fn(s.triGeom); fn(s.triShade); fn(s.instances); fn(s.sdfShapes); fn(s.sdfOps); fn(s.volumes); fn(s.volumeWindows); fn(s.volumeData);
    fn(s.materials); fn(s.textures); fn(s.texels); fn(s.shs);
}
}  // namespace
void SaveFlatScene(const FlatScene& scene, const std::string& path) {
    std::FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("flat scene: cannot open " + path);
    try {
        const ptgpu_flat_scene& v = scene.view;
        const uint32_t head[5] = {kFlatMagic, kFlatFormat, (uint32_t)PTGPU_ABI_VERSION, v.sceneTree, v.numSceneShapes};
        const double env[4] = {v.envColor[0], v.envColor[1], v.envColor[2], v.envTextureAngle};
        const int32_t envTex = v.envTexture;
        if (std::fwrite(head, 4, 5, f) != 5 || std::fwrite(env, 8, 4, f) != 4 || std::fwrite(&envTex, 4, 1, f) != 1) throw std::runtime_error("flat scene: write failed");
        each_array(const_cast<FlatScene&>(scene), [&](auto& vec) { put_vec(f, vec); });
    } catch (...) { std::fclose(f); throw; }
    std::fclose(f);
}
std::unique_ptr<FlatScene> LoadFlatScene(const std::string& path) {
    std::FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("flat scene: cannot open " + path);
    auto fs = std::make_unique<FlatScene>();
    try {
        uint32_t head[5]; double env[4]; int32_t envTex;
        if (std::fread(head, 4, 5, f) != 5 || std::fread(env, 8, 4, f) != 4 || std::fread(&envTex, 4, 1, f) != 1) throw std::runtime_error("flat scene: truncated file");
        if (head[0] != kFlatMagic) throw std::runtime_error("flat scene: bad magic");
        if (head[1] != kFlatFormat || head[2] != (uint32_t)PTGPU_ABI_VERSION) throw std::runtime_error("flat scene: format / ABI version mismatch");
        each_array(*fs, [&](auto& vec) { get_vec(f, vec); });
        if (fs->triShade.size() != fs->triGeom.size()) throw std::runtime_error("flat scene: triangle arrays differ in length");
        fs->view.envColor[0] = env[0]; fs->view.envColor[1] = env[1]; fs->view.envColor[2] = env[2];
        fs->view.envTextureAngle = env[3]; fs->view.envTexture = envTex;
        fs->Bind(head[3], head[4]);
    } catch (...) { std::fclose(f); throw; }
    std::fclose(f);
    return fs;
}

uint64_t FlatScene::Bytes() const {
    auto sz = [](const auto& v) { return (uint64_t)v.size() * sizeof(v[0]); };
    return sz(shapes) + sz(lights) + sz(trees) + sz(nodes) + sz(leafItems) + sz(spheres) + sz(cubes) + sz(planes) + sz(cylinders) +
           sz(meshes) + sz(triGeom) + sz(triShade) + sz(instances) + sz(sdfShapes) + sz(sdfOps) + sz(volumes) + sz(volumeWindows) +
           sz(volumeData) + sz(materials) + sz(textures) + sz(texels) + sz(shs);
}

namespace {
struct Flattener {
    FlatScene& f;
    std::map<const ColorTexture*, int32_t> texIds;
    std::map<const Mesh*, uint32_t> meshIds;
    size_t lastMat = 0;

    static void put3(float* dst, const Vector& v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }

    int32_t TextureId(const ITexture& t) {
        if (!t) return -1;
        auto it = texIds.find(t.get());
        if (it != texIds.end()) return it->second;
        ptgpu_texture pt;
        pt.width = t->Width; pt.height = t->Height; pt.texelOffset = f.texels.size() / 4;
        for (const Colour& c : t->Data) { f.texels.push_back(c.r); f.texels.push_back(c.g); f.texels.push_back(c.b); f.texels.push_back(1.0); }
        int32_t id = (int32_t)f.textures.size();
        f.textures.push_back(pt);
        texIds[t.get()] = id;
        return id;
    }
    std::vector<Material> seen;
    int32_t MaterialId(const Material& m) {
        if (lastMat < seen.size() && seen[lastMat].SameAs(m)) return (int32_t)lastMat;
        for (size_t i = 0; i < seen.size(); i++)
            if (seen[i].SameAs(m)) { lastMat = i; return (int32_t)i; }
        ptgpu_material pm;
        std::memset(&pm, 0, sizeof(pm));
        pm.color[0] = m.Color.r; pm.color[1] = m.Color.g; pm.color[2] = m.Color.b;
        pm.bumpMultiplier = m.BumpMultiplier; pm.emittance = m.Emittance; pm.index = m.Index; pm.gloss = m.Gloss;
        pm.tint = m.Tint; pm.reflectivity = m.Reflectivity; pm.transparent = m.Transparent ? 1 : 0;
        pm.texture = TextureId(m.Texture); pm.normalTexture = TextureId(m.NormalTexture);
        pm.bumpTexture = TextureId(m.BumpTexture); pm.glossTexture = TextureId(m.GlossTexture);
        seen.push_back(m);
        f.materials.push_back(pm);
        lastMat = seen.size() - 1;
        return (int32_t)lastMat;
    }
    // Append `t` to the global node / leaf-item arrays; itemBase is added to every leaf item.
    uint32_t AddTree(const Tree& t, uint32_t itemBase) {
        uint64_t nodeBase = f.nodes.size(), leafBase = f.leafItems.size();
        if (nodeBase + t.nodes.size() >= (1ull << 30) || leafBase + t.leafItems.size() >= (1ull << 30))
            throw std::runtime_error("kd-tree too large for 30-bit node/leaf indices");
        for (const ptgpu_node& n : t.nodes) {
            ptgpu_node o = n;
            uint32_t axis = n.a & 3u;
            if (axis) { o.a = (((n.a >> 2) + (uint32_t)nodeBase) << 2) | axis; o.b = n.b + (uint32_t)nodeBase; }
            else { o.a = ((n.a >> 2) + (uint32_t)leafBase) << 2; }
            f.nodes.push_back(o);
        }
        for (uint32_t it : t.leafItems) f.leafItems.push_back(it + itemBase);
        ptgpu_tree pt;
        put3(pt.bmin, t.box.Min); put3(pt.bmax, t.box.Max);
        pt.root = (uint32_t)nodeBase; pt.maxDepth = t.maxDepth;
        f.trees.push_back(pt);
        return (uint32_t)f.trees.size() - 1;
    }
    // anonymous: the triangles' own materials are never looked at (SphericalHarmonic: the Hit names the solid, SH.cs:54) - keep them out of materials[]
    uint32_t MeshId(const Mesh* m, bool anonymous = false) {
        auto it = meshIds.find(m);
        if (it != meshIds.end()) return it->second;
        if (!m->tree) throw std::runtime_error("Mesh not compiled: call Scene.Compile() first");
        ptgpu_mesh pm;
        pm.triFirst = (uint32_t)f.triGeom.size(); pm.triCount = (uint32_t)m->Triangles.size(); pm.pad = 0;
        for (const Triangle& t : m->Triangles) {
            ptgpu_tri_geom g; std::memset(&g, 0, sizeof(g));
            put3(g.v1, t.V1); put3(g.e1, Sub(t.V2, t.V1)); put3(g.e2, Sub(t.V3, t.V1));  // Triangle.cs:97-98
            f.triGeom.push_back(g);
            ptgpu_tri_shade s;
            put3(s.n1, t.N1); put3(s.n2, t.N2); put3(s.n3, t.N3);
            s.t1[0] = t.T1.x; s.t1[1] = t.T1.y; s.t2[0] = t.T2.x; s.t2[1] = t.T2.y; s.t3[0] = t.T3.x; s.t3[1] = t.T3.y;
            s.material = anonymous ? -1 : MaterialId(t.Mat);
            f.triShade.push_back(s);
        }
        pm.tree = AddTree(*m->tree, pm.triFirst);
        uint32_t id = (uint32_t)f.meshes.size();
        f.meshes.push_back(pm);
        meshIds[m] = id;
        return id;
    }
    ptgpu_shape Describe(const IShape* s, bool nested) {
        ptgpu_shape ps; ps.type = (uint32_t)s->Type(); ps.data = 0; ps.material = -1; ps.flags = s->IsClass() ? 1u : 0u;
        switch (s->Type()) {
            case PTGPU_SPHERE: {
                auto* q = static_cast<const Sphere*>(s);
                ptgpu_sphere d; std::memset(&d, 0, sizeof(d)); put3(d.center, q->Center); d.radius = q->Radius;
                ps.data = (uint32_t)f.spheres.size(); f.spheres.push_back(d); ps.material = MaterialId(q->Mat); break;
            }
            case PTGPU_CUBE: {
                auto* q = static_cast<const Cube*>(s);
                ptgpu_cube d; put3(d.min, q->Min); put3(d.max, q->Max);
                ps.data = (uint32_t)f.cubes.size(); f.cubes.push_back(d); ps.material = MaterialId(q->Mat); break;
            }
            case PTGPU_PLANE: {
                auto* q = static_cast<const Plane*>(s);
                ptgpu_plane d; put3(d.point, q->Point); put3(d.normal, q->Normal);
                ps.data = (uint32_t)f.planes.size(); f.planes.push_back(d); ps.material = MaterialId(q->Mat); break;
            }
            case PTGPU_CYLINDER: {
                auto* q = static_cast<const Cylinder*>(s);
                ptgpu_cylinder d{q->Radius, q->Z0, q->Z1};
                ps.data = (uint32_t)f.cylinders.size(); f.cylinders.push_back(d); ps.material = MaterialId(q->Mat); break;
            }
            case PTGPU_MESH: ps.data = MeshId(static_cast<const Mesh*>(s)); break;
            case PTGPU_TRANSFORMED: {
                auto* q = static_cast<const TransformedShape*>(s);
                ptgpu_instance d; std::memset(&d, 0, sizeof(d));
                std::memcpy(d.m, q->M.m, sizeof(d.m));
                Matrix inv = q->M.Inverse();  // TransformedShape.cs:45 recomputes Matrix.Inverse() per ray; same value every time
                std::memcpy(d.inv, inv.m, sizeof(d.inv));
                ptgpu_shape inner = Describe(q->Shape.get(), true);
                d.pad[0] = inner.type == PTGPU_TRANSFORMED ? 1u : 0u;  // nested: the device re-measures T level by level (nested_fold)
                d.shape = (uint32_t)f.shapes.size();
                f.shapes.push_back(inner);
                ps.data = (uint32_t)f.instances.size(); f.instances.push_back(d); break;
            }
            case PTGPU_SDF: {
                auto* q = static_cast<const SDFShape*>(s);
                ptgpu_sdf_shape d; d.progFirst = (uint32_t)f.sdfOps.size();
                q->Sdf->Emit(f.sdfOps);
                d.progCount = (uint32_t)f.sdfOps.size() - d.progFirst;
                Box b = q->Sdf->BoundingBox(); put3(d.bmin, b.Min); put3(d.bmax, b.Max);
                ps.data = (uint32_t)f.sdfShapes.size(); f.sdfShapes.push_back(d); ps.material = MaterialId(q->Mat); break;
            }
            case PTGPU_VOLUME: {
                auto* q = static_cast<const Volume*>(s);
                ptgpu_volume d; std::memset(&d, 0, sizeof(d));
                d.w = q->W; d.h = q->H; d.d = q->D; d.zscale = q->ZScale;
                d.windowFirst = (uint32_t)f.volumeWindows.size(); d.windowCount = (uint32_t)q->Windows.size();
                for (auto& w : q->Windows) f.volumeWindows.push_back(ptgpu_volume_window{w.Lo, w.Hi, MaterialId(w.VolumeWindowMaterial), 0});
                d.dataOffset = f.volumeData.size();
                f.volumeData.insert(f.volumeData.end(), q->Data.begin(), q->Data.end());
                put3(d.bmin, q->box.Min); put3(d.bmax, q->box.Max);
                ps.data = (uint32_t)f.volumes.size(); f.volumes.push_back(d); break;
            }
            case PTGPU_SH: {
                auto* q = static_cast<const SphericalHarmonic*>(s);
                ptgpu_sh d; std::memset(&d, 0, sizeof(d));
                d.l = q->L; d.m = q->M; d.mesh = MeshId(q->mesh.get(), true);
                d.positiveMaterial = MaterialId(q->PositiveMaterial); d.negativeMaterial = MaterialId(q->NegativeMaterial);
                ps.data = (uint32_t)f.shs.size(); f.shs.push_back(d); break;
            }
            default: throw std::runtime_error("unknown shape type");
        }
        return ps;
    }
};
}  // namespace

std::unique_ptr<FlatScene> Flatten(const Scene& scene) {
    if (!scene.tree) throw std::runtime_error("Scene not compiled");
    auto fs = std::make_unique<FlatScene>();
    FlatScene& f = *fs;
    Flattener fl{f, {}, {}, 0, {}};
    size_t n = scene.Shapes.size();
    f.shapes.resize(n);  // Scene.Shapes first, nested inner shapes after
    for (size_t i = 0; i < n; i++) {
        ptgpu_shape ps = fl.Describe(scene.Shapes[i].get(), false);
        f.shapes[i] = ps;
    }
    for (auto& l : scene.Lights)
        for (size_t i = 0; i < n; i++)
            if (scene.Shapes[i] == l) { f.lights.push_back((uint32_t)i); break; }
    uint32_t sceneTree = fl.AddTree(*scene.tree, 0);
    f.Bind(sceneTree, (uint32_t)n);
    ptgpu_flat_scene& v = f.view;
    v.envColor[0] = scene.Color.r; v.envColor[1] = scene.Color.g; v.envColor[2] = scene.Color.b;
    v.envTexture = fl.TextureId(scene.Texture);
    if (v.envTexture >= 0) {  // TextureId may have grown the arrays
        v.numTextures = (uint32_t)f.textures.size(); v.textures = f.textures.data();
        v.numTexels = f.texels.size() / 4; v.texels = f.texels.data();
    }
    v.envTextureAngle = scene.TextureAngle;
    return fs;
}

ptgpu_camera FlattenCamera(const Camera& c) {
    ptgpu_camera pc;
    Flattener::put3(pc.p, c.p); Flattener::put3(pc.u, c.u); Flattener::put3(pc.v, c.v); Flattener::put3(pc.w, c.w);
    pc.m = c.m; pc.focalDistance = c.focalDistance; pc.apertureRadius = c.apertureRadius;
    return pc;
}

// ---------------------------------------------------------------------------------------------------- render driver
static void Check(ptgpu_ctx* ctx, int rc, const char* what) {
    if (rc != PTGPU_OK) {
        const char* e = ptgpu_last_error(ctx);
        throw std::runtime_error(std::string(what) + ": " + (e ? e : "unknown error"));
    }
}
Renderer Renderer::NewRenderer(Scene& scene, Camera& camera, DefaultSampler& sampler, int w, int h, bool multithreaded) {  // Renderer.cs:35-56
    Renderer r;
    r.NumCPU = multithreaded ? 0 : 1;  // 0 = "ProcessorCount": the parallel path
    r.scene_ = &scene; r.camera_ = &camera; r.sampler_ = &sampler; r.w_ = w; r.h_ = h;
    return r;
}
Renderer::Renderer(Renderer&& o) noexcept
    : SamplesPerPixel(o.SamplesPerPixel), StratifiedSampling(o.StratifiedSampling), AdaptiveSamples(o.AdaptiveSamples),
      FireflySamples(o.FireflySamples), FireflyThreshold(o.FireflyThreshold), AdaptiveThreshold(o.AdaptiveThreshold), AdaptiveExponent(o.AdaptiveExponent),
      NumCPU(o.NumCPU), Device(o.Device), Devices(std::move(o.Devices)), Seed(o.Seed), scene_(o.scene_), camera_(o.camera_),
      sampler_(o.sampler_), w_(o.w_), h_(o.h_), passIndex_(o.passIndex_), ctx_(o.ctx_), flat_(std::move(o.flat_)) {
    o.ctx_ = nullptr;
}
Renderer::~Renderer() { if (ctx_) ptgpu_destroy(ctx_); }
void Renderer::EnsureUploaded() {
    if (!ctx_) {
        ptgpu_params p;
        std::memset(&p, 0, sizeof(p));
        p.device = Device;
        if (Devices.size() > PTGPU_MAX_DEVICES) throw std::runtime_error("Renderer.Devices: at most 8 GPUs");
        p.numDevices = (int)Devices.size();
        for (size_t k = 0; k < Devices.size(); k++) p.devices[k] = Devices[k];
        int rc = ptgpu_create(&p, &ctx_);
        if (rc != PTGPU_OK) { const char* e = ptgpu_last_error(nullptr); throw std::runtime_error(std::string("ptgpu_create: ") + (e ? e : "?")); }
    }
    if (!flat_) {
        scene_->Compile();  // Renderer.cs:208
        flat_ = Flatten(*scene_);
        Check(ctx_, ptgpu_upload_scene(ctx_, &flat_->view), "ptgpu_upload_scene");
    }
}
ptgpu_pass Renderer::MakePass() const {
    ptgpu_pass p;
    std::memset(&p, 0, sizeof(p));
    p.width = w_; p.height = h_; p.spp = SamplesPerPixel; p.stratified = StratifiedSampling ? 1 : 0;
    p.sampleBase = 0; p.sampleStride = 1;
    p.firstHitSamples = sampler_->FirstHitSamples; p.maxBounces = sampler_->MaxBounces;
    p.directLighting = sampler_->DirectLighting; p.softShadows = sampler_->SoftShadows;
    p.lightMode = sampler_->LightMode; p.specularMode = sampler_->SpecularMode;
    p.seed = Seed; p.passIndex = passIndex_;
    p.camera = FlattenCamera(*camera_);
    p.adaptiveSamples = AdaptiveSamples; p.fireflySamples = FireflySamples; p.fireflyThreshold = FireflyThreshold;
    p.serialRules = 0; p.adaptiveThreshold = AdaptiveThreshold; p.adaptiveExponent = AdaptiveExponent;
    p.flags = sampler_->RussianRoulette ? PTGPU_PASS_RUSSIAN_ROULETTE : 0;
    return p;
}
void Renderer::RenderParallel(float* out) {
    EnsureUploaded();
    ptgpu_pass p = MakePass();
    Check(ctx_, ptgpu_render_pass(ctx_, &p, out), "ptgpu_render_pass");
    passIndex_++;
}
void Renderer::Render(float* out) {  // Renderer.cs:80-198
    EnsureUploaded();
    ptgpu_pass p = MakePass();
    p.serialRules = 1;
    Check(ctx_, ptgpu_render_pass(ctx_, &p, out), "ptgpu_render_pass");
    passIndex_++;
}
std::vector<float> Renderer::Image(Channel channel) {
    std::vector<float> img((size_t)w_ * h_ * 3);
    if (!ctx_) throw std::runtime_error("nothing rendered yet");
    Check(ctx_, ptgpu_read_buffer(ctx_, (int)channel, img.data()), "ptgpu_read_buffer");
    return img;
}
ptgpu_counters Renderer::Counters() {
    ptgpu_counters c;
    std::memset(&c, 0, sizeof(c));
    if (ctx_) ptgpu_get_counters(ctx_, &c);
    return c;
}
void Renderer::IterativeRender(const std::string& pathTemplate, int iter) {  // Renderer.cs:702-765
    for (int i = 1; i <= iter; i++) {
        if (NumCPU == 1) Render(nullptr); else RenderParallel(nullptr);  // Renderer.cs:712-719
        std::string path = pathTemplate;
        size_t k = path.find("{0}");
        if (k != std::string::npos) path.replace(k, 3, std::to_string(i));
        std::vector<float> img = Image(ColorChannel);
        FILE* fp = std::fopen(path.c_str(), "wb");
        if (!fp) continue;
        std::fprintf(fp, "P6\n%d %d\n255\n", w_, h_);
        for (size_t q = 0; q < img.size(); q++) {  // Buffer.cs:155-160: Pow(1/2.2), *255, clamp
            double v = std::pow((double)img[q], 1.0 / 2.2) * 255;
            if (!(v == v)) v = 0;
            unsigned char b = (unsigned char)(v < 0 ? 0 : v > 255 ? 255 : v);
            std::fputc(b, fp);
        }
        std::fclose(fp);
    }
}

}  // namespace ptsharp
