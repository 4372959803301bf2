// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity unpinned (see orc_math.hpp header and DESIGN.md).
// CPU restatement of PTSharp's geometry layer: Box, Ray, Material, ColorTexture, Hit/HitInfo, every IShape,
// the SDF node types, Volume, the kd-tree builder (with the ConcurrentBag ordering) and recursive traversal,
// Scene and Camera.  Written in the reference's own recursive / virtual-dispatch style on purpose: it is the
// checker for the flat, iterative GPU path, not a second copy of it.
#pragma once
#include <atomic>
#include <cstdio>
#include <memory>
#include <vector>

#include "orc_math.hpp"

namespace orc {

struct IShape;
struct Ray;

// ---------------------------------------------------------------------------------------------- Ray
// Ray.cs:8-26
struct Ray {
    Vector Origin, Direction;
    Ray() {}
    Ray(const Vector& o, const Vector& d) : Origin(o), Direction(d) {}
    Vector Position(double t) const { return Origin.Add(Direction.MulScalar(t)); }  // Ray.cs:19
};

// ---------------------------------------------------------------------------------------------- Box
// Box.cs:5-115
struct Box {
    Vector Min, Max;
    Box() {}
    Box(const Vector& mn, const Vector& mx) : Min(mn), Max(mx) {}
    Vector Size() const { return Max.Sub(Min); }                                           // :56
    Vector Anchor(const Vector& a) const { return Min.Add(Size().Mul(a)); }                // :46
    Vector Center() const { return Anchor(Vector(0.5, 0.5, 0.5)); }                        // :48
    double OuterRadius() const { return Min.Sub(Center()).Length(); }                      // :50
    Box Extend(const Box& b) const { return Box(Min.Min(b.Min), Max.Max(b.Max)); }         // :58
    // Box.cs:72-94 — slab test in double on widened floats; NaN-propagating Math.Max/Min.
    void Intersect(const Ray& r, double& tmin, double& tmax) const {
        double x1 = (Min.X() - r.Origin.X()) / r.Direction.X();
        double y1 = (Min.Y() - r.Origin.Y()) / r.Direction.Y();
        double z1 = (Min.Z() - r.Origin.Z()) / r.Direction.Z();
        double x2 = (Max.X() - r.Origin.X()) / r.Direction.X();
        double y2 = (Max.Y() - r.Origin.Y()) / r.Direction.Y();
        double z2 = (Max.Z() - r.Origin.Z()) / r.Direction.Z();
        if (x1 > x2) { double t = x1; x1 = x2; x2 = t; }
        if (y1 > y2) { double t = y1; y1 = y2; y2 = t; }
        if (z1 > z2) { double t = z1; z1 = z2; z2 = t; }
        tmin = net_max(net_max(x1, y1), z1);
        tmax = net_min(net_min(x2, y2), z2);
    }
    // Box.cs:96-114; axis 1=X 2=Y 3=Z (Axis.cs)
    void Partition(int axis, double point, bool& left, bool& right) const {
        double mn = axis == 1 ? Min.X() : axis == 2 ? Min.Y() : Min.Z();
        double mx = axis == 1 ? Max.X() : axis == 2 ? Max.Y() : Max.Z();
        left = mn <= point;
        right = mx >= point;
    }
};

// Matrix.cs:157-173
static inline Box MulBox(const Matrix& M, const Box& box) {
    Vector r(M.m[0][0], M.m[1][0], M.m[2][0]);
    Vector u(M.m[0][1], M.m[1][1], M.m[2][1]);
    Vector b(M.m[0][2], M.m[1][2], M.m[2][2]);
    Vector t(M.m[0][3], M.m[1][3], M.m[2][3]);
    Vector xa = r.MulScalar(box.Min.X()), xb = r.MulScalar(box.Max.X());
    Vector ya = u.MulScalar(box.Min.Y()), yb = u.MulScalar(box.Max.Y());
    Vector za = b.MulScalar(box.Min.Z()), zb = b.MulScalar(box.Max.Z());
    Vector xa2 = xa.Min(xb), xb2 = xa.Max(xb);
    Vector ya2 = ya.Min(yb), yb2 = ya.Max(yb);
    Vector za2 = za.Min(zb), zb2 = za.Max(zb);
    Vector mn = xa2.Add(ya2).Add(za2).Add(t);
    Vector mx = xb2.Add(yb2).Add(zb2).Add(t);
    return Box(mn, mx);
}
// Matrix.cs:153
static inline Ray MulRay(const Matrix& M, const Ray& b) { return Ray(M.MulPosition(b.Origin), M.MulDirection(b.Direction)); }

// ------------------------------------------------------------------------------------------ Texture
// Texture.cs:96-252 (ColorTexture). Data holds Colours already raised to 2.2 by the loader (:163).
struct ColorTexture {
    int Width = 0, Height = 0;
    std::vector<Colour> Data;
    // Texture.cs:188-216
    Colour BilinearSample(double u, double v) const {
        if (u == 1) u -= EPS;
        if (v == 1) v -= EPS;
        double w = (double)Width - 1;
        double h = (double)Height - 1;
        int X, Y;
        double x, y;
        Modf(u * w, X, x);
        Modf(v * h, Y, y);
        int x0 = X, y0 = Y, x1 = x0 + 1, y1 = y0 + 1;
        const Colour& c00 = Data[(size_t)y0 * Width + x0];
        const Colour& c01 = Data[(size_t)y1 * Width + x0];
        const Colour& c10 = Data[(size_t)y0 * Width + x1];
        const Colour& c11 = Data[(size_t)y1 * Width + x1];
        Colour c(0, 0, 0);
        c = c.Add(c00.MulScalar((1 - x) * (1 - y)));
        c = c.Add(c10.MulScalar(x * (1 - y)));
        c = c.Add(c01.MulScalar((1 - x) * y));
        c = c.Add(c11.MulScalar(x * y));
        return c;
    }
    // Texture.cs:218-222
    static double Fract(double x) {
        int d; double f;
        Modf(x, d, f);
        return f;
    }
    // Texture.cs:224-229
    Colour Sample(double u, double v) const {
        u = Fract(Fract(u) + 1);
        v = Fract(Fract(v) + 1);
        return BilinearSample(u, 1 - v);
    }
    // Texture.cs:231-237
    Vector NormalSample(double u, double v) const {
        u = Fract(Fract(u) + 1);
        v = Fract(Fract(v) + 1);
        Colour c = BilinearSample(u, 1 - v);
        return Vector(c.r * 2 - 1, c.g * 2 - 1, c.b * 2 - 1).Normalize();
    }
    // Texture.cs:239-251 (y may reach Height when v wraps to 0 — SURVEY A.7; clamp the row read so the
    // oracle does not fault where .NET would throw).
    Vector BumpSample(double u, double v) const {
        u = Fract(Fract(u) + 1);
        v = Fract(Fract(v) + 1);
        v = 1 - v;
        int x = (int)(u * Width);
        int y = (int)(v * Height);
        int x1 = ClampInt(x - 1, 0, Width - 1), x2 = ClampInt(x + 1, 0, Width - 1);
        int y1 = ClampInt(y - 1, 0, Height - 1), y2 = ClampInt(y + 1, 0, Height - 1);
        int yr = ClampInt(y, 0, Height - 1), xr = ClampInt(x, 0, Width - 1);
        Colour cx = Data[(size_t)yr * Width + x1].Sub(Data[(size_t)yr * Width + x2]);
        Colour cy = Data[(size_t)y1 * Width + xr].Sub(Data[(size_t)y2 * Width + xr]);
        return Vector(cx.r, cy.r, 0);
    }
};

// ----------------------------------------------------------------------------------------- Material
// Material.cs:8-139
struct Material {
    Colour Color;
    const ColorTexture* Texture = nullptr;
    const ColorTexture* NormalTexture = nullptr;
    const ColorTexture* BumpTexture = nullptr;
    const ColorTexture* GlossTexture = nullptr;
    double BumpMultiplier = 0, Emittance = 0, Index = 0, Gloss = 0, Tint = 0, Reflectivity = 0;
    bool Transparent = false;
    int id = -1;  // oracle-side bookkeeping only (authoring id), not reference state
};

// --------------------------------------------------------------------------------------------- Hit
struct HitInfo {  // Hit.cs:58-75
    const IShape* Shape = nullptr;
    Vector Position, Normal;
    Ray ray;
    Material material;
    bool Inside = false;
};

struct Hit {  // Hit.cs:4-24
    const IShape* Shape = nullptr;
    double T = HIT_INF;
    std::shared_ptr<HitInfo> info;  // null unless pre-filled by TransformedShape (Hit.cs:10)
    // oracle-side bookkeeping for parity IDs: which triangle of which mesh / which top-level shape.
    int prim = -1;
    int top = -1;  // index in Scene.Shapes of the top-level shape whose Intersect returned this hit
    Hit() {}
    Hit(const IShape* s, double t) : Shape(s), T(t) {}
    bool Ok() const { return T < HIT_INF; }  // Hit.cs:22
    HitInfo Info(const Ray& r) const;         // Hit.cs:26-55
};
static inline Hit NoHit() { return Hit(nullptr, HIT_INF); }  // Hit.cs:24

// ------------------------------------------------------------------------------------------- IShape
enum ShapeKind { K_SPHERE = 1, K_CUBE, K_PLANE, K_CYLINDER, K_TRIANGLE, K_MESH, K_TRANSFORMED, K_SDF, K_VOLUME, K_SH };

// IShape.cs:3-11
struct IShape {
    virtual ~IShape() {}
    virtual int Kind() const = 0;
    // C# value types (struct Triangle/Mesh/Cylinder/TransformedShape) are re-boxed on every `new Hit(this,..)`,
    // so `hit.Shape != light` (Sampler.cs:264) is always true for them.
    virtual bool IsClass() const = 0;
    virtual void Compile() {}
    virtual Box BoundingBox() const = 0;
    virtual Hit Intersect(const Ray& r) const = 0;
    virtual Vector UVector(const Vector& p) const = 0;
    virtual Vector NormalAt(const Vector& p) const = 0;
    virtual Material MaterialAt(const Vector& p) const = 0;
    int sceneIndex = -1;  // bookkeeping: index in Scene.Shapes, -1 if nested
};

// Material.cs:124-138
static inline Material MaterialAtShape(const IShape* shape, const Vector& point) {
    Material material = shape->MaterialAt(point);
    Vector uv = shape->UVector(point);
    if (material.Texture) material.Color = material.Texture->Sample(uv.X(), uv.Y());
    if (material.GlossTexture) {
        Colour c = material.GlossTexture->Sample(uv.X(), uv.Y());
        material.Gloss = (c.r + c.g + c.b) / 3;
    }
    return material;
}

// ------------------------------------------------------------------------------------------- Sphere
struct Sphere : IShape {  // Sphere.cs:5-82
    Vector Center;
    double Radius;
    Material Mat;
    Box box;
    Sphere(const Vector& c, double r, const Material& m) : Center(c), Radius(r), Mat(m) {
        // Sphere.cs:29-32
        Vector mn(c.X() - r, c.Y() - r, c.Z() - r), mx(c.X() + r, c.Y() + r, c.Z() + r);
        box = Box(mn, mx);
    }
    int Kind() const override { return K_SPHERE; }
    bool IsClass() const override { return true; }
    Box BoundingBox() const override { return box; }
    Hit Intersect(const Ray& r) const override {  // Sphere.cs:40-60
        Vector to = r.Origin.Sub(Center);
        double b = to.Dot(r.Direction);
        double c = to.Dot(to) - Radius * Radius;
        double d = b * b - c;
        if (d > 0) {
            d = std::sqrt(d);
            double t1 = -b - d;
            if (t1 > EPS) return Hit(this, t1);
            double t2 = -b + d;
            if (t2 > EPS) return Hit(this, t2);
        }
        return NoHit();
    }
    Vector UVector(const Vector& p0) const override {  // Sphere.cs:62-69 (incl. the (X,0,Y) typo)
        Vector p = p0.Sub(Center);
        double u = std::atan2(p.Z(), p.X());
        double v = std::atan2(p.Y(), Vector(p.X(), 0, p.Y()).Length());
        u = 1 - (u + M_PI) / (2 * M_PI);
        v = (v + M_PI / 2) / M_PI;
        return Vector(u, v, 0);
    }
    Material MaterialAt(const Vector&) const override { return Mat; }
    Vector NormalAt(const Vector& p) const override { return p.Sub(Center).Normalize(); }  // :78-81
};

// --------------------------------------------------------------------------------------------- Cube
struct Cube : IShape {  // Cube.cs:5-69
    Vector Min, Max;
    Material Mat;
    Cube(const Vector& mn, const Vector& mx, const Material& m) : Min(mn), Max(mx), Mat(m) {}
    int Kind() const override { return K_CUBE; }
    bool IsClass() const override { return true; }
    Box BoundingBox() const override { return Box(Min, Max); }
    Hit Intersect(const Ray& r) const override {  // Cube.cs:35-47
        Vector n = Min.Sub(r.Origin).Div(r.Direction);
        Vector f = Max.Sub(r.Origin).Div(r.Direction);
        Vector n2 = n.Min(f), f2 = n.Max(f);
        double t0 = net_max(net_max(n2.X(), n2.Y()), n2.Z());
        double t1 = net_min(net_min(f2.X(), f2.Y()), f2.Z());
        if (t0 > 0 && t0 < t1) return Hit(this, t0);
        return NoHit();
    }
    Vector UVector(const Vector& p0) const override {  // Cube.cs:49-53
        Vector p = p0.Sub(Min).Div(Max.Sub(Min));
        return Vector(p.X(), p.Z(), 0);
    }
    Material MaterialAt(const Vector&) const override { return Mat; }
    Vector NormalAt(const Vector& p) const override {  // Cube.cs:57-69
        if (std::fabs(p.X() - Min.X()) < EPS) return Vector(-1, 0, 0);
        if (std::fabs(p.X() - Max.X()) < EPS) return Vector(1, 0, 0);
        if (std::fabs(p.Y() - Min.Y()) < EPS) return Vector(0, -1, 0);
        if (std::fabs(p.Y() - Max.Y()) < EPS) return Vector(0, 1, 0);
        if (std::fabs(p.Z() - Min.Z()) < EPS) return Vector(0, 0, -1);
        if (std::fabs(p.Z() - Max.Z()) < EPS) return Vector(0, 0, 1);
        return Vector(0, 1, 0);
    }
};

// -------------------------------------------------------------------------------------------- Plane
struct Plane : IShape {  // Plane.cs:5-70
    Vector Point, Normal;
    Material Mat;
    Plane(const Vector& p, const Vector& n, const Material& m) : Point(p), Normal(n.Normalize()), Mat(m) {}  // :26-29
    int Kind() const override { return K_PLANE; }
    bool IsClass() const override { return true; }
    Box BoundingBox() const override { return Box(Vector(-INF, -INF, -INF), Vector(INF, INF, INF)); }  // :33-36
    Hit Intersect(const Ray& ray) const override {  // Plane.cs:38-52
        double d = Normal.Dot(ray.Direction);
        if (std::fabs(d) < EPS) return NoHit();
        Vector a = Point.Sub(ray.Origin);
        double t = a.Dot(Normal) / d;
        if (t < EPS) return NoHit();
        return Hit(this, t);
    }
    Vector UVector(const Vector&) const override { return Vector(); }
    Material MaterialAt(const Vector&) const override { return Mat; }
    Vector NormalAt(const Vector&) const override { return Normal; }
};

// ----------------------------------------------------------------------------------------- Cylinder
struct Cylinder : IShape {  // Cylinder.cs:5-166 (a C# struct)
    double Radius, Z0, Z1;
    Material Mat;
    Cylinder(double r, double z0, double z1, const Material& m) : Radius(r), Z0(z0), Z1(z1), Mat(m) {}
    int Kind() const override { return K_CYLINDER; }
    bool IsClass() const override { return false; }
    Box BoundingBox() const override { double r = Radius; return Box(Vector(-r, -r, Z0), Vector(r, r, Z1)); }  // :37-41
    Hit Intersect(const Ray& ray) const override {  // Cylinder.cs:43-111
        double r = Radius;
        Vector o = ray.Origin, d = ray.Direction;
        double tTop = (Z1 - o.Z()) / d.Z();
        double tBottom = (Z0 - o.Z()) / d.Z();
        double a = d.X() * d.X() + d.Y() * d.Y();
        double b = 2 * (o.X() * d.X() + o.Y() * d.Y());
        double c = o.X() * o.X() + o.Y() * o.Y() - r * r;
        double discriminant = b * b - 4 * a * c;
        if (tTop > EPS && tTop > 0) {
            Vector p = o.Add(d.MulScalar(tTop));  // `o + d * tTop` (Vector.cs:281-284, 242-245)
            double dist = std::sqrt(p.X() * p.X() + p.Y() * p.Y());
            if (dist <= r) return Hit(this, tTop);
        }
        if (tBottom > EPS && tBottom > 0) {
            Vector p = o.Add(d.MulScalar(tBottom));
            double dist = std::sqrt(p.X() * p.X() + p.Y() * p.Y());
            if (dist <= r) return Hit(this, tBottom);
        }
        if (discriminant >= 0) {
            double sq = std::sqrt(discriminant);
            double t1 = (-b + sq) / (2 * a);
            double t2 = (-b - sq) / (2 * a);
            double tLateral = std::nan("");
            if (t1 > EPS && t1 > 0) tLateral = t1;
            else if (t2 > EPS && t2 > 0) tLateral = t2;
            if (!(tLateral != tLateral)) {
                Vector p = o.Add(d.MulScalar(tLateral));
                double z = p.Z();
                if (z >= Z0 && z <= Z1) return Hit(this, tLateral);
            }
        }
        return NoHit();
    }
    Vector UVector(const Vector& p) const override { return Vector(-p.Y(), p.X(), 0).Normalize(); }  // :114-118
    Material MaterialAt(const Vector&) const override { return Mat; }
    Vector NormalAt(const Vector& p) const override {  // Cylinder.cs:122-163
        double epsilon = 0.0001;
        if (std::fabs(p.Z() - Z0) > epsilon && std::fabs(p.Z() - Z1) > epsilon) {
            Vector center(0, 0, (Z0 + Z1) / 2);
            Vector toPoint = p.Sub(center);
            Vector normal = toPoint.Normalize();
            if (normal.Dot(p.Sub(Vector(0, 0, Z0))) < 0) normal = normal.Negate();
            return normal;
        }
        if (std::fabs(p.Z() - Z0) < epsilon) return Vector(0, 0, -1);
        if (std::fabs(p.Z() - Z1) < epsilon) return Vector(0, 0, 1);
        return Vector(0, 0, 0);
    }
};

// ----------------------------------------------------------------------------------------- Triangle
struct Triangle : IShape {  // Triangle.cs:8-243 (a C# struct)
    Material Mat;
    Vector V1, V2, V3, N1, N2, N3, T1, T2, T3;
    int index = -1;  // bookkeeping: position in Mesh.Triangles
    int Kind() const override { return K_TRIANGLE; }
    bool IsClass() const override { return false; }
    Box BoundingBox() const override {  // Triangle.cs:80-85
        return Box(V1.Min(V2).Min(V3), V1.Max(V2).Max(V3));
    }
    Hit Intersect(const Ray& r) const override {  // Triangle.cs:95-124
        Vector e1 = V2.Sub(V1);
        Vector e2 = V3.Sub(V1);
        Vector h = r.Direction.Cross(e2);
        double det = e1.Dot(h);
        if (det > -EPS && det < EPS) return NoHit();
        double invDet = 1.0 / det;
        Vector s = r.Origin.Sub(V1);
        double u = s.Dot(h) * invDet;
        if (u < 0 || u > 1) return NoHit();
        Vector q = s.Cross(e1);
        double v = r.Direction.Dot(q) * invDet;
        if (v < 0 || (u + v) > 1) return NoHit();
        double t = e2.Dot(q) * invDet;
        if (t < EPS) return NoHit();
        Hit hit(this, t);
        hit.prim = index;
        return hit;
    }
    void Barycentric(const Vector& p, double& u, double& v, double& w) const {  // Triangle.cs:208-223
        Vector v0 = V2.Sub(V1), v1 = V3.Sub(V1), v2 = p.Sub(V1);
        double d00 = v0.Dot(v0), d01 = v0.Dot(v1), d11 = v1.Dot(v1), d20 = v2.Dot(v0), d21 = v2.Dot(v1);
        double d = d00 * d11 - d01 * d01;
        v = (d11 * d20 - d01 * d21) / d;
        w = (d00 * d21 - d01 * d20) / d;
        u = 1 - v - w;
    }
    Vector UVector(const Vector& p) const override {  // Triangle.cs:128-136
        double u, v, w;
        Barycentric(p, u, v, w);
        Vector n;
        n = n.Add(T1.MulScalar(u));
        n = n.Add(T2.MulScalar(v));
        n = n.Add(T3.MulScalar(w));
        return Vector(n.X(), n.Y(), 0);
    }
    Material MaterialAt(const Vector&) const override { return Mat; }
    Vector NormalAt(const Vector& p) const override {  // Triangle.cs:142-189
        double u, v, w;
        Barycentric(p, u, v, w);
        Vector n = N1.MulScalar(u).Add(N2.MulScalar(v)).Add(N3.MulScalar(w));
        if (Mat.NormalTexture) {
            Vector b = T1.MulScalar(u).Add(T2.MulScalar(v)).Add(T3.MulScalar(w));
            Vector ns = Mat.NormalTexture->NormalSample(b.X(), b.Y());
            if (!ns.Equals(Vector())) {
                Vector dv1 = V2.Sub(V1), dv2 = V3.Sub(V1), dt1 = T2.Sub(T1), dt2 = T3.Sub(T1);
                Vector T = dv1.MulScalar(dt2.Y()).Sub(dv2.MulScalar(dt1.Y())).Normalize();
                Vector B = dv2.MulScalar(dt1.X()).Sub(dv1.MulScalar(dt2.X())).Normalize();
                Vector N = T.Cross(B);
                Matrix M;
                M.m[0][0] = T.X(); M.m[0][1] = B.X(); M.m[0][2] = N.X();
                M.m[1][0] = T.Y(); M.m[1][1] = B.Y(); M.m[1][2] = N.Y();
                M.m[2][0] = T.Z(); M.m[2][1] = B.Z(); M.m[2][2] = N.Z();
                M.m[3][3] = 1;
                n = M.MulDirection(ns);
            }
        }
        if (Mat.BumpTexture) {
            Vector b = T1.MulScalar(u).Add(T2.MulScalar(v)).Add(T3.MulScalar(w));
            Vector bump = Mat.BumpTexture->BumpSample(b.X(), b.Y());
            if (!bump.Equals(Vector())) {
                Vector dv1 = V2.Sub(V1), dv2 = V3.Sub(V1), dt1 = T2.Sub(T1), dt2 = T3.Sub(T1);
                Vector tangent = dv1.MulScalar(dt2.Y()).Sub(dv2.MulScalar(dt1.Y())).Normalize();
                Vector bitangent = dv2.MulScalar(dt1.X()).Sub(dv1.MulScalar(dt2.X())).Normalize();
                n = n.Add(tangent.MulScalar(bump.X() * Mat.BumpMultiplier));
                n = n.Add(bitangent.MulScalar(bump.Y() * Mat.BumpMultiplier));
            }
        }
        return n.Normalize();
    }
    Vector Normal() const { return V2.Sub(V1).Cross(V3.Sub(V1)).Normalize(); }  // Triangle.cs:198-203
    void FixNormals() {                                                          // Triangle.cs:224-237
        Vector n = Normal(), zero;
        if (N1.Equals(zero)) N1 = n;
        if (N2.Equals(zero)) N2 = n;
        if (N3.Equals(zero)) N3 = n;
    }
};

// --------------------------------------------------------------------------------------------- Tree
// Traversal-cost probe (tree-quality tests only; single-threaded; off by default).
struct TraversalProbe { bool on = false; long long nodes = 0, prims = 0; };
inline TraversalProbe& probe() { static TraversalProbe p; return p; }

// Tree.cs:8-267
struct Node {
    int Axis = 0;  // Axis.cs: 0 None, 1 X, 2 Y, 3 Z
    double Point = 0;
    std::vector<const IShape*> Shapes;
    std::unique_ptr<Node> Left, Right;

    Hit IntersectShapes(const Ray& r) const {  // Tree.cs:115-128
        Hit hit = NoHit();
        if (probe().on) { probe().nodes++; probe().prims += (long long)Shapes.size(); }
        for (const IShape* shape : Shapes) {
            Hit h = shape->Intersect(r);
            if (shape->sceneIndex >= 0) h.top = shape->sceneIndex;  // bookkeeping only
            if (h.T < hit.T) hit = h;
        }
        return hit;
    }
    Hit Intersect(const Ray& r, double tmin, double tmax) const {  // Tree.cs:67-113
        double tsplit;
        bool leftFirst;
        if (Axis != 0 && probe().on) probe().nodes++;
        switch (Axis) {
            case 0: return IntersectShapes(r);
            case 1:
                tsplit = (Point - r.Origin.X()) / r.Direction.X();
                leftFirst = (r.Origin.X() < Point) || (r.Origin.X() == Point && r.Direction.X() <= 0);
                break;
            case 2:
                tsplit = (Point - r.Origin.Y()) / r.Direction.Y();
                leftFirst = (r.Origin.Y() < Point) || (r.Origin.Y() == Point && r.Direction.Y() <= 0);
                break;
            default:
                tsplit = (Point - r.Origin.Z()) / r.Direction.Z();
                leftFirst = (r.Origin.Z() < Point) || (r.Origin.Z() == Point && r.Direction.Z() <= 0);
                break;
        }
        const Node* first = leftFirst ? Left.get() : Right.get();
        const Node* second = leftFirst ? Right.get() : Left.get();
        if (tsplit > tmax || tsplit <= 0) return first->Intersect(r, tmin, tmax);
        else if (tsplit < tmin) return second->Intersect(r, tmin, tmax);
        else {
            Hit h1 = first->Intersect(r, tmin, tsplit);
            if (h1.T <= tsplit) return h1;
            Hit h2 = second->Intersect(r, tsplit, net_min(tmax, h1.T));
            return h1.T <= h2.T ? h1 : h2;
        }
    }
    int PartitionScore(int axis, double point) const {  // Tree.cs:150-175
        int left = 0, right = 0;
        for (const IShape* s : Shapes) {
            bool l, r;
            s->BoundingBox().Partition(axis, point, l, r);
            if (l) left++;
            if (r) right++;
        }
        return left >= right ? left : right;
    }
    // Tree.cs:130-148 applied to a ConcurrentBag<double> filled single-threaded with
    // min0,max0,min1,max1,... (Tree.cs:212-220).  A single-thread ConcurrentBag enumerates LIFO
    // (SURVEY §8c U1), so ElementAt(k) is insertion index 2N-1-k.
    static double MedianOfBag(const std::vector<double>& inserted) {
        size_t count = inserted.size();
        if (count == 0) return 0;
        auto elementAt = [&](size_t k) { return inserted[count - 1 - k]; };
        if (count % 2 == 1) return elementAt(count / 2);
        double a = elementAt(count / 2 - 1);
        double b = elementAt(count / 2);
        return (a + b) / 2;
    }
    void Split(int depth) {  // Tree.cs:201-265
        if (Shapes.size() < 8) return;
        std::vector<double> xs, ys, zs;
        xs.reserve(Shapes.size() * 2); ys.reserve(Shapes.size() * 2); zs.reserve(Shapes.size() * 2);
        for (const IShape* s : Shapes) {
            Box box = s->BoundingBox();
            xs.push_back(box.Min.X()); xs.push_back(box.Max.X());
            ys.push_back(box.Min.Y()); ys.push_back(box.Max.Y());
            zs.push_back(box.Min.Z()); zs.push_back(box.Max.Z());
        }
        double mx = MedianOfBag(xs), my = MedianOfBag(ys), mz = MedianOfBag(zs);
        int best = (int)((double)Shapes.size() * 0.85);
        int bestAxis = 0;
        double bestPoint = 0.0;
        int sx = PartitionScore(1, mx);
        if (sx < best) { best = sx; bestAxis = 1; bestPoint = mx; }
        int sy = PartitionScore(2, my);
        if (sy < best) { best = sy; bestAxis = 2; bestPoint = my; }
        int sz = PartitionScore(3, mz);
        if (sz < best) { best = sz; bestAxis = 3; bestPoint = mz; }
        if (bestAxis == 0) return;
        // Partition (Tree.cs:177-199): ConcurrentBag.ToArray() of a single-thread bag = reverse insertion order.
        std::vector<const IShape*> l, r;
        for (const IShape* s : Shapes) {
            bool bl, br;
            s->BoundingBox().Partition(bestAxis, bestPoint, bl, br);
            if (bl) l.push_back(s);
            if (br) r.push_back(s);
        }
        Axis = bestAxis;
        Point = bestPoint;
        Left.reset(new Node());
        Right.reset(new Node());
        Left->Shapes.assign(l.rbegin(), l.rend());
        Right->Shapes.assign(r.rbegin(), r.rend());
        Left->Split(depth + 1);
        Right->Split(depth + 1);
        Shapes.clear();
        Shapes.shrink_to_fit();
    }
};

static inline Box BoxForShapes(const std::vector<const IShape*>& shapes) {  // Box.cs:20-32
    if (shapes.empty()) return Box();
    Box box = shapes[0]->BoundingBox();
    for (const IShape* s : shapes) box = box.Extend(s->BoundingBox());
    return box;
}

struct Tree {
    Box box;
    std::unique_ptr<Node> Root;
    static Tree* NewTree(const std::vector<const IShape*>& shapes) {  // Tree.cs:22-29
        Tree* t = new Tree();
        t->box = BoxForShapes(shapes);
        t->Root.reset(new Node());
        t->Root->Shapes = shapes;
        t->Root->Split(0);
        return t;
    }
    Hit Intersect(const Ray& r) const {  // Tree.cs:31-42
        double tmin, tmax;
        box.Intersect(r, tmin, tmax);
        if (tmax < tmin || tmax <= 0) return NoHit();
        return Root->Intersect(r, tmin, tmax);
    }
};

// --------------------------------------------------------------------------------------------- Mesh
struct Mesh : IShape {  // Mesh.cs:6-140 (a C# struct)
    std::vector<Triangle> Triangles;
    std::unique_ptr<Tree> tree;
    mutable bool haveBox = false;
    mutable Box box;
    int Kind() const override { return K_MESH; }
    bool IsClass() const override { return false; }
    void Compile() override {  // Mesh.cs:45-57
        if (!tree) {
            std::vector<const IShape*> shapes(Triangles.size());
            for (size_t i = 0; i < Triangles.size(); i++) shapes[i] = &Triangles[i];
            tree.reset(Tree::NewTree(shapes));
        }
    }
    Box BoundingBox() const override {  // Mesh.cs:88-103
        if (!haveBox) {
            Vector mn = Triangles[0].V1, mx = Triangles[0].V1;
            for (const Triangle& t : Triangles) {
                mn = mn.Min(t.V1).Min(t.V2).Min(t.V3);
                mx = mx.Max(t.V1).Max(t.V2).Max(t.V3);
            }
            box = Box(mn, mx);
            haveBox = true;
        }
        return box;
    }
    Hit Intersect(const Ray& r) const override { return tree->Intersect(r); }  // Mesh.cs:122-125
    Vector UVector(const Vector&) const override { return Vector(); }           // :127-130
    Material MaterialAt(const Vector&) const override { return Material(); }    // :132-135
    Vector NormalAt(const Vector&) const override { return Vector(); }          // :137-140
};

// ------------------------------------------------------------------------------- SphericalHarmonic
// SH.cs:7-338.  The marching-cubes mesh (SH.cs:20, MC.cs) is handed in from outside: the oracle restates what happens to a ray
// (Intersect / NormalAt / MaterialAt), not the authoring-time mesh generation, which tests/test_mc.py checks against its own
// numpy restatement of MC.cs.
struct SphericalHarmonic : IShape {
    int L, M;
    Material PositiveMaterial, NegativeMaterial;
    Mesh mesh;
    int Kind() const override { return K_SH; }
    bool IsClass() const override { return true; }
    void Compile() override { mesh.Compile(); }                                                 // SH.cs:24-27
    Box BoundingBox() const override { return Box(Vector(-1, -1, -1), Vector(1, 1, 1)); }         // SH.cs:29-33
    Hit Intersect(const Ray& r) const override {                                                  // SH.cs:47-55
        Hit hit = mesh.Intersect(r);
        if (!hit.Ok()) return NoHit();
        return Hit(this, hit.T);
    }
    Vector UVector(const Vector&) const override { return Vector(); }
    Material MaterialAt(const Vector& p) const override { return EvaluateHarmonic(p) < 0 ? NegativeMaterial : PositiveMaterial; }  // SH.cs:62-72
    Vector NormalAt(const Vector& p) const override {                                             // SH.cs:74-86
        const double e = 0.0001;
        double x = p.X(), y = p.Y(), z = p.Z();
        Vector n(Evaluate(Vector(x - e, y, z)) - Evaluate(Vector(x + e, y, z)),
                 Evaluate(Vector(x, y - e, z)) - Evaluate(Vector(x, y + e, z)),
                 Evaluate(Vector(x, y, z - e)) - Evaluate(Vector(x, y, z + e)));
        return n.Normalize();
    }
    double EvaluateHarmonic(const Vector& p) const { return Harmonic(p.Normalize()); }          // SH.cs:88-91
    double Evaluate(const Vector& p) const { return (double)p.Length() - std::fabs(Harmonic(p.Normalize())); }  // SH.cs:93-101
    // shFunc (SH.cs:213-338) and the functions it selects (SH.cs:103-211): float literals widened, products left to right
    double Harmonic(const Vector& d) const {
        const double X = d.X(), Y = d.Y(), Z = d.Z();
        if (L == 0 && M == 0) return 0.282095f;
        if (L == 1 && M == -1) return -0.488603f * Y;
        if (L == 1 && M == 0) return 0.488603f * Z;
        if (L == 1 && M == 1) return -0.488603f * X;
        if (L == 2 && M == -2) return 1.092548f * X * Y;
        if (L == 2 && M == -1) return -1.092548f * Y * Z;
        if (L == 2 && M == 0) return 0.315392f * (-X * X - Y * Y + 2.0f * Z * Z);
        if (L == 2 && M == 1) return -1.092548f * X * Z;
        if (L == 2 && M == 2) return 0.546274f * (X * X - Y * Y);
        if (L == 3 && M == -3) return -0.590044f * Y * (3.0f * X * X - Y * Y);
        if (L == 3 && M == -2) return 2.890611f * X * Y * Z;
        if (L == 3 && M == -1) return -0.457046f * Y * (4.0f * Z * Z - X * X - Y * Y);
        if (L == 3 && M == 0) return 0.373176f * Z * (2.0f * Z * Z - 3.0f * X * X - 3.0f * Y * Y);
        if (L == 3 && M == 1) return -0.457046f * X * (4.0f * Z * Z - X * X - Y * Y);
        if (L == 3 && M == 2) return 1.445306f * Z * (X * X - Y * Y);
        if (L == 3 && M == 3) return -0.590044f * X * (X * X - 3.0f * Y * Y);
        if (L == 4 && M == -4) return 2.503343f * X * Y * (X * X - Y * Y);
        if (L == 4 && M == -3) return -1.770131f * Y * Z * (3.0f * X * X - Y * Y);
        if (L == 4 && M == -2) return 0.946175f * X * Y * (7.0f * Z * Z - 1.0f);
        if (L == 4 && M == -1) return -0.669047f * Y * Z * (7.0f * Z * Z - 3.0f);
        if (L == 4 && M == 0) { double z2 = Z * Z; return 0.105786f * (35.0f * z2 * z2 - 30.0f * z2 + 3.0f); }
        if (L == 4 && M == 1) return -0.669047f * X * Z * (7.0f * Z * Z - 3.0f);
        if (L == 4 && M == 2) return 0.473087f * (X * X - Y * Y) * (7.0f * Z * Z - 1.0f);
        if (L == 4 && M == 3) return -1.770131f * X * Z * (X * X - 3.0f * Y * Y);
        if (L == 4 && M == 4) { double x2 = X * X, y2 = Y * Y; return 0.625836f * (x2 * (x2 - 3.0f * y2) - y2 * (3.0f * x2 - y2)); }
        return std::nan("");
    }
};

// --------------------------------------------------------------------------------- TransformedShape
struct TransformedShape : IShape {  // TransformedShape.cs:9-93 (a C# struct)
    IShape* Shape;
    Matrix M;
    TransformedShape(IShape* s, const Matrix& m) : Shape(s), M(m) {}
    int Kind() const override { return K_TRANSFORMED; }
    bool IsClass() const override { return false; }
    void Compile() override { Shape->Compile(); }
    Box BoundingBox() const override { return MulBox(M, Shape->BoundingBox()); }  // :36-39
    Hit Intersect(const Ray& r) const override {                                  // :43-72
        Matrix inv = M.Inverse();  // recomputed per call in the reference (:45)
        Ray shapeRay = MulRay(inv, r);
        Hit hit = Shape->Intersect(shapeRay);
        if (!hit.Ok()) return hit;
        const IShape* shape = hit.Shape;
        Vector shapePosition = shapeRay.Position(hit.T);
        Vector shapeNormal = shape->NormalAt(shapePosition);
        Vector position = M.MulPosition(shapePosition);
        Vector normal = M.Inverse().Transpose().MulDirection(shapeNormal);
        Material material = MaterialAtShape(shape, shapePosition);
        bool inside = false;
        if (shapeNormal.Dot(shapeRay.Direction) > 0) {
            normal = normal.Negate();
            inside = true;
        }
        auto info = std::make_shared<HitInfo>();
        info->Shape = shape;
        info->Position = position;
        info->Normal = normal;
        info->ray = Ray(position, normal);
        info->material = material;
        info->Inside = inside;
        hit.T = position.Sub(r.Origin).Length();
        hit.info = info;
        return hit;
    }
    Vector UVector(const Vector& p) const override { return Shape->UVector(p); }
    Vector NormalAt(const Vector& p) const override { return Shape->NormalAt(p); }
    Material MaterialAt(const Vector& p) const override { return Shape->MaterialAt(p); }
};

// ---------------------------------------------------------------------------------------------- SDF
struct SDF {  // SDF.cs:6-10
    virtual ~SDF() {}
    virtual double Evaluate(const Vector& p) const = 0;
    virtual Box BoundingBox() const = 0;
};
struct SphereSDF : SDF {  // SDF.cs:112-139
    double Radius, Exponent;
    SphereSDF(double r) : Radius(r), Exponent(2) {}
    double Evaluate(const Vector& p) const override { return p.LengthN(Exponent) - Radius; }
    Box BoundingBox() const override { double r = Radius; return Box(Vector(-r, -r, -r), Vector(r, r, r)); }
};
struct CubeSDF : SDF {  // SDF.cs:141-195
    Vector Size;
    CubeSDF(const Vector& s) : Size(s) {}
    double Evaluate(const Vector& p) const override {
        double x = p.X(), y = p.Y(), z = p.Z();
        if (x < 0) x = -x;
        if (y < 0) y = -y;
        if (z < 0) z = -z;
        x -= Size.X() / 2; y -= Size.Y() / 2; z -= Size.Z() / 2;
        double a = x;
        if (y > a) a = y;
        if (z > a) a = z;
        if (a > 0) a = 0;
        if (x < 0) x = 0;
        if (y < 0) y = 0;
        if (z < 0) z = 0;
        double b = std::sqrt(x * x + y * y + z * z);
        return a + b;
    }
    Box BoundingBox() const override {
        double x = Size.X() / 2, y = Size.Y() / 2, z = Size.Z() / 2;
        return Box(Vector(-x, -y, -z), Vector(x, y, z));
    }
};
struct CylinderSDF : SDF {  // SDF.cs:197-252
    double Radius, Height;
    CylinderSDF(double r, double h) : Radius(r), Height(h) {}
    Box BoundingBox() const override { double r = Radius, h = Height / 2; return Box(Vector(-r, -h, -r), Vector(r, h, r)); }
    double Evaluate(const Vector& p) const override {
        double x = std::sqrt(p.X() * p.X() + p.Z() * p.Z());
        double y = p.Y();
        if (x < 0) x = -x;
        if (y < 0) y = -y;
        x -= Radius;
        y -= Height / 2;
        double a = x;
        if (y > a) a = y;
        if (a > 0) a = 0;
        if (x < 0) x = 0;
        if (y < 0) y = 0;
        double b = std::sqrt(x * x + y * y);
        return a + b;
    }
};
struct CapsuleSDF : SDF {  // SDF.cs:254-285
    Vector A, B;
    double Radius, Exponent;
    CapsuleSDF(const Vector& a, const Vector& b, double r) : A(a), B(b), Radius(r), Exponent(2) {}
    double Evaluate(const Vector& p) const override {
        Vector pa = p.Sub(A), ba = B.Sub(A);
        double h = net_max(0, net_min(1, pa.Dot(ba) / ba.Dot(ba)));
        return pa.Sub(ba.MulScalar(h)).LengthN(Exponent) - Radius;
    }
    Box BoundingBox() const override {
        Vector a = A.Min(B), b = A.Max(B);
        return Box(a.SubScalar(Radius), b.AddScalar(Radius));
    }
};
struct TorusSDF : SDF {  // SDF.cs:287-319 (bbox quirk: Min.Z = Max.Z = +a)
    double MajorRadius, MinRadius, MajorExponent, MinorExponent;
    TorusSDF(double major, double minor) : MajorRadius(major), MinRadius(minor), MajorExponent(2), MinorExponent(2) {}
    double Evaluate(const Vector& p) const override {
        Vector q(Vector(p.X(), p.Y(), 0).LengthN(MajorExponent) - MajorRadius, p.Z(), 0);
        return q.LengthN(MinorExponent) - MinRadius;
    }
    Box BoundingBox() const override {
        double a = MinRadius, b = MinRadius + MajorRadius;
        return Box(Vector(-b, -b, a), Vector(b, b, a));
    }
};
struct TransformSDF : SDF {  // SDF.cs:321-355
    const SDF* Inner;
    Matrix M, Inv;
    TransformSDF(const SDF* s, const Matrix& m) : Inner(s), M(m), Inv(m.Inverse()) {}
    double Evaluate(const Vector& p) const override { return Inner->Evaluate(Inv.MulPosition(p)); }
    Box BoundingBox() const override { return MulBox(M, Inner->BoundingBox()); }
};
struct ScaleSDF : SDF {  // SDF.cs:357-384
    const SDF* Inner;
    double Factor;
    ScaleSDF(const SDF* s, double f) : Inner(s), Factor(f) {}
    double Evaluate(const Vector& p) const override { return Inner->Evaluate(p.DivScalar(Factor)) * Factor; }
    Box BoundingBox() const override { double f = Factor; return MulBox(Matrix::Scale(Vector(f, f, f)), Inner->BoundingBox()); }
};
struct UnionSDF : SDF {  // SDF.cs:386-437
    std::vector<const SDF*> Items;
    double Evaluate(const Vector& p) const override {
        double result = 0; int i = 0;
        for (const SDF* it : Items) { double d = it->Evaluate(p); if (i == 0 || d < result) result = d; i++; }
        return result;
    }
    Box BoundingBox() const override {
        Box result; int i = 0;
        for (const SDF* it : Items) { Box b = it->BoundingBox(); result = (i == 0) ? b : result.Extend(b); i++; }
        return result;
    }
};
struct DifferenceSDF : SDF {  // SDF.cs:439-482
    std::vector<const SDF*> Items;
    double Evaluate(const Vector& p) const override {
        double result = 0; int i = 0;
        for (const SDF* it : Items) {
            double d = it->Evaluate(p);
            if (i == 0) result = d;
            else if (-d > result) result = -d;
            i++;
        }
        return result;
    }
    Box BoundingBox() const override { return Items[0]->BoundingBox(); }
};
struct IntersectionSDF : SDF {  // SDF.cs:484-533
    std::vector<const SDF*> Items;
    double Evaluate(const Vector& p) const override {
        double result = 0; int i = 0;
        for (const SDF* it : Items) { double d = it->Evaluate(p); if (i == 0 || d > result) result = d; i++; }
        return result;
    }
    Box BoundingBox() const override {
        Box result; int i = 0;
        for (const SDF* it : Items) { Box b = it->BoundingBox(); result = (i == 0) ? b : result.Extend(b); i++; }
        return result;
    }
};
struct RepeatSDF : SDF {  // SDF.cs:535-559 (bbox quirk: a point at the origin)
    const SDF* Inner;
    Vector Step;
    RepeatSDF(const SDF* s, const Vector& st) : Inner(s), Step(st) {}
    double Evaluate(const Vector& p) const override { return Inner->Evaluate(p.Mod(Step).Sub(Step.DivScalar(2))); }
    Box BoundingBox() const override { return Box(); }
};

struct SDFShape : IShape {  // SDF.cs:12-110
    const SDF* Sdf;
    Material Mat;
    SDFShape(const SDF* s, const Material& m) : Sdf(s), Mat(m) {}
    int Kind() const override { return K_SDF; }
    bool IsClass() const override { return true; }
    Box BoundingBox() const override { return Sdf->BoundingBox(); }
    double Evaluate(const Vector& p) const { return Sdf->Evaluate(p); }
    Hit Intersect(const Ray& ray) const override {  // SDF.cs:32-76
        double epsilon = (double)0.00001f;
        double start = (double)0.0001f;
        double jumpSize = (double)0.001f;
        Box box = BoundingBox();
        double t1, t2;
        box.Intersect(ray, t1, t2);
        if (t2 < t1 || t2 < 0) return NoHit();
        double t = net_max(start, t1);
        bool jump = true;
        for (int i = 0; i < 1000; i++) {
            double d = Evaluate(ray.Position(t));
            if (jump && d < 0) {
                t -= jumpSize;
                jump = false;
                continue;
            }
            if (d < epsilon) return Hit(this, t);
            if (jump && d < jumpSize) d = jumpSize;
            t += d;
            if (t > t2) return NoHit();
        }
        return NoHit();
    }
    Vector UVector(const Vector&) const override { return Vector(); }
    Vector NormalAt(const Vector& p) const override {  // SDF.cs:83-92
        double e = 0.0001;
        double x = p.X(), y = p.Y(), z = p.Z();
        Vector n(Evaluate(Vector(x - e, y, z)) - Evaluate(Vector(x + e, y, z)),
                 Evaluate(Vector(x, y - e, z)) - Evaluate(Vector(x, y + e, z)),
                 Evaluate(Vector(x, y, z - e)) - Evaluate(Vector(x, y, z + e)));
        return n.Normalize();
    }
    Material MaterialAt(const Vector&) const override { return Mat; }
};

// ------------------------------------------------------------------------------------------- Volume
struct VolumeWindow { double Lo, Hi; Material Mat; };  // Volume.cs:8-20
struct Volume : IShape {                               // Volume.cs:6-197
    int W, H, D;
    double ZScale;
    std::vector<double> Data;
    std::vector<VolumeWindow> Windows;
    Box box;
    int Kind() const override { return K_VOLUME; }
    bool IsClass() const override { return true; }
    double Get(int x, int y, int z) const {  // :40-46
        if (x < 0 || y < 0 || z < 0 || x >= W || y >= H || z >= D) return 0;
        return Data[(size_t)x + (size_t)y * W + (size_t)z * W * H];
    }
    double Sample(double x, double y, double z) const {  // :73-104 (index bugs kept)
        z /= ZScale;
        x = ((x + 1) / 2) * (double)W;
        y = ((z + 1) / 2) * (double)H;
        z = ((z + 2) / 2) * (double)D;
        int x0 = (int)std::floor(x), y0 = (int)std::floor(y), z0 = (int)std::floor(z);
        int x1 = x0 + 1, y1 = y0 + 1, z1 = z0 + 1;
        double v000 = Get(x0, y0, z0), v001 = Get(x0, y0, z1), v010 = Get(x0, y1, z0), v011 = Get(x0, y1, z1);
        double v100 = Get(x1, y0, z0), v101 = Get(x1, y0, z1), v110 = Get(x1, y1, z0), v111 = Get(x1, y1, z1);
        x -= (double)x0; y -= (double)y0; z -= (double)z0;
        double c00 = v000 * (1 - x) + v100 * x;
        double c01 = v001 * (1 - x) + v101 * x;
        double c10 = v010 * (1 - x) + v110 * x;
        double c11 = v011 * (1 - x) + v111 * x;
        double c0 = c00 * (1 - y) + c10 * y;
        double c1 = c01 * (1 - y) + c11 * y;
        return c0 * (1 - z) + c1 * z;
    }
    Box BoundingBox() const override { return box; }
    int Sign(const Vector& a) const {  // :113-131 (`i` never incremented)
        double s = Sample(a.X(), a.Y(), a.Z());
        int i = 0;
        for (const VolumeWindow& w : Windows) {
            if (s < w.Lo) return i + 1;
            if (s > w.Hi) continue;
            return 0;
        }
        return (int)Windows.size() + 1;
    }
    Vector UVector(const Vector&) const override { return Vector(); }
    Vector NormalAt(const Vector& p) const override {  // :138-145
        double eps = (double)0.001f;
        Vector n(Sample(p.X() - eps, p.Y(), p.Z()) - Sample(p.X() + eps, p.Y(), p.Z()),
                 Sample(p.X(), p.Y() - eps, p.Z()) - Sample(p.X(), p.Y() + eps, p.Z()),
                 Sample(p.X(), p.Y(), p.Z() - eps) - Sample(p.X(), p.Y(), p.Z() + eps));
        return n.Normalize();
    }
    Material MaterialAt(const Vector& p) const override {  // :147-167
        double be = (double)1e9f;
        Material bm;
        double s = Sample(p.X(), p.Y(), p.Z());
        for (const VolumeWindow& w : Windows) {
            if (s >= w.Lo && s <= w.Hi) return w.Mat;
            double e = net_min(std::fabs(s - w.Lo), std::fabs(s - w.Hi));
            if (e < be) { be = e; bm = w.Mat; }
        }
        return bm;
    }
    Hit Intersect(const Ray& ray) const override {  // :169-197
        double tmin, tmax;
        box.Intersect(ray, tmin, tmax);
        double step = (double)(1.0f / 512.0f);
        double start = net_max(step, tmin);
        int sign = -1;
        for (double t = start; t <= tmax; t += step) {
            Vector p = ray.Position(t);
            int s = Sign(p);
            if (s == 0 || (sign >= 0 && s != sign)) {
                t -= step;
                step /= 64;
                t += step;
                for (int i = 0; i < 64; i++) {
                    if (Sign(ray.Position(t)) == 0) return Hit(this, t - step);
                    t += step;
                }
            }
            sign = s;
        }
        return NoHit();
    }
};

// ---------------------------------------------------------------------------------------- Hit.Info
inline HitInfo Hit::Info(const Ray& r) const {  // Hit.cs:26-55
    if (info) return *info;
    const IShape* shape = Shape;
    Vector position = r.Position(T);
    Vector normal = shape->NormalAt(position);
    Material material = MaterialAtShape(shape, position);
    bool inside = false;
    if (normal.Dot(r.Direction) > 0) {
        normal = normal.Negate();
        inside = true;
        int k = shape->Kind();
        if (k == K_VOLUME || k == K_SDF || k == K_SH) inside = false;  // Hit.cs:41-47
    }
    HitInfo hi;
    hi.Shape = shape;
    hi.Position = position;
    hi.Normal = normal;
    hi.ray = Ray(position, normal);
    hi.material = material;
    hi.Inside = inside;
    return hi;
}

// -------------------------------------------------------------------------------------------- Scene
struct Scene {  // Scene.cs:9-80
    Colour Color;
    const ColorTexture* Texture = nullptr;
    double TextureAngle = 0;
    std::unique_ptr<Tree> tree;
    std::atomic<long long> rays{0};
    std::vector<IShape*> Shapes;
    std::vector<IShape*> Lights;
    void Add(IShape* p) {  // Scene.cs:29-38
        p->sceneIndex = (int)Shapes.size();
        Shapes.push_back(p);
        if (p->MaterialAt(Vector()).Emittance > 0) Lights.push_back(p);
    }
    void Compile() {  // Scene.cs:48-68
        for (IShape* s : Shapes) s->Compile();
        if (!tree) {
            std::vector<const IShape*> v(Shapes.begin(), Shapes.end());
            tree.reset(Tree::NewTree(v));
        }
    }
    Hit Intersect(const Ray& r) {  // Scene.cs:75-79
        rays.fetch_add(1, std::memory_order_relaxed);
        return tree->Intersect(r);
    }
};

// ------------------------------------------------------------------------------------------- Camera
struct Camera {  // Camera.cs:9-120
    Vector p, u, v, w;
    double m = 0, focalDistance = 0, apertureRadius = 0;
    static Camera LookAt(const Vector& eye, const Vector& center, const Vector& up, double fovy) {  // :23-35
        Camera c;
        c.p = eye;
        c.w = center.Sub(eye).Normalize();
        c.u = up.Cross(c.w).Normalize();
        c.v = c.w.Cross(c.u).Normalize();
        c.m = 1 / std::tan(fovy * M_PI / 360);
        return c;
    }
    void SetFocus(const Vector& focalPoint, double aperture) {  // :39-43
        focalDistance = focalPoint.Sub(p).Length();
        apertureRadius = aperture;
    }
    // Camera.cs:98-119; r1,r2 are the two Random.Shared draws (only consumed when apertureRadius > 0).
    template <class Rng>
    Ray CastRay(int x, int y, int W, int H, double uu, double vv, Rng& rng) const {
        double aspect = W / (double)H;
        double px = ((x + uu - 0.5) / (W - 1.0)) * 2 - 1;
        double py = ((y + vv - 0.5) / (H - 1.0)) * 2 - 1;
        Vector d = Vector().Add(u.MulScalar(-px * aspect)).Add(v.MulScalar(-py)).Add(w.MulScalar(m)).Normalize();
        Vector P = p;
        if (apertureRadius > 0) {
            Vector focalPoint = p.Add(d.MulScalar(focalDistance));
            double angle = rng.NextDouble() * 2 * M_PI;
            double radius = rng.NextDouble() * apertureRadius;
            P = P.Add(u.MulScalar(std::cos(angle) * radius));
            P = P.Add(v.MulScalar(std::sin(angle) * radius));
            d = focalPoint.Sub(P).Normalize();
        }
        return Ray(P, d);
    }
};

}  // namespace orc
