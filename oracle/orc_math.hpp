// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity unpinned (the reference ships no tests, golden
// vectors or runnable binary in this environment; see DESIGN.md "Oracle").
//
// CPU restatement of PTSharp's numeric model.  Every function cites the reference file:line it
// follows (paths relative to /root/reference/PTSharpCore/).  Nothing under ptsharp_b200/ may include,
// link or execute this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do.
//
// Build flags that matter: -ffp-contract=off (no FMA contraction), SSE2 doubles/floats (x86-64 default).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

// Util.cs:10-11
static const double INF = 1e9;
static const double EPS = 1e-9;
// Hit.cs:6 — `1e9F` is exactly representable in float, widened to double.
static const double HIT_INF = (double)1e9f;

// .NET (Core 3.0+) System.Math.Min/Max on doubles: NaN-propagating, -0 < +0.
static inline bool is_negative(double v) { return std::signbit(v); }
static inline double net_max(double a, double b) {
    if (a != b) {
        if (!(a != a)) return b < a ? a : b;
        return a;
    }
    return is_negative(b) ? a : b;
}
static inline double net_min(double a, double b) {
    if (a != b) {
        if (!(a != a)) return a < b ? a : b;
        return a;
    }
    return is_negative(a) ? a : b;
}

// Vector.cs:193-543 — storage is System.Numerics.Vector3 (3 x float); every ctor casts to float
// (Vector.cs:222-227); getters widen to double (Vector.cs:204-220).
struct Vector {
    float x = 0.f, y = 0.f, z = 0.f;
    Vector() {}
    Vector(double X, double Y, double Z) : x((float)X), y((float)Y), z((float)Z) {}
    double X() const { return (double)x; }
    double Y() const { return (double)y; }
    double Z() const { return (double)z; }

    // Vector.cs:408-417 — element-wise: double arithmetic on widened floats, one rounding to float.
    Vector Add(const Vector& b) const { return Vector(X() + b.X(), Y() + b.Y(), Z() + b.Z()); }
    Vector Sub(const Vector& b) const { return Vector(X() - b.X(), Y() - b.Y(), Z() - b.Z()); }
    Vector Mul(const Vector& b) const { return Vector(X() * b.X(), Y() * b.Y(), Z() * b.Z()); }
    Vector Div(const Vector& b) const { return Vector(X() / b.X(), Y() / b.Y(), Z() / b.Z()); }
    // Vector.cs:420-426
    Vector Mod(const Vector& b) const {
        double mx = X() - b.X() * std::floor(X() / b.X());
        double my = Y() - b.Y() * std::floor(Y() / b.Y());
        double mz = Z() - b.Z() * std::floor(Z() / b.Z());
        return Vector(mx, my, mz);
    }
    // Vector.cs:429-438
    Vector AddScalar(double b) const { return Vector(X() + b, Y() + b, Z() + b); }
    Vector SubScalar(double b) const { return Vector(X() - b, Y() - b, Z() - b); }
    Vector MulScalar(double b) const { return Vector(X() * b, Y() * b, Z() * b); }
    Vector DivScalar(double b) const { return Vector(X() / b, Y() / b, Z() / b); }
    // Vector.cs:441-444
    Vector Min(const Vector& b) const { return Vector(net_min(X(), b.X()), net_min(Y(), b.Y()), net_min(Z(), b.Z())); }
    Vector Max(const Vector& b) const { return Vector(net_max(X(), b.X()), net_max(Y(), b.Y()), net_max(Z(), b.Z())); }
    // Vector.cs:396-399
    Vector Negate() const { return Vector(-X(), -Y(), -Z()); }
    Vector Abs() const { return Vector(std::fabs(X()), std::fabs(Y()), std::fabs(Z())); }

    // Vector.cs:370-373 -> System.Numerics.Vector3.Dot: float products, summed (xx+yy)+zz, no FMA
    // (dpps semantics).
    double Dot(const Vector& b) const {
        float xx = x * b.x, yy = y * b.y, zz = z * b.z;
        float s = xx + yy;
        s = s + zz;
        return (double)s;
    }
    // Vector.cs:382-386 -> Vector3.Cross, float, unfused.
    Vector Cross(const Vector& b) const {
        Vector r;
        float a0 = y * b.z, a1 = z * b.y;
        float b0 = z * b.x, b1 = x * b.z;
        float c0 = x * b.y, c1 = y * b.x;
        r.x = a0 - a1;
        r.y = b0 - b1;
        r.z = c0 - c1;
        return r;
    }
    // Vector.cs:356 -> Vector3.Length = MathF.Sqrt(Dot(v,v))
    double Length() const { return (double)std::sqrt((float)Dot(*this)); }
    // Vector.cs:389-393 -> Vector3.Normalize = value / value.Length() (float division per lane)
    Vector Normalize() const {
        float len = std::sqrt((float)Dot(*this));
        Vector r;
        r.x = x / len;
        r.y = y / len;
        r.z = z / len;
        return r;
    }
    // Vector.cs:359-367
    double LengthN(double n) const {
        if (n == 2) return Length();
        Vector a = Abs();
        return std::pow(std::pow(a.X(), n) + std::pow(a.Y(), n) + std::pow(a.Z(), n), 1 / n);
    }
    // Vector.cs:491-494
    double MinComponent() const { return net_min(net_min(X(), Y()), Z()); }
    double MaxComponent() const { return net_max(net_max(X(), Y()), Z()); }
    // Vector.cs:451-454
    bool Equals(const Vector& b) const { return X() == b.X() && Y() == b.Y() && Z() == b.Z(); }

    // Vector.cs:497 — this = normal n, argument = incident i:  i - n*(2*(n.i))
    Vector Reflect(const Vector& i) const { return i.Sub(MulScalar(2 * Dot(i))); }
    // Vector.cs:500-514
    Vector Refract(const Vector& i, double n1, double n2) const {
        double nr = n1 / n2;
        double cosI = -Dot(i);
        double sinT2 = nr * nr * (1 - cosI * cosI);
        if (sinT2 > 1) return Vector();
        double cosT = std::sqrt(1 - sinT2);
        return i.MulScalar(nr).Add(MulScalar(nr * cosI - cosT));
    }
    // Vector.cs:517-536
    double Reflectance(const Vector& i, double n1, double n2) const {
        double nr2 = (n1 * n1) / (n2 * n2);
        double cosI = -Dot(i);
        double sinT2 = nr2 * (1 - cosI * cosI);
        if (sinT2 > 1) return 1;
        double cosT = std::sqrt(1 - sinT2);
        double cosI_n1 = n1 * cosI;
        double cosT_n2 = n2 * cosT;
        double rOrth = (cosI_n1 - cosT_n2) / (cosI_n1 + cosT_n2);
        double rPar = (cosT_n2 - cosI_n1) / (cosT_n2 + cosI_n1);
        return (rOrth * rOrth + rPar * rPar) / 2;
    }
};

// Colour.cs:8-256 — 3 x double.
struct Colour {
    double r = 0, g = 0, b = 0;
    Colour() {}
    Colour(double R, double G, double B) : r(R), g(G), b(B) {}
    Colour Add(const Colour& o) const { return Colour(r + o.r, g + o.g, b + o.b); }        // :231
    Colour Sub(const Colour& o) const { return Colour(r - o.r, g - o.g, b - o.b); }        // :234
    Colour Mul(const Colour& o) const { return Colour(r * o.r, g * o.g, b * o.b); }        // :237
    Colour MulScalar(double s) const { return Colour(r * s, g * s, b * s); }                // :228
    Colour DivScalar(double s) const { return Colour(r / s, g / s, b / s); }                // :243
    Colour Pow(double e) const { return Colour(std::pow(r, e), std::pow(g, e), std::pow(b, e)); }  // :136
    // Colour.cs:219-224
    Colour Mix(const Colour& o, double pct) const { return MulScalar(1 - pct).Add(o.MulScalar(pct)); }
    double MaxComponent() const { return net_max(net_max(r, g), b); }                      // :255
    // Colour.cs:125-132 — 8-bit channels / 255.0f (float division), then Pow(2.2f).
    static Colour HexColor(int x) {
        float red = (float)((x >> 16) & 0xff) / 255.0f;
        float green = (float)((x >> 8) & 0xff) / 255.0f;
        float blue = (float)(x & 0xff) / 255.0f;
        return Colour(red, green, blue).Pow((double)2.2f);
    }
};

struct Box;

// Matrix.cs:8-231 — 16 x double, row-major Mrc.
struct Matrix {
    double m[4][4];
    Matrix() { std::memset(m, 0, sizeof(m)); }
    static Matrix FromRows(const double* v) {
        Matrix r;
        std::memcpy(r.m, v, sizeof(r.m));
        return r;
    }
    static Matrix Identity() {
        Matrix r;
        r.m[0][0] = r.m[1][1] = r.m[2][2] = r.m[3][3] = 1;
        return r;
    }
    // Matrix.cs:33-36 — ignores `this`.
    static Matrix Translate(const Vector& v) {
        Matrix r = Identity();
        r.m[0][3] = v.X(); r.m[1][3] = v.Y(); r.m[2][3] = v.Z();
        return r;
    }
    // Matrix.cs:38-41
    static Matrix Scale(const Vector& v) {
        Matrix r = Identity();
        r.m[0][0] = v.X(); r.m[1][1] = v.Y(); r.m[2][2] = v.Z();
        return r;
    }
    // Matrix.cs:44-54
    static Matrix Rotate(Vector v, double a) {
        v = v.Normalize();
        double s = std::sin(a), c = std::cos(a), k = 1 - c;
        double vx = v.X(), vy = v.Y(), vz = v.Z();
        Matrix r;
        r.m[0][0] = k * vx * vx + c;      r.m[0][1] = k * vx * vy + vz * s; r.m[0][2] = k * vz * vx - vy * s; r.m[0][3] = 0;
        r.m[1][0] = k * vx * vy - vz * s; r.m[1][1] = k * vy * vy + c;      r.m[1][2] = k * vy * vz + vx * s; r.m[1][3] = 0;
        r.m[2][0] = k * vz * vx + vy * s; r.m[2][1] = k * vy * vz - vx * s; r.m[2][2] = k * vz * vz + c;      r.m[2][3] = 0;
        r.m[3][0] = 0; r.m[3][1] = 0; r.m[3][2] = 0; r.m[3][3] = 1;
        return r;
    }
    // Matrix.cs:111-131 — left-to-right sums of four products.
    Matrix Mul(const Matrix& b) const {
        Matrix r;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++)
                r.m[i][j] = m[i][0] * b.m[0][j] + m[i][1] * b.m[1][j] + m[i][2] * b.m[2][j] + m[i][3] * b.m[3][j];
        return r;
    }
    // Matrix.cs:134-141
    Vector MulPosition(const Vector& b) const {
        double X = m[0][0] * b.X() + m[0][1] * b.Y() + m[0][2] * b.Z() + m[0][3];
        double Y = m[1][0] * b.X() + m[1][1] * b.Y() + m[1][2] * b.Z() + m[1][3];
        double Z = m[2][0] * b.X() + m[2][1] * b.Y() + m[2][2] * b.Z() + m[2][3];
        return Vector(X, Y, Z);
    }
    // Matrix.cs:144-150 — NB: re-normalises.
    Vector MulDirection(const Vector& b) const {
        double X = m[0][0] * b.X() + m[0][1] * b.Y() + m[0][2] * b.Z();
        double Y = m[1][0] * b.X() + m[1][1] * b.Y() + m[1][2] * b.Z();
        double Z = m[2][0] * b.X() + m[2][1] * b.Y() + m[2][2] * b.Z();
        return Vector(X, Y, Z).Normalize();
    }
    // Matrix.cs:176
    Matrix Transpose() const {
        Matrix r;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) r.m[i][j] = m[j][i];
        return r;
    }
    // Matrix.cs:179-193 — 24 signed triple products in the reference's order.
    double Determinant() const {
        const double M11 = m[0][0], M12 = m[0][1], M13 = m[0][2], M14 = m[0][3];
        const double M21 = m[1][0], M22 = m[1][1], M23 = m[1][2], M24 = m[1][3];
        const double M31 = m[2][0], M32 = m[2][1], M33 = m[2][2], M34 = m[2][3];
        const double M41 = m[3][0], M42 = m[3][1], M43 = m[3][2], M44 = m[3][3];
        return (M11 * M22 * M33 * M44 - M11 * M22 * M34 * M43 +
                M11 * M23 * M34 * M42 - M11 * M23 * M32 * M44 +
                M11 * M24 * M32 * M43 - M11 * M24 * M33 * M42 -
                M12 * M23 * M34 * M41 + M12 * M23 * M31 * M44 -
                M12 * M24 * M31 * M43 + M12 * M24 * M33 * M41 -
                M12 * M21 * M33 * M44 + M12 * M21 * M34 * M43 +
                M13 * M24 * M31 * M42 - M13 * M24 * M32 * M41 +
                M13 * M21 * M32 * M44 - M13 * M21 * M34 * M42 +
                M13 * M22 * M34 * M41 - M13 * M22 * M31 * M44 -
                M14 * M21 * M32 * M43 + M14 * M21 * M33 * M42 -
                M14 * M22 * M33 * M41 + M14 * M22 * M31 * M43 -
                M14 * M23 * M31 * M42 + M14 * M23 * M32 * M41);
    }
    // Matrix.cs:196-217
    Matrix Inverse() const {
        const double M11 = m[0][0], M12 = m[0][1], M13 = m[0][2], M14 = m[0][3];
        const double M21 = m[1][0], M22 = m[1][1], M23 = m[1][2], M24 = m[1][3];
        const double M31 = m[2][0], M32 = m[2][1], M33 = m[2][2], M34 = m[2][3];
        const double M41 = m[3][0], M42 = m[3][1], M43 = m[3][2], M44 = m[3][3];
        Matrix r;
        double d = Determinant();
        r.m[0][0] = (M23 * M34 * M42 - M24 * M33 * M42 + M24 * M32 * M43 - M22 * M34 * M43 - M23 * M32 * M44 + M22 * M33 * M44) / d;
        r.m[0][1] = (M14 * M33 * M42 - M13 * M34 * M42 - M14 * M32 * M43 + M12 * M34 * M43 + M13 * M32 * M44 - M12 * M33 * M44) / d;
        r.m[0][2] = (M13 * M24 * M42 - M14 * M23 * M42 + M14 * M22 * M43 - M12 * M24 * M43 - M13 * M22 * M44 + M12 * M23 * M44) / d;
        r.m[0][3] = (M14 * M23 * M32 - M13 * M24 * M32 - M14 * M22 * M33 + M12 * M24 * M33 + M13 * M22 * M34 - M12 * M23 * M34) / d;
        r.m[1][0] = (M24 * M33 * M41 - M23 * M34 * M41 - M24 * M31 * M43 + M21 * M34 * M43 + M23 * M31 * M44 - M21 * M33 * M44) / d;
        r.m[1][1] = (M13 * M34 * M41 - M14 * M33 * M41 + M14 * M31 * M43 - M11 * M34 * M43 - M13 * M31 * M44 + M11 * M33 * M44) / d;
        r.m[1][2] = (M14 * M23 * M41 - M13 * M24 * M41 - M14 * M21 * M43 + M11 * M24 * M43 + M13 * M21 * M44 - M11 * M23 * M44) / d;
        r.m[1][3] = (M13 * M24 * M31 - M14 * M23 * M31 + M14 * M21 * M33 - M11 * M24 * M33 - M13 * M21 * M34 + M11 * M23 * M34) / d;
        r.m[2][0] = (M22 * M34 * M41 - M24 * M32 * M41 + M24 * M31 * M42 - M21 * M34 * M42 - M22 * M31 * M44 + M21 * M32 * M44) / d;
        r.m[2][1] = (M14 * M32 * M41 - M12 * M34 * M41 - M14 * M31 * M42 + M11 * M34 * M42 + M12 * M31 * M44 - M11 * M32 * M44) / d;
        r.m[2][2] = (M12 * M24 * M41 - M14 * M22 * M41 + M14 * M21 * M42 - M11 * M24 * M42 - M12 * M21 * M44 + M11 * M22 * M44) / d;
        r.m[2][3] = (M14 * M22 * M31 - M12 * M24 * M31 - M14 * M21 * M32 + M11 * M24 * M32 + M12 * M21 * M34 - M11 * M22 * M34) / d;
        r.m[3][0] = (M23 * M32 * M41 - M22 * M33 * M41 - M23 * M31 * M42 + M21 * M33 * M42 + M22 * M31 * M43 - M21 * M32 * M43) / d;
        r.m[3][1] = (M12 * M33 * M41 - M13 * M32 * M41 + M13 * M31 * M42 - M11 * M33 * M42 - M12 * M31 * M43 + M11 * M32 * M43) / d;
        r.m[3][2] = (M13 * M22 * M41 - M12 * M23 * M41 - M13 * M21 * M42 + M11 * M23 * M42 + M12 * M21 * M43 - M11 * M22 * M43) / d;
        r.m[3][3] = (M12 * M23 * M31 - M13 * M22 * M31 + M13 * M21 * M32 - M11 * M23 * M32 - M12 * M21 * M33 + M11 * M22 * M33) / d;
        return r;
    }
};

// Util.cs:108-113
static inline void Modf(double input, int& dec, double& frac) {
    double tr = std::trunc(input);
    dec = (int)tr;
    frac = input - tr;
}
// Util.cs:131-138
static inline int ClampInt(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
// Util.cs:13
static inline double Radians(double deg) { return deg * M_PI / 180; }

}  // namespace orc
