// mc.cpp — host-side authoring, SURVEY 8f rank 4 (second half): marching cubes over an SDF (MC.cs:7-127) and the
// spherical-harmonic solids built on it (SH.cs:7-104).  Pure preprocessing: the output is an ordinary Mesh whose triangles
// come out in the reference's order (cells x-major, then y, then z; per cell the Lorensen-Cline case table in Bourke's order,
// mc_table.inc), so the reference builder grows the same kd-tree over it.  Nothing here is on the render path.
#include <cmath>
#include <functional>
#include <stdexcept>

#include "../csrc/sh_funcs.hpp"
#include "ptsharp.hpp"

namespace ptsharp {

namespace {

// ---- the case table ------------------------------------------------------------------------------------------------------
struct McTable {
    int count[256];      // triangles of the case
    int tri[256][15];    // cube edges, three per triangle
    int edges[256];      // bit e set: edge e is cut = its two corners lie on different sides (MC.cs:135-166 lists the same masks)
};
const int kPair[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};  // MC.cs:129-133
const McTable& Table() {
    static const McTable t = [] {
        static const char enc[] =
#include "mc_table.inc"
            ;
        McTable r{};
        int c = 0, n = 0;
        for (const char* p = enc;; p++) {
            if (*p == '.' || *p == 0) {
                if (n % 3 != 0 || c > 255) throw std::logic_error("mc_table.inc is malformed");
                r.count[c++] = n / 3; n = 0;
                if (*p == 0) break;
                continue;
            }
            if (n >= 15) throw std::logic_error("mc_table.inc is malformed");
            r.tri[c][n++] = *p <= '9' ? *p - '0' : *p - 'a' + 10;
        }
        if (c != 256) throw std::logic_error("mc_table.inc is malformed");
        for (int i = 0; i < 256; i++)
            for (int e = 0; e < 12; e++)
                if (((i >> kPair[e][0]) & 1) != ((i >> kPair[e][1]) & 1)) r.edges[i] |= 1 << e;
        return r;
    }();
    return t;
}

// MC.mcInterpolate (MC.cs:113-126)
Vector Interpolate(const Vector& p1, const Vector& p2, double v1, double v2, double x) {
    const double EPS = 1e-9;
    if (std::fabs(x - v1) < EPS) return p1;
    if (std::fabs(x - v2) < EPS) return p2;
    if (std::fabs(v1 - v2) < EPS) return p1;
    const double t = (x - v1) / (v2 - v1);
    return Vector(p1.X() + t * (p2.X() - p1.X()), p1.Y() + t * (p2.Y() - p1.Y()), p1.Z() + t * (p2.Z() - p1.Z()));
}

// MC.mcPolygonize (MC.cs:68-111): appends the triangles of one cell
void Polygonize(const Vector p[8], const double v[8], double x, std::vector<Triangle>& out) {
    const McTable& T = Table();
    int index = 0;
    for (int i = 0; i < 8; i++) if (v[i] < x) index |= 1 << i;
    if (T.edges[index] == 0) return;
    Vector points[12];
    for (int i = 0; i < 12; i++)
        if (T.edges[index] & (1 << i)) points[i] = Interpolate(p[kPair[i][0]], p[kPair[i][1]], v[kPair[i][0]], v[kPair[i][1]], x);
    for (int i = 0; i < T.count[index]; i++) {
        Triangle t;  // new Triangle(): `new Material()`, zero normals and texture coordinates
        t.V3 = points[T.tri[index][i * 3 + 0]];
        t.V2 = points[T.tri[index][i * 3 + 1]];
        t.V1 = points[T.tri[index][i * 3 + 2]];
        t.FixNormals();
        out.push_back(t);
    }
}

// SDF.Evaluate for the node types of SDF.cs, run from the linear program SDF::Emit produces (the device evaluates the same program,
// sdf_evaluate in csrc/pt_device.cuh): Vector lanes are floats, scalars doubles.
double LengthN(const Vector& p, double n) {  // Vector.cs:359-367
    if (n == 2) return (double)Length(p);
    return std::pow(std::pow(std::fabs(p.X()), n) + std::pow(std::fabs(p.Y()), n) + std::pow(std::fabs(p.Z()), n), 1 / n);
}
Vector MulPos(const double* m, const Vector& b) {  // Matrix.MulPosition
    return Vector(m[0] * b.X() + m[1] * b.Y() + m[2] * b.Z() + m[3], m[4] * b.X() + m[5] * b.Y() + m[6] * b.Z() + m[7],
                  m[8] * b.X() + m[9] * b.Y() + m[10] * b.Z() + m[11]);
}
double RunProgram(const std::vector<ptgpu_sdf_op>& prog, Vector p) {
    double vs[64];
    Vector ps[64];
    int nv = 0, np = 0;
    for (const ptgpu_sdf_op& op : prog) {
        if (nv >= 63 || np >= 63) throw std::runtime_error("SDF program too deep for the host evaluator");
        switch (op.op) {
            case PTGPU_SDF_SPHERE: vs[nv++] = LengthN(p, op.p[1]) - op.p[0]; break;
            case PTGPU_SDF_CUBE: {
                double x = std::fabs(p.X()), y = std::fabs(p.Y()), z = std::fabs(p.Z());
                x -= (double)(float)op.p[0] / 2; y -= (double)(float)op.p[1] / 2; z -= (double)(float)op.p[2] / 2;
                double a = x;
                if (y > a) a = y;
                if (z > a) a = z;
                if (a > 0) a = 0;
                if (x < 0) x = 0;
                if (y < 0) y = 0;
                if (z < 0) z = 0;
                vs[nv++] = a + std::sqrt(x * x + y * y + z * z);
                break;
            }
            case PTGPU_SDF_CYLINDER: {
                double x = std::sqrt(p.X() * p.X() + p.Z() * p.Z()), y = std::fabs(p.Y());
                x -= op.p[0]; y -= op.p[1] / 2;
                double a = x;
                if (y > a) a = y;
                if (a > 0) a = 0;
                if (x < 0) x = 0;
                if (y < 0) y = 0;
                vs[nv++] = a + std::sqrt(x * x + y * y);
                break;
            }
            case PTGPU_SDF_CAPSULE: {
                const Vector A(op.p[0], op.p[1], op.p[2]), B(op.p[3], op.p[4], op.p[5]);
                const Vector pa = Sub(p, A), ba = Sub(B, A);
                const double h = NetMax(0, NetMin(1, (double)Dot(pa, ba) / (double)Dot(ba, ba)));
                vs[nv++] = LengthN(Sub(pa, MulScalar(ba, h)), op.p[7]) - op.p[6];
                break;
            }
            case PTGPU_SDF_TORUS: {
                const Vector q(LengthN(Vector(p.X(), p.Y(), 0), op.p[2]) - op.p[0], p.Z(), 0);
                vs[nv++] = LengthN(q, op.p[3]) - op.p[1];
                break;
            }
            case PTGPU_SDF_PUSH_TRANSFORM: ps[np++] = p; p = MulPos(op.p, p); break;
            case PTGPU_SDF_PUSH_SCALE: ps[np++] = p; p = Vector(p.X() / op.p[0], p.Y() / op.p[0], p.Z() / op.p[0]); break;
            case PTGPU_SDF_PUSH_REPEAT: {
                ps[np++] = p;
                const Vector st(op.p[0], op.p[1], op.p[2]);
                const Vector q(p.X() - st.X() * std::floor(p.X() / st.X()), p.Y() - st.Y() * std::floor(p.Y() / st.Y()), p.Z() - st.Z() * std::floor(p.Z() / st.Z()));
                p = Sub(q, Vector(st.X() / 2, st.Y() / 2, st.Z() / 2));
                break;
            }
            case PTGPU_SDF_POP:
                p = ps[--np];
                if (op.n == 1) vs[nv - 1] = vs[nv - 1] * op.p[0];
                break;
            case PTGPU_SDF_UNION: case PTGPU_SDF_DIFFERENCE: case PTGPU_SDF_INTERSECTION: {
                const int base = nv - (int)op.n;
                double r = vs[base];
                for (int k = 1; k < (int)op.n; k++) {
                    const double q = vs[base + k];
                    if (op.op == PTGPU_SDF_UNION) { if (q < r) r = q; }
                    else if (op.op == PTGPU_SDF_DIFFERENCE) { if (-q > r) r = -q; }
                    else if (q > r) r = q;
                }
                nv = base; vs[nv++] = r;
                break;
            }
            default: throw std::runtime_error("unknown SDF op");
        }
    }
    return nv > 0 ? vs[nv - 1] : 0.0;
}

}  // namespace

// MC.NewSDFMesh (MC.cs:9-66) over any scalar field
std::shared_ptr<Mesh> MC::NewFieldMesh(const std::function<double(const Vector&)>& evaluate, const Box& box, double step) {
    const Vector mn = box.Min, size = box.Size();
    const int nx = (int)std::ceil(size.X() / step), ny = (int)std::ceil(size.Y() / step), nz = (int)std::ceil(size.Z() / step);
    const double sx = size.X() / nx, sy = size.Y() / ny, sz = size.Z() / nz;
    std::vector<Triangle> triangles;
    for (int x = 0; x < nx - 1; x++)
        for (int y = 0; y < ny - 1; y++)
            for (int z = 0; z < nz - 1; z++) {
                const double x0 = x * sx + mn.X(), y0 = y * sy + mn.Y(), z0 = z * sz + mn.Z();
                const double x1 = x0 + sx, y1 = y0 + sy, z1 = z0 + sz;
                const Vector p[8] = {Vector(x0, y0, z0), Vector(x1, y0, z0), Vector(x1, y1, z0), Vector(x0, y1, z0),
                                     Vector(x0, y0, z1), Vector(x1, y0, z1), Vector(x1, y1, z1), Vector(x0, y1, z1)};
                double v[8];
                for (int i = 0; i < 8; i++) v[i] = evaluate(p[i]);
                Polygonize(p, v, 0, triangles);
            }
    return Mesh::NewMesh(std::move(triangles));
}
std::shared_ptr<Mesh> MC::NewSDFMesh(const SDFPtr& sdf, const Box& box, double step) {
    std::vector<ptgpu_sdf_op> prog;
    sdf->Emit(prog);
    return NewFieldMesh([&](const Vector& p) { return RunProgram(prog, p); }, box, step);
}
int MC::CaseTriangles(int index, int out15[15]) {
    if (index < 0 || index > 255) return -1;
    const McTable& T = Table();
    for (int i = 0; i < T.count[index] * 3; i++) out15[i] = T.tri[index][i];
    return T.count[index];
}
int MC::CaseEdges(int index) { return (index < 0 || index > 255) ? -1 : Table().edges[index]; }

// SphericalHarmonic.NewSphericalHarmonic (SH.cs:14-22); `step` is the reference's 0.01F unless a test asks for a coarser mesh
ShapePtr SphericalHarmonic::NewSphericalHarmonic(int l, int m, const Material& pm, const Material& nm, double step) {
    if (!sh_supported(l, m)) throw std::runtime_error("unsupported spherical harmonic");  // the reference prints this and then fails on a null delegate
    auto sh = std::make_shared<SphericalHarmonic>();
    sh->L = l; sh->M = m; sh->PositiveMaterial = pm; sh->NegativeMaterial = nm;
    sh->mesh = MC::NewFieldMesh([&](const Vector& p) { return sh->Evaluate(p); }, sh->BoundingBox(), step);
    return sh;
}
double SphericalHarmonic::EvaluateHarmonic(const Vector& p) const {  // SH.cs:88-91
    const Vector d = Normalize(p);
    return sh_eval(L, M, d.X(), d.Y(), d.Z());
}
double SphericalHarmonic::Evaluate(const Vector& p) const { return (double)Length(p) - std::fabs(EvaluateHarmonic(p)); }  // SH.cs:93-101
Material SphericalHarmonic::MaterialAt(const Vector& p) const { return EvaluateHarmonic(p) < 0 ? NegativeMaterial : PositiveMaterial; }  // SH.cs:62-72

}  // namespace ptsharp
