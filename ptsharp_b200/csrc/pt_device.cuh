// pt_device.cuh — device-side numeric model, shape intersection, kd-tree traversal and surface evaluation.
//
// Numeric model: PTSharp's Vector stores three FP32 lanes; Add/Sub/Mul/Div/Min/Max/MulScalar
// compute in FP64 on widened lanes and round once to FP32 (Vector.cs:408-444); Dot/Cross/Normalize/Length are FP32
// System.Numerics.Vector3 calls (Vector.cs:356-393); scalars (t, tsplit, Fresnel terms) are FP64.  Because FP64 has
// >= 2*24+2 significand bits, fl32(fl64(a op b)) == fl32(a op b) for + - * / on FP32 inputs, so lane-wise ops on two
// Vectors are plain FP32 instructions here; only Vector x double-scalar products need the FP64 multiply.
// The file is compiled with -fmad=false: no a*b+c contraction anywhere, like the .NET JIT.
//
// Self-intersection: bounce and shadow rays start exactly on the surface (EPS = 1e-9 with FP32-stored
// positions), so the rate at which a ray re-hits the surface it left is set by these roundings and is part of the
// reference's converged image.  That is why this code mirrors the reference's arithmetic instead of using a
// conventional FP32 tracer with an origin offset.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ptgpu.h"
#include "mesh_derive.hpp"
#include "sh_funcs.hpp"

#define PT_D __device__ __forceinline__
#define PT_HD __host__ __device__ __forceinline__
#define PT_DN __device__ __noinline__
// k_shade was 215 KB of SASS with every libm call, Philox refill and texture fetch inlined at each call site, and ncu showed
// 41 % of its stall samples in `no_instructions` (instruction-cache misses of divergent warps).  The build keeps ONE copy of
// those bodies (PT_DC = noinline, scalar arguments and results only: nothing forces a value into local memory).
#define PT_DC __device__ __noinline__

// PT_NO_CULL = 1 builds the ARBITER of every shortcut this file takes (tests/test_gpu_parity.py::test_cull_matches_no_cull): no
// padded-bounds test ever skips a subtree, a mesh or an instance (every reference leaf is tested in full), the division-free
// triangle filter is replaced by the reference sequence, no walk is clipped to the running best and shadow rays take the full
// closest-hit walk.  What remains is Tree.Intersect / Node.Intersect / IntersectShapes as written (Tree.cs:31-128).
#ifndef PT_NO_CULL
#define PT_NO_CULL 0
#endif
#ifndef PT_CULL_BOUNDS
#define PT_CULL_BOUNDS (!PT_NO_CULL)   // padded-bounds tests may skip subtrees / meshes / instances
#endif
#ifndef PT_TRI_FILTER
#define PT_TRI_FILTER (!PT_NO_CULL)    // division-free triangle filter in front of the reference sequence
#endif

namespace pt {

static constexpr double kEPS = 1e-9;            // Util.cs:11
static constexpr double kINF = 1e9;             // Util.cs:10
static constexpr double kHitInf = 1000000000.0; // Hit.cs:6 (1e9F is exact)
static constexpr double kPi = 3.14159265358979323846;
static constexpr int kSceneStack = 32;          // scene-level kd stack entries
static constexpr int kMeshStack = 64;           // mesh-level kd stack entries
static constexpr int kSdfValueStack = 16;
static constexpr int kSdfPointStack = 8;

// ---------------------------------------------------------------------------------------------------- vectors
struct V3 { float x, y, z; };
PT_D V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
PT_D V3 v3d(double x, double y, double z) { V3 r; r.x = (float)x; r.y = (float)y; r.z = (float)z; return r; }  // new Vector(double,double,double)
PT_D V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
PT_D V3 vadd(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
PT_D V3 vsub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
PT_D V3 vmul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
PT_D V3 vdiv(V3 a, V3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
PT_D V3 vneg(V3 a) { return v3(-a.x, -a.y, -a.z); }
PT_D V3 vmuls(V3 a, double s) { return v3d((double)a.x * s, (double)a.y * s, (double)a.z * s); }  // MulScalar
PT_D V3 vdivs(V3 a, double s) { return v3d((double)a.x / s, (double)a.y / s, (double)a.z / s); }  // DivScalar
PT_D float vdotf(V3 a, V3 b) { float s = a.x * b.x + a.y * b.y; return s + a.z * b.z; }           // (xx+yy)+zz, unfused
PT_D double vdot(V3 a, V3 b) { return (double)vdotf(a, b); }
PT_D V3 vcross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
PT_D float vlenf(V3 a) { return sqrtf(vdotf(a, a)); }
PT_D V3 vnorm(V3 a) { float l = vlenf(a); return v3(a.x / l, a.y / l, a.z / l); }
PT_D bool veq(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
PT_D bool vzero(V3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }
PT_DC double2 sincos_c(double a) { double s_, c_; sincos(a, &s_, &c_); return make_double2(s_, c_); }  // (sin, cos)
PT_DC double acos_c(double x) { return acos(x); }
PT_DC V3 vnorm_c(V3 a) { return vnorm(a); }  // shading code: one copy of the three IEEE divisions + sqrt
PT_DC double atan2_c(double y, double x) { return atan2(y, x); }
PT_D float vaxis(V3 a, uint32_t axis) { return axis == 1 ? a.x : (axis == 2 ? a.y : a.z); }

// System.Math.Min/Max (.NET Core 3.0+): NaN-propagating, -0 < +0.  CUDA fmin/fmax drop NaNs, so hand-written.
PT_D double netmax(double a, double b) {
    if (a != b) return (a != a) ? a : (b < a ? a : b);
    return (__double2hiint(b) < 0) ? a : b;
}
PT_D double netmin(double a, double b) {
    if (a != b) return (a != a) ? a : (a < b ? a : b);
    return (__double2hiint(a) < 0) ? a : b;
}
// netmin(a, b) when b is known to be a positive number (a running best.T: EPS <= T <= 1e9, never NaN): one comparison.
// a < b -> a; a >= b -> b (a == b > 0: the same value either way); a NaN -> `a >= b` is false -> a, the NaN, as Math.Min does.
PT_D double netmin_best(double a, double bestT) { return !(a >= bestT) ? a : bestT; }
PT_D float netmaxf(float a, float b) {
    if (a != b) return (a != a) ? a : (b < a ? a : b);
    return (__float_as_int(b) < 0) ? a : b;
}
PT_D float netminf(float a, float b) {
    if (a != b) return (a != a) ? a : (a < b ? a : b);
    return (__float_as_int(a) < 0) ? a : b;
}
PT_D V3 vmin(V3 a, V3 b) { return v3(netminf(a.x, b.x), netminf(a.y, b.y), netminf(a.z, b.z)); }
PT_D V3 vmax(V3 a, V3 b) { return v3(netmaxf(a.x, b.x), netmaxf(a.y, b.y), netmaxf(a.z, b.z)); }

// Ray.Position (Ray.cs:19)
PT_D V3 ray_at(V3 o, V3 d, double t) { return vadd(o, vmuls(d, t)); }

// Matrix.MulPosition / MulDirection (Matrix.cs:134-150), m row-major
PT_D V3 mat_pos(const double* __restrict__ m, V3 b) {
    double X = m[0] * (double)b.x + m[1] * (double)b.y + m[2] * (double)b.z + m[3];
    double Y = m[4] * (double)b.x + m[5] * (double)b.y + m[6] * (double)b.z + m[7];
    double Z = m[8] * (double)b.x + m[9] * (double)b.y + m[10] * (double)b.z + m[11];
    return v3d(X, Y, Z);
}
PT_D V3 mat_dir(const double* __restrict__ m, V3 b) {
    double X = m[0] * (double)b.x + m[1] * (double)b.y + m[2] * (double)b.z;
    double Y = m[4] * (double)b.x + m[5] * (double)b.y + m[6] * (double)b.z;
    double Z = m[8] * (double)b.x + m[9] * (double)b.y + m[10] * (double)b.z;
    return vnorm(v3d(X, Y, Z));
}
// Matrix.Transpose().MulDirection (TransformedShape.cs:57): read m column-wise
PT_D V3 mat_dir_transposed(const double* __restrict__ m, V3 b) {
    double X = m[0] * (double)b.x + m[4] * (double)b.y + m[8] * (double)b.z;
    double Y = m[1] * (double)b.x + m[5] * (double)b.y + m[9] * (double)b.z;
    double Z = m[2] * (double)b.x + m[6] * (double)b.y + m[10] * (double)b.z;
    return vnorm(v3d(X, Y, Z));
}

// ---------------------------------------------------------------------------------------------------- device scene
// Derived per Volume at upload: the largest voxel of every 4x4x4 block of the grid, dilated by two voxels (vol_skip).
static constexpr int kVolBlock = 4;
struct VolBlocks { uint64_t first; int32_t nbx, nby, nbz; int32_t pad; };  // blocks -1 .. nb per axis; entry ((bz + 1) * (nby + 2) + by + 1) * (nbx + 2) + bx + 1

struct DScene {
    const ptgpu_shape* shapes;
    const uint32_t* lights;
    const ptgpu_tree* trees;
    const ptgpu_node* nodes;
    const uint32_t* leafItems;
    const ptgpu_sphere* spheres;
    const ptgpu_cube* cubes;
    const ptgpu_plane* planes;
    const ptgpu_cylinder* cylinders;
    const ptgpu_mesh* meshes;
    const float4* triGeom;          // 3 x float4 per triangle
    const ptgpu_tri_shade* triShade;
    const ptgpu_instance* instances;
    const ptgpu_sdf_shape* sdfShapes;
    const ptgpu_sdf_op* sdfOps;
    const ptgpu_volume* volumes;
    const ptgpu_sh* shs;
    const ptgpu_volume_window* volumeWindows;
    const double* volumeData;
    const ptgpu_material* materials;
    const ptgpu_texture* textures;
    const double4* texels;        // 4 doubles per texel (the reference's Colour precision, see ptgpu.h)
    // Derived at upload for mesh trees (see "mesh traversal" below):
    const uint4* meshNodes;         // 4 x uint4 per node: reference kd nodes, bounds-only nodes and micro leaves (see mesh_step)
    const float4* instBounds;       // 2 x float4 per TransformedShape of a Mesh: padded WORLD-space bounds of the instance (FP32 pre-test)
    // Scene.tree level: which of the first kMaskShapes scene shapes a ray still has to evaluate (see "candidate mask" at scene_advance)
    const uint4* sceneLeafMask;     // 2 x uint4 per Scene.tree node (node - root): the shapes in the leaf; bit 255 = it holds shapes >= kMaskShapes
    const float4* candBlocks;       // 2 x float4 per block of up to 16 instanced meshes: union of their instBounds; lo.w = first member, hi.w = count
    const float4* candMembers;      // 2 x float4 per member: its instBounds; lo.w = index in Scene.Shapes
    uint32_t numCandBlocks;
    uint32_t shadeSurfaces, shadeSub;  // split of the shade-order bins between surfaces and patches (shade_bin in ptgpu.cu)
    uint32_t maskOn;                // 0: Scene.tree repeats (almost) nothing - the mask would only cost its own upkeep
    uint32_t hasNested;             // a TransformedShape of a TransformedShape exists (ptgpu_instance.pad[0])
    uint32_t maskBase[8];           // bits of the scene shapes no bounds test can drop (everything but instanced meshes) + bit 255
    const float4* leafGeom;         // 3 x float4 per leaf triangle in sorted order: (V1, triangle id) (e1, position in the leaf) (e2, -)
    const VolBlocks* volBlocks;     // per Volume: where its table of block maxima sits in volBlockMax (see vol_skip)
    const double* volBlockMax;
    uint32_t sceneTree, numSceneShapes, numLights, numShapes;
    double envColor[3];
    int32_t envTexture;
    double envTextureAngle;
};

struct HitRec {
    double t;       // Hit.T
    double tInner;  // object-space T of the inner hit when shape is a TransformedShape
    int32_t shape;  // index in Scene.Shapes (top level), -1 = miss
    int32_t prim;   // global triangle index, -1 if the hit shape is not a triangle
};

// ---------------------------------------------------------------------------------------------------- Box.Intersect
// Box.cs:72-94
PT_D void box_intersect(const float* __restrict__ bmin, const float* __restrict__ bmax, V3 o, V3 d, double& tmin, double& tmax) {
    double ox = o.x, oy = o.y, oz = o.z, dx = d.x, dy = d.y, dz = d.z;
    double x1 = ((double)bmin[0] - ox) / dx, y1 = ((double)bmin[1] - oy) / dy, z1 = ((double)bmin[2] - oz) / dz;
    double x2 = ((double)bmax[0] - ox) / dx, y2 = ((double)bmax[1] - oy) / dy, z2 = ((double)bmax[2] - oz) / dz;
    if (x1 > x2) { double t = x1; x1 = x2; x2 = t; }
    if (y1 > y2) { double t = y1; y1 = y2; y2 = t; }
    if (z1 > z2) { double t = z1; z1 = z2; z2 = t; }
    tmin = netmax(netmax(x1, y1), z1);
    tmax = netmin(netmin(x2, y2), z2);
}

// ---------------------------------------------------------------------------------------------------- primitives
// Sphere.cs:40-60
PT_D double sphere_intersect(const ptgpu_sphere& s, V3 o, V3 d) {
    V3 to = vsub(o, ld3(s.center));
    double b = vdot(to, d);
    double c = vdot(to, to) - s.radius * s.radius;
    double disc = b * b - c;
    if (disc > 0) {
        disc = sqrt(disc);
        double t1 = -b - disc;
        if (t1 > kEPS) return t1;
        double t2 = -b + disc;
        if (t2 > kEPS) return t2;
    }
    return kHitInf;
}
// Cube.cs:35-47
PT_D double cube_intersect(const ptgpu_cube& c, V3 o, V3 d) {
    V3 n = vdiv(vsub(ld3(c.min), o), d);
    V3 f = vdiv(vsub(ld3(c.max), o), d);
    V3 n2 = vmin(n, f), f2 = vmax(n, f);
    double t0 = netmax(netmax((double)n2.x, (double)n2.y), (double)n2.z);
    double t1 = netmin(netmin((double)f2.x, (double)f2.y), (double)f2.z);
    if (t0 > 0 && t0 < t1) return t0;
    return kHitInf;
}
// Plane.cs:38-52
PT_D double plane_intersect(const ptgpu_plane& p, V3 o, V3 d) {
    V3 N = ld3(p.normal);
    double dd = vdot(N, d);
    if (fabs(dd) < kEPS) return kHitInf;
    V3 a = vsub(ld3(p.point), o);
    double t = vdot(a, N) / dd;
    if (t < kEPS) return kHitInf;
    return t;
}
// Cylinder.cs:43-111 — first test that passes wins, in the reference's order (caps, then far root, then near root).
PT_D double cylinder_intersect(const ptgpu_cylinder& cy, V3 o, V3 d) {
    double r = cy.radius;
    double ox = o.x, oy = o.y, oz = o.z, dx = d.x, dy = d.y, dz = d.z;
    double tTop = (cy.z1 - oz) / dz;
    double tBottom = (cy.z0 - oz) / dz;
    double a = dx * dx + dy * dy;
    double b = 2 * (ox * dx + oy * dy);
    double c = ox * ox + oy * oy - r * r;
    double discriminant = b * b - 4 * a * c;
    if (tTop > kEPS && tTop > 0) {
        V3 p = vadd(o, vmuls(d, tTop));
        double dist = sqrt((double)p.x * (double)p.x + (double)p.y * (double)p.y);
        if (dist <= r) return tTop;
    }
    if (tBottom > kEPS && tBottom > 0) {
        V3 p = vadd(o, vmuls(d, tBottom));
        double dist = sqrt((double)p.x * (double)p.x + (double)p.y * (double)p.y);
        if (dist <= r) return tBottom;
    }
    if (discriminant >= 0) {
        double sq = sqrt(discriminant);
        double t1 = (-b + sq) / (2 * a);
        double t2 = (-b - sq) / (2 * a);
        double tl = 0;
        bool have = false;
        if (t1 > kEPS && t1 > 0) { tl = t1; have = true; }
        else if (t2 > kEPS && t2 > 0) { tl = t2; have = true; }
        if (have) {
            V3 p = vadd(o, vmuls(d, tl));
            double z = p.z;
            if (z >= cy.z0 && z <= cy.z1) return tl;
        }
    }
    return kHitInf;
}
// Triangle.cs:95-124 (e1, e2 precomputed by the host exactly as V2.Sub(V1), V3.Sub(V1)) — the reference sequence,
// one FP64 division per call.  Kept as the arbiter for the borderline cases of the filtered version below.
PT_D double triangle_intersect_exact(V3 v1, V3 e1, V3 e2, V3 o, V3 d) {
    V3 h = vcross(d, e2);
    double det = vdot(e1, h);
    if (det > -kEPS && det < kEPS) return kHitInf;
    double invDet = 1.0 / det;
    V3 s = vsub(o, v1);
    double u = vdot(s, h) * invDet;
    if (u < 0 || u > 1) return kHitInf;
    V3 q = vcross(s, e1);
    double v = vdot(d, q) * invDet;
    if (v < 0 || (u + v) > 1) return kHitInf;
    double t = vdot(e2, q) * invDet;
    if (t < kEPS) return kHitInf;
    return t;
}
// Same result, bit for bit, without the division on the (overwhelmingly common) miss path.  In the reference
// det, a = s.h, b = d.q, c = e2.q are FP32 dot products and u = a*(1/det), v = b*(1/det), t = c*(1/det) are FP64, so
//   u < 0        <=>  a and det have opposite signs           (1/det keeps det's sign, the product cannot underflow)
//   u > 1        <=>  |a| > |det| with equal signs            (a, det are FP32: a != det implies |a/det - 1| >= 2^-24)
//   v < 0        <=>  b and det have opposite signs
//   u + v > 1    <=>  (a + b) / det > 1                       unless a + b is within 1e-13 relative of det
//   t < EPS      <=>  c / det < 1e-9                          unless within 1e-13 relative of it
// and the excluded borderline cases (a == det, near-equalities) are sent to the exact sequence above.
PT_D double triangle_intersect_regs(float4 A, float4 B4, float4 C4, V3 o, V3 d);
PT_D double triangle_intersect(const float4* __restrict__ g, V3 o, V3 d) {
#if !PT_TRI_FILTER
    const float4 A = __ldg(g), B4 = __ldg(g + 1), C4 = __ldg(g + 2);
    return triangle_intersect_exact(v3(A.x, A.y, A.z), v3(B4.x, B4.y, B4.z), v3(C4.x, C4.y, C4.z), o, d);
#else
    return triangle_intersect_regs(__ldg(g), __ldg(g + 1), __ldg(g + 2), o, d);
#endif
}
PT_D double triangle_intersect_regs(float4 A, float4 B4, float4 C4, V3 o, V3 d) {
    const V3 v1 = v3(A.x, A.y, A.z), e1 = v3(B4.x, B4.y, B4.z), e2 = v3(C4.x, C4.y, C4.z);
    // all four FP32 dot products up front (a warp pays for its slowest lane anyway), then one decision
    const V3 h = vcross(d, e2);
    const float det = vdotf(e1, h);
    const V3 s = vsub(o, v1);
    const float a = vdotf(s, h);
    const V3 q = vcross(s, e1);
    const float b = vdotf(d, q);
    const float c = vdotf(e2, q);
    const bool pos = det > 0.f;
    const float sa = pos ? a : -a, sb = pos ? b : -b, ad = fabsf(det);
    // misses that need no FP64 at all: |det| <= 1e-9f (1e-9f is the largest float below 1e-9), u < 0, u > 1, v < 0
    const bool quickMiss = (ad <= 1e-9f) | (sa < 0.f) | (sa > ad) | (sb < 0.f);
    const bool odd = !(det == det) | !(a == a) | !(b == b) | !(c == c) | (a == det);  // NaNs and u == 1 +- ulp: ask the reference sequence
    if (odd) return triangle_intersect_exact(v1, e1, e2, o, d);
    if (quickMiss) return kHitInf;
    const double detd = (double)det, adet = fabs(detd);
    const double sum = (double)a + (double)b;                       // exact or within 2^-53
    const double excess = pos ? (sum - detd) : (detd - sum);        // > 0  <=>  (a + b) / det > 1
    const double cs = pos ? (double)c : -(double)c;                 // c / det = cs / |det|
    const double lim = kEPS * adet;
    if (fabs(excess) <= 1e-13 * adet || fabs(cs - lim) <= 1e-13 * lim) return triangle_intersect_exact(v1, e1, e2, o, d);
    if (excess > 0 || cs < lim) return kHitInf;                     // u + v > 1, t < EPS
    return (double)c * (1.0 / detd);                                // the reference's t = e2.q * invDet
}

// ---------------------------------------------------------------------------------------------------- SDF
// Vector.LengthN (Vector.cs:359-367)
PT_D double length_n(V3 p, double n) {
    if (n == 2) return (double)vlenf(p);
    double ax = fabs((double)p.x), ay = fabs((double)p.y), az = fabs((double)p.z);
    return pow(pow(ax, n) + pow(ay, n) + pow(az, n), 1 / n);
}
// Evaluate the SDF program of one SDFShape at p (SDF.cs Evaluate methods, see ptgpu.h for the op encoding).
PT_DN double sdf_evaluate(const ptgpu_sdf_op* __restrict__ prog, uint32_t count, V3 p) {
    double vs[kSdfValueStack];
    V3 ps[kSdfPointStack];
    int nv = 0, np = 0;
    for (uint32_t i = 0; i < count; i++) {
        const ptgpu_sdf_op& op = prog[i];
        switch (op.op) {
            case PTGPU_SDF_SPHERE: vs[nv++] = length_n(p, op.p[1]) - op.p[0]; break;  // SDF.cs:130-133
            case PTGPU_SDF_CUBE: {                                                     // SDF.cs:156-188
                double x = p.x, y = p.y, z = p.z;
                if (x < 0) x = -x;
                if (y < 0) y = -y;
                if (z < 0) z = -z;
                // Size components are Vector lanes (FP32) widened
                x -= (double)(float)op.p[0] / 2; y -= (double)(float)op.p[1] / 2; z -= (double)(float)op.p[2] / 2;
                double a = x;
                if (y > a) a = y;
                if (z > a) a = z;
                if (a > 0) a = 0;
                if (x < 0) x = 0;
                if (y < 0) y = 0;
                if (z < 0) z = 0;
                vs[nv++] = a + sqrt(x * x + y * y + z * z);
                break;
            }
            case PTGPU_SDF_CYLINDER: {  // SDF.cs:226-251
                double x = sqrt((double)p.x * (double)p.x + (double)p.z * (double)p.z);
                double y = p.y;
                if (x < 0) x = -x;
                if (y < 0) y = -y;
                x -= op.p[0];
                y -= op.p[1] / 2;
                double a = x;
                if (y > a) a = y;
                if (a > 0) a = 0;
                if (x < 0) x = 0;
                if (y < 0) y = 0;
                vs[nv++] = a + sqrt(x * x + y * y);
                break;
            }
            case PTGPU_SDF_CAPSULE: {  // SDF.cs:272-278
                V3 A = v3d(op.p[0], op.p[1], op.p[2]), B = v3d(op.p[3], op.p[4], op.p[5]);
                V3 pa = vsub(p, A), ba = vsub(B, A);
                double h = netmax(0, netmin(1, vdot(pa, ba) / vdot(ba, ba)));
                vs[nv++] = length_n(vsub(pa, vmuls(ba, h)), op.p[7]) - op.p[6];
                break;
            }
            case PTGPU_SDF_TORUS: {  // SDF.cs:307-311
                V3 q = v3d(length_n(v3(p.x, p.y, 0.f), op.p[2]) - op.p[0], (double)p.z, 0.0);
                vs[nv++] = length_n(q, op.p[3]) - op.p[1];
                break;
            }
            case PTGPU_SDF_PUSH_TRANSFORM: ps[np++] = p; p = mat_pos(op.p, p); break;          // SDF.cs:340-344
            case PTGPU_SDF_PUSH_SCALE: ps[np++] = p; p = vdivs(p, op.p[0]); break;             // SDF.cs:371-374
            case PTGPU_SDF_PUSH_REPEAT: {                                                      // SDF.cs:549-553, Vector.cs:420-426
                ps[np++] = p;
                V3 st = v3d(op.p[0], op.p[1], op.p[2]);
                double mx = (double)p.x - (double)st.x * floor((double)p.x / (double)st.x);
                double my = (double)p.y - (double)st.y * floor((double)p.y / (double)st.y);
                double mz = (double)p.z - (double)st.z * floor((double)p.z / (double)st.z);
                p = vsub(v3d(mx, my, mz), vdivs(st, 2));
                break;
            }
            case PTGPU_SDF_POP:
                p = ps[--np];
                if (op.n == 1) vs[nv - 1] = vs[nv - 1] * op.p[0];
                break;
            case PTGPU_SDF_UNION: {  // SDF.cs:398-412
                int base = nv - (int)op.n;
                double result = vs[base];
                for (int k = 1; k < (int)op.n; k++) if (vs[base + k] < result) result = vs[base + k];
                nv = base; vs[nv++] = result;
                break;
            }
            case PTGPU_SDF_DIFFERENCE: {  // SDF.cs:452-471
                int base = nv - (int)op.n;
                double result = vs[base];
                for (int k = 1; k < (int)op.n; k++) if (-vs[base + k] > result) result = -vs[base + k];
                nv = base; vs[nv++] = result;
                break;
            }
            case PTGPU_SDF_INTERSECTION: {  // SDF.cs:493-509
                int base = nv - (int)op.n;
                double result = vs[base];
                for (int k = 1; k < (int)op.n; k++) if (vs[base + k] > result) result = vs[base + k];
                nv = base; vs[nv++] = result;
                break;
            }
            default: break;
        }
    }
    return nv > 0 ? vs[nv - 1] : 0.0;
}
// SDFShape.Intersect (SDF.cs:32-76)
PT_D double sdf_intersect(const DScene& S, const ptgpu_sdf_shape& sh, V3 o, V3 d) {
    const double epsilon = (double)0.00001f, start = (double)0.0001f, jumpSize = (double)0.001f;
    double t1, t2;
    box_intersect(sh.bmin, sh.bmax, o, d, t1, t2);
    if (t2 < t1 || t2 < 0) return kHitInf;
    double t = netmax(start, t1);
    bool jump = true;
    const ptgpu_sdf_op* prog = S.sdfOps + sh.progFirst;
    for (int i = 0; i < 1000; i++) {
        double dist = sdf_evaluate(prog, sh.progCount, ray_at(o, d, t));
        if (jump && dist < 0) { t -= jumpSize; jump = false; continue; }
        if (dist < epsilon) return t;
        if (jump && dist < jumpSize) dist = jumpSize;
        t += dist;
        if (t > t2) return kHitInf;
    }
    return kHitInf;
}
// SDFShape.NormalAt (SDF.cs:83-92)
PT_D V3 sdf_normal(const DScene& S, const ptgpu_sdf_shape& sh, V3 p) {
    const double e = 0.0001;
    double x = p.x, y = p.y, z = p.z;
    const ptgpu_sdf_op* prog = S.sdfOps + sh.progFirst;
    uint32_t n = sh.progCount;
    double nx = sdf_evaluate(prog, n, v3d(x - e, y, z)) - sdf_evaluate(prog, n, v3d(x + e, y, z));
    double ny = sdf_evaluate(prog, n, v3d(x, y - e, z)) - sdf_evaluate(prog, n, v3d(x, y + e, z));
    double nz = sdf_evaluate(prog, n, v3d(x, y, z - e)) - sdf_evaluate(prog, n, v3d(x, y, z + e));
    return vnorm(v3d(nx, ny, nz));
}

// ---------------------------------------------------------------------------------------------------- Volume
PT_D double vol_get(const ptgpu_volume& v, const double* __restrict__ data, int x, int y, int z) {  // Volume.cs:40-46
    if (x < 0 || y < 0 || z < 0 || x >= v.w || y >= v.h || z >= v.d) return 0;
    return __ldg(data + v.dataOffset + (size_t)x + (size_t)y * v.w + (size_t)z * v.w * v.h);
}
PT_DN double vol_sample(const ptgpu_volume& v, const double* __restrict__ data, double x, double y, double z) {  // Volume.cs:73-104 (index quirks kept)
    z /= v.zscale;
    x = ((x + 1) / 2) * (double)v.w;
    y = ((z + 1) / 2) * (double)v.h;
    z = ((z + 2) / 2) * (double)v.d;
    int x0 = (int)floor(x), y0 = (int)floor(y), z0 = (int)floor(z);
    int x1 = x0 + 1, y1 = y0 + 1, z1 = z0 + 1;
    double v000 = vol_get(v, data, x0, y0, z0), v001 = vol_get(v, data, x0, y0, z1), v010 = vol_get(v, data, x0, y1, z0), v011 = vol_get(v, data, x0, y1, z1);
    double v100 = vol_get(v, data, x1, y0, z0), v101 = vol_get(v, data, x1, y0, z1), v110 = vol_get(v, data, x1, y1, z0), v111 = vol_get(v, data, x1, y1, z1);
    x -= (double)x0; y -= (double)y0; z -= (double)z0;
    double c00 = v000 * (1 - x) + v100 * x;
    double c01 = v001 * (1 - x) + v101 * x;
    double c10 = v010 * (1 - x) + v110 * x;
    double c11 = v011 * (1 - x) + v111 * x;
    double c0 = c00 * (1 - y) + c10 * y;
    double c1 = c01 * (1 - y) + c11 * y;
    return c0 * (1 - z) + c1 * z;
}
PT_D int vol_sign(const DScene& S, const ptgpu_volume& v, V3 a) {  // Volume.cs:113-131 (`i` never advances)
    double s = vol_sample(v, S.volumeData, (double)a.x, (double)a.y, (double)a.z);
    for (uint32_t k = 0; k < v.windowCount; k++) {
        const ptgpu_volume_window& w = S.volumeWindows[v.windowFirst + k];
        if (s < w.lo) return 1;
        if (s > w.hi) continue;
        return 0;
    }
    return (int)v.windowCount + 1;
}
// t after k more `t += step` of the marching loop (Volume.cs:175), in closed form.  step is a power of two (1/512, then /64 per
// refinement) and at least one ulp of t, so inside a binade every one of those additions is exact and k of them equal ONE exact
// addition of k * step; only the addition that carries t into the next binade can round, and it is performed as such.
PT_D double vol_advance(double t, double step, long long k) {
    while (k > 0) {
        int e;
        frexp(t, &e);
        const double edge = ldexp(1.0, e);                          // the next power of two above t
        const long long j0 = (long long)ceil((edge - t) / step);    // additions until t reaches it (exact: multiples of ulp(t))
        if (k < j0) return t + (double)k * step;
        t += (double)(j0 - 1) * step;
        t += step;                                                  // the reference's own (possibly rounding) addition
        k -= j0;
    }
    return t;
}
// Empty-space skipping that cannot change the result.  While the loop is in its `sign == 1` state (the last Sign() saw a value
// below the first window, Volume.cs:118-119), a step whose Sample() is below windows[0].lo returns 1 again: nothing happens but
// `t += step`.  Sample() is a convex combination of the 8 voxels around the sample's grid position (zero outside the grid), so it
// is at most the largest voxel of the surrounding block; blockMax holds that maximum per 4x4x4 block, dilated by two voxels
// (one for the `+1` corners, one against the FP32 rounding of Ray.Position, which moves a sample by < 1e-4 voxels).  If the block
// the ray is in is below the window, every step up to the block's exit face (1e-3 voxels short of it) is skipped, with t advanced
// exactly as the skipped additions would have.  Returns the number of steps skipped (0: take a normal step).
#ifndef PT_VOL_SKIP
#define PT_VOL_SKIP (!PT_NO_CULL)
#endif
PT_D long long vol_skip(const DScene& S, const ptgpu_volume& v, uint32_t volIndex, V3 co, V3 cd, double t, double step, double tend) {
#if !PT_VOL_SKIP
    return 0;
#endif
    if (!(step >= 1e-12)) return 0;  // after a handful of refinements (step /= 64 each, Volume.cs:183) step nears the ulp of t: plain steps
    const VolBlocks vb = S.volBlocks[volIndex];
    const double lo = v.windowCount ? S.volumeWindows[v.windowFirst].lo : 1e300;
    // grid position of Ray.Position(t) as Volume.Sample maps it (Volume.cs:75-78, quirks included): linear in t
    const double zq0 = (double)co.z / v.zscale, zq1 = (double)cd.z / v.zscale;
    const double a0[3] = {(((double)co.x + 1) / 2) * (double)v.w, ((zq0 + 1) / 2) * (double)v.h, ((zq0 + 2) / 2) * (double)v.d};
    const double a1[3] = {((double)cd.x / 2) * (double)v.w, (zq1 / 2) * (double)v.h, (zq1 / 2) * (double)v.d};
    const int nb[3] = {vb.nbx, vb.nby, vb.nbz};
    int b[3];
    double texit = 1e300;
    const double m = 1e-3;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double p = a0[c] + a1[c] * t;
        const double bf = floor(p / kVolBlock);
        if (!(bf > -1e9 && bf < 1e9)) return 0;                       // NaN / huge: no skipping
        if (p < bf * kVolBlock + m || p > (bf + 1) * kVolBlock - m) return 0;  // too close to a block face
        b[c] = (int)bf;
        if (a1[c] > 0) texit = fmin(texit, ((bf + 1) * kVolBlock - m - a0[c]) / a1[c]);
        else if (a1[c] < 0) texit = fmin(texit, (bf * kVolBlock + m - a0[c]) / a1[c]);
    }
    double bm = 0;  // outside the table every voxel is outside the grid: Sample() = 0
    if (b[0] >= -1 && b[0] <= nb[0] && b[1] >= -1 && b[1] <= nb[1] && b[2] >= -1 && b[2] <= nb[2])
        bm = __ldg(S.volBlockMax + vb.first + ((size_t)(b[2] + 1) * (nb[1] + 2) + (b[1] + 1)) * (nb[0] + 2) + (b[0] + 1));
    if (!(bm < lo)) return 0;
    if (!(texit > t)) return 0;
    // samples t, t + step, ..., t + (k - 1) step lie inside the block; one fewer against the rounding of the quotient
    double kf = floor((texit - t) / step) - 1;
    // the loop ends (NoHit) at the first t > tend: never skip to within two steps of it, the ordinary steps finish the march
    kf = fmin(kf, floor((tend - t) / step) - 2);
    if (!(kf >= 2)) return 0;
    return (long long)fmin(kf, 1e15);
}
PT_D double volume_intersect(const DScene& S, const ptgpu_volume& v, V3 o, V3 d, uint32_t volIndex = 0xFFFFFFFFu) {  // Volume.cs:169-197
    double tmin, tmax;
    box_intersect(v.bmin, v.bmax, o, d, tmin, tmax);
    double step = (double)(1.0f / 512.0f);
    double start = netmax(step, tmin);
    int sign = -1;
    for (double t = start; t <= tmax; t += step) {
        if (sign == 1 && volIndex != 0xFFFFFFFFu) {
            const long long k = vol_skip(S, v, volIndex, o, d, t, step, tmax);
            if (k > 0) { t = vol_advance(t, step, k - 1); continue; }  // k steps: k - 1 here, the loop's own `t += step`
        }
        int s = vol_sign(S, v, ray_at(o, d, t));
        if (s == 0 || (sign >= 0 && s != sign)) {
            t -= step;
            step /= 64;
            t += step;
            for (int i = 0; i < 64; i++) {
                if (vol_sign(S, v, ray_at(o, d, t)) == 0) return t - step;
                t += step;
            }
        }
        sign = s;
    }
    return kHitInf;
}
PT_D V3 volume_normal(const DScene& S, const ptgpu_volume& v, V3 p) {  // Volume.cs:138-145
    const double eps = (double)0.001f;
    double x = p.x, y = p.y, z = p.z;
    const double* D = S.volumeData;
    double nx = vol_sample(v, D, x - eps, y, z) - vol_sample(v, D, x + eps, y, z);
    double ny = vol_sample(v, D, x, y - eps, z) - vol_sample(v, D, x, y + eps, z);
    double nz = vol_sample(v, D, x, y, z - eps) - vol_sample(v, D, x, y, z + eps);
    return vnorm(v3d(nx, ny, nz));
}
PT_D int32_t volume_material(const DScene& S, const ptgpu_volume& v, V3 p) {  // Volume.cs:147-167; -1 = `new Material()`
    double be = (double)1e9f;
    int32_t bm = -1;
    double s = vol_sample(v, S.volumeData, (double)p.x, (double)p.y, (double)p.z);
    for (uint32_t k = 0; k < v.windowCount; k++) {
        const ptgpu_volume_window& w = S.volumeWindows[v.windowFirst + k];
        if (s >= w.lo && s <= w.hi) return w.material;
        double e = netmin(fabs(s - w.lo), fabs(s - w.hi));
        if (e < be) { be = e; bm = w.material; }
    }
    return bm;
}

// ---------------------------------------------------------------------------------------------------- kd traversal
// Tree.Intersect / Node.Intersect (Tree.cs:31-113) in stack form: running best with strict `<`
// updates; on "both children" push (second, tsplit, tmax); on pop skip when best.T <= tsplit, otherwise continue with
// tmax = Math.Min(tmax, best.T).  Comparisons are written exactly as in the reference so NaNs take the same branches.
//
// Execution model.  A naive "one thread walks one ray to completion" kernel measured 2.2 active lanes per warp
// instruction on the 250k-triangle scene (ncu, profiles/r01_trace_naive.txt): per-ray cost is heavy-tailed (a few rays
// cross leaves of hundreds of triangles) and a warp runs as long as its slowest lane.  The tracer is therefore a
// persistent per-lane state machine with ray replacement: every lane owns one ray and advances it by one unit of work
// per state block (a few kd node steps, one leaf triangle, one scene shape); a lane whose ray finishes immediately
// pulls the next ray index from a global cursor (one atomic per coalesced group), so a long ray only ever occupies its
// own lane.  The arithmetic of each unit is untouched, so hits are bit-identical to the recursive reference.
struct KdCursor {  // one level of Tree.Intersect in flight
    uint32_t node;
    double tmin, tmax;
    int sp;
};

// Subtree culling.  The reference tree never splits off empty space (a split with an empty side scores N and is
// rejected, Tree.cs:228-255), so a ray crossing a hollow mesh walks dozens of cells whose triangles it cannot touch
// (measured: 229 triangle tests for rays that miss both meshes of C3).  Every node of a mesh tree therefore carries the
// bounding box of the triangles below it, padded by far more than the FP32 error of the triangle test; a subtree whose
// padded box the ray line misses for all t > 0 cannot change the running best, so skipping it (= returning NoHit from
// Node.Intersect) leaves every later comparison of the reference's control flow unchanged.
struct RayAux { float ix, iy, iz, pad; };  // 1/d per axis (inf where d == 0) and an origin-dependent extra padding
PT_D RayAux ray_aux(V3 o, V3 d) {
    RayAux a;
    a.ix = 1.0f / d.x; a.iy = 1.0f / d.y; a.iz = 1.0f / d.z;
    a.pad = 4e-6f * (fabsf(o.x) + fabsf(o.y) + fabsf(o.z));
    return a;
}
// The same slab test with fewer instructions for the mesh walk: t = lo * i - o * i by FMA, and the per-ray padding applied
// once to the interval ends as P = pad * max|i| (>= pad * |i| on every axis: only more conservative).
struct RayBox {
    float ix, iy, iz, cx, cy, cz, P;
};
PT_D RayBox ray_box(V3 o, V3 d) {
    RayBox a;
    a.ix = 1.0f / d.x; a.iy = 1.0f / d.y; a.iz = 1.0f / d.z;
    a.cx = o.x * a.ix; a.cy = o.y * a.iy; a.cz = o.z * a.iz;  // NaN when o = 0 and d = 0 on an axis: the min/max below ignore NaNs
    const float pad = 4e-6f * (fabsf(o.x) + fabsf(o.y) + fabsf(o.z));
    a.P = pad * fmaxf(fmaxf(fabsf(a.ix), fabsf(a.iy)), fabsf(a.iz)) * 1.0001f;
    return a;
}
PT_D bool box_line_hit_fast(float lox, float loy, float loz, float hix, float hiy, float hiz, const RayBox& r, float& tnear) {
#if !PT_CULL_BOUNDS
    tnear = -INFINITY;
    return true;
#endif
    const float x1 = __fmaf_rn(lox, r.ix, -r.cx), x2 = __fmaf_rn(hix, r.ix, -r.cx);
    const float y1 = __fmaf_rn(loy, r.iy, -r.cy), y2 = __fmaf_rn(hiy, r.iy, -r.cy);
    const float z1 = __fmaf_rn(loz, r.iz, -r.cz), z2 = __fmaf_rn(hiz, r.iz, -r.cz);
    const float tn = fmaxf(fmaxf(fminf(x1, x2), fminf(y1, y2)), fminf(z1, z2)) - r.P;
    const float tf = fminf(fminf(fmaxf(x1, x2), fmaxf(y1, y2)), fmaxf(z1, z2)) + r.P;
    const float slack = 2e-5f * (fabsf(tf) + fabsf(tn)) + 1e-30f;   // the FMA form rounds o * i and lo * i - that separately
    tnear = tn - slack;
    return !(tn > tf + slack) && !(tf < -slack);  // NaN comparisons are false -> treated as a hit
}
enum { KD_INTERIOR = 0, KD_LEAF = 1 };

// IShape.Intersect for the analytic shapes (everything except Mesh and TransformedShape).
PT_D double primitive_intersect(const DScene& S, const ptgpu_shape& sh, V3 o, V3 d) {
    switch (sh.type) {
        case PTGPU_SPHERE: return sphere_intersect(S.spheres[sh.data], o, d);
        case PTGPU_CUBE: return cube_intersect(S.cubes[sh.data], o, d);
        case PTGPU_PLANE: return plane_intersect(S.planes[sh.data], o, d);
        case PTGPU_CYLINDER: return cylinder_intersect(S.cylinders[sh.data], o, d);
        case PTGPU_SDF: return sdf_intersect(S, S.sdfShapes[sh.data], o, d);
        case PTGPU_VOLUME: return volume_intersect(S, S.volumes[sh.data], o, d, sh.data);
        default: return kHitInf;
    }
}

// ---------------------------------------------------------------------------------------------------- mesh traversal
// Tree.Intersect of a Mesh.tree on derived data (same tree, same visiting order, same arithmetic per triangle):
//
//  * 16-byte stack entries {tsplit, node, culled}: the `tmax` handed to the far child is recoverable as
//    netmin(tsplit of the entry below | root tmax, best.T) — at push time c.tmax is the root tmax, the tsplit of the
//    push directly below, or a netmin(.., best.T) of one of those from an earlier pop, and best.T only decreases — so
//    entry 0 is a sentinel holding the root tmax.
//  * Child bounds in the parent (64-byte node): whether a child's padded triangle bounds are missed by the ray line
//    (= its Node.Intersect returns NoHit, see box_line_hit) is known before descending, so a culled near child costs no
//    memory round trip and a culled far child is skipped when it is popped.  It is still pushed: its tsplit is the
//    `tmax` of the entries above it.
//  * Bounds-only nodes below the reference leaves.  Rays mostly cross LARGE leaves (the builder stops at 85 % overlap;
//    a C3 camera ray meets ~60 triangles in ~2 leaves).  The triangles of a leaf are sorted spatially and hung under a
//    small binary hierarchy of padded bounds ending in micro leaves of <= 4 triangles; a subtree the ray line misses is
//    skipped (each of its triangles would return NoHit).  This reorders the leaf, and the reference keeps the FIRST
//    shape in array order among equal T (strict <, Tree.cs:122), so every triangle carries its position in the
//    reference leaf and a tie inside the leaf goes to the lower position; a tie with a hit from an earlier leaf never
//    replaces it (bestPos = 0 at leaf entry).
PT_D void stk_put(uint4* e, double ts, uint32_t node, uint32_t culled) { *e = make_uint4((uint32_t)__double2loint(ts), (uint32_t)__double2hiint(ts), node, culled); }
PT_D double stk_t(const uint4& e) { return __hiloint2double((int)e.y, (int)e.x); }

PT_D bool box_line_hit(float lox, float loy, float loz, float hix, float hiy, float hiz, V3 o, const RayAux& ra, float* entry = nullptr) {
#if !PT_CULL_BOUNDS
    if (entry) *entry = -INFINITY;
    return true;
#endif
    // fminf/fmaxf ignore NaNs (0 * inf when the origin lies on a slab plane of a zero direction): conservative
    const float x1 = (lox - ra.pad - o.x) * ra.ix, x2 = (hix + ra.pad - o.x) * ra.ix;
    const float y1 = (loy - ra.pad - o.y) * ra.iy, y2 = (hiy + ra.pad - o.y) * ra.iy;
    const float z1 = (loz - ra.pad - o.z) * ra.iz, z2 = (hiz + ra.pad - o.z) * ra.iz;
    const float tn = fmaxf(fmaxf(fminf(x1, x2), fminf(y1, y2)), fminf(z1, z2));
    const float tf = fminf(fminf(fmaxf(x1, x2), fmaxf(y1, y2)), fmaxf(z1, z2));
    const float slack = 1e-5f * fabsf(tf) + 1e-30f;
    if (entry) *entry = tn - 1e-5f * fabsf(tn) - 1e-30f;  // a lower bound of the parameter at which the ray enters the padded box (NaN: no bound)
    return !(tn > tf + slack) && !(tf < -slack);  // NaN comparisons are false -> treated as a hit
}

// FP32 pre-test of Tree.Intersect's box test (Tree.cs:36-41): false only when the ray misses the mesh box by far more
// than any rounding, in which case the FP64 Box.Intersect gives tmax < tmin or tmax <= 0 as well (NoHit).
PT_D bool tree_box_maybe_hit(const ptgpu_tree& t, V3 o, const RayAux& ra) {
    const float ext = fmaxf(fmaxf(t.bmax[0] - t.bmin[0], t.bmax[1] - t.bmin[1]), t.bmax[2] - t.bmin[2]);
    const float mag = fmaxf(fmaxf(fmaxf(fabsf(t.bmin[0]), fabsf(t.bmax[0])), fmaxf(fabsf(t.bmin[1]), fabsf(t.bmax[1]))), fmaxf(fabsf(t.bmin[2]), fabsf(t.bmax[2])));
    const float p = 1e-4f * ext + 1e-5f * mag + 1e-7f;
    return box_line_hit(t.bmin[0] - p, t.bmin[1] - p, t.bmin[2] - p, t.bmax[0] + p, t.bmax[1] + p, t.bmax[2] + p, o, ra);
}

PT_D uint4 stk_entry(double ts, uint32_t node, uint32_t culled) { return make_uint4((uint32_t)__double2loint(ts), (uint32_t)__double2hiint(ts), node, culled); }

// Where the 16-byte stack entries live.  PtrStack: a plain array (local memory, or the per-ray global array of the scene level).
struct PtrStack {
    uint4* p;
    PT_D uint4 get(int i) const { return p[i]; }
    PT_D void put(int i, const uint4& v) { p[i] = v; }
    PT_D void reset() {}
    PT_D void shrink(int) {}
};
// Resume the nearest pending far child that can still hold a closer hit.  False = traversal finished.
template <class Stk>
PT_D bool mesh_pop_t(KdCursor& c, double bestT, Stk& stk) {
    while (c.sp > 0) {
        const uint4 e = stk.get(c.sp);
        --c.sp;
        const double ts = stk_t(e);
        if (bestT <= ts) continue;  // `if (h1.T <= tsplit) return h1`
        if (e.w) continue;          // its Node.Intersect returns NoHit
        c.node = e.z;
        c.tmin = ts;
        c.tmax = netmin_best(stk_t(stk.get(c.sp)), bestT);
        stk.shrink(c.sp);
        return true;
    }
    stk.shrink(0);
    return false;
}
PT_D bool mesh_pop(KdCursor& c, double bestT, uint4* stk) { PtrStack ps{stk}; return mesh_pop_t(c, bestT, ps); }

// Node.Intersect step on Scene.tree (16-byte reference nodes, no culling) with the 16-byte stack.
PT_D int scene_step(const ptgpu_node* __restrict__ nodes, KdCursor& c, V3 o, V3 d, uint4* stk, int stackEnt, uint32_t& leafFirst, uint32_t& leafCount) {
    const int4 raw = __ldg(reinterpret_cast<const int4*>(nodes + c.node));
    const uint32_t a = (uint32_t)raw.z, b = (uint32_t)raw.w;
    const uint32_t axis = a & 3u;
    if (axis == 0) { leafFirst = a >> 2; leafCount = b; return KD_LEAF; }
    const double split = __hiloint2double(raw.y, raw.x);
    const double oa = (double)vaxis(o, axis), da = (double)vaxis(d, axis);
    const double tsplit = (split - oa) / da;
    const bool leftFirst = (oa < split) || (oa == split && da <= 0);
    const uint32_t first = leftFirst ? (a >> 2) : b;
    const uint32_t second = leftFirst ? b : (a >> 2);
    if (tsplit > c.tmax || tsplit <= 0) c.node = first;
    else if (tsplit < c.tmin) c.node = second;
    else {
        if (c.sp + 1 < stackEnt) {
            // entry 0 (the sentinel with the root tmax, see mesh_pop) is written with the first push above it: until then c.tmax IS the
            // root tmax (only this branch changes it), and after a pop back to level 0 it is netmin(root tmax, best.T), which gives the
            // same netmin(.., best.T) at every later pop because best.T only decreases.  Rays that never push never touch their stack.
            if (c.sp == 0) stk_put(stk, c.tmax, 0u, 0u);
            c.sp++; stk_put(stk + c.sp, tsplit, second, 0u);
        }
        c.node = first;
        c.tmax = tsplit;
    }
    return KD_INTERIOR;
}

#ifdef PT_DEBUG_STEPS
__device__ unsigned long long g_dbg[8];  // 0 items, 1 real steps, 2 virtual steps, 3 leaf visits, 4 triangle tests, 5 pops, 6 pushes
#define DBG_ADD(i, v) atomicAdd(&g_dbg[i], (unsigned long long)(v))
#else
#define DBG_ADD(i, v)
#endif
enum { MESH_INTERIOR = 0, MESH_LEAF = 1, MESH_DONE = 2 };
// record layout constants (kNodeVirtual, kNodeRefLeaf, kRefLeaf, leaf_ref, kVirtualDepthMax ...): mesh_derive.hpp
static constexpr int kMeshStackEnt = kMeshStack + kVirtualDepthMax + 1;

// One step at c.node (whose bounds are known to be hit).  Node records (4 x uint4, q0 = {split, a, b}):
//   reference interior  a = left << 2 | axis (1..3), b = right                  -> Node.Intersect (Tree.cs:67-113)
//   bounds-only node    a = left << 2, b = kNodeVirtual | right [| kNodeRefLeaf]  -> both children hold triangles of ONE
//                       reference leaf; every child whose padded bounds the ray line hits is visited, in any order
//   micro leaf          no record: named in the parent's child reference (leaf_ref)
// kNodeRefLeaf / kRefLeafRoot mark the root of a reference leaf (where the tie-break position restarts).
// MESH_LEAF: triangles [tFirst, tFirst + tCount).
template <class Stk>
PT_D int mesh_step_t(const uint4* __restrict__ nodes, const RayBox& ra, KdCursor& c, V3 o, V3 d, Stk& stk, double bestT, uint32_t& bestPos, uint32_t& tFirst,
                     uint32_t& tCount) {
    if (c.node & kRefLeaf) {
        if (c.node & kRefLeafRoot) bestPos = 0;
        tFirst = c.node & kRefFirstMask; tCount = ((c.node >> 26) & 3u) + 1u;
        return MESH_LEAF;
    }
    const uint4* np = nodes + (size_t)c.node * 4;
    // all four quads up front: behind the leaf test the bounds would cost a second, serialised memory round trip
    const uint4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2), q3 = __ldg(np + 3);
    const uint32_t a = q0.z, b = q0.w;
    const uint32_t axis = a & 3u;
    if (axis == 0 && (b & kNodeRefLeaf)) bestPos = 0;
    DBG_ADD(axis == 0 ? 2 : 1, 1);
    float tnL, tnR;
    bool hitL = box_line_hit_fast(__uint_as_float(q1.x), __uint_as_float(q1.y), __uint_as_float(q1.z), __uint_as_float(q1.w), __uint_as_float(q2.x),
                                  __uint_as_float(q2.y), ra, tnL);
    bool hitR = box_line_hit_fast(__uint_as_float(q2.z), __uint_as_float(q2.w), __uint_as_float(q3.x), __uint_as_float(q3.y), __uint_as_float(q3.z),
                                  __uint_as_float(q3.w), ra, tnR);
    const uint32_t left = a >> 2, right = b & kNodeIndexMask;
    bool go;
    if (axis == 0) {
        // bounds-only node.  tn* are strict lower bounds of the T of any triangle below the child (the padded box contains
        // the triangles with a margin far above the FP32 error of the triangle test), so a child with best.T <= tn cannot
        // improve or tie the running best: it is dropped here, or when it is popped (its tn is stored as the entry's
        // `tsplit`, and mesh_pop skips entries with best.T <= tsplit).  Near child first.
        hitL = hitL && !(bestT <= (double)tnL);
        hitR = hitR && !(bestT <= (double)tnR);
        const bool leftNear = !(tnR < tnL);
        const uint32_t nearC = leftNear ? left : right, farC = leftNear ? right : left;
        const bool hitNear = leftNear ? hitL : hitR, hitFar = leftNear ? hitR : hitL;
        if (hitNear && hitFar) { c.sp++; stk.put(c.sp, stk_entry((double)(leftNear ? tnR : tnL), farC, 0u)); }
        c.node = hitNear ? nearC : farC;
        go = hitNear || hitFar;
    } else {
        const double split = __hiloint2double((int)q0.y, (int)q0.x);
        const double oa = (double)vaxis(o, axis), da = (double)vaxis(d, axis);
        const double tsplit = (split - oa) / da;
        const bool leftFirst = (oa < split) || (oa == split && da <= 0);
        const uint32_t first = leftFirst ? left : right, second = leftFirst ? right : left;
        const bool hitFirst = leftFirst ? hitL : hitR, hitSecond = leftFirst ? hitR : hitL;
        if (tsplit > c.tmax || tsplit <= 0) { c.node = first; go = hitFirst; }
        else if (tsplit < c.tmin) { c.node = second; go = hitSecond; }
        else if (!hitFirst) {
            // the near child returns NoHit: what the pop of (second, tsplit) would do, without the stack round trip
            if (bestT <= tsplit || !hitSecond) go = false;
            else { c.node = second; c.tmin = tsplit; c.tmax = netmin_best(c.tmax, bestT); go = true; }
        } else {
            c.sp++;
            stk.put(c.sp, stk_entry(tsplit, second, hitSecond ? 0u : 1u));
            c.node = first;
            c.tmax = tsplit;
            go = hitFirst;
        }
    }
    if (go) return MESH_INTERIOR;
    return mesh_pop_t(c, bestT, stk) ? MESH_INTERIOR : MESH_DONE;
}
PT_D int mesh_step(const uint4* __restrict__ nodes, const RayBox& ra, KdCursor& c, V3 o, V3 d, uint4* stk, double bestT, uint32_t& bestPos, uint32_t& tFirst,
                   uint32_t& tCount) {
    PtrStack ps{stk};
    return mesh_step_t(nodes, ra, c, o, d, ps, bestT, bestPos, tFirst, tCount);
}

// The triangles [tPos, tEnd) of a micro leaf, at most `budget` of them.  Triangles are stored in sorted order, and the
// reference keeps the FIRST shape in array order among equal T (Tree.cs:122), hence the position tie-break.
PT_D void leaf_work(const DScene& S, V3 o, V3 d, uint32_t& tPos, uint32_t tEnd, double& best, int32_t& prim, uint32_t& bestPos, int budget) {
#pragma unroll 1
    for (int k = 0; k < budget && tPos < tEnd; k++) {
        const float4* g = S.leafGeom + (size_t)tPos * 3;
        DBG_ADD(4, 1);
        const double t = triangle_intersect(g, o, d);
        if (t <= best && t < kHitInf) {  // rare: fetch the ids only for candidates (a T of INF never replaces NoHit)
            const uint32_t pos = __float_as_uint(__ldg(g + 1).w);
            if (t < best || pos < bestPos) { best = t; prim = (int32_t)__float_as_uint(__ldg(g).w); bestPos = pos; }
        }
        tPos++;
    }
}

// The analytic subset (the split tracer is only used for scenes without SDFShape / Volume).
PT_D double primitive_intersect_analytic(const DScene& S, const ptgpu_shape& sh, V3 o, V3 d) {
    switch (sh.type) {
        case PTGPU_SPHERE: return sphere_intersect(S.spheres[sh.data], o, d);
        case PTGPU_CUBE: return cube_intersect(S.cubes[sh.data], o, d);
        case PTGPU_PLANE: return plane_intersect(S.planes[sh.data], o, d);
        case PTGPU_CYLINDER: return cylinder_intersect(S.cylinders[sh.data], o, d);
        default: return kHitInf;
    }
}

// Shadow rays: Hit.T of the light itself along the ray, i.e. what light.Intersect(ray) returns when Tree.Intersect reaches it
// (the same device function, so the same bits).  Sphere / Cube / Plane only (the class-typed analytic lights); -1 = not
// evaluated up front (SDFShape / Volume lights: no cut-off for their shadow rays).
PT_D double light_hit_t(const DScene& S, int32_t lightShape, V3 o, V3 d) {
    const ptgpu_shape sh = S.shapes[lightShape];
    switch (sh.type) {
        case PTGPU_SPHERE: return sphere_intersect(S.spheres[sh.data], o, d);
        case PTGPU_CUBE: return cube_intersect(S.cubes[sh.data], o, d);
        case PTGPU_PLANE: return plane_intersect(S.planes[sh.data], o, d);
        default: return -1.0;
    }
}
// Beyond this parameter a shape's Hit cannot matter to a shadow ray whose light sits at tL (see scene_advance).
PT_D double shadow_clip(double tL) { return tL > 0 ? tL : 1e300; }  // the 1e-4 margin is applied where it is compared

#ifndef PT_LEAF_BURST
#define PT_LEAF_BURST 8
#endif
#ifndef PT_NODE_BURST
#define PT_NODE_BURST 4   // 4: +1.2 % over 8 on C3 (bench.py, 128 spp); 16: -8 %
#endif
#ifndef PT_SPLIT_FETCH_MIN
#define PT_SPLIT_FETCH_MIN 8   // mesh_walk refills idle lanes once this many wait (or nothing else is left to do)
#endif
#ifndef PT_MARCH_FETCH_MIN
#define PT_MARCH_FETCH_MIN 1   // march_items refills idle lanes once this many wait: a refill costs three loads, an idle lane a whole burst (8: C5 +15 % slower)
#endif
#ifndef PT_SDF_BURST
#define PT_SDF_BURST 2    // sphere-tracing steps between two refills of a warp of march_items (a lane that finishes idles to the end of the burst; 16: ncu 8 of 32 lanes per instruction)
#endif
#ifndef PT_VOL_BURST
#define PT_VOL_BURST 8    // Volume marching steps between two refills
#endif
enum { ST_IDLE = 0, ST_SCENE_NODE, ST_SCENE_LEAF, ST_MESH_NODE, ST_MESH_LEAF, ST_MESH_DONE, ST_FINISH, ST_EXIT };

// ---------------------------------------------------------------------------------------------------- split tracer
// Scene.Intersect (Scene.cs:75-79) as two kinds of kernels.
//
// A naive "one thread walks one ray to completion" kernel measured 2.2 active lanes per warp instruction on a 250k-triangle scene
// (per-ray cost is heavy-tailed and a warp runs as long as its slowest lane); round 1's single persistent state machine with ray
// replacement (LEAF / NODE / GLUE / MARCH classes voting per iteration) settled at a third of the lanes per class.  Here the scene
// level - Scene.tree, the analytic shapes, TransformedShape set-up and fold, the FP64 Box.Intersect of every deferred shape - runs
// as a streaming kernel (`scene_advance`, one thread per ray, every lane busy) that carries each ray up to the next shape whose
// Intersect is a loop of its own and emits a 48-byte work item for it:
//   Mesh      -> `mesh_walk`   (k_mesh):      Tree.Intersect of Mesh.tree; a persistent kernel that only knows NODE and LEAF work
//   SDFShape  -> `march_items<SDF>`    (k_march): the sphere-tracing loop of SDF.cs:47-74, one loop body for the whole warp
//   Volume    -> `march_items<VOLUME>` (k_march): the marching loop of Volume.cs:172-196
// The shape's Hit is written into the ray's record and the next `scene_advance` round folds it in and continues with the ray's
// remaining shapes.  A ray takes as many rounds as it enters deferred shapes.  The per-ray arithmetic and visiting order are
// those of Tree.Intersect (Tree.cs:31-128), every loop is reproduced step for step, so hits are bit-identical to the reference.
// Work items of the three kinds share one queue; the kind sits in the top two bits of the item's root word and every consumer
// kernel takes the items of its kind.
static constexpr uint32_t kItemKindShift = 30u, kItemMesh = 0u, kItemSdf = 1u, kItemVolume = 2u, kItemIndexMask = (1u << 30) - 1u;
// Per ray of the launch: what a ray that waits for a deferred shape carries from one scene_advance round to the next.  One 128-byte
// record (four 32-byte sectors): the rays of a RESUME round are a scattered subset of the launch, so their state is gathered, and
// fourteen separate arrays cost fourteen sector reads per ray (ncu on the instanced scene: 250 B of DRAM traffic per resumed ray).
struct alignas(32) RayState {
    double bestT, bestTInner;                                   // running Hit of Scene.tree's traversal
    int32_t bestShape, bestPrim; uint32_t scNode; int32_t scSp; // ... and the Scene.tree cursor
    double scTmin, scTmax;
    uint32_t sPos, sEnd, curShape; int32_t curInst;             // position in the current scene leaf
    double mBest; int32_t mPrim; int32_t pad0;                  // Hit of the pending deferred Intersect (written by k_mesh / k_march)
    uint32_t mask[8];                                           // candidate mask: scene shapes the ray still has to evaluate
};
static_assert(sizeof(RayState) == 128, "RayState is four sectors");
struct SplitState {
    RayState* state;                                                            // [ray], stateQuads x 16 bytes apart
    uint32_t stateQuads;                                                        // 8 with the candidate mask, 6 (three sectors) without it
    uint4* sceneStack; int stackEnt;                                            // [ray][stackEnt], entry 0 = sentinel
    uint4* meshStack;                                                           // PT_MESH_GSTACK: [k_mesh thread][kMeshStackEnt]
    unsigned long long* kindItems;                                              // [3] work items consumed per kind (counters; [0] is filled in by the host)
};
PT_D RayState* ray_state(const SplitState& W, uint32_t ray) { return reinterpret_cast<RayState*>(reinterpret_cast<uint4*>(W.state) + (size_t)ray * W.stateQuads); }
#ifndef PT_BEST_CLIP
#define PT_BEST_CLIP (!PT_NO_CULL)   // scene_advance: no mesh walk beyond the running best of the Scene.tree traversal (see there)
#endif
struct MeshQueue {       // work items: a = (co.xyz, ray)  b = (cd.xyz, kind | root node or shape data index)  c = (tmin, tmax)
    float4* a; float4* b; double2* c; uint32_t* count;
    float* lim;          // shadow launches only: the walk may stop at the first Hit with T < lim (a float at or below the light's tL; <= 0: never)
};

// Mesh.Intersect by one thread, start to end (the tail rounds of scenes where rays enter many meshes: a few thousand rays,
// latency-bound whatever the scheduling, not worth a k_mesh launch each).
PT_D void mesh_walk_single(const DScene& S, V3 co, V3 cd, uint32_t root, double tmin, double tmax, double& best, int32_t& prim, double anyLim = -1.0) {
    const RayBox ra = ray_box(co, cd);
    KdCursor mc; mc.node = root; mc.tmin = tmin; mc.tmax = tmax; mc.sp = 0;
    uint4 stk[kMeshStackEnt];
    stk_put(stk, tmax, 0u, 0u);
    best = kHitInf; prim = -1;
    uint32_t bestPos = 0;
    for (;;) {
        uint32_t first, count;
        const int r = mesh_step(S.meshNodes, ra, mc, co, cd, stk, best, bestPos, first, count);
        if (r == MESH_DONE) break;
        if (r == MESH_LEAF) {
            uint32_t tPos = first;
            leaf_work(S, co, cd, tPos, first + count, best, prim, bestPos, 4);
            if (best < anyLim) break;  // shadow ray: closer than the light (see scene_advance)
            if (!mesh_pop(mc, best, stk)) break;
        }
    }
}

#ifndef PT_SCENE_MASK
#define PT_SCENE_MASK (!PT_NO_CULL)   // candidate mask + mailboxing at the Scene.tree level (see scene_advance)
#endif
static constexpr uint32_t kMaskShapes = 255u;  // scene shapes with a bit of their own; bit 255 stands for all the others
static constexpr int kSceneBlock = 128;        // threads per block of every kernel that runs scene_advance (the mask sits in shared memory, one column per thread)
PT_D void save_ray_state(RayState* p, const HitRec& best, const KdCursor& sc, uint32_t sPos, uint32_t sEnd, uint32_t curShape, int32_t curInst, const uint32_t* mk) {
    uint4* rs = reinterpret_cast<uint4*>(p);  // six 128-bit stores; the fifth quad (mBest, mPrim) belongs to the consumer kernel
#if PT_SCENE_MASK
    if (mk) {
    rs[5] = make_uint4(mk[0], mk[kSceneBlock], mk[2 * kSceneBlock], mk[3 * kSceneBlock]);
    rs[6] = make_uint4(mk[4 * kSceneBlock], mk[5 * kSceneBlock], mk[6 * kSceneBlock], mk[7 * kSceneBlock]);
    }
#endif
    rs[0] = make_uint4((uint32_t)__double2loint(best.t), (uint32_t)__double2hiint(best.t), (uint32_t)__double2loint(best.tInner), (uint32_t)__double2hiint(best.tInner));
    rs[1] = make_uint4((uint32_t)best.shape, (uint32_t)best.prim, sc.node, (uint32_t)sc.sp);
    rs[2] = make_uint4((uint32_t)__double2loint(sc.tmin), (uint32_t)__double2hiint(sc.tmin), (uint32_t)__double2loint(sc.tmax), (uint32_t)__double2hiint(sc.tmax));
    rs[3] = make_uint4(sPos, sEnd, curShape, (uint32_t)curInst);
}
PT_D void save_hit(RayState* p, double t, int32_t prim) {  // one 128-bit store
    reinterpret_cast<uint4*>(p)[4] = make_uint4((uint32_t)__double2loint(t), (uint32_t)__double2hiint(t), (uint32_t)prim, 0u);
}

// Advance rays through Scene.tree until each either finishes (sink) or has to enter a Mesh (work item to `out`).
// MODE 0: rays [0, n) start.  MODE 1: the n rays named by the items of `in` continue after their mesh walk.  MODE 2: the
// same rays, whose mesh walks have NOT run yet, are carried to their end by this thread (mesh walks inline, no more items).
enum { SCENE_START = 0, SCENE_RESUME = 1, SCENE_FINISH = 2 };
// SHADOW = true: the rays are sampleLight's visibility rays (Sampler.cs:261-265), whose only use is `hit.Shape == light`.
// `lightOf(i)` names the light; its own Hit.T along the ray, tL, is evaluated from the ray every round (the light is an analytic
// class-typed shape) and the walk - same order, same arithmetic - stops at the first fold that leaves best.T < tL: the
// running best only decreases, so the closest hit can no longer be the light (exact any-hit cut-off).  A ray that
// misses the light altogether (tL = INF) is Black without a walk.  Meshes entered beyond tL (1e-4 relative margin, far above the
// FP32 error of any T) are not walked: their Hit has T > tL, so it can neither be the light nor hide a closer one; and mesh work
// items carry tL so the walk itself stops at the first triangle hit below it (k_mesh<true>).
// Hit.T of a TransformedShape whose Shape is another TransformedShape (instance.pad[0] != 0), from the T the innermost shape returned.
// Every level re-measures T in ITS caller's space from the transformed hit point (TransformedShape.cs:47-69): level k turns the T of
// level k+1 into |M_k * shapeRay_k.Position(T) - ray_k.Origin|.  tInner = the T the outermost level received - what its own
// `shapeRay.Position(hit.T)` is evaluated with (Hit.Info then takes NormalAt / MaterialAt of the INNERMOST shape at that point of the
// first shape space, as the reference does: hit.Shape is the innermost shape, hit.HitInfo the outermost level's).
static constexpr int kMaxInstanceDepth = 4;
PT_D double nested_fold(const DScene& S, int32_t outer, V3 o, V3 d, double tInnermost, double& tInner) {
    int n = 1;
    for (ptgpu_shape sh = S.shapes[S.instances[outer].shape]; sh.type == PTGPU_TRANSFORMED && n < kMaxInstanceDepth; sh = S.shapes[S.instances[sh.data].shape]) n++;
    double t = tInnermost;
    for (int k = n - 1; k >= 0; k--) {  // level k: its caller's ray is the world ray taken through the k levels above it (recomputed: no arrays, the path is rare)
        V3 ro = o, rd = d;
        int32_t cur = outer;
        for (int j = 0; j < k; j++) {
            const ptgpu_instance& up = S.instances[cur];
            ro = mat_pos(up.inv, ro); rd = mat_dir(up.inv, rd);
            cur = (int32_t)S.shapes[up.shape].data;
        }
        const ptgpu_instance& in = S.instances[cur];
        const V3 so = mat_pos(in.inv, ro), sd = mat_dir(in.inv, rd);
        if (k == 0) tInner = t;
        const V3 position = mat_pos(in.m, ray_at(so, sd, t));
        t = (double)vlenf(vsub(position, ro));
    }
    return t;
}
struct NoLight { PT_D int32_t operator()(uint32_t) const { return -1; } };
// TIER: what the scene holds, so that a scene runs the kernel without the code (and the registers) of shapes it does not have - with
// everything compiled in as runtime branches the analytic-shape scenes C1 / C2 traced 10 % slower and C3 3 %:
//   0  analytic shapes only (Sphere, Cube, Plane, Cylinder)                       1  ... and Meshes added to the Scene directly
//   2  everything but nesting: TransformedShape, SDFShape, Volume, SphericalHarmonic, the candidate mask
//   3  ... and TransformedShapes of TransformedShapes (nested_fold: 3-5 % of the instanced scene's pass when merely compiled in)
enum { TIER_ANALYTIC = 0, TIER_MESH = 1, TIER_FULL = 2, TIER_NESTED = 3 };
template <int MODE, int TIER, bool SHADOW = false, class Source, class Sink, class LightOf = NoLight>
PT_D void scene_advance(const DScene& S, const SplitState& W, uint32_t n, const MeshQueue& in, const MeshQueue& out, Source source, Sink sink, LightOf lightOf = LightOf()) {
    // Candidate mask (PT_SCENE_MASK).  The reference builder puts a shape into every Scene.tree leaf its box overlaps and stops splitting
    // at 85 % overlap, so leaves are large and repeat each other (the 200-instance scene: 501 items in 45 leaves, up to 41 per leaf, the
    // floor cube in all of them): a camera ray ran ~200 instance pre-tests and evaluated the floor in every leaf it crossed (ncu: 23 000
    // instructions per ray at the scene level).  Two facts make most of that a no-op: (1) a ray that misses an instanced mesh's padded
    // world bounds gets NoHit from it wherever it meets it (the pre-test below); (2) IShape.Intersect depends on the ray and the shape
    // only, so a second evaluation returns the Hit the first one returned, and the fold `h.T < best.T` (Tree.cs:121-125) cannot take it
    // again - best.T only decreases (also true where the first evaluation was cut short by PT_BEST_CLIP or by a shadow ray's light).
    // So each ray carries one bit per scene shape (the first 255; bit 255 = "the others", never cleared): set at the start for every
    // shape but the instanced meshes, whose bits come from testing the ray against their bounds, 16 instances per block with the
    // block's union first; cleared when the shape is evaluated.  A leaf none of whose shapes has its bit set is skipped as a whole
    // (sceneLeafMask), otherwise its items are visited in array order as before and those without a bit are passed over.  What is
    // evaluated, in which order, and every fold that can change best are those of the reference.
    constexpr bool MASK = TIER >= TIER_FULL;
    __shared__ uint32_t maskColumns[MASK ? 8 * kSceneBlock : 1];
    uint32_t* const mk = (MASK && PT_SCENE_MASK && S.maskOn) ? maskColumns + threadIdx.x : nullptr;
    const ptgpu_tree sceneTree = S.trees[S.sceneTree];
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        constexpr bool RESUME = MODE != SCENE_START;
        const uint32_t ray = RESUME ? __float_as_uint(in.a[k].w) : k;
        V3 o, d;
        source(ray, o, d);
        HitRec best;
        KdCursor sc;
        uint32_t sPos = 0, sEnd = 0, curShape = 0;
        int32_t curInst = -1, mPrim = -1;
        double mBest = kHitInf;
        V3 co = o, cd = d;
        uint4* sstk = W.sceneStack + (size_t)ray * W.stackEnt;
        const RayAux worldAux = ray_aux(o, d);
        const float dirLen = vlenf(d);  // 1 for every ray the renderer makes; ptgpu_intersect_batch takes directions as given
        const double tL = SHADOW ? light_hit_t(S, lightOf(ray), o, d) : -1.0;
        const double clipL = shadow_clip(tL);
        int st;
        if (SHADOW && MODE == SCENE_START && !(tL < kHitInf)) {  // the ray misses the light: hit.Shape != light whatever is hit
            best.t = kHitInf; best.tInner = 0; best.shape = -1; best.prim = -1;
            sink(ray, best);
            continue;
        }
        if (RESUME) {
            const uint4* rs = reinterpret_cast<const uint4*>(ray_state(W, ray));  // seven 128-bit loads from four consecutive sectors
            const uint4 r0 = rs[0], r1 = rs[1], r2 = rs[2], r3 = rs[3], r4 = rs[4];
#if PT_SCENE_MASK
            if (mk) {
                const uint4 r5 = rs[5], r6 = rs[6];
                mk[0] = r5.x; mk[kSceneBlock] = r5.y; mk[2 * kSceneBlock] = r5.z; mk[3 * kSceneBlock] = r5.w;
                mk[4 * kSceneBlock] = r6.x; mk[5 * kSceneBlock] = r6.y; mk[6 * kSceneBlock] = r6.z; mk[7 * kSceneBlock] = r6.w;
            }
#endif
            best.t = __hiloint2double((int)r0.y, (int)r0.x); best.tInner = __hiloint2double((int)r0.w, (int)r0.z);
            best.shape = (int32_t)r1.x; best.prim = (int32_t)r1.y; sc.node = r1.z; sc.sp = (int)r1.w;
            sc.tmin = __hiloint2double((int)r2.y, (int)r2.x); sc.tmax = __hiloint2double((int)r2.w, (int)r2.z);
            sPos = r3.x; sEnd = r3.y; curShape = r3.z; curInst = (int32_t)r3.w;
            const float4 ia = in.a[k], ib = in.b[k];  // the item holds the ray in the shape's space (TransformedShape.cs:45): no second Matrix.MulRay
            if (curInst >= 0) { co = v3(ia.x, ia.y, ia.z); cd = v3(ib.x, ib.y, ib.z); }
            if (MODE == SCENE_FINISH) {  // the pending walk / march first
                const double2 ic = in.c[k];
                const uint32_t tag = __float_as_uint(ib.w);
                if ((tag >> kItemKindShift) == kItemMesh)
                    mesh_walk_single(S, v3(ia.x, ia.y, ia.z), v3(ib.x, ib.y, ib.z), tag, ic.x, ic.y, mBest, mPrim, SHADOW ? (double)in.lim[k] : -1.0);
                else {  // a pending SDFShape / Volume march: the whole Intersect inline (same prologue, same loop)
                    ptgpu_shape msh;
                    msh.type = (tag >> kItemKindShift) == kItemSdf ? PTGPU_SDF : PTGPU_VOLUME; msh.data = tag & kItemIndexMask; msh.material = -1; msh.flags = 0;
                    mBest = primitive_intersect(S, msh, v3(ia.x, ia.y, ia.z), v3(ib.x, ib.y, ib.z));
                    mPrim = -1;
                }
            } else { mBest = __hiloint2double((int)r4.y, (int)r4.x); mPrim = (int32_t)r4.z; }
            st = ST_MESH_DONE;
        } else {
            best.t = kHitInf; best.tInner = 0; best.shape = -1; best.prim = -1;
            box_intersect(sceneTree.bmin, sceneTree.bmax, o, d, sc.tmin, sc.tmax);  // Tree.cs:36-41
            if (sc.tmax < sc.tmin || sc.tmax <= 0) st = ST_FINISH;
            else {
                sc.node = sceneTree.root; sc.sp = 0; st = ST_SCENE_NODE;  // sentinel: written by scene_step with the first push
#if PT_SCENE_MASK
                if (mk) {
#pragma unroll
                for (int w = 0; w < 8; w++) mk[w * kSceneBlock] = S.maskBase[w];
                const RayBox worldBox = ray_box(o, d);
                float tnear;
                for (uint32_t b = 0; b < S.numCandBlocks; b++) {  // the instanced meshes whose padded world bounds the ray line meets
                    const float4 blo = __ldg(S.candBlocks + 2 * b), bhi = __ldg(S.candBlocks + 2 * b + 1);
                    if (!box_line_hit_fast(blo.x, blo.y, blo.z, bhi.x, bhi.y, bhi.z, worldBox, tnear)) continue;
                    const uint32_t mFirst = __float_as_uint(blo.w), mEnd = mFirst + __float_as_uint(bhi.w);
                    for (uint32_t m = mFirst; m < mEnd; m++) {
                        const float4 lo = __ldg(S.candMembers + 2 * m), hi = __ldg(S.candMembers + 2 * m + 1);
                        if (box_line_hit_fast(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, worldBox, tnear)) {
                            const uint32_t shp = __float_as_uint(lo.w);
                            mk[(shp >> 5) * kSceneBlock] |= 1u << (shp & 31u);
                        }
                    }
                }
                }
#endif
            }
        }
        for (;;) {
            if (st == ST_MESH_DONE) {  // fold the shape's Hit into the leaf's running best (Tree.cs:121-125)
                double t = mBest, tInner = 0;
                if (TIER >= TIER_FULL && curInst >= 0) {
                    tInner = mBest;
                    if (mBest < kHitInf) {  // TransformedShape.cs:47-69
                        const ptgpu_instance& inst = S.instances[curInst];
                        if (TIER == TIER_NESTED && inst.pad[0]) t = nested_fold(S, curInst, o, d, mBest, tInner);
                        else {
                            V3 position = mat_pos(inst.m, ray_at(co, cd, mBest));
                            t = (double)vlenf(vsub(position, o));
                        }
                    }
                }
                if (TIER >= TIER_FULL && (curShape >> 31)) { curShape &= 0x7FFFFFFFu; mPrim = -1; }  // SphericalHarmonic: the Hit names the solid, not the triangle (SH.cs:54)
                if (t < best.t) { best.t = t; best.tInner = tInner; best.shape = (int32_t)curShape; best.prim = mPrim; }
                st = (SHADOW && best.t < tL) ? ST_FINISH : ST_SCENE_LEAF;  // occluded: the closest hit is closer than the light
            }
            if (st == ST_SCENE_NODE) {
                uint32_t first, count;
                while (scene_step(S.nodes, sc, o, d, sstk, W.stackEnt, first, count) != KD_LEAF) {}
#if PT_SCENE_MASK
                if (mk) {   // a leaf without a shape this ray still has to evaluate: nothing in it can change best
                    const uint4 l0 = __ldg(S.sceneLeafMask + 2 * (size_t)(sc.node - sceneTree.root)), l1 = __ldg(S.sceneLeafMask + 2 * (size_t)(sc.node - sceneTree.root) + 1);
                    const uint32_t live = (l0.x & mk[0]) | (l0.y & mk[kSceneBlock]) | (l0.z & mk[2 * kSceneBlock]) | (l0.w & mk[3 * kSceneBlock]) |
                                          (l1.x & mk[4 * kSceneBlock]) | (l1.y & mk[5 * kSceneBlock]) | (l1.z & mk[6 * kSceneBlock]) | (l1.w & mk[7 * kSceneBlock]);
                    if (!live) count = 0;
                }
#endif
                sPos = first; sEnd = first + count; st = ST_SCENE_LEAF;
            }
            if (st == ST_SCENE_LEAF) {
                if (sPos == sEnd) st = mesh_pop(sc, best.t, sstk) ? ST_SCENE_NODE : ST_FINISH;
                else {  // next shape of the leaf, in array order (Tree.cs:119-126)
                    curShape = __ldg(S.leafItems + sPos);
                    sPos++;
#if PT_SCENE_MASK
                    if (mk && curShape < kMaskShapes) {  // not a candidate, or evaluated in an earlier leaf: its Hit cannot change best
                        const uint32_t word = mk[(curShape >> 5) * kSceneBlock], bit = 1u << (curShape & 31u);
                        if (!(word & bit)) continue;
                        mk[(curShape >> 5) * kSceneBlock] = word & ~bit;
                    }
#endif
                    ptgpu_shape sh = S.shapes[curShape];
                    curInst = -1; co = o; cd = d;
                    mPrim = -1;
                    if (TIER >= TIER_FULL && sh.type == PTGPU_TRANSFORMED) {  // TransformedShape.cs:45
                        curInst = (int32_t)sh.data;
                        // FP32 pre-test in world space: a ray that misses the padded world bounds of the instanced mesh gives a
                        // shapeRay that misses the mesh's Box (Tree.cs:36-41) -> NoHit, without the FP64 transform
                        const float4 wlo = __ldg(S.instBounds + 2 * (size_t)sh.data), whi = __ldg(S.instBounds + 2 * (size_t)sh.data + 1);
#if PT_BEST_CLIP
                        // ... and an instance entered beyond the running best cannot change it: its Hit's T is the world-space distance to a
                        // point inside these bounds (TransformedShape.cs:69: parameter x |direction|), and the fold only takes T < best.T.
                        // That reads T as a distance, which only holds while the FP32 dot products behind every T are accurate: for an
                        // origin millions of scene sizes away they are not (the no-cull arbiter found a reference hit 8 000 units in front
                        // of its own mesh at |o| = 5e6), so the comparison is only made for origins within ~250 extents of the instance.
                        float entry;
                        const bool hitBounds = box_line_hit(wlo.x, wlo.y, wlo.z, whi.x, whi.y, whi.z, o, worldAux, &entry);
                        const bool nearOrigin = worldAux.pad <= 1e-3f * fmaxf(fmaxf(whi.x - wlo.x, whi.y - wlo.y), whi.z - wlo.z);
                        if (!hitBounds || (nearOrigin && (double)(entry * dirLen) > netmin_best(clipL, best.t) * (1.0 + 1e-4))) { mBest = kHitInf; st = ST_MESH_DONE; continue; }
#else
                        if (!box_line_hit(wlo.x, wlo.y, wlo.z, whi.x, whi.y, whi.z, o, worldAux)) { mBest = kHitInf; st = ST_MESH_DONE; continue; }
#endif
                        const ptgpu_instance& inst = S.instances[sh.data];
                        co = mat_pos(inst.inv, o); cd = mat_dir(inst.inv, d);
                        sh = S.shapes[inst.shape];
                        if (TIER == TIER_NESTED && inst.pad[0]) {  // flagged by the flattener: a TransformedShape of a TransformedShape - its Intersect runs the inner one on ITS shapeRay
                            while (sh.type == PTGPU_TRANSFORMED) {
                                const ptgpu_instance& in2 = S.instances[sh.data];
                                co = mat_pos(in2.inv, co); cd = mat_dir(in2.inv, cd);
                                sh = S.shapes[in2.shape];
                            }
                        }
                    }
                    if (TIER >= TIER_MESH && (sh.type == PTGPU_MESH || (TIER >= TIER_FULL && sh.type == PTGPU_SH))) {  // Mesh.Intersect -> its own Tree.Intersect, starting from NoHit
                        // (SphericalHarmonic.Intersect is mesh.Intersect with the Hit renamed to the solid itself, SH.cs:47-55)
                        const ptgpu_tree mt = S.trees[S.meshes[sh.type == PTGPU_SH ? S.shs[sh.data].mesh : sh.data].tree];
                        mBest = kHitInf;
                        double tmin = 0, tmax = -1;
                        const RayAux ra = ray_aux(co, cd);
                        if (tree_box_maybe_hit(mt, co, ra)) box_intersect(mt.bmin, mt.bmax, co, cd, tmin, tmax);
#if PT_BEST_CLIP
                        // The fold below only takes a mesh Hit with T < best.T (Tree.cs:121-125).  Every triangle lies inside the tree's
                        // box, so a box entered beyond best.T cannot give one: no walk.  Otherwise the walk's tmax is cut to best.T: it
                        // then visits a prefix of the cells the reference visits, in the same order, and a Hit below best.T lies in that
                        // prefix - so the Hit that gets folded is the same.  (1e-4 relative margin against the FP32 rounding of
                        // the triangle test's T, and only for origins within ~250 extents of the mesh - see the instance test above;
                        // object-space T of an instance is not comparable with best.T, so instances are left alone.)
                        const bool nearOrigin = ra.pad <= 1e-3f * fmaxf(fmaxf(mt.bmax[0] - mt.bmin[0], mt.bmax[1] - mt.bmin[1]), mt.bmax[2] - mt.bmin[2]);
                        if (curInst < 0 && nearOrigin && !(tmax < tmin || tmax <= 0)) {
                            const double lim = netmin_best(clipL, best.t) * (1.0 + 1e-4);
                            if (tmin > lim) tmax = -1;
                            else if (tmax > lim) tmax = lim;
                        }
#endif
                        if (sh.type == PTGPU_SH) curShape |= 0x80000000u;  // the fold drops the triangle index
                        if (tmax < tmin || tmax <= 0) st = ST_MESH_DONE;
                        else if (MODE == SCENE_FINISH) { mesh_walk_single(S, co, cd, mt.root, tmin, tmax, mBest, mPrim, (SHADOW && curInst < 0) ? tL : -1.0); st = ST_MESH_DONE; }
                        else {
                            auto g = cooperative_groups::coalesced_threads();
                            uint32_t base = 0;
                            if (g.thread_rank() == 0) base = atomicAdd(out.count, g.size());
                            const uint32_t slot = g.shfl(base, 0) + g.thread_rank();
                            out.a[slot] = make_float4(co.x, co.y, co.z, __uint_as_float(ray));
                            out.b[slot] = make_float4(cd.x, cd.y, cd.z, __uint_as_float(mt.root));
                            out.c[slot] = make_double2(tmin, tmax);
                            // a float at or below tL (object-space T of an instance is not comparable with tL: no cut-off there)
                            if (SHADOW) out.lim[slot] = (curInst < 0 && tL > 0) ? __double2float_rd(tL) : -1.0f;
                            save_ray_state(ray_state(W, ray), best, sc, sPos, sEnd, curShape, curInst, mk);
                            break;
                        }
                    } else if (TIER >= TIER_FULL && (sh.type == PTGPU_SDF || sh.type == PTGPU_VOLUME)) {
                        // SDFShape.Intersect / Volume.Intersect: the prologue here (SDF.cs:34-46, Volume.cs:171-175), the loop in march_items
                        double t0, t1;
                        bool go;
                        if (sh.type == PTGPU_SDF) {
                            const ptgpu_sdf_shape& q = S.sdfShapes[sh.data];
                            double b1, b2;
                            box_intersect(q.bmin, q.bmax, co, cd, b1, b2);
                            go = !(b2 < b1 || b2 < 0);
                            t0 = netmax((double)0.0001f, b1); t1 = b2;
                        } else {
                            const ptgpu_volume& q = S.volumes[sh.data];
                            double b1, b2;
                            box_intersect(q.bmin, q.bmax, co, cd, b1, b2);
                            t0 = netmax((double)(1.0f / 512.0f), b1); t1 = b2;
                            go = t0 <= t1;  // the `t <= tmax` test of the first iteration (false for NaN bounds as well)
                        }
                        mBest = kHitInf;
                        if (!go) st = ST_MESH_DONE;
                        else if (MODE == SCENE_FINISH) { mBest = primitive_intersect(S, sh, co, cd); st = ST_MESH_DONE; }
                        else {
                            auto g = cooperative_groups::coalesced_threads();
                            uint32_t base = 0;
                            if (g.thread_rank() == 0) base = atomicAdd(out.count, g.size());
                            const uint32_t slot = g.shfl(base, 0) + g.thread_rank();
                            const uint32_t kind = sh.type == PTGPU_SDF ? kItemSdf : kItemVolume;
                            out.a[slot] = make_float4(co.x, co.y, co.z, __uint_as_float(ray));
                            out.b[slot] = make_float4(cd.x, cd.y, cd.z, __uint_as_float((kind << kItemKindShift) | sh.data));
                            out.c[slot] = make_double2(t0, t1);
                            if (SHADOW) out.lim[slot] = -1.0f;
                            save_ray_state(ray_state(W, ray), best, sc, sPos, sEnd, curShape, curInst, mk);
                            break;
                        }
                    } else {
                        mBest = primitive_intersect_analytic(S, sh, co, cd);
                        st = ST_MESH_DONE;
                    }
                }
            }
            if (st == ST_FINISH) {
                if (!(best.t < kHitInf)) best.shape = -1;  // Hit.Ok (Hit.cs:22)
                sink(ray, best);
                break;
            }
        }
    }
}

// Mesh.Intersect for every work item of `q`; the Hit goes to W.mBest / W.mPrim of the item's ray.
#ifndef PT_MESH_GSTACK
#define PT_MESH_GSTACK 0
#endif
template <bool ANYHIT>
PT_D void mesh_walk(const DScene& S, const SplitState& W, const MeshQueue& q, uint32_t* __restrict__ cursor) {
    float anyLim = -1.0f;  // ANYHIT (shadow rays): stop at the first Hit below it
    const uint32_t n = *q.count;
    int st = ST_IDLE;
    uint32_t ray = 0;
    V3 co = v3(0, 0, 0), cd = v3(0, 0, 1);
    RayBox ra = ray_box(co, cd);
    KdCursor mc; mc.node = 0; mc.tmin = mc.tmax = 0; mc.sp = 0;
#if PT_MESH_GSTACK
    // the kd stack in GLOBAL memory, one contiguous 16-byte entry per level and thread: a push / pop is one sector, where the interleaved
    // layout of local memory turns it into four 4-byte touches in four lines
    PtrStack mStk{W.meshStack + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * kMeshStackEnt};
#else
    uint4 mLoc[kMeshStackEnt];
    PtrStack mStk{mLoc};
#endif
    uint32_t tPos = 0, tEnd = 0, mBestPos = 0;
    double mBest = kHitInf;
    int32_t mPrim = -1;
#ifdef PT_DEBUG_STEPS
    int dbgSteps = 0, dbgLeaves = 0; uint32_t dbgRoot = 0;
#endif
    for (;;) {
        const unsigned idleMask = __ballot_sync(0xFFFFFFFFu, st == ST_IDLE);
        const unsigned leafMask = __ballot_sync(0xFFFFFFFFu, st == ST_MESH_LEAF);
        const unsigned nodeMask = __ballot_sync(0xFFFFFFFFu, st == ST_MESH_NODE);
        const int nIdle = __popc(idleMask), nLeaf = __popc(leafMask), nNode = __popc(nodeMask);
        if (nIdle + nLeaf + nNode == 0) break;
        if (nIdle >= PT_SPLIT_FETCH_MIN || nLeaf + nNode == 0) {
            if (st == ST_IDLE) {
                auto g = cooperative_groups::coalesced_threads();
                uint32_t base = 0;
                if (g.thread_rank() == 0) base = atomicAdd(cursor, g.size());
                const uint32_t i = g.shfl(base, 0) + g.thread_rank();
                if (i >= n) st = ST_EXIT;
                else {
                    const float4 a = q.a[i], b = q.b[i];
                    if ((__float_as_uint(b.w) >> kItemKindShift) != kItemMesh) continue;  // an SDFShape / Volume item: march_items takes it; this lane stays idle for a turn
                    const double2 c = q.c[i];
                    co = v3(a.x, a.y, a.z); cd = v3(b.x, b.y, b.z); ray = __float_as_uint(a.w);
                    ra = ray_box(co, cd);
                    mc.node = __float_as_uint(b.w); mc.tmin = c.x; mc.tmax = c.y; mc.sp = 0;
                    if (ANYHIT) anyLim = q.lim[i];
#ifdef PT_DEBUG_STEPS
                    dbgRoot = mc.node;
                    DBG_ADD(0, 1);
#endif
                    mStk.reset();
                    mStk.put(0, stk_entry(mc.tmax, 0u, 0u));
                    mBest = kHitInf; mPrim = -1; mBestPos = 0;
                    st = ST_MESH_NODE;
                }
            }
        } else if (nNode >= nLeaf) {  // the class more lanes wait in
            if (st == ST_MESH_NODE) {
#pragma unroll 1
                for (int k = 0; k < PT_NODE_BURST && st == ST_MESH_NODE; k++) {
                    uint32_t first, count;
                    const int r = mesh_step_t(S.meshNodes, ra, mc, co, cd, mStk, mBest, mBestPos, first, count);
#ifdef PT_DEBUG_STEPS
                    dbgSteps++;
#endif
                    if (r == MESH_LEAF) { tPos = first; tEnd = first + count; st = ST_MESH_LEAF; }
                    else if (r == MESH_DONE) st = ST_MESH_DONE;
                }
            }
        } else {
            if (st == ST_MESH_LEAF) {
#ifdef PT_DEBUG_STEPS
                dbgLeaves++;
                DBG_ADD(3, 1);
#endif
                leaf_work(S, co, cd, tPos, tEnd, mBest, mPrim, mBestPos, PT_LEAF_BURST);
                if (ANYHIT && mBest < (double)anyLim) st = ST_MESH_DONE;  // Mesh.Intersect's T can only be <= this: closer than the light
                else if (tPos >= tEnd) st = mesh_pop_t(mc, mBest, mStk) ? ST_MESH_NODE : ST_MESH_DONE;
            }
        }
#ifdef PT_DEBUG_STEPS
        if (st == ST_MESH_DONE) {
            if (dbgSteps + dbgLeaves > 100000) printf("long ray: %d node steps, %d leaves, o=(%g,%g,%g) d=(%g,%g,%g) best=%g root=%u\n", dbgSteps, dbgLeaves, co.x, co.y, co.z, cd.x, cd.y, cd.z, mBest, dbgRoot);
            dbgSteps = dbgLeaves = 0;
        }
#endif
        if (st == ST_MESH_DONE) { save_hit(ray_state(W, ray), mBest, mPrim); st = ST_IDLE; }
    }
}

// SDFShape.Intersect (SDF.cs:32-76) / Volume.Intersect (Volume.cs:169-197) for every work item of kind KIND in `q`; the Hit's T goes
// to W.mBest of the item's ray.  Persistent warps, one item per lane, a finished lane takes the next item of its kind: every lane of
// a warp runs the same loop body (one SDF evaluation, or one Volume.Sign, per step) from its first step to its last, which is what
// the per-iteration class votes of round 1's single kernel could not give (MARCH / NODE / LEAF / GLUE lanes side by side).
// The loops are reproduced step for step - t advances by the reference's own additions - so T is bit-identical.
template <int KIND>
PT_D void march_items(const DScene& S, const SplitState& W, const MeshQueue& q, uint32_t* __restrict__ cursor) {
    const unsigned full = 0xFFFFFFFFu;
    const uint32_t n = *q.count;
    bool have = false, exhausted = false;
    uint32_t ray = 0, data = 0;
    V3 co = v3(0, 0, 0), cd = v3(0, 0, 1);
    double t = 0, tend = 0, step = 0;   // SDF: t, t2, -;  Volume: t, tmax, step
    int it = 0;                         // SDF: iteration counter;  Volume: refinement counter
    uint32_t flags = 0;                 // SDF: bit 0 = `jump`;  Volume: bits 0-15 sign + 1, bit 16 refining, bits 17-31 pending sign + 1
    uint32_t taken = 0;
    for (;;) {
        const unsigned idle = __ballot_sync(full, !have && !exhausted);
        const unsigned busy = __ballot_sync(full, have);
        if (!idle && !busy) {
            for (int o = 16; o; o >>= 1) taken += __shfl_xor_sync(full, taken, o);
            if ((threadIdx.x & 31) == 0 && taken) atomicAdd(W.kindItems + KIND, (unsigned long long)taken);
            break;
        }
        if (__popc(idle) >= PT_MARCH_FETCH_MIN || !busy) {
            while (!have && !exhausted) {
                auto g = cooperative_groups::coalesced_threads();
                uint32_t base = 0;
                if (g.thread_rank() == 0) base = atomicAdd(cursor, g.size());
                const uint32_t i = g.shfl(base, 0) + g.thread_rank();
                if (i >= n) { exhausted = true; break; }
                const float4 b = q.b[i];
                const uint32_t tag = __float_as_uint(b.w);
                if ((tag >> kItemKindShift) != (uint32_t)KIND) continue;  // a Mesh item or the other marcher's
                const float4 a = q.a[i];
                const double2 c = q.c[i];
                co = v3(a.x, a.y, a.z); cd = v3(b.x, b.y, b.z); ray = __float_as_uint(a.w); data = tag & kItemIndexMask;
                t = c.x; tend = c.y; it = 0;
                if (KIND == (int)kItemSdf) flags = 1u;                                 // jump = true
                else { flags = 0u; step = (double)(1.0f / 512.0f); }                   // sign = -1
                have = true; taken++;
            }
        }
        if (have) {
            bool done = false;
            double result = kHitInf;
            if (KIND == (int)kItemSdf) {  // the loop body of SDF.cs:47-74
                const ptgpu_sdf_shape& sh = S.sdfShapes[data];
#pragma unroll 1
                for (int k = 0; k < PT_SDF_BURST && !done; k++) {
                    if (it >= 1000) { done = true; break; }
                    it++;
                    double dist = sdf_evaluate(S.sdfOps + sh.progFirst, sh.progCount, ray_at(co, cd, t));
                    const bool jump = flags & 1u;
                    if (jump && dist < 0) { t -= (double)0.001f; flags = 0u; continue; }
                    if (dist < (double)0.00001f) { result = t; done = true; break; }
                    if (jump && dist < (double)0.001f) dist = (double)0.001f;
                    t += dist;
                    if (t > tend) done = true;
                }
            } else {  // Volume.cs:172-196, one Sign() per step
                const ptgpu_volume& v = S.volumes[data];
#pragma unroll 1
                for (int k = 0; k < PT_VOL_BURST && !done; k++) {
                    const bool refining = (flags >> 16) & 1u;
                    if (!refining) {
                        if (!(t <= tend)) { done = true; break; }  // the `t <= tmax` loop test
                        const int sign = (int)(flags & 0xFFFFu) - 1;
                        if (sign == 1) {  // steps that can only return 1 again (vol_skip)
                            const long long kskip = vol_skip(S, v, data, co, cd, t, step, tend);
                            if (kskip > 0) { t = vol_advance(t, step, kskip); continue; }
                        }
                        const int sg = vol_sign(S, v, ray_at(co, cd, t));
                        if (sg == 0 || (sign >= 0 && sg != sign)) {
                            t -= step; step /= 64; t += step;
                            flags = (flags & 0xFFFFu) | (1u << 16) | ((uint32_t)(sg + 1) << 17);
                            it = 0;
                        } else { flags = (uint32_t)(sg + 1); t += step; }
                    } else if (it < 64) {
                        if (vol_sign(S, v, ray_at(co, cd, t)) == 0) { result = t - step; done = true; break; }
                        t += step; it++;
                    } else {  // the refinement found nothing: `sign = s`, then the outer loop's `t += step` (step stays shrunk)
                        flags = flags >> 17;
                        t += step;
                    }
                }
            }
            if (done) { save_hit(ray_state(W, ray), result, -1); have = false; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------- textures
struct Col { double r, g, b; };
PT_D void modf_net(double in, int& dec, double& frac) { double tr = trunc(in); dec = (int)tr; frac = in - tr; }  // Util.cs:108-113
PT_D double fract_net(double x) { int d; double f; modf_net(x, d, f); return f; }                                 // Texture.cs:218-222
struct Texel { double x, y, z; };
PT_D Texel ld_texel(const double4* p) {  // two 128-bit loads (r, g) (b, -)
    const double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    Texel t; t.x = a.x; t.y = a.y; t.z = b.x; return t;
}
PT_D Col tex_bilinear(const double4* __restrict__ texels, const ptgpu_texture& tx, double u, double v) {  // Texture.cs:188-216
    if (u == 1) u -= kEPS;
    if (v == 1) v -= kEPS;
    double w = (double)tx.width - 1, h = (double)tx.height - 1;
    int X, Y; double x, y;
    modf_net(u * w, X, x);
    modf_net(v * h, Y, y);
    int x0 = X, y0 = Y, x1 = x0 + 1, y1 = y0 + 1;
    const double4* T = texels + tx.texelOffset;
    const Texel c00 = ld_texel(T + (size_t)y0 * tx.width + x0), c01 = ld_texel(T + (size_t)y1 * tx.width + x0);
    const Texel c10 = ld_texel(T + (size_t)y0 * tx.width + x1), c11 = ld_texel(T + (size_t)y1 * tx.width + x1);
    double w00 = (1 - x) * (1 - y), w10 = x * (1 - y), w01 = (1 - x) * y, w11 = x * y;
    Col c = {0, 0, 0};
    c.r = c.r + c00.x * w00; c.g = c.g + c00.y * w00; c.b = c.b + c00.z * w00;
    c.r = c.r + c10.x * w10; c.g = c.g + c10.y * w10; c.b = c.b + c10.z * w10;
    c.r = c.r + c01.x * w01; c.g = c.g + c01.y * w01; c.b = c.b + c01.z * w01;
    c.r = c.r + c11.x * w11; c.g = c.g + c11.y * w11; c.b = c.b + c11.z * w11;
    return c;
}
PT_DC Col tex_sample_c(const ptgpu_texture* __restrict__ textures, const double4* __restrict__ texels, int32_t id, double u, double v) {  // Texture.cs:224-229
    const ptgpu_texture tx = textures[id];
    u = fract_net(fract_net(u) + 1);
    v = fract_net(fract_net(v) + 1);
    return tex_bilinear(texels, tx, u, 1 - v);
}
PT_D Col tex_sample(const DScene& S, int32_t id, double u, double v) { return tex_sample_c(S.textures, S.texels, id, u, v); }
PT_D V3 tex_normal_sample(const DScene& S, int32_t id, double u, double v) {  // Texture.cs:231-237
    Col c = tex_sample(S, id, u, v);
    return vnorm_c(v3d(c.r * 2 - 1, c.g * 2 - 1, c.b * 2 - 1));
}
PT_D int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
PT_D V3 tex_bump_sample(const DScene& S, int32_t id, double u, double v) {  // Texture.cs:239-251 (row read clamped)
    const ptgpu_texture tx = S.textures[id];
    u = fract_net(fract_net(u) + 1);
    v = fract_net(fract_net(v) + 1);
    v = 1 - v;
    int x = (int)(u * tx.width), y = (int)(v * tx.height);
    int x1 = clampi(x - 1, 0, tx.width - 1), x2 = clampi(x + 1, 0, tx.width - 1);
    int y1 = clampi(y - 1, 0, tx.height - 1), y2 = clampi(y + 1, 0, tx.height - 1);
    int yr = clampi(y, 0, tx.height - 1), xr = clampi(x, 0, tx.width - 1);
    const double* T = reinterpret_cast<const double*>(S.texels + tx.texelOffset);  // the red channel of 4 neighbours
    double cx = __ldg(T + 4 * ((size_t)yr * tx.width + x1)) - __ldg(T + 4 * ((size_t)yr * tx.width + x2));
    double cy = __ldg(T + 4 * ((size_t)y1 * tx.width + xr)) - __ldg(T + 4 * ((size_t)y2 * tx.width + xr));
    return v3d(cx, cy, 0.0);
}

// ---------------------------------------------------------------------------------------------------- surfaces
struct Mat {  // Material.cs after MaterialAt
    double cr, cg, cb;
    double emittance, index, gloss, tint, reflectivity;
    int32_t transparent;
    int32_t id;
};
PT_D Mat mat_load(const DScene& S, int32_t id) {
    Mat m;
    if (id < 0) {  // `new Material()` (Mesh.cs:132-135, Volume.cs:150)
        m.cr = m.cg = m.cb = 0; m.emittance = m.index = m.gloss = m.tint = m.reflectivity = 0; m.transparent = 0; m.id = -1;
        return m;
    }
    const ptgpu_material& pm = S.materials[id];
    m.cr = pm.color[0]; m.cg = pm.color[1]; m.cb = pm.color[2];
    m.emittance = pm.emittance; m.index = pm.index; m.gloss = pm.gloss; m.tint = pm.tint; m.reflectivity = pm.reflectivity;
    m.transparent = pm.transparent; m.id = id;
    return m;
}

// Triangle.Barycentric (Triangle.cs:208-223)
PT_D void tri_barycentric(V3 v1, V3 e1, V3 e2, V3 p, double& u, double& v, double& w) {
    V3 v2 = vsub(p, v1);
    double d00 = vdot(e1, e1), d01 = vdot(e1, e2), d11 = vdot(e2, e2), d20 = vdot(v2, e1), d21 = vdot(v2, e2);
    double d = d00 * d11 - d01 * d01;
    v = (d11 * d20 - d01 * d21) / d;
    w = (d00 * d21 - d01 * d20) / d;
    u = 1 - v - w;
}
PT_D V3 tri_blend_uv(const ptgpu_tri_shade& s, double u, double v, double w) {  // T1*u + T2*v + T3*w with Vector rounding
    V3 t1 = v3(s.t1[0], s.t1[1], 0.f), t2 = v3(s.t2[0], s.t2[1], 0.f), t3 = v3(s.t3[0], s.t3[1], 0.f);
    return vadd(vadd(vmuls(t1, u), vmuls(t2, v)), vmuls(t3, w));
}
// Triangle.NormalAt (Triangle.cs:142-189)
// STIER (shade tier, like scene_advance's TIER): 0 analytic shapes without textures, 1 ... and Meshes, still without textures, 2 everything.
template <int STIER>
PT_D V3 tri_normal(const DScene& S, uint32_t tri, V3 p) {
    const float4* g = S.triGeom + (size_t)tri * 3;
    float4 a = __ldg(g), b4 = __ldg(g + 1), c4 = __ldg(g + 2);
    V3 v1 = v3(a.x, a.y, a.z), e1 = v3(b4.x, b4.y, b4.z), e2 = v3(c4.x, c4.y, c4.z);
    const ptgpu_tri_shade& s = S.triShade[tri];
    double u, v, w;
    tri_barycentric(v1, e1, e2, p, u, v, w);
    V3 n = vadd(vadd(vmuls(ld3(s.n1), u), vmuls(ld3(s.n2), v)), vmuls(ld3(s.n3), w));
    if (STIER < 2 || s.material < 0) return vnorm_c(n);  // `new Material()` / a scene without textures
    const ptgpu_material& pm = S.materials[s.material];
    if (pm.normalTexture >= 0) {
        V3 b = tri_blend_uv(s, u, v, w);
        V3 ns = tex_normal_sample(S, pm.normalTexture, (double)b.x, (double)b.y);
        if (!vzero(ns)) {
            V3 dt1 = v3(s.t2[0] - s.t1[0], s.t2[1] - s.t1[1], 0.f), dt2 = v3(s.t3[0] - s.t1[0], s.t3[1] - s.t1[1], 0.f);
            V3 T = vnorm_c(vsub(vmuls(e1, (double)dt2.y), vmuls(e2, (double)dt1.y)));
            V3 B = vnorm_c(vsub(vmuls(e2, (double)dt1.x), vmuls(e1, (double)dt2.x)));
            V3 N = vcross(T, B);
            double X = (double)T.x * (double)ns.x + (double)B.x * (double)ns.y + (double)N.x * (double)ns.z;
            double Y = (double)T.y * (double)ns.x + (double)B.y * (double)ns.y + (double)N.y * (double)ns.z;
            double Z = (double)T.z * (double)ns.x + (double)B.z * (double)ns.y + (double)N.z * (double)ns.z;
            n = vnorm_c(v3d(X, Y, Z));  // Matrix.MulDirection
        }
    }
    if (pm.bumpTexture >= 0) {
        V3 b = tri_blend_uv(s, u, v, w);
        V3 bump = tex_bump_sample(S, pm.bumpTexture, (double)b.x, (double)b.y);
        if (!vzero(bump)) {
            V3 dt1 = v3(s.t2[0] - s.t1[0], s.t2[1] - s.t1[1], 0.f), dt2 = v3(s.t3[0] - s.t1[0], s.t3[1] - s.t1[1], 0.f);
            V3 tangent = vnorm_c(vsub(vmuls(e1, (double)dt2.y), vmuls(e2, (double)dt1.y)));
            V3 bitangent = vnorm_c(vsub(vmuls(e2, (double)dt1.x), vmuls(e1, (double)dt2.x)));
            n = vadd(n, vmuls(tangent, (double)bump.x * pm.bumpMultiplier));
            n = vadd(n, vmuls(bitangent, (double)bump.y * pm.bumpMultiplier));
        }
    }
    return vnorm_c(n);
}

// SphericalHarmonic.Evaluate / NormalAt / MaterialAt (SH.cs:62-98): p.Length() - |Y(p.Normalize())| and its central differences.
PT_D double sh_harmonic(const ptgpu_sh& h, V3 p) { const V3 d = vnorm_c(p); return sh_eval(h.l, h.m, (double)d.x, (double)d.y, (double)d.z); }
PT_D double sh_evaluate(const ptgpu_sh& h, V3 p) { return (double)vlenf(p) - fabs(sh_harmonic(h, p)); }
PT_D V3 sh_normal(const ptgpu_sh& h, V3 p) {
    const double e = 0.0001;
    const double x = p.x, y = p.y, z = p.z;
    const double nx = sh_evaluate(h, v3d(x - e, y, z)) - sh_evaluate(h, v3d(x + e, y, z));
    const double ny = sh_evaluate(h, v3d(x, y - e, z)) - sh_evaluate(h, v3d(x, y + e, z));
    const double nz = sh_evaluate(h, v3d(x, y, z - e)) - sh_evaluate(h, v3d(x, y, z + e));
    return vnorm_c(v3d(nx, ny, nz));
}

// IShape.NormalAt for a non-transformed shape entry (prim = global triangle index for meshes).
template <int STIER>
PT_D V3 shape_normal(const DScene& S, const ptgpu_shape& sh, int32_t prim, V3 p) {
    if (STIER < 1 && sh.type > PTGPU_CYLINDER) return v3(0, 0, 0);
    if (STIER < 2 && sh.type > PTGPU_MESH) return v3(0, 0, 0);
    switch (sh.type) {
        case PTGPU_SPHERE: return vnorm_c(vsub(p, ld3(S.spheres[sh.data].center)));  // Sphere.cs:78-81
        case PTGPU_CUBE: {                                                          // Cube.cs:57-69
            const ptgpu_cube& c = S.cubes[sh.data];
            if (fabs((double)p.x - (double)c.min[0]) < kEPS) return v3(-1, 0, 0);
            if (fabs((double)p.x - (double)c.max[0]) < kEPS) return v3(1, 0, 0);
            if (fabs((double)p.y - (double)c.min[1]) < kEPS) return v3(0, -1, 0);
            if (fabs((double)p.y - (double)c.max[1]) < kEPS) return v3(0, 1, 0);
            if (fabs((double)p.z - (double)c.min[2]) < kEPS) return v3(0, 0, -1);
            if (fabs((double)p.z - (double)c.max[2]) < kEPS) return v3(0, 0, 1);
            return v3(0, 1, 0);
        }
        case PTGPU_PLANE: return ld3(S.planes[sh.data].normal);
        case PTGPU_CYLINDER: {  // Cylinder.cs:122-163
            const ptgpu_cylinder& c = S.cylinders[sh.data];
            const double epsilon = 0.0001;
            if (fabs((double)p.z - c.z0) > epsilon && fabs((double)p.z - c.z1) > epsilon) {
                V3 center = v3d(0, 0, (c.z0 + c.z1) / 2);
                V3 normal = vnorm_c(vsub(p, center));
                if (vdot(normal, vsub(p, v3d(0, 0, c.z0))) < 0) normal = vneg(normal);
                return normal;
            }
            if (fabs((double)p.z - c.z0) < epsilon) return v3(0, 0, -1);
            if (fabs((double)p.z - c.z1) < epsilon) return v3(0, 0, 1);
            return v3(0, 0, 0);
        }
        case PTGPU_MESH: return tri_normal<STIER>(S, (uint32_t)prim, p);  // hit.Shape is the Triangle
        case PTGPU_SDF: return sdf_normal(S, S.sdfShapes[sh.data], p);
        case PTGPU_VOLUME: return volume_normal(S, S.volumes[sh.data], p);
        case PTGPU_SH: return sh_normal(S.shs[sh.data], p);
        default: return v3(0, 0, 0);
    }
}

// Material.MaterialAt(shape, point) (Material.cs:124-138).  UVector is only evaluated when a texture needs it
// (it has no side effects in the reference).
template <int STIER>
PT_D Mat shape_material(const DScene& S, const ptgpu_shape& sh, int32_t prim, V3 p) {
    int32_t id = sh.material;
    if (STIER >= 1 && sh.type == PTGPU_MESH) id = S.triShade[prim].material;
    else if (STIER < 2) {}
    else if (sh.type == PTGPU_VOLUME) id = volume_material(S, S.volumes[sh.data], p);
    else if (sh.type == PTGPU_SH) { const ptgpu_sh& h = S.shs[sh.data]; id = sh_harmonic(h, p) < 0 ? h.negativeMaterial : h.positiveMaterial; }
    Mat m = mat_load(S, id);
    if (id < 0) return m;
    if (STIER < 2) return m;  // a scene without textures
    const ptgpu_material& pm = S.materials[id];
    if (pm.texture >= 0 || pm.glossTexture >= 0) {
        V3 uv = v3(0, 0, 0);
        switch (sh.type) {
            case PTGPU_SPHERE: {  // Sphere.cs:62-69 (sic: (p.X, 0, p.Y))
                V3 q = vsub(p, ld3(S.spheres[sh.data].center));
                double u = atan2_c((double)q.z, (double)q.x);
                double v = atan2_c((double)q.y, (double)vlenf(v3(q.x, 0.f, q.y)));
                u = 1 - (u + kPi) / (2 * kPi);
                v = (v + kPi / 2) / kPi;
                uv = v3d(u, v, 0);
                break;
            }
            case PTGPU_CUBE: {  // Cube.cs:49-53
                const ptgpu_cube& c = S.cubes[sh.data];
                V3 q = vdiv(vsub(p, ld3(c.min)), vsub(ld3(c.max), ld3(c.min)));
                uv = v3(q.x, q.z, 0.f);
                break;
            }
            case PTGPU_CYLINDER: uv = vnorm_c(v3d(-(double)p.y, (double)p.x, 0)); break;  // Cylinder.cs:114-118
            case PTGPU_MESH: {                                                           // Triangle.cs:128-136
                const float4* g = S.triGeom + (size_t)prim * 3;
                float4 a = __ldg(g), b4 = __ldg(g + 1), c4 = __ldg(g + 2);
                double u, v, w;
                tri_barycentric(v3(a.x, a.y, a.z), v3(b4.x, b4.y, b4.z), v3(c4.x, c4.y, c4.z), p, u, v, w);
                const ptgpu_tri_shade& s = S.triShade[prim];
                V3 n = v3(0, 0, 0);
                n = vadd(n, vmuls(v3(s.t1[0], s.t1[1], 0.f), u));
                n = vadd(n, vmuls(v3(s.t2[0], s.t2[1], 0.f), v));
                n = vadd(n, vmuls(v3(s.t3[0], s.t3[1], 0.f), w));
                uv = v3(n.x, n.y, 0.f);
                break;
            }
            default: break;  // Plane / SDFShape / Volume: new Vector()
        }
        if (pm.texture >= 0) { Col c = tex_sample(S, pm.texture, (double)uv.x, (double)uv.y); m.cr = c.r; m.cg = c.g; m.cb = c.b; }
        if (pm.glossTexture >= 0) { Col c = tex_sample(S, pm.glossTexture, (double)uv.x, (double)uv.y); m.gloss = (c.r + c.g + c.b) / 3; }
    }
    return m;
}

struct Surface {  // HitInfo (Hit.cs:58-75)
    V3 position, normal;
    Mat mat;
    bool inside;
};

// Hit.Info (Hit.cs:26-55) / the HitInfo TransformedShape.Intersect pre-fills (TransformedShape.cs:52-70).
template <int STIER>
PT_D Surface hit_info(const DScene& S, V3 o, V3 d, const HitRec& h) {
    Surface sf;
    ptgpu_shape sh = S.shapes[h.shape];
    const bool xf = STIER >= 2 && sh.type == PTGPU_TRANSFORMED;
    const ptgpu_instance* inst = nullptr;
    V3 so = o, sd = d;
    double t = h.t;
    if (xf) {  // the shape's own frame: one copy of shape_normal / shape_material serves both cases
        inst = S.instances + sh.data;
        so = mat_pos(inst->inv, o); sd = mat_dir(inst->inv, d);
        sh = S.shapes[inst->shape];
        if (inst->pad[0]) { while (sh.type == PTGPU_TRANSFORMED) sh = S.shapes[S.instances[sh.data].shape]; }  // nested: hit.Shape is the innermost shape (see nested_fold)
        t = h.tInner;
    }
    const V3 position = ray_at(so, sd, t);
    V3 normal = shape_normal<STIER>(S, sh, h.prim, position);
    sf.mat = shape_material<STIER>(S, sh, h.prim, position);
    sf.inside = false;
    if (xf) {
        sf.position = mat_pos(inst->m, position);
        V3 wn = mat_dir_transposed(inst->inv, normal);
        if (vdot(normal, sd) > 0) { wn = vneg(wn); sf.inside = true; }
        sf.normal = wn;
        return sf;
    }
    if (vdot(normal, d) > 0) {
        normal = vneg(normal);
        sf.inside = true;
        if (STIER >= 2 && (sh.type == PTGPU_VOLUME || sh.type == PTGPU_SDF || sh.type == PTGPU_SH)) sf.inside = false;  // Hit.cs:41-47
    }
    sf.position = position;
    sf.normal = normal;
    return sf;
}

}  // namespace pt
