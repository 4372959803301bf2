// ptgpu.cu — libptgpu: the wavefront path-tracing pipeline and its C ABI (include/ptgpu.h), sm_100a only.
//
// Pipeline per batch of camera samples (replaces Renderer.RenderParallel's pixel/spp loops, Renderer.cs:287-311,
// and DefaultSampler.sample's recursion, Sampler.cs:55-145, with an iterative throughput-weighted form):
//
//   k_raygen        Camera.CastRay (Camera.cs:98-119) + the fu/fv jitter of Renderer.cs:297-304, pixel-major   -> ray queue
//   for depth = 0 .. MaxBounces:
//     Scene.Intersect of every queued ray (the split tracer, pt_device.cuh):
//       k_scene_trace<START>    Scene.tree, analytic shapes, instance set-up; a work item per Mesh / SDFShape / Volume entered
//       rounds of  k_mesh (Mesh.tree walks) / k_march<SDF> / k_march<VOLUME> (the marching loops)  +  k_scene_trace<RESUME>
//                                                                                                   -> hit records
//     k_bin_count / k_bin_scan / k_bin_scatter    shade order: hit records by surface and patch
//     k_shade         Hit.Info, emission/termination, Ray.Bounce, sampleLights ray generation           -> next ray queue,
//                     (queue appends are warp-aggregated: one atomic per coalesced group)                 shadow queue, sum
//     k_scene_shadow<...> + the same rounds: sampleLight's visibility test (Sampler.cs:262-265) with the exact any-hit cut-off
//                                                                                                   -> sum buffer
//   k_add_sample    c /= spp; Buffer.AddSample (Welford, Buffer.cs:33-44); with several devices behind the handle it also sums
//                   their pass accumulators over NVLink peer access
//
// Kernels read their item counts from device memory, so a pass is issued without host synchronisation (scenes whose rays
// may enter more than four deferred shapes poll one queue count per round).  Random numbers: Philox4x32-10 keyed on
// (seed, pass | pixel, sample, path node, sub-stream, draw) — see rng_enter().
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#ifndef PT_SPLIT
#define PT_SPLIT 1
#endif
#ifndef PT_ANYHIT
#define PT_ANYHIT (!PT_NO_CULL)   // 0: shadow rays take the full closest-hit walk (A/B switch for the exact any-hit cut-off, see scene_advance)
#endif
#include "pt_device.cuh"
#include "sh_funcs.hpp"

namespace cg = cooperative_groups;
using namespace pt;

// ====================================================================================================== RNG
struct Rng {
    uint32_t k0, k1, c0, c1, c2, c3;
    uint32_t draw;
    uint32_t cache[4];
    uint32_t cachedBlock;
};
PT_DC uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
PT_D void rng_set_sample(Rng& r, uint32_t seed, uint32_t pass, uint32_t pixel, uint32_t sample) { r.k0 = seed; r.k1 = pass; r.c0 = pixel; r.c1 = sample; }
// Counter word 3: [31:20] first-hit index+1 | [19:14] depth | [13:6] sub-stream | [5:0] block of two draws.
// Word 2: one bit per depth >= 2 telling which BounceType branch the path took under SpecularModeAll.
PT_D void rng_enter(Rng& r, uint32_t pathBits, uint32_t first, uint32_t depth, uint32_t sub) {
    r.c2 = pathBits;
    r.c3 = ((first & 0xFFFu) << 20) | ((depth & 0x3Fu) << 14) | ((sub & 0xFFu) << 6);
    r.draw = 0;
    r.cachedBlock = 0xFFFFFFFFu;
}
PT_D double rng_next(Rng& r) {  // Random.Shared.NextDouble(): 53 random bits / 2^53
    uint32_t block = r.draw >> 1;
    if (block != r.cachedBlock) {
        const uint4 o = philox4x32_10(r.c0, r.c1, r.c2, r.c3 | (block & 0x3Fu), r.k0, r.k1);
        r.cache[0] = o.x; r.cache[1] = o.y; r.cache[2] = o.z; r.cache[3] = o.w;
        r.cachedBlock = block;
    }
    uint32_t hi = (r.draw & 1) ? r.cache[2] : r.cache[0], lo = (r.draw & 1) ? r.cache[3] : r.cache[1];
    r.draw++;
    unsigned long long bits = ((unsigned long long)hi << 32) | lo;
    return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
}

// ====================================================================================================== sampling
// Vector.RandomUnitVector (Vector.cs:339-347)
PT_D V3 random_unit_vector(Rng& rng) {
    double z = rng_next(rng) * 2.0 - 1.0;
    double a = rng_next(rng) * 2.0 * kPi;
    double r = sqrt(1.0 - z * z);
    const double2 sc = sincos_c(a);
    return v3d(r * sc.x, r * sc.y, z);
}
// Util.Cone (Util.cs:17-32)
PT_D V3 cone(V3 direction, double theta, double u, double v, Rng& rng) {
    if (theta < kEPS) return direction;
    theta = theta * (1 - (2 * acos_c(u) / kPi));
    const double2 m = sincos_c(theta);
    const double m1 = m.x, m2 = m.y;
    double a = v * 2 * kPi;
    V3 q = random_unit_vector(rng);
    V3 s = vcross(direction, q);
    V3 t = vcross(direction, s);
    const double2 sca = sincos_c(a);
    const double sa = sca.x, ca = sca.y;
    V3 d = v3(0, 0, 0);
    d = vadd(d, vmuls(s, m1 * ca));
    d = vadd(d, vmuls(t, m1 * sa));
    d = vadd(d, vmuls(direction, m2));
    return vnorm_c(d);
}
// Vector.Reflect (Vector.cs:497): n.Reflect(i) = i - n * (2 * n.i)
PT_D V3 reflect(V3 n, V3 i) { return vsub(i, vmuls(n, 2 * vdot(n, i))); }
// Vector.Refract (Vector.cs:500-514)
PT_D V3 refract(V3 n, V3 i, double n1, double n2) {
    double nr = n1 / n2;
    double cosI = -vdot(n, i);
    double sinT2 = nr * nr * (1 - cosI * cosI);
    if (sinT2 > 1) return v3(0, 0, 0);
    double cosT = sqrt(1 - sinT2);
    return vadd(vmuls(i, nr), vmuls(n, nr * cosI - cosT));
}
// Vector.Reflectance (Vector.cs:517-536)
PT_D double reflectance(V3 n, V3 i, double n1, double n2) {
    double nr2 = (n1 * n1) / (n2 * n2);
    double cosI = -vdot(n, i);
    double sinT2 = nr2 * (1 - cosI * cosI);
    if (sinT2 > 1) return 1;
    double cosT = sqrt(1 - sinT2);
    double a = n1 * cosI, b = n2 * cosT;
    double rOrth = (a - b) / (a + b);
    double rPar = (b - a) / (b + a);
    return (rOrth * rOrth + rPar * rPar) / 2;
}
// Ray.Bounce (Ray.cs:44-85).  mode: BounceType (0 Any, 1 Diffuse, 2 Specular).
PT_D void bounce(V3 rayDir, const Surface& sf, double u, double v, int mode, Rng& rng, V3& outO, V3& outD, bool& reflected, double& pOut) {
    V3 n = sf.normal;
    double n1 = 1.0, n2 = sf.mat.index;
    if (sf.inside) { double t = n1; n1 = n2; n2 = t; }
    double p = sf.mat.reflectivity >= 0 ? sf.mat.reflectivity : reflectance(n, rayDir, n1, n2);
    bool refl = false;
    if (mode == 0) refl = rng_next(rng) < p;
    else if (mode == 2) refl = true;
    if (refl || sf.mat.transparent) {  // one copy of cone() for the reflected and the refracted ray
        V3 base;
        if (refl) { outO = sf.position; base = reflect(n, rayDir); pOut = p; }
        else {
            base = refract(n, rayDir, n1, n2);
            outO = vadd(sf.position, vmuls(base, 1e-4));  // Ray.cs:78, the only epsilon offset in the code base
            pOut = 1 - p;
        }
        outD = cone(base, sf.mat.gloss, u, v, rng);
        reflected = true;
    } else {  // Ray.WeightedBounce (Ray.cs:28-35)
        double radius = sqrt(u);
        double theta = 2 * kPi * v;
        V3 s = vnorm_c(vcross(n, random_unit_vector(rng)));
        V3 t = vcross(n, s);
        const double2 sct = sincos_c(theta);
        const double st = sct.x, ct = sct.y;
        V3 d = v3(0, 0, 0);
        d = vadd(d, vmuls(s, radius * ct));
        d = vadd(d, vmuls(t, radius * st));
        d = vadd(d, vmuls(n, sqrt(1 - u)));
        outO = sf.position;
        outD = d;
        reflected = false;
        pOut = 1 - p;
    }
}

// ====================================================================================================== pipeline state
struct PassD {  // ptgpu_pass plus derived values, passed by value to kernels
    int32_t width, height, spp, stratified, sppRoot;
    int32_t subpixelJitter;  // 1: fu, fv = xi1, xi2 (adaptive / firefly passes, Renderer.cs:351-353, 432); 2: fu = (x + xi) * (1.0f / w) (serial firefly, Renderer.cs:97-98, 179-180)
    int32_t sampleBase, sampleStride;
    int32_t firstHitSamples, maxBounces, directLighting, softShadows, lightMode, specularMode;
    int32_t russianRoulette;  // PTGPU_PASS_RUSSIAN_ROULETTE (opt-in, not part of parity mode)
    uint32_t seed, passIndex;
    ptgpu_camera cam;
};

struct DeviceCounters { unsigned long long cameraSamples, segments, shadowRays, nanSamples; unsigned long long kindItems[3]; };  // kindItems: work items emitted for k_mesh / k_march<SDF> / k_march<VOLUME>

// Ray record: 52 bytes in three float4 streams + one u32 stream (SoA so each stream coalesces).
//   od0 = (o.x, o.y, o.z, bits(pixel))   od1 = (d.x, d.y, d.z, bits(meta))   bt = (beta.r, beta.g, beta.b, bits(pathBits))
//   meta = depth[5:0] | emission[6] | first[18:7]      smp = global sample index
struct RayQueue { float4* od0; float4* od1; float4* bt; uint32_t* smp; };
struct HitQueue { double* t; double* tInner; int32_t* shape; int32_t* prim; };
// Shadow record: 48 bytes.  so = (o.xyz, bits(pixel))  sd = (d.xyz, bits(light shape index))  sc = (contribution rgb, 0)
struct ShadowQueue { float4* so; float4* sd; float4* sc; };

PT_D uint32_t f2u(float f) { return __float_as_uint(f); }
PT_D float u2f(uint32_t u) { return __uint_as_float(u); }

PT_D void accumulate(float* __restrict__ sum, DeviceCounters* cnt, uint32_t pixel, float r, float g, float b) {
    if (!(isfinite(r) && isfinite(g) && isfinite(b))) { atomicAdd(&cnt->nanSamples, 1ull); return; }
    if (r != 0.f) atomicAdd(sum + (size_t)pixel * 3 + 0, r);
    if (g != 0.f) atomicAdd(sum + (size_t)pixel * 3 + 1, g);
    if (b != 0.f) atomicAdd(sum + (size_t)pixel * 3 + 2, b);
}

// Camera.CastRay (Camera.cs:98-119)
PT_D void cast_ray(const ptgpu_camera& cam, int x, int y, int w, int h, double u, double v, Rng& rng, V3& o, V3& d) {
    double aspect = (double)w / (double)h;
    double px = (((double)x + u - 0.5) / ((double)w - 1.0)) * 2 - 1;
    double py = (((double)y + v - 0.5) / ((double)h - 1.0)) * 2 - 1;
    V3 cu = ld3(cam.u), cv = ld3(cam.v), cw = ld3(cam.w), cp = ld3(cam.p);
    V3 dir = v3(0, 0, 0);
    dir = vadd(dir, vmuls(cu, -px * aspect));
    dir = vadd(dir, vmuls(cv, -py));
    dir = vadd(dir, vmuls(cw, cam.m));
    dir = vnorm_c(dir);
    V3 p = cp;
    if (cam.apertureRadius > 0) {
        V3 focalPoint = vadd(cp, vmuls(dir, cam.focalDistance));
        double angle = rng_next(rng) * 2 * kPi;
        double radius = rng_next(rng) * cam.apertureRadius;
        const double2 sca = sincos_c(angle);
        const double sa = sca.x, ca = sca.y;
        p = vadd(p, vmuls(cu, ca * radius));
        p = vadd(p, vmuls(cv, sa * radius));
        dir = vnorm_c(vsub(focalPoint, p));
    }
    o = p;
    d = dir;
}

// ====================================================================================================== kernels
// K1.  Camera samples [g0, g0+n) of the pass, PIXEL-MAJOR: g -> (pixel = g / slots, slot k = g % slots), global sample index
// sampleBase + k*sampleStride.  The samples of one pixel are neighbours in the queue, so a warp traces (nearly) one camera ray
// 32 times - the reference's jitter is 1/w of a pixel (below) - and the lanes walk the same nodes, hit the same triangle and
// run the same shading code; their bounce and shadow rays then leave from one point.  The order changes nothing else: every
// draw is keyed by (pixel, global sample index, place in the path tree).
// Non-stratified: fu = (x+xi1)/w, fv = (y+xi2)/h are passed where CastRay expects a
// sub-pixel offset — the reference's behaviour (Renderer.cs:297-304), reproduced on purpose.
#ifndef PT_PIXEL_MAJOR
#define PT_PIXEL_MAJOR 1   // 0: sample-major (pixel = g % npix), the order of round 1 (A/B switch)
#endif
__global__ void __launch_bounds__(256) k_raygen(PassD P, unsigned long long g0, uint32_t n, uint32_t slots, RayQueue q, uint32_t* __restrict__ count,
                                                 DeviceCounters* cnt, const uint32_t* __restrict__ pixelList, const uint32_t* __restrict__ listCount) {
    const uint32_t npix = (uint32_t)P.width * (uint32_t)P.height;
    if (pixelList) {  // sparse pass: one sample for each listed pixel; this batch covers list[g0, g0+n)
        const uint32_t lc = *listCount;
        n = lc > (uint32_t)g0 ? min(n, lc - (uint32_t)g0) : 0u;
    }
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned long long g = g0 + i;
        uint32_t pixel = PT_PIXEL_MAJOR ? (uint32_t)(g / slots) : (uint32_t)(g % npix), k = PT_PIXEL_MAJOR ? (uint32_t)(g % slots) : (uint32_t)(g / npix);
        if (pixelList) { pixel = pixelList[g0 + i]; k = 0; }
        int x = (int)(pixel % (uint32_t)P.width), y = (int)(pixel / (uint32_t)P.width);
        uint32_t sample = (uint32_t)(P.sampleBase + (int)k * P.sampleStride);
        Rng rng;
        rng_set_sample(rng, P.seed, P.passIndex, pixel, sample);
        rng_enter(rng, 0, 0, 0, 0);
        double fu, fv;
        if (P.stratified) {  // Renderer.cs:231-246: strata centres, no jitter; slot k -> (u, v) = (k / root, k % root)
            uint32_t s = sample % (uint32_t)(P.sppRoot * P.sppRoot);
            fu = ((double)(s / (uint32_t)P.sppRoot) + 0.5) / (double)P.sppRoot;
            fv = ((double)(s % (uint32_t)P.sppRoot) + 0.5) / (double)P.sppRoot;
        } else {
            double xo = rng_next(rng), yo = rng_next(rng);
            if (P.subpixelJitter == 1) { fu = xo; fv = yo; }
            else if (P.subpixelJitter == 2) { fu = ((double)x + xo) * (double)(1.0f / (float)P.width); fv = ((double)y + yo) * (double)(1.0f / (float)P.height); }
            else { fu = ((double)x + xo) / (double)P.width; fv = ((double)y + yo) / (double)P.height; }
        }
        V3 o, d;
        cast_ray(P.cam, x, y, P.width, P.height, fu, fv, rng, o, d);
        q.od0[i] = make_float4(o.x, o.y, o.z, u2f(pixel));
        q.od1[i] = make_float4(d.x, d.y, d.z, u2f(0u | (1u << 6)));  // depth 0, emission = true (Sampler.cs:42)
        q.bt[i] = make_float4(1.f, 1.f, 1.f, u2f(0u));
        q.smp[i] = sample;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { *count = n; atomicAdd(&cnt->cameraSamples, (unsigned long long)n); }
}

// Light geometry used by sampleLight (Sampler.cs:215-236): Sphere -> its centre/radius; Cylinder -> (0,0,(Z0+Z1)/2),
// Radius; anything else -> bounding box centre / outer radius.  Precomputed at upload into DLight.
struct DLight { float center[3]; uint32_t shape; double radius; uint32_t classTyped; uint32_t isCylinder; };

// Shade order.  k_shade runs one thread per hit record, and which code a record needs (Hit.Info of a triangle / cube / sphere,
// the BRDF of its material, NEE or not) follows the surface that was hit: in ray order a warp holds a mix of them (ncu: 9.4 of 32
// lanes per instruction, 41 % of the stall samples `no_instructions`: the divergent warps miss the instruction cache).  The hit
// records of a launch are therefore visited bin by bin (bin = shape type x material of the surface, 0 = miss); inside a bin
// the order stays close to ray order (2048-record chunks, warp-contiguous), so the gathers still touch whole sectors.
// Results do not depend on the order: every draw is keyed by its place in the path tree (rng_enter).
#ifndef PT_SHADE_SUB
#define PT_SHADE_SUB 32   // patches per surface bin (1 = order by surface only).  8-spp C3 pass with 32 surface bins: 1 -> 61.05 ms, 64 -> 60.54, 256 -> 60.9
#endif
#ifndef PT_BIN_INSTANCE_MUL
#define PT_BIN_INSTANCE_MUL 3u   // 0: every instance of a shape in the shape's bin
#endif
#ifndef PT_SURFACE_BINS
#define PT_SURFACE_BINS 128   // x PT_SHADE_SUB patches = 4096 bins (16 KB shared histogram).  C4 2-spp pass: 32 x 64 -> 161.5 ms, 128 x 32 -> 157.8 (instances hashed in: 165 without); C3 unchanged
#endif
static constexpr int kSurfaceBins = PT_SURFACE_BINS, kShadeSub = PT_SHADE_SUB, kShadeBins = kSurfaceBins * kShadeSub, kBinChunk = 2048;
// Bin of a hit record: surface (shape type x material, one per Mesh; 0 = miss) x patch.  The patch of a triangle hit is its index
// within the mesh scaled to kShadeSub (mesh triangles are stored in a spatial order, so a patch is a piece of surface a few
// thousand triangles large): the rays a launch appends then start patch by patch, and the mesh walks and shadow rays of the next
// launches work on one neighbourhood of the kd-tree at a time.
PT_D uint32_t shade_bin(const DScene& S, int32_t shape, int32_t prim) {
    if (shape < 0) return 0u;
    ptgpu_shape sh = S.shapes[shape];
    uint32_t which = 0u, sub = 0u;
    if (sh.type == PTGPU_TRANSFORMED) {
        which = (sh.data + 1u) * PT_BIN_INSTANCE_MUL;
        const ptgpu_instance& inst = S.instances[sh.data];
        sh = S.shapes[inst.shape];
        if (inst.pad[0]) { while (sh.type == PTGPU_TRANSFORMED) sh = S.shapes[S.instances[sh.data].shape]; }
    }  // instances of one mesh sit in different places
    if (sh.type == PTGPU_MESH) {
        which += sh.data * 7u;  // a Mesh carries its materials per triangle: one surface bin per mesh
        if (S.shadeSub > 1 && prim >= 0) {
            const ptgpu_mesh m = S.meshes[sh.data];
            sub = (uint32_t)((float)((uint32_t)prim - m.triFirst) * __fdividef((float)S.shadeSub, (float)(m.triCount ? m.triCount : 1u)));  // any monotone map will do
            if (sub >= S.shadeSub) sub = S.shadeSub - 1;
        }
    }
    // the 4096 bins are split between surfaces and patches per scene (upload_one): few surfaces -> finer patches
    const uint32_t surface = 1u + ((uint32_t)sh.type * 5u + (uint32_t)(sh.material + 1) + which) % (S.shadeSurfaces - 1u);
    return surface * S.shadeSub + sub;
}
// bins[b] += records of bin b
__global__ void __launch_bounds__(256) k_bin_count(DScene S, const int32_t* __restrict__ shape, const int32_t* __restrict__ prim, const uint32_t* __restrict__ count,
                                                  uint32_t* __restrict__ bins) {
    __shared__ uint32_t h[kShadeBins];
    for (int b = threadIdx.x; b < kShadeBins; b += 256) h[b] = 0;
    __syncthreads();
    const uint32_t n = *count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t b = shade_bin(S, shape[i], prim[i]);
        const unsigned m = __match_any_sync(__activemask(), b);
        if ((m & ((1u << (threadIdx.x & 31)) - 1u)) == 0) atomicAdd(&h[b], (uint32_t)__popc(m));
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kShadeBins; b += 256) if (h[b]) atomicAdd(&bins[b], h[b]);
}
// bins[b] = first slot of bin b (exclusive prefix sum), in place; one block of 1024 threads
__global__ void __launch_bounds__(1024) k_bin_scan(uint32_t* __restrict__ bins) {
    constexpr int kPer = (kShadeBins + 1023) / 1024;
    __shared__ uint32_t warpSum[32];
    uint32_t v[kPer], sum = 0;
#pragma unroll
    for (int k = 0; k < kPer; k++) { const int b = threadIdx.x * kPer + k; v[k] = b < kShadeBins ? bins[b] : 0u; sum += v[k]; }
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = sum;
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += t; }
    if (lane == 31) warpSum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warpSum[lane], wi = w;
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, d); if ((int)lane >= d) wi += t; }
        warpSum[lane] = wi - w;
    }
    __syncthreads();
    uint32_t run = warpSum[warp] + incl - sum;
#pragma unroll
    for (int k = 0; k < kPer; k++) { const int b = threadIdx.x * kPer + k; if (b < kShadeBins) bins[b] = run; run += v[k]; }
}
// perm[slot] = record index, bin by bin
__global__ void __launch_bounds__(256) k_bin_scatter(DScene S, const int32_t* __restrict__ shape, const int32_t* __restrict__ prim, const uint32_t* __restrict__ count,
                                                    uint32_t* __restrict__ cursors, uint32_t* __restrict__ perm) {
    __shared__ uint32_t h[kShadeBins];  // records of the chunk per bin, then the chunk's first slot in each bin
    const uint32_t n = *count;
    constexpr int kPer = kBinChunk / 256;
    for (int b = threadIdx.x; b < kShadeBins; b += 256) h[b] = 0;
    __syncthreads();
    for (uint32_t c0 = blockIdx.x * (uint32_t)kBinChunk; c0 < n; c0 += gridDim.x * (uint32_t)kBinChunk) {
        uint32_t bin[kPer], rank[kPer];
#pragma unroll
        for (int k = 0; k < kPer; k++) {
            const uint32_t i = c0 + k * 256 + threadIdx.x;
            bin[k] = i < n ? shade_bin(S, shape[i], prim[i]) : 0xFFFFFFFFu;
            const unsigned m = __match_any_sync(0xFFFFFFFFu, bin[k]);
            const int leader = __ffs(m) - 1;
            uint32_t off = 0;
            if ((int)(threadIdx.x & 31) == leader && i < n) off = atomicAdd(&h[bin[k]], (uint32_t)__popc(m));
            rank[k] = __shfl_sync(0xFFFFFFFFu, off, leader) + __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kPer; k++) {  // the first record of a bin in this chunk (rank 0) reserves the chunk's run in the bin
            const uint32_t i = c0 + k * 256 + threadIdx.x;
            if (i < n && rank[k] == 0) h[bin[k]] = atomicAdd(&cursors[bin[k]], h[bin[k]]);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kPer; k++) {
            const uint32_t i = c0 + k * 256 + threadIdx.x;
            if (i < n) perm[h[bin[k]] + rank[k]] = i;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kPer; k++) {  // leave the table zeroed for the next chunk (only the touched bins)
            const uint32_t i = c0 + k * 256 + threadIdx.x;
            if (i < n && rank[k] == 0) h[bin[k]] = 0;
        }
        __syncthreads();
    }
}

// K3.  One thread per hit record: Hit.Info, emission, then the (u, v, mode) loop of Sampler.cs:97-131.
#ifndef PT_SHADE_MINBLOCKS
#define PT_SHADE_MINBLOCKS 8   // 64 registers: 6.8 ms vs 7.0 ms (4 blocks, 128 registers) per 8-spp C3 pass with the shade order on
#endif
// ncu on the Cornell box (C2), where k_shade is half of the pass: `no_instruction` is its top stall (4.3 warps per issue) - a record is
// ~1 300 straight-line warp instructions spread over a 64 KB+ kernel, and 32 warps that each sit somewhere else in it miss the
// instruction caches for one another.  So the warps of a block start every record together (block-uniform trip count, one barrier
// per record) and the blocks are large: 16 warps then fetch the same lines at about the same time.  Per pass (same gpurun call):
// C2 46.1 -> 43.2 ms, C1 20.8 -> 19.7 ms, C3 57.1 -> 56.1 ms; 128 / 256 / 1024 threads with the barrier: 44.8 / 44.2 / 44.1 ms on C2,
// 512 threads without it 43.6 ms (C3 57.5).
#ifndef PT_SHADE_BLOCK
#define PT_SHADE_BLOCK 512
#endif
#ifndef PT_SHADE_SYNC
#define PT_SHADE_SYNC 1
#endif
template <int STIER>
__global__ void __launch_bounds__(PT_SHADE_BLOCK, PT_SHADE_MINBLOCKS * 128 / PT_SHADE_BLOCK) k_shade(DScene S, PassD P, const DLight* __restrict__ lights, RayQueue q, const uint32_t* __restrict__ count,
                                                HitQueue hq, RayQueue nq, uint32_t* __restrict__ ncount, ShadowQueue sq, uint32_t* __restrict__ scount,
                                                float* __restrict__ sum, DeviceCounters* cnt, uint32_t capRays, uint32_t capShadow,
                                                const uint32_t* __restrict__ perm, uint32_t* __restrict__ overflow) {
    const uint32_t n = *count;
#if PT_SHADE_SYNC
    for (uint32_t j0 = blockIdx.x * blockDim.x; j0 < n; j0 += gridDim.x * blockDim.x) {
        __syncthreads();
        const uint32_t j = j0 + threadIdx.x;
        if (j >= n) continue;
#else
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
#endif
        const uint32_t i = perm ? perm[j] : j;  // shade order (see k_bin_scatter)
        float4 a = q.od0[i], b = q.od1[i], c = q.bt[i];
        const V3 o = v3(a.x, a.y, a.z), d = v3(b.x, b.y, b.z);
        const uint32_t pixel = f2u(a.w), meta = f2u(b.w), pathBits = f2u(c.w), sample = q.smp[i];
        const uint32_t depth = meta & 63u, first = (meta >> 7) & 0xFFFu;
        const bool emission = (meta >> 6) & 1u;
        const float br = c.x, bg = c.y, bb = c.z;
        HitRec h;
        h.t = hq.t[i]; h.tInner = hq.tInner[i]; h.shape = hq.shape[i]; h.prim = hq.prim[i];
        if (h.shape < 0) {  // sampleEnvironment (Sampler.cs:177-189)
            double er = S.envColor[0], eg = S.envColor[1], eb = S.envColor[2];
            if (STIER >= 2 && S.envTexture >= 0) {
                double u = atan2_c((double)d.z, (double)d.x) + S.envTextureAngle;
                double v = atan2_c((double)d.y, (double)vlenf(v3(d.x, 0.f, d.z)));
                u = (u + kPi) / (2 * kPi);
                v = (v + kPi / 2) / kPi;
                Col e = tex_sample(S, S.envTexture, u, v);
                er = e.r; eg = e.g; eb = e.b;
            }
            accumulate(sum, cnt, pixel, br * (float)er, bg * (float)eg, bb * (float)eb);
            continue;
        }
        const Surface sf = hit_info<STIER>(S, o, d, h);
        const int samples = depth == 0 ? P.firstHitSamples : 1;
        const int nroot = (int)sqrt((double)samples);
        const float invnn = 1.0f / (float)(nroot * nroot);
        if (sf.mat.emittance > 0) {
            if (P.directLighting && !emission) continue;  // Sampler.cs:75-78
            float e = (float)(sf.mat.emittance * (double)samples) * invnn;  // Sampler.cs:79 then :144
            accumulate(sum, cnt, pixel, br * (float)sf.mat.cr * e, bg * (float)sf.mat.cg * e, bb * (float)sf.mat.cb * e);
        }
        int ma, mb;
        if (P.specularMode == PTGPU_SPECULAR_ALL || (depth == 0 && P.specularMode == PTGPU_SPECULAR_FIRST)) { ma = 1; mb = 2; }
        else { ma = 0; mb = 0; }
        Rng rng;
        rng_set_sample(rng, P.seed, P.passIndex, pixel, sample);
        uint32_t k = 0;
        for (int u = 0; u < nroot; u++) {
            for (int v = 0; v < nroot; v++) {
                for (int mode = ma; mode <= mb; mode++, k++) {
                    const uint32_t cFirst = depth == 0 ? (k + 1) : first;
                    const uint32_t cBits = depth == 0 ? 0u : (pathBits | ((uint32_t)(mode - ma) << ((depth - 1) & 31u)));
                    rng_enter(rng, cBits, cFirst, depth + 1, 0);
                    double fu = ((double)u + rng_next(rng)) / (double)nroot;
                    double fv = ((double)v + rng_next(rng)) / (double)nroot;
                    V3 no, nd;
                    bool reflected;
                    double p;
                    bounce(d, sf, fu, fv, mode, rng, no, nd, reflected, p);
                    if (mode == 0) p = 1;
                    if (!(p > 0)) continue;
                    float wr, wg, wb;
                    if (reflected) {  // indirect.Mix(Color.Mul(indirect), Tint) (Sampler.cs:113-114)
                        double tint = sf.mat.tint;
                        wr = (float)((1 - tint) + sf.mat.cr * tint); wg = (float)((1 - tint) + sf.mat.cg * tint); wb = (float)((1 - tint) + sf.mat.cb * tint);
                    } else { wr = (float)sf.mat.cr; wg = (float)sf.mat.cg; wb = (float)sf.mat.cb; }
                    const float pf = (float)p * invnn;
                    const float cr = br * wr * pf, cgn = bg * wg * pf, cb = bb * wb * pf;
                    // child path segment (traced only while depth+1 <= MaxBounces, Sampler.cs:57)
                    bool pushRay = (int)depth + 1 <= P.maxBounces;
                    float rr = 1.f;
                    if (pushRay && P.russianRoulette && depth >= 2) {
                        // Opt-in Russian roulette (ptgpu_pass.flags; the reference's own `russianRoulette` branch, Sampler.cs:133-142, is never
                        // enabled and does not terminate anything).  The child survives with probability q = the largest component of its local
                        // weight (material colour x branch probability), clamped to [minReflectance = 0.05, 1], and carries 1/q: unbiased.  The
                        // vertex's own next-event estimate below is not affected.  Sub-stream 255 is reserved for this draw.
                        const float q = fminf(fmaxf(fmaxf(wr, fmaxf(wg, wb)) * (float)p, 0.05f), 1.f);
                        rng_enter(rng, cBits, cFirst, depth + 1, 255);
                        if (rng_next(rng) >= (double)q) pushRay = false; else rr = 1.f / q;
                    }
                    if (pushRay) {
                        auto g2 = cg::coalesced_threads();
                        uint32_t base = 0;
                        if (g2.thread_rank() == 0) base = atomicAdd(ncount, g2.size());
                        base = g2.shfl(base, 0);
                        uint32_t slot = base + g2.thread_rank();
                        if (slot < capRays) {
                            nq.od0[slot] = make_float4(no.x, no.y, no.z, u2f(pixel));
                            nq.od1[slot] = make_float4(nd.x, nd.y, nd.z, u2f((depth + 1) | ((reflected ? 1u : 0u) << 6) | (cFirst << 7)));
                            nq.bt[slot] = make_float4(cr * rr, cgn * rr, cb * rr, u2f(cBits));
                            nq.smp[slot] = sample;
                        }
                    }
                    // next-event estimation for diffuse bounces (Sampler.cs:122-128, 191-296)
                    if (!reflected && P.directLighting && S.numLights > 0) {
                        const uint32_t nL = S.numLights;
                        const uint32_t loops = P.lightMode == PTGPU_LIGHT_ALL ? nL : 1u;
                        const float lscale = P.lightMode == PTGPU_LIGHT_ALL ? 1.0f / (float)nL : (float)nL;
                        for (uint32_t li = 0; li < loops; li++) {
                            uint32_t lightIndex = li;
                            if (P.lightMode == PTGPU_LIGHT_ALL) rng_enter(rng, cBits, cFirst, depth + 1, 1 + (li % 255u));
                            else {
                                rng_enter(rng, cBits, cFirst, depth + 1, 1);
                                int idx = (int)(rng_next(rng) * (double)nL);  // Random.Shared.Next(nLights)
                                lightIndex = (uint32_t)(idx >= (int)nL ? (int)nL - 1 : idx);
                            }
                            const DLight L = lights[lightIndex];
                            const V3 center = ld3(L.center);
                            const double radius = L.radius;
                            V3 point = center;
                            if (P.softShadows) {  // Sampler.cs:240-255
                                for (;;) {
                                    double x = rng_next(rng) * 2 - 1;
                                    double y = rng_next(rng) * 2 - 1;
                                    if (x * x + y * y <= 1) {
                                        V3 l = vnorm_c(vsub(center, sf.position));
                                        V3 uu = vnorm_c(vcross(l, random_unit_vector(rng)));
                                        V3 vv = vcross(l, uu);
                                        point = vadd(vadd(center, vmuls(uu, x * radius)), vmuls(vv, y * radius));
                                        break;
                                    }
                                }
                            }
                            V3 rayDirection = vnorm_c(vsub(point, sf.position));
                            double diffuse = vdot(rayDirection, sf.normal);
                            if (diffuse <= 0) continue;
                            if (!L.classTyped) continue;  // `hit.Shape != light` is always true for struct shapes
                            double coverage;
                            if (L.isCylinder) coverage = 1.0;
                            else {  // Sampler.cs:277-288
                                // theta = asin(R/hyp); adj = R/tan(theta); d = cos(theta)*adj; r = sin(theta)*adj; coverage = r*r/(d*d)
                                // is tan^2(theta) = R^2 / (hyp^2 - R^2): evaluated in that form (no asin/tan/sincos; differs from
                                // the reference's rounding by a few FP64 ulps, on a continuous weight)
                                double hyp = (double)vlenf(vsub(center, sf.position));
                                coverage = (radius * radius) / (hyp * hyp - radius * radius);
                                if (hyp < radius || hyp == radius) coverage = 1;  // asin(>1) is NaN -> 1; tan^2(pi/2) ~ 2.7e32 -> min(.., 1)
                                coverage = netmin(coverage, 1);
                            }
                            const ptgpu_shape lsh = S.shapes[L.shape];
                            Mat lm = shape_material<STIER>(S, lsh, -1, point);  // Material.MaterialAt(light, point)
                            float m = (float)(lm.emittance * diffuse * coverage) * lscale;
                            auto g3 = cg::coalesced_threads();
                            uint32_t base = 0;
                            if (g3.thread_rank() == 0) base = atomicAdd(scount, g3.size());
                            base = g3.shfl(base, 0);
                            uint32_t slot = base + g3.thread_rank();
                            if (slot < capShadow) {
                                sq.so[slot] = make_float4(sf.position.x, sf.position.y, sf.position.z, u2f(pixel));
                                sq.sd[slot] = make_float4(rayDirection.x, rayDirection.y, rayDirection.z, u2f(L.shape));
                                sq.sc[slot] = make_float4(cr * (float)lm.cr * m, cgn * (float)lm.cg * m, cb * (float)lm.cb * m, 0.f);
                            } else *overflow = 1u;  // dropped: the pass is reported and not added to the Buffer
                        }
                    }
                }
            }
        }
    }
}

// K2 / K4.  Scene.Intersect of the path segments and of sampleLight's visibility rays (Sampler.cs:261-265: closest hit, then identity
// with the light) - see "split tracer" in pt_device.cuh: the scene level as streaming kernels, the deferred shapes as persistent ones.
#ifndef PT_SCENE_MINBLOCKS
#define PT_SCENE_MINBLOCKS 8
#endif
template <int MODE, int TIER>
__global__ void __launch_bounds__(128, MODE == SCENE_FINISH ? 4 : PT_SCENE_MINBLOCKS) k_scene_trace(DScene S, SplitState W, RayQueue q, const uint32_t* __restrict__ count, MeshQueue in, MeshQueue out, HitQueue hq,
                                                      DeviceCounters* cnt) {
    constexpr bool RESUME = MODE != SCENE_START;
    const uint32_t n = RESUME ? *in.count : *count;
    scene_advance<MODE, TIER>(S, W, n, in, out,
                          [&](uint32_t i, V3& o, V3& d) { float4 a = q.od0[i], b = q.od1[i]; o = v3(a.x, a.y, a.z); d = v3(b.x, b.y, b.z); },
                          [&](uint32_t i, const HitRec& h) { hq.t[i] = h.t; hq.tInner[i] = h.tInner; hq.shape[i] = h.shape; hq.prim[i] = h.prim; });
    if (!RESUME && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&cnt->segments, (unsigned long long)n);
}
template <int MODE, int TIER>
__global__ void __launch_bounds__(128, MODE == SCENE_FINISH ? 4 : PT_SCENE_MINBLOCKS) k_scene_shadow(DScene S, SplitState W, ShadowQueue sq, const uint32_t* __restrict__ scount, uint32_t capShadow, uint32_t first,
                                                       uint32_t chunk, MeshQueue in, MeshQueue out, float* __restrict__ sum, DeviceCounters* cnt) {
    // The shadow queue may hold several times the rays the tracer has per-ray state for: it is traced in chunks of `chunk` records
    // starting at `first`; inside a chunk rays are numbered from 0 (state, work items), the queue is read at first + i.
    constexpr bool RESUME = MODE != SCENE_START;
    uint32_t n;
    if (RESUME) n = *in.count;
    else {
        const uint32_t total = min(*scount, capShadow);
        n = total > first ? min(total - first, chunk) : 0u;
    }
    scene_advance<MODE, TIER, PT_ANYHIT != 0>(S, W, n, in, out,
                          [&](uint32_t i, V3& o, V3& d) { float4 a = sq.so[first + i], b = sq.sd[first + i]; o = v3(a.x, a.y, a.z); d = v3(b.x, b.y, b.z); },
                          [&](uint32_t i, const HitRec& h) {
                              const uint32_t light = f2u(sq.sd[first + i].w);
                              if (h.shape >= 0 && (uint32_t)h.shape == light) {
                                  float4 c = sq.sc[first + i];
                                  accumulate(sum, cnt, f2u(sq.so[first + i].w), c.x, c.y, c.z);
                              }
                          },
                          [&](uint32_t i) { return (int32_t)f2u(sq.sd[first + i].w); });
    if (!RESUME && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&cnt->shadowRays, (unsigned long long)n);
}
// Small blocks: the warps of a launch finish at very different times (a few grazing rays take ~1000 steps) and a block's
// registers are only returned when its last warp exits.  The walk is latency-bound, so occupancy pays: 2-warp blocks
// (32 blocks/SM is the hardware limit) at 48 registers = 40 warps/SM measured 4 % faster than 32 warps at 57 registers.
#ifndef PT_MESH_BLOCK
#define PT_MESH_BLOCK 64
#endif
#ifndef PT_MESH_WARPS_PER_SM
#define PT_MESH_WARPS_PER_SM 40
#endif
// 40 warps x 48 registers is the register file: each of the four sub-partitions holds 16 K registers = 10 warps of 48 (42 warps would need 40 registers).
// ANYHIT: the items are Mesh.Intersect calls of shadow rays, each with the light's own T (see scene_advance in pt_device.cuh).
template <bool ANYHIT>
__global__ void __launch_bounds__(PT_MESH_BLOCK, PT_MESH_WARPS_PER_SM * 32 / PT_MESH_BLOCK) k_mesh(DScene S, SplitState W, MeshQueue q, uint32_t* __restrict__ cursor) {
    mesh_walk<ANYHIT>(S, W, q, cursor);
}

#ifndef PT_MARCH_MINBLOCKS
#define PT_MARCH_MINBLOCKS 16   // 64-thread blocks: 32 warps/SM at 64 registers.  The march loops are chains of dependent FP64 operations (ncu: `wait` is the top stall), so warps in flight pay: 6 blocks (91 registers) 213 ms per 2-spp C5 pass, 10: 175, 12: 169, 16: 166, 20: 170
#endif
template <int KIND>
__global__ void __launch_bounds__(64, PT_MARCH_MINBLOCKS) k_march(DScene S, SplitState W, MeshQueue q, uint32_t* __restrict__ cursor) {
    march_items<KIND>(S, W, q, cursor);
}

__global__ void k_clamp_count(uint32_t* count, uint32_t cap, uint32_t* overflow) {
    if (*count > cap) { *overflow = 1; *count = cap; }
}

// K5.  c /= spp; Buffer.AddSample (Renderer.cs:307-309, Buffer.cs:33-44), Welford state in FP64 like Colour.
struct PixelBuf { double* M; double* V; int32_t* samples; };
// `overflow` (one flag per lane, may be NULL): a queue overflowed in this pass, so its sum is incomplete and the Buffer is left alone.
// meanOut receives s * meanScale, added to what it holds when meanAdd is set (the stratified branch reports the mean over its strata).
struct SumSet { const float* p[PTGPU_MAX_DEVICES]; int n; };  // the pass accumulators of every device of the handle (peers read over NVLink)
__global__ void k_add_sample(SumSet sums, double divisor, uint32_t npix, PixelBuf pb, float* __restrict__ meanOut, const uint32_t* __restrict__ overflow,
                             int numFlags, float meanScale, int meanAdd) {
    for (int k = 0; k < numFlags; k++) if (overflow[k * 16]) return;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        int32_t ns = pb.samples[i] + 1;
        pb.samples[i] = ns;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float acc = sums.p[0][(size_t)i * 3 + c];
            for (int k = 1; k < sums.n; k++) acc += sums.p[k][(size_t)i * 3 + c];  // the multi-GPU reduce, fused into Buffer.AddSample
            double s = (double)acc / divisor;
            if (meanOut) meanOut[(size_t)i * 3 + c] = (meanAdd ? meanOut[(size_t)i * 3 + c] : 0.f) + (float)s * meanScale;
            if (ns == 1) { pb.M[(size_t)i * 3 + c] = s; continue; }
            double m = pb.M[(size_t)i * 3 + c];
            double M2 = m + (s - m) / (double)ns;
            pb.M[(size_t)i * 3 + c] = M2;
            pb.V[(size_t)i * 3 + c] = pb.V[(size_t)i * 3 + c] + (s - m) * (s - M2);
        }
    }
}
// Buffer.Color / Variance / StandardDeviation / Samples (Buffer.cs:46-57, 126-132) and the two derived channels of
// Buffer.Image: Albedo (CalculateAlbedo, Buffer.cs:240-282) and Normal (GetNormalAt / CalculateNormal, Buffer.cs:99-124, 222-233).
PT_D double clamp01_net(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }  // Math.Clamp(v, 0, 1): NaN passes through
__global__ void k_read_buffer(PixelBuf pb, int channel, int w, int h, float* __restrict__ out) {
    const uint32_t npix = (uint32_t)w * (uint32_t)h;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        int32_t ns = pb.samples[i];
        if (channel == 4) {  // color / max(r, g, b), clamped; NaN colour or zero maximum -> black
            const double r = pb.M[(size_t)i * 3], g = pb.M[(size_t)i * 3 + 1], b = pb.M[(size_t)i * 3 + 2];
            double o3[3] = {0, 0, 0};
            if (!(r != r || g != g || b != b)) {
                const double mx = netmax(netmax(r, g), b);
                if (mx != 0) { o3[0] = clamp01_net(r / mx); o3[1] = clamp01_net(g / mx); o3[2] = clamp01_net(b / mx); }
            }
            for (int c = 0; c < 3; c++) out[(size_t)i * 3 + c] = (float)o3[c];
            continue;
        }
        if (channel == 5) {  // Sobel-style differences of the neighbours' means; pixels outside the frame are `new Pixel()` (zero)
            const int x = (int)(i % (uint32_t)w), y = (int)(i / (uint32_t)w);
            auto M = [&](int xx, int yy, int c) -> double { return (xx < 0 || yy < 0 || xx >= w || yy >= h) ? 0.0 : pb.M[((size_t)yy * w + xx) * 3 + c]; };
            double nx = (M(x - 1, y - 1, 0) + 2 * M(x - 1, y, 0) + M(x - 1, y + 1, 0) - M(x + 1, y - 1, 0) - 2 * M(x + 1, y, 0) - M(x + 1, y + 1, 0)) / 8.0;
            double ny = (M(x - 1, y - 1, 1) + 2 * M(x, y - 1, 1) + M(x + 1, y - 1, 1) - M(x - 1, y + 1, 1) - 2 * M(x, y + 1, 1) - M(x + 1, y + 1, 1)) / 8.0;
            double nz = 1.0;
            const double length = sqrt(nx * nx + ny * ny + nz * nz);
            nx /= length; ny /= length; nz /= length;
            const V3 n = v3d(nx, ny, nz);  // new Vector(nx, ny, nz): FP32 storage
            out[(size_t)i * 3] = (float)(((double)n.x + 1.0) * 0.5);
            out[(size_t)i * 3 + 1] = (float)(((double)n.y + 1.0) * 0.5);
            out[(size_t)i * 3 + 2] = n.z;
            continue;
        }
        for (int c = 0; c < 3; c++) {
            double v;
            if (channel == 0) v = pb.M[(size_t)i * 3 + c];
            else if (channel == 3) v = (double)ns;
            else {
                v = ns < 2 ? 0.0 : pb.V[(size_t)i * 3 + c] / (double)(ns - 1);
                if (channel == 2) v = pow(v, (double)0.5f);
            }
            out[(size_t)i * 3 + c] = (float)v;
        }
    }
}

// Firefly pass (Renderer.cs:418-468).  Pixels whose StandardDeviation().MaxComponent() exceeds the threshold.
// mode 0: deviation > threshold (Renderer.cs:426 / :174).  mode 1: the serial Render()'s adaptive rule (Renderer.cs:153-158):
// (int)pow(clamp(deviation / threshold, 0, 1), exponent) >= 1.
__global__ void k_firefly_select(PixelBuf pb, double threshold, uint32_t npix, uint32_t* __restrict__ list, uint32_t* __restrict__ listCount, int mode = 0,
                                 double exponent = 1.0) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        int32_t ns = pb.samples[i];
        double mx = 0;
        for (int c = 0; c < 3; c++) {
            double v = ns < 2 ? 0.0 : pb.V[(size_t)i * 3 + c] / (double)(ns - 1);
            double sd = pow(v, (double)0.5f);
            mx = c == 0 ? sd : netmax(mx, sd);
        }
        bool pick;
        if (mode == 0) pick = mx > threshold;
        else {
            double v = mx / threshold;
            v = v < 0 ? 0.0 : (v > 1 ? 1.0 : v);  // Math.Clamp
            v = pow(v, exponent);
            pick = (int)v >= 1;
        }
        if (pick) {
            auto g = cg::coalesced_threads();
            uint32_t base = 0;
            if (g.thread_rank() == 0) base = atomicAdd(listCount, g.size());
            list[g.shfl(base, 0) + g.thread_rank()] = i;
        }
    }
}
// IsFirefly (Renderer.cs:474-497) + CalculateLocalDeviation (:499-537) for the new sample of every listed pixel, against the
// buffer as it stands at the start of this iteration (the reference reads its neighbours while other threads update them).
__global__ void k_firefly_decide(const float* __restrict__ sum, const uint32_t* __restrict__ list, const uint32_t* __restrict__ listCount, PixelBuf pb,
                                 int w, int h, uint8_t* __restrict__ reject) {
    const uint32_t n = *listCount;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t pixel = list[i];
        const int x = (int)(pixel % (uint32_t)w), y = (int)(pixel / (uint32_t)w);
        const double r = sum[(size_t)pixel * 3], g = sum[(size_t)pixel * 3 + 1], b = sum[(size_t)pixel * 3 + 2];
        const double brightness = r * 0.2126 + g * 0.7152 + b * 0.0722;
        bool rej = false;
        if (brightness > 0.9) {
            int x0 = max(0, x - 1), y0 = max(0, y - 1), x1 = min(w - 1, x + 1), y1 = min(h - 1, y + 1);
            double tr = 0, tg = 0, tb = 0;
            int count = 0;
            for (int j = y0; j <= y1; j++)
                for (int ii = x0; ii <= x1; ii++) {
                    size_t q = ((size_t)j * w + ii) * 3;
                    tr += pb.M[q]; tg += pb.M[q + 1]; tb += pb.M[q + 2];
                    count++;
                }
            double dr = fabs(r - tr / count), dg = fabs(g - tg / count), db = fabs(b - tb / count);
            rej = sqrt(dr * dr + dg * dg + db * db) > 0.2;
        }
        reject[i] = rej ? 1 : 0;
    }
}
// Accepted samples go through Buffer.AddSample and the pixel stays listed; a rejected sample ends the pixel's loop (`break`).
__global__ void k_firefly_apply(float* __restrict__ sum, const uint32_t* __restrict__ list, const uint32_t* __restrict__ listCount, const uint8_t* __restrict__ reject,
                                PixelBuf pb, uint32_t* __restrict__ nextList, uint32_t* __restrict__ nextCount) {
    const uint32_t n = *listCount;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t pixel = list[i];
        const bool rej = reject ? reject[i] : false;
        if (!rej) {
            int32_t ns = pb.samples[pixel] + 1;
            pb.samples[pixel] = ns;
            for (int c = 0; c < 3; c++) {
                double s = (double)sum[(size_t)pixel * 3 + c];
                if (ns == 1) { pb.M[(size_t)pixel * 3 + c] = s; continue; }
                double m = pb.M[(size_t)pixel * 3 + c];
                double M2 = m + (s - m) / (double)ns;
                pb.M[(size_t)pixel * 3 + c] = M2;
                pb.V[(size_t)pixel * 3 + c] = pb.V[(size_t)pixel * 3 + c] + (s - m) * (s - M2);
            }
            auto g = cg::coalesced_threads();
            uint32_t base = 0;
            if (g.thread_rank() == 0) base = atomicAdd(nextCount, g.size());
            nextList[g.shfl(base, 0) + g.thread_rank()] = pixel;
        }
        sum[(size_t)pixel * 3] = 0.f; sum[(size_t)pixel * 3 + 1] = 0.f; sum[(size_t)pixel * 3 + 2] = 0.f;
    }
}

// K6.  Test hook: Scene.Intersect + Hit.Info on caller-supplied rays, through the same split tracer as the pipeline.
struct BatchOut { int32_t* shape; int32_t* prim; double* t; float* normal3; float* position3; int32_t* inside; int32_t* material; };
template <int MODE, int TIER>
__global__ void __launch_bounds__(128) k_scene_batch(DScene S, SplitState W, uint32_t nStart, MeshQueue in, MeshQueue out, const float* __restrict__ o3,
                                                      const float* __restrict__ d3, BatchOut B) {
    constexpr bool RESUME = MODE != SCENE_START;
    const uint32_t n = RESUME ? *in.count : nStart;
    scene_advance<MODE, TIER>(S, W, n, in, out,
                          [&](uint32_t i, V3& o, V3& d) { o = v3(o3[3 * i], o3[3 * i + 1], o3[3 * i + 2]); d = v3(d3[3 * i], d3[3 * i + 1], d3[3 * i + 2]); },
                          [&](uint32_t i, const HitRec& h) {
                              V3 o = v3(o3[3 * i], o3[3 * i + 1], o3[3 * i + 2]), d = v3(d3[3 * i], d3[3 * i + 1], d3[3 * i + 2]);
                              B.shape[i] = h.shape;
                              B.t[i] = h.t;
                              int32_t localPrim = -1;
                              V3 nn = v3(0, 0, 0), pp = v3(0, 0, 0);
                              int32_t ins = 0, mat = -1;
                              if (h.shape >= 0) {
                                  if (h.prim >= 0) {
                                      ptgpu_shape sh = S.shapes[h.shape];
                                      while (sh.type == PTGPU_TRANSFORMED) sh = S.shapes[S.instances[sh.data].shape];
                                      localPrim = h.prim - (int32_t)S.meshes[sh.data].triFirst;
                                  }
                                  if (B.normal3 || B.position3 || B.inside || B.material) {  // Hit.Info only when asked for
                                      Surface sf = hit_info<2>(S, o, d, h);
                                      nn = sf.normal; pp = sf.position; ins = sf.inside ? 1 : 0; mat = sf.mat.id;
                                  }
                              }
                              B.prim[i] = localPrim;
                              if (B.normal3) { B.normal3[3 * i] = nn.x; B.normal3[3 * i + 1] = nn.y; B.normal3[3 * i + 2] = nn.z; }
                              if (B.position3) { B.position3[3 * i] = pp.x; B.position3[3 * i + 1] = pp.y; B.position3[3 * i + 2] = pp.z; }
                              if (B.inside) B.inside[i] = ins;
                              if (B.material) B.material[i] = mat;
                          });
}
__global__ void k_cast_rays(PassD P, int n, const int32_t* x, const int32_t* y, const double* fu, const double* fv, const int32_t* sample, float* o3, float* d3) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Rng rng;
        rng_set_sample(rng, P.seed, P.passIndex, (uint32_t)(y[i] * P.width + x[i]), (uint32_t)sample[i]);
        rng_enter(rng, 0, 0, 0, 0);
        rng_next(rng); rng_next(rng);  // the jitter draws come first in the stream
        V3 o, d;
        cast_ray(P.cam, x[i], y[i], P.width, P.height, fu[i], fv[i], rng, o, d);
        o3[3 * i] = o.x; o3[3 * i + 1] = o.y; o3[3 * i + 2] = o.z;
        d3[3 * i] = d.x; d3[3 * i + 1] = d.y; d3[3 * i + 2] = d.z;
    }
}
__global__ void k_keyed_draw(uint32_t seed, uint32_t pass, uint32_t pixel, uint32_t sample, uint32_t bits, uint32_t first, uint32_t depth, uint32_t sub,
                             uint32_t drawIndex, double* out) {
    Rng rng;
    rng_set_sample(rng, seed, pass, pixel, sample);
    rng_enter(rng, bits, first, depth, sub);
    double v = 0;
    for (uint32_t i = 0; i <= drawIndex; i++) v = rng_next(rng);
    *out = v;
}
// ====================================================================================================== host side
static constexpr int kMaxLanes = 8;
#ifndef PT_LANES
#define PT_LANES 2   // C3, 512 spp: 1 -> 362.2, 2 -> 367.7, 3 -> 363.8, 4 -> 362.7 Msamples/s (the tail of one batch's k_mesh overlaps the next batch)
#endif
struct Lane {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    RayQueue rq[2]{};
    HitQueue hq{};
    ShadowQueue sq{};
    uint64_t capRays = 0, capShadow = 0;   // records the lane's queues are allocated for (grown on demand up to ctx->capRays)
    uint32_t* counts = nullptr;   // [0],[1] ray queue counts, [2] shadow count, [3] overflow flag, [4] trace cursor, [5] shadow cursor,
                                  // [10],[11] mesh queue counts, [12] mesh cursor
    SplitState split{};           // split tracer (scene_advance / mesh_walk): per-ray scene-level state and the two mesh work queues
    MeshQueue mq[2]{};
    uint64_t splitCap = 0;
    uint32_t* perm = nullptr;     // shade order of the current launch (capRays entries)
    uint32_t* bins = nullptr;     // kShadeBins counters / cursors
};
struct ptgpu_ctx {
    int device = 0;
    int numSMs = 148;
    // multi-GPU inside the handle: the handle the caller holds is the root (devices[0]); peers are full contexts of their own
    // device (scene replica, queues, pass accumulator) driven by one host thread each during a pass
    std::vector<ptgpu_ctx*> peers;
    std::vector<float*> peerStage;        // root-side copies of the peers' accumulators when peer access is not available
    std::vector<char> peerDirect;         // the root can read peer k's memory (cudaDeviceEnablePeerAccess)
    cudaEvent_t evPeerDone = nullptr;     // on a peer: its share of the pass is in its dSum
    uint32_t* laneCounts = nullptr;       // kMaxLanes x 16 counters, lane k at + 16 k (so the overflow flags are one strided array)
    uint64_t overflowPasses = 0;
    uint64_t traceLaunches = 0;
    cudaStream_t stream = nullptr;
    std::string error;
    // scene
    std::vector<std::pair<void*, uint64_t>> sceneAllocs;   // (pointer, bytes) of the resident scene
    std::vector<std::pair<void*, uint64_t>> stage;         // pinned host staging buffers (pointer, capacity), reused across uploads
    size_t stageNext = 0;
    double deriveMs = 0;                                   // host time of the last derive_mesh
    std::vector<std::pair<void*, uint64_t>> scenePool;     // buffers of the previous scene, reused by the next upload of similar size
    DScene scene{};
    DLight* dLights = nullptr;
    bool haveScene = false;
    uint64_t meshNodeBytes = 0;
    uint64_t sceneBytes = 0;
    // queues: `numLanes` independent sets (own stream, queues, split-tracer state); the batches of a pass go round the
    // lanes so that the tail of one batch's kernel (a handful of long rays) overlaps the bulk of another batch's
    uint64_t capRays = 0;         // path records per lane
    Lane lanes[kMaxLanes];
    int numLanes = 1;
    cudaEvent_t evFork = nullptr;
    bool hasKind[3] = {false, false, false};  // deferred shapes in the scene: Mesh, SDFShape, Volume (directly or under a TransformedShape)
    uint64_t roundItems = 0;                    // profiling: work items (all kinds) of the rounds in which k_mesh ran
    double kindMs[3] = {0, 0, 0};               // profiling: device time / launches of k_mesh, k_march<SDF>, k_march<VOLUME> in the last pass
    uint64_t kindLaunches[3] = {0, 0, 0};
    int splitStackEnt = 2;
    int sceneTier = TIER_FULL;  // TIER_* of the uploaded scene (scene_advance)
    int shadeTier = 2;          // k_shade<STIER>: 0 analytic shapes, no textures; 1 ... and Meshes; 2 everything
    int splitRounds = 0;          // 0: no deferred shapes; > 0: every ray enters at most this many (fixed rounds, no host sync); -1: loop on the queue count
    uint32_t* dCounts = nullptr;  // [6] batch cursor, [8],[9] firefly list counts
    DeviceCounters* dCounters = nullptr;
    // image state
    int bufW = 0, bufH = 0;
    float* dSum = nullptr;
    float* dMean = nullptr;
    PixelBuf pb{};
    uint32_t* dList[2] = {nullptr, nullptr};  // firefly pixel lists (ping-pong)
    uint8_t* dReject = nullptr;
    // stats
    std::atomic<uint64_t> launches{0};  // (the lanes of a polled scene are driven by one host thread each)
    double lastPassMs = 0, traceMs = 0, shadeMs = 0, shadowMs = 0, raygenMs = 0;
    bool profiling = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evA = nullptr, evB = nullptr, evC = nullptr, evD = nullptr;
};

static std::string g_createError;
static std::mutex g_mu;

static std::mutex g_errMu;  // ctx->error may be written by the lane threads of a polled scene
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            { std::lock_guard<std::mutex> lk__(g_errMu); ctx->error = std::string(#call) + ": " + cudaGetErrorString(e__); } \
            return PTGPU_E_CUDA;                                                                         \
        }                                                                                                \
    } while (0)

static int fail(ptgpu_ctx* ctx, int code, const std::string& msg) {
    std::lock_guard<std::mutex> lk(g_errMu);
    ctx->error = msg;
    return code;
}

// Scene buffers come from a small pool: re-uploading a scene of the same shape (a new frame of an animation, bench.py's
// end-to-end step) reuses the previous allocations; cudaMalloc / cudaFree next to tens of GB of queues cost 0.2-1.5 s.
static int scene_alloc(ptgpu_ctx* ctx, uint64_t bytes, void** out) {
    int best = -1;
    for (size_t i = 0; i < ctx->scenePool.size(); i++) {
        const uint64_t b = ctx->scenePool[i].second;
        if (b >= bytes && b <= bytes + bytes / 4 + 4096 && (best < 0 || b < ctx->scenePool[best].second)) best = (int)i;
    }
    if (best >= 0) {
        *out = ctx->scenePool[best].first;
        ctx->sceneAllocs.push_back(ctx->scenePool[best]);
        ctx->scenePool.erase(ctx->scenePool.begin() + best);
        return PTGPU_OK;
    }
    void* p = nullptr;
    CK(cudaMalloc(&p, bytes));
    ctx->sceneAllocs.push_back({p, bytes});
    *out = p;
    return PTGPU_OK;
}
// Pinned host staging, slot by slot in request order (the same scene shape asks for the same sizes again).
static void* stage_alloc(ptgpu_ctx* ctx, uint64_t bytes) {
    const size_t slot = ctx->stageNext++;
    if (slot >= ctx->stage.size()) ctx->stage.push_back({nullptr, 0});
    auto& b = ctx->stage[slot];
    if (b.second < bytes) {
        if (b.first) cudaFreeHost(b.first);
        b = {nullptr, 0};
        const uint64_t cap = bytes + bytes / 8 + 4096;
        void* p = nullptr;
        if (cudaMallocHost(&p, cap) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        b = {p, cap};
    }
    return b.first;
}
template <class T>
static int upload(ptgpu_ctx* ctx, const T* host, uint64_t count, const T** dev) {
    *dev = nullptr;
    uint64_t bytes = (count ? count : 1) * sizeof(T);
    void* p = nullptr;
    int rc = scene_alloc(ctx, bytes, &p);
    if (rc != PTGPU_OK) return rc;
    ctx->sceneBytes += count * sizeof(T);
    if (count) CK(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    *dev = reinterpret_cast<const T*>(p);
    return PTGPU_OK;
}

// The resident scene's buffers go to the pool (release = true: to the driver).
static void free_scene(ptgpu_ctx* ctx, bool release = false) {
    for (auto& a : ctx->sceneAllocs) ctx->scenePool.push_back(a);
    ctx->sceneAllocs.clear();
    if (release) { for (auto& a : ctx->scenePool) cudaFree(a.first); ctx->scenePool.clear(); }
    ctx->haveScene = false;
    ctx->sceneBytes = 0;
    ctx->dLights = nullptr;
}
static void trim_scene_pool(ptgpu_ctx* ctx) {  // after an upload: what the new scene did not reuse goes back to the driver
    for (auto& a : ctx->scenePool) cudaFree(a.first);
    ctx->scenePool.clear();
}
static void free_split(Lane& L) {
    SplitState& W = L.split;
    void* ps[] = {W.state, W.sceneStack, W.meshStack, L.mq[0].a, L.mq[0].b, L.mq[0].c, L.mq[1].a, L.mq[1].b, L.mq[1].c, L.mq[0].lim, L.mq[1].lim};
    for (void* p : ps) cudaFree(p);
    W = SplitState{};
    L.mq[0] = L.mq[1] = MeshQueue{};
    L.splitCap = 0;
}
static void free_queues(ptgpu_ctx* ctx) {
    for (int k = 0; k < kMaxLanes; k++) {
        Lane& L = ctx->lanes[k];
        for (int i = 0; i < 2; i++) { cudaFree(L.rq[i].od0); cudaFree(L.rq[i].od1); cudaFree(L.rq[i].bt); cudaFree(L.rq[i].smp); L.rq[i] = RayQueue{}; }
        cudaFree(L.hq.t); cudaFree(L.hq.tInner); cudaFree(L.hq.shape); cudaFree(L.hq.prim); L.hq = HitQueue{};
        cudaFree(L.sq.so); cudaFree(L.sq.sd); cudaFree(L.sq.sc); L.sq = ShadowQueue{};
        cudaFree(L.perm); cudaFree(L.bins); L.perm = nullptr; L.bins = nullptr;
        L.capRays = 0; L.capShadow = 0;
        free_split(L);
    }
}
static void free_image(ptgpu_ctx* ctx) {
    cudaFree(ctx->dSum); cudaFree(ctx->dMean); cudaFree(ctx->pb.M); cudaFree(ctx->pb.V); cudaFree(ctx->pb.samples);
    cudaFree(ctx->dList[0]); cudaFree(ctx->dList[1]); cudaFree(ctx->dReject);
    ctx->dList[0] = ctx->dList[1] = nullptr; ctx->dReject = nullptr;
    ctx->dSum = ctx->dMean = nullptr; ctx->pb = PixelBuf{}; ctx->bufW = ctx->bufH = 0;
}

static int grid_for(ptgpu_ctx* ctx, int blocksPerSM) { return ctx->numSMs * blocksPerSM; }

// Per-ray state of the split tracer for launches of up to `cap` rays.
static int ensure_split(ptgpu_ctx* ctx, Lane& L, uint64_t cap, int stackEnt) {
    const uint32_t quads = (PT_SCENE_MASK && ctx->scene.maskOn) ? 8u : 6u;  // records are allocated at the larger size
    L.split.stateQuads = quads;
    if (cap <= L.splitCap && stackEnt == L.split.stackEnt) return PTGPU_OK;
    CK(cudaStreamSynchronize(L.stream));
    free_split(L);
    SplitState& W = L.split;
    CK(cudaMalloc(&W.state, cap * sizeof(RayState)));
    W.stateQuads = quads;
    CK(cudaMalloc(&W.sceneStack, cap * (uint64_t)stackEnt * sizeof(uint4)));
#if PT_MESH_GSTACK
    CK(cudaMalloc(&W.meshStack, (uint64_t)grid_for(ctx, PT_MESH_WARPS_PER_SM * 32 / PT_MESH_BLOCK) * PT_MESH_BLOCK * kMeshStackEnt * sizeof(uint4)));
#endif
    W.stackEnt = stackEnt;
    W.kindItems = ctx->dCounters->kindItems;
    for (int i = 0; i < 2; i++) {
        CK(cudaMalloc(&L.mq[i].a, cap * sizeof(float4))); CK(cudaMalloc(&L.mq[i].b, cap * sizeof(float4))); CK(cudaMalloc(&L.mq[i].c, cap * sizeof(double2)));
        CK(cudaMalloc(&L.mq[i].lim, cap * sizeof(float)));
        L.mq[i].count = L.counts + 10 + i;
    }
    L.splitCap = cap;
    return PTGPU_OK;
}
// One Scene.Intersect wavefront.  start(out) launches the scene kernel for the fresh rays; resume(in, out) launches it for the rays
// named by `in`'s items, after the consumer kernels of the deferred shapes (k_mesh, k_march<SDF>, k_march<VOLUME>: whichever kinds
// the scene holds) have written their Hits.  No rounds for scenes without deferred shapes, a fixed number when the scene bounds
// what a ray can enter, else until the queue is empty.
#ifndef PT_FINISH_MAX
#define PT_FINISH_MAX 65536   // pending items at or below which the remaining rounds run as one SCENE_FINISH launch
#endif
template <bool ANYHIT = false, class StartFn, class ResumeFn, class FinishFn>
static int run_split(ptgpu_ctx* ctx, Lane& L, cudaStream_t st, StartFn start, ResumeFn resume, FinishFn finish) {
    uint32_t* cursor = L.counts + 12;  // [12] k_mesh, [13] k_march<SDF>, [14] k_march<VOLUME>
    CK(cudaMemsetAsync(L.counts + 10, 0, 2 * sizeof(uint32_t), st));
    start(L.mq[0]);
    ctx->launches++;
    int cur = 0;
    static const bool detail = std::getenv("PTGPU_TRACE_DETAIL") != nullptr;
    const bool timed = ctx->profiling || detail;
    const int gridMesh = grid_for(ctx, PT_MESH_WARPS_PER_SM * 32 / PT_MESH_BLOCK), gridMarch = grid_for(ctx, PT_MARCH_MINBLOCKS);
    for (int round = 0; ctx->splitRounds < 0 || round < ctx->splitRounds; round++) {
        if (ctx->splitRounds < 0) {
            uint32_t pending = 0;
            CK(cudaMemcpyAsync(&pending, L.mq[cur].count, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (pending == 0) break;
            static const uint32_t finishMax = std::getenv("PTGPU_FINISH_MAX") ? (uint32_t)std::atoll(std::getenv("PTGPU_FINISH_MAX")) : (uint32_t)PT_FINISH_MAX;  // development override
            if (pending <= finishMax && round > 0) { finish(L.mq[cur]); ctx->launches++; break; }
        }
        CK(cudaMemsetAsync(L.mq[cur ^ 1].count, 0, sizeof(uint32_t), st));
        CK(cudaMemsetAsync(cursor, 0, 3 * sizeof(uint32_t), st));
        // per-launch device time of each consumer kernel when profiling (bench.py's roofline): kind 0 mesh, 1 SDF, 2 Volume
        uint32_t items = 0;
        if (timed) cudaMemcpyAsync(&items, L.mq[cur].count, 4, cudaMemcpyDeviceToHost, st);  // read by the time the first consumer is timed
        auto consumer = [&](int kind, auto launch) {
            if (timed) cudaEventRecord(ctx->evC, st);
            launch();
            ctx->launches++;
            if (timed) {
                cudaEventRecord(ctx->evD, st); cudaEventSynchronize(ctx->evD);
                float ms = 0; cudaEventElapsedTime(&ms, ctx->evC, ctx->evD);
                ctx->kindMs[kind] += ms; ctx->kindLaunches[kind]++;
                if (kind == 0) ctx->roundItems += items;  // all kinds; the marchers' own counts are subtracted in ptgpu_get_counters
                if (detail) fprintf(stderr, "%s round %d  queue items %u  %.3f ms\n", kind == 0 ? "k_mesh" : kind == 1 ? "k_march<SDF>" : "k_march<VOLUME>", round, items, ms);
            }
        };
        if (ctx->hasKind[0]) consumer(0, [&] { k_mesh<ANYHIT><<<gridMesh, PT_MESH_BLOCK, 0, st>>>(ctx->scene, L.split, L.mq[cur], cursor); });
        if (ctx->hasKind[1]) consumer(1, [&] { k_march<(int)kItemSdf><<<gridMarch, 64, 0, st>>>(ctx->scene, L.split, L.mq[cur], cursor + 1); });
        if (ctx->hasKind[2]) consumer(2, [&] { k_march<(int)kItemVolume><<<gridMarch, 64, 0, st>>>(ctx->scene, L.split, L.mq[cur], cursor + 2); });
        resume(L.mq[cur], L.mq[cur ^ 1]);
        ctx->launches++;
        cur ^= 1;
    }
    return PTGPU_OK;
}

extern "C" {

int ptgpu_abi_version(void) { return PTGPU_ABI_VERSION; }

// Diagnostic (not part of include/ptgpu.h): run the host-side derivation of ptgpu_upload_scene without a device and return
// its time in ms and an FNV-1a hash of the derived buffers (tools/derive_time.py; negative = failure).
extern "C" double ptgpu_debug_derive(const ptgpu_flat_scene* s, uint64_t* hashOut, uint64_t* nodeRecords, uint64_t* leafTriangles) {
    MeshDerived dv;
    std::string err;
    std::vector<void*> bufs;
    auto t0 = std::chrono::steady_clock::now();
    const bool ok = derive_mesh(s, dv, err, [&](uint64_t bytes) -> void* { void* p = std::malloc(bytes); bufs.push_back(p); return p; });
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (ok) {
        uint64_t h = 1469598103934665603ull;
        auto mix = [&](const void* p, uint64_t bytes) { const uint64_t* w = static_cast<const uint64_t*>(p); for (uint64_t i = 0; i < bytes / 8; i++) { h ^= w[i]; h *= 1099511628211ull; } };
        mix(dv.mn, dv.mnRecords * 64); mix(dv.lg, dv.lgCount * sizeof(ptgpu_tri_geom));
        for (const ptgpu_tree& t : dv.trees) { h ^= t.root; h *= 1099511628211ull; }
        if (hashOut) *hashOut = h;
        if (nodeRecords) *nodeRecords = dv.mnRecords;
        if (leafTriangles) *leafTriangles = dv.lgCount;
    }
    for (void* p : bufs) std::free(p);
    return ok ? ms : -1.0;
}

const char* ptgpu_last_error(ptgpu_ctx* ctx) {
    if (ctx) return ctx->error.c_str();
    std::lock_guard<std::mutex> lk(g_mu);
    return g_createError.c_str();
}

int ptgpu_abi_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(ptgpu_params); case 1: return (int)sizeof(ptgpu_pass); case 2: return (int)sizeof(ptgpu_camera);
        case 3: return (int)sizeof(ptgpu_counters); case 4: return (int)sizeof(ptgpu_flat_scene); case 5: return (int)sizeof(ptgpu_node);
        case 6: return (int)sizeof(ptgpu_tree); case 7: return (int)sizeof(ptgpu_shape); case 8: return (int)sizeof(ptgpu_sphere);
        case 9: return (int)sizeof(ptgpu_cube); case 10: return (int)sizeof(ptgpu_plane); case 11: return (int)sizeof(ptgpu_cylinder);
        case 12: return (int)sizeof(ptgpu_mesh); case 13: return (int)sizeof(ptgpu_tri_geom); case 14: return (int)sizeof(ptgpu_tri_shade);
        case 15: return (int)sizeof(ptgpu_instance); case 16: return (int)sizeof(ptgpu_sdf_op); case 17: return (int)sizeof(ptgpu_sdf_shape);
        case 18: return (int)sizeof(ptgpu_volume_window); case 19: return (int)sizeof(ptgpu_volume); case 20: return (int)sizeof(ptgpu_material);
        case 21: return (int)sizeof(ptgpu_texture); case 22: return (int)sizeof(ptgpu_sh);
        default: return -1;
    }
}

static int create_one(const ptgpu_params* params, int dev, ptgpu_ctx** out);

// ptgpu_params.numDevices > 1: one context per device; the first is the handle the caller holds (the root), the others its peers.
int ptgpu_create(const ptgpu_params* params, ptgpu_ctx** out) {
    if (!out) return PTGPU_E_ARG;
    *out = nullptr;
    const int nd = params ? params->numDevices : 0;
    if (nd < 0 || nd > PTGPU_MAX_DEVICES) {
        std::lock_guard<std::mutex> lk(g_mu);
        g_createError = "numDevices out of range";
        return PTGPU_E_ARG;
    }
    if (nd <= 1) return create_one(params, nd == 1 ? params->devices[0] : (params ? params->device : 0), out);
    for (int a = 0; a < nd; a++)
        for (int b = a + 1; b < nd; b++)
            if (params->devices[a] == params->devices[b]) {
                std::lock_guard<std::mutex> lk(g_mu);
                g_createError = "devices[] names a device twice";
                return PTGPU_E_ARG;
            }
    ptgpu_ctx* root = nullptr;
    int rc = create_one(params, params->devices[0], &root);
    if (rc != PTGPU_OK) return rc;
    for (int k = 1; k < nd; k++) {
        ptgpu_ctx* peer = nullptr;
        rc = create_one(params, params->devices[k], &peer);
        if (rc != PTGPU_OK) { ptgpu_destroy(root); return rc; }
        cudaEventCreateWithFlags(&peer->evPeerDone, cudaEventDisableTiming);
        root->peers.push_back(peer);
        // NVLink peer access root -> peer: the Buffer.AddSample kernel of the root reads the peers' accumulators in place
        int can = 0;
        cudaSetDevice(root->device);
        bool direct = false;
        if (cudaDeviceCanAccessPeer(&can, root->device, peer->device) == cudaSuccess && can) {
            const cudaError_t pe = cudaDeviceEnablePeerAccess(peer->device, 0);
            direct = pe == cudaSuccess || pe == cudaErrorPeerAccessAlreadyEnabled;
            cudaGetLastError();
        }
        if (std::getenv("PTGPU_NO_PEER_ACCESS")) direct = false;  // development switch: exercise the staged copy
        root->peerDirect.push_back(direct ? 1 : 0);
        root->peerStage.push_back(nullptr);
    }
    cudaSetDevice(root->device);
    *out = root;
    return PTGPU_OK;
}

static int create_one(const ptgpu_params* params, int dev, ptgpu_ctx** out) {
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        std::lock_guard<std::mutex> lk(g_mu);
        g_createError = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                        " (libptgpu has no CPU fallback)";
        return PTGPU_E_CUDA;
    }
    if (dev < 0 || dev >= ndev) {
        std::lock_guard<std::mutex> lk(g_mu);
        g_createError = "device ordinal out of range";
        return PTGPU_E_ARG;
    }
    ptgpu_ctx* ctx = new ptgpu_ctx();
    ctx->device = dev;
    auto bail = [&](const char* what, cudaError_t err) {
        std::lock_guard<std::mutex> lk(g_mu);
        g_createError = std::string(what) + ": " + cudaGetErrorString(err);
        delete ctx;
        return PTGPU_E_CUDA;
    };
    if ((e = cudaSetDevice(dev)) != cudaSuccess) return bail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, dev)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
    ctx->numSMs = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1); cudaEventCreate(&ctx->evA); cudaEventCreate(&ctx->evB); cudaEventCreate(&ctx->evC); cudaEventCreate(&ctx->evD);
    // queueCapacity = path records in flight over all lanes (default 2^27, allocated on demand); flags bits 0-3 = number of lanes (0 = default)
    {
        uint64_t total = (params && params->queueCapacity) ? params->queueCapacity : (1ull << 27);
        int lanes = (params && (params->flags & 15)) ? (params->flags & 15) : PT_LANES;
        if (const char* ev = std::getenv("PTGPU_LANES")) { int v = std::atoi(ev); if (v >= 1) lanes = v; }            // development overrides
        if (const char* ev = std::getenv("PTGPU_QUEUE_LOG2")) { int v = std::atoi(ev); if (v >= 16 && v <= 30) total = 1ull << v; }
        if (lanes > kMaxLanes) lanes = kMaxLanes;
        if (total < (1ull << 22)) lanes = 1;
        ctx->numLanes = lanes;
        ctx->capRays = total / (uint64_t)lanes;
        if (ctx->capRays > (1ull << 30)) ctx->capRays = 1ull << 30;
        if (ctx->capRays == 0) ctx->capRays = 1;
    }
    if ((e = cudaMalloc(&ctx->laneCounts, kMaxLanes * 16 * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMalloc", e);
    cudaMemset(ctx->laneCounts, 0, kMaxLanes * 16 * sizeof(uint32_t));
    for (int k = 0; k < ctx->numLanes; k++) {
        Lane& L = ctx->lanes[k];
        if ((e = cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
        if ((e = cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
        L.counts = ctx->laneCounts + 16 * k;
    }
    if ((e = cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaMalloc(&ctx->dCounts, 16 * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->dCounters, sizeof(DeviceCounters))) != cudaSuccess) return bail("cudaMalloc", e);
    cudaMemset(ctx->dCounts, 0, 16 * sizeof(uint32_t));
    cudaMemset(ctx->dCounters, 0, sizeof(DeviceCounters));
    *out = ctx;
    return PTGPU_OK;
}

void ptgpu_destroy(ptgpu_ctx* ctx) {
    if (!ctx) return;
    for (ptgpu_ctx* p : ctx->peers) ptgpu_destroy(p);
    ctx->peers.clear();
    cudaSetDevice(ctx->device);
    for (float* p : ctx->peerStage) cudaFree(p);
    ctx->peerStage.clear();
    cudaStreamSynchronize(ctx->stream);
    free_scene(ctx, true);
    for (auto& b : ctx->stage) if (b.first) cudaFreeHost(b.first);
    ctx->stage.clear();
    free_queues(ctx);
    free_image(ctx);
    for (int k = 0; k < kMaxLanes; k++) {
        Lane& L = ctx->lanes[k];
        if (L.stream) cudaStreamDestroy(L.stream);
        if (L.done) cudaEventDestroy(L.done);
    }
    cudaFree(ctx->laneCounts);
    if (ctx->evPeerDone) cudaEventDestroy(ctx->evPeerDone);
    if (ctx->evFork) cudaEventDestroy(ctx->evFork);
    cudaFree(ctx->dCounts);
    cudaFree(ctx->dCounters);
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1); cudaEventDestroy(ctx->evA); cudaEventDestroy(ctx->evB); cudaEventDestroy(ctx->evC); cudaEventDestroy(ctx->evD);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

uint64_t ptgpu_scene_bytes(ptgpu_ctx* ctx) { return ctx ? ctx->sceneBytes : 0; }

// Index ranges of a flat scene (a PTFS file written elsewhere, or a foreign flattener): everything the upload and the kernels
// dereference.  Returns an empty string when the scene is consistent.
static std::string validate_flat_scene(const ptgpu_flat_scene* s) {
    auto bad = [](const char* what, uint64_t i) { return std::string("flat scene: ") + what + " (element " + std::to_string(i) + ")"; };
    if (s->numTrees == 0 || s->sceneTree >= s->numTrees) return "flat scene: sceneTree out of range";
    if (s->numSceneShapes > s->numShapes) return "flat scene: numSceneShapes > numShapes";
    auto mat_ok = [&](int32_t m) { return m >= -1 && (m < 0 || (uint32_t)m < s->numMaterials); };
    for (uint32_t i = 0; i < s->numShapes; i++) {
        const ptgpu_shape& sh = s->shapes[i];
        uint32_t lim = 0;
        switch (sh.type) {
            case PTGPU_SPHERE: lim = s->numSpheres; break;
            case PTGPU_CUBE: lim = s->numCubes; break;
            case PTGPU_PLANE: lim = s->numPlanes; break;
            case PTGPU_CYLINDER: lim = s->numCylinders; break;
            case PTGPU_MESH: lim = s->numMeshes; break;
            case PTGPU_TRANSFORMED: lim = s->numInstances; break;
            case PTGPU_SDF: lim = s->numSdfShapes; break;
            case PTGPU_VOLUME: lim = s->numVolumes; break;
            case PTGPU_SH: lim = s->numShs; break;
            default: return bad("unknown shape type", i);
        }
        if (sh.data >= lim) return bad("shape data index out of range", i);
        if (!mat_ok(sh.material)) return bad("shape material out of range", i);
    }
    for (uint32_t i = 0; i < s->numLights; i++) if (s->lights[i] >= s->numSceneShapes) return bad("light index out of range", i);
    for (uint32_t i = 0; i < s->numInstances; i++) if (s->instances[i].shape >= s->numShapes) return bad("instance shape out of range", i);
    for (uint32_t i = 0; i < s->numMeshes; i++) {
        const ptgpu_mesh& m = s->meshes[i];
        if (m.tree >= s->numTrees || m.tree == s->sceneTree) return bad("mesh tree out of range", i);
        if ((uint64_t)m.triFirst + m.triCount > s->numTriangles) return bad("mesh triangle range out of range", i);
    }
    for (uint64_t i = 0; i < s->numTriangles; i++) {
        if (!mat_ok(s->triShade[i].material)) return bad("triangle material out of range", i);
    }
    for (uint32_t i = 0; i < s->numShs; i++) {
        const ptgpu_sh& h = s->shs[i];
        if (h.mesh >= s->numMeshes || !mat_ok(h.positiveMaterial) || !mat_ok(h.negativeMaterial)) return bad("SphericalHarmonic mesh / material out of range", i);
        if (!sh_supported(h.l, h.m)) return bad("unsupported spherical harmonic (l <= 4, |m| <= l)", i);
    }
    for (uint32_t i = 0; i < s->numTrees; i++) if (s->trees[i].root >= s->numNodes) return bad("tree root out of range", i);
    // which tree owns a node decides what its leaf items index: nodes of a tree are contiguous from its root to the next root
    std::vector<std::pair<uint32_t, uint32_t>> roots;  // (root, tree)
    for (uint32_t i = 0; i < s->numTrees; i++) roots.push_back({s->trees[i].root, i});
    std::sort(roots.begin(), roots.end());
    for (size_t r = 0; r < roots.size(); r++) {
        const uint64_t b = roots[r].first, e = r + 1 < roots.size() ? roots[r + 1].first : s->numNodes;
        const bool sceneTree = roots[r].second == s->sceneTree;
        for (uint64_t i = b; i < e; i++) {
            const ptgpu_node& n = s->nodes[i];
            if (n.a & 3u) {
                if ((n.a >> 2) <= i || (n.a >> 2) >= e || n.b <= i || n.b >= e) return bad("kd node child out of its tree (children follow their parent)", i);
            } else {
                if ((uint64_t)(n.a >> 2) + n.b > s->numLeafItems) return bad("kd leaf range out of range", i);
                for (uint32_t k = 0; k < n.b; k++) {
                    const uint32_t item = s->leafItems[(n.a >> 2) + k];
                    if (sceneTree ? item >= s->numSceneShapes : item >= s->numTriangles) return bad("kd leaf item out of range", i);
                }
            }
        }
    }
    for (uint32_t i = 0; i < s->numSdfShapes; i++)
        if ((uint64_t)s->sdfShapes[i].progFirst + s->sdfShapes[i].progCount > s->numSdfOps) return bad("SDF program range out of range", i);
    for (uint32_t i = 0; i < s->numVolumes; i++) {
        const ptgpu_volume& v = s->volumes[i];
        if (v.w <= 0 || v.h <= 0 || v.d <= 0) return bad("volume dimensions", i);
        if ((uint64_t)v.windowFirst + v.windowCount > s->numVolumeWindows) return bad("volume window range out of range", i);
        if (v.dataOffset + (uint64_t)v.w * v.h * v.d > s->numVolumeData) return bad("volume data range out of range", i);
    }
    for (uint32_t i = 0; i < s->numVolumeWindows; i++) if (!mat_ok(s->volumeWindows[i].material)) return bad("volume window material out of range", i);
    auto tex_ok = [&](int32_t t) { return t >= -1 && (t < 0 || (uint32_t)t < s->numTextures); };
    for (uint32_t i = 0; i < s->numMaterials; i++) {
        const ptgpu_material& m = s->materials[i];
        if (!tex_ok(m.texture) || !tex_ok(m.normalTexture) || !tex_ok(m.bumpTexture) || !tex_ok(m.glossTexture)) return bad("material texture out of range", i);
    }
    if (!tex_ok(s->envTexture)) return "flat scene: envTexture out of range";
    for (uint32_t i = 0; i < s->numTextures; i++) {
        const ptgpu_texture& t = s->textures[i];
        if (t.width < 1 || t.height < 1 || t.texelOffset + (uint64_t)t.width * t.height > s->numTexels) return bad("texture texel range out of range", i);
    }
    return std::string();
}

static int upload_one(ptgpu_ctx* ctx, const ptgpu_flat_scene* s, MeshDerived& dv, bool derive);

int ptgpu_upload_scene(ptgpu_ctx* ctx, const ptgpu_flat_scene* s) {
    if (!ctx || !s) return PTGPU_E_ARG;
    if (s->abiVersion != PTGPU_ABI_VERSION) return fail(ctx, PTGPU_E_ARG, "flat scene ABI version mismatch");
    {
        const std::string why = validate_flat_scene(s);
        if (!why.empty()) return fail(ctx, PTGPU_E_ARG, why);
    }
    MeshDerived dv;  // derived once on the host (pinned staging of the root), copied to every device of the handle
    int rc = upload_one(ctx, s, dv, true);
    for (size_t k = 0; rc == PTGPU_OK && k < ctx->peers.size(); k++) {
        rc = upload_one(ctx->peers[k], s, dv, false);
        if (rc != PTGPU_OK) ctx->error = "device " + std::to_string(ctx->peers[k]->device) + ": " + ctx->peers[k]->error;
    }
    cudaSetDevice(ctx->device);
    return rc;
}

static int upload_one(ptgpu_ctx* ctx, const ptgpu_flat_scene* s, MeshDerived& dv, bool derive) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());  // every lane idle: the previous scene's buffers are reused, not freed
    free_scene(ctx);
    // limits the kernels were compiled with
    for (uint32_t i = 0; i < s->numTrees; i++) {
        uint32_t lim = (i == s->sceneTree) ? (uint32_t)kSceneStack : (uint32_t)kMeshStack;
        if (s->trees[i].maxDepth >= lim) return fail(ctx, PTGPU_E_LIMIT, "kd-tree deeper than the traversal stack (" + std::to_string(s->trees[i].maxDepth) + ")");
    }
    for (uint32_t i = 0; i < s->numInstances; i++) {  // TransformedShape of TransformedShape ...: a chain of at most kMaxInstanceDepth, flagged in pad[0]
        int depth = 1;
        for (ptgpu_shape sh = s->shapes[s->instances[i].shape]; sh.type == PTGPU_TRANSFORMED; sh = s->shapes[s->instances[sh.data].shape])
            if (++depth > kMaxInstanceDepth) return fail(ctx, PTGPU_E_LIMIT, "TransformedShape nested deeper than " + std::to_string(kMaxInstanceDepth));
        if ((depth > 1) != (s->instances[i].pad[0] != 0)) return fail(ctx, PTGPU_E_ARG, "ptgpu_instance.pad[0] must be 1 exactly when the inner shape is a TransformedShape");
    }
    for (uint32_t i = 0; i < s->numSdfShapes; i++) {  // validate SDF programs against the device stacks
        int nv = 0, np = 0;
        for (uint32_t k = 0; k < s->sdfShapes[i].progCount; k++) {
            const ptgpu_sdf_op& op = s->sdfOps[s->sdfShapes[i].progFirst + k];
            if (op.op >= PTGPU_SDF_SPHERE && op.op <= PTGPU_SDF_TORUS) nv++;
            else if (op.op >= PTGPU_SDF_PUSH_TRANSFORM && op.op <= PTGPU_SDF_PUSH_REPEAT) np++;
            else if (op.op == PTGPU_SDF_POP) np--;
            else if (op.op >= PTGPU_SDF_UNION && op.op <= PTGPU_SDF_INTERSECTION) nv -= (int)op.n - 1;
            else return fail(ctx, PTGPU_E_ARG, "unknown SDF op");
            if (nv > kSdfValueStack || np > kSdfPointStack || nv < 0 || np < 0) return fail(ctx, PTGPU_E_LIMIT, "SDF program exceeds the evaluation stacks");
        }
        if (nv != 1) return fail(ctx, PTGPU_E_ARG, "SDF program does not reduce to one value");
    }
    DScene& D = ctx->scene;
    std::memset(&D, 0, sizeof(D));
    int rc;
#define UP(field, src, cnt) if ((rc = upload(ctx, src, cnt, &D.field)) != PTGPU_OK) return rc
    UP(shapes, s->shapes, s->numShapes);
    UP(lights, s->lights, s->numLights);
    UP(trees, s->trees, s->numTrees);
    UP(nodes, s->nodes, s->numNodes);
    UP(leafItems, s->leafItems, s->numLeafItems);
    UP(spheres, s->spheres, s->numSpheres);
    UP(cubes, s->cubes, s->numCubes);
    UP(planes, s->planes, s->numPlanes);
    UP(cylinders, s->cylinders, s->numCylinders);
    UP(meshes, s->meshes, s->numMeshes);
    {
        const float4* g = nullptr;
        if ((rc = upload(ctx, reinterpret_cast<const float4*>(s->triGeom), s->numTriangles * 3, &g)) != PTGPU_OK) return rc;
        D.triGeom = g;
    }
    UP(triShade, s->triShade, s->numTriangles);
    UP(instances, s->instances, s->numInstances);
    {   // padded world-space bounds of every instanced mesh (see scene_advance): the 8 corners of Mesh.tree's Box through Matrix
        std::vector<float> ib((size_t)(s->numInstances ? s->numInstances : 1) * 8, 0.f);
        for (uint32_t i = 0; i < s->numInstances; i++) {
            const ptgpu_instance& in = s->instances[i];
            const ptgpu_shape& inner = s->shapes[in.shape];
            float* o8 = &ib[(size_t)i * 8];
            double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
            if (inner.type == PTGPU_MESH || inner.type == PTGPU_SH) {
                const ptgpu_tree& t = s->trees[s->meshes[inner.type == PTGPU_SH ? s->shs[inner.data].mesh : inner.data].tree];
                for (int c = 0; c < 8; c++) {
                    const double p[3] = {(c & 1) ? t.bmax[0] : t.bmin[0], (c & 2) ? t.bmax[1] : t.bmin[1], (c & 4) ? t.bmax[2] : t.bmin[2]};
                    for (int r = 0; r < 3; r++) {
                        const double v = in.m[4 * r] * p[0] + in.m[4 * r + 1] * p[1] + in.m[4 * r + 2] * p[2] + in.m[4 * r + 3];
                        lo[r] = std::min(lo[r], v); hi[r] = std::max(hi[r], v);
                    }
                }
                const double ext = std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2]));
                for (int r = 0; r < 3; r++) {
                    const double padv = 1e-4 * ext + 1e-5 * std::max(std::fabs(lo[r]), std::fabs(hi[r])) + 1e-6;
                    o8[r] = std::nextafter((float)(lo[r] - padv), -INFINITY); o8[4 + r] = std::nextafter((float)(hi[r] + padv), INFINITY);
                }
            } else {
                for (int r = 0; r < 3; r++) { o8[r] = -3.0e38f; o8[4 + r] = 3.0e38f; }  // no Box test in the reference: never skipped
            }
        }
        const float4* dib = nullptr;
        if ((rc = upload(ctx, reinterpret_cast<const float4*>(ib.data()), (uint64_t)ib.size() / 4, &dib)) != PTGPU_OK) return rc;
        CK(cudaStreamSynchronize(ctx->stream));
        D.instBounds = dib;
        // Candidate mask of scene_advance: the instanced meshes among the first kMaskShapes scene shapes in Morton order of their
        // bounds' centres, 16 to a block under the union of their bounds; every other scene shape is a candidate from the start.
        struct Cand { uint32_t shape, inst; uint64_t key; };
        std::vector<Cand> cands;
        uint32_t base[8] = {0, 0, 0, 0, 0, 0, 0, 0x80000000u};
        const ptgpu_tree& stree = s->trees[s->sceneTree];
        auto spread = [](uint64_t v) { v &= 0x1FFFFF; v = (v | v << 32) & 0x1F00000000FFFFull; v = (v | v << 16) & 0x1F0000FF0000FFull; v = (v | v << 8) & 0x100F00F00F00F00Full;
                                       v = (v | v << 4) & 0x10C30C30C30C30C3ull; v = (v | v << 2) & 0x1249249249249249ull; return v; };
        for (uint32_t k = 0; k < s->numSceneShapes && k < kMaskShapes; k++) {
            const ptgpu_shape& sh = s->shapes[k];
            const float* b8 = sh.type == PTGPU_TRANSFORMED ? &ib[(size_t)sh.data * 8] : nullptr;
            if (!b8 || b8[0] < -1e38f) { base[k >> 5] |= 1u << (k & 31); continue; }
            uint64_t key = 0;
            for (int r = 0; r < 3; r++) {
                const double ext = std::max(1e-30, (double)stree.bmax[r] - (double)stree.bmin[r]);
                const double u = std::min(1.0, std::max(0.0, (0.5 * ((double)b8[r] + (double)b8[4 + r]) - (double)stree.bmin[r]) / ext));
                key |= spread((uint64_t)(u * 2097151.0)) << r;
            }
            cands.push_back(Cand{k, sh.data, key});
        }
        std::sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) { return a.key < b.key || (a.key == b.key && a.shape < b.shape); });
        auto u2f = [](uint32_t v) { float f; std::memcpy(&f, &v, 4); return f; };
        std::vector<float> blocks, members;
        for (size_t f0 = 0; f0 < cands.size(); f0 += 16) {
            const size_t f1 = std::min(cands.size(), f0 + 16);
            float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
            for (size_t m = f0; m < f1; m++) {
                const float* b8 = &ib[(size_t)cands[m].inst * 8];
                for (int r = 0; r < 3; r++) { lo[r] = std::min(lo[r], b8[r]); hi[r] = std::max(hi[r], b8[4 + r]); }
                const float rec[8] = {b8[0], b8[1], b8[2], u2f(cands[m].shape), b8[4], b8[5], b8[6], 0.f};
                members.insert(members.end(), rec, rec + 8);
            }
            const float rec[8] = {std::nextafter(lo[0], -INFINITY), std::nextafter(lo[1], -INFINITY), std::nextafter(lo[2], -INFINITY), u2f((uint32_t)f0),
                                  std::nextafter(hi[0], INFINITY), std::nextafter(hi[1], INFINITY), std::nextafter(hi[2], INFINITY), u2f((uint32_t)(f1 - f0))};
            blocks.insert(blocks.end(), rec, rec + 8);
        }
        D.numCandBlocks = (uint32_t)(blocks.size() / 8);
        if (blocks.empty()) { blocks.assign(8, 0.f); members.assign(8, 0.f); }
        const float4 *dcb = nullptr, *dcm = nullptr;
        if ((rc = upload(ctx, reinterpret_cast<const float4*>(blocks.data()), (uint64_t)blocks.size() / 4, &dcb)) != PTGPU_OK) return rc;
        if ((rc = upload(ctx, reinterpret_cast<const float4*>(members.data()), (uint64_t)members.size() / 4, &dcm)) != PTGPU_OK) return rc;
        D.candBlocks = dcb; D.candMembers = dcm;
        std::memcpy(D.maskBase, base, sizeof(base));
        // per Scene.tree node: the shapes of the leaf as a mask (bit 255: it holds shapes without a bit of their own)
        uint64_t end = s->numNodes;
        for (uint32_t k = 0; k < s->numTrees; k++) if (s->trees[k].root > stree.root && s->trees[k].root < end) end = s->trees[k].root;
        std::vector<uint32_t> lm((size_t)(end - stree.root) * 8, 0u);
        for (uint64_t i = stree.root; i < end; i++) {
            const ptgpu_node& n = s->nodes[i];
            if ((n.a & 3u) != 0) continue;
            uint32_t* w = &lm[(size_t)(i - stree.root) * 8];
            for (uint32_t k = 0; k < n.b; k++) {
                const uint32_t item = s->leafItems[(n.a >> 2) + k];
                if (item < kMaskShapes) w[item >> 5] |= 1u << (item & 31); else w[7] |= 0x80000000u;
            }
        }
        // worth its upkeep (two more quads of per-ray state, a bit test per leaf item) when the leaves repeat their shapes
        uint64_t sceneItems = 0;
        for (uint64_t i = stree.root; i < end; i++) if ((s->nodes[i].a & 3u) == 0) sceneItems += s->nodes[i].b;
        static const char* maskEnv = std::getenv("PTGPU_SCENE_MASK");  // development: 0 / 1 force it off / on
        D.hasNested = 0;
        for (uint32_t i = 0; i < s->numInstances; i++) if (s->instances[i].pad[0]) D.hasNested = 1;
        {   // which scene-kernel instantiation this scene needs
            int tier = TIER_ANALYTIC;
            for (uint32_t k = 0; k < s->numSceneShapes; k++) {
                const uint32_t ty = s->shapes[k].type;
                if (ty == PTGPU_MESH) tier = std::max(tier, (int)TIER_MESH);
                else if (ty != PTGPU_SPHERE && ty != PTGPU_CUBE && ty != PTGPU_PLANE && ty != PTGPU_CYLINDER) tier = TIER_FULL;
            }
            ctx->sceneTier = tier;
        }
        D.maskOn = maskEnv ? (uint32_t)std::atoi(maskEnv) : (sceneItems >= 16 && 2 * sceneItems >= 3 * (uint64_t)s->numSceneShapes ? 1u : 0u);
        if (D.maskOn && ctx->sceneTier < TIER_FULL) ctx->sceneTier = TIER_FULL;  // the mask lives in the full instantiations
        if (D.hasNested) ctx->sceneTier = TIER_NESTED;
        ctx->shadeTier = (s->numTextures > 0 || s->envTexture >= 0 || ctx->sceneTier >= TIER_FULL) ? 2 : ctx->sceneTier;
        {   // shade order (shade_bin): kShadeBins bins = surfaces x patches.  A scene with few surfaces gets finer patches (C3, two meshes:
            // 16 x 256, +2 % over 128 x 32), a scene with hundreds of instances keeps them apart (C4: 128 x 32).
            uint32_t surfaces = 16;
            while (surfaces < (uint32_t)kSurfaceBins && surfaces < 2u * (s->numSceneShapes + 1u)) surfaces <<= 1;
            D.shadeSurfaces = surfaces; D.shadeSub = (uint32_t)kShadeBins / surfaces;
        }
        const uint4* dlm = nullptr;
        if ((rc = upload(ctx, reinterpret_cast<const uint4*>(lm.data()), (uint64_t)lm.size() / 4, &dlm)) != PTGPU_OK) return rc;
        CK(cudaStreamSynchronize(ctx->stream));
        D.sceneLeafMask = dlm;
    }
    UP(sdfShapes, s->sdfShapes, s->numSdfShapes);
    UP(sdfOps, s->sdfOps, s->numSdfOps);
    UP(volumes, s->volumes, s->numVolumes);
    UP(shs, s->shs, s->numShs);
    UP(volumeWindows, s->volumeWindows, s->numVolumeWindows);
    UP(volumeData, s->volumeData, s->numVolumeData);
    {   // per Volume: the largest voxel of every 4x4x4 block, dilated by two voxels, never below 0 (voxels outside the grid read 0);
        // a NaN voxel makes its blocks +inf (never skipped).  See vol_skip in pt_device.cuh.
        std::vector<VolBlocks> vbs(s->numVolumes ? s->numVolumes : 1);
        std::vector<double> bmax;
        for (uint32_t i = 0; i < s->numVolumes; i++) {
            const ptgpu_volume& v = s->volumes[i];
            VolBlocks vb;
            vb.first = bmax.size(); vb.pad = 0;
            vb.nbx = (v.w + kVolBlock - 1) / kVolBlock; vb.nby = (v.h + kVolBlock - 1) / kVolBlock; vb.nbz = (v.d + kVolBlock - 1) / kVolBlock;
            const double* data = s->volumeData + v.dataOffset;
            for (int bz = -1; bz <= vb.nbz; bz++)
                for (int by = -1; by <= vb.nby; by++)
                    for (int bx = -1; bx <= vb.nbx; bx++) {
                        double mx = 0;
                        for (int z = std::max(0, bz * kVolBlock - 2); z <= std::min(v.d - 1, bz * kVolBlock + kVolBlock + 2); z++)
                            for (int y = std::max(0, by * kVolBlock - 2); y <= std::min(v.h - 1, by * kVolBlock + kVolBlock + 2); y++)
                                for (int x = std::max(0, bx * kVolBlock - 2); x <= std::min(v.w - 1, bx * kVolBlock + kVolBlock + 2); x++) {
                                    const double q = data[(size_t)x + (size_t)y * v.w + (size_t)z * v.w * v.h];
                                    if (!(q == q)) mx = INFINITY; else if (q > mx) mx = q;
                                }
                        bmax.push_back(mx);
                    }
            vbs[i] = vb;
        }
        if (bmax.empty()) bmax.push_back(0);
        const VolBlocks* dvb = nullptr;
        if ((rc = upload(ctx, vbs.data(), (uint64_t)vbs.size(), &dvb)) != PTGPU_OK) return rc;
        const double* dbm = nullptr;
        if ((rc = upload(ctx, bmax.data(), (uint64_t)bmax.size(), &dbm)) != PTGPU_OK) return rc;
        CK(cudaStreamSynchronize(ctx->stream));  // the vectors are locals
        D.volBlocks = dvb; D.volBlockMax = dbm;
    }
    UP(materials, s->materials, s->numMaterials);
    UP(textures, s->textures, s->numTextures);
    {
        const double4* t = nullptr;
        if ((rc = upload(ctx, reinterpret_cast<const double4*>(s->texels), s->numTexels, &t)) != PTGPU_OK) return rc;
        D.texels = t;
    }
#undef UP
    {   // derived mesh data (see derive_mesh): 64-byte node records, sorted leaf triangles, patched tree roots
        if (derive) {
            std::string err;
            ctx->stageNext = 0;  // pinned staging owned by the handle: the derivation writes straight into DMA-able memory
            auto t0 = std::chrono::steady_clock::now();
            if (!derive_mesh(s, dv, err, [&](uint64_t bytes) -> void* { return stage_alloc(ctx, bytes); })) return fail(ctx, PTGPU_E_LIMIT, err);
            ctx->deriveMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
        {
            const ptgpu_tree* dt = nullptr;
            if ((rc = upload(ctx, dv.trees.data(), (uint64_t)dv.trees.size(), &dt)) != PTGPU_OK) return rc;
            D.trees = dt;
        }
        const uint4* dmn = nullptr;
        if ((rc = upload(ctx, reinterpret_cast<const uint4*>(dv.mn), dv.mnRecords * 4, &dmn)) != PTGPU_OK) return rc;
        D.meshNodes = dmn;
        ctx->meshNodeBytes = dv.mnRecords * 64;
        const float4* dl = nullptr;
        if ((rc = upload(ctx, reinterpret_cast<const float4*>(dv.lg), dv.lgCount * 3, &dl)) != PTGPU_OK) return rc;
        CK(cudaStreamSynchronize(ctx->stream));  // dv.trees is a local; the pinned staging is rewritten by the next upload
        D.leafGeom = dl;
    }
    D.sceneTree = s->sceneTree; D.numSceneShapes = s->numSceneShapes; D.numLights = s->numLights; D.numShapes = s->numShapes;
    D.envColor[0] = s->envColor[0]; D.envColor[1] = s->envColor[1]; D.envColor[2] = s->envColor[2];
    D.envTexture = s->envTexture; D.envTextureAngle = s->envTextureAngle;
    // Light geometry (Sampler.cs:215-236), evaluated once with the reference's Vector rounding.
    std::vector<DLight> lights(s->numLights);
    for (uint32_t i = 0; i < s->numLights; i++) {
        const ptgpu_shape& sh = s->shapes[s->lights[i]];
        DLight L;
        std::memset(&L, 0, sizeof(L));
        L.shape = s->lights[i];
        L.classTyped = sh.flags & 1u;
        if (sh.type == PTGPU_SPHERE) {
            const ptgpu_sphere& sp = s->spheres[sh.data];
            L.center[0] = sp.center[0]; L.center[1] = sp.center[1]; L.center[2] = sp.center[2];
            L.radius = sp.radius;
        } else if (sh.type == PTGPU_CYLINDER) {
            const ptgpu_cylinder& cy = s->cylinders[sh.data];
            L.center[2] = (float)((cy.z0 + cy.z1) / 2); L.radius = cy.radius; L.isCylinder = 1;
        } else {
            float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
            if (sh.type == PTGPU_CUBE) { std::memcpy(mn, s->cubes[sh.data].min, 12); std::memcpy(mx, s->cubes[sh.data].max, 12); }
            else if (sh.type == PTGPU_PLANE) { for (int k = 0; k < 3; k++) { mn[k] = -1e9f; mx[k] = 1e9f; } }
            else if (sh.type == PTGPU_SDF) { std::memcpy(mn, s->sdfShapes[sh.data].bmin, 12); std::memcpy(mx, s->sdfShapes[sh.data].bmax, 12); }
            else if (sh.type == PTGPU_VOLUME) { std::memcpy(mn, s->volumes[sh.data].bmin, 12); std::memcpy(mx, s->volumes[sh.data].bmax, 12); }
            else if (sh.type == PTGPU_SH) { for (int k = 0; k < 3; k++) { mn[k] = -1.f; mx[k] = 1.f; } }  // SH.cs:29-33
            // struct-typed lights (Mesh, TransformedShape) never pass the identity test; geometry is irrelevant
            // Box.Center / OuterRadius (Box.cs:46-50): Min + (Max-Min)*0.5 ; |Min - Center|
            float c[3], dlt[3];
            for (int k = 0; k < 3; k++) {
                float size = mx[k] - mn[k];
                float half = (float)((double)size * 0.5);
                c[k] = mn[k] + half;
                dlt[k] = mn[k] - c[k];
            }
            float s2 = dlt[0] * dlt[0] + dlt[1] * dlt[1];
            s2 = s2 + dlt[2] * dlt[2];
            L.center[0] = c[0]; L.center[1] = c[1]; L.center[2] = c[2];
            L.radius = (double)std::sqrt(s2);
        }
        lights[i] = L;
    }
    {
        const DLight* dl = nullptr;
        if ((rc = upload(ctx, lights.data(), lights.size(), &dl)) != PTGPU_OK) return rc;
        ctx->dLights = const_cast<DLight*>(dl);
    }
    CK(cudaStreamSynchronize(ctx->stream));
    {   // bound on the deferred shapes (Mesh / SDFShape / Volume, directly or instanced) a ray can enter = such items over all
        // leaves of Scene.tree (a ray visits a leaf, and a shape of a leaf, at most once)
        const ptgpu_tree& stree = s->trees[s->sceneTree];
        uint64_t end = s->numNodes;
        for (uint32_t k = 0; k < s->numTrees; k++) if (s->trees[k].root > stree.root && s->trees[k].root < end) end = s->trees[k].root;
        uint64_t deferred = 0;
        bool seen[256] = {false};
        if (seen[0]) return PTGPU_E_STATE;  // (never: keeps the array referenced in builds without the mask)
        ctx->hasKind[0] = ctx->hasKind[1] = ctx->hasKind[2] = false;
        for (uint64_t i = stree.root; i < end; i++) {
            const ptgpu_node& n = s->nodes[i];
            if ((n.a & 3u) != 0) continue;
            for (uint32_t k = 0; k < n.b; k++) {
                ptgpu_shape sh = s->shapes[s->leafItems[(n.a >> 2) + k]];
                while (sh.type == PTGPU_TRANSFORMED) sh = s->shapes[s->instances[sh.data].shape];
                const int kind = (sh.type == PTGPU_MESH || sh.type == PTGPU_SH) ? 0 : sh.type == PTGPU_SDF ? 1 : sh.type == PTGPU_VOLUME ? 2 : -1;
                if (kind >= 0) {
                    ctx->hasKind[kind] = true;
#if PT_SCENE_MASK
                    const uint32_t item = s->leafItems[(n.a >> 2) + k];  // a shape with a mask bit is evaluated once per ray, whatever the number of leaves it sits in
                    if (ctx->scene.maskOn && item < kMaskShapes) { if (!seen[item]) { seen[item] = true; deferred++; } } else deferred++;
#else
                    deferred++;
#endif
                }
            }
        }
        ctx->splitRounds = deferred <= 4 ? (int)deferred : -1;
        ctx->splitStackEnt = (int)stree.maxDepth + 2;
    }
    // The mesh nodes are re-read by every ray while hundreds of MB of queue records stream through the L2 between two
    // visits: pin (a fraction of) the node array in L2 with an access-policy window on every stream that launches tracers.
    {
        // Measured on C3 (B200): a 4-32 MB window is neutral, 64 MB and more is 15-35 % slower (the carve-out starves the
        // triangles), so the window is opt-in (PTGPU_L2_WINDOW_MB).
        static const bool off = std::getenv("PTGPU_L2_WINDOW_MB") == nullptr;
        cudaDeviceProp prop;
        if (!off && ctx->meshNodeBytes && cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
            size_t persist = (size_t)prop.persistingL2CacheMaxSize;
            if (const char* ev = std::getenv("PTGPU_L2_WINDOW_MB")) persist = std::min<size_t>(persist, (size_t)std::atoi(ev) << 20);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist);
            cudaStreamAttrValue attr;
            std::memset(&attr, 0, sizeof(attr));
            const size_t win = (size_t)std::min<uint64_t>(std::min<uint64_t>(ctx->meshNodeBytes, persist), (uint64_t)prop.accessPolicyMaxWindowSize);  // the head of the array = the upper levels
            attr.accessPolicyWindow.base_ptr = const_cast<uint4*>(ctx->scene.meshNodes);
            attr.accessPolicyWindow.num_bytes = win;
            attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)persist / (double)win);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
            for (int k = 0; k < ctx->numLanes; k++) cudaStreamSetAttribute(ctx->lanes[k].stream, cudaStreamAttributeAccessPolicyWindow, &attr);
            cudaGetLastError();
        }
    }
    trim_scene_pool(ctx);
    ctx->haveScene = true;
    return PTGPU_OK;
}

static uint64_t pow2_at_least(uint64_t v) { uint64_t p = 1; while (p < v) p <<= 1; return p; }
// Queues of one lane for batches of up to `needRays` path records and `needShadow` shadow records (grown, never shrunk).
static int ensure_queues(ptgpu_ctx* ctx, Lane& L, uint64_t needRays, uint64_t needShadow) {
    if (needRays > L.capRays) {
        const uint64_t cap = std::min<uint64_t>(pow2_at_least(needRays), std::max<uint64_t>(ctx->capRays, needRays));
        CK(cudaStreamSynchronize(L.stream));
        for (int i = 0; i < 2; i++) { cudaFree(L.rq[i].od0); cudaFree(L.rq[i].od1); cudaFree(L.rq[i].bt); cudaFree(L.rq[i].smp); L.rq[i] = RayQueue{}; }
        cudaFree(L.hq.t); cudaFree(L.hq.tInner); cudaFree(L.hq.shape); cudaFree(L.hq.prim); L.hq = HitQueue{};
        L.capRays = 0;
        for (int i = 0; i < 2; i++) {
            CK(cudaMalloc(&L.rq[i].od0, cap * sizeof(float4)));
            CK(cudaMalloc(&L.rq[i].od1, cap * sizeof(float4)));
            CK(cudaMalloc(&L.rq[i].bt, cap * sizeof(float4)));
            CK(cudaMalloc(&L.rq[i].smp, cap * sizeof(uint32_t)));
        }
        CK(cudaMalloc(&L.hq.t, cap * sizeof(double)));
        CK(cudaMalloc(&L.hq.tInner, cap * sizeof(double)));
        CK(cudaMalloc(&L.hq.shape, cap * sizeof(int32_t)));
        CK(cudaMalloc(&L.hq.prim, cap * sizeof(int32_t)));
        cudaFree(L.perm); L.perm = nullptr;
        CK(cudaMalloc(&L.perm, cap * sizeof(uint32_t)));
        if (!L.bins) CK(cudaMalloc(&L.bins, kShadeBins * sizeof(uint32_t)));
        L.capRays = cap;
    }
    if (needShadow > L.capShadow) {
        const uint64_t cap = pow2_at_least(needShadow);
        CK(cudaStreamSynchronize(L.stream));
        cudaFree(L.sq.so); cudaFree(L.sq.sd); cudaFree(L.sq.sc);
        L.sq = ShadowQueue{};
        L.capShadow = 0;
        CK(cudaMalloc(&L.sq.so, cap * sizeof(float4)));
        CK(cudaMalloc(&L.sq.sd, cap * sizeof(float4)));
        CK(cudaMalloc(&L.sq.sc, cap * sizeof(float4)));
        L.capShadow = cap;
    }
    return PTGPU_OK;
}

static int ensure_image(ptgpu_ctx* ctx, int w, int h) {
    if (ctx->bufW == w && ctx->bufH == h && ctx->dSum) return PTGPU_OK;
    free_image(ctx);
    size_t npix = (size_t)w * h;
    CK(cudaMalloc(&ctx->dSum, npix * 3 * sizeof(float)));
    CK(cudaMalloc(&ctx->dMean, npix * 3 * sizeof(float)));
    CK(cudaMalloc(&ctx->pb.M, npix * 3 * sizeof(double)));
    CK(cudaMalloc(&ctx->pb.V, npix * 3 * sizeof(double)));
    CK(cudaMalloc(&ctx->pb.samples, npix * sizeof(int32_t)));
    CK(cudaMalloc(&ctx->dList[0], npix * sizeof(uint32_t)));
    CK(cudaMalloc(&ctx->dList[1], npix * sizeof(uint32_t)));
    CK(cudaMalloc(&ctx->dReject, npix));
    CK(cudaMemsetAsync(ctx->pb.M, 0, npix * 3 * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->pb.V, 0, npix * 3 * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->pb.samples, 0, npix * sizeof(int32_t), ctx->stream));
    ctx->bufW = w; ctx->bufH = h;
    return PTGPU_OK;
}

static int make_passd(ptgpu_ctx* ctx, const ptgpu_pass* p, PassD& P) {
    if (p->width <= 1 || p->height <= 1 || p->spp <= 0) return fail(ctx, PTGPU_E_ARG, "width/height must be > 1 and spp > 0");
    if (p->maxBounces < 0 || p->maxBounces > 62) return fail(ctx, PTGPU_E_LIMIT, "maxBounces must be in [0, 62]");
    if (p->firstHitSamples < 1) return fail(ctx, PTGPU_E_ARG, "firstHitSamples must be >= 1");
    int nroot = (int)std::sqrt((double)p->firstHitSamples);
    int modes0 = (p->specularMode == PTGPU_SPECULAR_NAIVE) ? 1 : 2;
    if (nroot * nroot * modes0 > 4094) return fail(ctx, PTGPU_E_LIMIT, "firstHitSamples too large for the 12-bit first-hit index");
    // Philox addressing (rng_enter): one path bit per depth under SpecularModeAll; sub-streams 1..254 name the light under
    // LightModeAll (255 = Russian roulette); global sample indices from 2^20 on belong to the adaptive / firefly extra samples.
    if (p->specularMode == PTGPU_SPECULAR_ALL && p->maxBounces > 32) return fail(ctx, PTGPU_E_LIMIT, "SpecularModeAll: maxBounces must be <= 32 (one path bit per depth)");
    if (p->lightMode == PTGPU_LIGHT_ALL && p->directLighting && ctx->scene.numLights > 254) return fail(ctx, PTGPU_E_LIMIT, "LightModeAll: more than 254 lights");
    {
        const long long stride = p->sampleStride ? p->sampleStride : 1;
        const long long last = (long long)p->sampleBase + (long long)(p->spp - 1) * stride;
        if (p->sampleBase < 0 || stride < 1 || last >= (1ll << 20)) return fail(ctx, PTGPU_E_LIMIT, "global sample indices must stay in [0, 2^20): spp x ranks too large for one pass");
    }
    P.russianRoulette = (p->flags & PTGPU_PASS_RUSSIAN_ROULETTE) ? 1 : 0;
    P.width = p->width; P.height = p->height; P.spp = p->spp; P.stratified = p->stratified; P.subpixelJitter = 0;
    P.sppRoot = (int)std::sqrt((double)p->spp);
    P.sampleBase = p->sampleBase; P.sampleStride = p->sampleStride ? p->sampleStride : 1;
    P.firstHitSamples = p->firstHitSamples; P.maxBounces = p->maxBounces; P.directLighting = p->directLighting;
    P.softShadows = p->softShadows; P.lightMode = p->lightMode; P.specularMode = p->specularMode;
    P.seed = p->seed; P.passIndex = p->passIndex; P.cam = p->camera;
    return PTGPU_OK;
}

// Issue every kernel of one pass on `stream`, adding radiance into d_sum.  nSlots = samples per pixel rendered.
static int run_pass(ptgpu_ctx* ctx, const PassD& P, int nSlots, float* d_sum, cudaStream_t callerStream, const uint32_t* pixelList = nullptr,
                    const uint32_t* listCount = nullptr) {
    const uint64_t npix = (uint64_t)P.width * P.height;
    const uint64_t total = npix * (uint64_t)nSlots;
    // worst-case queue growth per camera sample: n^2 * modes at depth 0, x modes per later depth
    const int nroot = (int)std::sqrt((double)P.firstHitSamples);
    const uint64_t modes0 = (P.specularMode == PTGPU_SPECULAR_NAIVE) ? 1 : 2;
    const uint64_t modesN = (P.specularMode == PTGPU_SPECULAR_ALL) ? 2 : 1;
    const bool prof = ctx->profiling;
    // a profiled pass, and a scene whose rounds are polled from the host (splitRounds < 0: issued lane by lane anyway), use ONE lane
    // with the capacity of all of them
    static const bool detailEnv = std::getenv("PTGPU_TRACE_DETAIL") != nullptr;
    const bool polled = ctx->splitRounds < 0;  // the host reads a queue count per round: every lane is then driven by its own host thread
    const bool oneLane = prof || (polled && (detailEnv || pixelList != nullptr));
    const uint64_t laneCap = std::min<uint64_t>(oneLane ? ctx->capRays * (uint64_t)ctx->numLanes : ctx->capRays, 1ull << 30);
    uint64_t grow = 1, maxGrow = 1;
    for (int depth = 0; depth < P.maxBounces; depth++) {
        grow *= (depth == 0) ? (uint64_t)nroot * nroot * modes0 : modesN;
        if (grow > maxGrow) maxGrow = grow;
        if (maxGrow > laneCap) break;
    }
    // shadow rays spawned by one shade pass <= children of that pass * lights sampled per child
    const uint64_t lightsPer = (P.directLighting && ctx->scene.numLights) ? (P.lightMode == PTGPU_LIGHT_ALL ? ctx->scene.numLights : 1) : 0;
    uint64_t childGrow = maxGrow;
    {   // children of the deepest shade pass are not traced but still do NEE
        uint64_t g = 1;
        for (int depth = 0; depth <= P.maxBounces; depth++) { g *= (depth == 0) ? (uint64_t)nroot * nroot * modes0 : modesN; if (g > childGrow) childGrow = g; if (g > (1ull << 40)) break; }
    }
    uint64_t batch = laneCap / maxGrow;
    if (batch == 0) return fail(ctx, PTGPU_E_LIMIT, "queue capacity too small for this sampler's branching factor");
    uint64_t capShadow = batch * childGrow * (lightsPer ? lightsPer : 1);
    // up to 4x the ray capacity (48-byte records); the tracer's per-ray state is sized for the rays and the shadow queue is traced in
    // chunks of that size
    const uint64_t shadowCeil = std::min<uint64_t>(laneCap * 4, 0xFFFF0000ull);  // slots are 32-bit
    if (capShadow > shadowCeil) {  // shrink the batch so the shadow queue stays bounded
        batch = shadowCeil / (childGrow * (lightsPer ? lightsPer : 1));
        if (batch == 0) return fail(ctx, PTGPU_E_LIMIT, "queue capacity too small for this sampler's light count");
        capShadow = batch * childGrow * (lightsPer ? lightsPer : 1);
    }
    if (capShadow == 0) capShadow = 1;
    if (batch > total) batch = total;
    // spread the pass over the lanes: at least one batch per lane when the pass is large enough to be worth it
    int lanesUsed = oneLane ? 1 : ctx->numLanes;
    {
        const uint64_t per = (total + (uint64_t)lanesUsed - 1) / (uint64_t)lanesUsed;
        const uint64_t minBatch = 1ull << 18;
        if (per < batch) batch = std::max<uint64_t>(per, std::min<uint64_t>(minBatch, batch));
        const uint64_t nBatches = (total + batch - 1) / batch;
        if (nBatches < (uint64_t)lanesUsed) lanesUsed = (int)nBatches;
    }
    int rc = PTGPU_OK;
    const uint64_t shadowThisBatch = std::max<uint64_t>(1, batch * childGrow * (lightsPer ? lightsPer : 1));  // most shadow records one shade launch can append
    for (int k = 0; k < lanesUsed; k++) {
        Lane& L = ctx->lanes[k];
        // queues grow with the largest batch seen so far (powers of two), so a small pass does not allocate the full capacity
        if ((rc = ensure_queues(ctx, L, batch * maxGrow, shadowThisBatch)) != PTGPU_OK) return rc;
        if ((rc = ensure_split(ctx, L, L.capRays, ctx->splitStackEnt)) != PTGPU_OK) return rc;
    }
    const int gridShade = grid_for(ctx, 8), gridGen = grid_for(ctx, 8), gridFinish = grid_for(ctx, 4);
    const int tier = ctx->sceneTier;  // which instantiation of the scene kernels runs (see scene_advance)
    // (per-batch timing locals live in run_batch)
    if (prof) {
        ctx->traceMs = ctx->shadeMs = ctx->shadowMs = ctx->raygenMs = 0; ctx->traceLaunches = 0;
        for (int k = 0; k < 3; k++) { ctx->kindMs[k] = 0; ctx->kindLaunches[k] = 0; }
        ctx->roundItems = 0;
        CK(cudaMemsetAsync(&ctx->dCounters->kindItems[0], 0, 3 * sizeof(unsigned long long), callerStream));
    }
    // fork: the lanes' streams continue from the caller's stream ...
    if (!prof) {
        CK(cudaEventRecord(ctx->evFork, callerStream));
        for (int k = 0; k < lanesUsed; k++) CK(cudaStreamWaitEvent(ctx->lanes[k].stream, ctx->evFork, 0));
    }
    auto run_batch = [&](uint64_t g0, uint64_t batchIndex) -> int {
        int rc = PTGPU_OK;
        float ms = 0;
        Lane& L = ctx->lanes[prof ? 0 : (int)(batchIndex % (uint64_t)lanesUsed)];
        cudaStream_t stream = prof ? callerStream : L.stream;
        uint32_t* counts = L.counts;
        uint32_t n = (uint32_t)std::min<uint64_t>(batch, total - g0);
        int cur = 0;
        if (prof) cudaEventRecord(ctx->evA, stream);
        k_raygen<<<gridGen, 256, 0, stream>>>(P, g0, n, (uint32_t)nSlots, L.rq[0], counts + 0, ctx->dCounters, pixelList, listCount);
        ctx->launches++;
        if (prof) { cudaEventRecord(ctx->evB, stream); cudaEventSynchronize(ctx->evB); cudaEventElapsedTime(&ms, ctx->evA, ctx->evB); ctx->raygenMs += ms; }
        for (int depth = 0; depth <= P.maxBounces; depth++) {
            CK(cudaMemsetAsync(counts + (cur ^ 1), 0, sizeof(uint32_t), stream));
            CK(cudaMemsetAsync(counts + 2, 0, sizeof(uint32_t), stream));
            CK(cudaMemsetAsync(counts + 4, 0, 2 * sizeof(uint32_t), stream));  // trace / shadow work cursors
            if (prof) cudaEventRecord(ctx->evA, stream);
            {
                const RayQueue rqc = L.rq[cur];
                uint32_t* cnt = counts + cur;
                rc = run_split(ctx, L, stream,
                               [&](const MeshQueue& out) { { if (tier == TIER_NESTED) k_scene_trace<SCENE_START, TIER_NESTED><<<gridShade, 128, 0, stream>>>(ctx->scene, L.split, rqc, cnt, out, out, L.hq, ctx->dCounters); else if (tier == TIER_FULL) k_scene_trace<SCENE_START, TIER_FULL><<<gridShade, 128, 0, stream>>>(ctx->scene, L.split, rqc, cnt, out, out, L.hq, ctx->dCounters); else if (tier == TIER_MESH) k_scene_trace<SCENE_START, TIER_MESH><<<gridShade, 128, 0, stream>>>(ctx->scene, L.split, rqc, cnt, out, out, L.hq, ctx->dCounters); else k_scene_trace<SCENE_START, TIER_ANALYTIC><<<gridShade, 128, 0, stream>>>(ctx->scene, L.split, rqc, cnt, out, out, L.hq, ctx->dCounters); } },
                               [&](const MeshQueue& in, const MeshQueue& out) { { if (tier == TIER_NESTED) k_scene_trace<SCENE_RESUME, TIER_NESTED><<<gridShade, 128, 0, stream>>>(ctx->scene, L.split, rqc, cnt, in, out, L.hq, ctx->dCounters); else if (tier == TIER_FULL) k_scene_trace<SCENE_RESUME, TIER_FULL><<<gridShade, 128, 0, stream>>>(ctx->scene, L.split, rqc, cnt, in, out, L.hq, ctx->dCounters); else if (tier == TIER_MESH) k_scene_trace<SCENE_RESUME, TIER_MESH><<<gridShade, 128, 0, stream>>>(ctx->scene, L.split, rqc, cnt, in, out, L.hq, ctx->dCounters); else k_scene_trace<SCENE_RESUME, TIER_ANALYTIC><<<gridShade, 128, 0, stream>>>(ctx->scene, L.split, rqc, cnt, in, out, L.hq, ctx->dCounters); } },
                               [&](const MeshQueue& in) { { if (tier == TIER_NESTED) k_scene_trace<SCENE_FINISH, TIER_NESTED><<<gridFinish, 128, 0, stream>>>(ctx->scene, L.split, rqc, cnt, in, in, L.hq, ctx->dCounters); else if (tier == TIER_FULL) k_scene_trace<SCENE_FINISH, TIER_FULL><<<gridFinish, 128, 0, stream>>>(ctx->scene, L.split, rqc, cnt, in, in, L.hq, ctx->dCounters); else if (tier == TIER_MESH) k_scene_trace<SCENE_FINISH, TIER_MESH><<<gridFinish, 128, 0, stream>>>(ctx->scene, L.split, rqc, cnt, in, in, L.hq, ctx->dCounters); else k_scene_trace<SCENE_FINISH, TIER_ANALYTIC><<<gridFinish, 128, 0, stream>>>(ctx->scene, L.split, rqc, cnt, in, in, L.hq, ctx->dCounters); } });
                if (rc != PTGPU_OK) return rc;
                if (prof) ctx->traceLaunches++;
            }
            if (prof) { cudaEventRecord(ctx->evB, stream); cudaEventSynchronize(ctx->evB); cudaEventElapsedTime(&ms, ctx->evA, ctx->evB); ctx->traceMs += ms; cudaEventRecord(ctx->evA, stream); }
            static const bool shadeOrder = !(std::getenv("PTGPU_SHADE_ORDER") && std::atoi(std::getenv("PTGPU_SHADE_ORDER")) == 0);  // development switch
            if (shadeOrder) {
                CK(cudaMemsetAsync(L.bins, 0, kShadeBins * sizeof(uint32_t), stream));
                k_bin_count<<<grid_for(ctx, 4), 256, 0, stream>>>(ctx->scene, L.hq.shape, L.hq.prim, counts + cur, L.bins);
                k_bin_scan<<<1, 1024, 0, stream>>>(L.bins);
                k_bin_scatter<<<grid_for(ctx, 4), 256, 0, stream>>>(ctx->scene, L.hq.shape, L.hq.prim, counts + cur, L.bins, L.perm);
                ctx->launches += 3;
            }
            {   // the instantiation for what the scene holds (shade tier, see tri_normal in pt_device.cuh)
                auto shade = ctx->shadeTier == 0 ? k_shade<0> : ctx->shadeTier == 1 ? k_shade<1> : k_shade<2>;
                shade<<<gridShade * 128 / PT_SHADE_BLOCK, PT_SHADE_BLOCK, 0, stream>>>(ctx->scene, P, ctx->dLights, L.rq[cur], counts + cur, L.hq, L.rq[cur ^ 1], counts + (cur ^ 1),
                                                                                     L.sq, counts + 2, d_sum, ctx->dCounters, (uint32_t)L.capRays, (uint32_t)L.capShadow,
                                                                                     shadeOrder ? L.perm : nullptr, counts + 3);
            }
            k_clamp_count<<<1, 1, 0, stream>>>(counts + (cur ^ 1), (uint32_t)L.capRays, counts + 3);
            if (prof) { cudaEventRecord(ctx->evB, stream); cudaEventSynchronize(ctx->evB); cudaEventElapsedTime(&ms, ctx->evA, ctx->evB); ctx->shadeMs += ms; cudaEventRecord(ctx->evA, stream); }
            ctx->launches += 2;
            if (lightsPer) {
                // in chunks of the tracer's per-ray state (the shadow queue of this launch can hold up to 4x the ray capacity)
                const uint32_t chunk = (uint32_t)std::min<uint64_t>(L.splitCap, 0xFFFFFFFFull);
                for (uint64_t first64 = 0; first64 < shadowThisBatch; first64 += chunk) {
                    const uint32_t first = (uint32_t)first64;
                    rc = run_split<PT_ANYHIT != 0>(ctx, L, stream,
                                   [&](const MeshQueue& out) { { if (tier == TIER_NESTED) k_scene_shadow<SCENE_START, TIER_NESTED><<<gridShade, 128, 0, stream>>>(ctx->scene, L.split, L.sq, counts + 2, (uint32_t)L.capShadow, first, chunk, out, out, d_sum, ctx->dCounters); else if (tier == TIER_FULL) k_scene_shadow<SCENE_START, TIER_FULL><<<gridShade, 128, 0, stream>>>(ctx->scene, L.split, L.sq, counts + 2, (uint32_t)L.capShadow, first, chunk, out, out, d_sum, ctx->dCounters); else if (tier == TIER_MESH) k_scene_shadow<SCENE_START, TIER_MESH><<<gridShade, 128, 0, stream>>>(ctx->scene, L.split, L.sq, counts + 2, (uint32_t)L.capShadow, first, chunk, out, out, d_sum, ctx->dCounters); else k_scene_shadow<SCENE_START, TIER_ANALYTIC><<<gridShade, 128, 0, stream>>>(ctx->scene, L.split, L.sq, counts + 2, (uint32_t)L.capShadow, first, chunk, out, out, d_sum, ctx->dCounters); } },
                                   [&](const MeshQueue& in, const MeshQueue& out) { { if (tier == TIER_NESTED) k_scene_shadow<SCENE_RESUME, TIER_NESTED><<<gridShade, 128, 0, stream>>>(ctx->scene, L.split, L.sq, counts + 2, (uint32_t)L.capShadow, first, chunk, in, out, d_sum, ctx->dCounters); else if (tier == TIER_FULL) k_scene_shadow<SCENE_RESUME, TIER_FULL><<<gridShade, 128, 0, stream>>>(ctx->scene, L.split, L.sq, counts + 2, (uint32_t)L.capShadow, first, chunk, in, out, d_sum, ctx->dCounters); else if (tier == TIER_MESH) k_scene_shadow<SCENE_RESUME, TIER_MESH><<<gridShade, 128, 0, stream>>>(ctx->scene, L.split, L.sq, counts + 2, (uint32_t)L.capShadow, first, chunk, in, out, d_sum, ctx->dCounters); else k_scene_shadow<SCENE_RESUME, TIER_ANALYTIC><<<gridShade, 128, 0, stream>>>(ctx->scene, L.split, L.sq, counts + 2, (uint32_t)L.capShadow, first, chunk, in, out, d_sum, ctx->dCounters); } },
                                   [&](const MeshQueue& in) { { if (tier == TIER_NESTED) k_scene_shadow<SCENE_FINISH, TIER_NESTED><<<gridFinish, 128, 0, stream>>>(ctx->scene, L.split, L.sq, counts + 2, (uint32_t)L.capShadow, first, chunk, in, in, d_sum, ctx->dCounters); else if (tier == TIER_FULL) k_scene_shadow<SCENE_FINISH, TIER_FULL><<<gridFinish, 128, 0, stream>>>(ctx->scene, L.split, L.sq, counts + 2, (uint32_t)L.capShadow, first, chunk, in, in, d_sum, ctx->dCounters); else if (tier == TIER_MESH) k_scene_shadow<SCENE_FINISH, TIER_MESH><<<gridFinish, 128, 0, stream>>>(ctx->scene, L.split, L.sq, counts + 2, (uint32_t)L.capShadow, first, chunk, in, in, d_sum, ctx->dCounters); else k_scene_shadow<SCENE_FINISH, TIER_ANALYTIC><<<gridFinish, 128, 0, stream>>>(ctx->scene, L.split, L.sq, counts + 2, (uint32_t)L.capShadow, first, chunk, in, in, d_sum, ctx->dCounters); } });
                    if (rc != PTGPU_OK) return rc;
                }
                if (prof) { cudaEventRecord(ctx->evB, stream); cudaEventSynchronize(ctx->evB); cudaEventElapsedTime(&ms, ctx->evA, ctx->evB); ctx->shadowMs += ms; }
            }
            cur ^= 1;
        }
        return PTGPU_OK;
    };
    if (polled && lanesUsed > 1) {  // one host thread per lane: a lane's count polls block only its own thread
        std::vector<std::thread> workers;
        std::vector<int> rcs((size_t)lanesUsed, PTGPU_OK);
        for (int k = 0; k < lanesUsed; k++)
            workers.emplace_back([&, k] {
                cudaSetDevice(ctx->device);
                uint64_t bi = (uint64_t)k;
                for (uint64_t g0 = (uint64_t)k * batch; g0 < total && rcs[(size_t)k] == PTGPU_OK; g0 += batch * (uint64_t)lanesUsed, bi += (uint64_t)lanesUsed) rcs[(size_t)k] = run_batch(g0, bi);
            });
        for (auto& w : workers) w.join();
        for (int r : rcs) if (r != PTGPU_OK) return r;
    } else {
        uint64_t batchIndex = 0;
        for (uint64_t g0 = 0; g0 < total; g0 += batch, batchIndex++)
            if ((rc = run_batch(g0, batchIndex)) != PTGPU_OK) return rc;
    }
    // ... and the caller's stream continues after all of them (join)
    if (!prof) {
        for (int k = 0; k < lanesUsed; k++) {
            CK(cudaEventRecord(ctx->lanes[k].done, ctx->lanes[k].stream));
            CK(cudaStreamWaitEvent(callerStream, ctx->lanes[k].done, 0));
        }
    }
    CK(cudaGetLastError());
    return PTGPU_OK;
}

// Queue-overflow flags (counts[3] of every lane; set by k_clamp_count for the ray queue and by k_shade for the shadow queue).
static int clear_overflow(ptgpu_ctx* ctx, cudaStream_t st) {
    for (int k = 0; k < ctx->numLanes; k++) CK(cudaMemsetAsync(ctx->lanes[k].counts + 3, 0, sizeof(uint32_t), st));
    return PTGPU_OK;
}
// After the work is known to be complete: true when a queue overflowed since the last clear.
static bool take_overflow(ptgpu_ctx* ctx) {
    uint32_t h[kMaxLanes * 16];
    if (cudaMemcpy(h, ctx->laneCounts, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return false; }
    bool any = false;
    for (int k = 0; k < ctx->numLanes; k++) any = any || h[k * 16 + 3] != 0;
    if (any) {
        ctx->overflowPasses++;
        for (int k = 0; k < ctx->numLanes; k++) cudaMemset(ctx->lanes[k].counts + 3, 0, sizeof(uint32_t));
    }
    return any;
}

// NULL = the legacy default stream (the stream torch's default stream is): the pass is ordered after everything already queued on
// the caller's stream (e.g. the zero-fill of d_sum_rgb) and the stream continues after it (e.g. with the NCCL reduce).
static cudaStream_t caller_stream(void* stream) { return stream ? (cudaStream_t)stream : cudaStreamLegacy; }

int ptgpu_accumulate_device(ptgpu_ctx* ctx, const ptgpu_pass* pass, float* d_sum_rgb, void* stream) {
    if (!ctx || !pass || !d_sum_rgb) return PTGPU_E_ARG;
    if (!ctx->haveScene) return fail(ctx, PTGPU_E_STATE, "no scene uploaded");
    CK(cudaSetDevice(ctx->device));
    PassD P;
    int rc = make_passd(ctx, pass, P);
    if (rc != PTGPU_OK) return rc;
    cudaStream_t st = caller_stream(stream);
    CK(cudaEventRecord(ctx->ev0, st));
    int nSlots = P.stratified ? P.sppRoot * P.sppRoot : P.spp;
    rc = run_pass(ctx, P, nSlots, d_sum_rgb, st);
    if (rc != PTGPU_OK) return rc;
    CK(cudaEventRecord(ctx->ev1, st));
    return PTGPU_OK;
}

static SumSet one_sum(const float* p) { SumSet s; std::memset(&s, 0, sizeof(s)); s.p[0] = p; s.n = 1; return s; }

int ptgpu_add_sample_device(ptgpu_ctx* ctx, int32_t width, int32_t height, const float* d_sum_rgb, double divisor, void* stream) {
    if (!ctx || !d_sum_rgb) return PTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_image(ctx, width, height);
    if (rc != PTGPU_OK) return rc;
    cudaStream_t st = caller_stream(stream);
    if (st != ctx->stream) CK(cudaStreamSynchronize(ctx->stream));  // buffer allocation memsets
    k_add_sample<<<grid_for(ctx, 4), 256, 0, st>>>(one_sum(d_sum_rgb), divisor, (uint32_t)((size_t)width * height), ctx->pb, ctx->dMean, nullptr, 0, 1.f, 0);
    ctx->launches++;
    CK(cudaGetLastError());
    return PTGPU_OK;
}

// The pass accumulator of a peer device (the root's comes with ensure_image).
static int ensure_sum(ptgpu_ctx* ctx, int w, int h) {
    if (ctx->bufW == w && ctx->bufH == h && ctx->dSum) return PTGPU_OK;
    free_image(ctx);
    CK(cudaMalloc(&ctx->dSum, (size_t)w * h * 3 * sizeof(float)));
    ctx->bufW = w; ctx->bufH = h;
    return PTGPU_OK;
}

// The main pass of a multi-device handle: device k of N draws samples k, k + N, k + 2N, ... of the pass's sample sequence into its
// own accumulator (one host thread per peer: a pass may synchronise its stream between rounds), then the root's Buffer.AddSample
// kernel sums the N accumulators — the peers' through NVLink peer access — and applies Welford: reduce and update in one kernel.
static int render_main_multi(ptgpu_ctx* ctx, const PassD& P, cudaStream_t st) {
    const int N = 1 + (int)ctx->peers.size();
    const size_t npix = (size_t)P.width * P.height;
    std::vector<int> rcs((size_t)N, PTGPU_OK);
    auto share = [&](int k) { return P.spp / N + (k < P.spp % N ? 1 : 0); };
    auto run_one = [&](ptgpu_ctx* c, int k) -> int {
        if (cudaSetDevice(c->device) != cudaSuccess) { c->error = "cudaSetDevice failed"; return PTGPU_E_CUDA; }
        int rc = c == ctx ? PTGPU_OK : ensure_sum(c, P.width, P.height);
        if (rc != PTGPU_OK) return rc;
        if (cudaMemsetAsync(c->dSum, 0, npix * 3 * sizeof(float), c->stream) != cudaSuccess) { c->error = "cudaMemsetAsync failed"; return PTGPU_E_CUDA; }
        if ((rc = clear_overflow(c, c->stream)) != PTGPU_OK) return rc;
        PassD Q = P;
        Q.spp = share(k); Q.sampleBase = P.sampleBase + k * P.sampleStride; Q.sampleStride = P.sampleStride * N;
        if (Q.spp > 0 && (rc = run_pass(c, Q, Q.spp, c->dSum, c->stream)) != PTGPU_OK) return rc;
        if (c != ctx && cudaEventRecord(c->evPeerDone, c->stream) != cudaSuccess) { c->error = "cudaEventRecord failed"; return PTGPU_E_CUDA; }
        return PTGPU_OK;
    };
    std::vector<std::thread> threads;
    for (int k = 1; k < N; k++) threads.emplace_back([&, k] { rcs[(size_t)k] = run_one(ctx->peers[(size_t)k - 1], k); });
    rcs[0] = run_one(ctx, 0);
    for (auto& t : threads) t.join();
    CK(cudaSetDevice(ctx->device));
    for (int k = 0; k < N; k++)
        if (rcs[(size_t)k] != PTGPU_OK) {
            if (k > 0) ctx->error = "device " + std::to_string(ctx->peers[(size_t)k - 1]->device) + ": " + ctx->peers[(size_t)k - 1]->error;
            return rcs[(size_t)k];
        }
    SumSet sums;
    std::memset(&sums, 0, sizeof(sums));
    sums.p[0] = ctx->dSum; sums.n = N;
    for (int k = 1; k < N; k++) {
        ptgpu_ctx* c = ctx->peers[(size_t)k - 1];
        CK(cudaStreamWaitEvent(st, c->evPeerDone, 0));
        if (ctx->peerDirect[(size_t)k - 1]) sums.p[k] = c->dSum;
        else {  // no peer access between these two devices: one device-to-device copy into a root-side buffer
            float*& stage = ctx->peerStage[(size_t)k - 1];
            if (!stage) CK(cudaMalloc(&stage, npix * 3 * sizeof(float)));
            CK(cudaMemcpyPeerAsync(stage, ctx->device, c->dSum, c->device, npix * 3 * sizeof(float), st));
            sums.p[k] = stage;
        }
    }
    k_add_sample<<<grid_for(ctx, 4), 256, 0, st>>>(sums, (double)P.spp, (uint32_t)npix, ctx->pb, ctx->dMean, ctx->laneCounts + 3, ctx->numLanes, 1.f, 0);
    ctx->launches++;
    CK(cudaStreamSynchronize(st));  // the peers' accumulators are free again
    bool overflow = false;
    for (int k = 1; k < N; k++) {
        ptgpu_ctx* c = ctx->peers[(size_t)k - 1];
        cudaSetDevice(c->device);
        overflow = take_overflow(c) || overflow;
    }
    CK(cudaSetDevice(ctx->device));
    if (overflow) return fail(ctx, PTGPU_E_LIMIT, "ray queue overflow on a peer device (internal sizing error)");
    return PTGPU_OK;
}

int ptgpu_render_pass(ptgpu_ctx* ctx, const ptgpu_pass* pass, float* out_mean_rgb) {
    if (!ctx || !pass) return PTGPU_E_ARG;
    if (!ctx->haveScene) return fail(ctx, PTGPU_E_STATE, "no scene uploaded");
    CK(cudaSetDevice(ctx->device));
    PassD P;
    int rc = make_passd(ctx, pass, P);
    if (rc != PTGPU_OK) return rc;
    rc = ensure_image(ctx, P.width, P.height);
    if (rc != PTGPU_OK) return rc;
    const size_t npix = (size_t)P.width * P.height;
    cudaStream_t st = ctx->stream;
    const uint32_t* flags = ctx->laneCounts + 3;
    const int nflags = ctx->numLanes;
    CK(cudaEventRecord(ctx->ev0, st));
    if ((rc = clear_overflow(ctx, st)) != PTGPU_OK) return rc;
    if (P.stratified) {
        // Renderer.cs:231-246: every stratum sample is its own Buffer.AddSample (root device only: one sample per pixel at a time)
        int nn = P.sppRoot * P.sppRoot;
        for (int k = 0; k < nn; k++) {
            PassD Q = P;
            Q.sampleBase = P.sampleBase + k * P.sampleStride;
            CK(cudaMemsetAsync(ctx->dSum, 0, npix * 3 * sizeof(float), st));
            rc = run_pass(ctx, Q, 1, ctx->dSum, st);
            if (rc != PTGPU_OK) return rc;
            k_add_sample<<<grid_for(ctx, 4), 256, 0, st>>>(one_sum(ctx->dSum), 1.0, (uint32_t)npix, ctx->pb, ctx->dMean, flags, nflags, 1.f / (float)nn, k > 0);
            ctx->launches++;
        }
    } else if (!ctx->peers.empty()) {
        rc = render_main_multi(ctx, P, st);
        if (rc != PTGPU_OK) return rc;
    } else {
        CK(cudaMemsetAsync(ctx->dSum, 0, npix * 3 * sizeof(float), st));
        rc = run_pass(ctx, P, P.spp, ctx->dSum, st);
        if (rc != PTGPU_OK) return rc;
        k_add_sample<<<grid_for(ctx, 4), 256, 0, st>>>(one_sum(ctx->dSum), (double)P.spp, (uint32_t)npix, ctx->pb, ctx->dMean, flags, nflags, 1.f, 0);
        ctx->launches++;
    }
    if (out_mean_rgb) CK(cudaMemcpyAsync(out_mean_rgb, ctx->dMean, npix * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    // Extra samples (root device) use their own ranges of the global sample index so that no Philox stream is reused.
    const int kAdaptiveBase = 1 << 20, kFireflyBase = 1 << 21;
    if (pass->serialRules) {  // Renderer.cs:150-191: the extra samples of the serial Render()
        if (pass->adaptiveSamples > 0 && pass->adaptiveExponent < 0) return fail(ctx, PTGPU_E_ARG, "serialRules: a negative AdaptiveExponent is not supported");
        uint32_t* lc = ctx->dCounts + 8;  // [8], [9]: list counts
        for (int stage = 0; stage < 2; stage++) {
            const int nExtra = stage == 0 ? pass->adaptiveSamples : pass->fireflySamples;
            if (nExtra <= 0) continue;
            PassD Q = P;
            Q.stratified = 0; Q.subpixelJitter = stage == 0 ? 1 : 2; Q.sampleStride = 1;
            CK(cudaMemsetAsync(lc, 0, 2 * sizeof(uint32_t), st));
            CK(cudaMemsetAsync(ctx->dSum, 0, npix * 3 * sizeof(float), st));
            if (stage == 0) k_firefly_select<<<grid_for(ctx, 4), 256, 0, st>>>(ctx->pb, pass->adaptiveThreshold, (uint32_t)npix, ctx->dList[0], lc, 1, pass->adaptiveExponent);
            else k_firefly_select<<<grid_for(ctx, 4), 256, 0, st>>>(ctx->pb, pass->fireflyThreshold, (uint32_t)npix, ctx->dList[0], lc, 0, 1.0);
            ctx->launches++;
            int cur = 0;
            for (int j = 0; j < nExtra; j++) {  // every listed pixel takes all nExtra samples, each its own AddSample
                Q.sampleBase = (stage == 0 ? kAdaptiveBase : kFireflyBase) + j;
                rc = run_pass(ctx, Q, 1, ctx->dSum, st, ctx->dList[cur], lc + cur);
                if (rc != PTGPU_OK) return rc;
                CK(cudaMemsetAsync(lc + (cur ^ 1), 0, sizeof(uint32_t), st));
                k_firefly_apply<<<grid_for(ctx, 4), 256, 0, st>>>(ctx->dSum, ctx->dList[cur], lc + cur, nullptr, ctx->pb, ctx->dList[cur ^ 1], lc + (cur ^ 1));
                ctx->launches++;
                cur ^= 1;
            }
        }
    } else if (pass->adaptiveSamples > 0) {  // Renderer.cs:340-364
        PassD Q = P;
        Q.stratified = 0; Q.subpixelJitter = 1; Q.sampleStride = 1;
        for (int j = 0; j < pass->adaptiveSamples; j++) {
            Q.sampleBase = kAdaptiveBase + j;
            CK(cudaMemsetAsync(ctx->dSum, 0, npix * 3 * sizeof(float), st));
            rc = run_pass(ctx, Q, 1, ctx->dSum, st);
            if (rc != PTGPU_OK) return rc;
            k_add_sample<<<grid_for(ctx, 4), 256, 0, st>>>(one_sum(ctx->dSum), 1.0, (uint32_t)npix, ctx->pb, nullptr, flags, nflags, 1.f, 0);
            ctx->launches++;
        }
    }
    if (!pass->serialRules && pass->fireflySamples > 0) {  // Renderer.cs:418-468
        PassD Q = P;
        Q.stratified = 0; Q.subpixelJitter = 1; Q.sampleStride = 1;
        uint32_t* lc = ctx->dCounts + 8;  // [8], [9]: list counts
        CK(cudaMemsetAsync(lc, 0, 2 * sizeof(uint32_t), st));
        CK(cudaMemsetAsync(ctx->dSum, 0, npix * 3 * sizeof(float), st));
        k_firefly_select<<<grid_for(ctx, 4), 256, 0, st>>>(ctx->pb, pass->fireflyThreshold, (uint32_t)npix, ctx->dList[0], lc);
        ctx->launches++;
        int cur = 0;
        for (int j = 0; j < pass->fireflySamples; j++) {
            Q.sampleBase = kFireflyBase + j;
            rc = run_pass(ctx, Q, 1, ctx->dSum, st, ctx->dList[cur], lc + cur);
            if (rc != PTGPU_OK) return rc;
            CK(cudaMemsetAsync(lc + (cur ^ 1), 0, sizeof(uint32_t), st));
            k_firefly_decide<<<grid_for(ctx, 4), 256, 0, st>>>(ctx->dSum, ctx->dList[cur], lc + cur, ctx->pb, P.width, P.height, ctx->dReject);
            k_firefly_apply<<<grid_for(ctx, 4), 256, 0, st>>>(ctx->dSum, ctx->dList[cur], lc + cur, ctx->dReject, ctx->pb, ctx->dList[cur ^ 1], lc + (cur ^ 1));
            ctx->launches += 2;
            cur ^= 1;
        }
    }
    CK(cudaEventRecord(ctx->ev1, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->lastPassMs = ms;
    // the Buffer.AddSample kernels saw the flag and left the Buffer alone: the pass can be repeated with a larger queueCapacity
    if (take_overflow(ctx)) return fail(ctx, PTGPU_E_LIMIT, "ray queue overflow (internal sizing error): the pass was not added to the Buffer");
    return PTGPU_OK;
}

int ptgpu_read_buffer(ptgpu_ctx* ctx, int32_t channel, float* out_rgb) {
    if (!ctx || !out_rgb || channel < 0 || channel > 5) return PTGPU_E_ARG;
    if (!ctx->dSum) return fail(ctx, PTGPU_E_STATE, "no pass rendered yet");
    CK(cudaSetDevice(ctx->device));
    size_t npix = (size_t)ctx->bufW * ctx->bufH;
    float* tmp = nullptr;
    CK(cudaMalloc(&tmp, npix * 3 * sizeof(float)));
    k_read_buffer<<<grid_for(ctx, 4), 256, 0, ctx->stream>>>(ctx->pb, channel, ctx->bufW, ctx->bufH, tmp);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(out_rgb, tmp, npix * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    CK(e);
    return PTGPU_OK;
}

// Exact copy of the Welford buffer out of / into the device (checkpoint and resume of an IterativeRender loop).
int ptgpu_export_buffer(ptgpu_ctx* ctx, int32_t* width, int32_t* height, double* M_rgb, double* V_rgb, int32_t* samples) {
    if (!ctx || !width || !height) return PTGPU_E_ARG;
    if (!ctx->dSum) return fail(ctx, PTGPU_E_STATE, "no pass rendered yet");
    CK(cudaSetDevice(ctx->device));
    *width = ctx->bufW; *height = ctx->bufH;
    const size_t npix = (size_t)ctx->bufW * ctx->bufH;
    CK(cudaStreamSynchronize(ctx->stream));
    if (M_rgb) CK(cudaMemcpy(M_rgb, ctx->pb.M, npix * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (V_rgb) CK(cudaMemcpy(V_rgb, ctx->pb.V, npix * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (samples) CK(cudaMemcpy(samples, ctx->pb.samples, npix * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return PTGPU_OK;
}
int ptgpu_import_buffer(ptgpu_ctx* ctx, int32_t width, int32_t height, const double* M_rgb, const double* V_rgb, const int32_t* samples) {
    if (!ctx || width <= 1 || height <= 1 || !M_rgb || !V_rgb || !samples) return PTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_image(ctx, width, height);
    if (rc != PTGPU_OK) return rc;
    const size_t npix = (size_t)width * height;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(ctx->pb.M, M_rgb, npix * 3 * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->pb.V, V_rgb, npix * 3 * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->pb.samples, samples, npix * sizeof(int32_t), cudaMemcpyHostToDevice));
    return PTGPU_OK;
}

int ptgpu_reset_buffer(ptgpu_ctx* ctx) {
    if (!ctx) return PTGPU_E_ARG;
    if (!ctx->dSum) return PTGPU_OK;
    CK(cudaSetDevice(ctx->device));
    size_t npix = (size_t)ctx->bufW * ctx->bufH;
    CK(cudaMemsetAsync(ctx->pb.M, 0, npix * 3 * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->pb.V, 0, npix * 3 * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->pb.samples, 0, npix * sizeof(int32_t), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PTGPU_OK;
}

int ptgpu_intersect_batch(ptgpu_ctx* ctx, int32_t n, const float* o3, const float* d3, int32_t* shape, int32_t* prim, double* t, float* normal3,
                          float* position3, int32_t* inside, int32_t* material) {
    if (!ctx || n < 0 || !o3 || !d3 || !shape || !prim || !t) return PTGPU_E_ARG;
    if (!ctx->haveScene) return fail(ctx, PTGPU_E_STATE, "no scene uploaded");
    if (n == 0) return PTGPU_OK;
    CK(cudaSetDevice(ctx->device));
    float *dO = nullptr, *dD = nullptr, *dN = nullptr, *dP = nullptr;
    int32_t *dS = nullptr, *dPr = nullptr, *dI = nullptr, *dM = nullptr;
    double* dT = nullptr;
    size_t N = (size_t)n;
    int rc = PTGPU_OK;
    auto cleanup = [&]() { cudaFree(dO); cudaFree(dD); cudaFree(dN); cudaFree(dP); cudaFree(dS); cudaFree(dPr); cudaFree(dI); cudaFree(dM); cudaFree(dT); };
#define CKC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { ctx->error = std::string(#call) + ": " + cudaGetErrorString(e__); cleanup(); return PTGPU_E_CUDA; } } while (0)
    CKC(cudaMalloc(&dO, N * 12)); CKC(cudaMalloc(&dD, N * 12));
    if (normal3) CKC(cudaMalloc(&dN, N * 12));
    if (position3) CKC(cudaMalloc(&dP, N * 12));
    if (inside) CKC(cudaMalloc(&dI, N * 4));
    if (material) CKC(cudaMalloc(&dM, N * 4));
    CKC(cudaMalloc(&dS, N * 4)); CKC(cudaMalloc(&dPr, N * 4)); CKC(cudaMalloc(&dT, N * 8));
    CKC(cudaMemcpyAsync(dO, o3, N * 12, cudaMemcpyHostToDevice, ctx->stream));
    CKC(cudaMemcpyAsync(dD, d3, N * 12, cudaMemcpyHostToDevice, ctx->stream));
    CKC(cudaMemsetAsync(ctx->dCounts + 6, 0, sizeof(uint32_t), ctx->stream));
    {
        Lane& L = ctx->lanes[0];
        int rcs = ensure_split(ctx, L, std::max<uint64_t>(N, L.splitCap), ctx->splitStackEnt);
        if (rcs != PTGPU_OK) { cleanup(); return rcs; }
        const BatchOut B{dS, dPr, dT, dN, dP, dI, dM};
        const int tier = ctx->sceneTier;
        rcs = run_split(ctx, L, ctx->stream,
                        [&](const MeshQueue& out) { { if (tier == TIER_NESTED) k_scene_batch<SCENE_START, TIER_NESTED><<<grid_for(ctx, 8), 128, 0, ctx->stream>>>(ctx->scene, L.split, (uint32_t)n, out, out, dO, dD, B); else if (tier == TIER_FULL) k_scene_batch<SCENE_START, TIER_FULL><<<grid_for(ctx, 8), 128, 0, ctx->stream>>>(ctx->scene, L.split, (uint32_t)n, out, out, dO, dD, B); else if (tier == TIER_MESH) k_scene_batch<SCENE_START, TIER_MESH><<<grid_for(ctx, 8), 128, 0, ctx->stream>>>(ctx->scene, L.split, (uint32_t)n, out, out, dO, dD, B); else k_scene_batch<SCENE_START, TIER_ANALYTIC><<<grid_for(ctx, 8), 128, 0, ctx->stream>>>(ctx->scene, L.split, (uint32_t)n, out, out, dO, dD, B); } },
                        [&](const MeshQueue& in, const MeshQueue& out) { { if (tier == TIER_NESTED) k_scene_batch<SCENE_RESUME, TIER_NESTED><<<grid_for(ctx, 8), 128, 0, ctx->stream>>>(ctx->scene, L.split, (uint32_t)n, in, out, dO, dD, B); else if (tier == TIER_FULL) k_scene_batch<SCENE_RESUME, TIER_FULL><<<grid_for(ctx, 8), 128, 0, ctx->stream>>>(ctx->scene, L.split, (uint32_t)n, in, out, dO, dD, B); else if (tier == TIER_MESH) k_scene_batch<SCENE_RESUME, TIER_MESH><<<grid_for(ctx, 8), 128, 0, ctx->stream>>>(ctx->scene, L.split, (uint32_t)n, in, out, dO, dD, B); else k_scene_batch<SCENE_RESUME, TIER_ANALYTIC><<<grid_for(ctx, 8), 128, 0, ctx->stream>>>(ctx->scene, L.split, (uint32_t)n, in, out, dO, dD, B); } },
                        [&](const MeshQueue& in) { { if (tier == TIER_NESTED) k_scene_batch<SCENE_FINISH, TIER_NESTED><<<grid_for(ctx, 4), 128, 0, ctx->stream>>>(ctx->scene, L.split, (uint32_t)n, in, in, dO, dD, B); else if (tier == TIER_FULL) k_scene_batch<SCENE_FINISH, TIER_FULL><<<grid_for(ctx, 4), 128, 0, ctx->stream>>>(ctx->scene, L.split, (uint32_t)n, in, in, dO, dD, B); else if (tier == TIER_MESH) k_scene_batch<SCENE_FINISH, TIER_MESH><<<grid_for(ctx, 4), 128, 0, ctx->stream>>>(ctx->scene, L.split, (uint32_t)n, in, in, dO, dD, B); else k_scene_batch<SCENE_FINISH, TIER_ANALYTIC><<<grid_for(ctx, 4), 128, 0, ctx->stream>>>(ctx->scene, L.split, (uint32_t)n, in, in, dO, dD, B); } });
        if (rcs != PTGPU_OK) { cleanup(); return rcs; }
    }
    CKC(cudaGetLastError());
    CKC(cudaMemcpyAsync(shape, dS, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaMemcpyAsync(prim, dPr, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaMemcpyAsync(t, dT, N * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (normal3) CKC(cudaMemcpyAsync(normal3, dN, N * 12, cudaMemcpyDeviceToHost, ctx->stream));
    if (position3) CKC(cudaMemcpyAsync(position3, dP, N * 12, cudaMemcpyDeviceToHost, ctx->stream));
    if (inside) CKC(cudaMemcpyAsync(inside, dI, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (material) CKC(cudaMemcpyAsync(material, dM, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
#undef CKC
    cleanup();
    return rc;
}

int ptgpu_cast_rays(ptgpu_ctx* ctx, const ptgpu_pass* pass, int32_t n, const int32_t* x, const int32_t* y, const double* fu, const double* fv,
                    const int32_t* sample, float* o3, float* d3) {
    if (!ctx || !pass || n < 0 || !x || !y || !fu || !fv || !sample || !o3 || !d3) return PTGPU_E_ARG;
    if (n == 0) return PTGPU_OK;
    CK(cudaSetDevice(ctx->device));
    PassD P;
    int rc = make_passd(ctx, pass, P);
    if (rc != PTGPU_OK) return rc;
    size_t N = (size_t)n;
    int32_t *dx = nullptr, *dy = nullptr, *ds = nullptr;
    double *du = nullptr, *dv = nullptr;
    float *dO = nullptr, *dD = nullptr;
    auto cleanup = [&]() { cudaFree(dx); cudaFree(dy); cudaFree(ds); cudaFree(du); cudaFree(dv); cudaFree(dO); cudaFree(dD); };
#define CKC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { ctx->error = std::string(#call) + ": " + cudaGetErrorString(e__); cleanup(); return PTGPU_E_CUDA; } } while (0)
    CKC(cudaMalloc(&dx, N * 4)); CKC(cudaMalloc(&dy, N * 4)); CKC(cudaMalloc(&ds, N * 4)); CKC(cudaMalloc(&du, N * 8)); CKC(cudaMalloc(&dv, N * 8));
    CKC(cudaMalloc(&dO, N * 12)); CKC(cudaMalloc(&dD, N * 12));
    CKC(cudaMemcpyAsync(dx, x, N * 4, cudaMemcpyHostToDevice, ctx->stream));
    CKC(cudaMemcpyAsync(dy, y, N * 4, cudaMemcpyHostToDevice, ctx->stream));
    CKC(cudaMemcpyAsync(ds, sample, N * 4, cudaMemcpyHostToDevice, ctx->stream));
    CKC(cudaMemcpyAsync(du, fu, N * 8, cudaMemcpyHostToDevice, ctx->stream));
    CKC(cudaMemcpyAsync(dv, fv, N * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_cast_rays<<<grid_for(ctx, 2), 128, 0, ctx->stream>>>(P, n, dx, dy, du, dv, ds, dO, dD);
    ctx->launches++;
    CKC(cudaGetLastError());
    CKC(cudaMemcpyAsync(o3, dO, N * 12, cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaMemcpyAsync(d3, dD, N * 12, cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
#undef CKC
    cleanup();
    return PTGPU_OK;
}

int ptgpu_keyed_draw(ptgpu_ctx* ctx, uint32_t seed, uint32_t pass, uint32_t pixel, uint32_t sample, uint32_t bits, uint32_t first, uint32_t depth,
                     uint32_t sub, uint32_t drawIndex, double* out) {
    if (!ctx || !out) return PTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    double* d = nullptr;
    CK(cudaMalloc(&d, 8));
    k_keyed_draw<<<1, 1, 0, ctx->stream>>>(seed, pass, pixel, sample, bits, first, depth, sub, drawIndex, d);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(out, d, 8, cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    CK(e);
    return PTGPU_OK;
}

int ptgpu_get_counters(ptgpu_ctx* ctx, ptgpu_counters* out) {
    if (!ctx || !out) return PTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    DeviceCounters dc;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(&dc, ctx->dCounters, sizeof(dc), cudaMemcpyDeviceToHost));
    std::memset(out, 0, sizeof(*out));
    out->cameraSamples = dc.cameraSamples; out->segments = dc.segments; out->shadowRays = dc.shadowRays; out->nanSamples = dc.nanSamples;
    out->kernelLaunches = ctx->launches.load();
#ifdef PT_DEBUG_STEPS
    {
        unsigned long long h[8];
        if (cudaMemcpyFromSymbol(h, g_dbg, sizeof(h)) == cudaSuccess && h[0])
            fprintf(stderr, "[dbg] mesh items %llu: per item %.2f reference-node steps, %.2f bounds-only steps, %.2f micro leaves, %.2f triangle tests\n", h[0],
                    (double)h[1] / h[0], (double)h[2] / h[0], (double)h[3] / h[0], (double)h[4] / h[0]);
    }
#endif
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->lastPassMs = ms;
    out->lastPassMs = ctx->lastPassMs;
    out->traceMs = ctx->traceMs; out->shadeMs = ctx->shadeMs; out->shadowMs = ctx->shadowMs; out->raygenMs = ctx->raygenMs;
    out->meshMs = ctx->kindMs[0]; out->meshLaunches = ctx->kindLaunches[0];
    out->meshItems = ctx->roundItems >= dc.kindItems[1] + dc.kindItems[2] ? ctx->roundItems - dc.kindItems[1] - dc.kindItems[2] : 0;
    out->sdfMs = ctx->kindMs[1]; out->sdfItems = dc.kindItems[1]; out->sdfLaunches = ctx->kindLaunches[1];
    out->volumeMs = ctx->kindMs[2]; out->volumeItems = dc.kindItems[2]; out->volumeLaunches = ctx->kindLaunches[2];
    out->traceLaunches = ctx->traceLaunches;
    take_overflow(ctx);  // an overflow of a ptgpu_accumulate_device pass is reported here
    out->queueOverflows = ctx->overflowPasses;
    out->devices = 1 + ctx->peers.size();
    for (ptgpu_ctx* p : ctx->peers) {  // the peers' shares of the passes
        ptgpu_counters pc;
        if (ptgpu_get_counters(p, &pc) != PTGPU_OK) { ctx->error = p->error; cudaSetDevice(ctx->device); return PTGPU_E_CUDA; }
        out->cameraSamples += pc.cameraSamples; out->segments += pc.segments; out->shadowRays += pc.shadowRays; out->nanSamples += pc.nanSamples;
        out->kernelLaunches += pc.kernelLaunches; out->queueOverflows += pc.queueOverflows;
    }
    if (!ctx->peers.empty()) CK(cudaSetDevice(ctx->device));
    return PTGPU_OK;
}

int ptgpu_reset_counters(ptgpu_ctx* ctx) {
    if (!ctx) return PTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemset(ctx->dCounters, 0, sizeof(DeviceCounters)));
    ctx->launches = 0;
    ctx->overflowPasses = 0;
    for (ptgpu_ctx* p : ctx->peers) { int rc = ptgpu_reset_counters(p); if (rc != PTGPU_OK) return rc; }
    if (!ctx->peers.empty()) CK(cudaSetDevice(ctx->device));
    return PTGPU_OK;
}

int ptgpu_set_profiling(ptgpu_ctx* ctx, int32_t on) {
    if (!ctx) return PTGPU_E_ARG;
    ctx->profiling = on != 0;
    return PTGPU_OK;
}

}  // extern "C"
