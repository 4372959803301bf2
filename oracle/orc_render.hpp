// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity unpinned (see orc_math.hpp header and DESIGN.md).
// CPU restatement of PTSharp's integrator and render driver: Ray.Bounce / Util.Cone / DefaultSampler
// (recursive, as in the reference), Pixel/Buffer (Welford) and Renderer.RenderParallel's tile schedule
// on a shared-FIFO thread pool (WorkStealingScheduler.cs:11-30).
//
// Random numbers.  The reference draws from the unseeded, per-thread Random.Shared everywhere
// (Renderer.cs:297, Sampler.cs:102,207,242, Vector.cs:341, Ray.cs:61), so no two of its runs agree bit for
// bit.  The oracle offers two interchangeable draw sources behind one interface:
//   * RNG_SEQUENTIAL — xoshiro256** (what .NET's Random.Shared is on 64-bit), seeded per 32x32 task, drawn in
//     the reference's own call order;
//   * RNG_KEYED      — Philox4x32-10 addressed by (pixel, sample, path-node, sub-stream, draw index).  This is
//     the stream layout the GPU path uses, which lets tests compare single camera samples, not just
//     converged images.  The integrator below tells the source *where* in the path tree it is
//     (Enter/EnterLight); a sequential source ignores that.
#pragma once
#include <algorithm>
#include <functional>
#include <thread>

#include "orc_scene.hpp"

namespace orc {

// ----------------------------------------------------------------------------------------------- RNG
// Philox4x32-10 (Salmon et al., SC'11); constants from the paper.
static inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int i = 0; i < 10; i++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { RNG_SEQUENTIAL = 0, RNG_KEYED = 1 };

// Sub-stream word (Philox counter word 3):  [31:20] first-hit index+1 | [19:14] depth | [13:6] sub | [5:0] block
static inline uint32_t stream_word(uint32_t first, uint32_t depth, uint32_t sub) {
    return ((first & 0xFFFu) << 20) | ((depth & 0x3Fu) << 14) | ((sub & 0xFFu) << 6);
}

struct Rng {
    int mode = RNG_SEQUENTIAL;
    // sequential
    uint64_t s[4] = {1, 2, 3, 4};
    // keyed
    uint32_t key[2] = {0, 0};
    uint32_t ctr[4] = {0, 0, 0, 0};
    uint32_t draw = 0;
    uint32_t cache[4];
    uint32_t cachedBlock = 0xFFFFFFFFu;

    static uint64_t splitmix(uint64_t& x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    void SeedSequential(uint64_t seed) {
        for (int i = 0; i < 4; i++) s[i] = splitmix(seed);
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t NextU64() {  // xoshiro256**
        uint64_t result = rotl(s[1] * 5, 7) * 9;
        uint64_t t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return result;
    }
    // keyed addressing ----------------------------------------------------------------------------
    void SetSample(uint32_t seed, uint32_t pass, uint32_t pixel, uint32_t sample) {
        key[0] = seed; key[1] = pass; ctr[0] = pixel; ctr[1] = sample;
    }
    // Position the stream on path node (pathBits, first, depth), sub-stream `sub`, draw 0.
    void Enter(uint32_t pathBits, uint32_t first, uint32_t depth, uint32_t sub) {
        ctr[2] = pathBits;
        ctr[3] = stream_word(first, depth, sub);
        draw = 0;
        cachedBlock = 0xFFFFFFFFu;
    }
    double NextDouble() {
        if (mode == RNG_SEQUENTIAL) return (double)(NextU64() >> 11) * (1.0 / 9007199254740992.0);
        uint32_t block = draw >> 1;
        if (block != cachedBlock) {
            uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3] | (block & 0x3Fu)};
            philox4x32_10(c, key, cache);
            cachedBlock = block;
        }
        uint32_t hi = cache[(draw & 1) * 2], lo = cache[(draw & 1) * 2 + 1];
        draw++;
        uint64_t bits = ((uint64_t)hi << 32) | lo;
        return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
    }
    // Random.Shared.Next(n) (Sampler.cs:207) — any unbiased index draw is statistically equivalent.
    int NextInt(int n) {
        int i = (int)(NextDouble() * n);
        return i >= n ? n - 1 : i;
    }
};

// Vector.cs:339-347
static inline Vector RandomUnitVector(Rng& rng) {
    double z = rng.NextDouble() * 2.0 - 1.0;
    double a = rng.NextDouble() * 2.0 * M_PI;
    double r = std::sqrt(1.0 - z * z);
    double x = std::sin(a);
    double y = std::cos(a);
    return Vector(r * x, r * y, z);
}

// Util.cs:17-32
static inline Vector Cone(const Vector& direction, double theta, double u, double v, Rng& rng) {
    if (theta < EPS) return direction;
    theta = theta * (1 - (2 * std::acos(u) / M_PI));
    double m1 = std::sin(theta);
    double m2 = std::cos(theta);
    double a = v * 2 * M_PI;
    Vector q = RandomUnitVector(rng);
    Vector s = direction.Cross(q);
    Vector t = direction.Cross(s);
    Vector d = Vector().Add(s.MulScalar(m1 * std::cos(a))).Add(t.MulScalar(m1 * std::sin(a))).Add(direction.MulScalar(m2)).Normalize();
    return d;
}

// Ray.cs:28-35
static inline Ray WeightedBounce(const Ray& n, double u, double v, Rng& rng) {
    double radius = std::sqrt(u);
    double theta = 2 * M_PI * v;
    Vector s = n.Direction.Cross(RandomUnitVector(rng)).Normalize();
    Vector t = n.Direction.Cross(s);
    Vector d = Vector().Add(s.MulScalar(radius * std::cos(theta))).Add(t.MulScalar(radius * std::sin(theta))).Add(n.Direction.MulScalar(std::sqrt(1 - u)));
    return Ray(n.Origin, d);
}

enum BounceType { BounceTypeAny = 0, BounceTypeDiffuse = 1, BounceTypeSpecular = 2 };  // BounceType.cs
enum LightMode { LightModeRandom = 0, LightModeAll = 1 };                                // LightMode.cs
enum SpecularMode { SpecularModeNaive = 0, SpecularModeFirst = 1, SpecularModeAll = 2 };   // SpecularMode.cs

// Ray.cs:44-85
static inline void Bounce(const Ray& self, const HitInfo& info, double u, double v, int bounceType, Rng& rng,
                          Ray& outRay, bool& reflected, double& pOut) {
    const Ray& n = info.ray;
    const Material& material = info.material;
    double n1 = 1.0, n2 = material.Index;
    if (info.Inside) { double t = n1; n1 = n2; n2 = t; }
    double p = material.Reflectivity >= 0 ? material.Reflectivity : n.Direction.Reflectance(self.Direction, n1, n2);
    bool reflect = false;
    switch (bounceType) {
        case BounceTypeAny: reflect = rng.NextDouble() < p; break;
        case BounceTypeDiffuse: reflect = false; break;
        case BounceTypeSpecular: reflect = true; break;
    }
    if (reflect) {
        Ray reflectedRay(n.Origin, n.Direction.Reflect(self.Direction));  // Ray.cs:21
        outRay = Ray(reflectedRay.Origin, Cone(reflectedRay.Direction, material.Gloss, u, v, rng));
        reflected = true;
        pOut = p;
    } else if (material.Transparent) {
        Ray refracted(n.Origin, n.Direction.Refract(self.Direction, n1, n2));  // Ray.cs:23
        refracted.Origin = refracted.Origin.Add(refracted.Direction.MulScalar(1e-4));
        outRay = Ray(refracted.Origin, Cone(refracted.Direction, material.Gloss, u, v, rng));
        reflected = true;
        pOut = 1 - p;
    } else {
        outRay = WeightedBounce(n, u, v, rng);
        reflected = false;
        pOut = 1 - p;
    }
}

// ------------------------------------------------------------------------------------------- Sampler
struct Counters {
    long long segments = 0;     // scene.Intersect calls issued by sample() (Sampler.cs:62)
    long long shadowRays = 0;   // scene.Intersect calls issued by sampleLight() (Sampler.cs:262)
    long long cameraSamples = 0;
};

struct DefaultSampler {  // Sampler.cs:10-297
    int FirstHitSamples = 1, MaxBounces = 4;
    bool DirectLighting = true, SoftShadows = true;
    int lightMode = LightModeRandom, specularMode = SpecularModeNaive;

    Colour Sample(Scene& scene, const Ray& ray, Rng& rng, Counters& cn) const {  // Sampler.cs:40-43
        return sample(scene, ray, true, FirstHitSamples, 0, rng, cn, 0, 0);
    }

    Colour sampleEnvironment(const Scene& scene, const Ray& ray) const {  // Sampler.cs:177-189
        if (scene.Texture) {
            Vector d = ray.Direction;
            double u = std::atan2(d.Z(), d.X()) + scene.TextureAngle;
            double v = std::atan2(d.Y(), Vector(d.X(), 0, d.Z()).Length());
            u = (u + M_PI) / (2 * M_PI);
            v = (v + M_PI / 2) / M_PI;
            return scene.Texture->Sample(u, v);
        }
        return scene.Color;
    }

    // Sampler.cs:55-145.  (pathBits, first) name this vertex in the path tree for the keyed RNG.
    Colour sample(Scene& scene, const Ray& ray, bool emission, int samples, int depth, Rng& rng, Counters& cn,
                  uint32_t pathBits, uint32_t first) const {
        if (depth > MaxBounces) return Colour(0, 0, 0);
        cn.segments++;
        Hit hit = scene.Intersect(ray);
        if (!hit.Ok()) return sampleEnvironment(scene, ray);
        HitInfo info = hit.Info(ray);
        const Material& material = info.material;
        Colour result(0, 0, 0);
        if (material.Emittance > 0) {
            if (DirectLighting && !emission) return Colour(0, 0, 0);
            result = result.Add(material.Color.MulScalar(material.Emittance * samples));
        }
        int n = (int)std::sqrt((double)samples);
        int ma, mb;
        if (specularMode == SpecularModeAll || (depth == 0 && specularMode == SpecularModeFirst)) {
            ma = BounceTypeDiffuse; mb = BounceTypeSpecular;
        } else {
            ma = BounceTypeAny; mb = BounceTypeAny;
        }
        int k = 0;
        for (int u = 0; u < n; u++) {
            for (int v = 0; v < n; v++) {
                for (int mode = ma; mode <= mb; mode++, k++) {
                    // name of the child vertex this iteration spawns
                    uint32_t cFirst = depth == 0 ? (uint32_t)(k + 1) : first;
                    uint32_t cBits = depth == 0 ? 0u : (pathBits | ((uint32_t)(mode - ma) << ((depth - 1) & 31)));
                    rng.Enter(cBits, cFirst, (uint32_t)depth + 1, 0);
                    double fu = ((double)u + rng.NextDouble()) / (double)n;
                    double fv = ((float)v + rng.NextDouble()) / (double)n;
                    Ray newRay;
                    bool reflected;
                    double p;
                    Bounce(ray, info, fu, fv, mode, rng, newRay, reflected, p);
                    if (mode == BounceTypeAny) p = 1;
                    if (p > 0 && reflected) {
                        Colour indirect = sample(scene, newRay, reflected, 1, depth + 1, rng, cn, cBits, cFirst);
                        Colour tinted = indirect.Mix(material.Color.Mul(indirect), material.Tint);
                        result = result.Add(tinted.MulScalar(p));
                    }
                    if (p > 0 && !reflected) {
                        Colour indirect = sample(scene, newRay, reflected, 1, depth + 1, rng, cn, cBits, cFirst);
                        Colour direct(0, 0, 0);
                        if (DirectLighting) direct = sampleLights(scene, info.ray, rng, cn, cBits, cFirst, depth + 1);
                        result = result.Add(material.Color.Mul(direct.Add(indirect)).MulScalar(p));
                    }
                }
            }
        }
        // Sampler.cs:133-142 (Russian roulette) is unreachable: russianRoulette is always false.
        return result.DivScalar(n * n);
    }

    Colour sampleLights(Scene& scene, const Ray& n, Rng& rng, Counters& cn, uint32_t bits, uint32_t first, int depth) const {  // Sampler.cs:191-210
        int nLights = (int)scene.Lights.size();
        if (nLights == 0) return Colour(0, 0, 0);
        if (lightMode == LightModeAll) {
            Colour result;
            for (int i = 0; i < nLights; i++) {
                rng.Enter(bits, first, (uint32_t)depth, 1 + (uint32_t)(i % 255));
                result = result.Add(sampleLight(scene, n, scene.Lights[i], rng, cn));
            }
            return result.DivScalar(nLights);
        } else {
            rng.Enter(bits, first, (uint32_t)depth, 1);
            int lightIndex = rng.NextInt(nLights);
            return sampleLight(scene, n, scene.Lights[lightIndex], rng, cn).MulScalar((double)nLights);
        }
    }

    Colour sampleLight(Scene& scene, const Ray& n, const IShape* light, Rng& rng, Counters& cn) const {  // Sampler.cs:212-296
        Vector center;
        double radius;
        if (light->Kind() == K_SPHERE) {
            const Sphere* s = static_cast<const Sphere*>(light);
            radius = s->Radius;
            center = s->Center;
        } else if (light->Kind() == K_CYLINDER) {
            const Cylinder* c = static_cast<const Cylinder*>(light);
            radius = c->Radius;
            center = Vector(0, 0, (c->Z0 + c->Z1) / 2);
        } else {
            Box box = light->BoundingBox();
            radius = box.OuterRadius();
            center = box.Center();
        }
        Vector point = center;
        if (SoftShadows) {
            while (true) {
                double x = rng.NextDouble() * 2 - 1;
                double y = rng.NextDouble() * 2 - 1;
                if (x * x + y * y <= 1) {
                    Vector l = center.Sub(n.Origin).Normalize();
                    Vector u = l.Cross(RandomUnitVector(rng)).Normalize();
                    Vector v = l.Cross(u);
                    point = center.Add(u.MulScalar(x * radius)).Add(v.MulScalar(y * radius));
                    break;
                }
            }
        }
        Vector rayDirection = point.Sub(n.Origin).Normalize();
        double diffuse = rayDirection.Dot(n.Direction);
        if (diffuse <= 0) return Colour(0, 0, 0);
        Ray ray(n.Origin, rayDirection);
        cn.shadowRays++;
        Hit hit = scene.Intersect(ray);
        // `hit.Shape != light` is a reference comparison; C# struct shapes are re-boxed per Hit.
        if (!hit.Ok() || hit.Shape != light || !light->IsClass()) return Colour(0, 0, 0);
        double coverage;
        if (light->Kind() == K_CYLINDER) {
            coverage = 1.0;
        } else {
            double hyp = center.Sub(n.Origin).Length();
            double theta = std::asin(radius / hyp);
            double adj = radius / std::tan(theta);
            double d = std::cos(theta) * adj;
            double r = std::sin(theta) * adj;
            coverage = (r * r) / (d * d);
            if (hyp < radius) coverage = 1;
            coverage = net_min(coverage, 1);
        }
        Material material = MaterialAtShape(light, point);
        double m = material.Emittance * diffuse * coverage;
        return material.Color.MulScalar(m);
    }
};

// -------------------------------------------------------------------------------------------- Buffer
struct Pixel {  // Buffer.cs:18-58
    int Samples = 0;
    Colour M, V;
    void AddSample(const Colour& sample) {  // :33-44
        Samples++;
        if (Samples == 1) { M = sample; return; }
        Colour m = M;
        M = M.Add(sample.Sub(M).DivScalar(Samples));
        V = V.Add(sample.Sub(m).Mul(sample.Sub(M)));
    }
    Colour Variance() const {  // :48-55
        if (Samples < 2) return Colour(0, 0, 0);
        return V.DivScalar((double)(Samples - 1));
    }
};

struct Buffer {  // Buffer.cs:60-133
    int W = 0, H = 0;
    std::vector<Pixel> Pixels;
    Buffer(int w, int h) : W(w), H(h), Pixels((size_t)w * h) {}
    void AddSample(int x, int y, const Colour& s) { Pixels[(size_t)y * W + x].AddSample(s); }
};

// ------------------------------------------------------------------------------------------ Renderer
struct RenderOptions {
    int SamplesPerPixel = 2;          // Renderer.cs:42
    bool StratifiedSampling = false;  // Renderer.cs:44
    int threads = 1;                  // Environment.ProcessorCount in the reference (Renderer.cs:260)
    int rngMode = RNG_SEQUENTIAL;
    uint32_t seed = 0x50545348u;      // "PTSH"
    uint32_t pass = 0;
    int sampleBase = 0;               // global index of this pass's first sample (keyed RNG)
    int sampleStride = 1;             // global index stride (a rank of a multi-process job draws every stride-th sample)
    int AdaptiveSamples = 0;          // Renderer.cs:23
    int FireflySamples = 0;           // Renderer.cs:27
    double FireflyThreshold = 1;      // Renderer.cs:47
    bool SerialRules = false;         // the extra samples follow the serial Render() (Renderer.cs:150-191) instead of RenderParallel
    double AdaptiveThreshold = 1, AdaptiveExponent = 1;  // Renderer.cs:44-45
    // optional window (bounded CPU-baseline samples): pixels outside it are skipped
    int x0 = 0, y0 = 0, x1 = -1, y1 = -1;
    // optional task sample (bounded CPU-baseline samples of a large frame): of the non-empty 32x32 tasks of the whole frame, in queue
    // order, only every taskStride-th one starting at taskOffset is rendered - a strided sample that covers the frame evenly
    int taskStride = 1, taskOffset = 0;
};

// One camera sample of the non-stratified branch (Renderer.cs:294-304), incl. the fu/fv quirk.
static inline Colour RenderOneSample(Scene& scene, const Camera& camera, const DefaultSampler& sampler, int w, int h,
                                     int x, int y, Rng& rng, Counters& cn) {
    double xOffset = rng.NextDouble();
    double yOffset = rng.NextDouble();
    double fu = (x + xOffset) / w;
    double fv = (y + yOffset) / h;
    Ray ray = camera.CastRay(x, y, w, h, fu, fv, rng);
    cn.cameraSamples++;
    return sampler.Sample(scene, ray, rng, cn);
}

// One pass of Renderer.RenderParallel (Renderer.cs:199-338): adds one Welford sample per pixel (the mean of spp
// radiance samples) in the default branch, or sppRoot^2 individual samples in the stratified branch.
static inline Counters RenderPass(Scene& scene, const Camera& camera, const DefaultSampler& sampler, Buffer& buf,
                                  const RenderOptions& opt) {
    int w = buf.W, h = buf.H;
    int spp = opt.SamplesPerPixel;
    int sppRoot = (int)std::sqrt((double)spp);
    scene.Compile();
    scene.rays = 0;
    int wx0 = opt.x0, wy0 = opt.y0, wx1 = opt.x1 < 0 ? w : opt.x1, wy1 = opt.y1 < 0 ? h : opt.y1;

    struct Task { int x0, y0, x1, y1; };
    std::vector<Task> tasks;
    // Renderer.cs:257-281: 256-px tiles cut into 32x32 sub-tiles, enqueued in this order (empty ones included).
    const int tile_size = 256, sub_tile_size = 32;
    int num_tiles_x = (w + tile_size - 1) / tile_size;
    int num_tiles_y = (h + tile_size - 1) / tile_size;
    for (int tile_index = 0; tile_index < num_tiles_x * num_tiles_y; tile_index++) {
        int tile_x = tile_index % num_tiles_x, tile_y = tile_index / num_tiles_x;
        int x_start = tile_x * tile_size, y_start = tile_y * tile_size;
        int x_end = std::min(x_start + tile_size, w), y_end = std::min(y_start + tile_size, h);
        for (int sy = 0; sy < tile_size; sy += sub_tile_size)
            for (int sx = 0; sx < tile_size; sx += sub_tile_size) {
                Task t;
                t.x0 = x_start + sx; t.y0 = y_start + sy;
                t.x1 = std::min(t.x0 + sub_tile_size, x_end);
                t.y1 = std::min(t.y0 + sub_tile_size, y_end);
                tasks.push_back(t);
            }
    }
    if (opt.taskStride > 1) {  // bounded sample: keep every taskStride-th task that has pixels
        std::vector<Task> kept;
        size_t k = 0;
        for (const Task& t : tasks) {
            if (t.x1 <= t.x0 || t.y1 <= t.y0) continue;
            if ((int)(k++ % (size_t)opt.taskStride) == opt.taskOffset) kept.push_back(t);
        }
        tasks.swap(kept);
    }
    std::atomic<size_t> next{0};
    int nthreads = std::max(1, opt.threads);
    std::vector<Counters> perThread((size_t)nthreads);
    auto worker = [&](int tid) {
        Counters cn;
        Rng rng;
        rng.mode = opt.rngMode;
        for (;;) {
            size_t ti = next.fetch_add(1);
            if (ti >= tasks.size()) break;
            const Task& t = tasks[ti];
            if (opt.rngMode == RNG_SEQUENTIAL) rng.SeedSequential(((uint64_t)opt.seed << 32) ^ ((uint64_t)opt.pass << 20) ^ ti);
            for (int y = t.y0; y < t.y1; y++) {
                for (int x = t.x0; x < t.x1; x++) {
                    if (x < wx0 || x >= wx1 || y < wy0 || y >= wy1) continue;
                    uint32_t pixel = (uint32_t)(y * w + x);
                    if (opt.StratifiedSampling) {  // Renderer.cs:231-254
                        int si = 0;
                        for (int u = 0; u < sppRoot; u++)
                            for (int v = 0; v < sppRoot; v++, si++) {
                                double fu = ((double)u + 0.5) / (double)sppRoot;
                                double fv = ((double)v + 0.5) / (double)sppRoot;
                                rng.SetSample(opt.seed, opt.pass, pixel, (uint32_t)(opt.sampleBase + si * opt.sampleStride));
                                rng.Enter(0, 0, 0, 0);
                                Ray ray = camera.CastRay(x, y, w, h, fu, fv, rng);
                                cn.cameraSamples++;
                                buf.AddSample(x, y, sampler.Sample(scene, ray, rng, cn));
                            }
                    } else {  // Renderer.cs:287-311
                        Colour c(0, 0, 0);
                        for (int p = 0; p < spp; p++) {
                            rng.SetSample(opt.seed, opt.pass, pixel, (uint32_t)(opt.sampleBase + p * opt.sampleStride));
                            rng.Enter(0, 0, 0, 0);
                            c = c.Add(RenderOneSample(scene, camera, sampler, w, h, x, y, rng, cn));
                        }
                        c = c.DivScalar(spp);
                        buf.AddSample(x, y, c);
                    }
                }
            }
        }
        perThread[(size_t)tid] = cn;
    };
    std::vector<std::thread> pool;
    for (int i = 1; i < nthreads; i++) pool.emplace_back(worker, i);
    worker(0);
    for (auto& th : pool) th.join();  // Renderer.cs:336 (Dispose joins the pool)

    // Parallel.For over pixel indices (Renderer.cs:349, 423): contiguous chunks handed to the same thread count.
    auto parallelFor = [&](size_t n, const std::function<void(size_t, size_t, int)>& body) {
        std::vector<std::thread> ths;
        size_t chunk = (n + (size_t)nthreads - 1) / (size_t)nthreads;
        for (int t = 1; t < nthreads; t++) ths.emplace_back([&, t] { size_t a = std::min(n, chunk * t), b = std::min(n, chunk * (t + 1)); if (a < b) body(a, b, t); });
        if (n > 0) body(0, std::min(n, chunk), 0);
        for (auto& th : ths) th.join();
    };
    const uint32_t kAdaptiveBase = 1u << 20, kFireflyBase = 1u << 21;  // global sample index ranges of the extra samples
    // One extra camera sample with uniform sub-pixel jitter: fu, fv = Random.Shared.NextDouble() (Renderer.cs:351-354, 432).
    auto extraSample = [&](int x, int y, uint32_t sampleIndex, Rng& rng, Counters& cn) {
        rng.SetSample(opt.seed, opt.pass, (uint32_t)(y * w + x), sampleIndex);
        rng.Enter(0, 0, 0, 0);
        double fu = rng.NextDouble(), fv = rng.NextDouble();
        Ray ray = camera.CastRay(x, y, w, h, fu, fv, rng);
        cn.cameraSamples++;
        return sampler.Sample(scene, ray, rng, cn);
    };
    if (opt.SerialRules) {
        // Renderer.cs:150-191 (the serial Render(), what IterativeRender runs when NumCPU == 1).  Per pixel, after its main sample:
        //   samples = AdaptiveSamples * (int)Math.Pow(Math.Clamp(StandardDeviation.MaxComponent / AdaptiveThreshold, 0, 1), AdaptiveExponent)
        //   more samples with fu, fv = rand.NextDouble() (:153-165); then, if the deviation exceeds FireflyThreshold, FireflySamples
        //   more with fu = (x + rand) * invWidth, invWidth = 1.0f / w (:97, :172-186) and no IsFirefly test.  Pixels are independent
        //   of each other here, so the two stages run frame-wide one after the other.
        auto deviation = [&](size_t i) { return buf.Pixels[i].Variance().Pow((double)0.5f).MaxComponent(); };
        for (int stage = 0; stage < 2; stage++) {
            const int nExtra = stage == 0 ? opt.AdaptiveSamples : opt.FireflySamples;
            if (nExtra <= 0) continue;
            std::vector<uint32_t> list;
            for (size_t i = 0; i < buf.Pixels.size(); i++) {
                int y = (int)(i / (size_t)w), x = (int)(i % (size_t)w);
                if (x < wx0 || x >= wx1 || y < wy0 || y >= wy1) continue;
                bool pick;
                if (stage == 0) {
                    double v = deviation(i) / opt.AdaptiveThreshold;
                    v = v < 0 ? 0.0 : (v > 1 ? 1.0 : v);
                    v = std::pow(v, opt.AdaptiveExponent);
                    pick = (int)v >= 1;
                } else pick = deviation(i) > opt.FireflyThreshold;
                if (pick) list.push_back((uint32_t)i);
            }
            std::vector<Colour> fresh(list.size());
            for (int j = 0; j < nExtra; j++) {
                parallelFor(list.size(), [&](size_t a, size_t b, int tid) {
                    Rng rng; rng.mode = opt.rngMode;
                    if (opt.rngMode == RNG_SEQUENTIAL) rng.SeedSequential(((uint64_t)opt.seed << 32) ^ ((uint64_t)opt.pass << 20) ^ (stage ? 0xF1000 : 0xA1000) ^ ((uint64_t)j << 8) ^ (uint64_t)tid);
                    Counters cn;
                    for (size_t i = a; i < b; i++) {
                        int y = (int)(list[i] / (uint32_t)w), x = (int)(list[i] % (uint32_t)w);
                        if (stage == 0) fresh[i] = extraSample(x, y, kAdaptiveBase + (uint32_t)j, rng, cn);
                        else {
                            rng.SetSample(opt.seed, opt.pass, list[i], kFireflyBase + (uint32_t)j);
                            rng.Enter(0, 0, 0, 0);
                            const double xo = rng.NextDouble(), yo = rng.NextDouble();
                            const double fu = ((double)x + xo) * (double)(1.0f / (float)w), fv = ((double)y + yo) * (double)(1.0f / (float)h);
                            Ray ray = camera.CastRay(x, y, w, h, fu, fv, rng);
                            cn.cameraSamples++;
                            fresh[i] = sampler.Sample(scene, ray, rng, cn);
                        }
                    }
                    perThread[(size_t)tid].segments += cn.segments; perThread[(size_t)tid].shadowRays += cn.shadowRays; perThread[(size_t)tid].cameraSamples += cn.cameraSamples;
                });
                for (size_t i = 0; i < list.size(); i++) buf.Pixels[list[i]].AddSample(fresh[i]);
            }
        }
    } else if (opt.AdaptiveSamples > 0) {
        // Renderer.cs:340-364: AdaptiveSamples more samples for EVERY pixel, each its own Buffer.AddSample.  (The second loop,
        // :376-388, re-renders as many samples only to fill a dictionary nothing reads: no effect on the Buffer, omitted.)
        parallelFor((size_t)w * h, [&](size_t a, size_t b, int tid) {
            Rng rng; rng.mode = opt.rngMode;
            if (opt.rngMode == RNG_SEQUENTIAL) rng.SeedSequential(((uint64_t)opt.seed << 32) ^ ((uint64_t)opt.pass << 20) ^ 0xA0000 ^ (uint64_t)tid);
            Counters cn;
            for (size_t i = a; i < b; i++) {
                int y = (int)(i / (size_t)w), x = (int)(i % (size_t)w);
                if (x < wx0 || x >= wx1 || y < wy0 || y >= wy1) continue;
                for (int j = 0; j < opt.AdaptiveSamples; j++) buf.AddSample(x, y, extraSample(x, y, kAdaptiveBase + (uint32_t)j, rng, cn));
            }
            perThread[(size_t)tid].segments += cn.segments; perThread[(size_t)tid].shadowRays += cn.shadowRays; perThread[(size_t)tid].cameraSamples += cn.cameraSamples;
        });
    }
    if (!opt.SerialRules && opt.FireflySamples > 0) {
        // Renderer.cs:418-468.  Every pixel whose StandardDeviation().MaxComponent() > FireflyThreshold takes up to
        // FireflySamples more samples and stops at the first one IsFirefly() rejects.  The reference lets every pixel's loop
        // read its neighbours' running means while other threads update them; here the iterations are synchronous (all
        // pixels draw sample j, decide against the buffer as it stood before sample j, then apply) so that a run is
        // reproducible.  `skippedPixels` is local to the pass, so its else-branch (:448-463) is unreachable.
        std::vector<uint32_t> list;
        for (size_t i = 0; i < buf.Pixels.size(); i++) {
            Colour sd = buf.Pixels[i].Variance().Pow((double)0.5f);  // Pixel.StandardDeviation (Buffer.cs:57)
            if (sd.MaxComponent() > opt.FireflyThreshold) list.push_back((uint32_t)i);
        }
        std::vector<Colour> fresh;
        std::vector<uint8_t> reject;
        for (int j = 0; j < opt.FireflySamples && !list.empty(); j++) {
            fresh.assign(list.size(), Colour());
            parallelFor(list.size(), [&](size_t a, size_t b, int tid) {
                Rng rng; rng.mode = opt.rngMode;
                if (opt.rngMode == RNG_SEQUENTIAL) rng.SeedSequential(((uint64_t)opt.seed << 32) ^ ((uint64_t)opt.pass << 20) ^ 0xF0000 ^ ((uint64_t)j << 8) ^ (uint64_t)tid);
                Counters cn;
                for (size_t i = a; i < b; i++) {
                    int y = (int)(list[i] / (uint32_t)w), x = (int)(list[i] % (uint32_t)w);
                    fresh[i] = extraSample(x, y, kFireflyBase + (uint32_t)j, rng, cn);
                }
                perThread[(size_t)tid].segments += cn.segments; perThread[(size_t)tid].shadowRays += cn.shadowRays; perThread[(size_t)tid].cameraSamples += cn.cameraSamples;
            });
            reject.assign(list.size(), 0);
            for (size_t i = 0; i < list.size(); i++) {  // IsFirefly + CalculateLocalDeviation (Renderer.cs:474-537)
                const Colour& smp = fresh[i];
                int y = (int)(list[i] / (uint32_t)w), x = (int)(list[i] % (uint32_t)w);
                double brightness = smp.r * 0.2126 + smp.g * 0.7152 + smp.b * 0.0722;
                if (brightness > 0.9) {
                    int sx = std::max(0, x - 1), sy = std::max(0, y - 1), ex = std::min(w - 1, x + 1), ey = std::min(h - 1, y + 1);
                    double tr = 0, tg = 0, tb = 0; int count = 0;
                    for (int jj = sy; jj <= ey; jj++)
                        for (int ii = sx; ii <= ex; ii++) { const Colour& c = buf.Pixels[(size_t)jj * w + ii].M; tr += c.r; tg += c.g; tb += c.b; count++; }
                    double dr = std::fabs(smp.r - tr / count), dg = std::fabs(smp.g - tg / count), db = std::fabs(smp.b - tb / count);
                    reject[i] = std::sqrt(dr * dr + dg * dg + db * db) > 0.2;
                }
            }
            std::vector<uint32_t> next;
            for (size_t i = 0; i < list.size(); i++)
                if (!reject[i]) { buf.Pixels[list[i]].AddSample(fresh[i]); next.push_back(list[i]); }
            list.swap(next);
        }
    }
    Counters total;
    for (const Counters& c : perThread) {
        total.segments += c.segments;
        total.shadowRays += c.shadowRays;
        total.cameraSamples += c.cameraSamples;
    }
    return total;
}

}  // namespace orc
