// Host-side derivation of what the mesh walk reads (see mesh_step in pt_device.cuh), from the flat scene of include/ptgpu.h.
// Plain C++ (no CUDA types) so it can be compiled and timed without a device; ptgpu_upload_scene calls it with a pinned
// staging allocator.  Runs on all host threads: the leaves of the reference trees (Tree.cs:201-265) are independent.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ptgpu.h"

#if defined(__CUDACC__)
#define PT_LAYOUT_HD __host__ __device__ __forceinline__
#else
#define PT_LAYOUT_HD inline
#endif

// ------------------------------------------------------------------------------------------------ record layout
// Node records (4 x uint4 = 16 words; words 0-1 = split, word 2 = a, word 3 = b, words 4-15 = both children's padded bounds):
//   reference interior  a = left << 2 | axis (1..3), b = right
//   bounds-only node    a = left << 2, b = kNodeVirtual | right [| kNodeRefLeaf]
// A child reference (30 bits) is either a node index (< 2^29) or a micro leaf named in place, saving the round trip
// to a record that would only hold (first, count):  bit 29 = 1 | bit 28 = root of a reference leaf | bits 27:26 = count - 1
// | bits 25:0 = first triangle in leafGeom.
static constexpr uint32_t kNodeVirtual = 0x80000000u, kNodeRefLeaf = 0x40000000u, kNodeIndexMask = 0x3FFFFFFFu;
static constexpr uint32_t kRefLeaf = 1u << 29, kRefLeafRoot = 1u << 28, kRefFirstMask = (1u << 26) - 1u;
PT_LAYOUT_HD uint32_t leaf_ref(uint32_t first, uint32_t count, bool root) { return kRefLeaf | (root ? kRefLeafRoot : 0u) | ((count - 1u) << 26) | first; }
static constexpr int kVirtualDepthMax = 12;  // levels of bounds-only nodes below a reference leaf (4 * 2^12 triangles)

// ------------------------------------------------------------------------------------------------ derivation
struct MeshDerived {
    uint32_t* mn = nullptr;          // node records, 16 words each: [0, numNodes) mirror the reference nodes, bounds-only nodes follow
    uint64_t mnRecords = 0;
    ptgpu_tri_geom* lg = nullptr;    // leaf triangles in sorted order; pad0 = triangle id, pad1 = position in the reference leaf
    uint64_t lgCount = 0;
    std::vector<ptgpu_tree> trees;   // trees[] with the mesh roots replaced by their child references
    int virtualDepth = 0;
    double ms = 0;                   // host time of the derivation
};

namespace mesh_derive_detail {

inline int host_threads() {
    if (const char* ev = std::getenv("PTGPU_HOST_THREADS")) { const int v = std::atoi(ev); if (v > 0) return v; }
    const unsigned hc = std::thread::hardware_concurrency();
    return (int)std::min(32u, std::max(1u, hc));
}

// f(begin, end, worker) over [0, n) in chunks handed out by an atomic counter
template <class F>
inline void parallel_for(uint64_t n, uint64_t chunk, int threads, F f) {
    if (n == 0) return;
    const uint64_t chunks = (n + chunk - 1) / chunk;
    const int nt = (int)std::min<uint64_t>((uint64_t)threads, chunks);
    if (nt <= 1) { f((uint64_t)0, n, 0); return; }
    std::atomic<uint64_t> next{0};
    auto body = [&](int w) {
        for (;;) {
            const uint64_t c = next.fetch_add(1);
            if (c >= chunks) return;
            f(c * chunk, std::min(n, (c + 1) * chunk), w);
        }
    };
    std::vector<std::thread> pool;
    for (int w = 1; w < nt; w++) pool.emplace_back(body, w);
    body(0);
    for (auto& t : pool) t.join();
}

// records of the bounds-only subtree over n > 4 triangles, root included (halving split, micro leaves of <= 4)
inline uint32_t virt_records(uint32_t n) {
    const uint32_t a = (n + 1) / 2, b = n - a;
    return 1u + (a > 4 ? virt_records(a) : 0u) + (b > 4 ? virt_records(b) : 0u);
}

inline uint32_t float_bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float bits_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t spread10(uint32_t x) {
    x &= 1023u; x = (x | (x << 16)) & 0x030000FFu; x = (x | (x << 8)) & 0x0300F00Fu; x = (x | (x << 4)) & 0x030C30C3u;
    return (x | (x << 2)) & 0x09249249u;
}

// Padded bounds of n triangles: 1e-4 of the box size plus 1e-5 of the coordinate magnitude (the FP32 triangle test errs by
// ~1e-7 of |origin - vertex|; the origin-dependent part is added per ray, see box_line_hit in pt_device.cuh).
template <class GetTri>
inline void padded_bounds(uint32_t n, GetTri tri, float* lo, float* hi) {
    const float BIG = 3.0e38f;
    for (int c = 0; c < 3; c++) { lo[c] = BIG; hi[c] = -BIG; }
    for (uint32_t k = 0; k < n; k++) {
        const ptgpu_tri_geom& g = tri(k);
        for (int c = 0; c < 3; c++) {
            const float p0 = g.v1[c], p1 = g.v1[c] + g.e1[c], p2 = g.v1[c] + g.e2[c];
            lo[c] = std::min(lo[c], std::min(p0, std::min(p1, p2)));
            hi[c] = std::max(hi[c], std::max(p0, std::max(p1, p2)));
        }
    }
    const float ext = std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2]));
    for (int c = 0; c < 3; c++) {
        const float padv = 1e-4f * ext + 1e-5f * std::max(std::fabs(lo[c]), std::fabs(hi[c])) + 1e-7f;
        lo[c] -= padv; hi[c] += padv;
    }
}

}  // namespace mesh_derive_detail

// `alloc(bytes)` returns host staging memory that stays valid until the caller has copied it (pinned in ptgpu_upload_scene).
// False + `err` when a compile-time limit of the walk is exceeded.
inline bool derive_mesh(const ptgpu_flat_scene* s, MeshDerived& out, std::string& err, const std::function<void*(uint64_t)>& alloc) {
    using namespace mesh_derive_detail;
    const int threads = host_threads();
    const bool trace = std::getenv("PTGPU_DERIVE_TRACE") != nullptr;
    auto tprev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (!trace) return; auto t = std::chrono::steady_clock::now(); std::fprintf(stderr, "derive_mesh: %-28s %7.1f ms\n", what, std::chrono::duration<double, std::milli>(t - tprev).count()); tprev = t; };
    const uint64_t nn = s->numNodes;
    const float BIG = 3.0e38f;
    // nodes of a tree are contiguous from its root up to the next tree's root
    std::vector<uint8_t> isMeshNode(nn, 0);
    for (uint32_t m = 0; m < s->numMeshes; m++) {
        const ptgpu_tree& t = s->trees[s->meshes[m].tree];
        uint64_t end = nn;
        for (uint32_t k = 0; k < s->numTrees; k++) if (s->trees[k].root > t.root && s->trees[k].root < end) end = s->trees[k].root;
        std::memset(isMeshNode.data() + t.root, 1, (size_t)(end - t.root));
    }
    // the reference leaves of the mesh trees, in node order; offsets of their triangles and bounds-only records
    std::vector<uint32_t> leaves;
    for (uint64_t i = 0; i < nn; i++) if (isMeshNode[i] && (s->nodes[i].a & 3u) == 0) leaves.push_back((uint32_t)i);
    const uint64_t nl = leaves.size();
    lap("mesh nodes + leaf list");
    std::vector<uint64_t> lgOff(nl + 1, 0), vOff(nl + 1, 0);
    parallel_for(nl, 4096, threads, [&](uint64_t b, uint64_t e, int) {
        for (uint64_t k = b; k < e; k++) {
            const uint32_t count = s->nodes[leaves[k]].b;
            lgOff[k + 1] = count ? count : 1;                       // an empty mesh gets one degenerate triangle
            vOff[k + 1] = count > 4 ? virt_records(count) - 1u : 0u;  // the root of the subtree is the leaf's own record
        }
    });
    for (uint64_t k = 0; k < nl; k++) { lgOff[k + 1] += lgOff[k]; vOff[k + 1] += vOff[k]; }
    lap("offsets");
    out.lgCount = lgOff[nl];
    out.mnRecords = nn + vOff[nl];
    if (out.lgCount > (uint64_t)kRefFirstMask) { err = "mesh too large for the 26-bit leaf triangle index"; return false; }
    if (out.mnRecords >= (uint64_t)kRefLeaf) { err = "mesh too large for the 29-bit node index"; return false; }
    out.mn = static_cast<uint32_t*>(alloc(std::max<uint64_t>(out.mnRecords, 1) * 64));
    out.lg = static_cast<ptgpu_tri_geom*>(alloc(std::max<uint64_t>(out.lgCount, 1) * sizeof(ptgpu_tri_geom)));
    if (!out.mn || !out.lg) { err = "out of host staging memory"; return false; }
    uint32_t* mn = out.mn;
    ptgpu_tri_geom* lg = out.lg;
    parallel_for(out.mnRecords, 1 << 16, threads, [&](uint64_t b, uint64_t e, int) { std::memset(mn + b * 16, 0, (size_t)(e - b) * 64); });

    lap("staging alloc + clear");
    std::vector<float> nb(nn * 8);             // padded triangle bounds below every reference node
    std::vector<uint32_t> ref(nn);             // how a parent (or the tree) refers to reference node i
    parallel_for(nn, 1 << 16, threads, [&](uint64_t b, uint64_t e, int) { for (uint64_t i = b; i < e; i++) ref[i] = (uint32_t)i; });
    std::vector<int> depthOf((size_t)threads, 0);

    lap("nb/ref alloc");
    // pass 1: reference leaves -> padded bounds, triangles sorted along a Morton curve (each carrying its global index and
    // its position in the reference leaf, the tie-break of Tree.cs:122) and, for more than 4 triangles, a bounds-only binary
    // hierarchy rooted at the leaf's own record and ending in micro leaves of <= 4 triangles
    parallel_for(nl, 256, threads, [&](uint64_t kb, uint64_t ke, int worker) {
        std::vector<std::pair<uint32_t, uint32_t>> order;  // (morton, position in leaf)
        std::vector<float> cen;
        struct Build { uint64_t idx; uint32_t t0, t1; int depth; };
        std::vector<Build> todo;
        int vdepth = 0;
        for (uint64_t k = kb; k < ke; k++) {
            const uint64_t i = leaves[k];
            const ptgpu_node& n = s->nodes[i];
            const uint32_t first = n.a >> 2, count = n.b;
            {
                float lo[3], hi[3];
                padded_bounds(count, [&](uint32_t q) -> const ptgpu_tri_geom& { return s->triGeom[s->leafItems[first + q]]; }, lo, hi);
                float* o = &nb[i * 8];
                o[0] = lo[0]; o[1] = lo[1]; o[2] = lo[2]; o[3] = 0; o[4] = hi[0]; o[5] = hi[1]; o[6] = hi[2]; o[7] = 0;
            }
            const uint32_t t0 = (uint32_t)lgOff[k];
            if (count == 0) {  // cannot come out of Node.Split (an empty side is rejected) except for an empty mesh
                std::memset(&lg[t0], 0, sizeof(ptgpu_tri_geom));
                ref[i] = leaf_ref(t0, 1, true);
                continue;
            }
            float lo[3] = {BIG, BIG, BIG}, hi[3] = {-BIG, -BIG, -BIG};
            cen.resize((size_t)count * 3);
            for (uint32_t q = 0; q < count; q++) {
                const ptgpu_tri_geom& g = s->triGeom[s->leafItems[first + q]];
                for (int c = 0; c < 3; c++) {
                    const float v = g.v1[c] + (g.e1[c] + g.e2[c]) * (1.0f / 3.0f);
                    cen[(size_t)q * 3 + c] = v; lo[c] = std::min(lo[c], v); hi[c] = std::max(hi[c], v);
                }
            }
            order.clear();
            for (uint32_t q = 0; q < count; q++) {
                uint32_t qz[3];
                for (int c = 0; c < 3; c++) {
                    const float e = hi[c] - lo[c];
                    const float f = e > 0 ? (cen[(size_t)q * 3 + c] - lo[c]) / e : 0.f;
                    qz[c] = (uint32_t)std::min(1023.f, std::max(0.f, f * 1023.f));
                }
                order.push_back({spread10(qz[0]) | (spread10(qz[1]) << 1) | (spread10(qz[2]) << 2), q});
            }
            std::sort(order.begin(), order.end());
            for (uint32_t q = 0; q < count; q++) {
                const uint32_t pos = order[q].second, tri = s->leafItems[first + pos];
                ptgpu_tri_geom g = s->triGeom[tri];
                g.pad0 = bits_float(tri); g.pad1 = bits_float(pos); g.pad2 = 0.f;
                lg[t0 + q] = g;
            }
            if (count <= 4) { ref[i] = leaf_ref(t0, count, true); continue; }
            mn[i * 16 + 3] = kNodeRefLeaf;
            uint64_t nextRec = nn + vOff[k];  // this leaf's bounds-only records are [nn + vOff[k], nn + vOff[k + 1])
            todo.clear();
            todo.push_back({i, t0, t0 + count, 0});
            while (!todo.empty()) {
                const Build bld = todo.back(); todo.pop_back();
                vdepth = std::max(vdepth, bld.depth + 1);
                const uint32_t cnt = bld.t1 - bld.t0, mid = bld.t0 + (cnt + 1) / 2;
                uint32_t refs[2];
                const uint32_t lo2[2] = {bld.t0, mid}, hi2[2] = {mid, bld.t1};
                for (int side = 0; side < 2; side++) {
                    const uint32_t cn = hi2[side] - lo2[side];
                    if (cn <= 4) refs[side] = leaf_ref(lo2[side], cn, false);
                    else {
                        refs[side] = (uint32_t)nextRec++;
                        todo.push_back({refs[side], lo2[side], hi2[side], bld.depth + 1});
                    }
                }
                uint32_t* o = &mn[bld.idx * 16];
                o[2] = refs[0] << 2; o[3] = (o[3] & kNodeRefLeaf) | kNodeVirtual | refs[1];
                float llo[3], lhi[3], rlo[3], rhi[3];
                padded_bounds(mid - bld.t0, [&](uint32_t q) -> const ptgpu_tri_geom& { return lg[bld.t0 + q]; }, llo, lhi);
                padded_bounds(bld.t1 - mid, [&](uint32_t q) -> const ptgpu_tri_geom& { return lg[mid + q]; }, rlo, rhi);
                const float pk[12] = {llo[0], llo[1], llo[2], lhi[0], lhi[1], lhi[2], rlo[0], rlo[1], rlo[2], rhi[0], rhi[1], rhi[2]};
                for (int q = 0; q < 12; q++) o[4 + q] = float_bits(pk[q]);
            }
        }
        depthOf[(size_t)worker] = std::max(depthOf[(size_t)worker], vdepth);
    });
    lap("pass 1 (leaves)");
    for (int v : depthOf) out.virtualDepth = std::max(out.virtualDepth, v);
    if (out.virtualDepth > kVirtualDepthMax) { err = "a kd leaf holds more triangles than the bounds-only hierarchy supports"; return false; }

    // interior bounds: nodes are stored parent-before-child, so one reverse sweep folds children into parents.
    // Scene-tree nodes get an unbounded box (never culled).
    for (uint64_t ii = nn; ii-- > 0;) {
        float* o = &nb[ii * 8];
        const ptgpu_node& n = s->nodes[ii];
        if (!isMeshNode[ii]) { o[0] = o[1] = o[2] = -BIG; o[4] = o[5] = o[6] = BIG; o[3] = o[7] = 0; }
        else if ((n.a & 3u) != 0) {
            const float* l = &nb[(uint64_t)(n.a >> 2) * 8];
            const float* r = &nb[(uint64_t)n.b * 8];
            for (int c = 0; c < 3; c++) { o[c] = std::min(l[c], r[c]); o[4 + c] = std::max(l[4 + c], r[4 + c]); }
            o[3] = o[7] = 0;
        }
    }
    lap("interior bounds sweep");
    // pass 2: reference interior nodes with both children's padded bounds
    parallel_for(nn, 1 << 15, threads, [&](uint64_t b, uint64_t e, int) {
        for (uint64_t i = b; i < e; i++) {
            if (!isMeshNode[i]) continue;
            const ptgpu_node& n = s->nodes[i];
            if ((n.a & 3u) == 0) continue;
            uint32_t* o = &mn[i * 16];
            std::memcpy(o, &n.split, 8);
            o[2] = (ref[n.a >> 2] << 2) | (n.a & 3u); o[3] = ref[n.b];
            const float* l = &nb[(uint64_t)(n.a >> 2) * 8];
            const float* r = &nb[(uint64_t)n.b * 8];
            const float pk[12] = {l[0], l[1], l[2], l[4], l[5], l[6], r[0], r[1], r[2], r[4], r[5], r[6]};
            for (int q = 0; q < 12; q++) o[4 + q] = float_bits(pk[q]);
        }
    });
    lap("pass 2 (interior records)");
    // the roots the split tracer / trace_rays start from
    out.trees.assign(s->trees, s->trees + s->numTrees);
    for (uint32_t m = 0; m < s->numMeshes; m++) out.trees[s->meshes[m].tree].root = ref[s->trees[s->meshes[m].tree].root];
    return true;
}
