// mesh_build.h — host-side tree construction for mesh_bvh.cuh (plain C++, no CUDA): runs once per mesh, like the reference's
// TriangleBvh::build (external/RayTracer/src/bvh.cu:531-606, called from the RayTracerImpl constructor, raytracer.cu:37).
//
// The reference partitions by the median of the axis with the largest centroid variance into a 4-wide tree with <= 8 triangles
// per leaf.  Here: binned surface-area heuristic (16 bins, three axes) for the triangle tree — rays prune by box entry distance, so
// tight low-overlap boxes pay on every one of the millions of queries per frame — and object-median splits for the vertex tree
// (balanced, nearest-neighbour pruning only needs compact boxes).  Splits are deterministic: ties are broken by primitive index.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "mesh_bvh.cuh"

namespace ntx {
namespace mesh {

struct Prim {
    float lo[3], hi[3], c[3];
};

struct BuildResult {
    std::vector<Node> nodes;   // nodes[0] is the root
    std::vector<int> order;    // leaf slot -> primitive index
    int depth = 0;
};

class TreeBuilder {
public:
    TreeBuilder(const std::vector<Prim>& prims, int leaf_max, bool sah) : prims_(prims), leaf_max_(leaf_max), sah_(sah) {}

    BuildResult run() {
        BuildResult r;
        const int n = (int)prims_.size();
        r.order.resize(n);
        for (int i = 0; i < n; i++) r.order[i] = i;
        order_ = &r.order;
        nodes_ = &r.nodes;
        depth_ = 0;
        float lo[3], hi[3];
        const int root = n > 0 ? build(0, n, 1, lo, hi) : leaf_link(0, 0);
        if (root < 0) {   // everything fits one leaf (or the mesh is empty): wrap it so that node 0 exists
            Node nd;
            std::memset(&nd, 0, sizeof(nd));
            set_box(nd.lo0, nd.hi0, lo, hi, n > 0);
            set_box(nd.lo1, nd.hi1, lo, hi, false);
            nd.c0 = root;
            nd.c1 = leaf_link(0, 0);
            r.nodes.push_back(nd);
        }
        r.depth = depth_;
        return r;
    }

private:
    static void set_box(float* dlo, float* dhi, const float* lo, const float* hi, bool valid) {
        for (int k = 0; k < 3; k++) {
            dlo[k] = valid ? lo[k] : INFINITY;
            dhi[k] = valid ? hi[k] : -INFINITY;
        }
    }
    static float half_area(const float* lo, const float* hi) {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx * dy + dy * dz + dz * dx;
    }

    int build(int begin, int end, int depth, float* lo, float* hi) {
        std::vector<int>& order = *order_;
        depth_ = std::max(depth_, depth);
        const int count = end - begin;
        float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int k = 0; k < 3; k++) { lo[k] = INFINITY; hi[k] = -INFINITY; }
        for (int i = begin; i < end; i++) {
            const Prim& p = prims_[order[i]];
            for (int k = 0; k < 3; k++) {
                lo[k] = std::min(lo[k], p.lo[k]); hi[k] = std::max(hi[k], p.hi[k]);
                clo[k] = std::min(clo[k], p.c[k]); chi[k] = std::max(chi[k], p.c[k]);
            }
        }
        if (count <= leaf_max_) return leaf_link(begin, count);

        int mid = -1;
        if (sah_ && depth < kSahDepth && count > 2 * leaf_max_) mid = sah_split(begin, end, clo, chi);
        if (mid <= begin || mid >= end) {   // object median along the widest centroid axis
            int axis = 0;
            for (int k = 1; k < 3; k++) if (chi[k] - clo[k] > chi[axis] - clo[axis]) axis = k;
            mid = begin + count / 2;
            std::nth_element(order.begin() + begin, order.begin() + mid, order.begin() + end, [&](int a, int b) {
                const float ca = prims_[a].c[axis], cb = prims_[b].c[axis];
                return ca < cb || (ca == cb && a < b);
            });
        }
        const int me = (int)nodes_->size();
        nodes_->emplace_back();
        float l0[3], h0[3], l1[3], h1[3];
        const int c0 = build(begin, mid, depth + 1, l0, h0);
        const int c1 = build(mid, end, depth + 1, l1, h1);
        Node& nd = (*nodes_)[me];
        std::memset(&nd, 0, sizeof(nd));
        set_box(nd.lo0, nd.hi0, l0, h0, true);
        set_box(nd.lo1, nd.hi1, l1, h1, true);
        nd.c0 = c0;
        nd.c1 = c1;
        return me;
    }

    // Returns the partition point of the cheapest of 3 x 15 binned planes, or -1 when the centroids do not spread.
    int sah_split(int begin, int end, const float* clo, const float* chi) {
        constexpr int kBins = 16;
        std::vector<int>& order = *order_;
        float best_cost = INFINITY;
        int best_axis = -1, best_bin = -1;
        for (int axis = 0; axis < 3; axis++) {
            const float ext = chi[axis] - clo[axis];
            if (!(ext > 0.0f)) continue;
            const float scale = kBins / ext;
            int cnt[kBins] = {0};
            float blo[kBins][3], bhi[kBins][3];
            for (int b = 0; b < kBins; b++) for (int k = 0; k < 3; k++) { blo[b][k] = INFINITY; bhi[b][k] = -INFINITY; }
            for (int i = begin; i < end; i++) {
                const Prim& p = prims_[order[i]];
                const int b = std::min(kBins - 1, std::max(0, (int)((p.c[axis] - clo[axis]) * scale)));
                cnt[b]++;
                for (int k = 0; k < 3; k++) { blo[b][k] = std::min(blo[b][k], p.lo[k]); bhi[b][k] = std::max(bhi[b][k], p.hi[k]); }
            }
            float rarea[kBins];
            int rcnt[kBins];
            float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
            int c = 0;
            for (int b = kBins - 1; b > 0; b--) {
                for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], blo[b][k]); hi[k] = std::max(hi[k], bhi[b][k]); }
                c += cnt[b];
                rarea[b] = c ? half_area(lo, hi) : 0.0f;
                rcnt[b] = c;
            }
            for (int k = 0; k < 3; k++) { lo[k] = INFINITY; hi[k] = -INFINITY; }
            c = 0;
            for (int b = 0; b < kBins - 1; b++) {   // plane between bin b and b + 1
                for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], blo[b][k]); hi[k] = std::max(hi[k], bhi[b][k]); }
                c += cnt[b];
                if (c == 0 || rcnt[b + 1] == 0) continue;
                const float cost = half_area(lo, hi) * c + rarea[b + 1] * rcnt[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
            }
        }
        if (best_axis < 0) return -1;
        const float scale = kBins / (chi[best_axis] - clo[best_axis]);
        auto it = std::stable_partition(order.begin() + begin, order.begin() + end, [&](int a) {
            const int b = std::min(kBins - 1, std::max(0, (int)((prims_[a].c[best_axis] - clo[best_axis]) * scale)));
            return b <= best_bin;
        });
        return (int)(it - order.begin());
    }

    const std::vector<Prim>& prims_;
    int leaf_max_;
    bool sah_;
    std::vector<int>* order_ = nullptr;
    std::vector<Node>* nodes_ = nullptr;
    int depth_ = 0;
};

// triangles [m,3] index into vertices [n,3]; returns false on an out-of-range index or a non-finite coordinate
inline bool build_triangle_tree(const float* vertices, uint32_t n_vertices, const int32_t* triangles, uint32_t n_triangles, BuildResult& tree,
                                std::vector<Tri>& tris) {
    for (uint64_t i = 0; i < (uint64_t)n_vertices * 3; i++) if (!std::isfinite(vertices[i])) return false;
    std::vector<Prim> prims(n_triangles);
    for (uint32_t i = 0; i < n_triangles; i++) {
        Prim& p = prims[i];
        for (int k = 0; k < 3; k++) { p.lo[k] = INFINITY; p.hi[k] = -INFINITY; p.c[k] = 0.0f; }
        for (int c = 0; c < 3; c++) {
            const int32_t vi = triangles[3 * i + c];
            if (vi < 0 || (uint32_t)vi >= n_vertices) return false;
            for (int k = 0; k < 3; k++) {
                const float x = vertices[3 * (size_t)vi + k];
                p.lo[k] = std::min(p.lo[k], x); p.hi[k] = std::max(p.hi[k], x);
            }
        }
        for (int k = 0; k < 3; k++) p.c[k] = 0.5f * (p.lo[k] + p.hi[k]);
    }
    tree = TreeBuilder(prims, kTriLeafMax, true).run();
    tris.resize(n_triangles);
    for (uint32_t s = 0; s < n_triangles; s++) {
        const int i = tree.order[s];
        Tri& t = tris[s];
        std::memset(&t, 0, sizeof(t));
        for (int k = 0; k < 3; k++) {
            t.a[k] = vertices[3 * (size_t)triangles[3 * i + 0] + k];
            t.b[k] = vertices[3 * (size_t)triangles[3 * i + 1] + k];
            t.c[k] = vertices[3 * (size_t)triangles[3 * i + 2] + k];
        }
        t.idx = i;
    }
    return true;
}

inline bool build_point_tree(const float* points, uint32_t n_points, BuildResult& tree, std::vector<Point>& pts) {
    for (uint64_t i = 0; i < (uint64_t)n_points * 3; i++) if (!std::isfinite(points[i])) return false;
    std::vector<Prim> prims(n_points);
    for (uint32_t i = 0; i < n_points; i++)
        for (int k = 0; k < 3; k++) prims[i].lo[k] = prims[i].hi[k] = prims[i].c[k] = points[3 * (size_t)i + k];
    tree = TreeBuilder(prims, kPointLeafMax, false).run();
    pts.resize(n_points);
    for (uint32_t s = 0; s < n_points; s++) {
        const int i = tree.order[s];
        for (int k = 0; k < 3; k++) pts[s].p[k] = points[3 * (size_t)i + k];
        pts[s].idx = i;
    }
    return true;
}

}  // namespace mesh
}  // namespace ntx
