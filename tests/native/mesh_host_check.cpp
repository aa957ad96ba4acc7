// tests/native/mesh_host_check.cpp — TEST INFRASTRUCTURE: compiles the product's tree builder (csrc/mesh_build.h) and the per-query
// traversals of csrc/mesh_bvh.cuh for the HOST, so that `pytest -m "not gpu"` can check the tree logic (every primitive in exactly
// one leaf, conservative pruning, tie rules) against the exhaustive oracle without a GPU.  Built by tests/_util.build_mesh_host_check()
// into tests/native/_build/ (git-ignored); nothing in the product loads it.
#include <cstdint>
#include <vector>

#include "../../nerf_texture_b200/csrc/mesh_build.h"

using namespace ntx::mesh;

extern "C" {

// returns tree depth (< 0: invalid mesh); stats[0] = nodes, [1] = leaf slots covered exactly once (must equal n), [2] = boxes that fail to contain their primitives
int hostcheck_tree(const float* vertices, uint32_t n_v, const int32_t* triangles, uint32_t n_t, int points, int64_t* stats) {
    BuildResult tree;
    std::vector<Tri> tris;
    std::vector<Point> pts;
    const uint32_t n = points ? n_v : n_t;
    if (points) { if (!build_point_tree(vertices, n_v, tree, pts)) return -1; }
    else if (!build_triangle_tree(vertices, n_v, triangles, n_t, tree, tris)) return -1;
    std::vector<int> seen(n, 0);
    int64_t bad_box = 0;
    for (const Node& nd : tree.nodes) {
        for (int c = 0; c < 2; c++) {
            const int link = c ? nd.c1 : nd.c0;
            const float* lo = c ? nd.lo1 : nd.lo0;
            const float* hi = c ? nd.hi1 : nd.hi0;
            if (link >= 0) continue;
            for (int i = 0; i < leaf_count(link); i++) {
                const int s = leaf_first(link) + i;
                seen[tree.order[s]]++;
                for (int k = 0; k < 3; k++) {
                    if (points) { if (pts[s].p[k] < lo[k] || pts[s].p[k] > hi[k]) bad_box++; }
                    else for (const float* v : {tris[s].a, tris[s].b, tris[s].c}) if (v[k] < lo[k] || v[k] > hi[k]) bad_box++;
                }
            }
        }
    }
    int64_t once = 0;
    for (uint32_t i = 0; i < n; i++) once += seen[i] == 1;
    stats[0] = (int64_t)tree.nodes.size(); stats[1] = once; stats[2] = bad_box;
    return tree.depth;
}

int hostcheck_trace(const float* vertices, uint32_t n_v, const int32_t* triangles, uint32_t n_t, const float* rays_o, const float* rays_d,
                    uint32_t N, float slack, float* depth, int64_t* face) {
    BuildResult tree;
    std::vector<Tri> tris;
    if (!build_triangle_tree(vertices, n_v, triangles, n_t, tree, tris)) return -1;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)N; i++) {
        const Hit h = trace_one(tree.nodes.data(), tris.data(), rays_o + 3 * i, rays_d + 3 * i, slack);
        depth[i] = h.t;
        face[i] = h.face;
    }
    return tree.depth;
}

int hostcheck_knn(const float* points, uint32_t n_p, const float* queries, uint32_t N, uint32_t K, float r, float* dists, int64_t* idxs) {
    BuildResult tree;
    std::vector<Point> pts;
    if (K > 32 || !build_point_tree(points, n_p, tree, pts)) return -1;
    const float r2 = r * r;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)N; i++) {
        float bd[32]; int bi[32];
        const int found = knn_one<32>(tree.nodes.data(), pts.data(), queries + 3 * i, r2, (int)K, bd, bi);
        for (uint32_t s = 0; s < K; s++) {   // the neighbours sit in the LAST K slots of the list (mesh_bvh.cuh)
            dists[(size_t)i * K + s] = (int)s < found ? bd[32 - K + s] : -1.0f;
            idxs[(size_t)i * K + s] = (int)s < found ? bi[32 - K + s] : -1;
        }
    }
    return tree.depth;
}

}
