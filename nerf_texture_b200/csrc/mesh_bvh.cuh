// mesh_bvh.cuh — data layout and per-query traversals of the mesh front end:
//   * nearest ray/triangle hit            (reference: external/RayTracer/src/bvh.cu:259-301 ray_intersect, :695-721 raytrace_kernel,
//                                          include/raytracing/triangle.cuh:27-39 Triangle::ray_intersect)
//   * K nearest mesh vertices in a radius (reference call sites tools/map.py:396,456: frnn.frnn_grid_points, K = 8, r = 100)
//
// Not a translation of the reference's 4-wide BVH of 8-triangle leaves: a binary tree whose 64-byte nodes carry BOTH children's
// boxes (one 4 x 128-bit fetch decides both), SAH-built leaves of <= 4 triangles stored as 48-byte records in leaf order, and the
// same node format over the mesh vertices for the neighbour search.  Results are those of an exhaustive scan with the reference's
// formulas: every pruning test here is conservative, and every arithmetic step that decides a result is written with explicit
// round-to-nearest intrinsics (no FMA contraction) so that host restatement and kernel agree bit for bit.
//
// The traversal functions compile for host as well: tests/native/mesh_host_check.cpp instantiates them with g++ to check the tree
// logic against exhaustive scans without a GPU.  The product library only ever runs them on the device (csrc/mesh.cu).
#pragma once

#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define NTX_HD __host__ __device__ __forceinline__
#else
#define NTX_HD inline
#endif

namespace ntx {
namespace mesh {

#if defined(__CUDA_ARCH__)
#define NTX_MUL(a, b) __fmul_rn((a), (b))
#define NTX_ADD(a, b) __fadd_rn((a), (b))
#define NTX_SUB(a, b) __fsub_rn((a), (b))
#define NTX_DIV(a, b) __fdiv_rn((a), (b))
#define NTX_SQRT(a) __fsqrt_rn((a))
#else  // host: built with -ffp-contract=off
#define NTX_MUL(a, b) ((a) * (b))
#define NTX_ADD(a, b) ((a) + (b))
#define NTX_SUB(a, b) ((a) - (b))
#define NTX_DIV(a, b) ((a) / (b))
#define NTX_SQRT(a) sqrtf((a))
#endif

constexpr float kMaxDist = 10.0f;       // bvh.cu:36 MAX_DIST: depth of a ray that hits nothing
constexpr float kNoHit = 1e6f;          // triangle.cuh:37
constexpr int kTriLeafMax = 4;
constexpr int kPointLeafMax = 8;
constexpr int kStackDepth = 64;         // tree depth is bounded by the builder (kSahDepth + log2 of the largest median-split range)
constexpr int kSahDepth = 30;

// Child link: >= 0 inner node index; < 0 leaf, ~link = first << 4 | count (count 0 = empty child).
NTX_HD int leaf_link(int first, int count) { return ~((first << 4) | count); }
NTX_HD int leaf_first(int link) { return (~link) >> 4; }
NTX_HD int leaf_count(int link) { return (~link) & 15; }

struct alignas(16) Node {   // 64 B
    float lo0[3], hi0[3], lo1[3], hi1[3];
    int c0, c1;
    int pad[2];
};
static_assert(sizeof(Node) == 64, "node layout");

struct alignas(16) Tri {    // 48 B, in leaf order; idx = position in the caller's triangle array
    float a[3], b[3], c[3];
    int idx;
    int pad[2];
};
static_assert(sizeof(Tri) == 48, "triangle layout");

struct alignas(16) Point {  // 16 B, in leaf order
    float p[3];
    int idx;
};

#if defined(__CUDA_ARCH__)
NTX_HD void fetch_node(const Node* nodes, int i, Node& n) {
    const float4* p = reinterpret_cast<const float4*>(nodes + i);
    float4 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2), q3 = __ldg(p + 3);
    n.lo0[0] = q0.x; n.lo0[1] = q0.y; n.lo0[2] = q0.z; n.hi0[0] = q0.w;
    n.hi0[1] = q1.x; n.hi0[2] = q1.y; n.lo1[0] = q1.z; n.lo1[1] = q1.w;
    n.lo1[2] = q2.x; n.hi1[0] = q2.y; n.hi1[1] = q2.z; n.hi1[2] = q2.w;
    n.c0 = __float_as_int(q3.x); n.c1 = __float_as_int(q3.y);
}
NTX_HD void fetch_tri(const Tri* tris, int i, Tri& t) {
    const float4* p = reinterpret_cast<const float4*>(tris + i);
    float4 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2);
    t.a[0] = q0.x; t.a[1] = q0.y; t.a[2] = q0.z; t.b[0] = q0.w;
    t.b[1] = q1.x; t.b[2] = q1.y; t.c[0] = q1.z; t.c[1] = q1.w;
    t.c[2] = q2.x; t.idx = __float_as_int(q2.y);
}
NTX_HD void fetch_point(const Point* pts, int i, Point& q) {
    float4 v = __ldg(reinterpret_cast<const float4*>(pts + i));
    q.p[0] = v.x; q.p[1] = v.y; q.p[2] = v.z; q.idx = __float_as_int(v.w);
}
#else
NTX_HD void fetch_node(const Node* nodes, int i, Node& n) { n = nodes[i]; }
NTX_HD void fetch_tri(const Tri* tris, int i, Tri& t) { t = tris[i]; }
NTX_HD void fetch_point(const Point* pts, int i, Point& q) { q = pts[i]; }
#endif

// dot / cross exactly as a scalar evaluation of Eigen's fixed-size expressions: ((x0*y0 + x1*y1) + x2*y2), (a1*b2 - a2*b1, ...)
NTX_HD float dot3(const float* x, const float* y) {
    return NTX_ADD(NTX_ADD(NTX_MUL(x[0], y[0]), NTX_MUL(x[1], y[1])), NTX_MUL(x[2], y[2]));
}
NTX_HD void cross3(const float* a, const float* b, float* r) {
    r[0] = NTX_SUB(NTX_MUL(a[1], b[2]), NTX_MUL(a[2], b[1]));
    r[1] = NTX_SUB(NTX_MUL(a[2], b[0]), NTX_MUL(a[0], b[2]));
    r[2] = NTX_SUB(NTX_MUL(a[0], b[1]), NTX_MUL(a[1], b[0]));
}

// triangle.cuh:27-39.  Returns t, or kNoHit; n = (b - a) x (c - a), not normalised.
NTX_HD float tri_intersect(const Tri& tr, const float* ro, const float* rd, float* n) {
    float v1v0[3], v2v0[3], rov0[3], q[3];
    for (int k = 0; k < 3; k++) {
        v1v0[k] = NTX_SUB(tr.b[k], tr.a[k]);
        v2v0[k] = NTX_SUB(tr.c[k], tr.a[k]);
        rov0[k] = NTX_SUB(ro[k], tr.a[k]);
    }
    cross3(v1v0, v2v0, n);
    cross3(rov0, rd, q);
    const float d = NTX_DIV(1.0f, dot3(rd, n));
    const float u = NTX_MUL(d, -dot3(q, v2v0));
    const float v = NTX_MUL(d, dot3(q, v1v0));
    float t = NTX_MUL(d, -dot3(n, rov0));
    if (u < 0.0f || u > 1.0f || v < 0.0f || NTX_ADD(u, v) > 1.0f || t < 0.0f) t = kNoHit;
    return t;
}

// Entry distance of the ray into a box, or +inf if it misses the slab interval [0, t_limit].
//  * d = 0 on an axis gives inv = +-inf and bounds of +-inf — except 0 * inf = NaN when the origin lies exactly IN a box face; that
//    bound is then the mirror of the other one, i.e. the slab does not constrain the ray (rays aimed at mesh vertices and edges from
//    a symmetric origin do this all the time).  A remaining NaN (flat box, both faces) is dropped by fminf/fmaxf.
//  * each bound carries 2 roundings: a ray through a box CORNER (a mesh vertex) has entry == exit in the reals, so the interval test
//    is widened by a few ulp.  Both rules only ever keep more boxes: they never cut a hit the exhaustive scan would find.
NTX_HD float box_entry(const float* lo, const float* hi, const float* ro, const float* inv, float t_limit, float eps_abs) {
    float tn = 0.0f, tf = t_limit;
    for (int k = 0; k < 3; k++) {
        float t0 = (lo[k] - ro[k]) * inv[k], t1 = (hi[k] - ro[k]) * inv[k];
        if (t0 != t0) t0 = -t1;
        if (t1 != t1) t1 = -t0;
        tn = fmaxf(tn, fminf(t0, t1));
        tf = fminf(tf, fmaxf(t0, t1));
    }
    return tn <= tf * 1.000001f + eps_abs ? tn : INFINITY;
}

struct Hit {
    float t;      // kMaxDist when nothing was hit (bvh.cu:263)
    int face;     // caller's triangle index, -1 when nothing was hit
    int slot;     // position in the leaf-ordered triangle array
};

// ---- the two steps of a nearest-hit traversal -------------------------------------------------------------------------------
// Inner node: which children can still hold a hit nearer than `best_t` (plus slack: the t a triangle test returns carries its own
// rounding error, the box test must not cut it).  Returns the number of children to visit (0, 1, 2), nearer one in *c_near.
NTX_HD int trace_node_step(const Node& nd, const float* ro, const float* inv, float best_t, float slack_abs, int* c_near, int* c_far) {
    const float limit = best_t + (slack_abs + 4e-6f * best_t);
    float e0 = box_entry(nd.lo0, nd.hi0, ro, inv, limit, slack_abs);
    float e1 = box_entry(nd.lo1, nd.hi1, ro, inv, limit, slack_abs);
    int c0 = nd.c0, c1 = nd.c1;
    if (e1 < e0) { const float te = e0; e0 = e1; e1 = te; const int tc = c0; c0 = c1; c1 = tc; }
    *c_near = c0;
    *c_far = c1;
    return (e0 < INFINITY) + (e1 < INFINITY);
}

// Leaf: nearest hit with t < kMaxDist.  Ties in t go to the lowest triangle index (what a scan in index order with `t < mint` finds).
NTX_HD void trace_leaf_step(const Tri* tris, int link, const float* ro, const float* rd, Hit& best) {
    const int first = leaf_first(link), count = leaf_count(link);
    for (int i = 0; i < count; i++) {
        Tri tr;
        fetch_tri(tris, first + i, tr);
        float n[3];
        const float t = tri_intersect(tr, ro, rd, n);
        if (t < best.t || (t == best.t && best.face >= 0 && tr.idx < best.face)) {
            best.t = t; best.face = tr.idx; best.slot = first + i;
        }
    }
}

// One query, start to end.
NTX_HD Hit trace_one(const Node* nodes, const Tri* tris, const float* ro, const float* rd, float slack_abs) {
    Hit best{kMaxDist, -1, -1};
    const float inv[3] = {1.0f / rd[0], 1.0f / rd[1], 1.0f / rd[2]};
    int stack[kStackDepth];
    int sp = 0;
    int cur = 0;   // node 0 is the root
    for (;;) {
        if (cur >= 0) {
            Node nd;
            fetch_node(nodes, cur, nd);
            int c_near, c_far;
            const int n = trace_node_step(nd, ro, inv, best.t, slack_abs, &c_near, &c_far);
            if (n > 0) {
                cur = c_near;
                if (n > 1 && sp < kStackDepth) stack[sp++] = c_far;
                continue;
            }
        } else {
            trace_leaf_step(tris, cur, ro, rd, best);
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return best;
}

// triangle.cuh:23-25 normal(): (b - a) x (c - a), Eigen normalized() = v / sqrt(v.v) when v.v > 0
NTX_HD void tri_normal(const Tri& tr, float* n) {
    float e1[3], e2[3];
    for (int k = 0; k < 3; k++) { e1[k] = NTX_SUB(tr.b[k], tr.a[k]); e2[k] = NTX_SUB(tr.c[k], tr.a[k]); }
    cross3(e1, e2, n);
    const float z = dot3(n, n);
    if (z > 0.0f) {
        const float len = NTX_SQRT(z);
        n[0] = NTX_DIV(n[0], len); n[1] = NTX_DIV(n[1], len); n[2] = NTX_DIV(n[2], len);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// K nearest points with squared distance < r2, ascending by (distance, index).
// Distances are ((dx*dx + dy*dy) + dz*dz) of d = q - p; the box bound uses the same operation order on the clamped offsets, and
// rounding is monotonic, so bound <= distance of every point in the box holds in floating point, not just in the reals.
//
// The candidate list is K (compile-time) slots sorted by (distance, index) and only ever indexed by unrolled loop counters, so it
// lives in registers.  A query for k_want < K neighbours pre-fills the first K - k_want slots with entries that beat everything
// (distance -1): the real neighbours then occupy the last k_want slots and "the worst entry still of interest" is always slot K-1.
// Empty slots hold (+inf, INT_MAX), which every real candidate beats.
NTX_HD float sq_dist3(float dx, float dy, float dz) { return NTX_ADD(NTX_ADD(NTX_MUL(dx, dx), NTX_MUL(dy, dy)), NTX_MUL(dz, dz)); }

NTX_HD float box_sq_dist(const float* lo, const float* hi, const float* q) {
    float d[3];
    for (int k = 0; k < 3; k++) d[k] = fmaxf(fmaxf(NTX_SUB(lo[k], q[k]), NTX_SUB(q[k], hi[k])), 0.0f);
    return sq_dist3(d[0], d[1], d[2]);
}

constexpr int kEmptyIdx = 0x7fffffff;

template <int K>
NTX_HD void knn_list_init(float* bd, int* bi, int k_want) {
#pragma unroll
    for (int s = 0; s < K; s++) {
        const bool filler = s < K - k_want;
        bd[s] = filler ? -1.0f : INFINITY;
        bi[s] = filler ? 0 : kEmptyIdx;
    }
}

// s-th neighbour (0-based, s < k_want) after the search: distance / index, or -1 / -1 when fewer were in range.  `s` must be an
// unrolled loop counter of a loop over 0..K-1 for the list to stay in registers: iterate slots, not neighbours (see the kernels).
NTX_HD bool knn_slot_valid(int bi_s) { return bi_s != kEmptyIdx; }

NTX_HD int knn_node_step(const Node& nd, const float* q, float r2, float worst, int* c_near, int* c_far) {
    float e0 = (nd.c0 < 0 && leaf_count(nd.c0) == 0) ? INFINITY : box_sq_dist(nd.lo0, nd.hi0, q);
    float e1 = (nd.c1 < 0 && leaf_count(nd.c1) == 0) ? INFINITY : box_sq_dist(nd.lo1, nd.hi1, q);
    int c0 = nd.c0, c1 = nd.c1;
    if (e1 < e0) { const float te = e0; e0 = e1; e1 = te; const int tc = c0; c0 = c1; c1 = tc; }
    *c_near = c0;
    *c_far = c1;
    // beyond the radius nothing counts; with the list full, nothing beyond its last entry (equal distance: a lower index still does)
    return (e0 < r2 && e0 <= worst) + (e1 < r2 && e1 <= worst);
}

template <int K>
NTX_HD void knn_leaf_step(const Point* pts, int link, const float* q, float r2, float* bd, int* bi) {
    const int first = leaf_first(link), count = leaf_count(link);
    for (int i = 0; i < count; i++) {
        Point p;
        fetch_point(pts, first + i, p);
        const float d2 = sq_dist3(NTX_SUB(q[0], p.p[0]), NTX_SUB(q[1], p.p[1]), NTX_SUB(q[2], p.p[2]));
        if (!(d2 < r2)) continue;
        if (!(d2 < bd[K - 1] || (d2 == bd[K - 1] && p.idx < bi[K - 1]))) continue;
        float cd = d2;
        int ci = p.idx;
#pragma unroll
        for (int s = 0; s < K; s++) {   // sorted insertion: the carried entry sinks to its slot, the rest shift down, the last falls out
            const bool before = cd < bd[s] || (cd == bd[s] && ci < bi[s]);
            const float td = bd[s];
            const int ti = bi[s];
            bd[s] = before ? cd : td; bi[s] = before ? ci : ti;
            cd = before ? td : cd;    ci = before ? ti : ci;
        }
    }
}

// One query, start to end.  Returns the number of neighbours found (<= k_want); they are in slots K - k_want ... K - 1.
template <int K>
NTX_HD int knn_one(const Node* nodes, const Point* pts, const float* q, float r2, int k_want, float* bd, int* bi) {
    knn_list_init<K>(bd, bi, k_want);
    int stack[kStackDepth];
    int sp = 0;
    int cur = 0;
    for (;;) {
        if (cur >= 0) {
            Node nd;
            fetch_node(nodes, cur, nd);
            int c_near, c_far;
            const int n = knn_node_step(nd, q, r2, bd[K - 1], &c_near, &c_far);
            if (n > 0) {
                cur = c_near;
                if (n > 1 && sp < kStackDepth) stack[sp++] = c_far;
                continue;
            }
        } else {
            knn_leaf_step<K>(pts, cur, q, r2, bd, bi);
        }
        if (sp == 0) break;   // a popped subtree is re-tested against the current list when its node is fetched
        cur = stack[--sp];
    }
    int found = 0;
#pragma unroll
    for (int s = 0; s < K; s++) found += (s >= K - k_want) && knn_slot_valid(bi[s]);
    return found;
}

}  // namespace mesh
}  // namespace ntx
