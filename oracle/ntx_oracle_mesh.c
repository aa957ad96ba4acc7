/* oracle/ntx_oracle_mesh.c — CPU restatement of the mesh front end of the texture field.
 *
 * TEST INFRASTRUCTURE ONLY (same rule as ntx_oracle.c): only tests/, __graft_entry__.smoke() and bench.py's CPU legs use it.
 *
 * What it restates (paths relative to /root/reference):
 *   external/RayTracer/include/raytracing/triangle.cuh:27-39   Triangle::ray_intersect (the ray/triangle formula)
 *   external/RayTracer/include/raytracing/triangle.cuh:23-25   Triangle::normal
 *   external/RayTracer/src/bvh.cu:259-301                       ray_intersect: nearest t below MAX_DIST = 10 (:36), `t < mint`
 *   external/RayTracer/src/bvh.cu:695-721                       raytrace_kernel: depth, position = ro + t rd, normal, face_idx
 *   tools/map.py:454-500                                        MeshProjector.knn   (K nearest vertices -> coarse normal)
 *   tools/map.py:414-433                                        MeshProjector.project (two traces along +-normal -> p_sur, sdf)
 *
 * PARITY UNPINNED, for two reasons that cannot be fixed in this image:
 *   - external/RayTracer needs Eigen, which is absent, so the reference's BVH cannot be compiled here and there are no golden vectors
 *     (test_data/ holds three .obj files and no script).  The BVH only prunes: its result is the exhaustive scan below, up to (a) which
 *     of two triangles with the SAME t wins (traversal order there, lowest index here) and (b) nvcc's FMA contraction inside the Eigen
 *     expressions, which a restatement cannot reproduce; here every product and sum is rounded separately, left to right.
 *   - frnn (github.com/lxxue/FRNN, unpinned, readme.md:37) is not in the tree at all.  Its published contract — the K nearest points of
 *     points2 within radius r of each query, ascending, squared distances, -1 padding — is restated as an exhaustive scan; r = 100
 *     at both call sites (tools/map.py:396,456) exceeds every scene, so K neighbours always exist there.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define ORC_MAX_DIST 10.0f

static inline float dot3(const float* x, const float* y) { return (x[0] * y[0] + x[1] * y[1]) + x[2] * y[2]; }
static inline void cross3(const float* a, const float* b, float* r)
{
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}

/* triangle.cuh:27-39 */
static inline float tri_ray(const float* a, const float* b, const float* c, const float* ro, const float* rd)
{
    float v1v0[3], v2v0[3], rov0[3], n[3], q[3];
    for (int k = 0; k < 3; k++) { v1v0[k] = b[k] - a[k]; v2v0[k] = c[k] - a[k]; rov0[k] = ro[k] - a[k]; }
    cross3(v1v0, v2v0, n);
    cross3(rov0, rd, q);
    const float d = 1.0f / dot3(rd, n);
    const float u = d * -dot3(q, v2v0);
    const float v = d * dot3(q, v1v0);
    float t = d * -dot3(n, rov0);
    if (u < 0.0f || u > 1.0f || v < 0.0f || (u + v) > 1.0f || t < 0.0f) t = 1e6f;
    return t;
}

static void trace_one(const float* vertices, const int32_t* triangles, uint32_t n_tri, const float* ro, const float* rd, float* pos,
                      float* nrm, float* depth, int64_t* face)
{
    float mint = ORC_MAX_DIST;   /* bvh.cu:263 */
    int64_t best = -1;
    for (uint32_t i = 0; i < n_tri; i++) {
        const float* a = vertices + 3 * (size_t)triangles[3 * i];
        const float* b = vertices + 3 * (size_t)triangles[3 * i + 1];
        const float* c = vertices + 3 * (size_t)triangles[3 * i + 2];
        const float t = tri_ray(a, b, c, ro, rd);
        if (t < mint) { mint = t; best = i; }   /* bvh.cu:275 */
    }
    *depth = mint;                                             /* bvh.cu:705 */
    for (int k = 0; k < 3; k++) pos[k] = ro[k] + mint * rd[k]; /* bvh.cu:709 */
    if (best >= 0) {                                           /* bvh.cu:712-714; triangle.cuh:23-25 */
        const float* a = vertices + 3 * (size_t)triangles[3 * best];
        const float* b = vertices + 3 * (size_t)triangles[3 * best + 1];
        const float* c = vertices + 3 * (size_t)triangles[3 * best + 2];
        float e1[3], e2[3], n[3];
        for (int k = 0; k < 3; k++) { e1[k] = b[k] - a[k]; e2[k] = c[k] - a[k]; }
        cross3(e1, e2, n);
        const float z = dot3(n, n);
        if (z > 0.0f) { const float len = sqrtf(z); n[0] /= len; n[1] /= len; n[2] /= len; }
        nrm[0] = n[0]; nrm[1] = n[1]; nrm[2] = n[2];
        *face = best;
    } else {
        nrm[0] = nrm[1] = nrm[2] = 0.0f;                       /* bvh.cu:716; face_idx keeps the caller's -1 (raytracer.py:37) */
    }
}

void orc_mesh_trace(const float* vertices, const int32_t* triangles, uint32_t n_tri, const float* rays_o, const float* rays_d, uint32_t N,
                    float* positions, float* normals, float* depth, int64_t* face_idx)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)N; i++)
        trace_one(vertices, triangles, n_tri, rays_o + 3 * i, rays_d + 3 * i, positions + 3 * i, normals + 3 * i, depth + i, face_idx + i);
}

/* K nearest of `points` with squared distance < r*r, ascending by (distance, index); -1 padding */
static void knn_one(const float* points, uint32_t n_points, const float* q, uint32_t K, float r2, float* bd, int64_t* bi)
{
    uint32_t found = 0;
    for (uint32_t s = 0; s < K; s++) { bd[s] = -1.0f; bi[s] = -1; }
    for (uint32_t i = 0; i < n_points; i++) {
        const float dx = q[0] - points[3 * (size_t)i], dy = q[1] - points[3 * (size_t)i + 1], dz = q[2] - points[3 * (size_t)i + 2];
        const float d2 = (dx * dx + dy * dy) + dz * dz;
        if (!(d2 < r2)) continue;
        if (found == K && !(d2 < bd[K - 1])) continue;   /* scan is in index order: an equal distance never displaces an earlier index */
        uint32_t s = found < K ? found++ : K - 1;
        while (s > 0 && d2 < bd[s - 1]) { bd[s] = bd[s - 1]; bi[s] = bi[s - 1]; s--; }
        bd[s] = d2; bi[s] = i;
    }
}

void orc_points_knn(const float* points, uint32_t n_points, const float* queries, uint32_t N, uint32_t K, float r, float* dists, int64_t* idxs)
{
    const float r2 = r * r;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)N; i++) knn_one(points, n_points, queries + 3 * i, K, r2, dists + (size_t)i * K, idxs + (size_t)i * K);
}

static inline float norm3(const float* v) { return sqrtf((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }

/* tools/map.py:454-500 with the arguments project() passes: use_dir_vec=True, weighting='Shepard', no consistency checks.
 * K <= 32.  Sums run in index order (torch's reduction order over a K-long dimension is unspecified: tolerance, not bit-exact). */
static void coarse_normal(const float* vertices, const float* vertex_normals, uint32_t n_vertices, const float* x, uint32_t K, float r2,
                          float dir_vec_wdist, float* normal)
{
    float d2[32], dis[33], nrm[33][3];
    int64_t idx[32];
    knn_one(vertices, n_vertices, x, K, r2, d2, idx);
    float mean_dir[3] = {0, 0, 0}, ntest[3] = {0, 0, 0};
    for (uint32_t k = 0; k < K; k++) {
        int64_t j = idx[k] < 0 ? (int64_t)n_vertices - 1 : idx[k];   /* torch indexing with -1 */
        dis[k] = sqrtf(d2[k]);                                       /* :458 (NaN for the -1 padding) */
        float dvo[3], dv[3];
        for (int c = 0; c < 3; c++) { nrm[k][c] = vertex_normals[3 * j + c]; dvo[c] = x[c] - vertices[3 * j + c]; }   /* :459-460 */
        const float len = norm3(dvo) + 1e-5f;
        for (int c = 0; c < 3; c++) dv[c] = dvo[c] / len;            /* :461 */
        const float w = 1.0f / (dis[k] + 1e-7f);                      /* :474 */
        for (int c = 0; c < 3; c++) { mean_dir[c] += w * dv[c]; ntest[c] += nrm[k][c]; }   /* :476-477 */
    }
    for (int c = 0; c < 3; c++) ntest[c] /= (float)K;
    if (dot3(mean_dir, ntest) < 0.0f) for (int c = 0; c < 3; c++) mean_dir[c] = -mean_dir[c];   /* :478 */
    {
        const float len = norm3(mean_dir) + 1e-5f;                    /* :479 */
        for (int c = 0; c < 3; c++) nrm[K][c] = mean_dir[c] / len;    /* :480 */
    }
    dis[K] = dir_vec_wdist < 1e-5f ? 1e-5f : dir_vec_wdist;           /* :481-482 */
    float w[33], wsum = 0.0f;
    for (uint32_t k = 0; k <= K; k++) { w[k] = 1.0f / (dis[k] + 1e-7f); wsum += w[k]; }   /* :487 */
    float acc[3] = {0, 0, 0};
    for (uint32_t k = 0; k <= K; k++) {
        const float wk = w[k] / wsum;                                 /* :496 */
        const float len = norm3(nrm[k]) + 1e-5f;                      /* :497 */
        for (int c = 0; c < 3; c++) acc[c] += (nrm[k][c] / len) * wk; /* :498 */
    }
    const float len = norm3(acc) + 1e-5f;
    for (int c = 0; c < 3; c++) normal[c] = acc[c] / len;             /* :499 */
}

/* tools/map.py:414-433 (up to the tbn gather and h_mask, which stay torch one-liners on the outputs) */
void orc_mesh_project(const float* vertices, const float* vertex_normals, uint32_t n_vertices, const int32_t* triangles, uint32_t n_tri,
                      const float* xyz, uint32_t N, uint32_t K, float r, float dir_vec_wdist, float* p_sur, float* sdf, float* normal,
                      int64_t* face_idx)
{
    const float r2 = r * r;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < (int64_t)N; i++) {
        const float* x = xyz + 3 * i;
        float n[3], nn[3], p1[3], p2[3], fn[3], d1, d2;
        int64_t f1 = -1, f2 = -1;
        coarse_normal(vertices, vertex_normals, n_vertices, x, K, r2, dir_vec_wdist, n);
        for (int c = 0; c < 3; c++) nn[c] = -n[c];
        trace_one(vertices, triangles, n_tri, x, n, p1, fn, &d1, &f1);    /* :419 inner */
        trace_one(vertices, triangles, n_tri, x, nn, p2, fn, &d2, &f2);   /* :420 outer */
        const int cond = d1 < d2;                                          /* :421 */
        for (int c = 0; c < 3; c++) { p_sur[3 * i + c] = cond ? p1[c] : p2[c]; normal[3 * i + c] = n[c]; }
        sdf[i] = cond ? -d1 : d2;                                          /* :423 */
        face_idx[i] = cond ? f1 : f2;                                      /* :425 */
    }
}
