// mesh.cu — the mesh front end of the texture field behind the C ABI: nearest ray/triangle hits
// (external/RayTracer, `RayTracer.trace`), K nearest mesh vertices (frnn.frnn_grid_points at tools/map.py:396,456) and the whole of
// MeshProjector.project (tools/map.py:414-433) as ONE kernel per batch of samples.
//
// Trees in the 64-byte two-box node format of mesh_bvh.cuh, read through the read-only path (the trees of the reference's meshes —
// 10^4..10^6 triangles, <= 100 MB — live in the 126 MB L2 after the first wave; one 4 x LDG.128 node fetch decides both children).
// One thread per query: neighbouring samples of a camera ray walk nearly the same nodes, and the per-thread state (a 64-entry
// stack in local memory, the K-entry neighbour list in registers) is small.  Two other mappings were built, verified bit-exact and
// measured slower on B200 ("Tried and lost"; profiles/r02_mesh_ab_*): lanes pulling queries from a per-warp chunk, and
// 8 lanes per query on an 8-wide tree.
#include "common.cuh"
#include "mesh_build.h"

#include <mutex>
#include <new>

namespace ntx {
namespace mesh {

constexpr uint32_t kMeshMagic = 0x4d53484eu;

struct MeshHandle {
    uint32_t magic;
    int device;
    uint32_t n_vertices, n_triangles;
    uint32_t tri_nodes_n, pt_nodes_n, tri_depth, pt_depth;
    float slack;          // pruning slack of trace_one: 1e-5 of the mesh's largest extent
    Node* tri_nodes;
    Tri* tris;
    Node* pt_nodes;
    Point* pts;
    float* vertices;      // [n_vertices, 3] in the caller's order (the fused projection gathers neighbours by index)
};

static void free_handle(MeshHandle* h) {
    if (!h) return;
    cudaFree(h->tri_nodes); cudaFree(h->tris); cudaFree(h->pt_nodes); cudaFree(h->pts); cudaFree(h->vertices);
    h->magic = 0;
    delete h;
}

template <typename T>
static bool upload(T** dst, const std::vector<T>& src) {
    *dst = nullptr;
    const size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    if (cudaMalloc((void**)dst, bytes) != cudaSuccess) return false;
    if (!src.empty() && cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    return true;
}

__device__ __forceinline__ void load3(const float* p, uint32_t i, float* v) {
    v[0] = p[3 * (size_t)i]; v[1] = p[3 * (size_t)i + 1]; v[2] = p[3 * (size_t)i + 2];
}

// bvh.cu:695-721 raytrace_kernel.  positions may alias rays_o and normals rays_d (raytracer.py:52-54 `inplace`): a thread reads its
// ray before it writes anything.
__global__ void __launch_bounds__(128) mesh_trace_kernel(uint32_t N, const float* rays_o, const float* rays_d, float* positions, float* normals,
                                                         float* depth, long long* face_idx, const Node* __restrict__ nodes,
                                                         const Tri* __restrict__ tris, float slack) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float ro[3], rd[3];
    load3(rays_o, i, ro);
    load3(rays_d, i, rd);
    const Hit h = trace_one(nodes, tris, ro, rd, slack);
    depth[i] = h.t;
    for (int k = 0; k < 3; k++) positions[3 * (size_t)i + k] = NTX_ADD(ro[k], NTX_MUL(h.t, rd[k]));
    float n[3] = {0.0f, 0.0f, 0.0f};
    if (h.face >= 0) {
        Tri tr;
        fetch_tri(tris, h.slot, tr);
        tri_normal(tr, n);
        face_idx[i] = h.face;   // a miss leaves the caller's value (the reference pre-fills -1, raytracer.py:37)
    }
    for (int k = 0; k < 3; k++) normals[3 * (size_t)i + k] = n[k];
}

template <int K>
__global__ void __launch_bounds__(128) mesh_knn_kernel(uint32_t N, const float* queries, float r2, int k_want, const Node* __restrict__ nodes,
                                                       const Point* __restrict__ pts, float* dists, long long* idxs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float q[3];
    load3(queries, i, q);
    float bd[K];
    int bi[K];
    knn_one<K>(nodes, pts, q, r2, k_want, bd, bi);
    const size_t row = (size_t)i * k_want;
#pragma unroll
    for (int s = 0; s < K; s++) {
        if (s >= K - k_want) {
            const bool ok = knn_slot_valid(bi[s]);
            dists[row + s - (K - k_want)] = ok ? bd[s] : -1.0f;
            idxs[row + s - (K - k_want)] = ok ? bi[s] : -1;
        }
    }
}

__device__ __forceinline__ float norm3(const float* v) { return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

// tools/map.py:454-500 (knn with use_dir_vec=True, weighting='Shepard'): the terms one neighbour contributes.  The arithmetic is plain
// fp32 (the reference's is a chain of ~30 torch kernels over [N,K,3] temporaries whose reduction order is not specified — parity is to
// tolerance there, see the tests).  sums = {mean_dir[3], normal_test[3], weighted normals[3], weight sum}
__device__ __forceinline__ void neighbour_terms(const float* x, bool ok, float d2, int idx, const float* __restrict__ vertices,
                                                const float* __restrict__ vertex_normals, uint32_t n_vertices, float* sums) {
    // fewer than K vertices in the radius: frnn pads with -1, which torch's indexing wraps to the last vertex, sqrt(-1) = NaN (:458)
    const uint32_t j = ok ? (uint32_t)idx : n_vertices - 1;
    const float dis = ok ? sqrtf(d2) : NAN;
    float vn[3], dvo[3];
    for (int c = 0; c < 3; c++) { vn[c] = __ldg(vertex_normals + 3 * (size_t)j + c); dvo[c] = x[c] - __ldg(vertices + 3 * (size_t)j + c); }
    const float len = norm3(dvo) + 1e-5f;                 // :461
    const float w = 1.0f / (dis + 1e-7f);                 // :474 and :487 (Shepard): the same weight twice
    const float nl = norm3(vn) + 1e-5f;                   // :497
    for (int c = 0; c < 3; c++) {
        sums[c] += w * (dvo[c] / len);                    // :476
        sums[3 + c] += vn[c];                             // :477
        sums[6 + c] += (vn[c] / nl) * w;                  // :498 before the division by the weight sum
    }
    sums[9] += w;
}

__device__ __forceinline__ void finish_normal(float* sums, int k_want, float dir_vec_wdist, float* n) {
    float* mean_dir = sums;
    float* ntest = sums + 3;
    float* acc = sums + 6;
    float wsum = sums[9];
    for (int c = 0; c < 3; c++) ntest[c] /= (float)k_want;
    if (mean_dir[0] * ntest[0] + mean_dir[1] * ntest[1] + mean_dir[2] * ntest[2] < 0.0f)
        for (int c = 0; c < 3; c++) mean_dir[c] = -mean_dir[c];  // :478
    {
        const float len = norm3(mean_dir) + 1e-5f;                // :479
        for (int c = 0; c < 3; c++) mean_dir[c] /= len;
        const float w = 1.0f / (fmaxf(dir_vec_wdist, 1e-5f) + 1e-7f);   // :481-482, :487
        const float nl = norm3(mean_dir) + 1e-5f;                 // :497 (normalised a second time, as the reference does)
        for (int c = 0; c < 3; c++) acc[c] += (mean_dir[c] / nl) * w;
        wsum += w;
    }
    for (int c = 0; c < 3; c++) acc[c] /= wsum;                   // :496
    const float len = norm3(acc) + 1e-5f;                         // :499
    for (int c = 0; c < 3; c++) n[c] = acc[c] / len;
}

// tools/map.py:421-425 after the two casts of :419-420
__device__ __forceinline__ void project_outputs(size_t i, const float* x, const float* n, const Hit& h1, const Hit& h2, float* p_sur, float* sdf,
                                                float* normal_out, long long* face_idx) {
    const bool cond = h1.t < h2.t;                                 // :421
    const float t = cond ? h1.t : h2.t;
    for (int c = 0; c < 3; c++) {
        p_sur[3 * i + c] = NTX_ADD(x[c], NTX_MUL(t, cond ? n[c] : -n[c]));   // :422
        normal_out[3 * i + c] = n[c];
    }
    sdf[i] = cond ? -h1.t : h2.t;                                  // :423
    face_idx[i] = cond ? h1.face : h2.face;                        // :425
}

// MeshProjector.project, one thread per sample: the neighbour list never leaves the registers, both casts run back to back
template <int K>
__global__ void __launch_bounds__(128) mesh_project_kernel(uint32_t N, const float* xyz, int k_want, float r2, float dir_vec_wdist,
                                                           const float* __restrict__ vertices, const float* __restrict__ vertex_normals,
                                                           uint32_t n_vertices, const Node* __restrict__ pt_nodes, const Point* __restrict__ pts,
                                                           const Node* __restrict__ tri_nodes, const Tri* __restrict__ tris, float slack,
                                                           float* p_sur, float* sdf, float* normal_out, long long* face_idx) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float x[3];
    load3(xyz, i, x);
    float bd[K];
    int bi[K];
    knn_one<K>(pt_nodes, pts, x, r2, k_want, bd, bi);
    float sums[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int s = 0; s < K; s++)
        if (s >= K - k_want) neighbour_terms(x, knn_slot_valid(bi[s]), bd[s], bi[s], vertices, vertex_normals, n_vertices, sums);
    float n[3], nn[3];
    finish_normal(sums, k_want, dir_vec_wdist, n);
    for (int c = 0; c < 3; c++) nn[c] = -n[c];
    const Hit h1 = trace_one(tri_nodes, tris, x, n, slack);        // :419 inner
    const Hit h2 = trace_one(tri_nodes, tris, x, nn, slack);       // :420 outer
    project_outputs(i, x, n, h1, h2, p_sur, sdf, normal_out, face_idx);
}

static MeshHandle* checked(const void* mesh, const char* who) {
    MeshHandle* h = (MeshHandle*)mesh;
    if (!h || h->magic != kMeshMagic) { set_error("%s: not a mesh handle", who); return nullptr; }
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev != h->device) { set_error("%s: the mesh lives on device %d, the current device is %d", who, h->device, dev); return nullptr; }
    return h;
}

}  // namespace mesh
}  // namespace ntx

using namespace ntx;
using namespace ntx::mesh;

extern "C" {

int ntx_mesh_create(const float* vertices, uint32_t n_vertices, const int32_t* triangles, uint32_t n_triangles, void** mesh_out) {
    NTX_REQUIRE(mesh_out != nullptr, NTX_ERR_INVALID_ARGUMENT, "ntx_mesh_create: mesh_out is null");
    *mesh_out = nullptr;
    NTX_REQUIRE(vertices != nullptr && n_vertices > 0, NTX_ERR_INVALID_ARGUMENT, "ntx_mesh_create: no vertices");
    NTX_REQUIRE(n_triangles == 0 || triangles != nullptr, NTX_ERR_INVALID_ARGUMENT, "ntx_mesh_create: triangles is null");
    NTX_REQUIRE(n_vertices < (1u << 27) && n_triangles < (1u << 27), NTX_ERR_UNSUPPORTED, "ntx_mesh_create: more than 2^27 primitives");
    BuildResult tri_tree, pt_tree;
    std::vector<Tri> tris;
    std::vector<Point> pts;
    NTX_REQUIRE(build_triangle_tree(vertices, n_vertices, triangles, n_triangles, tri_tree, tris), NTX_ERR_INVALID_ARGUMENT,
                "ntx_mesh_create: non-finite vertex or triangle index out of range");
    NTX_REQUIRE(build_point_tree(vertices, n_vertices, pt_tree, pts), NTX_ERR_INVALID_ARGUMENT, "ntx_mesh_create: non-finite vertex");
    NTX_REQUIRE(tri_tree.depth < kStackDepth && pt_tree.depth < kStackDepth, NTX_ERR_UNSUPPORTED, "ntx_mesh_create: tree deeper than %d", kStackDepth);
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = 0; i < n_vertices; i++)
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], vertices[3 * (size_t)i + k]); hi[k] = std::max(hi[k], vertices[3 * (size_t)i + k]); }
    MeshHandle* h = new (std::nothrow) MeshHandle();
    NTX_REQUIRE(h != nullptr, NTX_ERR_CUDA, "ntx_mesh_create: out of host memory");
    std::memset(h, 0, sizeof(*h));
    h->magic = kMeshMagic;
    cudaGetDevice(&h->device);
    h->n_vertices = n_vertices; h->n_triangles = n_triangles;
    h->tri_nodes_n = (uint32_t)tri_tree.nodes.size(); h->pt_nodes_n = (uint32_t)pt_tree.nodes.size();
    h->tri_depth = tri_tree.depth; h->pt_depth = pt_tree.depth;
    h->slack = 1e-5f * std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2]));
    std::vector<float> verts(vertices, vertices + 3 * (size_t)n_vertices);
    const bool ok = upload(&h->tri_nodes, tri_tree.nodes) && upload(&h->tris, tris) && upload(&h->pt_nodes, pt_tree.nodes) && upload(&h->pts, pts) &&
                    upload(&h->vertices, verts);
    if (!ok) {
        set_error("ntx_mesh_create: %s", cudaGetErrorString(cudaGetLastError()));
        free_handle(h);
        return NTX_ERR_CUDA;
    }
    *mesh_out = h;
    return NTX_OK;
}

int ntx_mesh_destroy(void* mesh) {
    if (!mesh) return NTX_OK;
    MeshHandle* h = (MeshHandle*)mesh;
    NTX_REQUIRE(h->magic == kMeshMagic, NTX_ERR_INVALID_ARGUMENT, "ntx_mesh_destroy: not a mesh handle");
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(h->device);
    free_handle(h);
    if (prev >= 0) cudaSetDevice(prev);
    return NTX_OK;
}

int ntx_mesh_info(const void* mesh, uint32_t* out6) {
    const MeshHandle* h = (const MeshHandle*)mesh;
    NTX_REQUIRE(h && h->magic == kMeshMagic && out6, NTX_ERR_INVALID_ARGUMENT, "ntx_mesh_info: not a mesh handle");
    out6[0] = h->n_vertices; out6[1] = h->n_triangles; out6[2] = h->tri_nodes_n; out6[3] = h->tri_depth; out6[4] = h->pt_nodes_n; out6[5] = h->pt_depth;
    return NTX_OK;
}

int ntx_mesh_trace(const void* mesh, const float* rays_o, const float* rays_d, float* positions, float* normals, float* depth, int64_t* face_idx,
                   uint32_t N, ntx_stream_t stream) {
    MeshHandle* h = checked(mesh, "ntx_mesh_trace");
    if (!h) return NTX_ERR_INVALID_ARGUMENT;
    if (N == 0) return NTX_OK;
    NTX_REQUIRE(rays_o && rays_d && positions && normals && depth && face_idx, NTX_ERR_INVALID_ARGUMENT, "ntx_mesh_trace: null pointer");
    const uint32_t block = (uint32_t)tunables().mesh_block;
    mesh_trace_kernel<<<ceil_div(N, block), block, 0, (cudaStream_t)stream>>>(N, rays_o, rays_d, positions, normals, depth, (long long*)face_idx,
                                                                          h->tri_nodes, h->tris, h->slack);
    return check_launch("ntx_mesh_trace");
}

int ntx_mesh_knn(const void* mesh, const float* queries, uint32_t N, uint32_t K, float r, float* dists, int64_t* idxs, ntx_stream_t stream) {
    MeshHandle* h = checked(mesh, "ntx_mesh_knn");
    if (!h) return NTX_ERR_INVALID_ARGUMENT;
    NTX_REQUIRE(K >= 1 && K <= 32, NTX_ERR_UNSUPPORTED, "ntx_mesh_knn: K must be in 1..32, got %u", K);
    NTX_REQUIRE(r > 0.0f, NTX_ERR_INVALID_ARGUMENT, "ntx_mesh_knn: r must be positive");
    if (N == 0) return NTX_OK;
    NTX_REQUIRE(queries && dists && idxs, NTX_ERR_INVALID_ARGUMENT, "ntx_mesh_knn: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const float r2 = r * r;
    const uint32_t block = (uint32_t)tunables().mesh_block;
    const dim3 grid(ceil_div(N, block));
    if (K <= 8) {
        mesh_knn_kernel<8><<<grid, block, 0, st>>>(N, queries, r2, (int)K, h->pt_nodes, h->pts, dists, (long long*)idxs);
    } else if (K <= 16) {
        mesh_knn_kernel<16><<<grid, block, 0, st>>>(N, queries, r2, (int)K, h->pt_nodes, h->pts, dists, (long long*)idxs);
    } else {
        mesh_knn_kernel<32><<<grid, block, 0, st>>>(N, queries, r2, (int)K, h->pt_nodes, h->pts, dists, (long long*)idxs);
    }
    return check_launch("ntx_mesh_knn");
}

int ntx_mesh_project(const void* mesh, const float* vertex_normals, const float* xyz, uint32_t N, uint32_t K, float r, float dir_vec_wdist,
                     float* p_sur, float* sdf, float* normal, int64_t* face_idx, ntx_stream_t stream) {
    MeshHandle* h = checked(mesh, "ntx_mesh_project");
    if (!h) return NTX_ERR_INVALID_ARGUMENT;
    NTX_REQUIRE(K >= 1 && K <= 16, NTX_ERR_UNSUPPORTED, "ntx_mesh_project: K must be in 1..16, got %u", K);
    NTX_REQUIRE(r > 0.0f, NTX_ERR_INVALID_ARGUMENT, "ntx_mesh_project: r must be positive");
    NTX_REQUIRE(h->n_triangles > 0, NTX_ERR_INVALID_ARGUMENT, "ntx_mesh_project: the mesh has no triangles");
    if (N == 0) return NTX_OK;
    NTX_REQUIRE(vertex_normals && xyz && p_sur && sdf && normal && face_idx, NTX_ERR_INVALID_ARGUMENT, "ntx_mesh_project: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const float r2 = r * r;
    const uint32_t block = (uint32_t)tunables().mesh_block;
    const dim3 grid(ceil_div(N, block));
    if (K <= 8) {
        mesh_project_kernel<8><<<grid, block, 0, st>>>(N, xyz, (int)K, r2, dir_vec_wdist, h->vertices, vertex_normals, h->n_vertices, h->pt_nodes, h->pts,
                                                     h->tri_nodes, h->tris, h->slack, p_sur, sdf, normal, (long long*)face_idx);
    } else {
        mesh_project_kernel<16><<<grid, block, 0, st>>>(N, xyz, (int)K, r2, dir_vec_wdist, h->vertices, vertex_normals, h->n_vertices, h->pt_nodes, h->pts,
                                                      h->tri_nodes, h->tris, h->slack, p_sur, sdf, normal, (long long*)face_idx);
    }
    return check_launch("ntx_mesh_project");
}

}  // extern "C"
