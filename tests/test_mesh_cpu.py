"""CPU tests of the mesh front end: the oracle against known answers, and the product's tree builder + traversal
logic (compiled for the host by tests/native/mesh_host_check.cpp) against the oracle's exhaustive scans — bit for bit."""
import ctypes as C

import numpy as np
import pytest

import _util as U


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def host():
    return U.build_mesh_host_check()


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.int32)


# ------------------------------------------------------------------------------------------------ oracle: known answers
def test_oracle_trace_known_answers():
    """unit square in z = 0 made of two triangles: straight-down rays hit at t = height exactly (all quantities are dyadic)"""
    O = U.oracle()
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], np.float32)
    f = np.array([[0, 1, 2], [1, 3, 2]], np.int32)
    o = np.array([[0.25, 0.25, 1.0], [0.75, 0.75, 2.0], [0.25, 0.25, -0.5], [2.0, 2.0, 1.0], [0.25, 0.25, 12.0], [0.5, 0.5, 1.0]], np.float32)
    d = np.array([[0, 0, -1], [0, 0, -1], [0, 0, 1], [0, 0, -1], [0, 0, -1], [0, 0, -1]], np.float32)
    pos, nrm, depth, face = O.mesh_trace(v, f, o, d)
    assert depth.tolist() == [1.0, 2.0, 0.5, 10.0, 10.0, 1.0]          # miss -> MAX_DIST (bvh.cu:36,263); a hit beyond 10 is a miss
    assert face.tolist() == [0, 1, 0, -1, -1, 0]                        # the diagonal belongs to both: the first in index order wins
    assert pos[0].tolist() == [0.25, 0.25, 0.0] and pos[3].tolist() == [2.0, 2.0, -9.0]   # position = o + depth * d, also for a miss (bvh.cu:709)
    assert nrm[0].tolist() == [0.0, 0.0, 1.0] and nrm[3].tolist() == [0.0, 0.0, 0.0]
    # a ray pointing away from the plane and one parallel to it hit nothing (t < 0 resp. 1/0)
    _, _, depth, face = O.mesh_trace(v, f, np.array([[0.25, 0.25, 1], [0.25, 0.25, 1]], np.float32), np.array([[0, 0, 1], [1, 0, 0]], np.float32))
    assert depth.tolist() == [10.0, 10.0] and face.tolist() == [-1, -1]


def test_oracle_knn_known_answers():
    O = U.oracle()
    pts = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [4, 0, 0], [1, 0, 0]], np.float32)
    d, i = O.points_knn(pts, np.array([[0.75, 0, 0]], np.float32), 4, 100.0)
    assert i[0].tolist() == [1, 4, 0, 2] and d[0].tolist() == [0.0625, 0.0625, 0.5625, 1.5625]   # squared, ascending, equal distance: lower index first
    d, i = O.points_knn(pts, np.array([[0.75, 0, 0]], np.float32), 4, 1.0)
    assert i[0].tolist() == [1, 4, 0, -1] and d[0, 3] == -1.0                                      # strictly inside the radius, -1 padding


def test_oracle_project_on_a_sphere():
    """on a finely tessellated sphere the projection is radial: sdf = |x| - R (negative inside), p_sur = R x/|x|, normal = x/|x|"""
    O = U.oracle()
    v, f, vn = U.bumpy_sphere(60, 90, bump=0.0, radius=0.7)
    rng = np.random.default_rng(3)
    x = rng.normal(size=(400, 3))
    x = (x / np.linalg.norm(x, axis=1, keepdims=True) * rng.uniform(0.5, 0.9, (400, 1))).astype(np.float32)
    x = x[np.abs(x[:, 2]) / np.linalg.norm(x, axis=1) < 0.9]            # stay away from the open poles
    p_sur, sdf, normal, face = O.mesh_project(v, vn, f, x)
    rad = np.linalg.norm(x, axis=1)
    assert (face >= 0).all()
    assert np.abs(sdf - (rad - 0.7)).max() < 5e-3
    nerr = np.abs(normal - x / rad[:, None]).max(1)                    # samples close to the surface see their neighbours edge-on: noisy
    assert np.median(nerr) < 0.02 and nerr.max() < 0.2
    assert np.abs(np.linalg.norm(p_sur, axis=1) - 0.7).max() < 5e-3


# ------------------------------------------------------------------------------------------------ host logic: trees and traversals
@pytest.mark.parametrize("shape", [(2, 3), (6, 8), (20, 30), (80, 120)])
def test_trees_cover_every_primitive_once(host, shape):
    v, f, _ = U.bumpy_sphere(*shape)
    for points, n in ((0, len(f)), (1, len(v))):
        st = np.zeros(3, np.int64)
        depth = host.hostcheck_tree(_p(v), len(v), _p(f), len(f), points, _p(st))
        assert 1 <= depth < 64
        assert st[1] == n and st[2] == 0, "every primitive in exactly one leaf, inside its leaf's box"
        assert st[0] >= 1


def test_tree_rejects_bad_input(host):
    v, f, _ = U.bumpy_sphere(6, 8)
    st = np.zeros(3, np.int64)
    bad = f.copy(); bad[3, 1] = len(v)
    assert host.hostcheck_tree(_p(v), len(v), _p(bad), len(bad), 0, _p(st)) < 0
    vb = v.copy(); vb[5, 2] = np.nan
    assert host.hostcheck_tree(_p(vb), len(vb), _p(f), len(f), 0, _p(st)) < 0


@pytest.mark.parametrize("shape,n_rays", [((20, 30), 60000), ((80, 120), 30000)])
def test_host_trace_equals_exhaustive_scan(host, shape, n_rays):
    O = U.oracle()
    v, f, _ = U.bumpy_sphere(*shape)
    rng = np.random.default_rng(shape[0])
    o, d = U.random_rays(rng, n_rays)
    _, _, depth, face = O.mesh_trace(v, f, o, d)
    hd, hf = np.empty(n_rays, np.float32), np.empty(n_rays, np.int64)
    assert host.hostcheck_trace(_p(v), len(v), _p(f), len(f), _p(o), _p(d), n_rays, C.c_float(1e-5), _p(hd), _p(hf)) > 0
    assert (face >= 0).mean() > 0.2
    assert np.array_equal(_bits(hd), _bits(depth)) and np.array_equal(hf, face)


def test_host_trace_ties_degenerates_and_in_plane_rays(host):
    O = U.oracle()
    v, f, _ = U.bumpy_sphere(20, 30, bump=0.0)
    rng = np.random.default_rng(1)
    o, d = U.adversarial_rays(rng, v, f)
    f2 = np.concatenate([f, f[:50], np.array([[0, 0, 1], [2, 2, 2]], np.int32)]).astype(np.int32)   # duplicated and degenerate triangles
    _, _, depth, face = O.mesh_trace(v, f2, o, d)
    n = len(o)
    hd, hf = np.empty(n, np.float32), np.empty(n, np.int64)
    for slack in (0.0, 1e-5):
        host.hostcheck_trace(_p(v), len(v), _p(f2), len(f2), _p(o), _p(d), n, C.c_float(slack), _p(hd), _p(hf))
        assert np.array_equal(_bits(hd), _bits(depth)) and np.array_equal(hf, face)


@pytest.mark.parametrize("K,r", [(8, 100.0), (5, 100.0), (8, 0.3), (1, 100.0), (32, 100.0)])
def test_host_knn_equals_exhaustive_scan(host, K, r):
    O = U.oracle()
    v, _, _ = U.bumpy_sphere(40, 60)
    rng = np.random.default_rng(K)
    q = rng.uniform(-1, 1, (20000, 3)).astype(np.float32)
    od, oi = O.points_knn(v, q, K, r)
    hd, hi = np.empty((len(q), K), np.float32), np.empty((len(q), K), np.int64)
    assert host.hostcheck_knn(_p(v), len(v), _p(q), len(q), K, C.c_float(r), _p(hd), _p(hi)) > 0
    assert np.array_equal(_bits(hd), _bits(od)) and np.array_equal(hi, oi)


def test_host_knn_ties_duplicates_and_tiny_clouds(host):
    O = U.oracle()
    rng = np.random.default_rng(2)
    pts = np.round(rng.uniform(-1, 1, (3000, 3)) * 8) / 8
    pts = np.concatenate([pts, pts[:500]]).astype(np.float32)          # lattice points with exact duplicates
    q = (np.round(rng.uniform(-1, 1, (8000, 3)) * 16) / 16).astype(np.float32)
    for cloud in (pts, pts[:1], pts[:3], pts[:9]):
        od, oi = O.points_knn(cloud, q, 8, 100.0)
        hd, hi = np.empty((len(q), 8), np.float32), np.empty((len(q), 8), np.int64)
        host.hostcheck_knn(_p(cloud), len(cloud), _p(q), len(q), 8, C.c_float(100.0), _p(hd), _p(hi))
        assert np.array_equal(_bits(hd), _bits(od)) and np.array_equal(hi, oi)


def test_mesh_create_validates_before_touching_the_gpu():
    L = U.ntx()
    lib = L.lib()
    v, f, _ = U.bumpy_sphere(6, 8)
    h = C.c_void_p()
    bad = f.copy(); bad[0, 0] = -1
    assert lib.ntx_mesh_create(_p(v), len(v), _p(bad), len(bad), C.byref(h)) == -1 and not h.value
    assert b"out of range" in lib.ntx_last_error()
    vb = v.copy(); vb[0, 0] = np.inf
    assert lib.ntx_mesh_create(_p(vb), len(vb), _p(f), len(f), C.byref(h)) == -1
    assert lib.ntx_mesh_create(None, 0, None, 0, C.byref(h)) == -1
    assert lib.ntx_mesh_trace(None, None, None, None, None, None, None, 4, None) == -1      # not a handle
    assert lib.ntx_mesh_destroy(None) == 0


def test_trees_stay_shallow_and_exact_on_hostile_vertex_distributions(host):
    """geometric progressions (every split isolates one primitive), a single line, all-equal points, huge coordinate ranges: the builder
    must stay below the traversal stack (64) — ntx_mesh_create refuses deeper trees — and the searches must stay exact"""
    O = U.oracle()
    rng = np.random.default_rng(5)
    n = 3000
    geo = np.stack([2.0 ** -np.arange(n, dtype=np.float64) * 1e3, np.zeros(n), np.zeros(n)], 1)
    geo[1500:, 0] = 0.0                                                   # half of them underflow to the same point
    line = np.stack([np.linspace(-1, 1, n), np.zeros(n), np.zeros(n)], 1)
    same = np.ones((n, 3)) * 0.25
    wide = rng.uniform(-1, 1, (n, 3)) * np.array([1e6, 1e-6, 1.0])
    q = rng.uniform(-1, 1, (3000, 3)).astype(np.float32)
    for cloud in (geo, line, same, wide):
        v = cloud.astype(np.float32)
        f = np.stack([np.arange(n - 2), np.arange(1, n - 1), np.arange(2, n)], 1).astype(np.int32)   # a strip of (mostly degenerate) triangles
        for points in (0, 1):
            st = np.zeros(3, np.int64)
            depth = host.hostcheck_tree(_p(v), n, _p(f), len(f), points, _p(st))
            assert 1 <= depth < 64, depth
            assert st[1] == (n if points else len(f)) and st[2] == 0
        od, oi = O.points_knn(v, q, 8, 100.0)
        hd, hi = np.empty((len(q), 8), np.float32), np.empty((len(q), 8), np.int64)
        host.hostcheck_knn(_p(v), n, _p(q), len(q), 8, C.c_float(100.0), _p(hd), _p(hi))
        assert np.array_equal(_bits(hd), _bits(od)) and np.array_equal(hi, oi)
        o, d = U.random_rays(rng, 3000)
        _, _, depth_o, face_o = O.mesh_trace(v, f, o, d)
        gd, gf = np.empty(len(o), np.float32), np.empty(len(o), np.int64)
        host.hostcheck_trace(_p(v), n, _p(f), len(f), _p(o), _p(d), len(o), C.c_float(1e-5), _p(gd), _p(gf))
        assert np.array_equal(_bits(gd), _bits(depth_o)) and np.array_equal(gf, face_o)


def test_property_random_small_meshes_and_clouds(host):
    """hypothesis: any small mesh / cloud (coordinates on a coarse lattice, so coincident vertices, coplanar and degenerate triangles, rays
    through vertices and exact distance ties are common), any K and radius — tree search == exhaustive scan, bit for bit"""
    from hypothesis import given, settings, strategies as st
    O = U.oracle()

    @settings(max_examples=60, deadline=None)
    @given(seed=st.integers(0, 2 ** 31 - 1), n_v=st.integers(3, 40), n_t=st.integers(1, 60), K=st.integers(1, 8), lattice=st.sampled_from([2, 4, 64]),
           r=st.sampled_from([0.25, 0.5, 1.0, 100.0]))
    def check(seed, n_v, n_t, K, lattice, r):
        rng = np.random.default_rng(seed)
        v = (np.round(rng.uniform(-1, 1, (n_v, 3)) * lattice) / lattice).astype(np.float32)
        f = rng.integers(0, n_v, (n_t, 3)).astype(np.int32)
        o = (np.round(rng.uniform(-1.5, 1.5, (300, 3)) * lattice) / lattice).astype(np.float32)
        tgt = v[rng.integers(0, n_v, 300)] + (rng.uniform(-0.1, 0.1, (300, 3)) * (rng.random((300, 1)) < 0.5)).astype(np.float32)
        d = (tgt - o).astype(np.float32)
        d[:50] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 50)]
        _, _, depth_o, face_o = O.mesh_trace(v, f, o, d)
        gd, gf = np.empty(300, np.float32), np.empty(300, np.int64)
        assert host.hostcheck_trace(_p(v), n_v, _p(f), n_t, _p(o), _p(d), 300, C.c_float(1e-5), _p(gd), _p(gf)) > 0
        assert np.array_equal(_bits(gd), _bits(depth_o)) and np.array_equal(gf, face_o)
        od, oi = O.points_knn(v, o, K, r)
        hd, hi = np.empty((300, K), np.float32), np.empty((300, K), np.int64)
        assert host.hostcheck_knn(_p(v), n_v, _p(o), 300, K, C.c_float(r), _p(hd), _p(hi)) > 0
        assert np.array_equal(_bits(hd), _bits(od)) and np.array_equal(hi, oi)

    check()
