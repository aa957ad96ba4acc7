"""Drop-in for the one entry point of `frnn` (github.com/lxxue/FRNN, an un-vendored dependency of the reference, readme.md:37) that the
reference calls: `frnn.frnn_grid_points` at tools/map.py:396 and :456.

    dists, idxs, nn, grid = frnn.frnn_grid_points(points1, points2, lengths1, lengths2, K, r, grid=None, return_nn=False, return_sorted=True)

points1 [B,P1,3] queries, points2 [B,P2,3] the searched cloud (the mesh vertices) -> squared distances [B,P1,K] ascending, indices
[B,P1,K] into points2 (-1 padding, distances -1), the gathered neighbours [B,P1,K,3] if return_nn, and a `grid` object to pass back in
so that the search structure over points2 is built once (here: the vertex tree of libntx's mesh handle instead of frnn's uniform grid).
PARITY UNPINNED: frnn itself is not in the tree; this follows its documented contract (exact K nearest within r).  Results are always
sorted (return_sorted=False only permits any order)."""
import torch

from nerf_texture_b200.mesh import Mesh

__all__ = ["frnn_grid_points"]


class _Grid:
    """What frnn returns as `grid`: one search structure per batch element of points2."""

    def __init__(self, points2, lengths2):
        self.meshes = []
        for b in range(points2.shape[0]):
            n = points2.shape[1] if lengths2 is None else int(lengths2[b])
            self.meshes.append(Mesh(points2[b, :n], None, device=points2.device))
        self.shape = tuple(points2.shape)


def frnn_grid_points(points1, points2, lengths1=None, lengths2=None, K=-1, r=-1, grid=None, return_nn=False, return_sorted=True,
                     radius_cell_ratio=2.0, filename=None):
    if points1.dim() != 3 or points2.dim() != 3 or points1.shape[0] != points2.shape[0] or points1.shape[2] != 3 or points2.shape[2] != 3:
        raise ValueError("points1 and points2 must be [B, P, 3] with the same batch size")
    if not points1.is_cuda or not points2.is_cuda:
        raise TypeError("frnn_grid_points: CUDA tensors only (there is no CPU path)")
    if K < 1 or K > 32:
        raise ValueError("K must be in 1..32")
    if torch.is_tensor(r):
        if r.numel() != 1 and not bool((r == r.flatten()[0]).all()):
            raise ValueError("one radius for all batch elements")
        r = float(r.flatten()[0])
    if grid is None:
        grid = _Grid(points2.detach(), lengths2)
    elif not isinstance(grid, _Grid) or grid.shape[0] != points2.shape[0]:
        raise ValueError("grid was not returned by frnn_grid_points for this batch")
    B, P1 = points1.shape[0], points1.shape[1]
    dists = torch.full((B, P1, K), -1.0, dtype=torch.float32, device=points1.device)
    idxs = torch.full((B, P1, K), -1, dtype=torch.int64, device=points1.device)
    for b in range(B):
        n = P1 if lengths1 is None else int(lengths1[b])
        if n:
            d, i = grid.meshes[b].knn(points1[b, :n].detach(), K=K, r=r)
            dists[b, :n], idxs[b, :n] = d, i
    nn = None
    if return_nn:
        nn = torch.stack([points2[b][idxs[b].clamp(min=0)] for b in range(B)])
        nn = torch.where((idxs >= 0).unsqueeze(-1), nn, torch.zeros_like(nn))
    return dists, idxs, nn, grid
