"""Drop-in `gridencoder.grid_clustering` (reference: gridencoder/grid_clustering.py:93-208).

`GridEncoder_clustering` is the hash-grid encoder plus one DEC-style soft-clustering head per level whose KL loss is a
regulariser on the table rows.  The encode path is the same libntx kernel as `GridEncoder`; the clustering loss is a
few small dense torch ops on one level's rows and is not a hot-path item, so it stays in torch.
"""
import numpy as np
import torch
import torch.nn as nn

from .grid import GridEncoder, grid_encode  # noqa: F401  (tools/map.py imports grid_encode from here too)


class ClusteringLayer(nn.Module):
    """Student-t soft assignment of rows to `n_clusters` centres (alpha = degrees of freedom)."""

    def __init__(self, n_clusters=4, hidden=2, cluster_centers=None, alpha=1.0):
        super().__init__()
        self.n_clusters, self.alpha, self.hidden = n_clusters, alpha, hidden
        if cluster_centers is None:
            dev = "cuda" if torch.cuda.is_available() else "cpu"
            cluster_centers = torch.empty(n_clusters, hidden, dtype=torch.float, device=dev).uniform_(-1e-4, 1e-4)
        self.cluster_centers = nn.Parameter(cluster_centers)
        self.kl_loss = nn.KLDivLoss(reduction="mean")

    def forward(self, x):
        d2 = ((x.unsqueeze(1) - self.cluster_centers) ** 2).sum(2)          # [N, K]
        q = (1.0 / (1.0 + d2 / self.alpha)) ** (float(self.alpha + 1) / 2)
        return q / q.sum(dim=1, keepdim=True)

    @staticmethod
    def _target(q):
        p = (q ** 2) / q.sum(0)
        return (p / p.sum(dim=1, keepdim=True)).detach()

    def clustering_loss(self, x):
        q = self(x)
        return self.kl_loss(q.log(), self._target(q))


class GridEncoder_clustering(GridEncoder):
    def __init__(self, input_dim=3, num_levels=4, level_dim=2, per_level_scale=2, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, gridtype="hash", align_corners=False):
        super().__init__(input_dim, num_levels, level_dim, per_level_scale, base_resolution, log2_hashmap_size, desired_resolution,
                         gridtype, align_corners)
        self.cluster_layers = nn.ModuleList([ClusteringLayer() for _ in range(num_levels)])
        self.kl_loss = nn.KLDivLoss(reduction="mean")

    def clustering_loss(self, pick_level=True):
        levels = np.random.choice(np.arange(self.num_levels), [1]) if pick_level else np.arange(self.num_levels)
        loss = 0.0
        for i in levels:
            rows = self.embeddings[int(self.offsets[i]): int(self.offsets[i + 1])]
            q = self.cluster_layers[i](rows)
            loss = loss + self.kl_loss(q.log(), ClusteringLayer._target(q))
        return loss
