"""Import-time stand-in for the simple-knn extension (scene/gaussian_model.py:20 imports it unconditionally;
it is only *called* when a scene is created from a raw point cloud, which the prune / distill / render
scripts never do)."""
import torch


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    """Mean squared distance to the 3 nearest neighbours, brute force in chunks (init-time only)."""
    P = points.shape[0]
    out = torch.empty(P, device=points.device, dtype=points.dtype)
    step = max(1, min(P, (1 << 26) // max(P, 1)))
    for s in range(0, P, step):
        d = torch.cdist(points[s:s + step], points)
        out[s:s + step] = (d.topk(4, dim=1, largest=False).values[:, 1:] ** 2).mean(dim=1)
    return out
