"""Synthetic "KITTI-shaped" scenes for tests and benchmarks: there is no dataset in
this environment, so inputs follow the reference's input contract instead -- rect-camera coordinates
inside PC_AREA_SCOPE x in [-40,40], y in [-1,3], z in [0,70.4] (lib/config.py:26-28), 16384 points
(lib/config.py:68), short clouds padded by duplicating points (lib/datasets/kitti_rcnn_dataset.py:339-341),
pixel coordinates inside the 375x1242 valid area of the zero-padded 384x1280 canvas
(lib/datasets/kitti_dataset.py:54-55).  Everything is generated on the CPU from a seeded torch.Generator.
"""
import math

import torch

SCOPE = ((-40.0, 40.0), (-1.0, 3.0), (0.0, 70.4))
IMG_H, IMG_W = 384, 1280
VALID_H, VALID_W = 375, 1242


def lidar_scene(seed: int, n: int = 16384, dup_frac: float = 0.02) -> torch.Tensor:
    """One (n,3) fp32 cloud: range pdf ~ 1/rho on [2.5,70] m, azimuth +-40 deg, 85 % ground plane at
    y = 1.65 +- 0.05, 15 % on ~20 car-sized boxes (1.5 x 1.6 x 3.9 m), the last `dup_frac` of the points are
    exact copies of earlier ones."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(n, generator=g)
    rho = 2.5 * (70.0 / 2.5) ** u
    theta = (torch.rand(n, generator=g) * 80.0 - 40.0) * (math.pi / 180.0)
    x, z = rho * torch.sin(theta), rho * torch.cos(theta)
    y = 1.65 + 0.05 * torch.randn(n, generator=g)
    n_obj = int(0.15 * n)
    centres_rho = 5.0 + 55.0 * torch.rand(20, generator=g)
    centres_th = (torch.rand(20, generator=g) * 70.0 - 35.0) * (math.pi / 180.0)
    which = torch.randint(0, 20, (n_obj,), generator=g)
    box = (torch.rand(n_obj, 3, generator=g) - 0.5) * torch.tensor([1.6, 1.5, 3.9])
    x[:n_obj] = centres_rho[which] * torch.sin(centres_th[which]) + box[:, 0]
    y[:n_obj] = 0.9 + box[:, 1]
    z[:n_obj] = centres_rho[which] * torch.cos(centres_th[which]) + box[:, 2]
    pts = torch.stack([x, y, z], dim=1)
    for a, (lo, hi) in enumerate(SCOPE):
        pts[:, a].clamp_(lo, hi)
    pts = pts[torch.randperm(n, generator=g)]
    n_dup = int(dup_frac * n)
    if n_dup:
        src = torch.randint(0, n - n_dup, (n_dup,), generator=g)
        pts[n - n_dup:] = pts[src]
    return pts.float().contiguous()


def uniform_scene(seed: int, n: int = 16384) -> torch.Tensor:
    """Stress variant: uniform in the scope box, so small balls are mostly empty (no early exit)."""
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(n, 3, generator=g)
    for a, (lo, hi) in enumerate(SCOPE):
        pts[:, a] = lo + (hi - lo) * pts[:, a]
    return pts.float().contiguous()


MEAN = (0.485, 0.456, 0.406)  # lib/datasets/kitti_dataset.py:24-25
STD = (0.229, 0.224, 0.225)


def host_image_prep(image_u8: torch.Tensor, hw=(IMG_H, IMG_W)) -> torch.Tensor:
    """What the reference's data loader does on the host (lib/datasets/kitti_dataset.py:44-55, then the cast and permute of
    lib/net/train_functions.py:37): uint8 (B,h,w,3) -> float64 /255, -mean, /std, zero-padded canvas -> fp32 (B,3,H,W)."""
    im = image_u8.to(torch.float64) / 255.0
    im = im - torch.tensor(MEAN, dtype=torch.float64)
    im = im / torch.tensor(STD, dtype=torch.float64)
    canvas = torch.zeros(image_u8.shape[0], hw[0], hw[1], 3, dtype=torch.float64)
    canvas[:, :image_u8.shape[1], :image_u8.shape[2]] = im
    return canvas.float().permute(0, 3, 1, 2).contiguous()


def batch(first_seed: int, b: int, n: int = 16384, kind: str = "lidar", with_u8: bool = False):
    """-> dict(points (B,n,3), image (B,3,384,1280), xy (B,n,2) pixel coords); scene i uses seed first_seed+i.
    The image is a random decoded camera frame (uint8, 375x1242) put through the reference's host-side preparation;
    with_u8=True also returns the frame itself as image_u8 (B,375,1242,3) for the device-side preparation."""
    make = lidar_scene if kind == "lidar" else uniform_scene
    pts = torch.stack([make(first_seed + i, n) for i in range(b)])
    g = torch.Generator().manual_seed(first_seed + 7919)
    xy = torch.stack([torch.rand(b, n, generator=g) * (VALID_W - 1), torch.rand(b, n, generator=g) * (VALID_H - 1)], dim=2)
    image_u8 = torch.randint(0, 256, (b, VALID_H, VALID_W, 3), generator=g, dtype=torch.uint8)
    out = {"points": pts, "image": host_image_prep(image_u8), "xy": xy.float().contiguous()}
    if with_u8:
        out["image_u8"] = image_u8.contiguous()
    return out
