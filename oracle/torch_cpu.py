"""oracle/torch_cpu.py -- TEST / BASELINE INFRASTRUCTURE.  A pure-PyTorch CPU implementation of the nine
pointnet2 ops behind the reference's `*_wrapper` names (BASELINE.json north_star: "a pure-PyTorch CPU
implementation of the same ops, timed on the box's host cores").  Vectorised torch, no custom kernels:
an FPS loop with running min + argmax, masked first-k ball query, topk(3) three_nn, torch.gather.

It is a throughput baseline, NOT a bit-exactness reference: torch's argmax / topk break ties by lowest
index and its distance arithmetic does not reproduce the reference's FMA contraction.  The C restatement
(oracle/pointnet2_oracle.c) is the exactness oracle.
"""
import torch


def _sqdist(a, b):
    # (B,M,1,3) - (B,1,N,3) -> (B,M,N)
    d = a.unsqueeze(2) - b.unsqueeze(1)
    return (d * d).sum(-1)


def furthest_point_sampling_wrapper(b, n, m, xyz, temp, idx):
    last = torch.zeros(b, dtype=torch.long)
    idx[:, 0] = 0
    ar = torch.arange(b)
    for j in range(1, m):
        p = xyz[ar, last].unsqueeze(1)  # (B,1,3)
        d = ((xyz - p) ** 2).sum(-1)
        torch.minimum(temp, d, out=temp)
        last = temp.argmax(dim=1)
        idx[:, j] = last.int()
    return 1


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    out.copy_(torch.gather(points, 2, idx.long().unsqueeze(1).expand(b, c, npoints)))
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    grad_points.scatter_add_(2, idx.long().unsqueeze(1).expand(b, c, npoints), grad_out)
    return 1


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx, chunk=1024):
    r2 = float(radius) * float(radius)
    ar = torch.arange(n, dtype=torch.int32).view(1, 1, n)
    for s in range(0, m, chunk):
        d2 = _sqdist(new_xyz[:, s:s + chunk], xyz)  # (B,chunk,N)
        key = torch.where(d2 < r2, ar, torch.full_like(ar, n))  # misses sort last
        k = min(nsample, n)
        first = torch.topk(key, k, dim=2, largest=False, sorted=True).values  # ascending hit indices
        if k < nsample:
            first = torch.cat([first, first.new_full((b, first.shape[1], nsample - k), n)], dim=2)
        head = first[:, :, :1]
        first = torch.where(first == n, head.expand_as(first), first)  # pad with the first hit
        first = torch.where(first == n, torch.zeros_like(first), first)  # no hit at all -> 0
        idx[:, s:s + chunk] = first
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    flat = idx.long().view(b, 1, npoints * nsample).expand(b, c, npoints * nsample)
    out.copy_(torch.gather(points, 2, flat).view(b, c, npoints, nsample))
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    flat = idx.long().view(b, 1, npoints * nsample).expand(b, c, npoints * nsample)
    grad_points.scatter_add_(2, flat, grad_out.reshape(b, c, npoints * nsample))
    return 1


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx, chunk=2048):
    for s in range(0, n, chunk):
        d2 = _sqdist(unknown[:, s:s + chunk], known)
        v, i = torch.topk(d2, 3, dim=2, largest=False, sorted=True)
        dist2[:, s:s + chunk] = v
        idx[:, s:s + chunk] = i.int()


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    flat = idx.long().view(b, 1, n * 3).expand(b, c, n * 3)
    g = torch.gather(points, 2, flat).view(b, c, n, 3)
    out.copy_((g * weight.unsqueeze(1)).sum(-1))


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    flat = idx.long().view(b, 1, n * 3).expand(b, c, n * 3)
    contrib = (grad_out.unsqueeze(-1) * weight.unsqueeze(1)).reshape(b, c, n * 3)
    grad_points.scatter_add_(2, flat, contrib)
