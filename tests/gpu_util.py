"""Helpers for the -m gpu tests: run an op set (ours or the reference's kernels) on numpy inputs."""
import numpy as np
import torch


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


class OpRunner:
    """numpy-in / numpy-out front end over a backend exposing the reference's nine `*_wrapper` functions."""

    def __init__(self, backend):
        self.be = backend

    def fps(self, xyz, m, temp=None, return_temp=False):
        x = dev(xyz)
        B, N, _ = x.shape
        t = torch.full((B, N), 1e10, device="cuda") if temp is None else dev(temp)
        idx = torch.zeros((B, m), dtype=torch.int32, device="cuda")
        self.be.furthest_point_sampling_wrapper(B, N, m, x, t, idx)
        torch.cuda.synchronize()
        return (idx.cpu().numpy(), t.cpu().numpy()) if return_temp else idx.cpu().numpy()

    def ball_query(self, radius, nsample, xyz, new_xyz):
        x, q = dev(xyz), dev(new_xyz)
        B, N, _ = x.shape
        M = q.shape[1]
        idx = torch.zeros((B, M, nsample), dtype=torch.int32, device="cuda")
        self.be.ball_query_wrapper(B, N, M, radius, nsample, q, x, idx)
        torch.cuda.synchronize()
        return idx.cpu().numpy()

    def gather(self, points, idx):
        p, i = dev(points), dev(idx)
        B, C, N = p.shape
        M = i.shape[1]
        out = torch.empty((B, C, M), device="cuda")
        self.be.gather_points_wrapper(B, C, N, M, p, i, out)
        return out.cpu().numpy()

    def gather_grad(self, grad_out, idx, n):
        g, i = dev(grad_out), dev(idx)
        B, C, M = g.shape
        out = torch.zeros((B, C, n), device="cuda")
        self.be.gather_points_grad_wrapper(B, C, n, M, g, i, out)
        return out.cpu().numpy()

    def group(self, points, idx):
        p, i = dev(points), dev(idx)
        B, C, N = p.shape
        _, M, ns = i.shape
        out = torch.empty((B, C, M, ns), device="cuda")
        self.be.group_points_wrapper(B, C, N, M, ns, p, i, out)
        return out.cpu().numpy()

    def group_grad(self, grad_out, idx, n):
        g, i = dev(grad_out), dev(idx)
        B, C, M, ns = g.shape
        out = torch.zeros((B, C, n), device="cuda")
        self.be.group_points_grad_wrapper(B, C, n, M, ns, g, i, out)
        return out.cpu().numpy()

    def three_nn(self, unknown, known):
        u, k = dev(unknown), dev(known)
        B, n, _ = u.shape
        m = k.shape[1]
        d2 = torch.empty((B, n, 3), device="cuda")
        idx = torch.empty((B, n, 3), dtype=torch.int32, device="cuda")
        self.be.three_nn_wrapper(B, n, m, u, k, d2, idx)
        return d2.cpu().numpy(), idx.cpu().numpy()

    def three_interpolate(self, points, idx, weight):
        p, i, w = dev(points), dev(idx), dev(weight)
        B, C, m = p.shape
        n = i.shape[1]
        out = torch.empty((B, C, n), device="cuda")
        self.be.three_interpolate_wrapper(B, C, m, n, p, i, w, out)
        return out.cpu().numpy()

    def three_interpolate_grad(self, grad_out, idx, weight, m):
        g, i, w = dev(grad_out), dev(idx), dev(weight)
        B, C, n = g.shape
        out = torch.zeros((B, C, m), device="cuda")
        self.be.three_interpolate_grad_wrapper(B, C, n, m, g, i, w, out)
        return out.cpu().numpy()
