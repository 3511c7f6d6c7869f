"""Pose and ray-generation primitives (reference: model/common.py:12-39,175-215,255-308,
model/poses_retriever.py:6-32, model/training.py:474-487) on the fused CUDA kernels."""
import torch
import torch.nn as nn

from . import _lib as L

__all__ = ["make_c2w", "Exp", "vec2skew", "convert3x4_4x4", "PoseRetriever", "arange_pixels",
           "get_world_cameraOrigin_cameraRay", "pixels_from_indices"]


class _PoseFn(torch.autograd.Function):
    """c2w = [Exp(r) | t; 0 0 0 1] @ init   (cope_pose_fwd / cope_pose_bwd)"""

    @staticmethod
    def forward(ctx, r, t, init):
        r, t, init = r.contiguous(), t.contiguous(), init.contiguous()
        out = torch.empty(4, 4, dtype=torch.float32, device=r.device)
        L.call("cope_pose_fwd", L.ptr(r), L.ptr(t), L.ptr(init), L.ptr(out), L.stream())
        ctx.save_for_backward(r, t, init)
        return out

    @staticmethod
    def backward(ctx, d_c2w):
        r, t, init = ctx.saved_tensors
        dr, dt = torch.empty_like(r), torch.empty_like(t)
        L.call("cope_pose_bwd", L.ptr(r), L.ptr(t), L.ptr(init), L.ptr(d_c2w.contiguous()), L.ptr(dr), L.ptr(dt),
               L.stream())
        return dr, dt, None


def make_c2w(r, t):
    """model/common.py:279-288 — so(3) exp + raw translation."""
    return _PoseFn.apply(r, t, torch.eye(4, dtype=torch.float32, device=r.device))


def Exp(r):
    """model/common.py:268-277."""
    return make_c2w(r, torch.zeros(3, dtype=torch.float32, device=r.device))[:3, :3]


def vec2skew(v):
    """model/common.py:255-265."""
    z = torch.zeros(1, dtype=torch.float32, device=v.device)
    return torch.stack([torch.cat([z, -v[2:3], v[1:2]]), torch.cat([v[2:3], z, -v[0:1]]),
                        torch.cat([-v[1:2], v[0:1], z])], dim=0)


def convert3x4_4x4(inp):
    """model/common.py:290-308 (tensor branch)."""
    if inp.dim() == 3:
        out = torch.cat([inp, torch.zeros_like(inp[:, 0:1])], dim=1)
        out[:, 3, 3] = 1.0
        return out
    return torch.cat([inp, torch.tensor([[0, 0, 0, 1]], dtype=inp.dtype, device=inp.device)], dim=0)


class PoseRetriever(nn.Module):
    """model/poses_retriever.py:6-32 — per-camera learnable (r, t) composed with a fixed init_c2w."""

    def __init__(self, num_cams, learn_R=True, learn_t=True, init_c2w=None):
        super().__init__()
        self.num_cams = num_cams
        if init_c2w is not None:
            self.init_c2w = nn.Parameter(init_c2w, requires_grad=False)
        else:
            self.init_c2w = nn.Parameter(torch.eye(4).float().unsqueeze(0).repeat(num_cams, 1, 1), requires_grad=False)
        self.r = nn.Parameter(torch.zeros(size=(num_cams, 3), dtype=torch.float32), requires_grad=learn_R)
        self.t = nn.Parameter(torch.zeros(size=(num_cams, 3), dtype=torch.float32), requires_grad=learn_t)

    def forward(self, cam_id):
        cam_id = int(cam_id)
        return _PoseFn.apply(self.r[cam_id], self.t[cam_id], self.init_c2w[cam_id])


def arange_pixels(resolution=(128, 128), batch_size=1, image_range=(-1., 1.), device=torch.device("cpu")):
    """model/common.py:12-39 (kept for callers that want the full grid; the train step uses pixels_from_indices)."""
    h, w = resolution
    rows, cols = torch.meshgrid(torch.arange(0, h, device=device), torch.arange(0, w, device=device), indexing="ij")
    loc = torch.stack([cols, rows], dim=-1).long().view(1, -1, 2).repeat(batch_size, 1, 1)
    sc = loc.clone().float()
    scale = image_range[1] - image_range[0]
    off = scale / 2
    sc[:, :, 0] = scale * sc[:, :, 0] / (w - 1) - off
    sc[:, :, 1] = scale * sc[:, :, 1] / (h - 1) - off
    return loc, sc


def pixels_from_indices(ray_idx, h, w):
    """Normalised (x, y) of flat pixel ids — the rows `arange_pixels(...)[1][:, ray_idx]` would select, without
    building the h*w grid every step."""
    col = (ray_idx % w).float()
    row = torch.div(ray_idx, w, rounding_mode="floor").float()
    return torch.stack([2.0 * col / (w - 1) - 1.0, 2.0 * row / (h - 1) - 1.0], dim=-1).unsqueeze(0)


class _RayGenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pixels, camera_mat, world_mat, scale_mat):
        pix = pixels.reshape(-1, 2).contiguous().float()
        cam, wm, sc = (m.reshape(4, 4).contiguous().float() for m in (camera_mat, world_mat, scale_mat))
        n, dev = pix.shape[0], pix.device
        o = torch.empty(n, 3, dtype=torch.float32, device=dev)
        d = torch.empty(n, 3, dtype=torch.float32, device=dev)
        nv = torch.empty(n, 1, dtype=torch.float32, device=dev)
        L.call("cope_raygen_fwd", L.ptr(pix), L.ptr(cam), L.ptr(wm), L.ptr(sc), n, L.ptr(o), L.ptr(d), L.ptr(nv), L.stream())
        ctx.save_for_backward(pix, cam, wm, sc)
        ctx.wshape = world_mat.shape
        return o, d, nv

    @staticmethod
    def backward(ctx, d_o, d_d, d_n):
        pix, cam, wm, sc = ctx.saved_tensors
        dw = torch.empty(4, 4, dtype=torch.float32, device=pix.device)
        ws = torch.empty(16, dtype=torch.float32, device=pix.device)
        L.call("cope_raygen_bwd", L.ptr(pix), L.ptr(cam), L.ptr(wm), L.ptr(sc), pix.shape[0],
               L.ptr(d_o.contiguous()) if d_o is not None else None,
               L.ptr(d_d.contiguous()) if d_d is not None else None,
               L.ptr(d_n.contiguous()) if d_n is not None else None, L.ptr(dw), L.ptr(ws), L.stream())
        return None, None, dw.reshape(ctx.wshape), None


def get_world_cameraOrigin_cameraRay(pixels, camera_mat, world_mat, scale_mat):
    """model/training.py:474-487 — rays (origin, unit direction, |direction|) of normalised pixels (1,N,2)."""
    return _RayGenFn.apply(pixels, camera_mat, world_mat, scale_mat)
