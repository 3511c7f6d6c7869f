"""PyTorch-CPU restatement of the cope-nerf NeuS render/train hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the
reference lines it follows (paths relative to /root/reference).  State is kept
in plain dicts of tensors keyed like the reference's state_dict
(`lin{l}.weight_g`, `lin{l}.weight_v`, `lin{l}.bias`, `variance`, `r`, `t`), so
checkpoints and fixtures are interchangeable with the reference modules.

Parity status: PINNED against the imported reference (tests/golden/*.npz made by
tests/golden/make_golden.py; checked by tests/test_oracle_golden.py).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

__all__ = [
    "embed", "embed_dim", "init_sdf_params", "init_color_params", "init_variance_params",
    "wn_weight", "sdf_forward", "sdf_value", "sdf_gradient", "color_forward", "inv_s_from_variance",
    "cdf_from_weights", "search_cdf", "sample_pdf", "up_sample", "cat_z_vals", "coarse_z", "render_core",
    "render",
    "vec2skew", "so3_exp", "make_c2w", "pose_forward", "camera_matrix", "pixel_grid", "patch_indices",
    "ray_generation", "near_far", "cos_anneal_ratio", "eikonal_loss", "rgb_l1_loss", "sdf_flow_loss",
    "train_step", "render_image", "DEFAULT_CFG",
    "MOTION_CFG", "init_motion_params", "motion_forward", "euler_xyz_to_matrix", "consecutive_relative_pose",
    "relative_camera_pose", "w2c_mappings",
    "warp_pixel", "flow_forward_prediction", "flow_rgb_loss", "sdf_consistency_loss", "stage1_losses",
    "smoothness_loss", "edge_smoothness_loss", "stage1_step", "refine_uv", "compute_loss_and_warp_image",
]

# configs/default.yaml:103-156 (no scene config overrides any of these shapes)
DEFAULT_CFG = dict(
    sdf=dict(d_out=257, d_in=4, d_hidden=256, n_layers=8, skip_in=(4,), multires=6, bias=0.5,
             scale=1.0, geometric_init=True, weight_norm=True),
    color=dict(d_feature=256, mode="idr", d_in=11, d_out=3, d_hidden=256, n_layers=4, weight_norm=True,
               multires_view=4, squeeze_out=True),
    variance=dict(init_val=0.3),
    renderer=dict(n_samples=64, n_importance=64, n_outside=0, up_sample_steps=4, perturb=1.0),
)


# --------------------------------------------------------------------------- embedder
def embed_dim(d, n_freqs):
    return d * (1 + 2 * n_freqs)


def embed(x, n_freqs):
    """model/neus_embedder.py:6-51 — [x | sin(2^0 x) | cos(2^0 x) | ... ], blocks d wide."""
    if n_freqs <= 0:
        return x
    cols = [x]
    for k in range(n_freqs):
        f = float(2 ** k)
        cols.append(torch.sin(x * f))
        cols.append(torch.cos(x * f))
    return torch.cat(cols, dim=-1)


# --------------------------------------------------------------------------- parameters
def _wn_split(weight):
    """nn.utils.weight_norm(dim=0): g = row norms (out,1), v = weight (neus_fields.py:261-262)."""
    return weight.norm(dim=1, keepdim=True).clone(), weight.clone()


def init_sdf_params(d_in=4, d_out=257, d_hidden=256, n_layers=8, skip_in=(4,), multires=6, bias=0.5,
                    scale=1.0, geometric_init=True, weight_norm=True, inside_outside=False):
    """model/neus_fields.py:205-264.  Consumes the global torch RNG in the same order as the
    reference constructor (one nn.Linear per layer, then the geometric re-initialisation)."""
    dims = [d_in] + [d_hidden] * n_layers + [d_out]
    if multires > 0:
        dims[0] = embed_dim(d_in, multires)
    n_lin = len(dims) - 1
    p = {}
    for l in range(n_lin):
        out_dim = dims[l + 1] - dims[0] if (l + 1) in skip_in else dims[l + 1]
        lin = torch.nn.Linear(dims[l], out_dim)
        w, b = lin.weight.data, lin.bias.data
        if geometric_init:
            if l == n_lin - 1:
                sign = -1.0 if inside_outside else 1.0
                torch.nn.init.normal_(w, mean=sign * np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                torch.nn.init.constant_(b, -sign * bias)
            elif multires > 0 and l == 0:
                torch.nn.init.constant_(b, 0.0)
                torch.nn.init.constant_(w[:, 4:], 0.0)
                torch.nn.init.normal_(w[:, :4], 0.0, np.sqrt(2) / np.sqrt(out_dim))
            elif multires > 0 and l in skip_in:
                torch.nn.init.constant_(b, 0.0)
                torch.nn.init.normal_(w, 0.0, np.sqrt(2) / np.sqrt(out_dim))
                torch.nn.init.constant_(w[:, -(dims[0] - 4):], 0.0)
            else:
                torch.nn.init.constant_(b, 0.0)
                torch.nn.init.normal_(w, 0.0, np.sqrt(2) / np.sqrt(out_dim))
        g, v = _wn_split(w)
        p[f"lin{l}.bias"] = b.clone()
        p[f"lin{l}.weight_g"] = g
        p[f"lin{l}.weight_v"] = v
    return p


def init_color_params(d_feature=256, mode="idr", d_in=11, d_out=3, d_hidden=256, n_layers=4,
                      weight_norm=True, multires_view=4, squeeze_out=True, use_negative_ray_vector=False):
    """model/neus_fields.py:307-344 (default nn.Linear init, then weight_norm)."""
    dims = [d_in + d_feature] + [d_hidden] * n_layers + [d_out]
    if multires_view > 0:
        dims[0] += embed_dim(3, multires_view) - 3
    p = {}
    for l in range(len(dims) - 1):
        lin = torch.nn.Linear(dims[l], dims[l + 1])
        g, v = _wn_split(lin.weight.data)
        p[f"lin{l}.bias"] = lin.bias.data.clone()
        p[f"lin{l}.weight_g"] = g
        p[f"lin{l}.weight_v"] = v
    return p


def init_variance_params(init_val=0.3):
    """model/neus_fields.py:459-462."""
    return {"variance": torch.tensor(init_val)}


def wn_weight(p, l):
    """torch._weight_norm(v, g, dim=0) = v * (g / ||v||_row)."""
    v = p[f"lin{l}.weight_v"]
    return v * (p[f"lin{l}.weight_g"] / v.norm(dim=1, keepdim=True))


def _n_lin(p):
    return len([k for k in p if k.endswith(".bias")])


# --------------------------------------------------------------------------- fields
def sdf_forward(p, x, multires=6, skip_in=(4,), scale=1.0):
    """model/neus_fields.py:268-283 — PE, 9 weight-normed linears, softplus(beta=100), skip concat/sqrt2."""
    x = x * scale
    e = embed(x, multires)
    h = e
    n = _n_lin(p)
    for l in range(n):
        if l in skip_in:
            h = torch.cat([h, e], dim=1) / np.sqrt(2)
        h = F.linear(h, wn_weight(p, l), p[f"lin{l}.bias"])
        if l < n - 1:
            h = F.softplus(h, beta=100)
    return torch.cat([h[:, :1] / scale, h[:, 1:]], dim=-1)


def sdf_value(p, x, **kw):
    """model/neus_fields.py:285-286."""
    return sdf_forward(p, x, **kw)[:, :1]


def sdf_gradient(p, x, create_graph=True, **kw):
    """model/neus_fields.py:291-303 — d sdf / d(x,y,z,t) by autograd on a fresh forward; (P,1,4)."""
    with torch.enable_grad():
        x.requires_grad_(True)
        y = sdf_value(p, x, **kw)
        (g,) = torch.autograd.grad(y, x, torch.ones_like(y), create_graph=create_graph, retain_graph=True)
    return g.unsqueeze(1)


def color_forward(p, points, normals, view_dirs, features, multires_view=4, squeeze_out=True):
    """model/neus_fields.py:346-374, mode 'idr': cat[points, PE(view), normals, feature] -> ReLU MLP -> sigmoid."""
    h = torch.cat([points, embed(view_dirs, multires_view), normals, features], dim=-1)
    n = _n_lin(p)
    for l in range(n):
        h = F.linear(h, wn_weight(p, l), p[f"lin{l}.bias"])
        if l < n - 1:
            h = F.relu(h)
    return torch.sigmoid(h) if squeeze_out else h


def inv_s_from_variance(pv):
    """model/neus_fields.py:464-465 + model/neus_renderer.py:360: exp(10 v) clipped to [1e-3, 1e3]; (1,1)."""
    return (torch.ones([1, 1]) * torch.exp(pv["variance"] * 10.0)).clip(1e-3, 1e3)


# --------------------------------------------------------------------------- sampling
def cdf_from_weights(weights):
    """model/neus_renderer.py:42-45."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    c = torch.cumsum(pdf, -1)
    return torch.cat([torch.zeros_like(c[..., :1]), c], -1)


def search_cdf(cdf, bins, n_samples):
    """model/neus_renderer.py:47-70 with det=True; returns (samples, inds int64)."""
    u = torch.linspace(0.5 / n_samples, 1.0 - 0.5 / n_samples, steps=n_samples)
    u = u.expand(list(cdf.shape[:-1]) + [n_samples]).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    lo = (inds - 1).clamp(min=0)
    hi = inds.clamp(max=cdf.shape[-1] - 1)
    c0, c1 = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
    b0, b1 = torch.gather(bins, 1, lo), torch.gather(bins, 1, hi)
    den = c1 - c0
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    return b0 + (u - c0) / den * (b1 - b0), inds


def sample_pdf(bins, weights, n_samples):
    """model/neus_renderer.py:39-70 (det=True is the only mode up_sample uses, :223)."""
    return search_cdf(cdf_from_weights(weights), bins, n_samples)[0]


def up_sample(z, sdf, n_importance, inv_s, return_aux=False):
    """model/neus_renderer.py:178-224 (inside_sphere forced to ones, :187)."""
    n = z.shape[0]
    s0, s1 = sdf[:, :-1], sdf[:, 1:]
    z0, z1 = z[:, :-1], z[:, 1:]
    mid = (s0 + s1) * 0.5
    cos = (s1 - s0) / (z1 - z0 + 1e-5)
    prev = torch.cat([torch.zeros([n, 1]), cos[:, :-1]], dim=-1)
    cos = torch.minimum(prev, cos).clip(-1e3, 0.0)
    dist = z1 - z0
    p_cdf = torch.sigmoid((mid - cos * dist * 0.5) * inv_s)
    n_cdf = torch.sigmoid((mid + cos * dist * 0.5) * inv_s)
    alpha = (p_cdf - n_cdf + 1e-5) / (p_cdf + 1e-5)
    trans = torch.cumprod(torch.cat([torch.ones([n, 1]), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    w = alpha * trans
    cdf = cdf_from_weights(w)
    new_z, inds = search_cdf(cdf, z, n_importance)
    if return_aux:
        return new_z.detach(), dict(weights=w, cdf=cdf, inds=inds)
    return new_z.detach()


def cat_z_vals(sdf_params, rays_o, rays_d, t, z, new_z, sdf, last, **kw):
    """model/neus_renderer.py:282-298 — sort the union; unless last, query SDF at the new points."""
    n, s = z.shape
    k = new_z.shape[1]
    zz, idx = torch.sort(torch.cat([z, new_z], dim=-1), dim=-1)
    if not last:
        pts = (rays_o[:, None, :] + rays_d[:, None, :] * new_z[..., :, None]).reshape(-1, 3)
        x = torch.cat([pts, t.unsqueeze(0).repeat(pts.shape[0], 1)], dim=-1)
        new_sdf = sdf_value(sdf_params, x, **kw).reshape(n, k)
        sdf = torch.gather(torch.cat([sdf, new_sdf], dim=-1), 1, idx)
    return zz, sdf


# --------------------------------------------------------------------------- render_core / forward
def render_core(P, rays_o, rays_d, rays_d_norm, t, z, sample_dist, cos_anneal=0.0, eval_mode=False,
                sdf_kw=None, color_kw=None):
    """model/neus_renderer.py:307-450.  P = dict(sdf=..., color=..., variance=...)."""
    sdf_kw, color_kw = sdf_kw or {}, color_kw or {}
    n, s = z.shape
    tail = torch.tensor([float(torch.as_tensor(sample_dist).detach())]).expand(n, 1)
    dists = torch.cat([z[:, 1:] - z[:, :-1], tail], -1)
    mid_z = z + dists * 0.5
    pts = rays_o[:, None, :] + rays_d[:, None, :] * mid_z[..., :, None]
    dirs = rays_d[:, None, :].expand(pts.shape).reshape(-1, 3)
    pts = pts.reshape(-1, 3)
    x = torch.cat([pts, t.unsqueeze(0).repeat(pts.shape[0], 1)], dim=-1)

    y = sdf_forward(P["sdf"], x, **sdf_kw)
    sdf, feat = y[:, :1], y[:, 1:]
    grad = sdf_gradient(P["sdf"], x.detach(), **sdf_kw).squeeze(1)         # (P,4); :356
    normals, flows = grad[:, :3], grad[:, 3:]
    rgb = color_forward(P["color"], x, grad, dirs, feat, **color_kw).reshape(n, s, 3)

    inv_s = inv_s_from_variance(P["variance"]).expand(n * s, 1)
    true_cos = (dirs * normals).sum(-1, keepdim=True)
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal) + F.relu(-true_cos) * cos_anneal)
    d = dists.reshape(-1, 1)
    p_cdf = torch.sigmoid((sdf - iter_cos * d * 0.5) * inv_s)
    n_cdf = torch.sigmoid((sdf + iter_cos * d * 0.5) * inv_s)
    alpha = ((p_cdf - n_cdf + 1e-5) / (p_cdf + 1e-5)).reshape(n, s).clip(0.0, 1.0)
    trans = torch.cumprod(torch.cat([torch.ones([n, 1]), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    w = alpha * trans
    color = (rgb * w[:, :, None]).sum(dim=1)
    depth = (z * w).sum(dim=1).unsqueeze(-1)                                # z_vals, not mid (:417)
    weighted_z = depth.detach().clone()
    if eval_mode:
        depth = depth / rays_d_norm
    return dict(
        color=color, depth_pred=depth, weighted_z_vals=weighted_z, sdf=sdf, dists=dists,
        normals=normals.reshape(n, s, 3), sdf_flows=flows.reshape(n, s, 1),
        sampled_points=pts.reshape(n, s, 3), s_val=1.0 / inv_s, mid_z_vals=mid_z, weights=w,
        cdf=p_cdf.reshape(n, s), inside_sphere=torch.ones(n, s),
        weight_inside=w.sum(-1).detach(), weight_outside=torch.zeros(n),
        sampled_color=rgb, gradients=grad,
    )


def coarse_z(near, far, n_samples, t_rand=None):
    """model/neus_renderer.py:466-483.  t_rand (N,n_samples) = the train-mode jitter, None = eval."""
    u = torch.linspace(0.0, 1.0, n_samples)
    z = near * (1.0 - u[None, :]) + far * u[None, :]
    if t_rand is not None:
        mids = 0.5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * t_rand
    return z


def render(P, rays_o, rays_d, rays_d_norm, t, near, far, cfg=None, cos_anneal=0.0, eval_mode=False,
           t_rand=None, return_z=False):
    """model/neus_renderer.py:453-584 (n_outside = 0, naive_render False, it >= importance_sampling_start).

    Train mode draws torch.rand([N, n_samples]) from the CPU generator (:482) unless t_rand is given."""
    cfg = cfg or DEFAULT_CFG["renderer"]
    n = rays_o.shape[0]
    n_samples, n_imp, steps = cfg["n_samples"], cfg["n_importance"], cfg["up_sample_steps"]
    sample_dist = (far[0, 0] - near[0, 0]) / n_samples
    if not eval_mode and t_rand is None:
        t_rand = torch.rand([n, n_samples])
    z = coarse_z(near, far, n_samples, None if eval_mode else t_rand)
    if n_imp > 0:
        with torch.no_grad():
            pts = (rays_o[:, None, :] + rays_d[:, None, :] * z[..., :, None]).reshape(-1, 3)
            x = torch.cat([pts, t.unsqueeze(0).repeat(pts.shape[0], 1)], dim=-1)
            sdf = sdf_value(P["sdf"], x).reshape(n, n_samples)
            for i in range(steps):
                new_z = up_sample(z, sdf, n_imp // steps, 64 * 2 ** i)
                z, sdf = cat_z_vals(P["sdf"], rays_o, rays_d, t, z, new_z, sdf, last=(i + 1 == steps))
    rc = render_core(P, rays_o, rays_d, rays_d_norm, t, z, sample_dist, cos_anneal, eval_mode)
    w = rc["weights"]
    s_tot = z.shape[1]
    out = dict(
        sdf=rc["sdf"], color_fine=rc["color"], depth_pred=rc["depth_pred"],
        weighted_z_vals=rc["weighted_z_vals"],
        s_val=rc["s_val"].reshape(n, s_tot).mean(dim=-1, keepdim=True), cdf_fine=rc["cdf"],
        weight_sum=w.sum(dim=-1, keepdim=True), weight_max=torch.max(w, dim=-1, keepdim=True)[0],
        normals=rc["normals"], sdf_flows=rc["sdf_flows"], sampled_points=rc["sampled_points"], weights=w,
        inside_sphere=rc["inside_sphere"], weight_inside=rc["weight_inside"],
        weight_outside=rc["weight_outside"],
    )
    if return_z:
        out["z_vals"] = z
    return out


# --------------------------------------------------------------------------- poses and rays
def vec2skew(v):
    """model/common.py:255-265."""
    z = torch.zeros(1, dtype=torch.float32)
    return torch.stack([torch.cat([z, -v[2:3], v[1:2]]),
                        torch.cat([v[2:3], z, -v[0:1]]),
                        torch.cat([-v[1:2], v[0:1], z])], dim=0)


def so3_exp(r):
    """model/common.py:268-277 (Rodrigues with ||r|| + 1e-15)."""
    k = vec2skew(r)
    n = r.norm() + 1e-15
    return torch.eye(3) + (torch.sin(n) / n) * k + ((1 - torch.cos(n)) / n ** 2) * (k @ k)


def make_c2w(r, t):
    """model/common.py:279-308 — [R | t; 0 0 0 1]; so(3) exp + raw translation."""
    top = torch.cat([so3_exp(r), t.unsqueeze(1)], dim=1)
    return torch.cat([top, torch.tensor([[0.0, 0.0, 0.0, 1.0]])], dim=0)


def pose_forward(pp, cam_id):
    """model/poses_retriever.py:25-32; pp = dict(r=(C,3), t=(C,3), init_c2w=(C,4,4))."""
    cam_id = int(cam_id)
    return make_c2w(pp["r"][cam_id], pp["t"][cam_id]) @ pp["init_c2w"][cam_id]


def camera_matrix(fx, fy, w, h):
    """dataloading/dataset.py:108-111."""
    return torch.tensor([[2 * fx / w, 0, 0, 0], [0, -2 * fy / h, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]],
                        dtype=torch.float32)


def pixel_grid(h, w):
    """model/common.py:12-39 — integer (col,row) and normalised [-1,1] coords of the full h*w grid."""
    rows, cols = torch.meshgrid(torch.arange(0, h), torch.arange(0, w), indexing="ij")
    loc = torch.stack([cols, rows], dim=-1).long().view(1, -1, 2)
    sc = loc.clone().float()
    sc[:, :, 0] = 2.0 * sc[:, :, 0] / (w - 1) - 1.0
    sc[:, :, 1] = 2.0 * sc[:, :, 1] / (h - 1) - 1.0
    return loc, sc


def patch_indices(h, w, patch_size, n_points):
    """model/training.py:413-436 (CPU randperm)."""
    n_patches = n_points // patch_size ** 2
    ha, wa = h - patch_size + 1, w - patch_size + 1
    n_patches = min(n_patches, ha * wa)
    corners = torch.randperm(ha * wa)[:n_patches]
    rows, cols = corners // wa, corners % wa
    off = torch.arange(patch_size).repeat(patch_size, 1)
    off = (off + off.t() * w).flatten()
    return ((rows * w + cols).unsqueeze(1) + off.view(-1)).flatten()


def ray_generation(pixels, camera_mat, world_mat, scale_mat):
    """model/training.py:474-487 + model/common.py:175-215.
    pixels (1,N,2) normalised; camera/scale (1,4,4); world (4,4).  Returns o (N,3), d (N,3), |d| (N,1)."""
    n = pixels.shape[1]
    m = torch.inverse(scale_mat) @ torch.inverse(world_mat) @ torch.inverse(camera_mat)
    origin = torch.zeros(1, 4, n)
    origin[:, -1] = 1.0
    cam_w = (m @ origin)[:, :3].permute(0, 2, 1)
    px = pixels.permute(0, 2, 1)
    px = torch.cat([px, torch.ones_like(px)], dim=1)            # (x, y, 1, 1): depth = 1
    pix_w = (m @ px)[:, :3].permute(0, 2, 1)
    v = pix_w - cam_w
    nv = v.norm(2, 2)
    return cam_w.reshape(-1, 3), (v / nv.unsqueeze(-1)).reshape(-1, 3), nv.view(-1, 1)


def near_far(rays_o, rays_d, depth_range):
    """model/training.py:101-118 — the sphere bounds are overwritten by constants (kept on-graph)."""
    a = torch.sum(rays_d ** 2, dim=-1, keepdim=True)
    b = 2.0 * torch.sum(rays_o * rays_d, dim=-1, keepdim=True)
    mid = 0.5 * (-b) / a
    return (mid - 1.0) * 0 + depth_range[0], (mid + 1.0) * 0 + depth_range[1]


def cos_anneal_ratio(it, anneal_end):
    """model/training.py:120-124."""
    return 1.0 if anneal_end == 0.0 else float(min(1.0, it / anneal_end))


# --------------------------------------------------------------------------- losses / step
def eikonal_loss(normals):
    """train.py:526."""
    return torch.mean((torch.linalg.norm(normals.reshape(-1, 3), ord=2, dim=-1) - 1.0) ** 2)


def rgb_l1_loss(rgb, rgb_gt):
    """model/training.py:508."""
    return torch.sum(torch.abs(rgb - rgb_gt)) / float(rgb.shape[0])


def sdf_flow_loss(out, ang_vel, vel):
    """train.py:467-477 — |(w x p + v).n + d sdf/dt| weighted by detached render weights."""
    pts = out["sampled_points"].reshape(-1, 3)
    nrm = out["normals"].reshape(-1, 3)
    fl = out["sdf_flows"].reshape(-1)
    w = out["weights"].reshape(-1).detach()
    flow = torch.linalg.cross(ang_vel.expand_as(pts), pts) + vel
    return torch.sum(torch.abs(torch.sum(flow * nrm, dim=-1) + fl) * w) / (torch.sum(w) + 1e-10)


def train_step(P, pose, pixels, camera_mat, scale_mat, rgb_gt, t, depth_range, cam_id=0, cos_anneal=0.5,
               rgb_weight=0.33333, eikonal_weight=0.1, t_rand=None, eval_mode=False, cfg=None):
    """One reference training iteration without the optimiser (train.py:425-532 restricted to the rgb +
    eikonal terms): pose -> rays -> near/far -> NeuSRenderer.forward -> loss.  Caller runs .backward()."""
    world = pose_forward(pose, cam_id)
    o, d, dn = ray_generation(pixels, camera_mat, world, scale_mat)
    near, far = near_far(o, d, depth_range)
    out = render(P, o, d, dn, t, near, far, cfg=cfg, cos_anneal=cos_anneal, eval_mode=eval_mode,
                 t_rand=t_rand, return_z=True)
    l_rgb = rgb_l1_loss(out["color_fine"], rgb_gt)
    l_eik = eikonal_loss(out["normals"])
    loss = rgb_weight * l_rgb + eikonal_weight * l_eik
    return loss, dict(out=out, rays_o=o, rays_d=d, rays_d_norm=dn, loss_rgb=l_rgb, loss_eikonal=l_eik)


def render_image(P, world_mat, camera_mat, scale_mat, h, w, t, depth_range, cos_anneal=1.0, chunk=1024, cfg=None, flow=None):
    """Evaluation image render, restating model/training.py:210-283 (render_visdata: rgb / depth / weighted-z /
    depth_highest_weight / normal maps and, with flow = (motion_params, time_step, next_time_step, n_sub), the predicted forward
    optical flow of :203-208, 265-283, 296-297).
    1024-ray chunks as the reference (:210); each chunk: rays (:213), near/far (:217), renderer(eval=True) (:220),
    arg-max-weight depth (:236-243), normal = sum_s normals * weights (:256-262), scene-flow integration of every sample point
    over the sub-steps (:269-272), weight average (:273-275), projection and flow (:277-280)."""
    _, sc = pixel_grid(h, w)
    rgb, depth, wz, dhw, nrm, flw = [], [], [], [], [], []
    ang_list, vel_list = [], []
    if flow is not None:
        mp, t0, t1, n_sub = flow
        for tt in torch.linspace(float(t0), float(t1), int(n_sub) + 1)[:-1]:          # :205-208
            a_t, v_t = motion_forward(mp, tt.view(-1, 1))
            ang_list.append(a_t); vel_list.append(v_t)
    with torch.no_grad():
        for i in range(0, h * w, chunk):
            pix = sc[:, i:i + chunk]
            if pix.shape[1] == 0:
                break                                   # the reference's range(0, n // 1024 + 1) yields an empty last chunk
            o, d, dn = ray_generation(pix, camera_mat, world_mat, scale_mat)
            near, far = near_far(o, d, depth_range)
            out = render(P, o, d, dn, t, near, far, cfg=cfg, cos_anneal=cos_anneal, eval_mode=True)
            wts = out["weights"]
            n = wts.shape[0]
            _, mi = torch.max(wts, dim=1)
            pts = out["sampled_points"].reshape(-1, 3)
            pc = (world_mat @ torch.cat([pts, torch.ones_like(pts[:, [0]])], dim=-1).T).T[:, :3].view(n, wts.shape[1], 3)
            dhw.append(-pc[:, :, -1][torch.arange(n), mi])
            nn_ = (out["normals"] * wts[:, :, None]).sum(dim=1)
            nrm.append((world_mat[:3, :3] @ nn_.T).T)
            rgb.append(out["color_fine"]); depth.append(out["depth_pred"]); wz.append(out["weighted_z_vals"])
            if flow is not None:
                pts_sf = torch.clone(pts)
                interval = (float(t1) - float(t0)) / int(n_sub)
                for k in range(int(n_sub)):
                    sf = torch.linalg.cross(ang_list[k].expand_as(pts_sf), pts_sf) + vel_list[k]
                    pts_sf = pts_sf + interval * sf
                pts_sf = torch.sum(wts.view(n, -1, 1) * pts_sf.view(n, -1, 3), dim=1)
                pm = (scale_mat[0, :3, :3] @ camera_mat[0, :3, :3] @ pts_sf.T).T
                pm = pm[:, :2] / pm[:, 2:]
                fl = pm - pix[0]
                flw.append(torch.stack([fl[:, 0] * (w / 2), fl[:, 1] * (h / 2)], dim=-1))     # :296-297
    res = dict(rgb=torch.cat(rgb), depth_pred=torch.cat(depth), weighted_z_vals=torch.cat(wz),
               depth_highest_weight=torch.cat(dhw), normal=torch.cat(nrm))
    if flow is not None:
        res["flow_pred"] = torch.cat(flw)
    return res


# --------------------------------------------------------------------------- continuous pose model (MotionNetwork)
MOTION_CFG = dict(d_out=6, d_in=1, d_hidden=256, n_layers=4, skip_in=(2,), multires=6, bias=0.5, scale=1.0,
                  geometric_init=False, weight_norm=True)           # configs/default.yaml:113-123


def init_motion_params(d_in=1, d_out=6, d_hidden=256, n_layers=4, skip_in=(2,), multires=6, bias=0.5, scale=1.0,
                       geometric_init=False, weight_norm=True, inside_outside=False):
    """model/neus_fields.py:79-138: one nn.Linear per layer (default init; the geometric branch as in the SDF net but with
    3-wide coordinate blocks, :114-133), weight-normed."""
    dims = [d_in] + [d_hidden] * n_layers + [d_out]
    if multires > 0:
        dims[0] = embed_dim(d_in, multires)
    n_lin = len(dims) - 1
    p = {}
    for l in range(n_lin):
        out_dim = dims[l + 1] - dims[0] if (l + 1) in skip_in else dims[l + 1]
        lin = torch.nn.Linear(dims[l], out_dim)
        w, b = lin.weight.data, lin.bias.data
        if geometric_init:
            if l == n_lin - 1:
                sign = -1.0 if inside_outside else 1.0
                torch.nn.init.normal_(w, mean=sign * np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                torch.nn.init.constant_(b, -sign * bias)
            elif multires > 0 and l == 0:
                torch.nn.init.constant_(b, 0.0)
                torch.nn.init.constant_(w[:, 3:], 0.0)
                torch.nn.init.normal_(w[:, :3], 0.0, np.sqrt(2) / np.sqrt(out_dim))
            elif multires > 0 and l in skip_in:
                torch.nn.init.constant_(b, 0.0)
                torch.nn.init.normal_(w, 0.0, np.sqrt(2) / np.sqrt(out_dim))
                torch.nn.init.constant_(w[:, -(dims[0] - 3):], 0.0)
            else:
                torch.nn.init.constant_(b, 0.0)
                torch.nn.init.normal_(w, 0.0, np.sqrt(2) / np.sqrt(out_dim))
        g, v = _wn_split(w)
        p[f"lin{l}.bias"] = b.clone()
        p[f"lin{l}.weight_g"] = g
        p[f"lin{l}.weight_v"] = v
    return p


def motion_forward(p, t, multires=6, skip_in=(2,), scale=1.0):
    """model/neus_fields.py:185-201: PE(t) -> weight-normed linears with LeakyReLU(0.2), skip concat / sqrt2 -> 6 outputs
    * scale -> (angular velocity, velocity)."""
    e = embed(t, multires)
    x = e
    n = _n_lin(p)
    for l in range(n):
        if l in skip_in:
            x = torch.cat([x, e], dim=1) / np.sqrt(2)
        x = F.linear(x, wn_weight(p, l), p[f"lin{l}.bias"])
        if l < n - 1:
            x = F.leaky_relu(x, 0.2)
    x = x * scale
    return x[:, :3], x[:, 3:]


def _axis_rot(axis, angle):
    """utils_poses/pose_pytorch3d.py:62-75."""
    c, s = torch.cos(angle), torch.sin(angle)
    one, zero = torch.ones_like(angle), torch.zeros_like(angle)
    if axis == "X":
        flat = (one, zero, zero, zero, c, -s, zero, s, c)
    elif axis == "Y":
        flat = (c, zero, s, zero, one, zero, -s, zero, c)
    else:
        flat = (c, -s, zero, s, c, zero, zero, zero, one)
    return torch.stack(flat, -1).reshape(angle.shape + (3, 3))


def euler_xyz_to_matrix(a):
    """utils_poses/pose_pytorch3d.py:8-19 with convention 'XYZ': Rx(a0) @ Ry(a1) @ Rz(a2)."""
    return _axis_rot("X", a[..., 0]) @ _axis_rot("Y", a[..., 1]) @ _axis_rot("Z", a[..., 2])


def consecutive_relative_pose(p, target_cam_idx, total_nb_images, nb_sample_timestep, **kw):
    """model/neus_fields.py:142-160."""
    ref = target_cam_idx + 1.0
    time_step = target_cam_idx / (total_nb_images - 1) * 2 - 1
    next_time_step = ref / (total_nb_images - 1) * 2 - 1
    n = int(nb_sample_timestep * (ref - target_cam_idx))
    lst = torch.linspace(time_step, next_time_step, n + 1)[:-1]
    dt = lst[1] - lst[0]
    ang, vel = motion_forward(p, lst.view(-1, 1), **kw)
    R_list = euler_xyz_to_matrix(ang * dt)
    V_list = vel * dt
    R, T = torch.eye(3), torch.zeros(3)
    for k in range(lst.shape[0]):
        T = (R_list[k] @ T.view(3, 1) + V_list[k].view(3, 1)).view(3)
        R = R @ R_list[k]
    pose = torch.eye(4)
    pose = torch.cat([torch.cat([R, T.view(3, 1)], dim=1), pose[3:]], dim=0)
    return dt, pose


def relative_camera_pose(p, target_cam_idx, final_ref_cam_idx, total_nb_images, nb_sample_timestep, **kw):
    """model/neus_fields.py:162-168."""
    out, dt = [], None
    for cam in range(target_cam_idx, final_ref_cam_idx):
        dt, pose = consecutive_relative_pose(p, cam, total_nb_images, nb_sample_timestep, **kw)
        out.append(pose)
    return dt, out


def w2c_mappings(rel):
    """model/neus_fields.py:172-183."""
    w2c = [torch.eye(4)]
    for r in rel:
        w2c.append(r @ w2c[-1])
    return torch.stack(w2c)


# --------------------------------------------------------------------------- stage-1 auxiliary losses
def warp_pixel(src_frame, uv):
    """train.py:235-244 (normalize_pix=True): bilinear, border-clamped sample of src_frame [1,3,H,W] at pixel coordinates
    uv [N,2] (x, y) -> [N,3]."""
    _, _, h, w = src_frame.shape
    gx = uv[:, 0] / ((w - 1) / 2) - 1
    gy = uv[:, 1] / ((h - 1) / 2) - 1
    grid = torch.stack([gx, gy], dim=-1).view(1, -1, 1, 2)
    return F.grid_sample(src_frame, grid, mode="bilinear", padding_mode="border", align_corners=True)[0, :, :, 0].T


def flow_forward_prediction(pts, weights, n_rays, w2c_t, ref_camera_mat, scale_mat, norm_pix, img_hw):
    """train.py:487-494 for one reference frame: rigid map of the sampled points, render-weight average per ray, projection
    with scale_mat[:3,:3] @ ref_camera_mat[:3,:3], flow in pixels relative to the ray's own normalised pixel."""
    h, w = img_hw
    pts_map = (w2c_t[:3, :3] @ pts.T + w2c_t[:3, 3:]).T
    wpm = torch.sum(weights.view(n_rays, -1, 1) * pts_map.view(n_rays, -1, 3), dim=1)
    pm = (scale_mat[0, :3, :3] @ ref_camera_mat[:3, :3] @ wpm.T).T
    pm = pm[:, :2] / pm[:, 2:]
    d = pm - norm_pix
    return torch.stack([d[:, 0] * (w / 2), d[:, 1] * (h / 2)], dim=-1)


def flow_rgb_loss(flow_list, pix, ref_imgs, rgb_gt):
    """train.py:506-517: warp every reference frame to the sampled pixels, masked L1 against the target colours, summed
    over the frames and divided by 3 (a constant in the reference, whatever the number of valid frames)."""
    total = 0.0
    for t, flow in enumerate(flow_list):
        ref = ref_imgs[t].unsqueeze(0).float()
        corr = pix + flow
        with torch.no_grad():
            lim = torch.tensor([ref.shape[3], ref.shape[2]]).float()
            mask = ((corr >= 0) & (corr < lim)).all(dim=1, keepdim=True)
        warped = warp_pixel(ref, corr)
        total = total + torch.sum(torch.abs(warped - rgb_gt) * mask) / (torch.sum(mask) + 1e-10)
    return total / 3.0


def sdf_consistency_loss(sdf_params, pts, sdf, cw2, world_time_step, **kw):
    """train.py:502-505: the SDF at the world time step, queried at the sampled points mapped by cw2, against the rendered one."""
    pw = (cw2[:3, :3] @ pts.T + cw2[:3, 3:]).T
    x = torch.cat([pw, torch.ones_like(pw[:, :1]) * world_time_step], dim=1)
    return torch.mean(torch.abs(sdf_value(sdf_params, x, **kw) - sdf))


def stage1_losses(sdf_params, motion_params, out, rgb_gt, query_time_step, image_idx, ref_image_idx_list, nb_valid,
                  total_nb_images, nb_sample_timestep, ref_camera_mats, scale_mat, norm_pix, pix, img_hw, ref_imgs,
                  world_cam_idx, world_time_step, use_flow_rgb=True, use_consistency=True, consistency_pose_grad=True,
                  sdf_kw=None, motion_kw=None):
    """train.py:467-517 (the `not query_in_canonical_space` branch): SDF-flow loss, flow-RGB loss over the valid reference
    frames and SDF-consistency loss.  `out` holds sampled_points, normals, sdf_flows, weights, sdf as NeuSRenderer returns them.
    Returns dict(sdf_loss, flow_rgb_loss, sdf_consistency_loss, flow_fw_pred)."""
    sdf_kw, motion_kw = sdf_kw or {}, motion_kw or {}
    pts = out["sampled_points"].reshape(-1, 3)
    weights = out["weights"].reshape(-1)
    n_rays = rgb_gt.shape[0]
    ang, vel = motion_forward(motion_params, torch.as_tensor([float(query_time_step)]).view(-1, 1), **motion_kw)
    res = dict(sdf_loss=sdf_flow_loss(out, ang, vel), flow_rgb_loss=torch.zeros(()), sdf_consistency_loss=torch.zeros(()),
               flow_fw_pred=[])
    if (use_flow_rgb or use_consistency) and int(ref_image_idx_list[0]) > int(image_idx):
        _, c2c = relative_camera_pose(motion_params, int(image_idx), int(ref_image_idx_list[nb_valid - 1]), total_nb_images,
                                      nb_sample_timestep, **motion_kw)
        sel = [int(r) - int(image_idx) for r in ref_image_idx_list[:nb_valid]]
        w2c = w2c_mappings(c2c)[sel]
        flows = [flow_forward_prediction(pts, weights, n_rays, w2c[t], ref_camera_mats[t], scale_mat, norm_pix, img_hw)
                 for t in range(len(w2c))]
        res["flow_fw_pred"] = flows
        if use_consistency and int(image_idx) != world_cam_idx:
            with torch.set_grad_enabled(consistency_pose_grad):
                lo, hi = min(world_cam_idx, int(image_idx)), max(world_cam_idx, int(image_idx))
                _, rel = relative_camera_pose(motion_params, lo, hi, total_nb_images, nb_sample_timestep, **motion_kw)
                c2c_w = w2c_mappings(rel)[-1]
                cw2 = torch.inverse(c2c_w) if world_cam_idx <= int(image_idx) else c2c_w
                pw = (cw2[:3, :3] @ pts.T + cw2[:3, 3:]).T
            x = torch.cat([pw, torch.ones_like(pw[:, :1]) * world_time_step], dim=1)
            res["sdf_consistency_loss"] = torch.mean(torch.abs(sdf_value(sdf_params, x, **sdf_kw) - out["sdf"]))
        if use_flow_rgb:
            res["flow_rgb_loss"] = flow_rgb_loss(flows[:nb_valid], pix, ref_imgs, rgb_gt)
    return res


# --------------------------------------------------------------------------- depth-patch smoothness + the whole stage-1 step
def smoothness_loss(inputs):
    """model/losses.py:7-18 (SmoothnessLoss.forward): inputs [n, ps, ps, 1]."""
    m = lambda x: torch.mean(torch.abs(x))
    l1 = m(inputs[:, :, :-1] - inputs[:, :, 1:])
    l2 = m(inputs[:, :-1, :] - inputs[:, 1:, :])
    l3 = m(inputs[:, :-1, :-1] - inputs[:, 1:, 1:])
    l4 = m(inputs[:, 1:, :-1] - inputs[:, :-1, 1:])
    return (l1 + l2 + l3 + l4) / 4


def edge_smoothness_loss(inputs, weights, gamma=0.1):
    """model/losses.py:20-38 (EdgePreservingSmoothnessLoss.forward): inputs [n, ps, ps, 1], weights [n, ps, ps, 3]."""
    m = lambda x: torch.mean(torch.abs(x))
    bf = lambda x: torch.exp(-torch.abs(x).sum(-1) / gamma).unsqueeze(-1)
    w1 = bf(weights[:, :, :-1] - weights[:, :, 1:])
    w2 = bf(weights[:, :-1, :] - weights[:, 1:, :])
    w3 = bf(weights[:, :-1, :-1] - weights[:, 1:, 1:])
    w4 = bf(weights[:, 1:, :-1] - weights[:, :-1, 1:])
    l1 = m(w1 * (inputs[:, :, :-1] - inputs[:, :, 1:]))
    l2 = m(w2 * (inputs[:, :-1, :] - inputs[:, 1:, :]))
    l3 = m(w3 * (inputs[:, :-1, :-1] - inputs[:, 1:, 1:]))
    l4 = m(w4 * (inputs[:, 1:, :-1] - inputs[:, :-1, 1:]))
    return (l1 + l2 + l3 + l4) / 4


def stage1_step(P, motion_params, rays_o, rays_d, rays_d_norm, near, far, rgb_gt, query_time_step, image_idx,
                ref_image_idx_list, nb_valid, total_nb_images, nb_sample_timestep, ref_camera_mats, scale_mat, norm_pix, pix,
                img_hw, ref_imgs, world_cam_idx, world_time_step, weights, patch_size=4, s_level=0, cos_anneal=0.5,
                t_rand=None, consistency_pose_grad=True, cfg=None, sdf_kw=None, color_kw=None, motion_kw=None):
    """The reference's stage-1 training iteration from the renderer call to the total loss (train.py:441-531 +
    model/training.py:490-531): NeuSRenderer.forward, SDF-flow / flow-RGB / SDF-consistency losses on the UN-DETACHED renderer
    outputs, the two depth-patch smoothness terms on depth_pred, the eikonal term, and compute_loss's weighted sum.
    `weights` = dict(rgb, eikonal, sdf, flow_rgb, sdf_consistency, edge_aware_smoothness, smoothness).
    Returns (loss, parts dict, renderer output dict)."""
    t = torch.as_tensor([float(query_time_step)])
    out = render(P, rays_o, rays_d, rays_d_norm, t, near, far, cfg=cfg, cos_anneal=cos_anneal, eval_mode=False, t_rand=t_rand)
    aux = stage1_losses(P["sdf"], motion_params, out, rgb_gt, query_time_step, image_idx, ref_image_idx_list, nb_valid,
                        total_nb_images, nb_sample_timestep, ref_camera_mats, scale_mat, norm_pix, pix, img_hw, ref_imgs,
                        world_cam_idx, world_time_step, consistency_pose_grad=consistency_pose_grad, sdf_kw=sdf_kw,
                        motion_kw=motion_kw)
    ps = patch_size
    disp = out["depth_pred"].view(-1, ps, ps, 1)                                  # train.py:520-521
    edge = 1 / (2 ** s_level) * edge_smoothness_loss(disp, rgb_gt.view(-1, ps, ps, 3))
    smooth = 1 / (2 ** s_level) * smoothness_loss(disp)
    parts = dict(rgb=rgb_l1_loss(out["color_fine"], rgb_gt), eikonal=eikonal_loss(out["normals"]), sdf=aux["sdf_loss"],
                 flow_rgb=aux["flow_rgb_loss"], sdf_consistency=aux["sdf_consistency_loss"], edge_aware_smoothness=edge,
                 smoothness=smooth)
    loss = sum(weights[k] * parts[k] for k in parts)                              # model/training.py:525-531
    parts["flow_fw_pred"] = aux["flow_fw_pred"]
    return loss, parts, out


# --------------------------------------------------------------------------- pose refinement
def refine_uv(h, w):
    """utils_poses/pose_refinement.py:88-96: [3, H, W] = (col, row, 1) with col, row normalised to [-1, 1]."""
    rows, cols = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    return torch.stack([cols / ((w - 1) / 2) - 1, rows / ((h - 1) / 2) - 1, torch.ones_like(rows)])


def compute_loss_and_warp_image(images, next_images, depths, K_batch, uv_batch, relative_poses):
    """utils_poses/pose_refinement.py:34-61 with warp_pixel_fn = train.py:235-244 (normalize_pix=False)."""
    n = len(images)
    xyz = torch.inverse(K_batch) @ ((uv_batch * depths).view(n, 3, -1))
    txyz = relative_poses[:, :3, :3] @ xyz + relative_poses[:, :3, 3:]
    tuv = K_batch @ txyz
    tdepth = tuv[:, 2:3]
    tuv = (tuv[:, :2] / tdepth).view(uv_batch[:, :2].shape)
    valid = ((tuv[:, 0] >= -1) & (tuv[:, 0] <= 1) & (tuv[:, 1] >= -1) & (tuv[:, 1] <= 1)).float().unsqueeze(1)
    coord = torch.stack([tuv[:, 0], tuv[:, 1]], dim=-1)
    warped = F.grid_sample(next_images, coord, mode="bilinear", padding_mode="border", align_corners=True)
    loss = torch.sum(torch.abs(warped - images) * valid) / torch.sum(valid)
    return loss, warped
