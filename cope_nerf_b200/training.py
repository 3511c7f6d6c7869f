"""Step glue of the reference's `mdl.Trainer` that sits on the hot path (model/training.py:101-124, 413-487,
490-558) plus the loss lines of train.py:472-477, 526 — restricted to what the NeuS render/train step needs.
Everything here is thin host code over the CUDA kernels; dataset I/O, logging and visualisation are out of scope.
"""
import numpy as np
import torch

from .common import get_world_cameraOrigin_cameraRay, pixels_from_indices

__all__ = ["near_far_from_sphere", "get_cos_anneal_ratio", "get_patch_indices", "process_data", "sample_rays", "eikonal_loss",
           "rgb_l1_loss", "sdf_flow_loss", "neus_losses", "render_train_step", "build_networks", "DEFAULT_CFG", "render_image", "scene_flow_affine"]

# configs/default.yaml:103-156
DEFAULT_CFG = dict(
    neus_sdf_network=dict(d_out=257, d_in=4, d_hidden=256, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
                          geometric_init=True, weight_norm=True),
    neus_variance_network=dict(init_val=0.3),
    neus_rendering_network=dict(d_feature=256, mode="idr", d_in=11, d_out=3, d_hidden=256, n_layers=4,
                                weight_norm=True, multires_view=4, squeeze_out=True, use_negative_ray_vector=False),
    neus_renderer=dict(n_samples=64, n_importance=64, n_outside=0, up_sample_steps=4, perturb=1.0,
                       n_max_network_queries=64000, importance_sampling_start=0, naive_render=False),
)


def build_networks(cfg=None, device="cuda", precision=None):
    """train.py:39-52 — construct the networks from config sections (same kwargs) and wrap them in NeuSRenderer."""
    from .fields import SDFNetwork, RenderingNetwork, SingleVarianceNetwork
    from .renderer import NeuSRenderer
    cfg = cfg or DEFAULT_CFG
    sdf = SDFNetwork(**cfg["neus_sdf_network"]).to(device)
    var = SingleVarianceNetwork(**cfg["neus_variance_network"]).to(device)
    col = RenderingNetwork(**cfg["neus_rendering_network"]).to(device)
    if precision is not None:
        sdf.precision = col.precision = precision
    return NeuSRenderer(None, sdf, var, col, None, **cfg["neus_renderer"]).to(device)


def near_far_from_sphere(rays_o, rays_d, depth_range):
    """model/training.py:101-118: the sphere bounds are overwritten by the constant depth range."""
    near = torch.full((rays_o.shape[0], 1), float(depth_range[0]), dtype=torch.float32, device=rays_o.device)
    far = torch.full((rays_o.shape[0], 1), float(depth_range[1]), dtype=torch.float32, device=rays_o.device)
    return near, far


def get_cos_anneal_ratio(iter_step, anneal_end):
    """model/training.py:120-124."""
    return 1.0 if anneal_end == 0.0 else float(np.min([1.0, iter_step / anneal_end]))


def get_patch_indices(h, w, patch_size, n_points):
    """model/training.py:413-436 — CPU randperm of top-left corners, row-major pixel ids inside each patch."""
    n_patches = n_points // (patch_size ** 2)
    ha, wa = h - patch_size + 1, w - patch_size + 1
    n_patches = min(n_patches, ha * wa)
    corners = torch.randperm(ha * wa)[:n_patches]
    rows, cols = corners // wa, corners % wa
    off = torch.arange(patch_size).repeat(patch_size, 1)
    off = (off + off.t() * w).flatten()
    return ((rows * w + cols).unsqueeze(1) + off.view(-1)).flatten()


def process_data(img, camera_mat, world_mat, scale_mat, n_points, patch_size=1, corners=None, seed=0):
    """process_data (model/training.py:439-471) on the device, one launch for the pixel part: patch corners -> flat pixel
    ids, integer + normalised pixel coordinates, target colours (cope_sample_pixels), then the rays (cope_raygen_fwd).
    img [1,3,h,w] or [3,h,w] on the device.  `corners` = int64 top-left corner ids (e.g. the reference's CPU stream,
    `torch.randperm((h-ps+1)*(w-ps+1))[:n_points // ps**2]`: identical pixels); None draws distinct corners on the device
    from `seed` (no randperm over ~10^6 elements, no h*w pixel grid).
    Returns (sampled_pixel [N,2], normalized_sampled_pixel [N,2], rays_o, rays_d, rays_d_norm, rgb_gt [N,3], ray_idx [N])."""
    from . import _lib as L
    img3 = img.reshape(3, img.shape[-2], img.shape[-1]).contiguous().float()
    h, w = img3.shape[-2:]
    dev = img3.device
    n_patches = min(n_points // (patch_size ** 2), (h - patch_size + 1) * (w - patch_size + 1))
    n = n_patches * patch_size ** 2
    if corners is not None:
        corners = corners[:n_patches].to(dev, torch.int64).contiguous()
    ray_idx = torch.empty(n, dtype=torch.int64, device=dev)
    pix, npix = torch.empty(n, 2, dtype=torch.float32, device=dev), torch.empty(n, 2, dtype=torch.float32, device=dev)
    rgb_gt = torch.empty(n, 3, dtype=torch.float32, device=dev)
    L.call("cope_sample_pixels", L.ptr(corners), int(seed) & (2 ** 64 - 1), h, w, patch_size, n_patches, L.ptr(img3),
           L.ptr(ray_idx), L.ptr(pix), L.ptr(npix), L.ptr(rgb_gt), L.stream())
    o, d, dn = get_world_cameraOrigin_cameraRay(npix.unsqueeze(0), camera_mat, world_mat, scale_mat)
    return pix, npix, o, d, dn, rgb_gt, ray_idx


def sample_rays(ray_idx, h, w, camera_mat, world_mat, scale_mat):
    """process_data's ray part (model/training.py:439-471): pixel ids -> normalised pixels -> world rays."""
    pix = pixels_from_indices(ray_idx.to(camera_mat.device), h, w)
    return (pix,) + tuple(get_world_cameraOrigin_cameraRay(pix, camera_mat, world_mat, scale_mat))


def eikonal_loss(normals):
    """train.py:526."""
    return torch.mean((torch.linalg.norm(normals.reshape(-1, 3), ord=2, dim=-1) - 1.0) ** 2)


def rgb_l1_loss(rgb, rgb_gt):
    """model/training.py:508."""
    return torch.sum(torch.abs(rgb - rgb_gt)) / float(rgb.shape[0])


def sdf_flow_loss(out, ang_vel, vel):
    """train.py:467-477."""
    pts = out["sampled_points"].reshape(-1, 3)
    nrm = out["normals"].reshape(-1, 3)
    fl = out["sdf_flows"].reshape(-1)
    w = out["weights"].reshape(-1).detach()
    flow = torch.linalg.cross(ang_vel.expand_as(pts), pts) + vel
    return torch.sum(torch.abs(torch.sum(flow * nrm, dim=-1) + fl) * w) / (torch.sum(w) + 1e-10)


def neus_losses(out, rgb_gt, rgb_weight=0.33333, eikonal_weight=0.1):
    """compute_loss (model/training.py:490-549) restricted to the rgb + eikonal terms of BASELINE.json's metric."""
    l_rgb = rgb_l1_loss(out["color_fine"], rgb_gt)
    l_eik = eikonal_loss(out["normals"])
    return rgb_weight * l_rgb + eikonal_weight * l_eik, dict(loss_rgb=l_rgb, loss_eikonal=l_eik)


def render_train_step(renderer, pose, cam_id, pixels, camera_mat, scale_mat, rgb_gt, time_step, depth_range,
                      cos_anneal_ratio=0.5, it=1, rgb_weight=0.33333, eikonal_weight=0.1, loss_scale=1.0, fused=True,
                      sdf_weight=0.0, motion=None):
    """One training iteration without the optimiser: pose -> rays -> near/far -> NeuSRenderer -> loss -> backward
    (train.py:425-532 with the rgb + eikonal terms, plus the SDF-flow term when `motion` is given).
    fused=True renders and reduces the losses in one autograd node (NeuSRenderer.forward_losses); fused=False goes through
    the reference-shaped output dict and the torch loss expressions.  Returns (loss, outputs, rays)."""
    world = pose(cam_id)
    o, d, dn = get_world_cameraOrigin_cameraRay(pixels, camera_mat, world, scale_mat)
    near, far = near_far_from_sphere(o, d, depth_range)
    if fused:
        loss, out = renderer.forward_losses(o, d, dn, time_step, near, far, rgb_gt, cos_anneal_ratio=cos_anneal_ratio, it=it,
                                            rgb_weight=rgb_weight, eikonal_weight=eikonal_weight, sdf_weight=sdf_weight,
                                            motion=motion)
    else:
        out = renderer(o, d, dn, time_step, near, far, cos_anneal_ratio=cos_anneal_ratio, it=it, eval=False)
        loss, parts = neus_losses(out, rgb_gt, rgb_weight, eikonal_weight)
        if motion is not None:
            loss = loss + sdf_weight * sdf_flow_loss(out, motion.reshape(-1)[:3], motion.reshape(-1)[3:])
    (loss * loss_scale).backward()
    return loss.detach(), out, (o, d, dn)


def scene_flow_affine(motion_network, time_step, next_time_step, n_sub):
    """The scene-flow integration of model/training.py:202-207, 269-272 as ONE affine map F [3,4]: the reference moves every sample
    point through n_sub Euler sub-steps p <- p + dt (w_t x p + v_t), w_t / v_t = motion_network at linspace(time_step,
    next_time_step, n_sub + 1)[:-1], dt = (next_time_step - time_step) / n_sub.  Each sub-step is affine in p, so their composition is
    M_{n-1} ... M_0 with M_t = [[I + dt [w_t]x, dt v_t], [0, 1]] - one batched network call + one chain launch (cope_pose_chain_fwd)."""
    from .motion import _ChainFn
    dev = motion_network.lin0.bias.device
    ts = torch.linspace(float(time_step), float(next_time_step), int(n_sub) + 1)[:-1].view(-1, 1).to(dev)
    ang, vel = motion_network(ts)
    dt = (float(next_time_step) - float(time_step)) / int(n_sub)
    z = torch.zeros_like(ang[:, 0])
    skew = torch.stack([torch.stack([z, -ang[:, 2], ang[:, 1]], -1), torch.stack([ang[:, 2], z, -ang[:, 0]], -1),
                        torch.stack([-ang[:, 1], ang[:, 0], z], -1)], dim=1)                  # [w]x, model/common.py:255-265 layout
    M = torch.eye(4, dtype=torch.float32, device=dev).repeat(int(n_sub), 1, 1)
    M[:, :3, :3] += dt * skew
    M[:, :3, 3] = dt * vel
    return _ChainFn.apply(M)[-1][:3].contiguous()


def render_image(renderer, world_mat, camera_mat, scale_mat, h, w, time_step, depth_range, cos_anneal_ratio=1.0, it=1,
                 chunk=None, rays=None, flow=None):
    """Full-image evaluation render (model/training.py:157-283 render_visdata / eval.py:133-157 render_eval: rgb, depth, normal and
    predicted-optical-flow maps) without the reference's 1024-ray chunk loop and its per-chunk device -> host copies.

    chunk=None renders the whole pixel range in ONE pass (one launch per kernel of the pipeline) whenever its working set fits the
    free device memory, and otherwise in the fewest equal passes that do (sized with cope_render_mlp_infer_ws_floats);
    an integer forces that many rays per pass.  Every result stays on the device.

    Returns a dict of device tensors over the rendered pixels (row-major):
      rgb (HW,3), depth_pred (HW,1) [= sum w z / |d|, eval mode], weighted_z_vals (HW,1),
      depth_highest_weight (HW,) [-z of the arg-max-weight sample in the camera frame], normal (HW,3) [sum w n, rotated
      into the camera frame by world_mat[:3,:3]], and with `flow` = (motion_network, time_step, next_time_step, n_sub) also
      flow_pred (HW,2): the forward optical flow in pixels (:265-283, 296-297; n_sub = nb_sample_timestep * frame distance, :202).
    `rays` = (first, count) restricts the render to a contiguous pixel range (multi-GPU: one range per rank)."""
    from . import _lib as L
    from .common import get_world_cameraOrigin_cameraRay, pixels_from_indices
    dev = world_mat.device
    first, count = rays if rays is not None else (0, h * w)
    out = dict(rgb=torch.empty(count, 3, device=dev), depth_pred=torch.empty(count, 1, device=dev),
               weighted_z_vals=torch.empty(count, 1, device=dev), depth_highest_weight=torch.empty(count, device=dev),
               normal=torch.empty(count, 3, device=dev))
    wm = world_mat.detach().contiguous().float()
    F = KS = None
    if flow is not None:
        with torch.no_grad():
            F = scene_flow_affine(*flow)
            KS = (scale_mat.reshape(-1, 4, 4)[0, :3, :3] @ camera_mat.reshape(-1, 4, 4)[0, :3, :3]).contiguous().float()
        out["flow_pred"] = torch.empty(count, 2, device=dev)
    if chunk is None:
        # bytes per ray of one pass, from the library's own size query (the inference workspace: packed weights, H stack, colour
        # input, PE gradients, ... ~26 kB per sample point) x the 1.25 growth slack of the scratch buffer, plus the per-sample
        # outputs and sampling intermediates
        n_samples = renderer.n_samples + renderer.n_importance
        sn, cn = renderer.sdf_network, renderer.color_network
        probe = 1024 * n_samples
        ws = L.query("cope_render_mlp_infer_ws_floats", sn.desc, cn.desc, probe, sn.precision)
        per_ray = int(ws * 4 * 1.25 / 1024) + n_samples * 160 + 4096
        free = torch.cuda.mem_get_info(dev)[0] if torch.cuda.is_available() else 0
        budget = max(int(free * 0.6), per_ray * 1024)
        passes = max(1, -(-count * per_ray // budget))
        chunk = -(-count // passes)
    with torch.no_grad():
        for c0 in range(0, count, chunk):
            n = min(chunk, count - c0)
            idx = torch.arange(first + c0, first + c0 + n, device=dev)
            pix = pixels_from_indices(idx, h, w)
            o, d, dn = get_world_cameraOrigin_cameraRay(pix, camera_mat, world_mat, scale_mat)
            near, far = near_far_from_sphere(o, d, depth_range)
            ro = renderer(o, d, dn, time_step, near, far, cos_anneal_ratio=cos_anneal_ratio, it=it, eval=True)
            S = ro['weights'].shape[1]
            L.call("cope_eval_reduce", L.ptr(ro['weights']), L.ptr(ro.grad4), L.ptr(ro.pts4), L.ptr(wm), n, S,
                   L.ptr(out['normal'][c0:c0 + n]), L.ptr(out['depth_highest_weight'][c0:c0 + n]),
                   L.ptr(F), L.ptr(KS), L.ptr(pix.reshape(-1, 2).contiguous()) if F is not None else None, w / 2.0, h / 2.0,
                   L.ptr(out['flow_pred'][c0:c0 + n]) if F is not None else None, L.stream())
            out['rgb'][c0:c0 + n] = ro['color_fine']
            out['depth_pred'][c0:c0 + n] = ro['depth_pred']
            out['weighted_z_vals'][c0:c0 + n] = ro['weighted_z_vals']
    return out
