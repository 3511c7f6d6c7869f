"""Stage-1 auxiliary losses of the reference's training step (train.py:467-517) on libcope_b200.

  * rgb L1 + eikonal + SDF-flow loss        -> `NeuSRenderer.forward_losses` (renderer._RenderStepFn, cope_step_losses_*), or
                                               `step_losses` below on an existing output dict
  * flow-RGB loss (train.py:480-517)         -> `weighted_points` (cope_weighted_points_*) + `flow_rgb_loss` (cope_flow_rgb_*)
  * SDF-consistency loss (train.py:496-505)  -> `sdf_consistency_loss`: rigid map of the sampled points into the world frame
                                               + `SDFNetwork.sdf` with gradients (cope_sdf_fwd / cope_sdf_bwd)

Everything runs on the device without host synchronisation; there is no CPU fallback.
"""
import torch

from . import _lib as L

__all__ = ["step_losses", "packed_outputs", "weighted_points", "flow_rgb_loss", "sdf_consistency_loss", "projection_matrices", "rigid_inverse",
           "stage1_losses", "depth_smoothness_losses", "SmoothnessLoss", "EdgePreservingSmoothnessLoss", "Stage1Static"]


def _f32(*shape, device):
    return torch.empty(*shape, dtype=torch.float32, device=device)


class _StepLossFn(torch.autograd.Function):
    """rgb L1 + eikonal (+ SDF-flow) on tensors that already exist (the dict path of NeuSRenderer.forward)."""

    @staticmethod
    def forward(ctx, color, rgb_gt, grad4, pts4, weights, motion, w_sum_global, w_rgb, w_eik, w_flow):
        N, P, dev = color.shape[0], grad4.shape[0], color.device
        color, grad4 = color.contiguous().float(), grad4.contiguous().float()
        rgb_gt = rgb_gt.contiguous().float()
        mot = motion.detach().reshape(6).contiguous().float() if motion is not None else None
        if mot is not None:
            pts4, weights = pts4.contiguous().float(), weights.detach().reshape(-1).contiguous().float()
        losses, coef, ws = _f32(4, device=dev), _f32(4, device=dev), _f32(8, device=dev)
        L.call("cope_step_losses_fwd", L.ptr(color), L.ptr(rgb_gt), L.ptr(grad4), L.ptr(pts4) if mot is not None else None,
               L.ptr(weights) if mot is not None else None, L.ptr(mot), L.ptr(w_sum_global), N, P, float(w_rgb), float(w_eik),
               float(w_flow), L.ptr(losses), L.ptr(coef), L.ptr(ws), L.stream())
        ctx.save_for_backward(color, rgb_gt, grad4, coef, *([pts4, weights, mot] if mot is not None else []))
        ctx.has_motion = mot is not None
        ctx.motion_shape = motion.shape if motion is not None else None
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(losses)
        return losses[0].clone(), losses

    @staticmethod
    def backward(ctx, g, _unused=None):
        if g is None:
            return (None,) * 10
        color, rgb_gt, grad4, coef = ctx.saved_tensors[:4]
        pts4 = weights = mot = None
        if ctx.has_motion:
            pts4, weights, mot = ctx.saved_tensors[4:7]
        N, P, dev = color.shape[0], grad4.shape[0], color.device
        d_color, d_grad4 = _f32(N, 3, device=dev), _f32(P, 4, device=dev)
        d_pts4 = _f32(P, 4, device=dev) if (mot is not None and ctx.needs_input_grad[3]) else None
        d_motion = torch.zeros(6, dtype=torch.float32, device=dev) if (mot is not None and ctx.needs_input_grad[5]) else None
        L.call("cope_step_losses_bwd", L.ptr(color), L.ptr(rgb_gt), L.ptr(grad4), L.ptr(pts4), L.ptr(weights), L.ptr(mot), N, P,
               L.ptr(coef), L.ptr(g.reshape(1).float()), L.ptr(d_color), L.ptr(d_grad4), L.ptr(d_pts4), L.ptr(d_motion),
               L.stream())
        if d_motion is not None:
            d_motion = d_motion.reshape(ctx.motion_shape)
        return d_color, None, d_grad4, d_pts4, None, d_motion, None, None, None, None


def packed_outputs(out):
    """(grad4 [P,4], pts4 [P,4]) of a renderer output: the packed tensors a `RenderOutputs` carries, or — for a plain dict
    with the reference's keys — the concatenation of normals | sdf_flows and sampled_points | 0."""
    g4, p4 = getattr(out, "grad4", None), getattr(out, "pts4", None)
    if g4 is None:
        g4 = torch.cat([out["normals"].reshape(-1, 3), out["sdf_flows"].reshape(-1, 1)], dim=1)
    if p4 is None:
        p = out["sampled_points"].reshape(-1, 3)
        p4 = torch.cat([p, torch.zeros_like(p[:, :1])], dim=1)
    return g4, p4


def step_losses(out, rgb_gt, rgb_weight=1.0, eikonal_weight=0.1, sdf_weight=0.0, motion=None, w_sum_global=None):
    """total, parts = losses of model/training.py:508 (rgb L1), train.py:526 (eikonal) and train.py:467-477 (SDF-flow, when
    `motion` = (angular velocity | velocity) [6] is given) from a `NeuSRenderer.forward` output dict.
    parts = [total, rgb, eikonal, sdf_flow] (no gradient)."""
    grad4, pts4 = packed_outputs(out)
    return _StepLossFn.apply(out["color_fine"], rgb_gt, grad4, pts4, out["weights"], motion, w_sum_global,
                             rgb_weight, eikonal_weight, sdf_weight)


# ------------------------------------------------------------------------------------------------ flow-RGB
class _WeightedPointsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, pts4):
        N, S = weights.shape
        weights, pts4 = weights.contiguous().float(), pts4.contiguous().float()
        wp = _f32(N, 4, device=weights.device)
        L.call("cope_weighted_points_fwd", L.ptr(weights), L.ptr(pts4), N, S, L.ptr(wp), L.stream())
        ctx.save_for_backward(weights, pts4)
        return wp

    @staticmethod
    def backward(ctx, d_wp):
        weights, pts4 = ctx.saved_tensors
        N, S = weights.shape
        d_w = _f32(N, S, device=weights.device) if ctx.needs_input_grad[0] else None
        d_p = _f32(N * S, 4, device=weights.device) if ctx.needs_input_grad[1] else None
        L.call("cope_weighted_points_bwd", L.ptr(weights), L.ptr(pts4), L.ptr(d_wp.contiguous()), N, S, L.ptr(d_w), L.ptr(d_p),
               L.stream())
        return d_w, d_p


def weighted_points(weights, pts4):
    """wp [N,4] = sum_s weights[n,s] * (pts4[n,s,:3], 1).  For any rigid map (R, T): sum_s w (R p + T) = R wp[:3] + T wp[3]
    (train.py:488-489 evaluates the left-hand side for every reference frame)."""
    return _WeightedPointsFn.apply(weights, pts4)


class _FlowRgbFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, wp, w2c, KS, norm_pix, pix, ref_imgs, rgb_gt, want_flow):
        N, T = wp.shape[0], w2c.shape[0]
        H, W = ref_imgs.shape[-2:]
        dev = wp.device
        args = [t.contiguous().float() for t in (wp, w2c, KS, norm_pix, pix, ref_imgs, rgb_gt)]
        flow = _f32(T, N, 2, device=dev) if want_flow else None
        loss, ws = _f32(1, device=dev), _f32(2 * T + 1, device=dev)
        L.call("cope_flow_rgb_fwd", *[L.ptr(a) for a in args], N, T, H, W, L.ptr(flow), L.ptr(loss), L.ptr(ws), L.stream())
        ctx.save_for_backward(*args, ws)
        ctx.dims = (N, T, H, W)
        ctx.set_materialize_grads(False)
        if flow is not None:
            ctx.mark_non_differentiable(flow)
        return loss.reshape(()), flow

    @staticmethod
    def backward(ctx, g, _unused=None):
        if g is None:
            return (None,) * 8
        *args, ws = ctx.saved_tensors
        N, T, H, W = ctx.dims
        dev = ws.device
        d_wp = _f32(N, 4, device=dev)
        d_w2c = torch.zeros(T, 4, 4, dtype=torch.float32, device=dev) if ctx.needs_input_grad[1] else None
        L.call("cope_flow_rgb_bwd", *[L.ptr(a) for a in args], N, T, H, W, L.ptr(ws), L.ptr(g.reshape(1).float()), L.ptr(d_wp),
               L.ptr(d_w2c), L.stream())
        return d_wp, d_w2c, None, None, None, None, None, None


def projection_matrices(scale_mat, ref_camera_mats):
    """KS [T,3,3] = scale_mat[0,:3,:3] @ ref_camera_mat[t,:3,:3] (train.py:490)."""
    return scale_mat[0, :3, :3].unsqueeze(0) @ ref_camera_mats[:, :3, :3]


def flow_rgb_loss(wp, w2c, KS, norm_pix, pix, ref_imgs, rgb_gt, return_flow=False):
    """train.py:486-517 for T reference frames: wp [N,4] from `weighted_points`, w2c [T,4,4] world -> reference camera
    (MotionNetwork.compute_w2c_mappings rows), KS [T,3,3] (`projection_matrices`), norm_pix / pix [N,2] the rays' normalised
    and integer pixel coordinates (x, y), ref_imgs [T,3,H,W], rgb_gt [N,3].
    Returns sum_t masked-L1(warped_t, rgb_gt) / 3 (and the predicted pixel flow [T,N,2])."""
    loss, flow = _FlowRgbFn.apply(wp, w2c, KS, norm_pix, pix, ref_imgs, rgb_gt, return_flow)
    return (loss, flow) if return_flow else loss


# ------------------------------------------------------------------------------------------------ depth-patch smoothness
class _PatchSmoothFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, rgb, ps, gamma, w_edge, w_smooth):
        depth = depth.contiguous().float()
        n = depth.numel() // (ps * ps)
        if depth.numel() != n * ps * ps:
            raise L.CopeError(f"depth smoothness: {depth.numel()} depths do not form {ps} x {ps} patches")
        rgb = rgb.contiguous().float() if rgb is not None else None
        dev = depth.device
        losses, ws = _f32(3, device=dev), _f32(12, device=dev)
        L.call("cope_patch_smooth_fwd", L.ptr(depth), L.ptr(rgb), n, ps, float(gamma), float(w_edge), float(w_smooth),
               L.ptr(losses), L.ptr(ws), L.stream())
        ctx.save_for_backward(depth, *([rgb] if rgb is not None else []))
        ctx.cfg = (n, ps, float(gamma), float(w_edge), float(w_smooth))
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(losses)
        return losses[0].clone(), losses

    @staticmethod
    def backward(ctx, g, _unused=None):
        if g is None:
            return (None,) * 6
        depth = ctx.saved_tensors[0]
        rgb = ctx.saved_tensors[1] if len(ctx.saved_tensors) > 1 else None
        n, ps, gamma, w_edge, w_smooth = ctx.cfg
        d_depth = torch.empty_like(depth)
        L.call("cope_patch_smooth_bwd", L.ptr(depth), L.ptr(rgb), n, ps, gamma, w_edge, w_smooth, L.ptr(g.reshape(1).float()),
               L.ptr(d_depth), L.stream())
        return d_depth, None, None, None, None, None


def depth_smoothness_losses(depth_pred, rgb_gt, patch_size, edge_weight=1.0, smooth_weight=0.0, gamma=0.1):
    """train.py:519-525 in one launch each way: depth_pred [N,1] and rgb_gt [N,3] are viewed as patch_size x patch_size
    patches (the rays of `process_data` are patch-major).  Returns (edge_weight * edge + smooth_weight * smooth,
    parts = [total, edge, smooth]) with edge = EdgePreservingSmoothnessLoss(depth, rgb) (model/losses.py:20-38) and
    smooth = SmoothnessLoss(depth) (model/losses.py:7-18); the reference's 1 / 2**s factor belongs in the weights."""
    return _PatchSmoothFn.apply(depth_pred, rgb_gt, int(patch_size), gamma, edge_weight, smooth_weight)


class SmoothnessLoss(torch.nn.Module):
    """model/losses.py:7-18 (same constructor and call signature): inputs [n, ps, ps, 1]."""

    def __init__(self, patch_size):
        super().__init__()
        self.patch_size = patch_size

    def forward(self, inputs):
        return _PatchSmoothFn.apply(inputs, None, inputs.shape[1], 0.1, 0.0, 1.0)[0]


class EdgePreservingSmoothnessLoss(torch.nn.Module):
    """model/losses.py:20-38: inputs [n, ps, ps, 1], weights (the target colours) [n, ps, ps, 3]."""

    def __init__(self, patch_size, bilateral_gamma=0.1):
        super().__init__()
        self.patch_size = patch_size
        self.gamma = bilateral_gamma

    def forward(self, inputs, weights):
        return _PatchSmoothFn.apply(inputs, weights, inputs.shape[1], self.gamma, 1.0, 0.0)[0]


# ------------------------------------------------------------------------------------------------ SDF consistency
def sdf_consistency_loss(sdf_network, pts, sdf, cw2, world_time_step):
    """train.py:496-505: map the sampled points into the world frame with the rigid map cw2 [4,4], query the SDF there at
    the world time step (with gradients, cope_sdf_fwd / cope_sdf_bwd) and compare with the SDF rendered at the query time."""
    pts = pts.reshape(-1, 3)
    pw = pts @ cw2[:3, :3].T + cw2[:3, 3]
    x = torch.cat([pw, torch.full_like(pw[:, :1], float(world_time_step))], dim=1)
    sdf_w = sdf_network.sdf(x)
    return torch.mean(torch.abs(sdf_w - sdf.reshape(-1, 1)))


def rigid_inverse(m):
    """inverse of a rigid 4x4 map [R | T]: [R^T | -R^T T] (the reference calls torch.inverse, train.py:500; the chained
    relative poses are products of rotations and translations, so this is the same matrix without an LU factorisation and
    its host-side singularity check)."""
    r, t = m[:3, :3], m[:3, 3:]
    top = torch.cat([r.T, -(r.T @ t)], dim=1)
    return torch.cat([top, m[3:].detach()], dim=0)


def stage1_losses(out, rgb_gt, motion_network, sdf_network, query_time_step, image_idx, ref_image_idx_list, nb_valid,
                  total_nb_images, nb_sample_timestep, ref_camera_mats, scale_mat, norm_pix, pix, ref_imgs, world_cam_idx,
                  world_time_step, use_flow_rgb=True, use_consistency=True, consistency_pose_grad=True, include_sdf_loss=True):
    """The `not query_in_canonical_space` branch of train.py:467-517 on a NeuSRenderer output dict: SDF-flow loss,
    flow-RGB loss over the valid reference frames and SDF-consistency loss.  Same control flow and argument meaning as the
    reference (image / reference indices, number of valid next time steps, world camera / time step); the per-sample
    arithmetic runs in cope_step_losses_*, cope_weighted_points_*, cope_flow_rgb_* and the SDF kernels.
    `out` may come from `NeuSRenderer.forward` or from `NeuSRenderer.forward_losses`: both keep weights / sampled_points /
    sdf differentiable, so the flow-RGB and consistency terms back-propagate into the renderer as in the reference.  With
    `forward_losses(..., sdf_weight=w, motion=...)` the SDF-flow term is already inside the fused node: pass
    include_sdf_loss=False so it is not evaluated (and not counted) a second time.
    Returns dict(sdf_loss, flow_rgb_loss, sdf_consistency_loss, flow_fw_pred [T,N,2] or None)."""
    dev = rgb_gt.device
    image_idx = int(image_idx)
    refs = [int(r) for r in ref_image_idx_list]
    zero = torch.zeros((), dtype=torch.float32, device=dev)
    res = dict(sdf_loss=zero, flow_rgb_loss=zero, sdf_consistency_loss=zero, flow_fw_pred=None)
    if include_sdf_loss:
        tq = torch.as_tensor([float(query_time_step)], dtype=torch.float32, device=dev).view(-1, 1)
        ang, vel = motion_network(tq)
        res["sdf_loss"] = step_losses(out, rgb_gt, 0.0, 0.0, 1.0, motion=torch.cat([ang, vel], dim=1))[0]
    if (use_flow_rgb or use_consistency) and refs[0] > image_idx:
        need_cons = use_consistency and image_idx != world_cam_idx
        # ONE MotionNetwork call + ONE integration launch for every consecutive frame pair that either term needs (the reference
        # walks the pairs of the reference frames and of the world map in two separate Python loops, train.py:480-483, 498-501)
        first = min(world_cam_idx, image_idx) if need_cons else image_idx
        last = refs[nb_valid - 1] if use_flow_rgb else image_idx
        if need_cons:
            last = max(last, world_cam_idx, image_idx)
        _, rel_all = motion_network._relative_poses(first, last, total_nb_images, nb_sample_timestep)
        if need_cons:
            lo, hi = min(world_cam_idx, image_idx), max(world_cam_idx, image_idx)
            rel_w = rel_all[lo - first:hi - first]
            if not (consistency_pose_grad and torch.is_grad_enabled()):
                rel_w = rel_w.detach()
            c2c_w = motion_network.compute_w2c_mappings(rel_w)[-1]
            cw2 = rigid_inverse(c2c_w) if world_cam_idx <= image_idx else c2c_w
            res["sdf_consistency_loss"] = sdf_consistency_loss(sdf_network, packed_outputs(out)[1][:, :3], out["sdf"], cw2,
                                                               world_time_step)
        if use_flow_rgb:
            rel_f = rel_all[image_idx - first:refs[nb_valid - 1] - first]
            sel = torch.as_tensor([r - image_idx for r in refs[:nb_valid]], device=dev)
            w2c = motion_network.compute_w2c_mappings(rel_f)[sel]
            wp = weighted_points(out["weights"], packed_outputs(out)[1])
            KS = projection_matrices(scale_mat, ref_camera_mats[:nb_valid])
            res["flow_rgb_loss"], res["flow_fw_pred"] = flow_rgb_loss(wp, w2c, KS, norm_pix, pix, ref_imgs[:nb_valid], rgb_gt,
                                                                      return_flow=True)
    return res


# ------------------------------------------------------------------------------------------------ shape-static stage-1 terms
class Stage1Static:
    """The flow-RGB and SDF-consistency terms of train.py:480-517 in a form whose tensor shapes do NOT depend on the frame index,
    so that the whole stage-1 iteration (render, losses, backward, both optimisers) can be captured in ONE CUDA graph and replayed
    for every frame; `stage1_losses` above follows the reference's control flow (pose chains whose length changes with the frame)
    and is the parity reference of this class (tests/test_gpu_round2.py).

    The reference chains the relative poses rel_a .. rel_{b-1} for every pair (a, b) it needs (query frame -> each reference frame,
    world frame <-> query frame).  Here ONE chain over all consecutive pairs of the sequence is evaluated per step,
    G_k = rel_{k-1} ... rel_0 (one batched MotionNetwork call on the (N - 1) * n_sub fixed time samples, one integration launch,
    one chain launch), and every map the losses need is G_b G_a^-1 with the rigid inverse: the same matrix as the reference's
    product up to fp32 rounding.  In particular cw2 = G_world G_idx^-1 covers both branches of train.py:501 (the inverse when the
    world frame precedes the query frame, the plain product otherwise).  Frame indices arrive as DEVICE tensors:
        idx_t [1] int64, ref_idx_t [T] int64 (clamped to N - 1), ref_valid_t [T] float (1 = the frame exists and counts),
        cons_on_t [1] float (0 when the query frame is the world frame, train.py:496)."""

    def __init__(self, motion_network, total_nb_images, nb_sample_timestep, world_cam_idx, world_time_step):
        from .motion import MotionNetwork
        dev = motion_network.lin0.bias.device
        self.motion, self.n_img, self.n_sub = motion_network, int(total_nb_images), int(nb_sample_timestep)
        self.world_idx = torch.tensor([int(world_cam_idx)], dtype=torch.int64, device=dev)
        self.world_time_step = float(world_time_step)
        ts, dts = [], []
        for cam in range(self.n_img - 1):
            lst, dt = MotionNetwork._pair_times(cam, self.n_img, self.n_sub)
            ts.append(lst); dts.append(dt)
        self.ts = torch.cat(ts).view(-1, 1).to(dev)
        self.dts = torch.stack(dts).to(dev)
        far = torch.zeros(4, 4, dtype=torch.float32, device=dev)        # an invalid reference frame: every projection lands far outside
        far[0, 3] = far[1, 3] = 1e8
        far[2, 3] = far[3, 3] = 1.0
        self.far = far

    def global_chain(self, query_time_step=None):
        """G [N, 4, 4] (G_0 = I).  With `query_time_step` ([1] device tensor) the MotionNetwork is evaluated on the chain's time
        samples AND the query time in ONE batched call; returns (G, motion [1, 6] = angular velocity | velocity at the query
        time, the SDF-flow term's input, train.py:472)."""
        from .motion import _ChainFn, _IntegrateFn
        if query_time_step is None:
            ang, vel = self.motion(self.ts)
            wv, mq = torch.cat([ang, vel], dim=1), None
        else:
            ang, vel = self.motion(torch.cat([self.ts, query_time_step.reshape(1, 1).float()], dim=0))
            all_wv = torch.cat([ang, vel], dim=1)
            wv, mq = all_wv[:-1], all_wv[-1:]
        G = _ChainFn.apply(_IntegrateFn.apply(wv, self.dts, self.n_img - 1, self.n_sub))
        return G if query_time_step is None else (G, mq)

    def losses(self, out, rgb_gt, sdf_network, idx_t, ref_idx_t, ref_valid_t, cons_on_t, ref_camera_mats, scale_mat, norm_pix, pix,
               ref_imgs, use_flow_rgb=True, use_consistency=True, consistency_pose_grad=True, G=None):
        """Returns dict(flow_rgb_loss, sdf_consistency_loss, flow_fw_pred [T,N,2]); `out` from NeuSRenderer.forward / forward_losses.
        `G`: the chain of `global_chain` when the caller already evaluated it (together with the query-time motion)."""
        if G is None:
            G = self.global_chain()
        inv_i = rigid_inverse(G.index_select(0, idx_t)[0])
        zero = torch.zeros((), dtype=torch.float32, device=G.device)
        res = dict(flow_rgb_loss=zero, sdf_consistency_loss=zero, flow_fw_pred=None)
        pts4 = packed_outputs(out)[1]
        if use_consistency:
            cw2 = G.index_select(0, self.world_idx)[0] @ inv_i
            if not (consistency_pose_grad and torch.is_grad_enabled()):
                cw2 = cw2.detach()
            res["sdf_consistency_loss"] = cons_on_t.reshape(()) * sdf_consistency_loss(sdf_network, pts4[:, :3], out["sdf"], cw2,
                                                                                       self.world_time_step)
        if use_flow_rgb:
            w2c = G.index_select(0, ref_idx_t) @ inv_i
            w2c = torch.where(ref_valid_t.view(-1, 1, 1) > 0, w2c, self.far)
            wp = weighted_points(out["weights"], pts4)
            KS = projection_matrices(scale_mat, ref_camera_mats)
            res["flow_rgb_loss"], res["flow_fw_pred"] = flow_rgb_loss(wp, w2c, KS, norm_pix, pix, ref_imgs, rgb_gt, return_flow=True)
        return res
