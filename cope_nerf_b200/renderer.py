"""Drop-in NeuSRenderer (reference: model/neus_renderer.py:107-584) on libcope_b200.

`forward` keeps the reference signature and the 15 output keys.  Internally:
  coarse z (cope_coarse_z) -> [no grad] SDF queries + 4x (cope_upsample, cope_merge_z) -> one fused autograd
  node for render_core: cope_ray_points -> cope_sdf_fwd (value + analytic gradient in one pass) -> cope_color_fwd
  -> cope_composite_fwd, with a hand-written backward (composite -> colour -> SDF first+second order -> rays).
No host synchronisation happens inside forward/backward (sample_dist is read on the device).
"""
import torch
import torch.nn as nn

from . import _lib as L

__all__ = ["NeuSRenderer", "sample_pdf"]


def _f32(*shape, device):
    return torch.empty(*shape, dtype=torch.float32, device=device)


def sample_pdf(bins, weights, n_samples, det=True, return_inds=False):
    """model/neus_renderer.py:39-70 (det=True).  The CDF is built with torch ops (as the reference does) and the
    inverse-CDF step runs in cope_sample_cdf; indices are bit-exact w.r.t. torch.searchsorted(right=True)."""
    if not det:
        raise NotImplementedError("up_sample only ever uses det=True (model/neus_renderer.py:223)")
    w = weights + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1).contiguous()
    return sample_cdf(cdf, bins, n_samples, return_inds)


def sample_cdf(cdf, bins, n_samples, return_inds=False):
    n, s = cdf.shape
    out = _f32(n, n_samples, device=cdf.device)
    inds = torch.empty(n, n_samples, dtype=torch.int64, device=cdf.device) if return_inds else None
    L.call("cope_sample_cdf", L.ptr(cdf.contiguous()), L.ptr(bins.contiguous()), n, s, n_samples, L.ptr(out),
           L.ptr(inds), L.stream())
    return (out, inds) if return_inds else out


def _call_prec(sdf_net, sdf_flat, col_net, col_flat):
    """`prec` of a render-MLP call: COPE_FLAT_HAS_PACK only when BOTH flat buffers carry their once-per-step weight pack."""
    a, b = sdf_net.call_prec(sdf_flat), col_net.call_prec(col_flat)
    return a if a == b else sdf_net.precision


def _core_forward(rnd, sdf_flat, col_flat, variance, rays_o, rays_d, rays_d_norm, time_step, z, near, far, n_coarse,
                  cos_anneal, eval_mode):
    """render_core forward (model/neus_renderer.py:307-450): points -> fused SDF + colour MLPs -> compositing.
    Returns (outputs, saved): outputs = (color, depth, grad4, weights, sdf, pts4, wz, cdf, wsum, wmax, inv_s, dists, mid_z),
    grad4 = packed (normal | sdf flow) [P,4], pts4 = packed (x,y,z,t) [P,4]."""
    sdf_net, col_net = rnd.sdf_network, rnd.color_network
    dev = z.device
    N, S = z.shape
    P = N * S
    s = L.stream()
    rays_o, rays_d = rays_o.contiguous().float(), rays_d.contiguous().float()
    rays_d_norm = rays_d_norm.contiguous().float()
    z = z.contiguous()
    near, far = near.contiguous().float(), far.contiguous().float()
    tstep = time_step.reshape(-1)[:1].contiguous().float()
    prec_s, prec_c = sdf_net.precision, col_net.precision
    if prec_s != prec_c:
        raise L.CopeError("SDF and colour networks must use the same precision inside NeuSRenderer")

    pts = _f32(P, 4, device=dev)
    dists, mid_z = _f32(N, S, device=dev), _f32(N, S, device=dev)
    L.call("cope_ray_points", L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), L.ptr(tstep), L.ptr(near), L.ptr(far),
           n_coarse, N, S, 1, L.ptr(pts), L.ptr(dists), L.ptr(mid_z), s)

    sdf, grad, rgb = _f32(P, 1, device=dev), _f32(P, 4, device=dev), _f32(P, 3, device=dev)
    sdf_saved = _f32(L.query("cope_sdf_saved_floats", sdf_net.desc, P, 1, prec_s), device=dev)
    col_saved = _f32(L.query("cope_color_saved_floats", col_net.desc, P, prec_c), device=dev)
    ws = L.scratch(L.query("cope_render_mlp_ws_floats", sdf_net.desc, col_net.desc, P, prec_s), dev)
    L.call("cope_render_mlp_fwd", sdf_net.desc, L.ptr(sdf_flat), col_net.desc, L.ptr(col_flat), L.ptr(pts), L.ptr(rays_d), S,
           col_net.multires_view, P, L.ptr(sdf), L.ptr(grad), L.ptr(rgb), L.ptr(sdf_saved), L.ptr(col_saved), L.ptr(ws),
           _call_prec(sdf_net, sdf_flat, col_net, col_flat), s)

    weights, cdf = _f32(N, S, device=dev), _f32(N, S, device=dev)
    color, depth, wz = _f32(N, 3, device=dev), _f32(N, 1, device=dev), _f32(N, 1, device=dev)
    wsum, wmax, inv_s = _f32(N, 1, device=dev), _f32(N, 1, device=dev), _f32(1, device=dev)
    L.call("cope_composite_fwd", L.ptr(sdf), L.ptr(grad), L.ptr(rgb), L.ptr(z), L.ptr(dists), L.ptr(rays_d),
           L.ptr(rays_d_norm), L.ptr(variance), float(cos_anneal), int(eval_mode), N, S, L.ptr(weights),
           L.ptr(color), L.ptr(depth), L.ptr(wz), L.ptr(cdf), L.ptr(wsum), L.ptr(wmax), L.ptr(inv_s), s)
    saved = (sdf_flat, col_flat, variance, rays_d, rays_d_norm, z, dists, mid_z, pts, sdf, grad, rgb, sdf_saved, col_saved)
    return (color, depth, grad, weights, sdf, pts, wz, cdf, wsum, wmax, inv_s, dists, mid_z), saved


def _core_backward(rnd, cfg, saved, need_rays, need_var, d_color, d_depth, d_grad_in, d_weights, d_sdf_up, d_pts):
    """render_core backward: compositing -> colour net -> SDF net (first + second order) -> rays.
    Upstream gradients (each may be None): d_color [N,3], d_depth [N,1], d_grad_in [P,4] (normals | sdf_flows, read only),
    d_weights [N,S], d_sdf_up [P,1]; d_pts [P,4] is a buffer OWNED by the caller's backward (accumulated into here)."""
    (sdf_flat, col_flat, variance, rays_d, rays_d_norm, z, dists, mid_z, pts, sdf, grad, rgb, sdf_saved, col_saved) = saved
    sdf_net, col_net = rnd.sdf_network, rnd.color_network
    N, S, n_coarse, cos_anneal, eval_mode = cfg
    P, dev, s = N * S, z.device, L.stream()
    prec_s = sdf_net.precision
    d_grad = _f32(P, 4, device=dev)           # written (not accumulated) by cope_composite_bwd: no zero fill
    d_sdf, d_rgb = _f32(P, 1, device=dev), _f32(P, 3, device=dev)
    d_rays_d = _f32(N, 3, device=dev)
    # one zero-filled arena for everything the kernels accumulate into: d_variance | dW_colour | dW_sdf (one fill launch)
    n_c, n_s = col_flat.numel(), sdf_flat.numel()
    o_s = 64 + (n_c + 63) // 64 * 64           # every block 256-byte aligned (the weight-gradient kernels use 16-byte reductions)
    arena = torch.zeros(o_s + n_s, dtype=torch.float32, device=dev)
    d_var, d_col_flat, d_sdf_flat = arena[:1], arena[64:64 + n_c].view_as(col_flat), arena[o_s:].view_as(sdf_flat)
    cg = lambda t: L.ptr(t.contiguous().float()) if t is not None else None
    L.call("cope_composite_bwd", L.ptr(sdf), L.ptr(grad), L.ptr(rgb), L.ptr(z), L.ptr(dists), L.ptr(rays_d),
           L.ptr(rays_d_norm), L.ptr(variance), cos_anneal, eval_mode, N, S, cg(d_color),
           cg(d_depth), cg(d_weights), cg(d_grad_in), L.ptr(d_sdf), L.ptr(d_grad), L.ptr(d_rgb), L.ptr(d_var), L.ptr(d_rays_d), s)
    if d_sdf_up is not None:
        d_sdf.add_(d_sdf_up.reshape(P, 1))

    ws = L.scratch(L.query("cope_render_mlp_ws_floats", sdf_net.desc, col_net.desc, P, prec_s), dev)
    d_dirs_pp = None
    if need_rays:
        if d_pts is None:
            d_pts = torch.zeros(P, 4, dtype=torch.float32, device=dev)
        d_dirs_pp = _f32(P, 3, device=dev)
    else:
        d_pts = None
    L.call("cope_render_mlp_bwd", sdf_net.desc, L.ptr(sdf_flat), col_net.desc, L.ptr(col_flat), L.ptr(pts), L.ptr(rays_d), S,
           col_net.multires_view, P, L.ptr(sdf_saved), L.ptr(col_saved), L.ptr(d_sdf), L.ptr(d_grad), L.ptr(d_rgb),
           L.ptr(d_sdf_flat), L.ptr(d_col_flat), L.ptr(d_pts), L.ptr(d_dirs_pp), L.ptr(ws),
           _call_prec(sdf_net, sdf_flat, col_net, col_flat), s)
    d_rays_o = None
    if need_rays:
        d_rays_o = _f32(N, 3, device=dev)
        L.call("cope_ray_points_bwd", L.ptr(d_pts), L.ptr(mid_z), L.ptr(d_dirs_pp), N, S, L.ptr(d_rays_o),
               L.ptr(d_rays_d), s)
    else:
        d_rays_d = None
    d_variance = d_var.reshape(variance.shape) if need_var else None
    return d_sdf_flat, d_col_flat, d_variance, d_rays_o, d_rays_d


class _RenderCoreFn(torch.autograd.Function):
    """render_core (model/neus_renderer.py:307-450) as one autograd node."""

    @staticmethod
    def forward(ctx, rnd, sdf_flat, col_flat, variance, rays_o, rays_d, rays_d_norm, time_step, z, near, far,
                n_coarse, cos_anneal, eval_mode):
        outs, saved = _core_forward(rnd, sdf_flat, col_flat, variance, rays_o, rays_d, rays_d_norm, time_step, z, near, far,
                                    n_coarse, cos_anneal, eval_mode)
        ctx.rnd, ctx.cfg = rnd, (z.shape[0], z.shape[1], n_coarse, float(cos_anneal), int(eval_mode))
        ctx.save_for_backward(*saved)
        ctx.set_materialize_grads(False)          # unused outputs arrive as None instead of freshly filled zero tensors
        ctx.mark_non_differentiable(*outs[6:])    # wz, cdf, wsum, wmax, inv_s, dists, mid_z
        return outs

    @staticmethod
    def backward(ctx, d_color, d_depth, d_grad4, d_weights, d_sdf_up, d_pts4, *unused):
        need_rays = ctx.needs_input_grad[4] or ctx.needs_input_grad[5]
        # the core accumulates into d_pts: never into autograd's own buffer (d_grad4 is only read)
        d_pts = d_pts4.contiguous().clone() if (d_pts4 is not None and need_rays) else None
        d_sdf_flat, d_col_flat, d_variance, d_rays_o, d_rays_d = _core_backward(
            ctx.rnd, ctx.cfg, ctx.saved_tensors, need_rays, ctx.needs_input_grad[3], d_color, d_depth, d_grad4, d_weights,
            d_sdf_up, d_pts)
        return (None, d_sdf_flat, d_col_flat, d_variance, d_rays_o, d_rays_d, None, None, None, None, None, None,
                None, None)


class _RenderStepFn(torch.autograd.Function):
    """render_core AND the step's loss reductions as one autograd node (the fused training path): forward =
    _core_forward + cope_step_losses_fwd, backward = cope_step_losses_bwd + _core_backward.  No per-sample tensor
    crosses autograd for the fused terms, and the ~60 elementwise launches of the torch loss expressions become two.
    Losses: rgb L1 (model/training.py:508), eikonal (train.py:526) and, with `motion` = (angular velocity | velocity)
    [6], the SDF-flow loss (train.py:467-477).  Returns (total, parts[4] = total / rgb / eikonal / sdf-flow, *core outputs).

    The six differentiable core outputs (color, depth, grad4 = normals | sdf_flows, weights, sdf, pts4) stay differentiable:
    a loss built on them OUTSIDE the node (the flow-RGB term on weights / sampled_points, the SDF-consistency term on sdf, the
    depth-smoothness terms on depth_pred — train.py:488-525) sends its gradient into the same backward, where it is added to
    the fused terms' gradients before the one compositing / MLP backward pass."""

    @staticmethod
    def forward(ctx, rnd, sdf_flat, col_flat, variance, rays_o, rays_d, rays_d_norm, time_step, z, near, far,
                n_coarse, cos_anneal, rgb_gt, motion, w_sum_global, w_rgb, w_eik, w_flow):
        outs, saved = _core_forward(rnd, sdf_flat, col_flat, variance, rays_o, rays_d, rays_d_norm, time_step, z, near, far,
                                    n_coarse, cos_anneal, False)
        color, depth, grad4, weights, sdf, pts4 = outs[:6]
        N, S = z.shape
        dev = z.device
        rgb_gt = rgb_gt.contiguous().float()
        mot = motion.detach().reshape(6).contiguous().float() if motion is not None else None
        losses, coef, ws = _f32(4, device=dev), _f32(4, device=dev), _f32(8, device=dev)
        L.call("cope_step_losses_fwd", L.ptr(color), L.ptr(rgb_gt), L.ptr(grad4), L.ptr(pts4) if mot is not None else None,
               L.ptr(weights) if mot is not None else None, L.ptr(mot), L.ptr(w_sum_global), N, N * S,
               float(w_rgb), float(w_eik), float(w_flow), L.ptr(losses), L.ptr(coef), L.ptr(ws), L.stream())
        ctx.rnd, ctx.cfg = rnd, (N, S, n_coarse, float(cos_anneal), 0)
        ctx.has_motion = mot is not None
        ctx.save_for_backward(*saved, color, weights, rgb_gt, coef, *([mot] if mot is not None else []))
        ctx.set_materialize_grads(False)
        ctx.motion_shape = motion.shape if motion is not None else None
        total = losses[0].clone()                 # a differentiable output must not alias the non-differentiable parts
        ctx.mark_non_differentiable(losses, *outs[6:])    # wz, cdf, wsum, wmax, inv_s, dists, mid_z
        return (total, losses, *outs)

    @staticmethod
    def backward(ctx, g_total, _g_parts, d_color_up, d_depth, d_grad4_up, d_weights, d_sdf_up, d_pts4_up, *unused):
        n_saved = 14
        saved = ctx.saved_tensors[:n_saved]
        color, weights, rgb_gt, coef = ctx.saved_tensors[n_saved:n_saved + 4]
        mot = ctx.saved_tensors[n_saved + 4] if ctx.has_motion else None
        N, S = ctx.cfg[:2]
        P, dev = N * S, color.device
        need_rays = ctx.needs_input_grad[4] or ctx.needs_input_grad[5]
        grad4, pts4 = saved[10], saved[8]
        n_in = 19
        if all(g is None for g in (g_total, d_color_up, d_depth, d_grad4_up, d_weights, d_sdf_up, d_pts4_up)):
            return (None,) * n_in
        d_color = d_grad = d_pts = d_motion = None
        if g_total is not None:
            d_color, d_grad = _f32(N, 3, device=dev), _f32(P, 4, device=dev)
            d_pts = _f32(P, 4, device=dev) if (need_rays and mot is not None) else None
            d_motion = torch.zeros(6, dtype=torch.float32, device=dev) if (mot is not None and ctx.needs_input_grad[14]) else None
            L.call("cope_step_losses_bwd", L.ptr(color), L.ptr(rgb_gt), L.ptr(grad4), L.ptr(pts4) if mot is not None else None,
                   L.ptr(weights) if mot is not None else None, L.ptr(mot), N, P, L.ptr(coef),
                   L.ptr(g_total.reshape(1).float()), L.ptr(d_color), L.ptr(d_grad), L.ptr(d_pts), L.ptr(d_motion), L.stream())
        # gradients that arrive from losses built outside the node on the differentiable outputs
        if d_color_up is not None:
            d_color = d_color_up if d_color is None else d_color.add_(d_color_up)
        if d_grad4_up is not None:
            d_grad = d_grad4_up if d_grad is None else d_grad.add_(d_grad4_up)
        if d_pts4_up is not None and need_rays:
            d_pts = d_pts4_up.contiguous().clone() if d_pts is None else d_pts.add_(d_pts4_up)
        d_sdf_flat, d_col_flat, d_variance, d_rays_o, d_rays_d = _core_backward(
            ctx.rnd, ctx.cfg, saved, need_rays, ctx.needs_input_grad[3], d_color, d_depth, d_grad, d_weights, d_sdf_up, d_pts)
        if d_motion is not None:
            d_motion = d_motion.reshape(ctx.motion_shape)
        return (None, d_sdf_flat, d_col_flat, d_variance, d_rays_o, d_rays_d, None, None, None, None, None, None, None,
                None, d_motion, None, None, None, None)


def _unpack_views(grad4, pts4, n):
    """normals (N,S,3), sdf_flows (N,S,1), sampled_points (N,S,3) as strided views of the packed [P,4] tensors."""
    return grad4[:, :3].reshape(n, -1, 3), grad4[:, 3:].reshape(n, -1, 1), pts4[:, :3].reshape(n, -1, 3)


def _render_core_infer(rnd, sdf_flat, col_flat, variance, rays_o, rays_d, rays_d_norm, time_step, z, near, far, n_coarse,
                       cos_anneal, eval_mode):
    """render_core (model/neus_renderer.py:307-450) without an autograd node: the MLP stage goes through
    cope_render_mlp_infer, which keeps nothing for a backward pass.  Same 14 outputs as _RenderCoreFn.forward."""
    sdf_net, col_net = rnd.sdf_network, rnd.color_network
    dev = z.device
    N, S = z.shape
    P = N * S
    s = L.stream()
    rays_o, rays_d = rays_o.contiguous().float(), rays_d.contiguous().float()
    rays_d_norm = rays_d_norm.contiguous().float()
    z = z.contiguous()
    near, far = near.contiguous().float(), far.contiguous().float()
    tstep = time_step.reshape(-1)[:1].contiguous().float()
    prec = sdf_net.precision
    if prec != col_net.precision:
        raise L.CopeError("SDF and colour networks must use the same precision inside NeuSRenderer")
    pts = _f32(P, 4, device=dev)
    dists, mid_z = _f32(N, S, device=dev), _f32(N, S, device=dev)
    L.call("cope_ray_points", L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), L.ptr(tstep), L.ptr(near), L.ptr(far),
           n_coarse, N, S, 1, L.ptr(pts), L.ptr(dists), L.ptr(mid_z), s)
    sdf, grad, rgb = _f32(P, 1, device=dev), _f32(P, 4, device=dev), _f32(P, 3, device=dev)
    ws = L.scratch(L.query("cope_render_mlp_infer_ws_floats", sdf_net.desc, col_net.desc, P, prec), dev)
    L.call("cope_render_mlp_infer", sdf_net.desc, L.ptr(sdf_flat), col_net.desc, L.ptr(col_flat), L.ptr(pts), L.ptr(rays_d), S,
           col_net.multires_view, P, L.ptr(sdf), L.ptr(grad), L.ptr(rgb), L.ptr(ws), _call_prec(sdf_net, sdf_flat, col_net, col_flat), s)
    weights, cdf = _f32(N, S, device=dev), _f32(N, S, device=dev)
    color, depth, wz = _f32(N, 3, device=dev), _f32(N, 1, device=dev), _f32(N, 1, device=dev)
    wsum, wmax, inv_s = _f32(N, 1, device=dev), _f32(N, 1, device=dev), _f32(1, device=dev)
    L.call("cope_composite_fwd", L.ptr(sdf), L.ptr(grad), L.ptr(rgb), L.ptr(z), L.ptr(dists), L.ptr(rays_d),
           L.ptr(rays_d_norm), L.ptr(variance), float(cos_anneal), int(eval_mode), N, S, L.ptr(weights),
           L.ptr(color), L.ptr(depth), L.ptr(wz), L.ptr(cdf), L.ptr(wsum), L.ptr(wmax), L.ptr(inv_s), s)
    normals = grad[:, :3].reshape(N, S, 3).contiguous()
    flows = grad[:, 3:].clone().reshape(N, S, 1)
    points = pts[:, :3].reshape(N, S, 3).contiguous()
    return color, depth, normals, flows, weights, sdf, points, wz, cdf, wsum, wmax, inv_s, dists, mid_z, grad, pts


class RenderOutputs(dict):
    """The reference's 15-key output dict (model/neus_renderer.py:567-584) plus, as ATTRIBUTES, the packed tensors the keys
    are views of: grad4 [P,4] = (normals | sdf_flows), pts4 [P,4] = (sampled_points | t).  The loss / image-reduction kernels
    read the packed form."""
    grad4 = None
    pts4 = None


class _LazyOutputs(RenderOutputs):
    """Output dict of the fused training path: the per-key views / fills of `NeuSRenderer.forward` are only built when a
    key is read (logging), so the timed step launches none of them."""

    def __init__(self, n, color, depth, grad4, weights, sdf, pts4, wz, cdf, wsum, wmax, inv_s, parts):
        super().__init__()
        self.grad4, self.pts4 = grad4, pts4
        nrm = lambda: _unpack_views(grad4, pts4, n)
        self._make = {
            'sdf': lambda: sdf, 'color_fine': lambda: color, 'depth_pred': lambda: depth, 'weighted_z_vals': lambda: wz,
            's_val': lambda: (1.0 / inv_s).expand(n, 1).clone(), 'cdf_fine': lambda: cdf, 'weight_sum': lambda: wsum,
            'weight_max': lambda: wmax, 'normals': lambda: nrm()[0], 'sdf_flows': lambda: nrm()[1],
            'sampled_points': lambda: nrm()[2], 'weights': lambda: weights,
            'inside_sphere': lambda: torch.ones_like(weights), 'weight_inside': lambda: wsum.reshape(n),
            'weight_outside': lambda: torch.zeros(n, dtype=torch.float32, device=weights.device),
            'loss': lambda: parts[0], 'loss_rgb': lambda: parts[1],
            'loss_eikonal': lambda: parts[2], 'loss_sdf': lambda: parts[3],
        }

    def __missing__(self, k):
        v = self._make[k]()
        self[k] = v
        return v

    def __contains__(self, k):
        return k in self._make

    def keys(self):
        return self._make.keys()


class NeuSRenderer(nn.Module):
    """model/neus_renderer.py:107-584.  Constructor kwargs = the `neus_renderer` config section plus the five
    networks, as train.py:46-52 passes them."""

    def __init__(self, nerf, sdf_network, deviation_network, color_network, motion_network, n_samples, n_importance,
                 n_outside, up_sample_steps, perturb, n_max_network_queries, importance_sampling_start, naive_render):
        super().__init__()
        self.nerf = nerf
        self.sdf_network = sdf_network
        self.deviation_network = deviation_network
        self.color_network = color_network
        self.motion_network = motion_network
        self.n_samples = n_samples
        self.n_importance = n_importance
        self.n_outside = n_outside
        self.up_sample_steps = up_sample_steps
        self.perturb = perturb
        self.n_max_network_queries = n_max_network_queries
        self.importance_sampling_start = importance_sampling_start
        self.naive_render = naive_render
        if n_outside > 0 or naive_render:
            raise NotImplementedError("n_outside > 0 / naive_render are dead paths at every shipped config "
                                      "(configs/default.yaml:151,156) and are out of scope (SURVEY.md §8)")
        self.t_rand_override = None     # tests inject the CPU-RNG jitter the reference would draw

    # -- pieces, exposed for the parity tests ------------------------------------------------------------
    def coarse_z(self, near, far, n_samples, t_rand):
        n = near.shape[0]
        z = _f32(n, n_samples, device=near.device)
        L.call("cope_coarse_z", L.ptr(near.contiguous().float()), L.ptr(far.contiguous().float()), L.ptr(t_rand),
               n, n_samples, L.ptr(z), L.stream())
        return z

    def _points(self, rays_o, rays_d, z, tstep, near, far):
        n, s = z.shape
        pts = _f32(n * s, 4, device=z.device)
        L.call("cope_ray_points", L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), L.ptr(tstep), L.ptr(near), L.ptr(far),
               self.n_samples, n, s, 0, L.ptr(pts), None, None, L.stream())
        return pts

    def up_sample(self, rays_o, rays_d, z_vals, sdf, n_importance, inv_s, return_aux=False):
        """model/neus_renderer.py:178-224."""
        n, s = z_vals.shape
        new_z = _f32(n, n_importance, device=z_vals.device)
        cdf = _f32(n, s, device=z_vals.device) if return_aux else None
        inds = torch.empty(n, n_importance, dtype=torch.int64, device=z_vals.device) if return_aux else None
        L.call("cope_upsample", L.ptr(z_vals.contiguous()), L.ptr(sdf.reshape(n, s).contiguous()), n, s, n_importance,
               float(inv_s), L.ptr(new_z), L.ptr(cdf), L.ptr(inds), L.stream())
        return (new_z, cdf, inds) if return_aux else new_z

    def merge_z(self, z_vals, new_z, sdf=None, new_sdf=None):
        n, s = z_vals.shape
        k = new_z.shape[1]
        z_out = _f32(n, s + k, device=z_vals.device)
        sdf_out = _f32(n, s + k, device=z_vals.device) if sdf is not None else None
        L.call("cope_merge_z", L.ptr(z_vals.contiguous()), L.ptr(new_z.contiguous()),
               L.ptr(sdf.contiguous()) if sdf is not None else None,
               L.ptr(new_sdf.contiguous()) if new_sdf is not None else None, n, s, k, L.ptr(z_out), L.ptr(sdf_out),
               L.stream())
        return z_out, sdf_out

    def cat_z_vals(self, rays_o, rays_d, time_step, z_vals, new_z_vals, sdf, last=False, sdf_flat=None, pack_token=None):
        """model/neus_renderer.py:282-298 (sorted merge instead of a full sort)."""
        if last:
            return self.merge_z(z_vals, new_z_vals)[0], sdf
        n, k = new_z_vals.shape
        flat = sdf_flat if sdf_flat is not None else self.sdf_network.flat_weights().detach()
        tstep = time_step.reshape(-1)[:1].contiguous().float()
        near = far = tstep   # unused when use_mid = 0
        pts = self._points(rays_o.contiguous(), rays_d.contiguous(), new_z_vals, tstep, near, far)
        new_sdf = self.sdf_network.query_flat(flat, pts, pack_token).reshape(n, k)
        return self.merge_z(z_vals, new_z_vals, sdf.reshape(z_vals.shape), new_sdf)

    def sample_z(self, rays_o, rays_d, time_step, near, far, eval, sdf_flat, it=0):
        """Coarse + hierarchical depths of model/neus_renderer.py:456-525 (no gradients)."""
        n = rays_o.shape[0]
        use_imp = it >= self.importance_sampling_start and self.n_importance > 0
        n_samples = self.n_samples if use_imp else self.n_samples + self.n_importance
        t_rand = None
        if not eval:
            t_rand = self.t_rand_override
            if t_rand is None:
                t_rand = torch.rand([n, n_samples])            # CPU generator, as neus_renderer.py:482
            t_rand = t_rand.to(rays_o.device, torch.float32).contiguous()
        with torch.no_grad():
            tstep = time_step.reshape(-1)[:1].contiguous().float()
            near_c, far_c = near.contiguous().float(), far.contiguous().float()
            ro, rd = rays_o.detach().contiguous().float(), rays_d.detach().contiguous().float()
            z = self.coarse_z(near_c, far_c, n_samples, t_rand)
            if use_imp:
                flat = sdf_flat.detach()
                token = []          # the four queries below share one weight pack (cope_sdf_query, COPE_WS_HOLDS_PACK)
                sdf = self.sdf_network.query_flat(flat, self._points(ro, rd, z, tstep, near_c, far_c), token).reshape(n, n_samples)
                k = self.n_importance // self.up_sample_steps
                for i in range(self.up_sample_steps):
                    new_z = self.up_sample(ro, rd, z, sdf, k, 64 * 2 ** i)
                    z, sdf = self.cat_z_vals(ro, rd, tstep, z, new_z, sdf, last=(i + 1 == self.up_sample_steps),
                                             sdf_flat=flat, pack_token=token)
        return z, n_samples

    def forward_losses(self, rays_o, rays_d, ray_d_norm, time_step, near, far, rgb_gt, cos_anneal_ratio=0.0, it=-1,
                       rgb_weight=1.0, eikonal_weight=0.1, sdf_weight=0.0, motion=None, w_sum_global=None):
        """Fused training path: `forward` (eval=False) AND the rgb / eikonal / SDF-flow loss reductions of
        model/training.py:508, train.py:526, :467-477 as ONE autograd node (_RenderStepFn).  `motion` = the MotionNetwork's
        (angular velocity | velocity) at the query time, a [6] tensor (None: no SDF-flow term).  Returns
        (total_loss, dict) where the dict holds the same keys as `forward` (detached) plus loss_rgb / loss_eikonal /
        loss_sdf."""
        n = len(rays_o)
        sdf_flat = self.sdf_network.flat_weights()
        col_flat = self.color_network.flat_weights()
        z, n_coarse = self.sample_z(rays_o, rays_d, time_step, near, far, False, sdf_flat, it)
        (total, parts, color, depth, grad4, weights, sdf, pts4, wz, cdf, wsum, wmax, inv_s, dists, mid_z) = \
            _RenderStepFn.apply(self, sdf_flat, col_flat, self.deviation_network.variance, rays_o, rays_d, ray_d_norm,
                                time_step, z, near, far, n_coarse, cos_anneal_ratio, rgb_gt, motion, w_sum_global,
                                rgb_weight, eikonal_weight, sdf_weight)
        out = _LazyOutputs(n, color, depth, grad4, weights, sdf, pts4, wz, cdf, wsum, wmax, inv_s, parts)
        return total, out

    def forward(self, rays_o, rays_d, ray_d_norm, time_step, near, far, perturb_overwrite=-1, background_rgb=None,
                cos_anneal_ratio=0.0, it=-1, eval=False):
        if background_rgb is not None:
            raise NotImplementedError("background_rgb is always None in the reference's callers")
        n = len(rays_o)
        sdf_flat = self.sdf_network.flat_weights()
        col_flat = self.color_network.flat_weights()
        z, n_coarse = self.sample_z(rays_o, rays_d, time_step, near, far, eval, sdf_flat, it)
        if not torch.is_grad_enabled():
            # inference (torch.no_grad(): render_eval / render_visdata, model/training.py:210-283): nothing kept for backward
            (color, depth, normals, flows, weights, sdf, points, wz, cdf, wsum, wmax, inv_s, dists, mid_z, grad4, pts4) = \
                _render_core_infer(self, sdf_flat.detach(), col_flat.detach(), self.deviation_network.variance.detach(), rays_o,
                                   rays_d, ray_d_norm, time_step, z, near, far, n_coarse, cos_anneal_ratio, eval)
        else:
            (color, depth, grad4, weights, sdf, pts4, wz, cdf, wsum, wmax, inv_s, dists, mid_z) = \
                _RenderCoreFn.apply(self, sdf_flat, col_flat, self.deviation_network.variance, rays_o, rays_d, ray_d_norm,
                                    time_step, z, near, far, n_coarse, cos_anneal_ratio, eval)
            normals, flows, points = _unpack_views(grad4, pts4, n)
        out = RenderOutputs({
            'sdf': sdf,
            'color_fine': color,
            'depth_pred': depth,
            'weighted_z_vals': wz,
            's_val': (1.0 / inv_s).expand(n, 1).clone(),
            'cdf_fine': cdf,
            'weight_sum': wsum,
            'weight_max': wmax,
            'normals': normals,
            'sdf_flows': flows,
            'sampled_points': points,
            'weights': weights,
            'inside_sphere': torch.ones_like(weights),
            'weight_inside': wsum.reshape(n).detach(),
            'weight_outside': torch.zeros(n, dtype=torch.float32, device=weights.device),
        })
        out.grad4, out.pts4 = grad4, pts4
        return out
