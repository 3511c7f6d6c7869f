"""Drop-in SDFNetwork / RenderingNetwork / SingleVarianceNetwork (reference: model/neus_fields.py:205-374,
459-465) backed by libcope_b200.  Same constructor kwargs, same state_dict keys (`lin{l}.weight_g`,
`lin{l}.weight_v`, `lin{l}.bias`, `variance`), same forward signatures.

Differences that are deliberate:
  * `SDFNetwork.gradient` is an analytic reverse sweep in CUDA (no autograd.grad / create_graph); its backward
    is the forward-mode tangent pass that the eikonal term needs (cope_sdf_bwd).  Gradients of `gradient(x)`
    with respect to x itself (a Hessian-vector product) are not provided — the reference's render path calls it
    on `pts_time.detach()` (model/neus_renderer.py:356).
  * weight-norm is applied once per call into one flat fp32 buffer [W_l | b_l]* that every kernel reads.
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from .embedder import embed_dim

__all__ = ["SDFNetwork", "RenderingNetwork", "SingleVarianceNetwork", "WNLinear"]


class WNLinear(nn.Module):
    """Parameter holder with the names nn.utils.weight_norm(nn.Linear) produces (weight_g, weight_v, bias)."""

    def __init__(self, weight, bias, weight_norm=True):
        super().__init__()
        self.weight_norm = weight_norm
        self.bias = nn.Parameter(bias.detach().clone())
        if weight_norm:
            self.weight_g = nn.Parameter(weight.detach().norm(dim=1, keepdim=True))
            self.weight_v = nn.Parameter(weight.detach().clone())
        else:
            self.weight = nn.Parameter(weight.detach().clone())

    @property
    def out_features(self):
        return self.bias.shape[0]

    @property
    def in_features(self):
        return (self.weight_v if self.weight_norm else self.weight).shape[1]


def _ptr_array(ts):
    return (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


class _FlatWeights(torch.autograd.Function):
    """(v_0, g_0, b_0, v_1, ...) -> one flat fp32 buffer [W_0 | b_0 | W_1 | b_1 ...]: ONE launch for the whole network
    (cope_flat_weights_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, offsets, pack, *params):
        n = len(params) // 3
        params = tuple(p.contiguous() for p in params)
        vs, gs, bs = params[0::3], params[1::3], params[2::3]
        # pack = (desc, is_color, Lv, tail offset, tail floats): the bf16 operands of the tensor-core path are written ONCE per
        # step behind the fp32 parameters (cope_mlp_pack); every entry point that gets this buffer skips its own re-pack
        n_total = offsets[-1] if pack is None else pack[3] + pack[4]
        flat = torch.empty(n_total, dtype=torch.float32, device=params[0].device)
        meta = dict(rows=(C.c_int * n)(*[v.shape[0] for v in vs]), cols=(C.c_int * n)(*[v.shape[1] for v in vs]),
                    w_off=(C.c_int64 * n)(*offsets[0:2 * n:2]), b_off=(C.c_int64 * n)(*offsets[1:2 * n:2]))
        L.call("cope_flat_weights_fwd", n, _ptr_array(vs), _ptr_array(gs), _ptr_array(bs), meta["rows"], meta["cols"],
               meta["w_off"], meta["b_off"], flat, L.stream())
        if pack is not None:
            L.call("cope_mlp_pack", pack[0], pack[1], pack[2], flat, L.stream())
        ctx.meta, ctx.n = meta, n
        ctx.param_refs = params          # the Parameter objects themselves (their .grad / bucket mark are read in backward)
        ctx.save_for_backward(*params)
        return flat

    @staticmethod
    def backward(ctx, dflat):
        params = ctx.saved_tensors
        n, meta = ctx.n, ctx.meta
        vs, gs, bs = params[0::3], params[1::3], params[2::3]
        dflat = dflat.contiguous()
        # Parameters whose .grad lives in a dist.FlatGradBucket (marked by it) receive their gradient IN PLACE: the kernel adds into
        # the bucket's views, and autograd gets None - instead of ~80 temporaries and one elementwise add launch per tensor.
        refs = ctx.param_refs
        in_place = all(getattr(p, "_cope_grad_in_place", False) and p.grad is not None and p.grad.is_contiguous()
                       and p.grad.dtype == torch.float32 and p.grad.shape == p.shape for p in refs)
        if in_place:
            dvs, dgs, dbs = [p.grad for p in refs[0::3]], [p.grad for p in refs[1::3]], [p.grad for p in refs[2::3]]
        else:
            dvs = [torch.empty_like(v) for v in vs]
            dgs = [torch.empty_like(g) for g in gs]
            dbs = [torch.empty_like(b) for b in bs]
        L.call("cope_flat_weights_bwd", n, _ptr_array(vs), _ptr_array(gs), meta["rows"], meta["cols"], meta["w_off"],
               meta["b_off"], dflat, _ptr_array(dvs), _ptr_array(dgs), _ptr_array(dbs), int(in_place), L.stream())
        if in_place:
            return (None, None) + (None,) * len(params)
        grads = []
        for dv, dg, db in zip(dvs, dgs, dbs):
            grads += [dv, dg, db]
        return (None, None, *grads)


class _MlpBase(nn.Module):
    """Shared: layer bookkeeping + flat weight assembly."""

    def _finish(self, dims_in, dims_out, d_in, multires, skip_layer):
        self._dims_in, self._dims_out = list(dims_in), list(dims_out)
        self.desc = L.MlpDesc.make(dims_in, dims_out, d_in, multires, skip_layer)
        offs, off = [], 0
        for a, b in zip(dims_in, dims_out):
            offs += [off, off + a * b]
            off += a * b + b
        offs.append(off)
        self._offsets = tuple(offs)
        self.n_flat = off

    def _lins(self):
        return [getattr(self, f"lin{l}") for l in range(len(self._dims_in))]

    def _pack_spec(self):
        """(desc, is_color, Lv, tail offset, tail floats) of the once-per-step bf16 weight pack, or None (strict fp32 path, or a
        shape the tensor-core path does not take)."""
        if getattr(self, "precision", L.PREC_FP32) != L.PREC_BF16:
            return None
        cache = self.__dict__.setdefault("_pack_cache", {})
        if "spec" not in cache:
            is_color = 1 if hasattr(self, "multires_view") else 0
            Lv = int(getattr(self, "multires_view", 0))
            try:
                n = L.query("cope_mlp_pack_floats", self.desc, is_color, Lv)
                cache["spec"] = (self.desc, is_color, Lv, L.query("cope_mlp_pack_offset", self.desc), n)
            except L.CopeError:
                cache["spec"] = None
        return cache["spec"]

    def call_prec(self, flat):
        """`prec` argument of a kernel call that receives `flat`: the precision, plus COPE_FLAT_HAS_PACK when `flat` carries the
        packed operands behind the parameters."""
        prec = getattr(self, "precision", L.PREC_FP32)
        return prec | (L.FLAT_HAS_PACK if (prec == L.PREC_BF16 and flat.numel() > self.n_flat) else 0)

    def flat_weights(self):
        """Effective (weight-normalised) parameters as one differentiable flat tensor (followed, on the tensor-core path, by the
        packed bf16 operands: see _pack_spec)."""
        lins = self._lins()
        if all(m.weight_norm for m in lins):
            ps = []
            for m in lins:
                ps += [m.weight_v, m.weight_g, m.bias]
            return _FlatWeights.apply(self._offsets, self._pack_spec(), *ps)
        parts = []
        for m in lins:   # plain nn.Linear parameterisation (weight_norm=False): nothing to normalise
            w = m.weight_v * (m.weight_g / m.weight_v.norm(dim=1, keepdim=True)) if m.weight_norm else m.weight
            parts += [w.reshape(-1), m.bias]
        return torch.cat(parts)


# ------------------------------------------------------------------------------------------------ SDF
class _SdfFn(torch.autograd.Function):
    """y = SDFNetwork.forward(x) and grad = d y[:,0]/dx in one pass (cope_sdf_fwd / cope_sdf_bwd)."""

    @staticmethod
    def forward(ctx, net, flat, x, want_grad, prec):
        P = x.shape[0]
        d_out = net._dims_out[-1]
        dev = x.device
        x = x.contiguous().float()
        y = torch.empty(P, d_out, dtype=torch.float32, device=dev)
        grad = torch.empty(P, net.desc.d_in, dtype=torch.float32, device=dev) if want_grad else None
        saved = torch.empty(L.query("cope_sdf_saved_floats", net.desc, P, int(want_grad), prec & 0xFF), dtype=torch.float32,
                            device=dev)
        ws = L.scratch(L.query("cope_sdf_ws_floats", net.desc, P, prec & 0xFF), dev)
        L.call("cope_sdf_fwd", net.desc, L.ptr(flat), L.ptr(x), P, L.ptr(y), d_out, y.data_ptr() + 4, d_out,
               L.ptr(grad), L.ptr(saved), L.ptr(ws), prec, L.stream())
        ctx.net, ctx.prec, ctx.want_grad = net, prec, want_grad
        ctx.save_for_backward(flat, x, saved)
        return y, grad

    @staticmethod
    def backward(ctx, dy, dgrad):
        flat, x, saved = ctx.saved_tensors
        net, prec = ctx.net, ctx.prec
        P, dev = x.shape[0], x.device
        d_out = net._dims_out[-1]
        dy = dy.contiguous() if dy is not None else None
        dgrad = dgrad.contiguous() if (dgrad is not None and ctx.want_grad) else None
        dflat = torch.zeros_like(flat)
        dx = torch.empty_like(x) if ctx.needs_input_grad[2] else None
        ws = L.scratch(L.query("cope_sdf_ws_floats", net.desc, P, prec & 0xFF), dev)
        L.call("cope_sdf_bwd", net.desc, L.ptr(flat), L.ptr(x), P, L.ptr(saved),
               L.ptr(dy), d_out, (dy.data_ptr() + 4) if dy is not None else None, d_out, L.ptr(dgrad),
               L.ptr(dflat), L.ptr(dx), 0, L.ptr(ws), prec, L.stream())
        return None, dflat, dx, None, None


class _SdfValueFn(torch.autograd.Function):
    """sdf = SDFNetwork.sdf(x) with gradients on the tensor-core path (the SDF-consistency re-query, train.py:504): the forward
    is the fused chain in its value-only form (stores H_1..H_top, no reverse sweep), the backward one fused adjoint sweep + one
    batched weight-gradient launch."""

    @staticmethod
    def forward(ctx, net, flat, x):
        P, dev = x.shape[0], x.device
        x = x.contiguous().float()
        sdf = torch.empty(P, 1, dtype=torch.float32, device=dev)
        saved = torch.empty(L.query("cope_sdf_saved_floats", net.desc, P, 0, L.PREC_BF16), dtype=torch.float32, device=dev)
        ws = L.scratch(L.query("cope_sdf_ws_floats", net.desc, P, L.PREC_BF16), dev)
        L.call("cope_sdf_fwd", net.desc, L.ptr(flat), L.ptr(x), P, L.ptr(sdf), 1, None, 0, None, L.ptr(saved), L.ptr(ws),
               net.call_prec(flat), L.stream())
        ctx.net = net
        ctx.save_for_backward(flat, x, saved)
        return sdf

    @staticmethod
    def backward(ctx, d_sdf):
        flat, x, saved = ctx.saved_tensors
        net = ctx.net
        P, dev = x.shape[0], x.device
        dflat = torch.zeros_like(flat)
        dx = torch.empty_like(x) if ctx.needs_input_grad[2] else None
        ws = L.scratch(L.query("cope_sdf_ws_floats", net.desc, P, L.PREC_BF16), dev)
        L.call("cope_sdf_bwd", net.desc, L.ptr(flat), L.ptr(x), P, L.ptr(saved), L.ptr(d_sdf.contiguous()), 1, None, 0, None,
               L.ptr(dflat), L.ptr(dx), 0, L.ptr(ws), net.call_prec(flat), L.stream())
        return None, dflat, dx


class SDFNetwork(_MlpBase):
    """model/neus_fields.py:205-303."""

    def __init__(self, d_in, d_out, d_hidden, n_layers, skip_in=(4,), multires=0, bias=0.5, scale=1,
                 geometric_init=True, weight_norm=True, inside_outside=False):
        super().__init__()
        dims = [d_in] + [d_hidden for _ in range(n_layers)] + [d_out]
        self.d_in = d_in
        self.multires = multires
        self.embed_fn_fine = None
        if multires > 0:
            dims[0] = embed_dim(d_in, multires)
            from .embedder import get_embedder
            self.embed_fn_fine = get_embedder(multires, input_dims=d_in)[0]
        self.num_layers = len(dims)
        self.skip_in = tuple(skip_in)
        self.scale = scale
        if len(self.skip_in) > 1:
            raise NotImplementedError("cope_nerf_b200.SDFNetwork supports at most one skip connection")
        dims_in, dims_out = [], []
        for l in range(0, self.num_layers - 1):
            out_dim = dims[l + 1] - dims[0] if (l + 1) in self.skip_in else dims[l + 1]
            lin = nn.Linear(dims[l], out_dim)     # same RNG consumption as the reference constructor
            if geometric_init:
                if l == self.num_layers - 2:
                    sgn = -1.0 if inside_outside else 1.0
                    torch.nn.init.normal_(lin.weight, mean=sgn * np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                    torch.nn.init.constant_(lin.bias, -sgn * bias)
                elif multires > 0 and l == 0:
                    torch.nn.init.constant_(lin.bias, 0.0)
                    torch.nn.init.constant_(lin.weight[:, 4:], 0.0)
                    torch.nn.init.normal_(lin.weight[:, :4], 0.0, np.sqrt(2) / np.sqrt(out_dim))
                elif multires > 0 and l in self.skip_in:
                    torch.nn.init.constant_(lin.bias, 0.0)
                    torch.nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
                    torch.nn.init.constant_(lin.weight[:, -(dims[0] - 4):], 0.0)
                else:
                    torch.nn.init.constant_(lin.bias, 0.0)
                    torch.nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
            setattr(self, "lin" + str(l), WNLinear(lin.weight.data, lin.bias.data, weight_norm))
            dims_in.append(dims[l])
            dims_out.append(out_dim)
        self.precision = L.PREC_FP32
        self._finish(dims_in, dims_out, d_in, multires, self.skip_in[0] if self.skip_in else -1)

    # -- functional entry points used by NeuSRenderer (flat weights computed once per step) ----------------
    def apply_flat(self, flat, x, want_grad):
        if self.scale != 1:
            raise NotImplementedError("scale != 1 is not wired into the fused kernels (reference default is 1.0)")
        return _SdfFn.apply(self, flat, x, want_grad, self.call_prec(flat))

    def query_flat(self, flat, x, pack_token=None):
        """sdf only, no autograd state (cope_sdf_query).  `pack_token`: a mutable list shared by consecutive queries with the SAME
        `flat` on the same stream (the four sampling queries of one forward): from the second call on, the packed bf16 weights that
        the first call left at the head of the scratch buffer are reused instead of re-packed."""
        P = x.shape[0]
        x = x.contiguous().float()
        out = torch.empty(P, 1, dtype=torch.float32, device=x.device)
        ws = L.scratch(L.query("cope_sdf_query_ws_floats", self.desc, P, self.precision), x.device)
        prec = self.call_prec(flat)
        if pack_token is not None and prec == L.PREC_BF16:      # no once-per-step pack in `flat`: share one pack between the queries
            key = (ws.data_ptr(), flat.data_ptr(), flat._version, L.stream())
            if pack_token and pack_token[0] == key:
                prec |= L.WS_HOLDS_PACK
            pack_token[:] = [key]
        L.call("cope_sdf_query", self.desc, L.ptr(flat), L.ptr(x), P, L.ptr(out), L.ptr(ws), prec, L.stream())
        return out

    # -- reference API ------------------------------------------------------------------------------------
    def forward(self, inputs):
        return self.apply_flat(self.flat_weights(), inputs, False)[0]

    def sdf(self, x):
        if not torch.is_grad_enabled() or not (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return self.query_flat(self.flat_weights().detach(), x)
        if self.precision == L.PREC_BF16 and self.scale == 1 and x.shape[0] > 0:
            return _SdfValueFn.apply(self, self.flat_weights(), x)
        return self.forward(x)[:, :1]

    def sdf_hidden_appearance(self, x):
        return self.forward(x)

    def gradient(self, x):
        with torch.enable_grad():
            _, g = self.apply_flat(self.flat_weights(), x.detach(), True)
        return g.unsqueeze(1)


# ------------------------------------------------------------------------------------------------ colour
class _ColorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, flat, points, normals, view_dirs, feats, dirs_group, prec):
        P, dev = points.shape[0], points.device
        points, normals, view_dirs = points.contiguous().float(), normals.contiguous().float(), view_dirs.contiguous().float()
        feats = feats.contiguous().float()       # y[:, 1:] views arrive with row stride 257: densify
        rgb = torch.empty(P, net._dims_out[-1], dtype=torch.float32, device=dev)
        saved = torch.empty(L.query("cope_color_saved_floats", net.desc, P, prec & 0xFF), dtype=torch.float32, device=dev)
        ws = L.scratch(L.query("cope_color_ws_floats", net.desc, P, prec & 0xFF), dev)
        L.call("cope_color_fwd", net.desc, L.ptr(flat), L.ptr(points), L.ptr(view_dirs), dirs_group, net.multires_view,
               L.ptr(normals), L.ptr(feats), feats.stride(0), P, L.ptr(rgb), L.ptr(saved), L.ptr(ws), prec, L.stream())
        ctx.net, ctx.prec, ctx.dirs_group = net, prec, dirs_group
        ctx.save_for_backward(flat, view_dirs, saved)
        ctx.feat_shape = feats.shape
        return rgb

    @staticmethod
    def backward(ctx, d_rgb):
        flat, view_dirs, saved = ctx.saved_tensors
        net, prec = ctx.net, ctx.prec
        P, dev = d_rgb.shape[0], d_rgb.device
        need = ctx.needs_input_grad
        dflat = torch.zeros_like(flat)
        dx = torch.zeros(P, 4, dtype=torch.float32, device=dev) if need[2] else None
        dn = torch.zeros(P, 4, dtype=torch.float32, device=dev) if need[3] else None
        dd = torch.empty(P, 3, dtype=torch.float32, device=dev) if need[4] else None
        df = torch.empty(ctx.feat_shape, dtype=torch.float32, device=dev) if need[5] else None
        ws = L.scratch(L.query("cope_color_ws_floats", net.desc, P, prec & 0xFF), dev)
        L.call("cope_color_bwd", net.desc, L.ptr(flat), L.ptr(view_dirs), ctx.dirs_group, net.multires_view, P,
               L.ptr(saved), L.ptr(d_rgb.contiguous()), L.ptr(dflat), L.ptr(dx), L.ptr(dd), L.ptr(dn), L.ptr(df),
               df.shape[1] if df is not None else 0, L.ptr(ws), prec, L.stream())
        if dd is not None and ctx.dirs_group > 1:
            dd = dd.view(-1, ctx.dirs_group, 3).sum(1)
        return None, dflat, dx, dn, dd, df, None, None


class RenderingNetwork(_MlpBase):
    """model/neus_fields.py:307-374 (mode 'idr' only — the only mode any shipped config uses)."""

    def __init__(self, d_feature, mode, d_in, d_out, d_hidden, n_layers, weight_norm=True, multires_view=0,
                 squeeze_out=True, use_negative_ray_vector=False):
        super().__init__()
        if mode != "idr" or not squeeze_out or use_negative_ray_vector:
            raise NotImplementedError("cope_nerf_b200.RenderingNetwork implements mode='idr', squeeze_out=True, "
                                      "use_negative_ray_vector=False (configs/default.yaml:136-147)")
        self.mode, self.squeeze_out, self.use_negative_ray_vector = mode, squeeze_out, use_negative_ray_vector
        self.multires_view = multires_view
        dims = [d_in + d_feature] + [d_hidden for _ in range(n_layers)] + [d_out]
        self.embedview_fn = None
        if multires_view > 0:
            from .embedder import get_embedder
            self.embedview_fn = get_embedder(multires_view)[0]
            dims[0] += embed_dim(3, multires_view) - 3
        self.num_layers = len(dims)
        for l in range(0, self.num_layers - 1):
            lin = nn.Linear(dims[l], dims[l + 1])
            setattr(self, "lin" + str(l), WNLinear(lin.weight.data, lin.bias.data, weight_norm))
        self.precision = L.PREC_FP32
        self._finish(dims[:-1], dims[1:], 4, 0, -1)

    def apply_flat(self, flat, points, normals, view_dirs, feature_vectors, dirs_group=1):
        return _ColorFn.apply(self, flat, points, normals, view_dirs, feature_vectors, dirs_group, self.call_prec(flat))

    def forward(self, points, normals, view_dirs, feature_vectors):
        return self.apply_flat(self.flat_weights(), points, normals, view_dirs, feature_vectors, 1)


class SingleVarianceNetwork(nn.Module):
    """model/neus_fields.py:459-465.  The renderer reads `variance` directly on the device."""

    def __init__(self, init_val):
        super().__init__()
        self.register_parameter("variance", nn.Parameter(torch.tensor(init_val)))

    def forward(self, x):
        return torch.ones([len(x), 1], device=self.variance.device) * torch.exp(self.variance * 10.0)
