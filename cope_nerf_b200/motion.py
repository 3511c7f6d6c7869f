"""Drop-in MotionNetwork — the reference's continuous pose model (model/neus_fields.py:79-201): an MLP from the time step
to (angular velocity, velocity), integrated over the sub-steps of every consecutive frame pair into relative camera poses
and chained into world -> camera maps.

Same constructor kwargs, state_dict keys (`lin{l}.weight_g / weight_v / bias`) and method signatures as the reference.
The network is the strict-fp32 MLP path of libcope_b200 with a LeakyReLU(0.2) activation (cope_sdf_fwd / cope_sdf_bwd);
the reference's Python double loop over frame pairs x sub-steps (thousands of tiny launches per training step,
neus_fields.py:142-170) is ONE batched network call + one integration kernel, and `compute_w2c_mappings` is one chain kernel
(cope_pose_integrate_* / cope_pose_chain_*), each with a hand-written backward."""
import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from .embedder import embed_dim
from .fields import WNLinear, _MlpBase, _SdfFn

__all__ = ["MotionNetwork"]


class _IntegrateFn(torch.autograd.Function):
    """(wv [F*n_sub, 6], dt [F]) -> rel [F, 4, 4]   (compute_consecutive_relative_pose for all pairs)"""

    @staticmethod
    def forward(ctx, wv, dt, F, n_sub):
        wv = wv.contiguous().float()
        dt = dt.reshape(F).contiguous().float()
        rel = torch.empty(F, 4, 4, dtype=torch.float32, device=wv.device)
        L.call("cope_pose_integrate_fwd", L.ptr(wv), L.ptr(dt), F, n_sub, L.ptr(rel), L.stream())
        ctx.save_for_backward(wv, dt)
        ctx.shape = (F, n_sub)
        return rel

    @staticmethod
    def backward(ctx, d_rel):
        wv, dt = ctx.saved_tensors
        F, n_sub = ctx.shape
        d_wv = torch.empty_like(wv)
        d_dt = torch.empty(F, dtype=torch.float32, device=wv.device) if ctx.needs_input_grad[1] else None
        L.call("cope_pose_integrate_bwd", L.ptr(wv), L.ptr(dt), F, n_sub, L.ptr(d_rel.contiguous().float()), L.ptr(d_wv),
               L.ptr(d_dt), L.stream())
        return d_wv, d_dt, None, None


class _ChainFn(torch.autograd.Function):
    """rel [F, 4, 4] -> w2c [F+1, 4, 4], w2c_0 = I, w2c_{i+1} = rel_i @ w2c_i   (compute_w2c_mappings)"""

    @staticmethod
    def forward(ctx, rel):
        rel = rel.contiguous().float()
        F = rel.shape[0]
        w2c = torch.empty(F + 1, 4, 4, dtype=torch.float32, device=rel.device)
        L.call("cope_pose_chain_fwd", L.ptr(rel), F, L.ptr(w2c), L.stream())
        ctx.save_for_backward(rel, w2c)
        return w2c

    @staticmethod
    def backward(ctx, d_w2c):
        rel, w2c = ctx.saved_tensors
        F = rel.shape[0]
        d_rel = torch.zeros_like(rel)
        L.call("cope_pose_chain_bwd", L.ptr(rel), L.ptr(w2c), F, L.ptr(d_w2c.contiguous().float()), L.ptr(d_rel), L.stream())
        return d_rel


class MotionNetwork(_MlpBase):
    """model/neus_fields.py:79-201."""

    def __init__(self, d_in, d_out, d_hidden, n_layers, skip_in=(4,), multires=0, bias=0.5, scale=1, geometric_init=True,
                 weight_norm=True, inside_outside=False):
        super().__init__()
        dims = [d_in] + [d_hidden for _ in range(n_layers)] + [d_out]
        self.embed_fn_fine = None
        self.scale = scale
        self.d_in = d_in
        self.multires = multires
        if multires > 0:
            dims[0] = embed_dim(d_in, multires)
            from .embedder import get_embedder
            self.embed_fn_fine = get_embedder(multires, input_dims=d_in)[0]
        self.num_layers = len(dims)
        self.skip_in = tuple(skip_in)
        if len(self.skip_in) > 1:
            raise NotImplementedError("cope_nerf_b200.MotionNetwork supports at most one skip connection")
        dims_in, dims_out = [], []
        for l in range(0, self.num_layers - 1):
            out_dim = dims[l + 1] - dims[0] if (l + 1) in self.skip_in else dims[l + 1]
            lin = nn.Linear(dims[l], out_dim)           # same RNG consumption as the reference constructor (:107-112)
            if geometric_init:                           # (:114-133; the shipped config sets geometric_init False)
                if l == self.num_layers - 2:
                    sgn = -1.0 if inside_outside else 1.0
                    torch.nn.init.normal_(lin.weight, mean=sgn * np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                    torch.nn.init.constant_(lin.bias, -sgn * bias)
                elif multires > 0 and l == 0:
                    torch.nn.init.constant_(lin.bias, 0.0)
                    torch.nn.init.constant_(lin.weight[:, 3:], 0.0)
                    torch.nn.init.normal_(lin.weight[:, :3], 0.0, np.sqrt(2) / np.sqrt(out_dim))
                elif multires > 0 and l in self.skip_in:
                    torch.nn.init.constant_(lin.bias, 0.0)
                    torch.nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
                    torch.nn.init.constant_(lin.weight[:, -(dims[0] - 3):], 0.0)
                else:
                    torch.nn.init.constant_(lin.bias, 0.0)
                    torch.nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
            setattr(self, "lin" + str(l), WNLinear(lin.weight.data, lin.bias.data, weight_norm))
            dims_in.append(dims[l])
            dims_out.append(out_dim)
        self.precision = L.PREC_FP32
        self._dims_in, self._dims_out = list(dims_in), list(dims_out)
        self.desc = L.MlpDesc.make(dims_in, dims_out, d_in, multires, self.skip_in[0] if self.skip_in else -1,
                                   activation=L.ACT_LEAKY_RELU, act_param=0.2)
        offs, off = [], 0
        for a, b in zip(dims_in, dims_out):
            offs += [off, off + a * b]
            off += a * b + b
        offs.append(off)
        self._offsets = tuple(offs)
        self.n_flat = off

    # -- the network (:185-201) ---------------------------------------------------------------------------
    def forward(self, inputs):
        y = _SdfFn.apply(self, self.flat_weights(), inputs, False, L.PREC_FP32)[0] * self.scale
        return y[:, :3], y[:, 3:]

    # -- pose integration ---------------------------------------------------------------------------------
    @staticmethod
    def _pair_times(cam_idx, total_nb_images, nb_sample_timestep):
        """time samples of one consecutive pair, in the reference's arithmetic (:143-148)"""
        target = float(cam_idx)
        ref = target + 1.0
        time_step = target / (total_nb_images - 1) * 2 - 1
        next_time_step = ref / (total_nb_images - 1) * 2 - 1
        n = int(nb_sample_timestep * (ref - target))
        lst = torch.linspace(time_step, next_time_step, n + 1)[:-1]
        return lst, lst[1] - lst[0]

    def _relative_poses(self, first, last, total_nb_images, nb_sample_timestep):
        dev = self.lin0.bias.device
        ts, dts = [], []
        for cam in range(int(first), int(last)):
            lst, dt = self._pair_times(cam, total_nb_images, nb_sample_timestep)
            ts.append(lst); dts.append(dt)
        F = len(ts)
        if F == 0:
            return None, torch.empty(0, 4, 4, device=dev)
        n_sub = ts[0].shape[0]
        ang, vel = self.forward(torch.cat(ts).view(-1, 1).to(dev))
        rel = _IntegrateFn.apply(torch.cat([ang, vel], dim=1), torch.stack(dts).to(dev), F, n_sub)
        return dts[-1].to(dev), rel        # the reference returns the interval of the last pair (:165-167)

    def compute_consecutive_relative_pose(self, target_cam_idx, total_nb_images, nb_sample_timestep):
        dt, rel = self._relative_poses(int(target_cam_idx), int(target_cam_idx) + 1, total_nb_images, nb_sample_timestep)
        return dt, rel[0]

    def compute_relative_camera_pose(self, target_cam_idx, final_ref_cam_idx, total_nb_images, nb_sample_timestep):
        dt, rel = self._relative_poses(target_cam_idx, final_ref_cam_idx, total_nb_images, nb_sample_timestep)
        return dt, list(rel.unbind(0))

    def compute_w2c_mappings(self, relative_camera_pose):
        """world = the first camera of the list (:172-183)."""
        if len(relative_camera_pose) == 0:
            return torch.eye(4, dtype=torch.float32, device=self.lin0.bias.device).unsqueeze(0)
        rel = relative_camera_pose if torch.is_tensor(relative_camera_pose) else torch.stack(list(relative_camera_pose))
        return _ChainFn.apply(rel)
