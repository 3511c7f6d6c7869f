"""ctypes binding of libcope_b200.so (include/cope_b200.h).

There is NO fallback: if the library is missing, or a kernel entry point is called without a CUDA device,
this raises.  Building: `python -m cope_nerf_b200.build` (or `__graft_entry__.build()`).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcope_b200.so")
MAX_LIN = 12
PREC_FP32, PREC_BF16 = 0, 1
WS_HOLDS_PACK = 0x100      # cope_sdf_query: the scratch head still holds this network's packed weights
FLAT_HAS_PACK = 0x200      # any bf16 MLP entry point: the packed weights follow the flat parameters (cope_mlp_pack)
ACT_SOFTPLUS100, ACT_LEAKY_RELU = 0, 1

_f = C.c_void_p      # device pointers travel as integers
_i, _l, _fl = C.c_int, C.c_int64, C.c_float


class MlpDesc(C.Structure):
    _fields_ = [("n_lin", C.c_int32), ("d_in", C.c_int32), ("multires", C.c_int32), ("skip_layer", C.c_int32),
                ("dims_in", C.c_int32 * MAX_LIN), ("dims_out", C.c_int32 * MAX_LIN),
                ("activation", C.c_int32), ("act_param", C.c_float)]

    @staticmethod
    def make(dims_in, dims_out, d_in, multires, skip_layer, activation=0, act_param=0.0):
        d = MlpDesc()
        d.n_lin, d.d_in, d.multires, d.skip_layer = len(dims_in), d_in, multires, skip_layer
        d.activation, d.act_param = activation, act_param
        for k, (a, b) in enumerate(zip(dims_in, dims_out)):
            d.dims_in[k], d.dims_out[k] = a, b
        return d


_D = C.POINTER(MlpDesc)
_SIGS = {
    "cope_version": (C.c_int, []),
    "cope_last_error": (C.c_char_p, []),
    "cope_launch_count": (C.c_uint64, []),
    "cope_mlp_flat_floats": (_l, [_D]),
    "cope_mlp_pack_offset": (_l, [_D]),
    "cope_mlp_pack_floats": (_l, [_D, _i, _i]),
    "cope_mlp_pack": (_i, [_D, _i, _i, _f, _f]),
    "cope_weightnorm_fwd": (_i, [_f, _f, _f, _i, _i, _f]),
    "cope_weightnorm_bwd": (_i, [_f, _f, _f, _f, _f, _i, _i, _f]),
    "cope_flat_weights_fwd": (_i, [_i, _f, _f, _f, _f, _f, _f, _f, _f, _f]),
    "cope_flat_weights_bwd": (_i, [_i, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _i, _f]),
    "cope_embed_fwd": (_i, [_f, _l, _i, _i, _f, _f]),
    "cope_sdf_saved_floats": (_l, [_D, _l, _i, _i]),
    "cope_sdf_ws_floats": (_l, [_D, _l, _i]),
    "cope_sdf_query_ws_floats": (_l, [_D, _l, _i]),
    "cope_sdf_query": (_i, [_D, _f, _f, _l, _f, _f, _i, _f]),
    "cope_sdf_fwd": (_i, [_D, _f, _f, _l, _f, _i, _f, _i, _f, _f, _f, _i, _f]),
    "cope_sdf_bwd": (_i, [_D, _f, _f, _l, _f, _f, _i, _f, _i, _f, _f, _f, _i, _f, _i, _f]),
    "cope_color_saved_floats": (_l, [_D, _l, _i]),
    "cope_color_ws_floats": (_l, [_D, _l, _i]),
    "cope_color_fwd": (_i, [_D, _f, _f, _f, _i, _i, _f, _f, _i, _l, _f, _f, _f, _i, _f]),
    "cope_color_bwd": (_i, [_D, _f, _f, _i, _i, _l, _f, _f, _f, _f, _f, _f, _f, _i, _f, _i, _f]),
    "cope_render_mlp_ws_floats": (_l, [_D, _D, _l, _i]),
    "cope_render_mlp_fwd": (_i, [_D, _f, _D, _f, _f, _f, _i, _i, _l, _f, _f, _f, _f, _f, _f, _i, _f]),
    "cope_dbg_render_bwd_eb_offset": (_l, [_D, _D, _l]),
    "cope_render_mlp_infer_ws_floats": (_l, [_D, _D, _l, _i]),
    "cope_render_mlp_infer": (_i, [_D, _f, _D, _f, _f, _f, _i, _i, _l, _f, _f, _f, _f, _i, _f]),
    "cope_eval_reduce": (_i, [_f, _f, _f, _f, _l, _i, _f, _f, _f, _f, _f, _fl, _fl, _f, _f]),
    "cope_pose_integrate_fwd": (_i, [_f, _f, _i, _i, _f, _f]),
    "cope_pose_integrate_bwd": (_i, [_f, _f, _i, _i, _f, _f, _f, _f]),
    "cope_pose_chain_fwd": (_i, [_f, _i, _f, _f]),
    "cope_pose_chain_bwd": (_i, [_f, _f, _i, _f, _f, _f]),
    "cope_render_mlp_bwd": (_i, [_D, _f, _D, _f, _f, _f, _i, _i, _l, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _i, _f]),
    "cope_ray_points": (_i, [_f, _f, _f, _f, _f, _f, _i, _l, _i, _i, _f, _f, _f, _f]),
    "cope_ray_points_bwd": (_i, [_f, _f, _f, _l, _i, _f, _f, _f]),
    "cope_coarse_z": (_i, [_f, _f, _f, _l, _i, _f, _f]),
    "cope_sample_cdf": (_i, [_f, _f, _l, _i, _i, _f, _f, _f]),
    "cope_upsample": (_i, [_f, _f, _l, _i, _i, _fl, _f, _f, _f, _f]),
    "cope_merge_z": (_i, [_f, _f, _f, _f, _l, _i, _i, _f, _f, _f]),
    "cope_composite_fwd": (_i, [_f] * 8 + [_fl, _i, _l, _i] + [_f] * 8 + [_f]),
    "cope_composite_bwd": (_i, [_f] * 8 + [_fl, _i, _l, _i] + [_f] * 9 + [_f]),
    "cope_pose_fwd": (_i, [_f, _f, _f, _f, _f]),
    "cope_pose_bwd": (_i, [_f, _f, _f, _f, _f, _f, _f]),
    "cope_raygen_fwd": (_i, [_f, _f, _f, _f, _l, _f, _f, _f, _f]),
    "cope_raygen_bwd": (_i, [_f, _f, _f, _f, _l, _f, _f, _f, _f, _f, _f]),
    "cope_tc_pack": (_i, [_f, _i, _i, _i, _i, _i, _i, _f, _f]),
    "cope_tc_gemm": (_i, [_i, _i, _i, _f, _i, _f, _f, _i, _fl, _f, _i, _i, _f]),
    "cope_tc_wgrad_ws_floats": (_l, []),
    "cope_tc_wgrad": (_i, [_l, _i, _i, _i, _i, _f, _i, _f, _i, _f, _i, _f, _f]),
    "cope_step_losses_fwd": (_i, [_f] * 7 + [_l, _l, _fl, _fl, _fl, _f, _f, _f, _f]),
    "cope_step_losses_bwd": (_i, [_f] * 6 + [_l, _l] + [_f] * 6 + [_f]),
    "cope_weighted_points_fwd": (_i, [_f, _f, _l, _i, _f, _f]),
    "cope_weighted_points_bwd": (_i, [_f, _f, _f, _l, _i, _f, _f, _f]),
    "cope_flow_rgb_fwd": (_i, [_f] * 7 + [_l, _i, _i, _i, _f, _f, _f, _f]),
    "cope_flow_rgb_bwd": (_i, [_f] * 7 + [_l, _i, _i, _i, _f, _f, _f, _f, _f]),
    "cope_patch_smooth_fwd": (_i, [_f, _f, _l, _i, _fl, _fl, _fl, _f, _f, _f]),
    "cope_patch_smooth_bwd": (_i, [_f, _f, _l, _i, _fl, _fl, _fl, _f, _f, _f]),
    "cope_pose_refine_fwd": (_i, [_f, _f, _f, _f, _f, _i, _i, _i, _f, _f, _f, _f]),
    "cope_pose_refine_bwd": (_i, [_f, _f, _f, _f, _f, _i, _i, _i, _f, _f, _f, _f]),
    "cope_sample_pixels": (_i, [_f, C.c_uint64, _i, _i, _i, _i, _f, _f, _f, _f, _f, _f]),
    "cope_adam_step": (_i, [_f, _f, _f, _f, _l, _f, _fl, _fl, _fl, _fl, _fl, _f]),
    "cope_sgemm": (_i, [_i, _i, _i, _i, _i, _f, _i, _f, _i, _f, _i, _i, _f]),
}
EXPORTS = tuple(_SIGS)

_lib = None


class CopeError(RuntimeError):
    pass


def load():
    """Load the shared library (no GPU needed for this); raises loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CopeError(f"{LIB_PATH} is missing: run `python -m cope_nerf_b200.build`. "
                            "cope_nerf_b200 has no CPU / PyTorch fallback path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def ptr(t):
    """Marshal a tensor argument (None -> NULL).  Returns the (contiguous) TENSOR, not its address: `call` takes
    the address while holding a reference, so a temporary made by .contiguous()/.float() cannot be freed — and its
    block re-used by the next temporary — before the kernel is enqueued."""
    if t is None:
        return None
    if not t.is_cuda:
        raise CopeError("cope_nerf_b200 kernels need CUDA tensors (no CPU fallback); got a CPU tensor")
    return t if t.is_contiguous() else t.contiguous()


def stream():
    if not torch.cuda.is_available():
        raise CopeError("cope_nerf_b200 kernels need a CUDA device (sm_100a); there is no CPU fallback")
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    lib = load()
    keep = args                      # tensors stay referenced until the launch has been enqueued
    conv = []
    for a in args:
        if isinstance(a, torch.Tensor):
            if not (a.is_cuda and a.is_contiguous()):
                raise CopeError(f"{name}: tensor arguments must be contiguous CUDA tensors (use _lib.ptr)")
            conv.append(a.data_ptr())
        elif isinstance(a, C.Array):
            conv.append(C.cast(a, C.c_void_p))
        else:
            conv.append(a)
    rc = getattr(lib, name)(*conv)
    del keep
    if rc != 0:
        raise CopeError(f"{name} failed ({rc}): {lib.cope_last_error().decode()}")


def query(name, *args):
    n = getattr(load(), name)(*args)
    if n < 0:
        raise CopeError(f"{name} failed: {load().cope_last_error().decode()}")
    return n


_scratch = {}      # (device index, stream handle) -> [buffer, replayed_by_a_graph]
_retired = []      # buffers a captured CUDA graph still writes into: kept alive for the life of the process


def scratch(n_floats, device):
    """Grow-only fp32 scratch, ONE PER (device, stream): kernels on one stream reuse it in stream order, and two streams of
    one device (e.g. the per-GPU threads of a DataParallel-style caller, or a CUDA-graph capture stream next to eager work)
    never share it.  A buffer that was handed out while its stream was being captured is replayed into by that graph for as
    long as the graph lives, so it is never freed: when a later, larger request outgrows it, it is retired (kept
    referenced) and a new one is allocated, instead of being released under the graph."""
    if not torch.cuda.is_available():
        raise CopeError("cope_nerf_b200 kernels need a CUDA device (sm_100a); there is no CPU fallback")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)
    ent = _scratch.get(key)
    if ent is None or ent[0].numel() < n_floats:
        if ent is not None and ent[1]:
            _retired.append(ent[0])
        ent = [torch.empty(int(n_floats * 1.25) + 1024, dtype=torch.float32, device=device), False]
        _scratch[key] = ent
    if torch.cuda.is_current_stream_capturing():
        ent[1] = True
    return ent[0]
