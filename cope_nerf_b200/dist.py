"""Multi-GPU data parallelism over rays: one process per GPU, replicated weights, each rank renders its own ray
slice, ONE NCCL all-reduce of a flat fp32 gradient buffer per step.  Replaces the reference's single-process
nn.DataParallel wrap (train.py:54), which re-broadcasts all parameters every forward and gathers 15 output
tensors to GPU 0.

Host-side logic only (sharding arithmetic + the flat bucket); it is backend-agnostic so the N>1 path is covered
by world_size-2 gloo tests on CPU (tests/test_dist_cpu.py)."""
import os

import torch
import torch.distributed as dist

__all__ = ["init_from_env", "shard_range", "FlatGradBucket", "allreduce_scalar_"]


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).
    Returns (rank, world, local_rank).  world == 1 -> no process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_rays, rank, world, align=16):
    """Contiguous ray slice of `rank`.  Slices are multiples of `align` (= patch_size**2 = 16, so the 4x4 depth
    patches of train.py:519-525 never straddle ranks); the remainder goes to the last ranks one block at a time."""
    blocks = n_rays // align
    base, extra = divmod(blocks, world)
    counts = [(base + (1 if r >= world - extra else 0)) * align for r in range(world)]
    counts[-1] += n_rays - blocks * align
    start = sum(counts[:rank])
    return start, start + counts[rank]


class FlatGradBucket:
    """All parameter gradients live in ONE flat fp32 buffer (p.grad are views), so the data-parallel exchange is a
    single all-reduce of ~4 MB (1,003,155 floats for SDF + colour + variance + motion).

    Build it ONCE, before the training loop, and clear gradients with `bucket.zero_()` (or
    `optimizer.zero_grad(set_to_none=False)`) — NOT with the default `optimizer.zero_grad()`: set_to_none=True drops the
    views, autograd then allocates fresh `.grad` tensors outside the bucket and the all-reduce would exchange a stale
    buffer.  `allreduce_()` checks this and re-attaches (copying the stray gradients in) rather than exchanging garbage.
    Gradients that already exist at construction are copied into the bucket, not discarded.

    Parameters of a bucket are marked `_cope_grad_in_place`: the weight-norm backward of the network modules
    (fields._FlatWeights) then ACCUMULATES into the bucket's views inside its own kernel and reports no gradient to autograd,
    which would otherwise launch one elementwise add per parameter tensor (~80 per step).  Consequence: for such parameters use
    `loss.backward()` (not `torch.autograd.grad`, which would find no gradient), and gradient hooks on them do not fire."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_params = None
        self._attach()

    def _views(self):
        off = 0
        for p in self.params:
            yield p, self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def _attach(self):
        """Point every p.grad at its slice of the flat buffer; a gradient that lives elsewhere is copied in first.
        Returns the number of parameters that had to be re-attached."""
        moved = 0
        for p, view in self._views():
            g = p.grad
            if g is not None and g.data_ptr() == view.data_ptr() and g.shape == view.shape:
                continue
            with torch.no_grad():
                if g is not None:
                    view.copy_(g)
                else:
                    view.zero_()
            p.grad = view
            moved += 1
        for p in self.params:
            p._cope_grad_in_place = True      # fields._FlatWeights.backward adds straight into these views (no per-tensor add launch)
        return moved

    def zero_(self):
        self._attach()
        self.flat.zero_()

    def flatten_params_(self):
        """Move the PARAMETERS into one flat fp32 buffer laid out like the gradients (every `p.data` becomes a view of it, values
        kept), so that an optimiser can update all of them with one elementwise launch (optim.FlatAdam).  Do this before a CUDA
        graph of the step is captured: kernels captured earlier keep reading the old storage.  Idempotent."""
        if self.flat_params is None:
            if any(p.dtype != torch.float32 for p in self.params):
                raise ValueError("flatten_params_: the flat buffers are fp32; found a parameter of another dtype")
            fp = torch.empty_like(self.flat)
            off = 0
            with torch.no_grad():
                for p in self.params:
                    n = p.numel()
                    view = fp[off:off + n].view(p.shape)
                    view.copy_(p.data)
                    p.data = view
                    off += n
            self.flat_params = fp
        return self.flat_params

    def params_attached(self):
        """True while every parameter still is the view of the flat parameter buffer that flatten_params_ made it.  `module.half()`,
        `module.to(other_device)` or `p.data = ...` re-allocate the storage: an optimiser that updates the flat buffer would then
        silently stop training those parameters, so optim.FlatAdam checks this at every eager step."""
        if self.flat_params is None:
            return False
        base, off = self.flat_params.data_ptr(), 0
        for p in self.params:
            if p.data_ptr() != base + 4 * off or p.dtype != torch.float32:
                return False
            off += p.numel()
        return True

    def param_range(self, params):
        """[lo, hi) element range of `params` inside the flat buffers; they must be a contiguous run of the bucket's parameters,
        in the bucket's order (one optimiser / learning rate per run, as train.py:59-60 builds one Adam per network group)."""
        ids = [id(p) for p in self.params]
        want = [id(p) for p in params if p.requires_grad]
        if not want or want[0] not in ids:
            raise ValueError("param_range: parameters are not in this bucket")
        k = ids.index(want[0])
        if ids[k:k + len(want)] != want:
            raise ValueError("param_range: parameters are not a contiguous run of the bucket's parameters (same order)")
        lo = sum(p.numel() for p in self.params[:k])
        return lo, lo + sum(p.numel() for p in self.params[k:k + len(want)])

    def allreduce_(self, average=False):
        """Sum (or mean) over ranks, in place.  Local losses must already carry the 1/world factor of any
        shard-linear term (rgb: sum/N, eikonal: mean over points)."""
        self._attach()      # no-op when every p.grad is still the bucket's view (the normal case)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            if average:
                self.flat.div_(dist.get_world_size())
        return self.flat


def allreduce_scalar_(t):
    """Global normaliser for losses that divide by a batch-wide sum (sdf_loss / flow_rgb_loss, train.py:477,515)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
