"""Optimiser step of the training loop (train.py:59-60 builds `torch.optim.Adam` per parameter group, model/training.py:552-558
steps it) as ONE elementwise launch over flat buffers.

torch's fused multi-tensor Adam walks the ~80 parameter tensors of the SDF / colour / variance networks in 64 K-element chunks:
two launches of ~30 CTAs, ~50 us each (`profiles/r02_bench_launches_summary.txt`), 3.6 % of the 1024-ray training step for 28 MB
of traffic.  `FlatAdam` keeps parameters, gradients (dist.FlatGradBucket) and both moments in flat fp32 buffers and calls
`cope_adam_step` once (csrc/optim.cu).  Same update rule as `torch.optim.Adam` (no amsgrad); the step count lives on the device,
so the optimiser step can be captured into the CUDA graph of the training step like capturable Adam."""
import torch

from . import _lib as L

__all__ = ["FlatAdam"]


class FlatAdam(torch.optim.Optimizer):
    """Adam over a `dist.FlatGradBucket` (all of it, or the contiguous run `params` of its parameters — one FlatAdam per learning
    rate).  A `torch.optim.Optimizer`, so learning-rate schedulers attach to it; `param_groups[0]['lr']` is read at every
    `step()` (a step captured into a CUDA graph keeps the value it was captured with, like any host scalar).

    Construction moves the bucket's parameters into one flat buffer (`FlatGradBucket.flatten_params_`): build it before capturing
    a graph of the step.  There is no CPU path: `step()` raises `CopeError` without a CUDA device."""

    def __init__(self, bucket, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, params=None):
        plist = [p for p in (bucket.params if params is None else params) if p.requires_grad]
        super().__init__(plist, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise ValueError("FlatAdam: one parameter group per optimiser (build one FlatAdam per learning rate)")
        self.bucket = bucket
        flat_p = bucket.flatten_params_()
        lo, hi = (0, flat_p.numel()) if params is None else bucket.param_range(plist)
        self.range = (lo, hi)
        self.flat_param, self.flat_grad = flat_p[lo:hi], bucket.flat[lo:hi]
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self.step_t = torch.zeros(1, dtype=torch.float32, device=flat_p.device)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if not self.flat_param.is_cuda:
            raise L.CopeError("FlatAdam.step needs CUDA tensors (cope_adam_step; there is no CPU fallback)")
        if not self.bucket.params_attached():
            raise L.CopeError("FlatAdam.step: a parameter no longer lives in the bucket's flat buffer (module.half() / .to(device) / "
                              "p.data = ... after the optimiser was built?); rebuild the bucket and the optimiser")
        self.bucket._attach()       # gradients that strayed outside the bucket (optimizer.zero_grad(set_to_none=True)) come back first
        g = self.param_groups[0]
        self.step_t.add_(1.0)
        L.call("cope_adam_step", self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.flat_param.numel(), self.step_t,
               float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), L.stream())
        return loss

    def zero_grad(self, set_to_none=False):
        """Clears this optimiser's slice of the flat gradient buffer (the views stay attached; `set_to_none` is ignored on purpose)."""
        self.bucket._attach()
        self.flat_grad.zero_()

    def state_dict(self):
        return {"step": self.step_t.clone(), "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "range": self.range, "param_groups": [{k: v for k, v in self.param_groups[0].items() if k != "params"}]}

    def load_state_dict(self, sd):
        if tuple(sd["range"]) != tuple(self.range):
            raise ValueError(f"FlatAdam.load_state_dict: saved element range {sd['range']} != {self.range}")
        self.step_t.copy_(sd["step"]); self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.param_groups[0].update(sd["param_groups"][0])
