#!/usr/bin/env python
"""bench.py — cope-nerf NeuS training step (BASELINE.json metric: train rays/s, fwd + bwd + eikonal, 64+64 samples).

    python bench.py --gpus 1 --steps 10 --warmup 3                      # our arm
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...    # N ranks, rays sharded (weak scaling)
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1      # reference arm (CPU oracle port)

Workload (BASELINE.json configs[1]): Co3D-skateboard shapes — 717x1275 synthetic frame,
1024 rays per GPU in 4x4 patches, 64 coarse + 4x16 importance samples, SDF MLP 8x256 (PE L=6 on (x,y,z,t)),
colour MLP 4x256, random geometric init (seed 678), losses 0.33333*rgb_L1 + 0.1*eikonal, pose (r,t) trainable,
Adam step included.  One "step" = pose -> rays -> hierarchical sampling -> render -> loss -> backward ->
gradient all-reduce (N>1) -> optimiser.

`value`  : inputs (pixel coords, gt colours, jitter) already resident in HBM.
`e2e`    : the same step through the public module API with HOST inputs: per step pinned-host -> device copy of
           pixel coords + gt colours (+ the CPU-RNG jitter the reference draws, neus_renderer.py:482) and a
           device -> host read of the loss.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FLOP_PER_TRAIN_RAY = 1_117_315_072          # 112*F_sdfq + 128*(6*F_sdf + 3*F_col)
H, W = 717, 1275
DEPTH_RANGE = (0.01, 5.0)
METRIC = "train rays/s fwd+bwd+eikonal (NeuS 64+64 samples)"
MLP_CALLS = {"cope_sdf_query", "cope_sdf_fwd", "cope_sdf_bwd", "cope_color_fwd", "cope_color_bwd", "cope_render_mlp_fwd",
             "cope_render_mlp_bwd"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


def synth_inputs(n_rays, n_batches, seed):
    """Seeded synthetic frame + per-step ray batches (host tensors)."""
    from cope_nerf_b200 import training as T
    from cope_nerf_b200.common import pixels_from_indices
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(3, H, W, generator=g)
    flat = img.view(3, H * W).t().contiguous()
    torch.manual_seed(seed)
    batches = []
    for _ in range(n_batches):
        idx = T.get_patch_indices(H, W, 4, n_rays)
        batches.append(dict(pix=pixels_from_indices(idx, H, W).contiguous(), rgb=flat[idx].contiguous(),
                            t_rand=torch.rand(n_rays, 64)))
    return batches


def camera():
    f = 0.8 * W
    return torch.tensor([[2 * f / W, 0, 0, 0], [0, -2 * f / H, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]],
                        dtype=torch.float32).unsqueeze(0)


class ClockSampler:
    """SM clock + clock-event reasons of one GPU DURING a timed region.  In-process NVML thread (5 ms period: a 20-step region of
    60 ms still yields ~10 samples); `nvidia-smi -lms` as the fallback when NVML cannot be loaded (its start-up alone can outlast a
    short region, which is why it is only the fallback)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("sw_power_cap", 0x4))

    def __init__(self, gpu):
        self.gpu, self.proc, self.path = gpu, None, None
        self.nvml, self.handle, self.thread, self.samples, self.halt = None, None, None, [], None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:                 # the CUDA ordinal is not the NVML index under CUDA_VISIBLE_DEVICES: go through the UUID
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(gpu).uuid)).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu)
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _loop(self):
        nv, h = self.nvml, self.handle
        while not self.halt.is_set():
            try:
                self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                     int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))))
            except Exception:
                pass
            self.halt.wait(0.005)

    def start(self):
        if self.nvml is not None:
            import threading
            self.samples, self.halt = [], threading.Event()
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.thread is not None:
            self.halt.set()
            self.thread.join(timeout=2)
            self.thread = None
            sm = sorted(c for c, _ in self.samples)
            bits = 0
            for _, r in self.samples:
                bits |= r
            if sm:
                try:
                    mx = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
                except Exception:
                    mx = max(sm)
                out = dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=sorted(n for n, b in self.REASONS if bits & b),
                           samples=len(sm), source="nvml")
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1])); mx.append(float(f[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out = dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm), source="nvidia-smi")
        return out


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_train_rays_per_s(n_rays, steps, warmup, seed=678):
    """The oracle port of the reference's PyTorch CPU path on all host cores: same step, bounded ray sample."""
    import oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(seed)
    P = dict(sdf=O.init_sdf_params(**O.DEFAULT_CFG["sdf"]), color=O.init_color_params(**O.DEFAULT_CFG["color"]),
             variance=O.init_variance_params(**O.DEFAULT_CFG["variance"]))
    params = []
    for t in P.values():
        for k in t:
            t[k] = t[k].requires_grad_(True)
            params.append(t[k])
    pose = dict(r=(torch.randn(1, 3) * 0.05).requires_grad_(True), t=(torch.randn(1, 3) * 0.05).requires_grad_(True),
                init_c2w=torch.eye(4).unsqueeze(0))
    opt = torch.optim.Adam(params + [pose["r"], pose["t"]], lr=1e-3)
    batches = synth_inputs(n_rays, steps + warmup, seed)
    K, S = camera(), torch.eye(4).unsqueeze(0)
    times = []
    for i, b in enumerate(batches):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss, _ = O.train_step(P, pose, b["pix"], K, S, b["rgb"], torch.zeros(1), list(DEPTH_RANGE), cos_anneal=0.5,
                               t_rand=b["t_rand"])
        loss.backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times)
    return n_rays * len(times) / dt, dt / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_rays
    v, spt = cpu_train_rays_per_s(n, args.steps, args.warmup)
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": spt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.rays, args.gpus),
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"{n} of the {args.rays} rays x 64+64 samples per step, same nets/losses/Adam, oracle port of the "
                                   f"reference's PyTorch CPU path, torch threads = {cores}"},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_name(rays):
    return (f"Co3D-skateboard NeuS train step: {rays} rays/GPU (4x4 patches of a {H}x{W} synthetic frame) x 64+64 "
            "samples, up_sample 4 iters, SDF 8x256 PE6 4-D, colour 4x256, rgb+eikonal+pose grads, Adam")


def workload_config(rays, world):
    """`config` of the JSON line: the WORKLOAD only, identical for both arms (how an arm runs it is in `implementation` / `cpu_baseline`)."""
    return {"workload": workload_name(rays), "rays_per_gpu": rays,
            "parallelism": f"dp{world} (rays sharded over the GPUs, one flat gradient all-reduce per step)",
            "l2": "per-step working set (saved activations, >1 GB) exceeds the 126 MB L2; new ray batch every step"}


# ---------------------------------------------------------------------------------------------- GPU arm
def eager_cuda_rays_per_s(n_rays, steps, warmup, dev, seed=678):
    """Informative second baseline: the SAME oracle port of the reference's eager-PyTorch path, run on the
    B200 (`with torch.device(cuda)`: every factory call of the port lands on the GPU; ATen / cuBLAS fp32 kernels, autograd incl.
    the double backward, torch.optim.Adam) — what the reference's own code path does on this GPU, since it ships no kernel."""
    import oracle as O
    host = synth_inputs(n_rays, steps + warmup, seed)
    torch.manual_seed(seed)
    P0 = dict(sdf=O.init_sdf_params(**O.DEFAULT_CFG["sdf"]), color=O.init_color_params(**O.DEFAULT_CFG["color"]),
              variance=O.init_variance_params(**O.DEFAULT_CFG["variance"]))
    with torch.device(dev):
        P = {t: {k: v.to(dev) for k, v in d.items()} for t, d in P0.items()}
        params = []
        for t in P.values():
            for k in t:
                t[k] = t[k].requires_grad_(True)
                params.append(t[k])
        pose = dict(r=(torch.randn(1, 3) * 0.05).requires_grad_(True), t=(torch.randn(1, 3) * 0.05).requires_grad_(True),
                    init_c2w=torch.eye(4).unsqueeze(0))
        opt = torch.optim.Adam(params + [pose["r"], pose["t"]], lr=1e-3)
        batches = [{k: v.to(dev) for k, v in b.items()} for b in host]
        K, S, t0 = camera().to(dev), torch.eye(4).unsqueeze(0), torch.zeros(1)
        for i, b in enumerate(batches):
            if i == warmup:
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e0.record()
            opt.zero_grad()
            loss, _ = O.train_step(P, pose, b["pix"], K, S, b["rgb"], t0, list(DEPTH_RANGE), cos_anneal=0.5, t_rand=b["t_rand"])
            loss.backward()
            opt.step()
        e1 = torch.cuda.Event(enable_timing=True); e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return n_rays / (ms * 1e-3), ms


def mlp_traffic_from_profile():
    """DRAM bytes (read + write) of the MLP kernel group of ONE 1024-ray step, summed from the committed ncu --set full summary
    of this round (profiles/r02_step_kernels_ncu_full.csv, written by tools/ncu_summary.py from the .ncu-rep).  None when the
    file is absent: the number is a measurement, never a literal."""
    import csv
    path = os.path.join(ROOT, "profiles", "r02_step_kernels_ncu_full.csv")
    try:
        tot = 0.0
        with open(path) as f:
            for row in csv.DictReader(f):
                if row.get("mlp_group") == "1":
                    tot += float(row["dram_read_bytes"]) + float(row["dram_write_bytes"])
        return (tot / 1e9, os.path.relpath(path, ROOT)) if tot > 0 else (None, None)
    except Exception:
        return None, None


def stage1_step_ms(C, dev, prec, n_rays, steps, warmup, use_graph=True, optimizer="flat"):
    """The reference's WHOLE stage-1 iteration (train.py:407-532) on one GPU: on-device patch sampling + rays (process_data),
    fused render + rgb / eikonal / SDF-flow node, flow-RGB over the valid reference frames, SDF-consistency re-query (another
    131 072-point SDF forward + backward), depth-patch smoothness, backward into SDF / colour / variance / MotionNetwork, both
    Adam steps.  Synthetic 100-frame sequence, query frame 40, reference frames 41-43, world frame 50, 10 sub-steps per pair
    (configs/default.yaml).  The frame-dependent pieces go through losses.Stage1Static (frame indices as device tensors, one
    global pose chain), so the iteration is ONE CUDA graph that is valid for every frame.  Returns (ms per step, launch note)."""
    from cope_nerf_b200 import losses as CL
    torch.manual_seed(678)
    rnd = C.training.build_networks(device=dev, precision=prec)
    mot = C.MotionNetwork(d_out=6, d_in=1, d_hidden=256, n_layers=4, skip_in=[2], multires=6, bias=0.5, scale=1.0, geometric_init=False,
                          weight_norm=True).to(dev)
    g = torch.Generator().manual_seed(5)
    img = torch.rand(1, 3, H, W, generator=g).to(dev)
    refs = torch.rand(3, 3, H, W, generator=g).to(dev)
    Kc, Sc = camera().to(dev), torch.eye(4, device=dev).unsqueeze(0)
    Kr = Kc.repeat(3, 1, 1)
    world = torch.eye(4, device=dev)
    n_img, idx, ref_idx, world_idx, n_sub = 100, 40, [41, 42, 43], 50, 10
    st = CL.Stage1Static(mot, n_img, n_sub, world_idx, -1.0 + 2.0 * world_idx / (n_img - 1))
    w = dict(rgb=1.0, eik=0.1, sdf=0.1, frgb=7.5, cons=1.0, edge=1.0, smooth=1e-4)
    n_corner = (H - 3) * (W - 3)
    static = dict(corners=torch.zeros(n_rays // 16, dtype=torch.int64, device=dev), t_rand=torch.zeros(n_rays, 64, device=dev),
                  qts=torch.tensor([idx / (n_img - 1) * 2 - 1], device=dev), idx=torch.tensor([idx], device=dev),
                  ref=torch.tensor(ref_idx, device=dev), valid=torch.ones(3, device=dev), cons_on=torch.ones(1, device=dev))
    from cope_nerf_b200.dist import FlatGradBucket
    bucket = FlatGradBucket(list(rnd.parameters()) + list(mot.parameters()))      # one fill launch clears every gradient
    if optimizer == "flat":          # train.py:59-60: one Adam for the NeuS networks, one for the MotionNetwork
        opt = C.optim.FlatAdam(bucket, lr=1e-3, params=list(rnd.parameters()))
        mopt = C.optim.FlatAdam(bucket, lr=5e-4, params=list(mot.parameters()))
    else:
        opt = torch.optim.Adam(list(rnd.parameters()), lr=1e-3, fused=True, capturable=True)
        mopt = torch.optim.Adam(list(mot.parameters()), lr=5e-4, fused=True, capturable=True)

    def new_inputs():
        static["corners"].copy_(torch.randint(0, n_corner, (n_rays // 16,), device=dev))
        static["t_rand"].copy_(torch.rand(n_rays, 64, device=dev))

    def iteration():
        bucket.zero_()
        pix, npix, o, d, dn, rgb_gt, _ = C.training.process_data(img, Kc, world, Sc, n_rays, patch_size=4, corners=static["corners"])
        near, far = C.training.near_far_from_sphere(o, d, DEPTH_RANGE)
        rnd.t_rand_override = static["t_rand"]
        G, motion = st.global_chain(static["qts"])          # ONE MotionNetwork call: pose chain of the sequence + the query-time motion
        loss, out = rnd.forward_losses(o, d, dn, static["qts"], near, far, rgb_gt, cos_anneal_ratio=0.5, it=1, rgb_weight=w["rgb"],
                                       eikonal_weight=w["eik"], sdf_weight=w["sdf"], motion=motion)
        aux = st.losses(out, rgb_gt, rnd.sdf_network, static["idx"], static["ref"], static["valid"], static["cons_on"], Kr, Sc, npix, pix,
                        refs, consistency_pose_grad=False, G=G)
        sm, _ = CL.depth_smoothness_losses(out["depth_pred"], rgb_gt, 4, edge_weight=w["edge"], smooth_weight=w["smooth"])
        total = loss + w["frgb"] * aux["flow_rgb_loss"] + w["cons"] * aux["sdf_consistency_loss"] + sm
        total.backward()
        opt.step(); mopt.step()
        return total.detach()

    graph, note = None, "eager launches"
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            new_inputs(); iteration()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    if use_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                iteration()
            note = "one CUDA graph per stage-1 iteration (valid for every frame: indices are device tensors)"
        except Exception as e:      # pragma: no cover
            graph, note = None, f"eager launches (graph capture failed: {type(e).__name__}: {str(e)[:100]})"
            torch.cuda.synchronize()

    def step():
        new_inputs()
        graph.replay() if graph is not None else iteration()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, note


class Runner:
    """One configuration of the training step on this rank: networks, optimiser, synthetic batches, and the step as ONE CUDA
    graph (pose -> rays -> sampling -> render -> loss -> backward -> gradient all-reduce -> Adam step: optim.FlatAdam by default)."""

    def __init__(self, C, dev, prec, n_rays, world, rank, n_batches, loss_scale, graph_mode, optimizer="flat"):
        from cope_nerf_b200.dist import FlatGradBucket
        self.C, self.dev, self.n, self.world = C, dev, n_rays, world
        torch.manual_seed(678)
        self.rnd = C.training.build_networks(device=dev, precision=prec)
        self.pose = C.PoseRetriever(1).to(dev)
        with torch.no_grad():
            self.pose.r.copy_(torch.randn(1, 3) * 0.05); self.pose.t.copy_(torch.randn(1, 3) * 0.05)
        self.bucket = FlatGradBucket(list(self.rnd.parameters()) + [self.pose.r, self.pose.t])
        if optimizer == "flat":      # ONE elementwise launch over the flat parameter / gradient / moment buffers (cope_adam_step)
            self.opt = C.optim.FlatAdam(self.bucket, lr=1e-3)
        else:                        # torch's fused multi-tensor Adam: two ~50 us launches over the ~80 parameter tensors
            self.opt = torch.optim.Adam(self.bucket.params, lr=1e-3, fused=True, capturable=True)
        self.Kc, self.Sc = camera().to(dev), torch.eye(4, device=dev).unsqueeze(0)
        self.tstep = torch.zeros(1, device=dev)
        host = synth_inputs(n_rays, n_batches, seed=678 + 1000 * rank)
        self.pinned = [{k: v.pin_memory() for k, v in b.items()} for b in host]
        self.resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
        self.loss_scale = loss_scale
        self.graph, self.static, self.static_loss, self.tail_eager = None, None, None, False
        self.note = "eager launches"
        if graph_mode != "none":
            self._capture(graph_mode)

    def compute(self, b):
        """zero grads -> pose -> rays -> sampling -> render -> loss -> backward"""
        self.bucket.zero_()
        self.rnd.t_rand_override = b["t_rand"]
        loss, _, _ = self.C.training.render_train_step(self.rnd, self.pose, 0, b["pix"], self.Kc, self.Sc, b["rgb"], self.tstep,
                                                       DEPTH_RANGE, cos_anneal_ratio=0.5, it=1, loss_scale=self.loss_scale)
        return loss

    def exchange_and_update(self):
        self.bucket.allreduce_()
        self.opt.step()

    def step_eager(self, b):
        loss = self.compute(b)
        self.exchange_and_update()
        return loss

    def _capture(self, mode):
        """mode 'full': the whole step incl. the gradient exchange and the Adam step (device-resident step count) in one graph - used on ONE GPU, where
        the exchange is a no-op; mode 'compute' (N > 1): forward + backward as one graph, NCCL all-reduce and Adam launched eagerly
        behind the replay on the same stream.  Measured at 8 GPUs: 'compute' 2.92 ms per step; with the all-reduce captured inside the
        graph one run took 3.95 ms per step and the next one hung in the replayed collective, so NCCL is never captured by default."""
        self.static = {k: torch.empty_like(v) for k, v in self.resident[0].items()}
        for attempt in (("full", "compute") if mode == "full" else ("compute",)):
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for b in self.resident[:3]:
                        for k in self.static:
                            self.static[k].copy_(b[k])
                        self.step_eager(self.static)          # also creates Adam's state before the capture
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    loss = self.compute(self.static)
                    if attempt == "full":
                        self.exchange_and_update()
                self.graph, self.static_loss, self.tail_eager = g, loss, attempt != "full"
                self.note = ("one CUDA graph per step (fwd + bwd + gradient all-reduce + Adam step)" if attempt == "full" else
                             "one CUDA graph per step (fwd + bwd), eager all-reduce + Adam step")
                return
            except Exception as e:      # pragma: no cover - depends on the box
                self.graph = None
                self.note = f"eager launches (graph capture '{attempt}' failed: {type(e).__name__}: {str(e)[:100]})"
                torch.cuda.synchronize()

    def step(self, b):
        if self.graph is None:
            return self.step_eager(b)
        for k in self.static:
            self.static[k].copy_(b[k], non_blocking=True)
        self.graph.replay()
        if self.tail_eager:
            self.exchange_and_update()
        return self.static_loss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=1024, help="rays per GPU per step")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling (BASELINE.json configs[2]): --rays is the TOTAL batch, split over the ranks in 16-ray blocks")
    ap.add_argument("--precision", default=os.environ.get("COPE_PRECISION", "auto"), choices=["auto", "fp32", "bf16"])
    ap.add_argument("--cpu-rays", type=int, default=512, help="ray sample of the CPU baseline step (BASELINE.json configs[0]: 512 rays)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the strong-scaling / ray-sweep / eager-CUDA extras of the JSON line")
    ap.add_argument("--optimizer", default="flat", choices=["flat", "torch"],
                    help="flat: optim.FlatAdam = one cope_adam_step launch over flat buffers; torch: torch.optim.Adam(fused=True, capturable=True)")
    ap.add_argument("--no-sweep", action="store_true", help="skip only the 4K-32K ray sweep of the extras")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--graph-mode", default="auto", choices=["auto", "full", "compute", "none"],
                    help="auto: the whole step (incl. fused Adam) as one graph on one GPU; forward + backward as one graph with the NCCL "
                         "all-reduce and Adam launched eagerly behind it on N > 1 (NCCL captured inside a replayed graph stalled / hung at 8 ranks)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)
    if args.no_graph:
        args.graph_mode = "none"

    import torch.distributed as dist
    import cope_nerf_b200 as C
    from cope_nerf_b200 import _lib as L
    from cope_nerf_b200.dist import init_from_env, shard_range
    assert torch.cuda.is_available(), "bench.py (our arm) needs a CUDA device; there is no CPU fallback"
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):   # NCCL prints its version banner on stdout (env or nccl.conf):
        os.environ["NCCL_DEBUG"] = "WARN"                             # keep stdout to the one JSON line of the contract
    # NCCL prints its version banner on the C-level stdout when the communicator is created: point fd 1 at stderr until the first
    # collective has run, so that stdout carries exactly the one JSON line of the contract
    sys.stdout.flush()
    fd1 = os.dup(1)
    os.dup2(2, 1)
    try:
        rank, world, local = init_from_env("nccl")
        torch.cuda.set_device(local)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(fd1, 1)
        os.close(fd1)
    dev = torch.device("cuda", local)
    lib = C.load_library()
    prec = {"fp32": C.PREC_FP32, "bf16": C.PREC_BF16}.get(args.precision)
    if prec is None:
        prec = getattr(C, "DEFAULT_PRECISION", C.PREC_FP32)
    n = args.rays
    if args.strong and world > 1:
        lo, hi = shard_range(args.rays, rank, world)      # this rank's slice of the total batch (multiples of one 4x4 patch)
        n = hi - lo
    if args.graph_mode == "auto":
        args.graph_mode = "full" if world == 1 else "compute"
    total_rays = args.rays if (args.strong and world > 1) else n * world
    # rgb (sum/N) and eikonal (mean) are shard-linear: scale the local loss by this rank's share of the batch
    run = Runner(C, dev, prec, n, world, rank, args.steps + args.warmup, n / total_rays, args.graph_mode, args.optimizer)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(fn, batches, steps, warmup, sampler=None):
        for b in batches[:warmup]:
            fn(b)
        barrier()
        if sampler is not None:      # clocks are sampled between the two barriers, i.e. while the timed steps run on the GPU
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for b in batches[warmup:warmup + steps]:
            fn(b)
        e1.record()
        barrier()
        if sampler is not None:
            sampler.result = sampler.stop()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- value: device-resident inputs
    clocks = ClockSampler(local)
    ms_total = timed_region(run.step, run.resident, args.steps, args.warmup, sampler=clocks)
    clk = clocks.result
    ms_per_step = ms_total / args.steps
    value = total_rays * args.steps / (ms_total * 1e-3)
    # launches of OUR kernels per step: kernels inside a replayed graph do not pass through the library's launch counter, so one
    # eager step is counted (the graph replays exactly these launches)
    l0 = lib.cope_launch_count()
    run.step_eager(run.resident[0])
    launches = (lib.cope_launch_count() - l0) * args.steps

    # ---- roofline of the MLP kernel group: CUDA events around the cope_sdf_* / cope_color_* calls on the launching
    # stream, over the same K batches launched eagerly (events cannot be recorded inside a replayed graph)
    mlp_events = []
    real_call = L.call

    def timed_call(name, *a):
        if name in MLP_CALLS:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            real_call(name, *a)
            e1.record()
            mlp_events.append((e0, e1))
        else:
            real_call(name, *a)

    L.call = timed_call
    torch.cuda.synchronize()
    for b in run.resident[args.warmup:]:
        run.step_eager(b)
    torch.cuda.synchronize()
    L.call = real_call
    mlp_ms = sum(a.elapsed_time(b) for a, b in mlp_events)

    # ---- e2e: host inputs every step, loss read back every step
    def step_e2e(b):
        if run.graph is None:
            dev_b = {k: v.to(dev, non_blocking=True) for k, v in b.items()}
            return run.step_eager(dev_b).item()
        return run.step(b).item()        # pinned host -> static device buffers -> graph replay -> loss read-back

    ms_e2e = timed_region(step_e2e, run.pinned, args.steps, args.warmup)
    e2e = total_rays * args.steps / (ms_e2e * 1e-3)
    h2d = sum(v.numel() * v.element_size() for v in run.pinned[0].values())
    graph_note = run.note

    # ---- extras recorded in the same line (BASELINE.json configs[2] strong split, configs[4] ray sweep): every rank takes part
    extras = {}
    if not args.no_extras and not args.strong:
        del run
        torch.cuda.empty_cache()
        if world > 1:
            lo, hi = shard_range(1024, rank, world)
            rs = Runner(C, dev, prec, hi - lo, world, rank, 13, (hi - lo) / 1024.0, args.graph_mode, args.optimizer)
            ms = timed_region(rs.step, rs.resident, 10, 3)
            extras["strong"] = {"rays_total": 1024, "rays_per_gpu": hi - lo, "ms_per_step": ms / 10, "value": 1024 * 10 / (ms * 1e-3),
                                "unit": "rays/s", "launch": rs.note,
                                "note": "ONE 1024-ray batch split over the ranks in 16-ray blocks (configs[2]); latency-bound below ~1 wave of 128-point tiles per GPU"}
            del rs
            torch.cuda.empty_cache()
        sweep = []
        for nr in (() if args.no_sweep else (4096, 16384, 32768)):
            L._scratch.clear()                   # the grow-only per-stream scratch of the previous size
            torch.cuda.empty_cache()
            need = nr * 128 * 36 * 1024          # ~28 kB of saved stacks + workspace per sample point, 1.25x scratch slack
            free = torch.cuda.mem_get_info(dev)[0]
            fits = torch.tensor([1 if need <= 0.85 * free else 0], device=dev)
            if world > 1:                        # ONE decision for all ranks: a rank that skipped would leave the others in a collective
                dist.all_reduce(fits, op=dist.ReduceOp.MIN)
            if int(fits.item()) == 0:            # never walk into an out-of-memory condition on the box
                sweep.append({"rays_per_gpu": nr, "skipped": f"needs ~{need / 2**30:.0f} GiB of saved activation stacks, {free / 2**30:.0f} GiB free"})
                continue
            try:
                rw = Runner(C, dev, prec, nr, world, rank, 5, 1.0 / world, "none", args.optimizer)
                ms = timed_region(rw.step, rw.resident, 3, 2)
                sweep.append({"rays_per_gpu": nr, "ms_per_step": ms / 3, "value": nr * world * 3 / (ms * 1e-3),
                              "mlp_frac_of_sustained_bf16_upper_bound": nr * FLOP_PER_TRAIN_RAY / (ms / 3 * 1e-3) / 1e12 / peaks()["bf16_sustained"]})
                del rw
            except Exception as e:      # pragma: no cover - memory of the box
                sweep.append({"rays_per_gpu": nr, "error": f"{type(e).__name__}: {str(e)[:80]}"})
            torch.cuda.empty_cache()
        if not args.no_sweep:
            extras["sweep"] = {"unit": "rays/s (whole job)", "launch": "eager", "points": sweep,
                               "note": "training steps at 4K-32K rays per GPU x 128 samples (configs[4]; 32K rays/GPU = 256K rays at 8 GPUs); "
                                       "the fraction counts the WHOLE step against the MLP FLOPs, so it is a lower bound of the MLP group's own"}

    if rank == 0:
        pk = peaks()
        mlp_ms_step = mlp_ms / args.steps if args.steps else 0.0
        ach = (n * FLOP_PER_TRAIN_RAY / (mlp_ms_step * 1e-3) / 1e12) if mlp_ms_step > 0 else None
        peak = pk["bf16_sustained"]
        traffic, traffic_src = mlp_traffic_from_profile()
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak",
            "vs_baseline": None, "dtype": "f32" if prec == C.PREC_FP32 else "bf16", "data": "synthetic",
            "config": workload_config(n, world),
            "implementation": {"precision": "fp32 SIMT (strict parity)" if prec == C.PREC_FP32 else "bf16 tcgen05, fp32 accumulate",
                               "launch": graph_note,
                               "optimizer": "optim.FlatAdam (cope_adam_step, one launch over flat buffers)" if args.optimizer == "flat"
                               else "torch.optim.Adam(fused=True, capturable=True)"},
            "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"bound": "tensor", "kernel": "SDF+colour MLP kernels (cope_sdf_query, cope_render_mlp_fwd/bwd = sdf_chain_query / sdf_fused<FWD,TAN,ADJ> / color_fused / tc_wgrad)",
                         "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": (ach / peak) if ach else None,
                         "traffic": traffic, "traffic_unit": "GB of DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) per step of the MLP kernel group at 1024 rays",
                         "traffic_source": traffic_src,
                         "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({pk['src']})",
                         "measured": "CUDA events around the MLP entry points, same K batches launched eagerly after the timed region",
                         "mlp_ms_per_step": mlp_ms_step, "mlp_share_of_step": min(1.0, mlp_ms_step / ms_per_step) if ms_per_step else None,
                         "algorithmic_flop_per_ray": FLOP_PER_TRAIN_RAY},
        }
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            v, spt = cpu_train_rays_per_s(args.cpu_rays, 5, 3)
            cores = os.cpu_count() or 1
            line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": cores, "kind": "port",
                                    "sample": f"{args.cpu_rays} rays x 64+64 samples per step (3 warm-up + 5 timed), same "
                                              f"nets/losses/Adam, oracle port of the reference's PyTorch CPU path"}
        if world == 1 and not args.no_extras:
            try:
                ms1, note1 = stage1_step_ms(C, dev, prec, 1024, 10, 3, optimizer=args.optimizer)
                line["stage1"] = {"ms_per_step": ms1, "value": 1024 / (ms1 * 1e-3), "unit": "rays/s", "launch": note1,
                                  "note": "the reference's whole stage-1 iteration (train.py:407-532): render + rgb / eikonal / SDF-flow + flow-RGB (3 "
                                          "reference frames) + SDF-consistency re-query + depth-patch smoothness + MotionNetwork, both Adam steps"}
            except Exception as e:      # pragma: no cover
                line["stage1"] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
            try:
                v, ms = eager_cuda_rays_per_s(1024, 5, 3, dev)
                line["eager_cuda_baseline"] = {"value": v, "unit": "rays/s", "ms_per_step": ms, "kind": "port on cuda",
                                               "sample": "1024 rays x 64+64 samples per step (3 warm-up + 5 timed), the oracle port of the reference's "
                                                         "eager PyTorch path (ATen / cuBLAS fp32, autograd double backward, Adam) on this B200; informative"}
            except Exception as e:      # pragma: no cover
                line["eager_cuda_baseline"] = {"error": f"{type(e).__name__}: {str(e)[:120]}"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
