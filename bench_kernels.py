#!/usr/bin/env python
"""Per-kernel roofline sweep (BASELINE.json configs[4]): 4K-256K rays x 128 samples on one B200.

For each HBM-bound kernel: achieved GB/s = algorithmic bytes (DESIGN.md §4 / SURVEY.md §8d) / CUDA-event time, against
the measured copy bandwidth in MEASURED_PEAKS.json.  For the fused SDF query chain: TFLOP/s against the bf16 peak.
Prints one JSON line per (kernel, rays).  Inputs are larger than L2 from 64K rays up; smaller sizes are L2-resident
and flagged so."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import cope_nerf_b200 as C  # noqa: E402
from cope_nerf_b200 import _lib as L  # noqa: E402
from bench import peaks  # noqa: E402


ITERS, WARM = None, None     # --iters / --warm: short runs for the ncu pass


def timeit(fn, iters=20, warm=3):
    iters, warm = (ITERS or iters), (WARM if WARM is not None else warm)
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def eval_image(rnd, pk, dev, h, w, chunk):
    """BASELINE.json configs[3]: full-image evaluation render (rgb + depth + normals, no backward) of an h x w frame,
    ScanNet shapes (484 x 648, configs/Scannet/scene0079_00.yaml:11-12), through training.render_image."""
    from bench import camera
    f = 0.8 * w
    K = torch.tensor([[2 * f / w, 0, 0, 0], [0, -2 * f / h, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]], dtype=torch.float32, device=dev)[None]
    world = torch.eye(4, device=dev)
    world[:3, 3] = torch.tensor([0.02, -0.03, 0.05], device=dev)
    S = torch.eye(4, device=dev)[None]
    t0 = torch.zeros(1, device=dev)
    # multi-GPU (torchrun): contiguous pixel ranges, one per rank; device time = max over ranks
    import torch.distributed as dist
    nrank = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    first, last = rank * (h * w) // nrank, (rank + 1) * (h * w) // nrank
    fn = lambda: C.training.render_image(rnd, world, K, S, h, w, t0, (0.01, 5.0), chunk=(chunk or None), rays=(first, last - first))
    if nrank > 1:
        dist.barrier()
    t = timeit(fn, iters=3, warm=1)
    if nrank > 1:
        tt = torch.tensor([t], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = tt.item()
        if rank != 0:
            return
    rays = h * w
    flop = rays * 440_983_552          # 112 F_sdfq + 128 (2 F_sdf + F_col) per evaluation ray
    print(json.dumps({"kernel": "eval_image_render", "n_gpus": nrank, "height": h, "width": w, "rays": rays, "chunk": chunk, "ms": t * 1e3,
                      "rays_per_s": rays / t, "achieved": flop / t / 1e12, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                      "frac": flop / t / 1e12 / (pk["bf16_sustained"] * nrank), "bound": "tensor", "peak_source": pk["src"],
                      "note": "pose -> rays -> 64+64 hierarchical sampling -> SDF value + gradient -> colour -> compositing -> "
                              "normal / arg-max-depth maps; every result stays on the device"}))


def stage1_losses_row(rnd, dev, n_rays=1024, S=128, T=3, h=717, w=1275):
    """the stage-1 auxiliary losses of train.py:467-517 at the training shapes (1024 rays x 128 samples, three
    717 x 1275 reference frames, 100-frame sequence), forward + backward through losses.stage1_losses: (a) the fused kernels alone
    (SDF-flow + flow-RGB incl. the MotionNetwork pose chain), (b) with the SDF-consistency re-query (131 072-point SDF forward +
    backward)."""
    from cope_nerf_b200 import losses as CL
    from cope_nerf_b200.motion import MotionNetwork
    from cope_nerf_b200.renderer import RenderOutputs
    torch.manual_seed(11)
    mot = MotionNetwork(d_out=6, d_in=1, d_hidden=256, n_layers=4, skip_in=[2], multires=6, bias=0.5, scale=1.0, geometric_init=False,
                        weight_norm=True).to(dev)
    P = n_rays * S
    pts4 = torch.cat([torch.randn(P, 3, device=dev) * 0.4 - torch.tensor([0, 0, 2.0], device=dev), torch.zeros(P, 1, device=dev)], -1)
    out = RenderOutputs({"color_fine": torch.rand(n_rays, 3, device=dev),
                         "weights": (torch.softmax(torch.randn(n_rays, S, device=dev), -1) * 0.9).requires_grad_(True),
                         "sdf": (torch.randn(P, 1, device=dev) * 0.2).requires_grad_(True)})
    out.grad4, out.pts4 = torch.randn(P, 4, device=dev).requires_grad_(True), pts4.requires_grad_(True)
    f = 0.8 * w
    K = torch.tensor([[2 * f / w, 0, 0, 0], [0, -2 * f / h, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]], dtype=torch.float32, device=dev)
    Kr, Sc = K[None].repeat(T, 1, 1), torch.eye(4, device=dev)[None]
    pix = torch.stack([torch.randint(0, w, (n_rays,), device=dev), torch.randint(0, h, (n_rays,), device=dev)], -1).float()
    npix = torch.stack([2 * pix[:, 0] / (w - 1) - 1, 2 * pix[:, 1] / (h - 1) - 1], -1)
    refs, gt = torch.rand(T, 3, h, w, device=dev), torch.rand(n_rays, 3, device=dev)

    def run(consistency):
        for t in (out["weights"], out["sdf"], out.grad4, out.pts4):
            t.grad = None
        mot.zero_grad(); rnd.sdf_network.zero_grad()
        res = CL.stage1_losses(out, gt, mot, rnd.sdf_network, -0.3, 40, [41, 42, 43], T, 100, 10, Kr, Sc, npix, pix, refs, 39, -0.2,
                               use_consistency=consistency)
        (res["sdf_loss"] + res["flow_rgb_loss"] + res["sdf_consistency_loss"]).backward()
    for name, cons in (("stage1_sdfflow_flowrgb_fwd_bwd", False), ("stage1_all_fwd_bwd", True)):
        t = timeit(lambda: run(cons), iters=10, warm=2)
        print(json.dumps({"kernel": name, "rays": n_rays, "samples": S, "ref_frames": T, "ms": t * 1e3,
                          "note": "losses.stage1_losses forward + backward: MotionNetwork pose chain (3 + 1 relative poses x 10 sub-steps), "
                                  "cope_step_losses / cope_weighted_points / cope_flow_rgb" +
                                  (", SDF-consistency re-query of 131072 points (cope_sdf_fwd / cope_sdf_bwd, bf16 layer-by-layer path)"
                                   if cons else "")}))


def motion_chain(dev, n_img, n_sub):
    """relative poses of all n_img - 1 consecutive frame pairs (n_sub sub-steps each) chained into
    world -> camera maps, forward + backward to the MotionNetwork parameters.  (The CPU port of the reference's Python
    double loop is timed by tests/time_motion_oracle.py — the oracle is test infrastructure and is not imported here.)"""
    from cope_nerf_b200.motion import MotionNetwork
    torch.manual_seed(5)
    m = MotionNetwork(d_out=6, d_in=1, d_hidden=256, n_layers=4, skip_in=[2], multires=6, bias=0.5, scale=1.0,
                      geometric_init=False, weight_norm=True).to(dev)          # configs/default.yaml:113-123
    wgt = torch.randn(n_img, 4, 4, device=dev)

    def ours():
        m.zero_grad()
        _, rel = m.compute_relative_camera_pose(0, n_img - 1, n_img, n_sub)
        (m.compute_w2c_mappings(rel) * wgt).sum().backward()
    t = timeit(ours, iters=10, warm=2)
    print(json.dumps({"kernel": "motion_pose_chain_fwd_bwd", "frames": n_img, "sub_steps": n_sub, "ms": t * 1e3,
                      "pairs_per_s": (n_img - 1) / t,
                      "note": "one batched MotionNetwork call (fp32 SIMT GEMMs, LeakyReLU) + cope_pose_integrate + cope_pose_chain, "
                              "forward and backward; latency-bound (3x3 / 4x4 fp32 arithmetic), no roofline"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, nargs="*", default=[4096, 16384, 65536, 262144])
    ap.add_argument("--eval-image", type=int, nargs=2, default=[484, 648], metavar=("H", "W"))
    ap.add_argument("--eval-chunk", type=int, default=16384, help="rays per pass of the evaluation render; 0 = as few passes as fit the free memory")
    ap.add_argument("--eval-only", action="store_true", help="only the evaluation-image row (the one that runs under torchrun)")
    ap.add_argument("--hbm-only", action="store_true", help="only the HBM-bound kernels (sampling, compositing, losses): the ncu pass")
    ap.add_argument("--iters", type=int, default=0)
    ap.add_argument("--warm", type=int, default=-1)
    args = ap.parse_args()
    global ITERS, WARM
    ITERS, WARM = (args.iters or None), (args.warm if args.warm >= 0 else None)
    if args.hbm_only:
        args.eval_image = [0, 0]
    from cope_nerf_b200.dist import init_from_env
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    rank, nrank, local = init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pk = peaks()
    torch.manual_seed(678)
    rnd = C.training.build_networks(device=dev, precision=C.PREC_BF16)
    if args.eval_image[0] > 0:
        eval_image(rnd, pk, dev, args.eval_image[0], args.eval_image[1], args.eval_chunk)
        if args.eval_only or nrank > 1:
            return
        motion_chain(dev, 100, 10)
        stage1_losses_row(rnd, dev)
    S = 128
    for N in args.rays:
        P = N * S
        z = torch.sort(torch.rand(N, S, device=dev) * 4.9 + 0.01, dim=-1)[0].contiguous()
        dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full((N, 1), 0.078, device=dev)], -1).contiguous()
        sdf = ((1.5 - z) * 0.7 + 0.02 * torch.randn(N, S, device=dev)).reshape(P, 1).contiguous()
        grad = torch.randn(P, 4, device=dev)
        rgb = torch.rand(P, 3, device=dev)
        rays_d = torch.nn.functional.normalize(torch.randn(N, 3, device=dev), dim=-1)
        dn = torch.ones(N, 1, device=dev)
        var = rnd.deviation_network.variance.detach()
        f = lambda *s: torch.empty(*s, device=dev)
        weights, color, depth, wz, cdf, wsum, wmax, inv_s = f(N, S), f(N, 3), f(N, 1), f(N, 1), f(N, S), f(N, 1), f(N, 1), f(1)
        st = L.stream()

        def comp_fwd():
            L.call("cope_composite_fwd", sdf, grad, rgb, z, dists, rays_d, dn, var, 0.5, 0, N, S, weights, color, depth, wz,
                   None, wsum, wmax, inv_s, st)
        d_color, d_depth, d_w = torch.rand(N, 3, device=dev), torch.rand(N, 1, device=dev), torch.rand(N, S, device=dev)
        d_sdf, d_grad, d_rgb, d_var, d_rd = f(P, 1), torch.zeros(P, 4, device=dev), f(P, 3), torch.zeros(1, device=dev), f(N, 3)

        def comp_bwd():
            L.call("cope_composite_bwd", sdf, grad, rgb, z, dists, rays_d, dn, var, 0.5, 0, N, S, d_color, d_depth, d_w,
                   None, d_sdf, d_grad, d_rgb, d_var, d_rd, st)
        new_z = f(N, 16)

        def ups():
            L.call("cope_upsample", z, sdf, N, S - 16, 16, 512.0, new_z, None, None, st)
        z112, s112 = z[:, :112].contiguous(), sdf.reshape(N, S)[:, :112].contiguous()
        zo, so = f(N, S), f(N, S)

        def mrg():
            L.call("cope_merge_z", z112, new_z, s112, new_z, N, 112, 16, zo, so, st)
        x = torch.cat([torch.randn(P, 3, device=dev) * 0.6, torch.zeros(P, 1, device=dev)], -1)
        flat = rnd.sdf_network.flat_weights().detach()
        # render-MLP stage through the C ABI (fused chains + batched weight gradients): SDF value + gradient + colour,
        # forward and backward incl. the double backward, 6 F_sdf + 3 F_col FLOP per point
        mlp = None
        if P <= 16384 * 128 and not args.hbm_only:
            sn, cn = rnd.sdf_network, rnd.color_network
            cflat = cn.flat_weights().detach()
            o_sdf, o_grad, o_rgb = f(P, 1), f(P, 4), f(P, 3)
            sdf_saved = f(L.query("cope_sdf_saved_floats", sn.desc, P, 1, C.PREC_BF16))
            col_saved = f(L.query("cope_color_saved_floats", cn.desc, P, C.PREC_BF16))
            ws = f(L.query("cope_render_mlp_ws_floats", sn.desc, cn.desc, P, C.PREC_BF16))
            dirs_pp = torch.nn.functional.normalize(torch.randn(N, 3, device=dev), dim=-1)
            g_sdf, g_grad0, g_rgb = torch.randn(P, 1, device=dev) * 1e-3, torch.randn(P, 4, device=dev) * 1e-3, torch.randn(P, 3, device=dev) * 1e-3
            dWs, dWc, dx, ddirs, g_grad = torch.zeros_like(flat), torch.zeros_like(cflat), torch.zeros(P, 4, device=dev), f(P, 3), f(P, 4)

            def mlp():
                L.call("cope_render_mlp_fwd", sn.desc, flat, cn.desc, cflat, x, dirs_pp, S, cn.multires_view, P, o_sdf, o_grad, o_rgb,
                       sdf_saved, col_saved, ws, C.PREC_BF16, st)
                g_grad.copy_(g_grad0)
                L.call("cope_render_mlp_bwd", sn.desc, flat, cn.desc, cflat, x, dirs_pp, S, cn.multires_view, P, sdf_saved, col_saved,
                       g_sdf, g_grad, g_rgb, dWs, dWc, dx, ddirs, ws, C.PREC_BF16, st)
        # loss reductions (cope_step_losses_*: rgb L1 + eikonal + SDF-flow; cope_weighted_points_fwd: flow-RGB's per-sample part)
        gt, mot6 = torch.rand(N, 3, device=dev), torch.randn(6, device=dev)
        l_out, l_coef, l_ws, l_g = f(4), f(4), f(8), torch.ones(1, device=dev)
        dl_color, dl_grad, dl_pts, dl_mot, wpts = f(N, 3), f(P, 4), f(P, 4), torch.zeros(6, device=dev), f(N, 4)

        def loss_fwd():
            L.call("cope_step_losses_fwd", color, gt, grad, x, weights, mot6, None, N, P, 0.33333, 0.1, 0.1, l_out, l_coef, l_ws, st)

        def loss_bwd():
            L.call("cope_step_losses_bwd", color, gt, grad, x, weights, mot6, N, P, l_coef, l_g, dl_color, dl_grad, dl_pts, dl_mot, st)

        def wpts_fwd():
            L.call("cope_weighted_points_fwd", weights, x, N, S, wpts, st)
        rows = [
            ("step_losses_fwd", loss_fwd, P * 36 + N * 24, "hbm"),               # grad4 16 + pts4 16 + w 4 per sample in
            ("step_losses_bwd", loss_bwd, P * 68 + N * 36, "hbm"),               # + d_grad4 16 + d_pts4 16 out
            ("weighted_points_fwd", wpts_fwd, P * 20 + N * 16, "hbm"),
            ("composite_fwd", comp_fwd, N * (S * 36 + 44), "hbm"),                # z,dists,sdf 12 + grad 16 + rgb 12 in; w 4 out
            ("composite_bwd", comp_bwd, N * (S * 68 + 44), "hbm"),                # re-read 32 + weights/d_weights 8, d_sdf 4 + d_normal 12 + d_rgb 12 out (the kernel's own traffic: 40 in + d_w 4, d_grad4 16 + d_sdf 4 + d_rgb 12 out = 76)
            ("upsample", ups, N * ((S - 16) * 8 + 64), "hbm"),
            ("merge_z", mrg, N * (2 * (112 + 16) * 4 + 2 * S * 4), "hbm"),
            ("sdf_query_chain", None if args.hbm_only else (lambda: rnd.sdf_network.query_flat(flat, x)), P * 918016, "tensor"),
            ("render_mlp_fwd_bwd", mlp, P * (6 * 1049088 + 3 * 543744), "tensor"),
        ]
        for name, fn, work, bound in rows:
            if fn is None or (name == "sdf_query_chain" and N > 65536):
                continue
            t = timeit(fn)
            if bound == "hbm":
                ach, peak, unit = work / t / 1e9, pk["hbm"], "GB/s"
            else:
                ach, peak, unit = work / t / 1e12, pk["bf16"], "TFLOP/s"
            print(json.dumps({"kernel": name, "rays": N, "samples": S, "us": t * 1e6, "achieved": ach, "peak": peak, "unit": unit,
                              "frac": ach / peak, "bound": bound, "bytes_or_flops": work,
                              "l2_resident": bool(bound == "hbm" and work < 100e6), "peak_source": pk["src"]}))


if __name__ == "__main__":
    main()
