"""The reference's whole stage-1 training iteration (train.py:441-531) through the REAL renderer: the losses built on the
un-detached renderer outputs — flow-RGB on weights / sampled_points (train.py:488), SDF-consistency on sdf (:505), the
depth-patch smoothness terms on depth_pred (:519-525) — must send their gradients through `_RenderCoreFn.backward` /
`_RenderStepFn.backward` (composite_bwd's d_depth / d_weights inputs, the d_sdf add, d_pts4 -> ray_points_bwd) into every
SDF / colour / variance / motion / pose parameter.

Fixture: tests/golden/stage1_render_small.npz, produced by make_golden.py from the imported reference NeuSRenderer + the
reference's own source lines train.py:467-526 + the imported mdl.Trainer.compute_loss.  fp32 path: <= 1e-3 relative;
bf16 tensor-core path (full-size networks, checked against the CPU oracle): cosine similarity > 0.999."""
import pytest
import torch

import cope_nerf_b200 as C
import oracle as O
from cope_nerf_b200 import losses as CL
from cope_nerf_b200.common import get_world_cameraOrigin_cameraRay
from conftest import cos_sim, load_golden, rel_err, unflatten
from test_gpu_parity import SMALL_CFG, cu, full_params, renderer_from

pytestmark = pytest.mark.gpu
DEV = "cuda"
MCFG = dict(d_out=6, d_in=1, d_hidden=64, n_layers=4, skip_in=[2], multires=6, bias=0.5, scale=1.0, geometric_init=False,
            weight_norm=True)
PART_KEYS = (("rgb", "loss_rgb"), ("eikonal", "loss_eikonal"), ("sdf", "loss_sdf"), ("flow_rgb", "loss_flow_rgb"),
             ("sdf_consistency", "sdf_consistency_loss"), ("edge_aware_smoothness", "edge_aware_smoothness_loss"),
             ("smoothness", "smoothness_loss"))


def _stage1_iteration(rnd, mot, pose, g, w, fused, cos_anneal):
    """One stage-1 iteration on the device.  `w`: the reference's loss weights (model/training.py:499-505 names)."""
    ps, s_level = int(g["ps"]), int(g["s_level"])
    world = pose(0)
    o, d, dn = get_world_cameraOrigin_cameraRay(cu(g["norm_pix"])[None], cu(g["K"]), world, torch.eye(4, device=DEV)[None])
    near, far = C.training.near_far_from_sphere(o, d, [float(v) for v in g["depth_range"]])
    rnd.t_rand_override = g["t_rand"]
    rgb_gt = cu(g["rgb_gt"])
    qts = cu(g["query_time_step"]).float()
    ang, vel = mot(qts.view(-1, 1))
    motion = torch.cat([ang, vel], dim=1)
    if fused:
        base, out = rnd.forward_losses(o, d, dn, qts, near, far, rgb_gt, cos_anneal_ratio=cos_anneal, it=1,
                                       rgb_weight=w["rgb_weight"], eikonal_weight=w["eikonal_weight"],
                                       sdf_weight=w["sdf_weight"], motion=motion)
        parts3 = out["loss_rgb"], out["loss_eikonal"], out["loss_sdf"]
    else:
        out = rnd(o, d, dn, qts, near, far, cos_anneal_ratio=cos_anneal, it=1, eval=False)
        base, p = CL.step_losses(out, rgb_gt, w["rgb_weight"], w["eikonal_weight"], w["sdf_weight"], motion=motion)
        parts3 = p[1], p[2], p[3]
    aux = CL.stage1_losses(out, rgb_gt, mot, rnd.sdf_network, float(g["query_time_step"]), int(g["image_idx"]),
                           [int(v) for v in g["ref_idx"]], int(g["nb_valid"]), int(g["total_nb_images"]),
                           int(g["nb_sample_timestep"]), cu(g["Kr"]), torch.eye(4, device=DEV)[None], cu(g["norm_pix"]),
                           cu(g["pix"]), cu(g["refs"]), int(g["world_cam_idx"]), float(g["world_time_step"]),
                           include_sdf_loss=False)
    f = 1.0 / 2 ** s_level                                                   # train.py:523,525
    sm, smp = CL.depth_smoothness_losses(out["depth_pred"], rgb_gt, ps, edge_weight=w["edge_aware_smoothness_weight"] * f,
                                         smooth_weight=w["smoothness_weight"] * f)
    loss = base + w["flow_rgb_weight"] * aux["flow_rgb_loss"] + w["sdf_consistency_weight"] * aux["sdf_consistency_loss"] + sm
    parts = dict(rgb=parts3[0], eikonal=parts3[1], sdf=parts3[2], flow_rgb=aux["flow_rgb_loss"],
                 sdf_consistency=aux["sdf_consistency_loss"], edge_aware_smoothness=smp[1] * f, smoothness=smp[2] * f)
    return loss, parts, out, aux


@pytest.mark.parametrize("fused", [False, True])
def test_stage1_iteration_through_renderer_golden(fused):
    g = load_golden("stage1_render_small")
    rnd = renderer_from({t: unflatten(g, f"param.{t}.") for t in ("sdf", "color", "variance")}, SMALL_CFG)
    mot = C.MotionNetwork(**MCFG).to(DEV)
    mot.load_state_dict(unflatten(g, "param.motion."))
    pose = C.PoseRetriever(1, init_c2w=g["init_c2w"].clone()).to(DEV)
    with torch.no_grad():
        pose.r.copy_(g["r"]); pose.t.copy_(g["t"])
    w = {k[2:]: float(v) for k, v in g.items() if k.startswith("w.")}
    loss, parts, out, aux = _stage1_iteration(rnd, mot, pose, g, w, fused, float(g["cos_anneal"]))
    loss.backward()
    assert rel_err(out["color_fine"], g["color"]) < 1e-4 and rel_err(out["depth_pred"], g["depth"]) < 1e-4
    assert rel_err(out["weights"], g["weights"]) < 1e-3
    assert rel_err(aux["flow_fw_pred"], g["flow_fw_pred"]) < 1e-3
    for a, b in PART_KEYS:
        assert rel_err(parts[a], g[f"loss.{b}"]) < 1e-3, (a, float(parts[a]), float(g[f"loss.{b}"]))
    assert rel_err(loss, g["loss.loss"]) < 1e-3
    worst = 0.0
    for tag, net in (("sdf", rnd.sdf_network), ("color", rnd.color_network), ("variance", rnd.deviation_network), ("motion", mot)):
        for k, p in net.named_parameters():
            ref = g[f"grad.{tag}.{k}"]
            assert p.grad is not None, f"{tag}.{k} received no gradient"
            e = rel_err(p.grad, ref)
            worst = max(worst, e)
            assert e < 1e-3, f"{tag}.{k}: rel {e:.2e} (fused={fused})"
    e_r, e_t = rel_err(pose.r.grad, g["dr"]), rel_err(pose.t.grad, g["dt"])
    assert e_r < 1e-3 and e_t < 1e-3, (e_r, e_t)
    print(f"stage-1 iteration (fused={fused}): worst parameter-gradient rel err {worst:.2e}, pose {e_r:.2e} / {e_t:.2e}")


def test_renderer_backward_each_output_vs_oracle():
    """Upstream gradients on depth_pred, weights, sdf and sampled_points ONE AT A TIME (plus eval_mode=1 for the
    rays_d_norm division of depth): parameter / variance / ray gradients of the real renderer against oracle autograd."""
    g = load_golden("render_small")
    sp = load_golden("small_weights")
    P = dict(sdf=unflatten(sp, "sdf."), color=unflatten(sp, "color."), variance=unflatten(sp, "variance."))
    n = g["rays_o"].shape[0]
    torch.manual_seed(3)
    probes = dict(depth_pred=torch.randn(n, 1), weights=torch.randn(n, 128), sdf=torch.randn(n * 128, 1),
                  sampled_points=torch.randn(n, 128, 3), normals=torch.randn(n, 128, 3), sdf_flows=torch.randn(n, 128, 1))
    for key, eval_mode in [(k, False) for k in probes] + [("depth_pred", True)]:
        rnd = renderer_from(P, SMALL_CFG)
        rnd.t_rand_override = g["t_rand"]
        ro, rd = cu(g["rays_o"]).requires_grad_(True), cu(g["rays_d"]).requires_grad_(True)
        out = rnd(ro, rd, cu(g["rays_d_norm"]), cu(g["t"]), cu(g["near"]), cu(g["far"]), cos_anneal_ratio=0.3, it=1, eval=eval_mode)
        (out[key] * cu(probes[key])).sum().backward()
        Pg = {t: {k: v.clone().requires_grad_(True) for k, v in P[t].items()} for t in P}
        oro, ord_ = g["rays_o"].clone().requires_grad_(True), g["rays_d"].clone().requires_grad_(True)
        oo = O.render(Pg, oro, ord_, g["rays_d_norm"], g["t"], g["near"], g["far"], cos_anneal=0.3, eval_mode=eval_mode,
                      t_rand=None if eval_mode else g["t_rand"])
        (oo[key] * probes[key]).sum().backward()
        for tag, net in (("sdf", rnd.sdf_network), ("color", rnd.color_network), ("variance", rnd.deviation_network)):
            for k, p in net.named_parameters():
                ref = Pg[tag][k].grad
                if ref is None or ref.abs().max() == 0:
                    assert p.grad is None or p.grad.abs().max().item() < 1e-6, (key, tag, k)
                    continue
                e = rel_err(p.grad, ref)
                assert e < 1e-3, f"d {key} (eval={eval_mode}) -> {tag}.{k}: rel {e:.2e}"
        if oro.grad is not None and oro.grad.abs().max() > 0:
            assert rel_err(ro.grad, oro.grad) < 1e-3 and rel_err(rd.grad, ord_.grad) < 1e-3, key


def test_stage1_iteration_bf16_full_size_vs_oracle():
    """The same iteration on the tensor-core path with the full-size networks (fused render + loss node), against the CPU
    oracle's stage1_step: losses within 2 % (bf16 activations), parameter gradients cos > 0.999."""
    g = load_golden("stage1_render_small")
    P = full_params(678, perturb=0.01)
    torch.manual_seed(77)
    mp = O.init_motion_params(**{**MCFG, "skip_in": (2,)})
    with torch.no_grad():
        for k in ("lin4.weight_g", "lin4.bias"):
            mp[k].mul_(3.0).add_(0.1)
    rnd = renderer_from(P, C.training.DEFAULT_CFG)
    rnd.sdf_network.precision = rnd.color_network.precision = C.PREC_BF16
    mot = C.MotionNetwork(**MCFG).to(DEV)
    mot.load_state_dict(mp)
    pose = C.PoseRetriever(1, init_c2w=g["init_c2w"].clone()).to(DEV)
    with torch.no_grad():
        pose.r.copy_(g["r"]); pose.t.copy_(g["t"])
    w = {k[2:]: float(v) for k, v in g.items() if k.startswith("w.")}
    loss, parts, out, aux = _stage1_iteration(rnd, mot, pose, g, w, True, float(g["cos_anneal"]))
    loss.backward()
    # oracle
    Pg = {t: {k: v.clone().requires_grad_(True) for k, v in P[t].items()} for t in P}
    mg = {k: v.clone().requires_grad_(True) for k, v in mp.items()}
    po = dict(r=g["r"].clone().requires_grad_(True), t=g["t"].clone().requires_grad_(True), init_c2w=g["init_c2w"].clone())
    oo, od, on = O.ray_generation(g["norm_pix"][None], g["K"], O.pose_forward(po, 0), torch.eye(4)[None])
    near, far = O.near_far(oo, od, [float(v) for v in g["depth_range"]])
    wo = dict(rgb=w["rgb_weight"], eikonal=w["eikonal_weight"], sdf=w["sdf_weight"], flow_rgb=w["flow_rgb_weight"],
              sdf_consistency=w["sdf_consistency_weight"], edge_aware_smoothness=w["edge_aware_smoothness_weight"],
              smoothness=w["smoothness_weight"])
    ol, op, oout = O.stage1_step(Pg, mg, oo, od, on, near, far, g["rgb_gt"], float(g["query_time_step"]), int(g["image_idx"]),
                                 [int(v) for v in g["ref_idx"]], int(g["nb_valid"]), int(g["total_nb_images"]),
                                 int(g["nb_sample_timestep"]), g["Kr"], torch.eye(4)[None], g["norm_pix"], g["pix"],
                                 (int(g["H"]), int(g["W"])), g["refs"], int(g["world_cam_idx"]), float(g["world_time_step"]), wo,
                                 patch_size=int(g["ps"]), s_level=int(g["s_level"]), cos_anneal=float(g["cos_anneal"]),
                                 t_rand=g["t_rand"], motion_kw=dict(multires=6, skip_in=(2,)))
    ol.backward()
    assert rel_err(loss, ol) < 2e-2, (float(loss), float(ol))
    assert rel_err(out["color_fine"], oout["color_fine"]) < 2e-2
    report = {}
    for tag, net, ref in (("sdf", rnd.sdf_network, Pg["sdf"]), ("color", rnd.color_network, Pg["color"]), ("motion", mot, mg)):
        cs = {k: cos_sim(p.grad, ref[k].grad) for k, p in net.named_parameters() if ref[k].grad is not None and ref[k].grad.abs().max() > 0}
        report[tag] = min(cs.values())
        bad = {k: v for k, v in cs.items() if v <= 0.999}
        assert not bad, f"{tag}: cos <= 0.999 for {bad}"
    print("stage-1 bf16 iteration: min cos per network", {k: round(v, 5) for k, v in report.items()},
          "variance rel", rel_err(rnd.deviation_network.variance.grad, Pg["variance"]["variance"].grad))
    assert rel_err(rnd.deviation_network.variance.grad, Pg["variance"]["variance"].grad) < 3e-2
