"""Pins the oracle (oracle/neus_oracle.py) to the reference: replays the fixtures that
tests/golden/make_golden.py produced from the imported reference.  CPU only."""
import torch

import oracle as O
from conftest import assert_close, load_golden, unflatten

KEYS = ["sdf", "color_fine", "depth_pred", "weighted_z_vals", "s_val", "cdf_fine", "weight_sum", "weight_max",
        "normals", "sdf_flows", "sampled_points", "weights", "inside_sphere", "weight_inside", "weight_outside"]


def test_embedder():
    g = load_golden("embed")
    assert torch.equal(O.embed(g["x4"], 6), g["e6"])
    assert torch.equal(O.embed(g["x3"], 4), g["e4"])


def test_constructor_rng_stream():
    """Reference constructors under seed 678 (configs/default.yaml:51) == oracle init under the same seed."""
    g = load_golden("init_seed678")
    torch.manual_seed(678)
    P = dict(sdf=O.init_sdf_params(**O.DEFAULT_CFG["sdf"]), color=O.init_color_params(**O.DEFAULT_CFG["color"]),
             variance=O.init_variance_params(**O.DEFAULT_CFG["variance"]))
    n = 0
    for tag, p in P.items():
        for k, v in p.items():
            assert_close(v.double().sum(), g[f"{tag}.{k}.sum"], 1e-12, f"{tag}.{k}.sum")
            assert torch.equal(v.flatten()[:4], g[f"{tag}.{k}.head"].flatten())
            n += 1
    assert n == 27 + 15 + 1
    assert sum(v.numel() for v in P["sdf"].values()) == 529050
    assert sum(v.numel() for v in P["color"].values()) == 273926


def test_full_size_fields():
    g = load_golden("full_fields_seed678")
    torch.manual_seed(678)
    sdf = O.init_sdf_params(**O.DEFAULT_CFG["sdf"])
    col = O.init_color_params(**O.DEFAULT_CFG["color"])
    y = O.sdf_forward(sdf, g["x"])
    grad = O.sdf_gradient(sdf, g["x"].clone()).squeeze(1)
    assert_close(y, g["y"], 1e-6, "y")
    assert_close(grad, g["grad"], 1e-5, "grad")
    assert_close(O.color_forward(col, g["x"], grad.detach(), g["dirs"], y[:, 1:].detach()), g["rgb"], 1e-6, "rgb")


def test_small_fields_and_eikonal_double_backward(small_params):
    g = load_golden("small_fields")
    P = small_params
    y = O.sdf_forward(P["sdf"], g["x"])
    grad = O.sdf_gradient(P["sdf"], g["x"].clone()).squeeze(1)
    assert_close(y, g["y"], 1e-6, "y")
    assert_close(grad, g["grad"], 1e-5, "grad")
    assert_close(O.color_forward(P["color"], g["x"], grad.detach(), g["dirs"], y[:, 1:].detach()), g["rgb"], 1e-6)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P["sdf"].items()}
    (O.sdf_gradient(Pg, g["x"].clone()).squeeze(1)[:, :3].norm(dim=-1) - 1).pow(2).mean().backward()
    for k, v in Pg.items():
        got = v.grad if v.grad is not None else torch.zeros_like(v)
        assert_close(got, g[f"eik.{k}"], 1e-5, k)


def test_sample_pdf_bit_exact():
    g = load_golden("sample_pdf")
    cdf = O.cdf_from_weights(g["weights"])
    assert torch.equal(cdf, g["cdf"])
    smp, inds = O.search_cdf(g["cdf"], g["bins"], 16)
    assert torch.equal(inds, g["inds"]) and inds.dtype == torch.int64
    assert torch.equal(smp, g["samples"])
    assert torch.equal(O.sample_pdf(g["bins"], g["weights"], 16), g["samples"])


def test_up_sample_and_cat(small_params):
    g = load_golden("up_sample")
    for S, inv_s in ((64, 64), (80, 128), (96, 256), (112, 512)):
        nz, aux = O.up_sample(g[f"z{S}"], g[f"sdf{S}"], 16, inv_s, return_aux=True)
        assert torch.equal(aux["inds"], g[f"inds{S}"])
        assert torch.equal(aux["cdf"], g[f"cdf{S}"])
        assert torch.equal(nz, g[f"new_z{S}"])
    with torch.no_grad():
        z2, s2 = O.cat_z_vals(small_params["sdf"], g["rays_o"], g["rays_d"], g["t"], g["z64"], g["new_z64"],
                              g["sdf64"], last=False)
    assert torch.equal(z2, g["cat_z"])
    assert_close(s2, g["cat_sdf"], 1e-6)
    assert (z2[:, 1:] > z2[:, :-1]).all(), "fixture must have no exact ties (SURVEY gotcha 10)"


def test_renderer_forward_eval_and_train(small_params):
    g = load_golden("render_small")
    a = (g["rays_o"], g["rays_d"], g["rays_d_norm"], g["t"], g["near"], g["far"])
    oe = O.render(small_params, *a, cos_anneal=0.5, eval_mode=True)
    ot = O.render(small_params, *a, cos_anneal=0.3, eval_mode=False, t_rand=g["t_rand"])
    assert [k for k in oe.keys()] == KEYS
    for k in KEYS:
        assert_close(oe[k], g[f"eval.{k}"], 2e-5, f"eval.{k}")
        assert_close(ot[k], g[f"train.{k}"], 2e-5, f"train.{k}")
    # the train-mode jitter comes from the CPU generator (neus_renderer.py:482)
    torch.manual_seed(77)
    o2 = O.render(small_params, *a, cos_anneal=0.3, eval_mode=False)
    assert_close(o2["color_fine"], g["train.color_fine"], 2e-5)


def test_poses_and_rays():
    g = load_golden("poses_rays")
    pose = dict(r=g["r"], t=g["t"], init_c2w=g["init_c2w"])
    for cam in range(3):
        assert_close(O.pose_forward(pose, cam), g[f"c2w{cam}"], 1e-7, f"c2w{cam}")
    H, W = int(g["H"]), int(g["W"])
    torch.manual_seed(9)
    assert torch.equal(O.patch_indices(H, W, 4, 64), g["idx"])
    assert torch.equal(O.pixel_grid(H, W)[1][:, g["idx"]], g["pix"])
    S = torch.eye(4).unsqueeze(0)
    o, d, n = O.ray_generation(g["pix"], g["K"], O.pose_forward(pose, 2), S)
    assert_close(o, g["ray_o2"], 1e-6); assert_close(d, g["ray_d2"], 1e-6); assert_close(n, g["ray_n2"], 1e-6)
    pg = {k: v.clone().requires_grad_(k in ("r", "t")) for k, v in pose.items()}
    o, d, _ = O.ray_generation(g["pix"], g["K"], O.pose_forward(pg, 1), S)
    ((o * g["wgt_o"]).sum() + (d * g["wgt_d"]).sum()).backward()
    assert_close(pg["r"].grad, g["dr1"], 1e-5, "dr"); assert_close(pg["t"].grad, g["dt1"], 1e-5, "dt")
    # r = 0 (the zero-initialised common case, SURVEY gotcha 15): R = I exactly
    assert torch.equal(O.so3_exp(torch.zeros(3)), torch.eye(3))


def test_full_step_parameter_gradients(small_params):
    g = load_golden("step_small")
    P = {t: {k: v.clone().requires_grad_(True) for k, v in small_params[t].items()} for t in small_params}
    pose = dict(r=g["r"].clone().requires_grad_(True), t=g["tr"].clone().requires_grad_(True),
                init_c2w=torch.eye(4).unsqueeze(0))
    H, W = 60, 80
    K = O.camera_matrix(0.8 * W, 0.8 * W, W, H).unsqueeze(0)
    loss, aux = O.train_step(P, pose, g["pix"], K, torch.eye(4).unsqueeze(0), g["rgb_gt"], g["t"], [0.01, 5.0],
                             cos_anneal=0.5, t_rand=g["t_rand"])
    loss.backward()
    assert_close(loss, g["loss"], 1e-6, "loss")
    assert_close(aux["out"]["color_fine"], g["color"], 1e-5); assert_close(aux["out"]["depth_pred"], g["depth"], 1e-5)
    for tag in ("sdf", "color", "variance"):
        for k, v in P[tag].items():
            assert_close(v.grad, g[f"grad.{tag}.{k}"], 2e-4, f"{tag}.{k}")
    assert_close(pose["r"].grad, g["dr"], 2e-4, "dr"); assert_close(pose["t"].grad, g["dt"], 2e-4, "dt")


def test_eval_image_render(small_params):
    """Evaluation image render (model/training.py:210-262 replayed with the imported reference in make_golden.py):
    rgb, depth, weighted z, arg-max-weight depth and the weighted normal map, in the reference's chunking."""
    g = load_golden("eval_image_small")
    H, W, ch = int(g["H"]), int(g["W"]), int(g["chunk"])
    out = O.render_image(small_params, g["world"], g["K"], torch.eye(4).unsqueeze(0), H, W, g["t"], [0.01, 5.0], cos_anneal=1.0,
                         chunk=ch)
    for k in ("rgb", "depth_pred", "weighted_z_vals", "depth_highest_weight", "normal"):
        assert out[k].shape == g[k].shape
        assert_close(out[k], g[k], 2e-5, k)
    # chunking is only a memory device: one chunk == the reference's chunks
    out1 = O.render_image(small_params, g["world"], g["K"], torch.eye(4).unsqueeze(0), H, W, g["t"], [0.01, 5.0], cos_anneal=1.0,
                          chunk=H * W)
    assert_close(out1["rgb"], g["rgb"], 2e-5, "rgb one chunk")


def _motion_params(g):
    return {k[len("param."):]: v for k, v in g.items() if k.startswith("param.")}


def test_motion_network_and_pose_integration():
    """Continuous pose model (model/neus_fields.py:79-201): constructor RNG stream, forward, relative poses over 7 frame
    pairs x 10 sub-steps, the w2c chain and every parameter gradient, against the imported reference (make_golden.py)."""
    g = load_golden("motion_small")
    P = _motion_params(g)
    cfg = dict(O.MOTION_CFG, d_hidden=64)
    torch.manual_seed(31)
    init = O.init_motion_params(**cfg)
    for k in ("lin0.weight_v", "lin2.weight_g", "lin3.bias"):      # lin4 was rescaled in the generator
        assert torch.equal(init[k], P[k]), k
    a, v = O.motion_forward(P, g["t_query"])
    assert_close(a, g["ang"], 1e-6, "ang"); assert_close(v, g["vel"], 1e-6, "vel")
    n_img, n_sub, first, last = int(g["n_img"]), int(g["n_sub"]), int(g["first"]), int(g["last"])
    Pg = {k: x.clone().requires_grad_(True) for k, x in P.items()}
    dt, rel = O.relative_camera_pose(Pg, first, last, n_img, n_sub)
    w2c = O.w2c_mappings(rel)
    assert_close(dt, g["dt"], 0, "dt")
    assert_close(torch.stack(rel), g["rel"], 1e-6, "rel"); assert_close(w2c, g["w2c"], 1e-6, "w2c")
    (w2c * g["wgt"]).sum().backward()
    for k in P:
        assert_close(Pg[k].grad, g[f"grad.{k}"], 2e-5, f"grad {k}")


def _stage1_oracle(g, requires_grad=True):
    """Replays tests/golden/stage1_small.npz (the reference's own train.py:467-517 lines, executed by make_golden.py)."""
    mk = (lambda v: v.clone().requires_grad_(True)) if requires_grad else (lambda v: v.clone())
    mp = {k: mk(v) for k, v in unflatten(g, "motion.").items()}
    sp = {k: mk(v) for k, v in unflatten(g, "sdfnet.").items()}
    lv = {k: mk(v) for k, v in unflatten(g, "in.").items()}
    small = dict(multires=6, skip_in=(4,), scale=1.0)
    out = {"sampled_points": lv["pts"], "normals": lv["normals"], "sdf_flows": lv["sdf_flows"], "weights": lv["weights"],
           "sdf": lv["sdf"]}
    res = O.stage1_losses(sp, mp, out, g["rgb_gt"], float(g["query_time_step"]), int(g["image_idx"]),
                          [int(v) for v in g["ref_idx"]], int(g["nb_valid"]), int(g["total_nb_images"]),
                          int(g["nb_sample_timestep"]), g["Kr"], g["scale"], g["norm_pix"], g["pix"], (int(g["H"]), int(g["W"])),
                          g["refs"], int(g["world_cam_idx"]), float(g["world_time_step"]), sdf_kw=small)
    return res, mp, sp, lv


def test_stage1_auxiliary_losses():
    g = load_golden("stage1_small")
    res, mp, sp, lv = _stage1_oracle(g)
    for k in ("sdf_loss", "flow_rgb_loss", "sdf_consistency_loss"):
        assert_close(res[k], g[k], 1e-6, k)
    assert_close(torch.stack(res["flow_fw_pred"]), g["flow_fw_pred"], 1e-5, "flow_fw_pred")
    w = g["loss_weights"]
    (w[0] * res["sdf_loss"] + w[1] * res["flow_rgb_loss"] + w[2] * res["sdf_consistency_loss"]).backward()
    for k, v in lv.items():
        assert_close(v.grad, g[f"grad.{k}"], 2e-5, f"grad {k}")
    n = 0
    for tag, params in (("motion", mp), ("sdfnet", sp)):
        for k, v in params.items():
            if f"grad.{tag}.{k}" in g:
                assert_close(v.grad, g[f"grad.{tag}.{k}"], 5e-5, f"grad {tag}.{k}")
                n += 1
    assert n == 15 + 27


def test_stage1_iteration_through_renderer():
    """tests/golden/stage1_render_small.npz: the imported reference NeuSRenderer + the reference's own train.py:467-526 lines
    + the imported mdl.Trainer.compute_loss, back-propagated into every SDF / colour / variance / motion / pose parameter
    through depth_pred, weights, sdf and sampled_points.  Replayed by oracle.stage1_step."""
    g = load_golden("stage1_render_small")
    mk = lambda v: v.clone().requires_grad_(True)
    P = {t: {k: mk(v) for k, v in unflatten(g, f"param.{t}.").items()} for t in ("sdf", "color", "variance")}
    mp = {k: mk(v) for k, v in unflatten(g, "param.motion.").items()}
    pose = dict(r=mk(g["r"]), t=mk(g["t"]), init_c2w=g["init_c2w"].clone())
    o, d, dn = O.ray_generation(g["norm_pix"][None], g["K"], O.pose_forward(pose, 0), torch.eye(4)[None])
    near, far = O.near_far(o, d, [float(v) for v in g["depth_range"]])
    w = dict(rgb=float(g["w.rgb_weight"]), eikonal=float(g["w.eikonal_weight"]), sdf=float(g["w.sdf_weight"]),
             flow_rgb=float(g["w.flow_rgb_weight"]), sdf_consistency=float(g["w.sdf_consistency_weight"]),
             edge_aware_smoothness=float(g["w.edge_aware_smoothness_weight"]), smoothness=float(g["w.smoothness_weight"]))
    loss, parts, out = O.stage1_step(P, mp, o, d, dn, near, far, g["rgb_gt"], float(g["query_time_step"]), int(g["image_idx"]),
                                     [int(v) for v in g["ref_idx"]], int(g["nb_valid"]), int(g["total_nb_images"]),
                                     int(g["nb_sample_timestep"]), g["Kr"], torch.eye(4)[None], g["norm_pix"], g["pix"],
                                     (int(g["H"]), int(g["W"])), g["refs"], int(g["world_cam_idx"]), float(g["world_time_step"]),
                                     w, patch_size=int(g["ps"]), s_level=int(g["s_level"]), cos_anneal=float(g["cos_anneal"]),
                                     t_rand=g["t_rand"])
    assert_close(loss, g["loss.loss"], 2e-4, "loss")
    for a, b in (("rgb", "loss_rgb"), ("eikonal", "loss_eikonal"), ("sdf", "loss_sdf"), ("flow_rgb", "loss_flow_rgb"),
                 ("sdf_consistency", "sdf_consistency_loss"), ("edge_aware_smoothness", "edge_aware_smoothness_loss"),
                 ("smoothness", "smoothness_loss")):
        assert_close(parts[a], g[f"loss.{b}"], 2e-5, a)
    assert float(g["weight_sum"].max()) > 0.9 and float(g["loss.edge_aware_smoothness_loss"]) > 1e-3   # a non-degenerate scene
    loss.backward()
    from conftest import rel_err
    n = 0
    for tag, params in (("sdf", P["sdf"]), ("color", P["color"]), ("variance", P["variance"]), ("motion", mp)):
        for k, v in params.items():
            assert rel_err(v.grad, g[f"grad.{tag}.{k}"]) < 1e-3, f"{tag}.{k}"
            n += 1
    assert n == 27 + 15 + 1 + 15
    assert rel_err(pose["r"].grad, g["dr"]) < 1e-3 and rel_err(pose["t"].grad, g["dt"]) < 1e-3


def test_pretrained_sdf_weights_fixture():
    """pretrained_sdf/model.pt (train.py:41-43), the reference's only fixture: forward, analytic gradient and the eikonal +
    value double backward of the oracle against the imported reference's numbers (tests/golden/pretrained_sdf.npz)."""
    g = load_golden("pretrained_sdf")
    P = {k: v.clone().requires_grad_(True) for k, v in unflatten(g, "param.").items()}
    assert len(P) == 27
    x = g["x"]
    assert_close(O.sdf_forward(P, x), g["y"], 1e-5, "pretrained fwd")
    og = O.sdf_gradient(P, x.clone()).squeeze(1)
    assert_close(og, g["grad"], 1e-5, "pretrained grad")
    assert abs(float(og[:, :3].norm(dim=-1).mean()) - 1.0) < 0.05          # a trained SDF: |grad| ~ 1 (the plane sdf = y + 1)
    ((og[:, :3].norm(dim=-1) - 1).pow(2).mean() + (O.sdf_forward(P, x) * g["wy"]).sum() / x.shape[0]).backward()
    from conftest import rel_err
    for k, v in P.items():
        assert abs(float(v.grad.double().norm()) - float(g[f"gnorm.{k}"])) <= 1e-3 * float(g[f"gnorm.{k}"]) + 1e-9, k
        if f"grad.{k}" in g:
            assert rel_err(v.grad, g[f"grad.{k}"]) < 1e-3, k


def test_pose_refinement_warp():
    """utils_poses/pose_refinement.py:34-61 + :117-126 (fixture: the reference's own lines with its PoseRetriever)."""
    g = load_golden("pose_refine_small")
    B, _, H, W = g["images"].shape
    pose = dict(r=g["r"].clone().requires_grad_(True), t=g["t"].clone().requires_grad_(True), init_c2w=torch.eye(4).repeat(B, 1, 1))
    rel = torch.stack([O.pose_forward(pose, i) for i in range(B)])
    assert_close(rel, g["rel"], 1e-6, "relative poses")
    uv = O.refine_uv(H, W).unsqueeze(0).repeat(B, 1, 1, 1)
    lp, wp = O.compute_loss_and_warp_image(g["images"], g["next_images"], g["depths"], g["K"], uv, rel)
    ln, wn = O.compute_loss_and_warp_image(g["next_images"], g["images"], g["next_depths"], g["K"], uv, torch.inverse(rel))
    assert_close(lp, g["loss_pos"], 1e-6, "loss pos"); assert_close(ln, g["loss_neg"], 1e-6, "loss neg")
    assert_close(wp, g["warped_pos"], 1e-6, "warped pos"); assert_close(wn, g["warped_neg"], 1e-6, "warped neg")
    ((lp + ln) / 2).backward()
    assert_close(pose["r"].grad, g["dr"], 1e-5, "dr"); assert_close(pose["t"].grad, g["dt"], 1e-5, "dt")


def test_eval_flow_map():
    """model/training.py:203-208, 265-283, 296-297: predicted forward optical flow of the evaluation render."""
    g = load_golden("eval_flow_small")
    P = {t: unflatten(g, f"param.{t}.") for t in ("sdf", "color", "variance")}
    out = O.render_image(P, g["world"], g["K"], torch.eye(4).unsqueeze(0), int(g["H"]), int(g["W"]), g["t0"], [0.5, 3.5],
                         cos_anneal=1.0, chunk=50, flow=(unflatten(g, "param.motion."), g["t0"], g["t1"], int(g["n_sub"])))
    assert_close(out["rgb"], g["rgb"], 5e-4, "rgb")
    assert (out["flow_pred"] - g["flow_pred"]).abs().max() <= 5e-4 * g["flow_pred"].abs().max()
