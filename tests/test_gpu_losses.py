"""GPU parity of the loss kernels (SURVEY.md 8 a17 + 8f rank 2: cope_step_losses_*, cope_weighted_points_*, cope_flow_rgb_*) and of
the fused render + loss training node, through the Python module API -> ctypes C-ABI, against the CPU oracle and the fixture
that make_golden.py produced by executing the reference's own train.py:467-517 lines.  fp32 path: <= 1e-3 relative."""
import pytest
import torch

import cope_nerf_b200 as C
import oracle as O
from cope_nerf_b200 import losses as CL
from conftest import assert_close, cos_sim, load_golden, rel_err, unflatten
from test_gpu_parity import SMALL_CFG, cu, renderer_from

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand_out(n, s, seed, with_grad=True):
    torch.manual_seed(seed)
    mk = lambda t: t.requires_grad_(with_grad)
    grad4 = mk(torch.randn(n * s, 4))
    pts4 = mk(torch.cat([torch.randn(n * s, 3) * 0.5, torch.full((n * s, 1), 0.2)], dim=1))
    return dict(color_fine=mk(torch.rand(n, 3)), _grad4=grad4, _pts4=pts4,
                weights=mk(torch.softmax(torch.randn(n, s), dim=1) * 0.8), rgb_gt=torch.rand(n, 3))


def _as_outputs(d):
    """a RenderOutputs (15-key dict + packed attributes) from the test's tensors"""
    from cope_nerf_b200.renderer import RenderOutputs
    out = RenderOutputs({k: v for k, v in d.items() if not k.startswith("_")})
    out.grad4, out.pts4 = d["_grad4"], d["_pts4"]
    return out


def _oracle_view(o):
    n = o["color_fine"].shape[0]
    return {"color_fine": o["color_fine"], "normals": o["_grad4"][:, :3].reshape(n, -1, 3),
            "sdf_flows": o["_grad4"][:, 3:].reshape(n, -1, 1), "sampled_points": o["_pts4"][:, :3].reshape(n, -1, 3),
            "weights": o["weights"]}


@pytest.mark.parametrize("n,s,with_motion", [(7, 5, False), (33, 128, True), (1024, 128, True), (1, 1, True)])
def test_step_losses_vs_oracle(n, s, with_motion):
    ref = _rand_out(n, s, 100 + n)
    if n == 7:
        with torch.no_grad():
            ref["_grad4"][3, :3] = 0.0            # |n| = 0: torch's norm backward returns 0 there
            ref["color_fine"][2] = ref["rgb_gt"][2]   # |x| at 0: sign(0) = 0
    dev = {k: v.detach().to(DEV).requires_grad_(v.requires_grad) for k, v in ref.items()}
    mot_r = (torch.randn(6) * 0.7).requires_grad_(True) if with_motion else None
    mot_d = mot_r.detach().to(DEV).requires_grad_(True) if with_motion else None
    w = (0.33333, 0.1, 0.7 if with_motion else 0.0)
    ov = _oracle_view(ref)
    l_rgb, l_eik = O.rgb_l1_loss(ov["color_fine"], ref["rgb_gt"]), O.eikonal_loss(ov["normals"])
    l_flow = O.sdf_flow_loss(ov, mot_r[:3], mot_r[3:]) if with_motion else torch.zeros(())
    tot_r = w[0] * l_rgb + w[1] * l_eik + w[2] * l_flow
    (tot_r * 1.7).backward()
    tot, parts = CL.step_losses(_as_outputs(dev), dev["rgb_gt"], w[0], w[1], w[2], motion=mot_d)
    (tot * 1.7).backward()
    assert rel_err(tot, tot_r) < 1e-5 and rel_err(parts[0], tot_r) < 1e-5
    assert rel_err(parts[1], l_rgb) < 1e-5 and rel_err(parts[2], l_eik) < 1e-5
    if with_motion:
        assert rel_err(parts[3], l_flow) < 1e-5
        assert rel_err(mot_d.grad, mot_r.grad) < 1e-4
        assert rel_err(dev["_pts4"].grad, ref["_pts4"].grad) < 1e-4
        assert dev["weights"].grad is None        # the reference detaches the weights (train.py:477)
    assert rel_err(dev["color_fine"].grad, ref["color_fine"].grad) < 1e-5
    assert rel_err(dev["_grad4"].grad, ref["_grad4"].grad) < 1e-5


def test_step_losses_global_normaliser():
    """Sharded rays: the SDF-flow term divides by the weight sum over ALL ranks."""
    ref = _rand_out(16, 8, 5, with_grad=False)
    dev = {k: v.to(DEV) for k, v in ref.items()}
    mot = torch.randn(6)
    wsum = torch.tensor([37.5], device=DEV)
    _, parts = CL.step_losses(_as_outputs(dev), dev["rgb_gt"], 0.0, 0.0, 1.0, motion=mot.to(DEV), w_sum_global=wsum)
    local = O.sdf_flow_loss(_oracle_view(ref), mot[:3], mot[3:]) * (ref["weights"].sum() + 1e-10) / (37.5 + 1e-10)
    assert rel_err(parts[3], local) < 1e-5


def _stage1_fixture():
    g = load_golden("stage1_small")
    mcfg = dict(d_out=6, d_in=1, d_hidden=64, n_layers=4, skip_in=[2], multires=6, bias=0.5, scale=1.0,
                geometric_init=False, weight_norm=True)
    mot = C.MotionNetwork(**mcfg).to(DEV)
    mot.load_state_dict(unflatten(g, "motion."))
    sdf = C.SDFNetwork(**SMALL_CFG["neus_sdf_network"]).to(DEV)
    sdf.load_state_dict(unflatten(g, "sdfnet."))
    return g, mot, sdf


def test_weighted_points_and_flow_rgb_golden():
    """Flow-RGB loss on the reference's own numbers: value, predicted flow, and the gradients w.r.t. points, weights and the
    MotionNetwork parameters (through cope_flow_rgb_bwd -> cope_pose_chain_bwd -> cope_pose_integrate_bwd -> the MLP)."""
    g, mot, sdf = _stage1_fixture()
    n, S = int(g["n"]), int(g["S"])
    pts = cu(g["in.pts"]).reshape(-1, 3)
    pts4 = torch.cat([pts, torch.zeros_like(pts[:, :1])], dim=1).requires_grad_(True)
    weights = cu(g["in.weights"]).requires_grad_(True)
    nb = int(g["nb_valid"])
    _, c2c = mot.compute_relative_camera_pose(1, int(g["ref_idx"][nb - 1]), int(g["total_nb_images"]), int(g["nb_sample_timestep"]))
    sel = torch.as_tensor([int(r) - 1 for r in g["ref_idx"][:nb]], device=DEV)
    w2c = mot.compute_w2c_mappings(c2c)[sel]
    wp = CL.weighted_points(weights, pts4)
    ref_wp = torch.cat([(g["in.weights"][..., None] * g["in.pts"]).sum(1), g["in.weights"].sum(1, keepdim=True)], dim=1)
    assert_close(wp, ref_wp, 1e-6, "weighted points")
    KS = CL.projection_matrices(cu(g["scale"]), cu(g["Kr"])[:nb])
    loss, flow = CL.flow_rgb_loss(wp, w2c, KS, cu(g["norm_pix"]), cu(g["pix"]), cu(g["refs"])[:nb], cu(g["rgb_gt"]), return_flow=True)
    assert_close(flow, g["flow_fw_pred"], 1e-4, "flow_fw_pred")
    assert rel_err(loss, g["flow_rgb_loss"]) < 1e-5
    # gradient of the flow-RGB term alone: from the oracle on the same fixture
    mp = {k: v.clone().requires_grad_(True) for k, v in unflatten(g, "motion.").items()}
    lv = {k: v.clone().requires_grad_(True) for k, v in unflatten(g, "in.").items()}
    _, rel_o = O.relative_camera_pose(mp, 1, int(g["ref_idx"][nb - 1]), int(g["total_nb_images"]), int(g["nb_sample_timestep"]),
                                      multires=6, skip_in=(2,))
    w2c_o = O.w2c_mappings(rel_o)[sel.cpu()]
    fl = [O.flow_forward_prediction(lv["pts"].reshape(-1, 3), lv["weights"].reshape(-1), n, w2c_o[t], g["Kr"][t], g["scale"],
                                    g["norm_pix"], (int(g["H"]), int(g["W"]))) for t in range(nb)]
    O.flow_rgb_loss(fl, g["pix"], g["refs"], g["rgb_gt"]).backward()
    loss.backward()
    assert rel_err(pts4.grad[:, :3], lv["pts"].grad.reshape(-1, 3)) < 1e-3
    assert rel_err(weights.grad, lv["weights"].grad) < 1e-3
    for k, p in mot.named_parameters():
        assert rel_err(p.grad, mp[k].grad) < 2e-3, k


def test_stage1_losses_golden():
    """train.py:467-517 end to end (SDF-flow + flow-RGB + SDF-consistency) against the fixture made from the reference's lines:
    the three values and every gradient (inputs, MotionNetwork parameters, SDF-network parameters)."""
    g, mot, sdf = _stage1_fixture()
    n, S = int(g["n"]), int(g["S"])
    lv = {k: cu(v).requires_grad_(True) for k, v in unflatten(g, "in.").items()}
    out = {"color_fine": torch.zeros(n, 3, device=DEV), "normals": lv["normals"], "sdf_flows": lv["sdf_flows"],
           "sampled_points": lv["pts"], "weights": lv["weights"], "sdf": lv["sdf"]}      # a plain reference-shaped dict
    res = CL.stage1_losses(out, cu(g["rgb_gt"]), mot, sdf, float(g["query_time_step"]), int(g["image_idx"]),
                           [int(v) for v in g["ref_idx"]], int(g["nb_valid"]), int(g["total_nb_images"]),
                           int(g["nb_sample_timestep"]), cu(g["Kr"]), cu(g["scale"]), cu(g["norm_pix"]), cu(g["pix"]), cu(g["refs"]),
                           int(g["world_cam_idx"]), float(g["world_time_step"]))
    for k in ("sdf_loss", "flow_rgb_loss", "sdf_consistency_loss"):
        assert rel_err(res[k], g[k]) < 1e-4, (k, res[k].item(), g[k].item())
    assert_close(res["flow_fw_pred"], g["flow_fw_pred"], 1e-4, "flow_fw_pred")
    w = g["loss_weights"]
    (w[0] * res["sdf_loss"] + w[1] * res["flow_rgb_loss"] + w[2] * res["sdf_consistency_loss"]).backward()
    for k, v in lv.items():
        assert rel_err(v.grad, g[f"grad.{k}"]) < 1e-3, k
    for k, p in mot.named_parameters():
        assert rel_err(p.grad, g[f"grad.motion.{k}"]) < 2e-3, k
    for k, p in sdf.named_parameters():
        assert rel_err(p.grad, g[f"grad.sdfnet.{k}"]) < 1e-3, k


def test_flow_rgb_border_and_empty_mask():
    """Rays whose correspondence leaves the frame are masked out (and clamp to the border in the warp); with no valid ray the
    loss is 0 and so are the gradients."""
    torch.manual_seed(3)
    n, H, W = 40, 12, 16
    wp = torch.cat([torch.randn(n, 2) * 0.3, -2.0 - torch.rand(n, 1), torch.ones(n, 1)], dim=1)
    K = O.camera_matrix(0.8 * W, 0.8 * W, W, H)
    w2c = torch.eye(4).unsqueeze(0)
    pix = torch.stack([torch.randint(0, W, (n,)), torch.randint(0, H, (n,))], -1).float()
    npix = torch.stack([2 * pix[:, 0] / (W - 1) - 1, 2 * pix[:, 1] / (H - 1) - 1], -1)
    refs, gt = torch.rand(1, 3, H, W), torch.rand(n, 3)
    for shift in (0.0, 1e4):                       # 1e4: every correspondence far outside
        wpr = wp.clone().requires_grad_(True)
        pts_map = wpr[:, :3] + torch.tensor([shift, 0.0, 0.0])
        pm = (K[:3, :3] @ pts_map.T).T
        pm = pm[:, :2] / pm[:, 2:]
        d = pm - npix
        flow = torch.stack([d[:, 0] * (W / 2), d[:, 1] * (H / 2)], -1)
        lo = O.flow_rgb_loss([flow], pix, refs, gt)
        lo.backward()
        wpd = cu(wp).requires_grad_(True)
        w2 = w2c.clone(); w2[0, 0, 3] = shift
        ld = CL.flow_rgb_loss(wpd, cu(w2), cu(K[:3, :3]).unsqueeze(0), cu(npix), cu(pix), cu(refs), cu(gt))
        ld.backward()
        assert abs(ld.item() - lo.item()) < 1e-5 * max(1.0, abs(lo.item()))
        if shift:
            assert ld.item() == 0.0 and wpd.grad.abs().max().item() == 0.0
        else:
            assert rel_err(wpd.grad[:, :3], wpr.grad[:, :3]) < 1e-3


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_fused_step_matches_dict_path(small_params, prec):
    """NeuSRenderer.forward_losses (render + losses as one autograd node) against forward + the torch loss expressions on the
    same renderer: identical loss and parameter / pose gradients (same kernels underneath, different reduction order only)."""
    g = load_golden("step_small")
    K = cu(O.camera_matrix(0.8 * 80, 0.8 * 80, 80, 60).unsqueeze(0))
    grads = {}
    for fused in (False, True):
        r = renderer_from(small_params, SMALL_CFG)
        if prec == "bf16":
            r.sdf_network.precision = r.color_network.precision = C.PREC_BF16
        pose = C.PoseRetriever(1).to(DEV)
        with torch.no_grad():
            pose.r.copy_(g["r"]); pose.t.copy_(g["tr"])
        r.t_rand_override = g["t_rand"]
        loss, out, _ = C.training.render_train_step(r, pose, 0, cu(g["pix"]), K, torch.eye(4, device=DEV).unsqueeze(0),
                                                    cu(g["rgb_gt"]), cu(g["t"]), (0.01, 5.0), cos_anneal_ratio=0.5, fused=fused)
        grads[fused] = dict(loss=loss, color=out["color_fine"].detach(), r=pose.r.grad, t=pose.t.grad,
                            **{k: p.grad for k, p in r.named_parameters()})
        if fused:
            assert set(out.keys()) >= {"sdf", "color_fine", "depth_pred", "normals", "sdf_flows", "weights", "sampled_points",
                                       "inside_sphere", "weight_outside", "s_val", "loss_rgb", "loss_eikonal"}
            assert out["normals"].shape == (cu(g["pix"]).shape[1], 128, 3)
    tol = 1e-4 if prec == "fp32" else 2e-2        # bf16 wgrad tiles are reduced with L2 atomics: run-to-run order noise
    for k, v in grads[True].items():
        assert rel_err(v, grads[False][k]) < tol, (k, rel_err(v, grads[False][k]))


def test_fused_step_with_sdf_flow_term(small_params):
    """The SDF-flow loss inside the fused node: loss value and gradients (networks, pose, motion vector) against the oracle's
    train_step + sdf_flow_loss."""
    g = load_golden("step_small")
    Kc = O.camera_matrix(0.8 * 80, 0.8 * 80, 80, 60).unsqueeze(0)
    torch.manual_seed(8)
    mot = torch.randn(6) * 0.5
    Pg = {t: {k: v.clone().requires_grad_(True) for k, v in small_params[t].items()} for t in small_params}
    po = dict(r=g["r"].clone().requires_grad_(True), t=g["tr"].clone().requires_grad_(True), init_c2w=torch.eye(4).unsqueeze(0))
    mo = mot.clone().requires_grad_(True)
    lo, aux = O.train_step(Pg, po, g["pix"], Kc, torch.eye(4).unsqueeze(0), g["rgb_gt"], g["t"], [0.01, 5.0], cos_anneal=0.5,
                           t_rand=g["t_rand"])
    lo = lo + 0.5 * O.sdf_flow_loss(aux["out"], mo[:3], mo[3:])
    lo.backward()
    r = renderer_from(small_params, SMALL_CFG)
    pose = C.PoseRetriever(1).to(DEV)
    with torch.no_grad():
        pose.r.copy_(g["r"]); pose.t.copy_(g["tr"])
    r.t_rand_override = g["t_rand"]
    md = cu(mot).requires_grad_(True)
    loss, out, _ = C.training.render_train_step(r, pose, 0, cu(g["pix"]), cu(Kc), torch.eye(4, device=DEV).unsqueeze(0),
                                                cu(g["rgb_gt"]), cu(g["t"]), (0.01, 5.0), cos_anneal_ratio=0.5, sdf_weight=0.5,
                                                motion=md)
    assert rel_err(loss, lo) < 1e-4
    assert rel_err(md.grad, mo.grad) < 1e-3
    for tag, net in (("sdf", r.sdf_network), ("color", r.color_network), ("variance", r.deviation_network)):
        for k, p in net.named_parameters():
            assert rel_err(p.grad, Pg[tag][k].grad) < 1e-3, (tag, k)
    assert rel_err(pose.r.grad, po["r"].grad) < 2e-3 and rel_err(pose.t.grad, po["t"].grad) < 2e-3


def test_process_data_on_device():
    """cope_sample_pixels against the reference's process_data arithmetic (model/training.py:413-471): with the reference's
    own CPU randperm prefix as corners the pixel ids / coordinates / colours are bit-exact; with device-drawn corners the
    patches are distinct, in range and row-major."""
    h, w, ps, n = 60, 80, 4, 256
    torch.manual_seed(12)
    img = torch.rand(1, 3, h, w)
    K = O.camera_matrix(0.8 * w, 0.8 * w, w, h).unsqueeze(0)
    torch.manual_seed(99)
    idx_ref = O.patch_indices(h, w, ps, n)
    torch.manual_seed(99)
    corners = torch.randperm((h - ps + 1) * (w - ps + 1))[:n // ps ** 2]
    loc, sc = O.pixel_grid(h, w)
    world = torch.eye(4, device=DEV)
    pix, npix, o, d, dn, rgb, idx = C.training.process_data(cu(img), cu(K), world, torch.eye(4, device=DEV).unsqueeze(0), n, ps,
                                                          corners=corners)
    assert torch.equal(idx.cpu(), idx_ref)
    assert torch.equal(npix.cpu(), sc[0, idx_ref]) and torch.equal(pix.cpu(), loc[0, idx_ref].float())
    assert torch.equal(rgb.cpu(), img.view(3, h * w).t()[idx_ref])
    ro, rd, rn = O.ray_generation(sc[:, idx_ref], K, torch.eye(4), torch.eye(4).unsqueeze(0))
    assert_close(d, rd, 1e-5); assert_close(o, ro, 1e-6)
    for hh, ww, pp, nn in ((717, 1275, 4, 1024), (9, 7, 3, 45), (5, 5, 1, 25), (4, 4, 4, 16)):
        im = torch.rand(3, hh, ww, device=DEV)
        seen = set()
        for seed in (0, 1, 12345678901234):
            pix, npix, *_, rgb, idx = C.training.process_data(im, cu(O.camera_matrix(ww, ww, ww, hh).unsqueeze(0)), world,
                                                              torch.eye(4, device=DEV).unsqueeze(0), nn, pp, seed=seed)
            k = idx.shape[0] // pp ** 2
            assert k == min(nn // pp ** 2, (hh - pp + 1) * (ww - pp + 1))
            pt = idx.view(k, pp, pp).cpu()
            corner = pt[:, 0, 0]
            assert corner.unique().numel() == k, "corners must be distinct"
            r0, c0 = corner // ww, corner % ww
            assert (r0 <= hh - pp).all() and (c0 <= ww - pp).all()
            off = torch.arange(pp).view(1, pp, 1) * ww + torch.arange(pp).view(1, 1, pp)
            assert torch.equal(pt, corner.view(-1, 1, 1) + off)
            assert torch.equal(rgb, im.view(3, -1).t()[idx])
            seen.add(tuple(corner.tolist()))
        if (hh - pp + 1) * (ww - pp + 1) > 1 and k < (hh - pp + 1) * (ww - pp + 1):
            assert len(seen) > 1, "different seeds must give different patches"
