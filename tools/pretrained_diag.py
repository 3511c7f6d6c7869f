"""Diagnostic: bf16 parameter-gradient cosine per tensor with pretrained_sdf weights, split by loss term (rgb only / eikonal only / both)
and by ray count.  python tools/pretrained_diag.py"""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
import cope_nerf_b200 as C, oracle as O
from conftest import cos_sim, rel_err
from test_gpu_round2 import _pretrained_params
from test_gpu_parity import renderer_from, cu

g, P = _pretrained_params()
Kc = O.camera_matrix(0.8 * 1275, 0.8 * 1275, 1275, 717).unsqueeze(0)
import os
print('COPE_NO_FUSED =', os.environ.get('COPE_NO_FUSED'))
for n in (int(os.environ.get('DIAG_N', '256')),):
    torch.manual_seed(23)
    pix = (torch.rand(1, n, 2) * 2 - 1) * 0.8
    pix[0, :, 1] = pix[0, :, 1].abs()
    rgb_gt, t, t_rand = torch.rand(n, 3), torch.tensor([0.3]), torch.rand(n, 64)
    r0, t0 = torch.randn(1, 3) * 0.05, torch.randn(1, 3) * 0.05
    for w_rgb, w_eik in ((0.33333, 0.0), (0.0, 0.1), (0.33333, 0.1)):
        Pg = {k: {a: v.clone().requires_grad_(True) for a, v in P[k].items()} for k in P}
        po = dict(r=r0.clone().requires_grad_(True), t=t0.clone().requires_grad_(True), init_c2w=torch.eye(4).unsqueeze(0))
        lo, aux = O.train_step(Pg, po, pix, Kc, torch.eye(4).unsqueeze(0), rgb_gt, t, [0.01, 5.0], cos_anneal=0.5, t_rand=t_rand,
                               rgb_weight=w_rgb, eikonal_weight=w_eik)
        lo.backward()
        res = {}
        for prec in (C.PREC_FP32, C.PREC_BF16):
            r = renderer_from(P, C.training.DEFAULT_CFG)
            r.sdf_network.precision = r.color_network.precision = prec
            pose = C.PoseRetriever(1).to('cuda')
            with torch.no_grad():
                pose.r.copy_(r0); pose.t.copy_(t0)
            r.t_rand_override = t_rand
            loss, out, _ = C.training.render_train_step(r, pose, 0, cu(pix), cu(Kc), torch.eye(4, device='cuda')[None], cu(rgb_gt), cu(t),
                                                        (0.01, 5.0), rgb_weight=w_rgb, eikonal_weight=w_eik)
            cs = {k: cos_sim(p.grad, Pg['sdf'][k].grad) for k, p in r.sdf_network.named_parameters() if Pg['sdf'][k].grad.abs().max() > 0}
            res[prec] = cs
            if prec == C.PREC_BF16:
                oo = aux['out']
                fe = {k: rel_err(out[k], oo[k]) for k in ('sdf', 'normals', 'sdf_flows', 'weights', 'color_fine', 'depth_pred')}
                print("   forward rel err:", {k: f"{v:.2e}" for k, v in fe.items()}, "sdf abs mean", f"{float((out['sdf'].cpu() - oo['sdf']).abs().mean()):.2e}")
                nrm = out['normals'].reshape(-1, 3).norm(dim=-1)
                on = aux['out']['normals'].reshape(-1, 3).norm(dim=-1)
                print(f"   |n|-1: oracle mean {float((on - 1).abs().mean()):.4e}  bf16-vs-oracle |n| abs err mean {float((nrm.cpu() - on).abs().mean()):.4e}")
        low = sorted(res[C.PREC_BF16].items(), key=lambda kv: kv[1])[:6]
        print(f"n={n} w_rgb={w_rgb} w_eik={w_eik}: fp32 min cos {min(res[C.PREC_FP32].values()):.6f}; bf16 lowest:",
              ", ".join(f"{k} {v:.5f}" for k, v in low))
