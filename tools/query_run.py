"""Profiling helper: three no-grad SDF queries of P points (default 65536) through SDFNetwork.query_flat: python tools/query_run.py [P]"""
import sys, torch
sys.path.insert(0, '.')
import cope_nerf_b200 as C
dev = torch.device('cuda'); torch.manual_seed(678)
rnd = C.training.build_networks(device=dev, precision=C.PREC_BF16)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
x = torch.cat([torch.randn(P, 3, device=dev) * 0.6, torch.zeros(P, 1, device=dev)], -1)
flat = rnd.sdf_network.flat_weights().detach()
for _ in range(3): rnd.sdf_network.query_flat(flat, x)
torch.cuda.synchronize(); print("ok")
