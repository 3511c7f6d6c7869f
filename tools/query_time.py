"""Profiling helper: CUDA-event time of the no-grad SDF query chain at a few sizes: python tools/query_time.py [P ...]"""
import sys, torch
sys.path.insert(0, '.')
import cope_nerf_b200 as C
dev = torch.device('cuda'); torch.manual_seed(678)
rnd = C.training.build_networks(device=dev, precision=C.PREC_BF16)
flat = rnd.sdf_network.flat_weights().detach()
for P in [int(a) for a in sys.argv[1:]] or [16384, 65536]:
    x = torch.cat([torch.randn(P, 3, device=dev) * 0.6, torch.zeros(P, 1, device=dev)], -1)
    for _ in range(5): rnd.sdf_network.query_flat(flat, x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): rnd.sdf_network.query_flat(flat, x)
    e1.record(); torch.cuda.synchronize()
    print(f"P={P}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per query (incl. the weight pack launch)")
