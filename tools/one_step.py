"""Profiling helper: N eager training steps (1024 rays, bf16 path) for `ncu --metrics gpu__time_duration.sum` launch lists:
    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file launches.csv python tools/one_step.py 2"""
import sys, torch
sys.path.insert(0, '.')
import cope_nerf_b200 as C, bench
dev = torch.device('cuda')
torch.manual_seed(678)
rnd = C.training.build_networks(device=dev, precision=C.PREC_BF16)
pose = C.PoseRetriever(1).to(dev)
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
host = bench.synth_inputs(1024, nsteps, 678)
res = [{k: v.to(dev) for k, v in b.items()} for b in host]
Kc, Sc = bench.camera().to(dev), torch.eye(4, device=dev)[None]
t0 = torch.zeros(1, device=dev)
for b in res:
    rnd.zero_grad(); pose.zero_grad()
    rnd.t_rand_override = b['t_rand']
    C.training.render_train_step(rnd, pose, 0, b['pix'], Kc, Sc, b['rgb'], t0, (0.01, 5.0))
torch.cuda.synchronize()
print("ok")
