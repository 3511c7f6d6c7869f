#!/usr/bin/env python
"""CPU timing of the oracle port of the reference's MotionNetwork pose loop (model/neus_fields.py:142-183): relative poses of
n_img - 1 consecutive frame pairs x n_sub sub-steps, chained into world -> camera maps, forward + backward.  The number that
sits beside bench_kernels.py's `motion_pose_chain_fwd_bwd` row (profiles/README.md).

    python tests/time_motion_oracle.py [n_img] [n_sub]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n_sub = int(sys.argv[2]) if len(sys.argv) > 2 else 10
torch.manual_seed(5)
P = {k: v.requires_grad_(True) for k, v in O.init_motion_params(**O.MOTION_CFG).items()}
wgt = torch.randn(n_img, 4, 4)
t0 = time.perf_counter()
(O.w2c_mappings(O.relative_camera_pose(P, 0, n_img - 1, n_img, n_sub)[1]) * wgt).sum().backward()
print(f"oracle port: {n_img - 1} pairs x {n_sub} sub-steps, forward + backward: {(time.perf_counter() - t0) * 1e3:.1f} ms "
      f"on {os.cpu_count()} host cores")
