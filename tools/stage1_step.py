"""Profiling helper: a few eager stage-1 iterations (bench.stage1_step_ms without the graph) for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/stage1_step.py"""
import sys, torch
sys.path.insert(0, '.')
import cope_nerf_b200 as C, bench
ms, note = bench.stage1_step_ms(C, torch.device('cuda'), C.PREC_BF16, 1024, 1, 0, use_graph=False)
print("ok", ms, note)
