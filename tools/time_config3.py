"""configs[3] of the bench alone: 20 models x one 200k-point scene through ppf_registration on one GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import objective_slam_b200 as ppf
for i in range(2):
    print({k: v for k, v in bench.config3(ppf, torch, bench.load_synth(), None, 1).items() if k != "workload"})
