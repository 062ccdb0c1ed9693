"""DepthWarpingLayer forward / backward HBM microbenchmark alone (bench.py's warp_layer arm): python tools/time_warp.py"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
print(json.dumps(bench.warp_layer_bench(torch.device("cuda", 0), bench.measured_peaks())))
