"""Runs bench.py's 512x512 sampling leg alone: python tools/sampling_bench.py [batch] [repeats]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    print(json.dumps(bench.sampling_leg(torch.device("cuda", 0), batch=batch)))
