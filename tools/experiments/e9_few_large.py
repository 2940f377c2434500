"""Runs only bench.py's few-large-members leg (the silesia shape): python tools/experiments/e9_few_large.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
h = bench.Harness(0)
print(json.dumps(bench.run_few_large(h)))
print("parallel / fallback streams:", h.ctx.parallel_streams)
