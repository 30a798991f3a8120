import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import topk_perf as t
t.run('bpr', 37888, 2_000_000, 128, reps=1)
