"""Runs tests/test_gpu_zz_reference_vectors.py's check without pytest/torch (seconds on a fresh GPU box):
    python scripts/dev_ref_vectors_gpu.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
t0 = time.time()
import pred_occ_planner_b200 as p  # noqa: E402
from tests import test_gpu_zz_reference_vectors as t  # noqa: E402

for cell in (p.POP_CELL_F32, p.POP_CELL_U8):
    t.test_astar_batch_equals_reference_outputs(cell)
    print(f"cell {cell}: CUDA A* == reference outputs on 40 problems ({time.time() - t0:.1f} s)", flush=True)
t.test_replan_cycle_against_the_reference_planner()
print(f"pop_replan_cycle agrees with the reference planner's recorded outputs on 24 agents ({time.time() - t0:.1f} s)", flush=True)
