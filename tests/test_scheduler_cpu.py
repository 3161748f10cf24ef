"""Host logic of multi-task pretraining on CPU: the task order mirrors the reference's BatchSchedulerSampler
(datasets/multi_task_scheduler.py:34-80; fixture tests/golden/task_schedule_golden.json produced from the unmodified
reference by tools/make_golden.py --schedule), and the per-step segment selection of the flat gradient buffer."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def test_task_order_matches_reference_sampler():
    from ctrlora_b200.scheduler import TaskSchedule
    g = json.load(open(os.path.join(GOLD, "task_schedule_golden.json")))
    for case in g["cases"]:
        np.random.seed(case["seed"])
        sched = TaskSchedule(case["tasks"], case["largest"], case["batch_size"], shuffle=case["shuffle"])
        got = list(sched)
        assert got == case["task_per_batch"], case
        assert len(sched) == len(case["task_per_batch"])


def test_active_segments():
    from ctrlora_b200.train import active_segments
    layout = {"base": (0, 100), "tasks": ["canny", "depth", "seg"], "lora": {"canny": (100, 10), "depth": (110, 10), "seg": (120, 10)}}
    assert active_segments(layout, ["depth"]) == [(0, 100, "base"), (110, 10, "depth")]
    # ranks on different tasks (the reference's permutation is per-rank, multi_task_scheduler.py:59): union, task order
    assert active_segments(layout, ["seg", "canny", "seg", "canny"]) == [(0, 100, "base"), (100, 10, "canny"), (120, 10, "seg")]
