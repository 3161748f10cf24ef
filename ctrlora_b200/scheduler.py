"""Task order of multi-task pretraining: a host-side mirror of the reference's BatchSchedulerSampler
(datasets/multi_task_scheduler.py:34-80), reduced to what the training step consumes -- WHICH task each mini-batch
belongs to.  The reference walks rounds; in every round it draws `np.random.permutation(n_tasks)` (un-seeded, per rank)
when shuffling, else `arange`, and emits one whole mini-batch per task in that order, `ceil(largest / batch_size)` rounds
per epoch.  Sample indices inside a batch come from the per-task samplers and do not matter to the kernels (inputs are
synthetic here); the task does: it selects the LoRA set (`switch_lora`) and the all-reduce / AdamW segments of the step.
"""
import math

import numpy as np


class TaskSchedule:
    def __init__(self, tasks, largest_dataset_size, batch_size, shuffle=True):
        self.tasks = list(tasks)
        self.rounds = math.ceil(largest_dataset_size / batch_size)   # range(0, largest * n, batch * n) in the reference
        self.shuffle = shuffle

    def __len__(self):
        return self.rounds * len(self.tasks)

    def __iter__(self):
        """Yields the task name of every mini-batch of one epoch (consumes np.random exactly like the reference)."""
        n = len(self.tasks)
        for _ in range(self.rounds):
            perm = np.random.permutation(n) if self.shuffle else np.arange(n)
            for i in perm:
                yield self.tasks[int(i)]
