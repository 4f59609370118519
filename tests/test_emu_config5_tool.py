"""tools/config5_run.py (the device-side generator + row-sharded step of BASELINE config 5) executed end to end on the CPU
emulator at toy size, one rank: the edge-list construction, the eight operands of the two RowBlockGraphs, the parameter / feature
shards and one eager step.  (The tool itself only runs on GPUs; a typo in it costs GPU-minutes.)"""
import json
import os
import runpy
import sys

import pytest
import torch

from tests.cuemu import harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, *a):
        pass

    def elapsed_time(self, other):
        return 1.0


def test_config5_tool_runs_at_toy_size(monkeypatch, capsys):
    harness.set_order("fwd")
    harness.emulated_device(monkeypatch)
    real_gen = torch.Generator
    monkeypatch.setattr(torch, "Generator", lambda device=None: real_gen())
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "max_memory_allocated", lambda d=None: 0)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    for k, v in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")):
        monkeypatch.setenv(k, v)
    monkeypatch.setattr(sys, "argv", ["config5_run.py", "--users", "300", "--items", "120", "--edges", "2500", "--embed-size", "64",
                                      "--dv", "64", "--dt", "64", "--steps", "1", "--batch", "64", "--no-graph"])
    runpy.run_path(os.path.join(ROOT, "tools", "config5_run.py"), run_name="__main__")
    out = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    res = json.loads(out[-1])
    assert res["n_gpus"] == 1 and len(res["loss_first_step"]) == 5
    assert all(x == x and abs(x) < 1e6 for x in res["loss_first_step"]) and res["loss_first_step"][0] > 0
