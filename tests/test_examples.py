"""CPU: the example scripts run end to end (on the emulated backend: `ms.make` is patched to inject the test world)."""
import json
import os
import runpy
import sys

import pytest

import maniskill_b200 as ms
from emu_world import EmuBackendWorld

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_demo_random_action_records_a_rollout(tmp_path, monkeypatch, capsys):
    pytest.importorskip("cv2")
    make = ms.make
    monkeypatch.setattr(ms, "make", lambda *a, **k: make(*a, world_factory=EmuBackendWorld, **k))
    monkeypatch.setattr(sys, "argv", ["demo_random_action.py", "-e", "PickCube-v1", "-n", "2", "-o", "state", "--record-dir", str(tmp_path),
                                      "--render-mode", "sensors", "-s", "1"])
    runpy.run_path(os.path.join(ROOT, "examples", "demo_random_action.py"), run_name="__main__")
    out = capsys.readouterr().out
    assert "50 control steps" in out and "success rate" in out
    meta = json.load(open(tmp_path / "trajectory.json"))
    assert meta["env_info"]["env_id"] == "PickCube-v1" and [e["elapsed_steps"] for e in meta["episodes"]] == [50, 50]
    assert (tmp_path / "0.mp4").stat().st_size > 1000
