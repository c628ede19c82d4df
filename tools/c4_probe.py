"""BASELINE config C4 alone (ATSP n=1000, batch 64, 100 starts, greedy; per-step key-streaming path):
   python tools/c4_probe.py            # wall-clock of one rollout
   ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 120 --csv python tools/c4_probe.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.config_sweep import run  # noqa: E402

if __name__ == "__main__":
    run("atsp", 1000, int(os.environ.get("B", 64)), 1, 100, "greedy", reps=1)
