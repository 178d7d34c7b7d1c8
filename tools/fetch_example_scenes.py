#!/usr/bin/env python3
"""Copies the two example scenes BASELINE.json's configs name (inputs, not code) from the reference
checkout into tests/golden/ so that GPU-box tests, smoke() and bench.py can run without /root/reference."""
import os
import shutil

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for name in ("cornell_box.psy", "cube.psy"):
    shutil.copyfile(os.path.join("/root/reference/example_scenes", name), os.path.join(REPO, "tests", "golden", name))
print("ok")
