"""Runs the capture-under-garbage-collection regression test with the guard removed: the test must FAIL then
(evidence that kliff_b200.utils.capture_without_gc is what keeps the capture valid).
Usage: python scripts/debug/capture_guard_off.py"""
import contextlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import pytest

import kliff_b200.legacy.fused as fused
import kliff_b200.legacy.loss as loss

fused.capture_without_gc = contextlib.nullcontext
loss.capture_without_gc = contextlib.nullcontext
rc = pytest.main(["tests/test_gpu_api.py", "-q", "-k", "eager_garbage_collector", "--timeout", "60", "-p", "no:cacheprovider"])
print("pytest rc with the guard removed:", int(rc))
