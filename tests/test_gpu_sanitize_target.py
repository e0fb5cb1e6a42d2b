"""tools/sanitize_target.py — the small-shape pass over every kernel variant that is meant to run under
compute-sanitizer — must at least run clean and correct on its own (compute-sanitizer itself is closed on this pool)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sanitize_target_runs_clean():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_target.py")], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "sanitize target done" in out.stdout and "ok=False" not in out.stdout
