"""The library's own out-of-bounds net (compute-sanitizer is closed on the GPU pool: profiles/r2_sanitizer.md).  With TMQ_GUARD_BYTES set,
every device allocation of libtmq carries NaN-filled red zones (csrc/tmq_internal.h): an out-of-bounds or uninitialised READ poisons a
result and fails the oracle comparison inside tools/sanitize_target.py, an out-of-bounds WRITE is found by tmq_guard_check().  The target
runs the whole hot path on 8^3 x 16, plainly and with the ghost-zone path forced in every halo mode and partition."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hot_path_under_guarded_allocations():
    env = dict(os.environ, TMQ_GUARD_BYTES="4096")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_target.py"), "all", "4"], capture_output=True, text=True, env=env, timeout=1500)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.count("sanitize target:") == 16                      # plain + 5 modes x 3 partitions
    assert "guard check: 0 corrupted allocations" in p.stdout and "TMQ_GUARD_BYTES=4096" in p.stdout


def test_guard_detects_an_out_of_bounds_write():
    """the net is not vacuous: a deliberate write one element past a device buffer is reported"""
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import tmq
c = tmq.Context((4, 4, 4, 4))
p = c.dev_malloc(1024)
c.h2d(p, np.zeros(129))          # 1032 bytes into a 1024-byte allocation
c.sync()
bad, msg = tmq.guard_check()
print("BAD", bad, msg)
''' % (os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200"), os.path.join(ROOT, "tests"))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, TMQ_GUARD_BYTES="4096"), timeout=300)
    assert "BAD 1 out-of-bounds write behind an allocation of 1024 bytes" in p.stdout, p.stdout + p.stderr
