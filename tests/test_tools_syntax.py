"""The measurement scripts (bench.py, tools/*.py, __graft_entry__.py) only run on the GPU box; here they must at least parse
and bench.py must expose the flags the driver passes."""
import ast
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scripts_parse():
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")] + sorted(glob.glob(os.path.join(ROOT, "tools", "*.py")))
    assert len(files) >= 8
    for f in files:
        ast.parse(open(f).read(), filename=f)


def test_bench_flags():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out.stdout
