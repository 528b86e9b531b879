"""A short fixed-seed run of tools/fuzz_tree_host.py: random tree parameters, sizes around every boundary and tiny rings,
host memory / file / tee / sharded forms of the tree digest against the oracle, through the CPU test double."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tree_host_logic_random_cases(mock_lib):
    env = dict(os.environ, FUZZ_ITERS="40", FUZZ_SEED="4")
    env.pop("MXD_MOCK_SANITIZE", None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_tree_host.py")], capture_output=True, text=True,
                         timeout=900, env=env, cwd=ROOT)
    assert out.returncode == 0 and "0 mismatches" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
