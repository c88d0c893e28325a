"""bench.py's reference arm is the CPU oracle and nothing else: it must not import the product package or map its CUDA
library (the driver records which .so files each arm's process loaded)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_never_maps_the_product_library():
    code = ("import sys, bench; r = bench.cpu_reference(1, 1, 1); maps = open('/proc/self/maps').read(); "
            "print('MAPPED' if 'liblicos_b200' in maps else 'CLEAN', 'licos_b200' in sys.modules, r['kind'])")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.split()[-3:] == ["CLEAN", "False", "port"], out.stdout


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2"], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
