"""CPU: bench.py parses its contract flags and describes its workload without touching a GPU."""
import importlib.util
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_help_lists_the_contract_flags():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--workload", "--numbering"):
        assert flag in r.stdout


def test_workload_description_and_stage_bytes():
    b = _bench()
    args = types.SimpleNamespace(workload="box", gas="air5", gpus=8, ppc=31, numbering="morton")
    cfg = b.workload_config(args, 8_000_000, 248_000_000)
    assert "BASELINE configs[4]" in cfg["workload"] and cfg["partition"] == "2x2x2 bricks" and "z-order" in cfg["cell_numbering"]
    assert b.procs_for(4) == (2, 2, 1)
    # 96 + 168 + 71 + 40 = 375 B per parcel-step for 5-species air
    assert sum(b.STAGE_BYTES_AIR.values()) == 375.0
    cp = b.case_parameters("air5", 200, 31)
    assert abs(cp["fnum"] * 31 - 1e20 * 0.004 ** 3) < 1e-6 * 1e20 * 0.004 ** 3


def test_capsule_workload_description():
    b = _bench()
    args = types.SimpleNamespace(workload="capsule", gas="air5", gpus=8, ppc=31, cells=200, numbering="morton")
    cfg = b.workload_config(args, 8_000_000, 250_000_000)
    assert "BASELINE configs[3]" in cfg["workload"] and cfg["partition"].startswith("2x2x2 bricks") and cfg["gas"] == "air5"
