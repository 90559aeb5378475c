"""bench.py's reference arm (the CPU oracle timed on the host cores) runs without a GPU and prints exactly one
JSON line with the contract's keys; non-zero ranks of a multi-process launch print nothing."""
import json
import os
import subprocess
import sys

from tests.conftest import ROOT


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--small", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_prints_one_contract_line():
    out = _run()
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "agent-replans/s"
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None and d["scaling"] == "weak"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "agents" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.strip() == ""


def test_reference_arm_does_not_load_libpop():
    """The CPU arm is the checker's code only: configuration from the oracle, world from the pure-Python generator —
    libpop.so (the product) is never mapped into that process."""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--small', '--steps', '1', '--warmup', '0'];"
            "runpy.run_path(%r, run_name='__main__');"
            "maps = open('/proc/self/maps').read(); assert 'liboracle.so' in maps; assert 'libpop.so' not in maps" % os.path.join(ROOT, "bench.py"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
