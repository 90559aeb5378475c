"""The C-ABI library loads on a CPU-only box and exports every symbol include/pop.h declares;
without a GPU it refuses to create a handle instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import pytest

from tests.conftest import ROOT, _has_gpu


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pop.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pop_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_exported(pop):
    lib = pop.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libpop.so does not export {n}"
    bound = {s[0] for s in pop._lib.SYMBOLS}
    assert bound == set(names), f"ctypes table and header differ: {bound ^ set(names)}"


def test_struct_layouts(pop):
    assert C.sizeof(pop._lib.PopTrajRecord) == 1536
    assert pop.TRAJ_DTYPE.itemsize == 1536
    assert C.sizeof(pop._lib.PopCylinder) == pop.CYL_DTYPE.itemsize == 96
    cfg = pop.default_config()
    assert cfg.abi_version == pop.load().pop_abi_version() == 1
    # defaults are the reference's simulator parameters
    assert (cfg.nx, cfg.ny, cfg.nz, cfg.nt) == (66, 66, 20, 9)
    assert abs(cfg.resolution - 0.15) < 1e-7 and cfg.check_num == 1 and cfg.allocate_num == 10000


@pytest.mark.skipif(_has_gpu(), reason="only meaningful without a CUDA device")
def test_no_cpu_fallback(pop):
    with pytest.raises(pop.PopError) as e:
        pop.PopEngine(pop.default_config())
    assert e.value.code == -2  # POP_E_NO_DEVICE


def test_invalid_config_rejected(pop):
    cfg = pop.default_config()
    cfg.nx = 0
    with pytest.raises(pop.PopError) as e:
        pop.PopEngine(cfg)
    assert e.value.code == -1
    cfg = pop.default_config()
    cfg.replan_risk_rate = 0.1  # Gaussian resampling of peers is not reproducible → refused
    with pytest.raises(pop.PopError) as e:
        pop.PopEngine(cfg)
    assert e.value.code == -6


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle."""
    pkg = os.path.join(ROOT, "pred_occ_planner_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.lower(), f"{f} mentions the oracle"
