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
    cfg.replan_risk_rate = 0.1  # resampled peer particles carry fractional risks: u8 cells are refused
    cfg.cell_dtype = pop.POP_CELL_U8
    with pytest.raises(pop.PopError) as e:
        pop.PopEngine(cfg)
    assert e.value.code == -6
    cfg.cell_dtype = pop.POP_CELL_F32
    cfg.num_resample = 17
    with pytest.raises(pop.PopError) as e:
        pop.PopEngine(cfg)
    assert e.value.code == -1


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle."""
    pkg = os.path.join(ROOT, "pred_occ_planner_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.lower(), f"{f} mentions the oracle"


def test_default_configs_carry_the_reference_values(pop):
    """Defaults that need no device: the particle map's compile-time constants (map_parameters.h), the raycast
    map's "grid_map/..." parameters and the simulator camera."""
    cfg = pop.default_config()
    d = pop.PopDspConfig()
    pop.load().pop_dsp_default_config(C.byref(cfg), C.byref(d))
    assert (d.half_fov_h, d.half_fov_v, d.angle_resolution, d.max_particle_num_voxel) == (43, 29, 1, 7)
    V = cfg.nx * cfg.ny * cfg.nz
    assert d.pyramid_capacity == int(V * 7 + 1e5) // (360 * 180) * 2 == 20        # SAFE_PARTICLE_NUM_PYRAMID
    assert d.max_points == 5000 and d.new_born_particle_number_each_point == 20 and d.max_points_per_pyramid == 100
    assert abs(d.new_born_particle_weight - 1e-4) < 1e-9 and abs(d.prediction_future_time[5] - 1.8) < 1e-6
    g = pop.PopGridmapConfig()
    pop.load().pop_gridmap_default_config(C.byref(g))
    assert (g.p_hit, g.p_miss, g.p_min, g.p_max, g.p_occ) == (0.70, 0.35, 0.12, 0.97, 0.80)
    cam = pop._lib.PopCamera()
    pop.load().pop_camera_default(C.byref(cam), 640, 480)
    assert abs(cam.fx - 387.229248046875) < 1e-4 and cam.sensing_horizon == 6.0
    assert list(cam.cam2body) == [0, 0, 1, 0, -1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 0, 1]   # pcl_render_node.cpp:338


@pytest.mark.skipif(_has_gpu(), reason="only meaningful without a CUDA device")
def test_raycast_map_has_no_cpu_fallback(pop):
    from pred_occ_planner_b200 import gridmap
    with pytest.raises(pop.PopError) as e:
        gridmap.GridMaps(n_maps=1)
    assert e.value.code == -2  # POP_E_NO_DEVICE


def test_oracle_default_config_equals_libpops(pop, oracle):
    """bench.py's CPU arm takes its configuration from the oracle (it must not load libpop): both statements of the
    reference's simulator defaults are the same bytes."""
    from tests import oracle_lib
    a, b = pop.default_config(), oracle_lib.default_config()
    assert bytes(a) == bytes(b)
