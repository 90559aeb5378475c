"""Records what the reference's own raycast map (plan_env/src/grid_map.cpp compiled unmodified into
oracle/_ref/libref_gridmap.so) does with five depth-camera frames, with and without its depth filter: the projected
points of every frame, and SHA-256 digests of occupancy_buffer_ (log-odds doubles), occupancy_buffer_inflate_ and the
local bounds after every frame, plus the log-odds of a fixed sample of voxels.  -> tests/golden/ref_gridmap_pins.npz
Run in the container that has /root/reference (make -C oracle first)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pred_occ_planner_b200 import gridmap
from tests import ref_gridmap_lib as rg

out = {}
cam, cloud, cfg = rg.camera(), rg.scene(), rg.config()
rng = np.random.default_rng(5)
for use_filter in (1, 0):
    dp = gridmap.depth_params(cam.height, cam.width, cam.fx, cam.fy, cam.cx, cam.cy, use_filter=bool(use_filter),
                              skip_pixel=2, margin=2, mindist=0.2, maxdist=5.0)
    ref = rg.RefGridMap(cfg, dp)
    sample = rng.integers(0, int(np.prod(ref.dims)), 4096)
    out[f"f{use_filter}_sample"] = sample
    for k in range(rg.N_FRAMES):
        img, R, pos = rg.frame(cam, cloud, k)
        pts = ref.project(img, R, pos)
        ref.raycast()
        occ, inf, b = ref.buffers()
        out[f"f{use_filter}_k{k}_pts"] = pts
        out[f"f{use_filter}_k{k}_bounds"] = b
        out[f"f{use_filter}_k{k}_occ_sha"] = np.frombuffer(bytes.fromhex(rg.digest(occ)), np.uint8)
        out[f"f{use_filter}_k{k}_inf_sha"] = np.frombuffer(bytes.fromhex(rg.digest(inf)), np.uint8)
        out[f"f{use_filter}_k{k}_occ_sample"] = occ.ravel()[sample]
        print("filter", use_filter, "frame", k, "points", len(pts), "occupied", int((occ > np.log(cfg.p_occ / (1 - cfg.p_occ))).sum()),
              "inflated", int(inf.sum()), "bounds", b.tolist())
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_gridmap_pins.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
