"""Small end-to-end pass over the kernels (ground-truth SOGM with byte cells, A*, corridors, QP, separation, FSM,
isTrajSafe; sensor, particle map, raycast map, trajectory sampling) meant to be run under
`compute-sanitizer --tool memcheck` (development aid)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import pred_occ_planner_b200 as p
from pred_occ_planner_b200 import dsp, gridmap, worldgen
from tests import cases

# ---- ground-truth map + planning cycle, byte cells, agents near the map edge (clipped footprints) ----
n = 6
c2 = cases.cfg_small_c2(p.POP_CELL_U8, n_local=n, n_total=n)
e2 = p.PopEngine(c2, device=0)
cl, cy = cases.world_moving(5, (12.0, 12.0, 3.0), 4, 6)
e2.set_world_cloud(cl); e2.set_cylinders(cy)
e2.set_body_particles(*worldgen.body_table(n, [0.4, 0.4, 0.45]))
e2.set_trajectories(np.zeros(n, p.TRAJ_DTYPE))
s0, g0 = worldgen.circle_agents(n, 4.5)
for k in range(3):
    if k > 0:
        e2.swarm_advance(10.0 + 0.1 * k); s0 = e2.swarm_states()
    st, tr = e2.replan_cycle(s0, g0, s0[:, :3].astype(np.float32), 10.0 + 0.1 * k, 10.02 + 0.1 * k)
    e2.traj_publish_local()
    print("cycle", k, st.tolist())
print("traj safe", e2.is_traj_safe(10.25, 1.0)[0].tolist())
m0 = e2.agent_map(0)
edge = np.array([[s0[0, 0] + 5.9, s0[0, 1], 1.0], [s0[0, 0] - 5.9, s0[0, 1] + 5.9, 1.0], [s0[0, 0], s0[0, 1], 1.0]])
print("edge queries", m0.getClearOcccupancy(edge, 0).tolist(), m0.getInflateOccupancy(edge).tolist())
e2.close()

cfg = cases.dsp_cfg_ref(n_local=2, n_total=2)
d = cases.dsp_params(cfg, n_randoms=1 << 16)
eng = p.PopEngine(cfg, device=0)
maps = dsp.DspMaps(eng, d)
cam = maps.set_camera(96, 72)
cloud, cyl = worldgen.make_world(11, arena=(10.0, 10.0, 3.0), n_static=6, n_moving=0, height=3.0)
gx, gy = np.meshgrid(np.arange(-5, 5, 0.15), np.arange(-5, 5, 0.15), indexing="ij")
cloud = np.concatenate([cloud, np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size)], axis=1)]).astype(np.float32)
eng.set_world_cloud(cloud)
for k in range(3):
    pos = np.array([[-4.0 + 0.1 * k, -0.5, 1.0], [-4.0 + 0.1 * k, 0.8, 1.1]])
    quat = np.array([[1.0, 0, 0, 0], [np.cos(0.1), 0, 0, np.sin(0.1)]])
    counts, depth = maps.render(pos, quat, want_depth=True)
    ok = maps.cloudCallbackRendered(pos.astype(np.float32), quat.astype(np.float32), 10.0 + 0.1 * k)
    print("frame", k, counts.tolist(), ok.tolist(), maps.stats(0).tolist())
print("occupied", len(maps.getOccupancyMap(0)))
start = np.zeros((2, 9)); start[:, :3] = pos; start[:, 3] = 1.0
goal = np.array([[4.0, 0, 1.0], [4.0, 1.0, 1.0]])
st, tr = eng.plan_cycle(start, goal, 10.32)
print("plan", st.tolist())
pva, valid = eng.traj_sample_batch([10.4, 10.6])
eng.traj_publish_local(); eng.swarm_advance(10.42); print("states", eng.swarm_states()[:, :3].round(3).tolist())
eng.close()
c = gridmap.default_gridmap_config()
c.map_size[0], c.map_size[1], c.map_size[2] = 20.0, 20.0, 4.0
c.local_map_margin = 5
gm = gridmap.GridMaps(c, n_maps=2)
P = gridmap.depth_params(72, 96, cam.fx, cam.fy, cam.cx, cam.cy, use_filter=False, skip_pixel=2)
gm.setDepthParams(P)
c2b = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
img = np.where(depth >= 65535, 0, depth).astype(np.uint16)
for k in range(2):
    gm.updateOccupancyFromDepth(img, np.stack([c2b, c2b]), pos)
print("gridmap occupied", int((gm.buffers(0)[0] > 0).sum()))
gm.close()
print("done")
