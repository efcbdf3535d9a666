"""Host-side timing of every call of the pipelined e2e loop (tuning only)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from slamrs_b200 import GpuPlacement, GridMapSlam
from slamrs_b200.workloads import WORKLOADS

wl = WORKLOADS["c3"]
sim = wl.simulator()
WARM = int(sys.argv[2]) if len(sys.argv) > 2 else 5
scans = [sim.next_scan(wl.speed_left, wl.speed_right) for _ in range(WARM + 25)]
slam = GridMapSlam(wl.slam_config(), GpuPlacement(device=0))
for obs, odo in scans[:WARM]:
    slam.update(obs, odo)
bufs = [torch.empty(slam.grid_w * slam.grid_h, dtype=torch.float64).pin_memory().numpy() for _ in range(2)]
mode = sys.argv[1] if len(sys.argv) > 1 else "async"
rows = []
extra = []
t_all = time.perf_counter()
for i, (obs, odo) in enumerate(scans[WARM:WARM + 20]):
    t0 = time.perf_counter()
    if mode == "split":
        a, d, v = slam._scan_arrays(obs)
        from slamrs_b200.slam import _ptr
        slam._L.slamrs_gpu_upload_scan(slam._h, _ptr(a), _ptr(d), _ptr(v), a.size); ta = time.perf_counter()
        slam.step_async(odo); tb = time.perf_counter()
        slam.sync(); extra.append(((ta - t0) * 1e3, (tb - ta) * 1e3, (time.perf_counter() - tb) * 1e3))
    else:
        slam.update(obs, odo)
    t1 = time.perf_counter(); slam.estimated_pose()
    t2 = time.perf_counter()
    if mode in ("async", "split"):
        slam.map_wait(); t3 = time.perf_counter(); slam.estimated_likelihood_async(bufs[i & 1])
    else:
        t3 = t2; slam.estimated_likelihood(bufs[i & 1])
    t4 = time.perf_counter()
    rows.append([(t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3])
slam.map_wait()
print(mode, "total ms/step", (time.perf_counter() - t_all) * 1e3 / 20)
for r in rows: print("  update %.3f pose %.3f wait %.3f readout %.3f" % tuple(r))
for e in extra: print('  upload %.3f step_async %.3f sync %.3f' % e)
slam.close()
