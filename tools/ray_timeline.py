"""Reads the per-item timeline a -DSLAMRS_RAY_TRACE build dumps (SLAMRS_RAY_TRACE_LOG) for the last
k_ray_update_half launch: per phase mean / max durations, rounds per CTA and the launch's makespan."""
import sys
import numpy as np

log = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 16)
t0c = log[:, 0].astype(np.int64)
live = (t0c > 0) & (log[:, 7] > 0)
# only the items of the LAST launch: stamps within 1 ms of the newest one
newest = log[live][:, 7].astype(np.int64).max()
live &= (newest - log[:, 7].astype(np.int64)) < 1_000_000
ends = np.sort(log[live][:, 7].astype(np.int64))
cut = ends[0]
for a, b in zip(ends[:-1], ends[1:]):      # the last launch = everything after the last gap of > 60 us between item ends
    if b - a > 60_000:
        cut = b
live &= log[:, 7].astype(np.int64) >= cut
idx = np.nonzero(live)[0]
L = log[idx].astype(np.int64)
start = L[:, 0].min()
names = ["setup", "beam select", "walk", "band scan", "exchange", "owner wait", "write-back", "spill/commit"]
print(f"items {len(idx)}  (work ids {idx.min()}..{idx.max()}), makespan {(L[:, 8].max() - start) / 1e3:.1f} us")
info = L[:, 9]
fused = (info >> 40) & 1
upper = (info >> 41) & 1
cta = (info >> 16) & 0xffff
smid = info & 0xffff
nbeam = (info >> 44) & 0xfff
remote = (info >> 42) & 1
for k, n in enumerate(names):
    d = (L[:, k + 1] - L[:, k]) / 1e3
    print(f"  {n:14s} mean {d.mean():7.2f} us  p50 {np.median(d):7.2f}  max {d.max():7.2f}")
tot = (L[:, 8] - L[:, 0]) / 1e3
print(f"  {'item total':14s} mean {tot.mean():7.2f} us  p50 {np.median(tot):7.2f}  max {tot.max():7.2f}")
print(f"  fused items {int(fused.sum())}, owners {int((1 - fused).sum())}; beams per half item mean {nbeam.mean():.0f}")
if remote.any():
    r = remote == 1
    print(f"  pulled by a peer: {int(r.sum())} items, total mean {tot[r].mean():.2f} us (others {tot[~r].mean():.2f}); they end at "
          f"{((L[r, 8] - start) / 1e3).min():.1f}..{((L[r, 8] - start) / 1e3).max():.1f} us")
# rounds
order = np.argsort(L[:, 0])
per_cta = {}
for i in order:
    per_cta.setdefault(int(cta[i]), []).append(i)
rounds = np.array([len(v) for v in per_cta.values()])
print(f"  CTAs that worked {len(per_cta)}; items per CTA: " + ", ".join(f"{r}:{int((rounds == r).sum())}" for r in sorted(set(rounds))))
for r in range(rounds.max()):
    its = [v[r] for v in per_cta.values() if len(v) > r]
    s = (L[its, 0] - start) / 1e3
    e = (L[its, 8] - start) / 1e3
    print(f"  round {r}: {len(its)} items, start {s.min():.1f}..{s.max():.1f} us, end {e.min():.1f}..{e.max():.1f} us, dur mean {(e - s).mean():.1f}")
    for k, n in enumerate(names):
        d = (L[its, k + 1] - L[its, k]) / 1e3
        print(f"      {n:14s} mean {d.mean():7.2f}  max {d.max():7.2f}")
