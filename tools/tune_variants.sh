#!/bin/bash
# A/B of tuning variants of the product library (built here with `python slamrs_b200/build.py -DNAME=V
# --out=variants/<tag>.so`). Usage (under gpurun): bash tools/tune_variants.sh <out tag> <variant tags...>
# Every variant replaces the in-tree library for one short bench run; the in-tree library is restored.
set -u
OUT=gpurun_out/tune_$1.log; shift
LIB=slamrs_b200/libslamrs_gpu.so
cp $LIB variants/main.so
: > $OUT
for v in main "$@"; do
  cp variants/$v.so $LIB
  echo "== $v" >> $OUT
  python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-e2e --no-full-copy --no-strict --no-eager ${BENCH_EXTRA:-} 2>> $OUT.err | python -c '
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]; s = d.get("strict_order_of_work") or {}
        print(json.dumps({"ms": round(d["ms_per_step"], 4), "copy_ms": round(r["ms_per_launch"], 4), "MB": round(r["bytes_per_launch"] / 1e6, 1),
                          "frac": round(r["frac"], 3), "phases": {k: round(v, 4) for k, v in d["phases_ms_per_step"].items()},
                          "strict_ms": round(s.get("ms_per_step", 0), 4), "strict_ray": round(s.get("phases_ms_per_step", {}).get("ray_update", 0), 4)}))
' >> $OUT
done
cp variants/main.so $LIB
cat $OUT
