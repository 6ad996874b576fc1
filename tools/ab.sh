# A/B of kernel variants on one box: bash tools/ab.sh lib1.so lib2.so ...  (paths relative to audiopure_b200/)
for i in 1 2; do
  for lib in "$@"; do
    AP_LIB=$PWD/audiopure_b200/$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_${lib%.so}_$i.json 2>/dev/null
  done
done
python - "$@" <<PY
import json, sys
for i in (1, 2):
    for lib in sys.argv[1:]:
        n = lib[:-3]
        try:
            d = json.loads(open("gpurun_out/ab_%s_%d.json" % (n, i)).read().strip().splitlines()[-1])
            print(n, i, round(d["value"], 1), round(d["roofline"]["avg_launch_ms"], 4), round(d["roofline"]["tail_kernel_ms_per_launch"], 3), d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"])
        except Exception as e:
            print(n, i, "failed", e)
PY
