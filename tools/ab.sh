# A/B of kernel variants on ONE box, interleaved, two rounds:  bash tools/ab.sh lib1.so lib2.so@16 ...
# (library paths relative to audiopure_b200/; an optional @N sets AP_DEBUG=N, honoured by --ablation builds only)
mkdir -p gpurun_out
for i in 1 2; do
  for v in "$@"; do
    lib=${v%@*}; dbg=0; case "$v" in *@*) dbg=${v#*@};; esac
    AP_LIB=$PWD/audiopure_b200/$lib AP_DEBUG=$dbg python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-certify \
      --no-same-box-peak > gpurun_out/ab_${lib%.so}_${dbg}_$i.json 2>/dev/null
  done
done
python - "$@" <<PY
import json, sys
print("| variant | round | clips/s | layer ms/launch | tail ms/launch | SM MHz | W max |")
print("|---|---|---|---|---|---|---|")
for i in (1, 2):
    for v in sys.argv[1:]:
        lib, _, dbg = v.partition("@")
        dbg = dbg or "0"
        try:
            d = json.loads(open("gpurun_out/ab_%s_%s_%d.json" % (lib[:-3], dbg, i)).read().strip().splitlines()[-1])
            print("| %s | %d | %.1f | %.4f | %.3f | %s | %s |" % (v, i, d["value"], d["roofline"]["avg_launch_ms"],
                  d["roofline"]["tail_kernel_ms_per_launch"], d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"]))
        except Exception as e:
            print("|", v, "|", i, "| failed", e, "|")
PY
