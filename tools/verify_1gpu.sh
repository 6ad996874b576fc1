# Round-end style verification on ONE GPU (under gpurun): full GPU test suite, smoke, the default bench line and the TF32 bench line.
python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/r02_pytest_final.log 2>&1; tail -3 gpurun_out/r02_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -c 200 gpurun_out/r02_bench_1gpu.err; tail -c 400 gpurun_out/r02_bench_1gpu.json
python bench.py --precision tf32 --batch 32 --steps 6 --warmup 3 --no-cpu-baseline --no-certify --no-same-box-peak 2>/dev/null | tail -1 > gpurun_out/r02_bench_tf32.json; cut -c1-200 gpurun_out/r02_bench_tf32.json
