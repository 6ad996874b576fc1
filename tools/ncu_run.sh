set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:layer_kernel -s 40 -c 2 -f -o gpurun_out/r01b_layer python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tail_kernel -s 1 -c 1 -f -o gpurun_out/r01b_tail python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:layer_kernel -s 40 -c 1 -f -o gpurun_out/r01b_layer_tf32 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --precision tf32 --batch 32 > gpurun_out/ncu_b4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tail_kernel -s 1 -c 1 -f -o gpurun_out/r01b_tail_tf32 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --precision tf32 --batch 32 > gpurun_out/ncu_b5.log 2>&1
ls -la gpurun_out/*.ncu-rep
