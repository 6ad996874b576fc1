# r02 profile captures (run under gpurun, ONE GPU).  Numbers printed by a run under ncu are never bench values.
set -x
B="python bench.py --no-cpu-baseline --no-certify --no-same-box-peak"
# every launch of the run (weight packing and cuDNN autotuning included; tools/launch_list.py cuts out one steady-state step)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv $B --steps 2 --warmup 3 > gpurun_out/ncu_r02_1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:layer_kernel -s 40 -c 2 -f -o gpurun_out/r02_layer $B --steps 1 --warmup 1 > gpurun_out/ncu_r02_2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tail_kernel -s 1 -c 1 -f -o gpurun_out/r02_tail $B --steps 1 --warmup 1 > gpurun_out/ncu_r02_3.log 2>&1
ncu --set full --clock-control none -k regex:"logmel_kernel|prologue_kernel|axpbz_kernel|smooth_inputs_kernel|vote_counts_kernel" -c 12 -f -o gpurun_out/r02_small python tools/small_kernels.py > gpurun_out/ncu_r02_4.log 2>&1
ls -la gpurun_out/*.ncu-rep
