# usage: bash tools/verify_multi.sh N  (under gpurun --gpus N): sharded-vs-unsharded bit-equality checks, then bench.py at N GPUs.
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 tools/multigpu_check.py > gpurun_out/r02_multigpu_check_${N}gpu.log 2>&1; echo "multigpu_check rc=$?"; grep -c "bit-equal: True" gpurun_out/r02_multigpu_check_${N}gpu.log; grep "MULTIGPU" gpurun_out/r02_multigpu_check_${N}gpu.log
python bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench rc=$?"
tail -c 900 gpurun_out/r02_bench_${N}gpu.json
