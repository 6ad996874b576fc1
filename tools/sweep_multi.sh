# usage: bash tools/sweep_multi.sh N   (under gpurun --gpus N)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29530 tools/sweep.py --batches 64,256,1024 --tstars 1,2,3,5,10 --json gpurun_out/r02_sweep_${N}gpu.jsonl 2>&1 | grep -v -i "warn\|OMP_NUM\|\*\*\*\*" | tee gpurun_out/r02_sweep_ddpm_${N}gpu.md | tail -4
$TR --master-port 29531 tools/sweep.py --purifier sde --batches 256 --tstars 5,10,25 --json gpurun_out/r02_sweep_${N}gpu.jsonl 2>&1 | grep -v -i "warn\|OMP_NUM\|\*\*\*\*" | tee gpurun_out/r02_sweep_sde_${N}gpu.md | tail -4
$TR --master-port 29532 tools/multigpu_check.py 2>&1 | grep "MULTIGPU\|equal" | tail -3
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/r02_nccl_${N}gpu.%h.%p.log python bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
tail -c 1200 gpurun_out/r02_bench_${N}gpu.json
cat gpurun_out/r02_nccl_${N}gpu.*.log 2>/dev/null | grep -i "nvls\|nranks\|NCCL version\|Connected all\|via P2P\|comm 0x" | sort | uniq -c | sort -rn | head -12 > gpurun_out/r02_nccl_${N}gpu_summary.txt
rm -f gpurun_out/r02_nccl_${N}gpu.*.log
