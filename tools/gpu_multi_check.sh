# usage: gpu_multi_check.sh N   -- bench at N GPUs (torchrun), JSON + stderr into gpurun_out/
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
echo "rc=$?"; tail -c 1500 gpurun_out/r02_bench_${N}gpu.json; tail -5 gpurun_out/r02_bench_${N}gpu.err
