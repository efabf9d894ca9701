set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02d_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02d_bench_reference.json 2>> gpurun_out/r02d_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02d_ncu_launch.log 2>&1
python tools/profile_vote.py 10000 50000 8 2 > gpurun_out/r02d_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vote_kernel_grouped -s 2 -c 1 -o gpurun_out/r02d_vote -f python tools/profile_vote.py 10000 50000 8 2 > gpurun_out/r02d_ncu.log 2>&1
tail -3 gpurun_out/r02d_tests.log; cut -c1-400 gpurun_out/r02d_bench.json
