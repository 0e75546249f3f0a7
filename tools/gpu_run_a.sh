set -x
nproc; grep -m1 "model name" /proc/cpuinfo; free -g | head -2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python tools/cold_trace.py > gpurun_out/cold_trace.log 2>&1
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_r02a.log 2>&1; tail -5 gpurun_out/pytest_gpu_r02a.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; tail -c 3000 gpurun_out/bench_r02a.json; tail -5 gpurun_out/bench_r02a.err
( time python bench.py --impl reference --steps 4 --warmup 1 ) > gpurun_out/bench_ref_r02a.json 2>&1; tail -c 1500 gpurun_out/bench_ref_r02a.json
