set -x
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -k "group or cli_gpus" > gpurun_out/pytest_group_1gpu.log 2>&1; tail -15 gpurun_out/pytest_group_1gpu.log
