D=gpurun_out/r2y
mkdir -p $D
timeout 600 python -m pytest tests/test_gpu_next_samplers.py -q -k "dpm" 2>&1 | tail -25 > $D/pytest_dpm.log
cat $D/pytest_dpm.log
