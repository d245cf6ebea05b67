mkdir -p gpurun_out/r2b
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -s 2>&1 | tail -25 > gpurun_out/r2b/pytest_tc.log
echo "pytest_tc rc=$?" >> gpurun_out/r2b/pytest_tc.log
if grep -q "failed\|error\|Error" gpurun_out/r2b/pytest_tc.log; then tail -30 gpurun_out/r2b/pytest_tc.log; exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2b/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2b/smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/r2b/bench_cfg2.json 2> gpurun_out/r2b/bench_cfg2.err
KDB200_ATTN_ONESHOT=1 timeout 200 python bench.py --no-extras > gpurun_out/r2b/bench_cfg2_oneshot.json 2> gpurun_out/r2b/bench_cfg2_oneshot.err
timeout 400 python bench.py --config cfg3 > gpurun_out/r2b/bench_cfg3.json 2> gpurun_out/r2b/bench_cfg3.err
timeout 200 python tools/profile_forward.py > gpurun_out/r2b/fwd_sw.txt 2>&1
timeout 200 python tools/profile_forward.py --config na > gpurun_out/r2b/fwd_na.txt 2>&1
timeout 200 python tools/profile_forward.py --config c5 --batch 16 > gpurun_out/r2b/fwd_c5.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_pipe --launch-skip 36 -c 12 -f -o gpurun_out/r2b/attn_sw python tools/profile_forward.py > gpurun_out/r2b/ncu_attn_sw.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_pipe --launch-skip 36 -c 12 -f -o gpurun_out/r2b/attn_c5 python tools/profile_forward.py --config c5 --batch 16 > gpurun_out/r2b/ncu_attn_c5.log 2>&1
tail -4 gpurun_out/r2b/pytest.log; tail -2 gpurun_out/r2b/smoke.log
