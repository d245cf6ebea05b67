mkdir -p gpurun_out/r2a
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/r2a/pytest.log
echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
timeout 400 python bench.py > gpurun_out/r2a/bench_cfg2.json 2> gpurun_out/r2a/bench_cfg2.err
KDB200_GEMM_ROLES_LO=1 timeout 200 python bench.py --no-extras > gpurun_out/r2a/bench_cfg2_roles_lo.json 2> gpurun_out/r2a/bench_cfg2_roles_lo.err
KDB200_ATTN_PERSIST=1 timeout 200 python bench.py --no-extras > gpurun_out/r2a/bench_cfg2_attn_persist.json 2> gpurun_out/r2a/bench_cfg2_attn_persist.err
timeout 200 python tools/profile_forward.py > gpurun_out/r2a/fwd_sw.txt 2>&1
timeout 200 python tools/profile_forward.py --config na > gpurun_out/r2a/fwd_na.txt 2>&1
timeout 400 python bench.py --config cfg3 > gpurun_out/r2a/bench_cfg3.json 2> gpurun_out/r2a/bench_cfg3.err
tail -5 gpurun_out/r2a/pytest.log; cat gpurun_out/r2a/bench_cfg2_roles_lo.json | head -c 300; echo; cat gpurun_out/r2a/bench_cfg2_attn_persist.json | head -c 300
