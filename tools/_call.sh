D=gpurun_out/r2n
mkdir -p $D
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -k "ffn_fused" 2>&1 | tail -15 > $D/pytest_ffn.log
cat $D/pytest_ffn.log
if grep -q "failed\|error" $D/pytest_ffn.log; then exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $D/pytest.log
cat $D/pytest.log
timeout 300 python bench.py --cpu-seconds 3 --parity-seconds 8 > $D/bench_cfg2_fused.json 2> $D/err1; head -c 200 $D/bench_cfg2_fused.json; echo
KDB200_NO_FFN_FUSE=1 timeout 300 python bench.py --no-extras > $D/bench_cfg2_unfused.json 2> $D/err2; head -c 200 $D/bench_cfg2_unfused.json; echo
timeout 200 python tools/profile_forward.py > $D/fwd_sw.txt 2>&1; grep "ffn\|total" $D/fwd_sw.txt
