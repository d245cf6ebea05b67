D=gpurun_out/r2v
mkdir -p $D
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -k "ffn_fused" 2>&1 | tail -4 > $D/pytest_ffn.log
cat $D/pytest_ffn.log
if grep -q "failed\|error" $D/pytest_ffn.log; then exit 1; fi
KDB200_FFN_TRACE=1 timeout 100 python tools/ffn_probe.py > $D/ffn_trace.txt 2>&1
grep -A31 "FFN trace M=131072" $D/ffn_trace.txt | head -32 | tail -31 | cut -c1-90
timeout 100 python tools/ffn_probe.py 2>&1 | grep "dbg="
timeout 300 python bench.py --no-extras > $D/bench_cfg2.json 2> $D/err1; head -c 130 $D/bench_cfg2.json | cut -c50-130; echo
timeout 600 python -m pytest tests/test_gpu_bf16_parity.py tests/test_gpu_parity.py -x -q 2>&1 | tail -2
