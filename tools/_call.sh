D=gpurun_out/r2d
mkdir -p $D
timeout 120 tools/bin/hw_probe > $D/hw_probe.txt 2>&1
timeout 500 python bench.py --config cfg3 > $D/bench_cfg3.json 2> $D/bench_cfg3.err
timeout 500 python bench.py --config cfg4 > $D/bench_cfg4.json 2> $D/bench_cfg4.err
timeout 700 python bench.py --config cfg5 > $D/bench_cfg5.json 2> $D/bench_cfg5.err
timeout 120 python tools/attn_trace.py > $D/attn_trace.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_pipe --launch-skip 36 -c 12 -f -o $D/attn_sw python tools/profile_forward.py > $D/ncu_attn_sw.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_pipe --launch-skip 36 -c 12 -f -o $D/attn_c5 python tools/profile_forward.py --config c5 --batch 16 > $D/ncu_attn_c5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_pipe --launch-skip 36 -c 4 -f -o $D/attn_na python tools/profile_forward.py --config na > $D/ncu_attn_na.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_persist --launch-skip 156 -c 12 -f -o $D/gemm_sw python tools/profile_forward.py > $D/ncu_gemm_sw.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 204 -c 80 --csv --log-file $D/launches_sw.csv python tools/profile_forward.py > $D/ncu_ll_sw.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 204 -c 80 --csv --log-file $D/launches_c5.csv python tools/profile_forward.py --config c5 --batch 16 > $D/ncu_ll_c5.log 2>&1
cat $D/hw_probe.txt; head -c 300 $D/bench_cfg5.json
