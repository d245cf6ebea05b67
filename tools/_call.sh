D=gpurun_out/r2w
mkdir -p $D
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $D/pytest.log; tail -2 $D/pytest.log
timeout 300 python __graft_entry__.py smoke > $D/smoke.log 2>&1; tail -1 $D/smoke.log
timeout 400 python bench.py > $D/bench_cfg2.json 2> $D/err2; head -c 130 $D/bench_cfg2.json | cut -c50-130; echo
timeout 400 python bench.py --config cfg3 --cpu-seconds 8 --parity-seconds 15 > $D/bench_cfg3.json 2> $D/err3; head -c 160 $D/bench_cfg3.json | cut -c80-160; echo
timeout 400 python bench.py --config cfg4 --cpu-seconds 8 --parity-seconds 15 > $D/bench_cfg4.json 2> $D/err4; head -c 190 $D/bench_cfg4.json | cut -c100-190; echo
timeout 600 python bench.py --config cfg5 --cpu-seconds 8 --parity-seconds 20 > $D/bench_cfg5.json 2> $D/err5; head -c 170 $D/bench_cfg5.json | cut -c80-170; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 180 -c 70 --csv --log-file $D/launches_sw.csv python tools/profile_forward.py > $D/fwd_sw_under_ncu.txt 2>&1
timeout 200 python tools/profile_forward.py > $D/fwd_sw.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ffn_fused --launch-skip 6 -c 1 -f -o /tmp/ffn python tools/profile_forward.py > $D/ncu_ffn.log 2>&1
ncu -i /tmp/ffn.ncu-rep --page raw --csv > $D/ncu_ffn_raw.csv 2>/dev/null
ls -la $D | head -20
