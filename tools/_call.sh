D=gpurun_out/r2m
mkdir -p $D
L=k-diffusion_b200/k_diffusion/_lib/libkdb200.so
cp $L /tmp/new.so
timeout 300 python bench.py --cpu-seconds 3 --parity-seconds 5 > $D/bench_new_iss2.json 2> $D/err1; head -c 150 $D/bench_new_iss2.json; echo
KDB200_GEMM_ISSUERS=1 timeout 300 python bench.py --no-extras > $D/bench_new_iss1.json 2> $D/err2; head -c 150 $D/bench_new_iss1.json; echo
cp tools/bin/old/libkdb200.so $L
timeout 300 python bench.py --cpu-seconds 3 --parity-seconds 5 > $D/bench_old.json 2> $D/err3; head -c 150 $D/bench_old.json; echo
cp /tmp/new.so $L
timeout 300 python bench.py --no-extras > $D/bench_new_iss2_again.json 2> $D/err4; head -c 150 $D/bench_new_iss2_again.json; echo
