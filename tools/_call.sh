D=gpurun_out/r2h
mkdir -p $D
timeout 120 tools/bin/mma_issue_bench > $D/mma_issue_bench.txt 2>&1
cat $D/mma_issue_bench.txt
