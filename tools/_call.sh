D=gpurun_out/r3a
mkdir -p $D
timeout 200 python tools/fa2_compare.py > $D/fa2_compare.txt 2>&1
cat $D/fa2_compare.txt | cut -c1-400
