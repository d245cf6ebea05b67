D=gpurun_out/r3b
mkdir -p $D
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $D/pytest.log; tail -2 $D/pytest.log
timeout 100 python tools/fa2_compare.py > $D/fa2_compare.txt 2>&1; cut -c1-330 $D/fa2_compare.txt
