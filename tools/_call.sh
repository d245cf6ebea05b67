#!/bin/bash
# Scratch wrapper for one GPU call:   gpurun --timeout 900 -- 'bash tools/_call.sh r9z'
# Writes everything under gpurun_out/<tag>/ (the only directory that comes back; keep it far below 64 MiB:
# export ncu reports as --page raw/source --csv and leave the .ncu-rep on the box).
D=gpurun_out/${1:-scratch}
mkdir -p $D
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $D/pytest.log; tail -2 $D/pytest.log
timeout 300 python bench.py > $D/bench.json 2> $D/bench.err; cut -c1-400 $D/bench.json
