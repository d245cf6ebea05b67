D=gpurun_out/r2z
mkdir -p $D
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --config cfg4 --no-extras > $D/bench_cfg4_4gpu.json 2> $D/err4
head -c 300 $D/bench_cfg4_4gpu.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 4 --config cfg5 --no-extras > $D/bench_cfg5_4gpu.json 2> $D/err5
head -c 300 $D/bench_cfg5_4gpu.json; echo
tail -3 $D/err4 $D/err5
