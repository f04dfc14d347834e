mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/r6_multi_tests.log 2>&1
tail -3 gpurun_out/r6_multi_tests.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 1 --warmup 1 --no-secondary ) > gpurun_out/r6_bench_n2.log 2>&1
tail -c 1500 gpurun_out/r6_bench_n2.log
