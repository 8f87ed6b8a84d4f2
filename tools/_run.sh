nvidia-smi -L | wc -l > gpurun_out/n4.log
timeout 300 python bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
timeout 300 python bench.py --gpus 4 --workload batch --steps 3 --warmup 2 --batch-pairs 200000 > gpurun_out/bench_batch_n4.json 2> gpurun_out/bench_batch_n4.err
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "multi_gpu" 2>&1 | tail -3 >> gpurun_out/n4.log
