nvidia-smi -L | wc -l > gpurun_out/n8.log
timeout 300 python bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
timeout 300 python bench.py --gpus 8 --workload batch --steps 3 --warmup 2 --batch-pairs 200000 > gpurun_out/bench_batch_n8.json 2> gpurun_out/bench_batch_n8.err
timeout 200 python bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench_ref_n8.json 2>&1
