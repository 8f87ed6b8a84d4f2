python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/pytest9.log
python bench.py --workload 2gb-full --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gbfull.json 2> gpurun_out/bench_2gbfull.err
python bench.py --workload 2gb --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gb.json 2> gpurun_out/bench_2gb.err
B=oracle/_ref/bdna; export NW_CUDA_TRACE=1
for i in 1 2 3; do NW_CUDA_MODE=full fast-needleman-wunsch_b200/bin/cuda.e $B/2gb-1.bdna $B/2gb-2.bdna >> gpurun_out/driver4.log 2>&1; done
