python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err
python bench.py --workload mid --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mid.json 2>&1
python bench.py --workload big --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_big.json 2>&1
B=oracle/_ref/bdna; export NW_CUDA_TRACE=1
for i in 1 2 3; do echo "== cuda.e 64gb boundary" >> gpurun_out/driver7.log; NW_CUDA_MODE=boundary fast-needleman-wunsch_b200/bin/cuda.e $B/64gb-1.bdna $B/64gb-2.bdna >> gpurun_out/driver7.log 2>&1; done
