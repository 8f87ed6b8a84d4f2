timeout 600 python -m pytest tests -x -q -m gpu -k "batch or driver or error or device_res" 2>&1 | tail -3 > gpurun_out/pytest14.log
for c in 4 5; do NW_CUDA_BATCH_CTAS_PER_SM=$c python bench.py --workload batch --steps 3 --warmup 2 --batch-pairs 200000 >> gpurun_out/batch4.log 2>&1; done
python bench.py --workload batch --steps 3 --warmup 2 --batch-pairs 1000000 >> gpurun_out/batch4.log 2>&1
B=oracle/_ref/bdna; export NW_CUDA_TRACE=1
for i in 1 2 3; do for m in boundary full; do echo "== cuda.e 2gb $m" >> gpurun_out/driver5.log; NW_CUDA_MODE=$m fast-needleman-wunsch_b200/bin/cuda.e $B/2gb-1.bdna $B/2gb-2.bdna >> gpurun_out/driver5.log 2>&1; done; done
for i in 1 2; do echo "== cuda.e 64gb boundary" >> gpurun_out/driver5.log; NW_CUDA_MODE=boundary fast-needleman-wunsch_b200/bin/cuda.e $B/64gb-1.bdna $B/64gb-2.bdna >> gpurun_out/driver5.log 2>&1; done
