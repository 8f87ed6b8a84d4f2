for r in 8 16; do for c in 6 7 8; do echo "R=$r ctas/sm=$c" >> gpurun_out/batch2.log; NW_CUDA_BATCH_R=$r NW_CUDA_BATCH_CTAS_PER_SM=$c python bench.py --workload batch --steps 3 --warmup 2 --batch-pairs 200000 >> gpurun_out/batch2.log 2>&1; done; done
echo "default 1M pairs" >> gpurun_out/batch2.log
python bench.py --workload batch --steps 3 --warmup 2 --batch-pairs 1000000 >> gpurun_out/batch2.log 2>&1
