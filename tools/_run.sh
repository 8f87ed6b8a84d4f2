timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest17.log
B=oracle/_ref/bdna; export NW_CUDA_TRACE=1
for i in 1 2 3; do echo "== cuda.e 2gb full" >> gpurun_out/driver6.log; NW_CUDA_MODE=full fast-needleman-wunsch_b200/bin/cuda.e $B/2gb-1.bdna $B/2gb-2.bdna 2>&1 | grep -E "^[0-9]+$|Score|plan_create|sync|table_to" >> gpurun_out/driver6.log; done
for i in 1 2; do echo "== cuda.e mid full" >> gpurun_out/driver6.log; NW_CUDA_MODE=full fast-needleman-wunsch_b200/bin/cuda.e $B/mid1.bdna $B/mid2.bdna 2>&1 | grep -E "^[0-9]+$|Score|plan_create|sync|table_to" >> gpurun_out/driver6.log; done
echo "== cuda.e mid full NO_STREAMED" >> gpurun_out/driver6.log; NW_CUDA_NO_STREAMED=1 NW_CUDA_MODE=full fast-needleman-wunsch_b200/bin/cuda.e $B/mid1.bdna $B/mid2.bdna 2>&1 | grep -E "^[0-9]+$|Score|plan_create|sync|table_to" >> gpurun_out/driver6.log
echo "== sentinel mid 16 thr" >> gpurun_out/driver6.log; OMP_NUM_THREADS=16 oracle/_ref/sentinel-otf-blocked-mt.e $B/mid1.bdna $B/mid2.bdna >> gpurun_out/driver6.log 2>&1
