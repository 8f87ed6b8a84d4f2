timeout 300 python -m pytest tests -x -q -m gpu -k "full_table or streamed or driver" 2>&1 | tail -3 > gpurun_out/pytest19.log
B=oracle/_ref/bdna; export NW_CUDA_TRACE=1
for nt in 0 1; do for th in 8 12 16; do echo "== mid full NO_NT=$nt threads=$th" >> gpurun_out/driver8.log; NW_CUDA_NO_NT=$nt NW_CUDA_COPY_THREADS=$th NW_CUDA_MODE=full fast-needleman-wunsch_b200/bin/cuda.e $B/mid1.bdna $B/mid2.bdna 2>&1 | grep -E "^[0-9]+$|table_to" >> gpurun_out/driver8.log; done; done
for nt in 0 1; do echo "== 2gb full NO_NT=$nt" >> gpurun_out/driver8.log; NW_CUDA_NO_NT=$nt NW_CUDA_MODE=full fast-needleman-wunsch_b200/bin/cuda.e $B/2gb-1.bdna $B/2gb-2.bdna 2>&1 | grep -E "^[0-9]+$|table_to" >> gpurun_out/driver8.log; done
