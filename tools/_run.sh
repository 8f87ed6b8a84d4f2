timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/pytest12.log
for k in 0 1; do echo "K2=$k" >> gpurun_out/quick8.log; NW_CUDA_K2=$k timeout 300 python tools/quick.py 2gb,mid,big,64gb 0,4,8,16 4 >> gpurun_out/quick8.log 2>&1; NW_CUDA_K2=$k timeout 100 python tools/micro1.py >> gpurun_out/quick8.log 2>&1; done
