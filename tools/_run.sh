python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "boundar or checkpoint or edge or random or column or score" 2>&1 | tail -3 > gpurun_out/pytest4.log
timeout 300 python tools/quick.py 2gb,mid,big,64gb 0,8 4 > gpurun_out/quick4.log 2>&1
timeout 100 python tools/micro1.py >> gpurun_out/quick4.log 2>&1
