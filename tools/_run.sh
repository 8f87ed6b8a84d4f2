timeout 400 python tools/stress.py 1 150 > gpurun_out/stress.log 2>&1
timeout 400 python tools/stress.py 2 150 >> gpurun_out/stress.log 2>&1
