python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/pytest10.log
timeout 300 python tools/quick.py 2gb,mid,64gb 0 4 > gpurun_out/quick5.log 2>&1
timeout 100 python tools/micro1.py >> gpurun_out/quick5.log 2>&1
python bench.py --workload batch --steps 3 --warmup 2 --batch-pairs 200000 >> gpurun_out/quick5.log 2>&1
python tools/full.py 8 2>&1 | head -1 >> gpurun_out/quick5.log
