timeout 300 python -m pytest tests -x -q -m gpu -k "score" 2>&1 | tail -3 > gpurun_out/pytest20.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final3.json 2>&1
