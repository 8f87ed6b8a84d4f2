python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 >> gpurun_out/smoke.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_default_ref.json 2>&1
