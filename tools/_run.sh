python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_check.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 >> gpurun_out/final_check.log
python bench.py > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err
