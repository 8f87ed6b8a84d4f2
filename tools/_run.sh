timeout 600 python -m pytest tests -x -q -m gpu -k "traceback or abi" 2>&1 | tail -8 > gpurun_out/pytest15.log
