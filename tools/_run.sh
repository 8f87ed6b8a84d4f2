python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full or edge or random or column or tiles" 2>&1 | tail -8 > gpurun_out/pytest8.log
for tb in 2 32; do for w in 8 4; do echo "warps2=$w tile_blocks=$tb" >> gpurun_out/full6.log; NW_CUDA_FULL_WARPS=$w NW_CUDA_TILE_BLOCKS=$tb python tools/full.py 8 2>&1 >> gpurun_out/full6.log; done; done
