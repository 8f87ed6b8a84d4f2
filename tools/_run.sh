export NW_CUDA_K2=0 NW_CUDA_LIB=$PWD/build/libnw_t.so
timeout 60 python tools/micro3.py 2>&1 | sort | uniq -c > gpurun_out/abl7.log
