for v in build/libnw_x7.so build/libnw_x8.so build/libnw_x9.so; do
  export NW_CUDA_LIB=$PWD/$v
  timeout 200 python tools/micro2.py 2>&1 | grep 64gb
done > gpurun_out/xbisect3.log 2>&1
