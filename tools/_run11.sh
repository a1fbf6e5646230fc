cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
DFMIR_WGRAD_PAIR=1 timeout 180 python -m pytest tests/test_gpu_umma.py -x -q -m gpu -k "u2 or u6" > gpurun_out/t_wpair.log 2>&1
echo "rc=$?" >> gpurun_out/t_wpair.log
tail -12 gpurun_out/t_wpair.log
DFMIR_WGRAD_PAIR=1 timeout 120 python tools/profile_conv.py > gpurun_out/pconv_wpair.log 2>&1; echo "rc=$?" >> gpurun_out/pconv_wpair.log
timeout 120 python tools/profile_conv.py > gpurun_out/pconv_nowpair.log 2>&1
DFMIR_WGRAD_PAIR=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/b2d_wpair.log 2>&1
