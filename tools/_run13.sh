cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_umma.py tests/test_gpu_nets.py -x -q -m gpu > gpurun_out/t_wh.log 2>&1
echo "rc=$?" >> gpurun_out/t_wh.log
tail -6 gpurun_out/t_wh.log
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/b2d_wh.log 2>&1
timeout 300 python tools/kernel_breakdown.py --out gpurun_out/bd_2d_wh.txt > /dev/null 2>&1
