cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
DFMIR_UMMA_PAIR=2 timeout 180 python -m pytest tests/test_gpu_umma.py -x -q -m gpu -k "u0 or u3 or u1 or u2" > gpurun_out/t_pair2.log 2>&1
echo "rc=$?" >> gpurun_out/t_pair2.log
tail -12 gpurun_out/t_pair2.log
DFMIR_UMMA_PAIR=2 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/b2d_pair2.log 2>&1
DFMIR_UMMA_PAIR=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/b2d_pair1.log 2>&1
DFMIR_UMMA_PAIR=2 timeout 300 python tools/kernel_breakdown.py --out gpurun_out/bd_2d_pair2.txt > /dev/null 2>&1
