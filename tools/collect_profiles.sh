#!/bin/bash
# Round-2 evidence: `ncu --set full` captures of the dominant kernels, summarised on the box (the .ncu-rep files are too
# large to bring back together), ncu launch lists of the 2-D and 3-D steps, compute-sanitizer logs.
#   gpurun --timeout 1800 -- bash tools/collect_profiles.sh
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
sum() { python tools/ncu_summary.py $O/$1.ncu-rep > $O/$1_ncu.txt 2>&1; rm -f $O/$1.ncu-rep; }
$NCU -k regex:conv_umma_pair -c 4 -o $O/r2_conv python tools/profile_conv.py 16 2 > $O/r2_conv_run.log 2>&1; sum r2_conv
$NCU -k regex:conv_wgrad_pair -c 2 -o $O/r2_wgrad python tools/profile_conv.py 16 2 > /dev/null 2>&1; sum r2_wgrad
$NCU -k regex:dmarch -c 2 -o $O/r2_conv3d python tools/profile_conv3d.py 36 16 > $O/r2_conv3d_run.log 2>&1; sum r2_conv3d
$NCU -k regex:dmarch -c 1 -o $O/r2_conv3d_32 python tools/profile_conv3d.py 32 32 160 1 > /dev/null 2>&1; sum r2_conv3d_32
$NCU -k regex:conv_wgrad_halo -c 1 -o $O/r2_wgrad3d python tools/profile_conv3d.py 36 16 > /dev/null 2>&1; sum r2_wgrad3d
ncu --set full --clock-control none -k "regex:ncc_|warp_|vecint|resize|grad_|fused_reg" -c 17 -o $O/r2_regops python tools/prof_ops3d.py > /dev/null 2>&1; sum r2_regops
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches.csv python tools/one_step.py > $O/r2_onestep.log 2>&1
python tools/launch_summary.py $O/r2_launches.csv > $O/r2_launches_summary.txt 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_umma.py -q -x -k "conv3d and (v9 or v12)" > $O/r2_racecheck_dmarch.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_umma.py tests/test_gpu_ops.py tests/test_gpu_optim.py -q -x -k "(conv3d and (v9 or v10 or v11 or v13)) or ncc or adam" > $O/r2_memcheck.log 2>&1
tail -n 3 $O/r2_racecheck_dmarch.log; tail -n 3 $O/r2_memcheck.log; du -sh $O
