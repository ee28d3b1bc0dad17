set -x
python profiles/profile_step.py --stages > gpurun_out/plain_r1.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1.csv python profiles/profile_step.py > gpurun_out/ncu_run_r1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_proj_tc -s 1 -c 1 -f -o gpurun_out/prof_r1_k_proj_tc python profiles/profile_step.py > gpurun_out/ncu_full_k_proj_tc.log 2>&1
tail -1 gpurun_out/ncu_full_k_proj_tc.log
tail -4 gpurun_out/plain_r1.log
