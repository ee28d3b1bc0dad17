# Round-1 profiling pass (one B200, under gpurun): plain run, per-launch device times of one
# warm-up + one measured 107-fold batch, then one `--set full` capture per main kernel.
set -x
python profiles/profile_step.py --stages > gpurun_out/plain_r1.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1.csv python profiles/profile_step.py > gpurun_out/ncu_run_r1.log 2>&1
for spec in "k_proj_tc:1:1" "k_gram_tc:1:1" "k_eig_tile:4:2" "k_sgemm:30:3" "k_svm_fit:1:1" "k_chol_inv:12:1" "k_gram_tn:3:2"; do
  k=$(echo $spec | cut -d: -f1); s=$(echo $spec | cut -d: -f2); c=$(echo $spec | cut -d: -f3)
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o gpurun_out/prof_r1_$k python profiles/profile_step.py > gpurun_out/ncu_full_$k.log 2>&1
  tail -1 gpurun_out/ncu_full_$k.log
done
tail -4 gpurun_out/plain_r1.log
