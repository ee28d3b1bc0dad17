set -x
python profiles/profile_step.py --stages > gpurun_out/plain_r1f.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1f.csv python profiles/profile_step.py > gpurun_out/ncu_run_r1f.log 2>&1
for spec in "k_proj_tc:1" "k_gram_tc:1" "k_eig_tile:5" "k_sgemm:25" "k_svm_fit:1" "k_chol_inv:12"; do
  k=${spec%%:*}; s=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o gpurun_out/prof_r1_$k python profiles/profile_step.py > gpurun_out/ncu_full_$k.log 2>&1
  tail -2 gpurun_out/ncu_full_$k.log
done
tail -4 gpurun_out/plain_r1f.log
