# C-SVC decoder inside the headline workload: plain run, launch list, one
# `--set full` capture of the SMO kernel and the kernel-matrix kernel.
set -x
python profiles/profile_step.py --stages --decoder svc_rbf --class-weight balanced > gpurun_out/plain_svc.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_svc.csv python profiles/profile_step.py --decoder svc_rbf --class-weight balanced > gpurun_out/ncu_run_svc.log 2>&1
for spec in "k_svc_smo:1:1" "k_svc_kmat:1:1" "k_svc_predict:1:1"; do
  k=$(echo $spec | cut -d: -f1); s=$(echo $spec | cut -d: -f2); c=$(echo $spec | cut -d: -f3)
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o gpurun_out/prof_r1_$k python profiles/profile_step.py --decoder svc_rbf --class-weight balanced > gpurun_out/ncu_full_$k.log 2>&1
  tail -1 gpurun_out/ncu_full_$k.log
done
tail -4 gpurun_out/plain_svc.log
