set -x
python profiles/profile_step.py --stages --folds 138 > gpurun_out/plain_r2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r2.csv python profiles/profile_step.py --folds 138 > gpurun_out/ncu_run_r2.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_svm_fit" -s 2 -c 1 -f -o gpurun_out/prof_r2_svm_fit python profiles/profile_step.py --folds 138 > gpurun_out/ncu_full_r2_svm_fit.log 2>&1
tail -1 gpurun_out/ncu_full_r2_svm_fit.log
