# Re-capture of the kernels changed after the full pass (ncu_all_r2.sh): launch list + `--set full`
# of the three tcgen05 kernels and the SVM kernel.  Same script, same skip counts.
set -x
python profiles/profile_step.py --stages --folds 138 > gpurun_out/plain_r2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r2.csv python profiles/profile_step.py --folds 138 > gpurun_out/ncu_run_r2.log 2>&1
for spec in "k_gram_tc:gram_tc:2:1" "k_proj_tc\(:proj_tc:2:1" "k_gemm_tc_nt:gemm_tc_nt:19:1" "k_svm_fit:svm_fit:2:1"; do
  k=$(echo "$spec" | cut -d: -f1); tag=$(echo "$spec" | cut -d: -f2); s=$(echo "$spec" | cut -d: -f3); c=$(echo "$spec" | cut -d: -f4)
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$k" -s $s -c $c -f \
      -o gpurun_out/prof_r2_$tag python profiles/profile_step.py --folds 138 > gpurun_out/ncu_full_r2_$tag.log 2>&1
  tail -1 gpurun_out/ncu_full_r2_$tag.log
done
