# Round-2 profiling pass (one B200, under gpurun).  Order as the recipe asks: the plain run first,
# then the per-launch device times of one warm-up + one measured 138-fold batch, then ONE
# `--set full` capture per kernel (named by its demangled name, so the <double> and <float>
# instances of the tile Jacobi are captured separately), including the streaming kernels.  The script
# runs two warm-up batches (cold solves, then the one-time move to the all-trials eigenbasis) and one
# measured batch: the skip counts select the measured batch's launches (k_eig_tile<double> #8 = the
# warm-started target view solve with eigenvectors).
set -x
python profiles/profile_step.py --stages --folds 138 > gpurun_out/plain_r2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r2.csv python profiles/profile_step.py --folds 138 > gpurun_out/ncu_run_r2.log 2>&1
for spec in "k_eig_tile<double>:eig_tile_f64:8:1" "k_eig_tile<float>:eig_tile_f32:4:1" "k_gram_tc:gram_tc:2:1" \
            "k_proj_tc\(:proj_tc:2:1" "k_split_tf32_batched:split_tf32_batched:2:1" "k_class_mean:class_mean:2:1" \
            "k_colsum:colsum:5:1" "k_gram_tn<double>:gram_tn_f64:5:1" "k_svm_fit:svm_fit:2:1" \
            "k_chol_inv:chol_inv:22:1" "k_gemm_tc_nt:gemm_tc_nt:19:1" "k_sgemm:sgemm:52:1"; do
  k=$(echo "$spec" | cut -d: -f1); tag=$(echo "$spec" | cut -d: -f2); s=$(echo "$spec" | cut -d: -f3); c=$(echo "$spec" | cut -d: -f4)
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$k" -s $s -c $c -f \
      -o gpurun_out/prof_r2_$tag python profiles/profile_step.py --folds 138 > gpurun_out/ncu_full_r2_$tag.log 2>&1
  tail -1 gpurun_out/ncu_full_r2_$tag.log
done
tail -4 gpurun_out/plain_r2.log
