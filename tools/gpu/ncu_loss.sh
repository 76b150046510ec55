timeout 400 ncu --set full --clock-control none --import-source on -k regex:'gemm_tn|grad_tiles|soft_pass2|col_lse' -s 14 -c 7 \
    -o gpurun_out/r2b_prof_loss python tools/loss_profile.py 32768 > /dev/null 2> gpurun_out/r2b_ncu_loss.err; echo "loss rc=$?"
tail -n 2 gpurun_out/r2b_ncu_loss.err; ls -la gpurun_out/r2b_prof_loss.ncu-rep
