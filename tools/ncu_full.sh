set -x
N="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
P="python tools/profile_step.py --config cfg2 --steps 1"
timeout 400 $N -k regex:"gemm_tc_kernel<.int.3" -s 11 -c 1 -o gpurun_out/full_da -f $P > gpurun_out/full_da.log 2>&1
timeout 400 $N -k regex:"gemm_tn_tc_kernel<.bool.1" -s 11 -c 1 -o gpurun_out/full_dw2f -f $P > gpurun_out/full_dw2f.log 2>&1
timeout 400 $N -k regex:"gemm_tn_tc_kernel<.bool.0" -s 18 -c 1 -o gpurun_out/full_dw1f -f $P > gpurun_out/full_dw1f.log 2>&1
timeout 400 $N -k regex:"dwconv_patch_wgrad_kernel<.int.8|dwconv_patch_kernel<.int.8" -s 4 -c 2 -o gpurun_out/full_dwbwd -f $P > gpurun_out/full_dwbwd.log 2>&1
timeout 400 $N -k regex:"initial_conv_wgrad|stem_bwd|ln_rows_bwd" -s 15 -c 3 -o gpurun_out/full_misc -f $P > gpurun_out/full_misc.log 2>&1
ls -la gpurun_out/
