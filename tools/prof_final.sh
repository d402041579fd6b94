# final round-2 captures (one gpurun call, 1 GPU): launch lists of the SimGCL / XSimGCL / NGCF steps and ncu --set full of the new kernels
set -x
for w in simgcl ngcf; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${w}_final.csv python tools/prof_r2.py $w > gpurun_out/prof_${w}.log 2>&1
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:nce_flash -c 2 -o gpurun_out/nce_flash_final -f python tools/prof_r2.py infonce > gpurun_out/prof_nce.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:ngcf_dense -c 6 -o gpurun_out/ngcf_dense_final -f python tools/prof_r2.py ngcf > gpurun_out/prof_ngcf2.log 2>&1
ls -la gpurun_out/*final*
