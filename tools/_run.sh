set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
$B > gpurun_out/plain_r01g4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01g4.csv $B > gpurun_out/ncu_l.log 2>&1
$B > gpurun_out/plain_r01g4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'pheno_lut_g4|counts_g4' -s 6 -c 2 -f -o gpurun_out/prof_g4_config2 $B > gpurun_out/ncu_a.log 2>&1
$B --workload config4 > gpurun_out/plain_r01g4_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'pheno_lut_g4|counts_g4' -s 6 -c 2 -f -o gpurun_out/prof_g4_config4 $B --workload config4 > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/ncu_a.log gpurun_out/ncu_b.log gpurun_out/ncu_l.log
ls -la gpurun_out/*.ncu-rep
