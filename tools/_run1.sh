mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu --eps 0 --quick --no-parity --e2e-sweeps 1"
$B > gpurun_out/r02_plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_list.log 2>&1
for cfg in "c2 row" "c2 transposed" "c3 row" "c4shard row"; do
  set -- $cfg
  CMD="$B --config $1 --layout $2"
  $CMD > gpurun_out/r02_plain_$1_$2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:lasso_fused -s 4 -c 1 -f -o /tmp/r02_prof_$1_$2 $CMD > gpurun_out/r02_ncu_$1_$2.log 2>&1
  ncu -i /tmp/r02_prof_$1_$2.ncu-rep --page raw --csv > gpurun_out/r02_prof_$1_$2_raw.csv 2>/dev/null
  tail -1 gpurun_out/r02_plain_$1_$2.log | cut -c1-120
done
ncu -i /tmp/r02_prof_c2_row.ncu-rep --page source --csv > gpurun_out/r02_prof_c2_row_src.csv 2>/dev/null
cp /tmp/r02_prof_c2_row.ncu-rep gpurun_out/
du -sh gpurun_out
