#!/bin/bash
# Round-2 ncu captures (one GPU, under gpurun): launch list of the default bench, then `--set full` captures of
# the kernels of the timed region.  The same command runs once without ncu first and must exit 0.
CMD="python bench.py --steps 3 --warmup 2 --cpu-seconds 0 --saturated-agents 0"
O=gpurun_out
$CMD > $O/r2_plain.json 2> $O/r2_plain.err || { echo "plain run failed"; tail -5 $O/r2_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv \
    --log-file $O/r2_launches.csv $CMD > $O/r2_ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_tick_loop|k_label_query|k_label_paths' -s 6 -c 6 \
    -f -o $O/r2_prof_main $CMD > $O/r2_ncu_main.log 2>&1
echo "main rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_recount|k_travel_times' -c 2 \
    -f -o $O/r2_prof_aux $CMD > $O/r2_ncu_aux.log 2>&1
echo "aux rc=$?"
ncu --set full --clock-control none -k regex:k_build_labels -s 280 -c 2 \
    -f -o $O/r2_prof_labels $CMD > $O/r2_ncu_labels.log 2>&1
echo "labels rc=$?"
for f in main aux labels; do
  ncu -i $O/r2_prof_$f.ncu-rep --page raw --csv > $O/r2_prof_${f}_raw.csv 2>/dev/null
done
ncu -i $O/r2_prof_main.ncu-rep --page source --csv > $O/r2_prof_main_source.csv 2>/dev/null
ls -la $O/r2_prof_* | head -20
# keep the transfer small: the reports with source can be large
du -sm $O/*.ncu-rep
