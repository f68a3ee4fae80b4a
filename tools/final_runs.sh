# Round-end measurement set (1 GPU).  Every ncu command runs only after the same command exited 0 without ncu.
timeout 400 python bench.py > gpurun_out/final_regional1m.json 2> gpurun_out/final_regional1m.err; echo rc=$?
timeout 300 python bench.py --workload brussels100k --cpu-seconds 0 --saturated-agents 0 > gpurun_out/final_brussels100k.json 2>/dev/null; echo rc=$?
timeout 300 python bench.py --workload brussels100k --no-reroute --cpu-seconds 0 --saturated-agents 0 > gpurun_out/final_brussels100k_noreroute.json 2>/dev/null; echo rc=$?
timeout 300 python bench.py --workload grid_zone --steps 4 --warmup 2 --cpu-seconds 0 --saturated-agents 0 > gpurun_out/final_grid_zone.json 2>/dev/null; echo rc=$?
timeout 300 python bench.py --workload rgg_hub --steps 10 --warmup 3 --cpu-seconds 0 --saturated-agents 0 > gpurun_out/final_rgg_hub.json 2>/dev/null; echo rc=$?
timeout 300 python bench.py --workload regional1m --steps 4 --warmup 2 --cpu-seconds 0 --saturated-agents 0 --opt tree_sides=1 > gpurun_out/final_regional1m_cold.json 2>/dev/null; echo rc=$?
for f in regional1m brussels100k brussels100k_noreroute grid_zone rgg_hub regional1m_cold; do python tools/summ.py < gpurun_out/final_$f.json | cut -c1-330; done
timeout 300 python bench.py --steps 2 --warmup 3 --cpu-seconds 0 --saturated-agents 0 > /dev/null 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/launches_r1_final2.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 0 --saturated-agents 0 > gpurun_out/ncu_final2.log 2>&1; echo rc=$?
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_tick_loop -s 3 -c 1 -o gpurun_out/prof_tick_loop_r1 -f python bench.py --steps 1 --warmup 3 --cpu-seconds 0 --saturated-agents 0 > gpurun_out/ncu_tick.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
