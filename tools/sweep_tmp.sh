run() { timeout 200 python bench.py --steps 2 --warmup 1 --cpu-seconds 0 --saturated-agents 0 "$@" 2>/dev/null | python tools/summ2.py; }
echo R128; run --workload regional1m --opt tree_sides=1 --opt tree_threads=128 --opt tree_scratch_bytes=64e9
echo R64; run --workload regional1m --opt tree_sides=1 --opt tree_threads=64 --opt tree_scratch_bytes=80e9
echo G128; run --workload grid_zone --opt tree_threads=128 --opt tree_scratch_bytes=90e9
echo B128; run --workload brussels100k --opt tree_threads=128
