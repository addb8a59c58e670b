# seeding kernels of one chunk under ncu: per-kernel time, instructions, lanes, DRAM bytes (serialised, cold cache)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${1:-ncu_seed}
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,launch__grid_size,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:'k_seed' --launch-skip 14 -c 7 --csv --log-file gpurun_out/${TAG}_metrics.csv \
  python bench_seed.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}.log 2>&1
tail -1 gpurun_out/${TAG}.log | cut -c1-200
python tools/ncu_metrics_table.py gpurun_out/${TAG}_metrics.csv 1 > gpurun_out/${TAG}_table.txt; cat gpurun_out/${TAG}_table.txt
