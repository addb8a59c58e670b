# one resident config-2 step under ncu: per-kernel time, instructions, lanes, DRAM bytes (serialised, cold cache)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${1:-ncu}
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:'k_myers|k_pack|k_align_prep' -c 80 --csv --log-file gpurun_out/${TAG}_metrics.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --in-flight 1 ${@:2} > gpurun_out/${TAG}.log 2>&1
tail -2 gpurun_out/${TAG}.log | cut -c1-300
python tools/ncu_metrics_table.py gpurun_out/${TAG}_metrics.csv 1 > gpurun_out/${TAG}_table.txt; cat gpurun_out/${TAG}_table.txt
