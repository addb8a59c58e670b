# ncu --set full of one kernel (regex $1) in a microbench point: $2 = length, $3 = divergence
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$4
ncu --set full --import-source on --clock-control none -k regex:"$1" -c 2 -o gpurun_out/${TAG} -f python bench_kernels.py --lengths $2 --divs $3 --pairs 65536 --check 2 > gpurun_out/${TAG}.log 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
python - "$TAG" <<'P'
import csv,sys
rows=list(csv.reader(open('gpurun_out/%s_raw.csv'%sys.argv[1])))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct']
stall=[h for h in hdr if 'smsp__average_warp' in h and 'issue_stalled' in h and h.endswith('.ratio')] or [h for h in hdr if 'warp_issue_stalled' in h and 'per_warp_active' in h]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    for w in want: print(w, '=', d.get(w))
    st=sorted(((float(d[h].replace(',','')) if d[h] not in ('','n/a') else 0,h) for h in stall), reverse=True)[:8]
    for v,h in st: print('   stall', h.split('issue_stalled_')[-1][:40], round(v,2))
P
