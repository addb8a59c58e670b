# kernels of ONE lf_gpu_align_chains call on the config-2 chunk with the reference's chains: launch list with durations (serialised by ncu)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,launch__grid_size,smsp__inst_executed.sum --clock-control none -k regex:'k_emit_slots|k_ksw_extend|k_chain|k_extend_prep|k_myers_large|k_myers_group' --launch-skip 120 -c 60 --csv --log-file gpurun_out/r03k_chain_call_launches.csv \
  python tools/gpu/chain_trace_real.py > gpurun_out/r03k_chain_call.log 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r03k_chain_call_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
out={}
for r in rows[1:]:
    out.setdefault(r[ii],{'k':r[ki][:60]})[r[mi]]=r[vi]
for i,v in list(out.items())[-45:]:
    print(i, v['k'], 'us', v.get('gpu__time_duration.sum'), 'grid', v.get('launch__grid_size'), 'inst', v.get('smsp__inst_executed.sum'))
P
