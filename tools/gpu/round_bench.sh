# the measurements of a round: default bench + reference arm, the other real-chain workloads, ncu metrics pass of a step
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${1:-r03}
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_ref.err
python bench.py > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench.err
for wl in config3_real config4_real config2_model; do python bench.py --workload $wl --steps 10 > gpurun_out/${TAG}_bench_1gpu_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err; done
python - "$TAG" <<'P'
import json,sys,glob
for f in sorted(glob.glob('gpurun_out/%s_bench_*.json'%sys.argv[1])):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if d.get('impl')=='reference': print(f,'reference', round(d['value'],1),'Mbp/s', d['config']); continue
        print(f, d['config']['workload'], 'value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['kernel_ms'],3),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']),'ms',round(d['e2e']['ms_per_step'],2),'single ms',round(d['e2e']['single_call']['ms_per_step'],2), d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f,'ERR',e)
P
