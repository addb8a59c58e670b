set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.txt 2>&1; tail -5 gpurun_out/s1_pytest.txt
for gk in 7 0 1 3; do
LF_GROUPK=$gk python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s1_bench_gk$gk.json 2> gpurun_out/s1_bench_gk$gk.err
done
LF_GROUPK=7 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sv-frac 0 > gpurun_out/s1_bench_gk7_sv0.json 2>/dev/null
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/s1_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['kernel_ms'],3),'frac',round(d['roofline']['frac'],3),'e2e ms',round(d['e2e']['ms_per_step'],2),'single',round(d['e2e']['single_call']['ms_per_step'],2), d['e2e_chains']['host_phase_ms'])
        print('   ', d['class_timeline_ms'])
    except Exception as e: print(f, 'ERR', e)
P
