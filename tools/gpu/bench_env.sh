# usage: bench_env.sh TAG "ENV=.. ENV=.." [bench args]   -- one short bench run with environment overrides, summary line
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1; ENVS=$2; shift 2
env $ENVS python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
python - "$TAG" <<'P'
import json,sys
f='gpurun_out/%s.json'%sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['kernel_ms'],3),'frac',round(d['roofline']['frac'],3),'e2e ms',round(d['e2e']['ms_per_step'],2),'single',round(d['e2e']['single_call']['ms_per_step'],2), d['e2e']['host_phase_ms'], d['config']['workload'])
    print('   ', {k:v[1] for k,v in d['class_timeline_ms'].items()})
except Exception as e: print(f, 'ERR', e, open('gpurun_out/%s.err'%sys.argv[1]).read()[-2000:])
P
