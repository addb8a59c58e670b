# stream priority levels x calls in flight on the default workload, repeated (same box)
cd $GRAFT_REPO_ROOT
for rep in 1 2; do for inf in 4 2; do for p in 1 0; do
  bash tools/gpu/bench_env.sh xp2_${inf}_${p}_$rep "LF_STREAM_PRIO=$p LF_BENCH_NO_SEEDING=1" --in-flight $inf --steps 12 > /dev/null
  python - <<P
import json
d=json.loads(open("gpurun_out/xp2_${inf}_${p}_$rep.json").read().strip().splitlines()[-1]); e=d["e2e"]
print("in-flight $inf prio $p rep $rep: in-flight ms", round(e["in_flight_tried"]["ms_per_step"],2), "single ms", round(e["single_call"]["ms_per_step"],2))
P
done; done; done
