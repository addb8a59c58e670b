# sweep of the per-class sort resolution (LF_BUCKETS) and the L2 fetch granularity on the default workload
cd $GRAFT_REPO_ROOT
i=0
while read -r envs; do
  i=$((i+1)); echo "== $envs"
  bash tools/gpu/bench_env.sh ${1:-xb}_$i "LF_X=1 $envs" --in-flight 1
done <<'L'

LF_L2_FETCH=32
LF_L2_FETCH=128
LF_BUCKETS=0=-1
LF_BUCKETS=0=0
LF_BUCKETS=0=2
LF_BUCKETS=0=-1,2=-1
LF_BUCKETS=0=2,2=2
LF_BUCKETS=0=-1,2=-1,4=-1,6=-1
LF_BUCKETS=0=2,2=2,4=2,6=2,18=2
LF_BUCKETS=0=3,2=3,4=3,6=3,18=3
LF_BUCKETS=0=-1 LF_L2_FETCH=32
L
