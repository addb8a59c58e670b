# calls in flight x class streams per context (LF_STREAMS): e2e ms per chunk
cd $GRAFT_REPO_ROOT
for inf in 4 2 3; do for st in 15 8 6 4; do
  echo "== in-flight $inf LF_STREAMS=$st"; bash tools/gpu/bench_env.sh xst_${inf}_${st} "LF_STREAMS=$st LF_BENCH_NO_SEEDING=1" --in-flight $inf --steps 12 | head -1 | cut -c1-140
done; done
