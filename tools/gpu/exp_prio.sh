# A/B of the stream priority levels (LF_STREAM_PRIO) on the three real-chain workloads, same box
cd $GRAFT_REPO_ROOT
for wl in config2_real config3_real config4_real; do for p in 1 0 1 0; do
  echo "== $wl LF_STREAM_PRIO=$p"; bash tools/gpu/bench_env.sh xprio_${wl}_$p "LF_STREAM_PRIO=$p LF_BENCH_NO_SEEDING=1" --workload $wl --steps 10 | head -1 | cut -c1-190
done; done
