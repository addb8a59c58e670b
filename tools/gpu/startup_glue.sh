# CUDA start-up inside the drop-in program: a small dataset, stderr of lordfast_gpu with its own timings, with strace-free wall clocks
cd $GRAFT_REPO_ROOT
T=$(mktemp -d); python integration/make_dataset.py $T --ref-len 1000000 --reads 2000 --read-len 5000 --seed 3 2>&1 | tail -2
(cd $T && $GRAFT_REPO_ROOT/oracle/_ref/lordfast --index ref.fa > /dev/null 2>&1)
for pw in 0 1 0 1; do
  echo "== LF_PREWARM_SYNC=$pw"; ( cd $T; if [ $pw = 1 ]; then export LF_PREWARM_SYNC=1; fi; LF_INIT_TRACE=1 $GRAFT_REPO_ROOT/integration/_build/lordfast_gpu --search ref.fa --seq reads.fa -t 16 -o out.sam 2>&1 | grep -E "lf_gpu|mapping|processed" | cut -c1-300 )
done
