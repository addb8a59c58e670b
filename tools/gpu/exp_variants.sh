# A/B of library variants built with `make -C lordfast_b200/csrc variant NAME=.. DEFS=..`: exp_variants.sh TAG name1 name2 ...
cd $GRAFT_REPO_ROOT
TAG=$1; shift
echo "== default"; bash tools/gpu/bench_env.sh ${TAG}_default "LF_X=1" --in-flight 1
for v in "$@"; do echo "== $v"; bash tools/gpu/bench_env.sh ${TAG}_$v "LFGPU_LIB=build/liblfgpu_$v.so" --in-flight 1; done
