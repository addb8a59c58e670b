# round-2 closing measurements on one GPU: whole GPU suite, benches, whole-program comparison with and without GPU seeding
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -v "^\[" | tail -6 > gpurun_out/r03c_pytest_gpu.txt; cat gpurun_out/r03c_pytest_gpu.txt
python tools/bench_lordfast_e2e.py --reads 20000 --replicate 5 > gpurun_out/r03c_lordfast_e2e_config2_x5.json 2> gpurun_out/r03c_e2e.err; tail -c 1500 gpurun_out/r03c_lordfast_e2e_config2_x5.json
LF_GPU_SEED=0 python tools/bench_lordfast_e2e.py --reads 20000 --replicate 5 --repeat 1 > gpurun_out/r03c_lordfast_e2e_config2_x5_cpu_seeding.json 2>> gpurun_out/r03c_e2e.err; tail -c 700 gpurun_out/r03c_lordfast_e2e_config2_x5_cpu_seeding.json
python bench_seed.py > gpurun_out/r03c_bench_seed_1gpu.json 2> gpurun_out/r03c_seed.err
bash tools/gpu/round_bench.sh r03c
