# closing evidence of round 2 on one GPU: memcheck of the new kernels, ncu --set full of the alignment kernels of a step
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
# 1. compute-sanitizer memcheck: seeding (small case on the GPU), pack + alignment + chain operator on a small chunk
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python - > gpurun_out/r03f_memcheck.txt 2>&1 <<'P'
import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import numpy as np
from lordfast_b200 import api, sim, fmindex
import test_seed
ref, reads, off = test_seed.small_case(seed=23, ref_len=40_000, n_reads=20, read_len=1500)
g = api.LfGpu(sim.pack_pac(ref), len(ref))
g.seed_init(fmindex.build(test_seed.CODE[ref], k_cache=10))
r = g.seed_batch(reads, off, sampling_count=300)
print("seeds", len(r[0]), len(r[2]))
w = sim.make_workload(300_000, 300, 3000, 0.12, 0.15, seed=5, sv_frac=0.3)
g2 = api.LfGpu(w.pac, len(w.ref))
seeds, chains = api.workload_chains(w)
recs, text, st = g2.align_chains(w.reads, w.read_off.astype(np.uint64), np.array([0], dtype=np.int64), np.array([len(w.ref)], dtype=np.int32), seeds, chains)
print("records", len(recs), "text", len(text))
P
echo "memcheck rc=$?" >> gpurun_out/r03f_memcheck.txt; tail -6 gpurun_out/r03f_memcheck.txt
# 2. ncu --set full of the alignment kernels of one resident step
ncu --set full --clock-control none --import-source on -k regex:'k_myers_bandreg|k_pack_reads' --launch-skip 0 -c 10 -o gpurun_out/r03f_full python bench.py --steps 1 --warmup 0 --no-cpu-baseline --in-flight 1 > gpurun_out/r03f_full.log 2>&1
ncu -i gpurun_out/r03f_full.ncu-rep --page raw --csv > gpurun_out/r03f_full_raw.csv 2>/dev/null; ls -la gpurun_out/r03f_full*; rm -f gpurun_out/r03f_full.ncu-rep   # the report itself is too big to travel back; the raw page has every metric
