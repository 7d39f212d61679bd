#!/usr/bin/env bash
# One gpurun call that produces the round's evidence (each leg under its own timeout, logs under gpurun_out/):
#   3. pytest -m gpu in ONE process (as the driver runs it)
#   4. bench.py (N = 1, defaults)
#   1. ncu launch list (gpu__time_duration.sum) of a 2-PC-step slice of the bench command
#   2. ncu --set full capture of the first igemm / GroupNorm / attention launches of one score-network forward
# (tests and bench first: they matter most if the box goes away)
#     gpurun --timeout 1500 -- 'bash tools/round_gpu_run.sh'        (LEGS=ncu: legs 1-2 only; LEGS=tests: legs 3-4 only)
set -u
O=gpurun_out
mkdir -p $O
TAG=${TAG:-r2_final}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1

if [ "${LEGS:-all}" != "ncu" ]; then
echo "== leg 3: pytest -m gpu (one process)"; date +%T
timeout 600 python -m pytest tests/ -q -s -m gpu --durations=12 -p no:cacheprovider > $O/${TAG}_pytest_gpu.log 2>&1
echo "rc=$?"
tail -25 $O/${TAG}_pytest_gpu.log

echo "== leg 4: bench.py"; date +%T
timeout 600 python bench.py > $O/${TAG}_bench_n1.log 2> $O/${TAG}_bench_n1.err
echo "rc=$?"
tail -c 3000 $O/${TAG}_bench_n1.log
date +%T
fi

if [ "${LEGS:-all}" != "tests" ]; then
echo "== leg 1: ncu launch list of the bench slice"; date +%T
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv \
    --log-file $O/${TAG}_ncu_launches_bench_slice.csv \
    python bench.py --steps 1 --warmup 1 --num-scales 2 --skip-train --skip-cpu --skip-extras > $O/${TAG}_ncu_launches_bench_slice.log 2>&1
echo "rc=$?"

echo "== leg 2: ncu --set full, igemm + GroupNorm kernels of the first forward"; date +%T
timeout 420 ncu --set full --clock-control none --import-source on -k 'regex:igemm_kernel|igemm_halo_kernel|gn_apply|gn_stats|attention_fwd' -c 34 \
    -o /tmp/${TAG}_ncu_full_forward -f python tools/quick_bench.py --infer > $O/${TAG}_ncu_full_forward.log 2>&1
echo "rc=$?"
# the .ncu-rep (70+ MB) exceeds what gpurun copies back: keep its raw page as CSV
ncu -i /tmp/${TAG}_ncu_full_forward.ncu-rep --page raw --csv > $O/${TAG}_ncu_full_forward_raw.csv 2>/dev/null
ls -la $O/${TAG}_ncu_full_forward_raw.csv
fi

