#!/bin/bash
# Round artefacts on ONE B200 (run under gpurun from the repo root): bench lines of every workload, the reference arm,
# the ncu launch list of the bench command and one `ncu --set full` capture of the steady-state half-step kernels.
# Everything lands in gpurun_out/ with the prefix given as $1 (default r02).
P=${1:-r02}
O=gpurun_out
set -x
timeout 600 python bench.py --steps 20 --warmup 3 > $O/${P}_bench_n1.log 2>&1 || exit 1
timeout 300 python bench.py --workload netflix --steps 5 --warmup 3 --no-e2e > $O/${P}_bench_n1_netflix.log 2>&1
timeout 300 python bench.py --workload powerlaw256 --steps 3 --warmup 3 --no-e2e > $O/${P}_bench_n1_powerlaw256.log 2>&1
timeout 300 python bench.py --workload rank --steps 5 --warmup 3 > $O/${P}_bench_n1_rank.log 2>&1
timeout 300 python bench.py --algo tcgen05_direct --steps 5 --warmup 3 --no-e2e > $O/${P}_bench_n1_direct.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${P}_bench_reference.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"wmf|als_half_step|tc_|dual" -c 400 --csv \
    --log-file $O/${P}_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e > $O/${P}_ncu_launches_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"als_half_step_(dual|tc)_kernel" -s 22 -c 6 \
    -o $O/${P}_prof_half_step python scripts/tc_profile.py --steady > $O/${P}_ncu_full_run.log 2>&1
ls -la $O | grep ${P}_
