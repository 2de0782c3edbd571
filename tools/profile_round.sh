#!/bin/bash
# Round-end evidence: GPU tests, bench line, ncu launch list, one full capture of the field kernel.
# usage: tools/profile_round.sh <tag>      (outputs under gpurun_out/<tag>_*)
tag=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:field_forward_kernel -s 3 -c 3 \
  -o gpurun_out/${tag}_field_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 300 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/${tag}_train_bench.json 2>/dev/null; python tools/print_train.py < gpurun_out/${tag}_train_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/${tag}_train_launches.csv \
  python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/${tag}_train_ncu.log 2>&1; echo "ncu train rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:composite_kernel -s 2 -c 1 \
  -o gpurun_out/${tag}_composite_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_comp.log 2>&1; echo "ncu composite rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:img_wgrad_kernel -s 60 -c 2 \
  -o gpurun_out/${tag}_wgrad_full -f python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/${tag}_ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:img_gemm_kernel -s 80 -c 4 \
  -o gpurun_out/${tag}_gemm_full -f python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/${tag}_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
