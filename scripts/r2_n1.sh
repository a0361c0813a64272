#!/bin/bash
# Round 2, 1-GPU call = what the driver runs at round end (smoke, pytest -m gpu, bench, reference arm) plus the
# profiler evidence for profiles/: launch list of the default bench command and ncu --set full captures of the
# fused kernel at CaffeNet size (LDG and TMA variants).
OUT=gpurun_out
mkdir -p $OUT
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2_smoke_n1.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/r2_smoke_n1.log
echo "== smoke under ncu (launch list; the 2-rank half must be skipped, not time out)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/r2_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2_smoke_ncu.log 2>&1; echo "rc=$?"; tail -1 $OUT/r2_smoke_ncu.log; grep -c fused_sync $OUT/r2_smoke_launches.csv
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > $OUT/r2_pytest_n1.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r2_pytest_n1.log
echo "== bench default"; timeout 900 python bench.py > $OUT/r2_bench_n1.json 2> $OUT/r2_bench_n1.err; echo "bench rc=$?"; cut -c1-1500 $OUT/r2_bench_n1.json; tail -3 $OUT/r2_bench_n1.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference > $OUT/r2_bench_ref_n1.json 2> $OUT/r2_bench_ref_n1.err; echo "rc=$?"; cut -c1-800 $OUT/r2_bench_ref_n1.json
echo "== ncu launch list (default bench command, autotune off)"
COS_BENCH_NO_AUTOTUNE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/r2_launches_caffenet_n1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-kernels --no-extras > $OUT/r2_ncu_bench.log 2>&1; echo "rc=$?"; wc -l $OUT/r2_launches_caffenet_n1.csv
echo "== ncu --set full: fused kernel at CaffeNet size, LDG then TMA"
for k in 0 1; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_sync -s 4 -c 2 -o $OUT/r2_prof_caffenet_n1_k$k -f python bench.py --sweep --sizes 232.5 --variants $( [ $k = 0 ] && echo ldg || echo tma ) --steps 3 --warmup 3 > $OUT/r2_ncu_full_k$k.log 2>&1; echo "rc=$?"
done
ls -la $OUT/*.ncu-rep 2>/dev/null
