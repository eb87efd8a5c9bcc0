#!/bin/bash
# Round-2 record run on ONE GPU (gpurun --timeout 2400 -- bash tools/record_run_1gpu.sh): the record -- smoke, full bench (+ reference arm), ncu captures,
# launch list, compute-sanitizer (default build, then the per-thread-arrival debug build)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_c13_smoke.log 2>&1; tail -1 gpurun_out/r2_c13_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/r2_c13_bench.json 2> gpurun_out/r2_c13_bench.err
tail -c 300 gpurun_out/r2_c13_bench.err
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r2_c13_bench_ref.json 2> gpurun_out/r2_c13_bench_ref.err
tail -c 400 gpurun_out/r2_c13_bench_ref.json
# launch list of the default run (every launch with its device time)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_c13_launches.log 2>&1
wc -l gpurun_out/r2_launches_bench.csv
# top kernel of the headline at the bench size (one launch over 1M graphs)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_tc_kernel -s 1 -c 1 \
    -o gpurun_out/r2_fused_tc_c2_1m python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2_c13_ncu_c2.log 2>&1
# C5 aggregation kernels, hub hints off / on (second step's launches)
for mb in 0 40; do
GNNB_HUB_L2_MB=$mb timeout 600 ncu --set full --clock-control none -k regex:agg_ -s 6 -c 6 \
    -o gpurun_out/r2_agg_c5_hub$mb python bench.py --workload c5_gcn_large --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_c13_ncu_c5_$mb.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -4
# compute-sanitizer over every kernel family
for tool in memcheck synccheck racecheck; do
  echo "== $tool" >> gpurun_out/r2_sanitizer_default.txt
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_smoke.py >> gpurun_out/r2_sanitizer_default.txt 2>&1
done
grep -E "SUMMARY|md5" gpurun_out/r2_sanitizer_default.txt | tail -20
# the same racecheck on the build where every worker thread arrives on the hand-off mbarriers itself
GNNB_NVCC_EXTRA=-DGNNB_TC_THREAD_ARRIVALS=1 python -m gnn_builder_b200.build --force > gpurun_out/r2_c13_rebuild.log 2>&1
echo "== racecheck, -DGNNB_TC_THREAD_ARRIVALS=1" > gpurun_out/r2_sanitizer_thread_arrivals.txt
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py >> gpurun_out/r2_sanitizer_thread_arrivals.txt 2>&1
grep -E "SUMMARY|md5" gpurun_out/r2_sanitizer_thread_arrivals.txt | tail -8
