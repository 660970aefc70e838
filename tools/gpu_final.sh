#!/bin/bash
# round-end measurement set (one GPU).  usage: gpu_final.sh [tests]   ("tests": smoke + pytest + sanitizers first)
# Summaries are produced ON the box (ncu is there) into gpurun_out/profiles/ and the .ncu-rep files are dropped:
# only gpurun_out/ travels back and it is capped at 64 MiB.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > gpurun_out/gpu.txt
if [ "$1" = tests ]; then
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 2>&1 | tail -3 | tee gpurun_out/pytest.log
echo "== sanitizer"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -1 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -1 gpurun_out/sanitize_racecheck.log
fi
echo "== bench default"; timeout 900 python bench.py 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.log; cut -c1-300 gpurun_out/bench_default.log
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference.log; cut -c1-200 gpurun_out/bench_reference.log
for w in poisson er cari; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 3 2>gpurun_out/bench_$w.err | tail -1 > gpurun_out/bench_$w.log; cut -c1-160 gpurun_out/bench_$w.log
done
timeout 900 python bench.py --workload rmat --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>gpurun_out/bench_rmat.err | tail -1 > gpurun_out/bench_rmat.log; cut -c1-160 gpurun_out/bench_rmat.log
echo "== ncu launch lists (rect, rmat)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_rect.csv \
  python bench.py --workload rect --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_ll.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_rmat.csv \
  python bench.py --workload rmat --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_ll_rmat.log 2>&1
echo "== ncu full: rect sort / long / copy kernels, ER fused, Poisson tiny, R-MAT long"
timeout 900 ncu --set full --clock-control none -k regex:"k_esc_numeric_warp|k_bitonic_numeric_cta|k_copy_rows|k_long_chunk_sort|k_long_merge|k_long_reduce|k_flops" -s 60 -c 30 -f -o gpurun_out/prof_rect_numeric \
  python bench.py --workload rect --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fused_light" -s 3 -c 1 -f -o gpurun_out/prof_er_fused \
  python bench.py --workload er --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_fused_tiny4" -s 3 -c 1 -f -o gpurun_out/prof_poisson_tiny \
  python bench.py --workload poisson --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_c.log 2>&1
timeout 1200 ncu --set full --clock-control none -k regex:"k_long_chunk_sort|k_long_merge|k_long_reduce" -s 40 -c 8 -f -o gpurun_out/prof_rmat_long \
  python bench.py --workload rmat --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_rmat_long.log 2>&1
ls -la gpurun_out/*.ncu-rep
python tools/summarize_profiles.py r02 gpurun_out/profiles > gpurun_out/summarize.log 2>&1; tail -3 gpurun_out/summarize.log
# keep only what fits the 64 MiB cap: the summaries, the logs, and the two small captures
rm -f gpurun_out/prof_rect_numeric.ncu-rep gpurun_out/prof_rmat_long.ncu-rep
du -sh gpurun_out
