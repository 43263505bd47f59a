#!/bin/bash
# One GPU-box visit: parity suite, bench lines (headline + A/B legs), launch list under ncu.  Everything lands in gpurun_out/<tag>_*.
# usage (from the repo root on the box):  bash tools/gpu_round.sh <tag> [quick]
tag=${1:-rXX}; quick=${2:-}
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
nproc >> $out/${tag}_gpu.txt
t0=$(date +%s)
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --durations=12 > $out/${tag}_pytest.log 2>&1; echo "pytest exit $? after $(( $(date +%s) - t0 )) s" >> $out/${tag}_pytest.log
tail -25 $out/${tag}_pytest.log
timeout 400 python bench.py > $out/${tag}_bench_ours.json 2> $out/${tag}_bench_ours.err; echo "bench exit $?"
if [ -z "$quick" ]; then
timeout 200 python bench.py --separate-conversion --no-packed --steps 300 --cpu-bscans 8 > $out/${tag}_bench_sepconv.json 2>> $out/${tag}_bench_ours.err
timeout 200 python bench.py --no-numa --no-packed --steps 200 --cpu-bscans 8 > $out/${tag}_bench_nonuma.json 2>> $out/${tag}_bench_ours.err
timeout 300 python bench.py --workload 2048x1024x128-16bit --no-packed --steps 300 --cpu-bscans 8 > $out/${tag}_bench_ours_2048.json 2>> $out/${tag}_bench_ours.err
timeout 300 python bench.py --workload 2048x1024x512-16bit-config4 --no-packed --steps 60 --cpu-bscans 16 > $out/${tag}_bench_config4.json 2>> $out/${tag}_bench_ours.err
timeout 200 python bench.py --mode cufft --no-packed --steps 300 --cpu-bscans 8 > $out/${tag}_bench_cufft.json 2>> $out/${tag}_bench_ours.err
timeout 200 python bench.py --mode split --no-packed --steps 300 --cpu-bscans 8 > $out/${tag}_bench_split.json 2>> $out/${tag}_bench_ours.err
timeout 120 python tools/gpu_check.py display > $out/${tag}_display.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench_ours.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench_steps3.csv python bench.py --steps 3 --warmup 1 --no-packed --cpu-bscans 8 > $out/${tag}_bench_under_ncu.log 2>&1
fi
for f in $out/${tag}_bench_*.json; do echo "== $f"; cut -c1-1500 $f; done
echo "total $(( $(date +%s) - t0 )) s"
