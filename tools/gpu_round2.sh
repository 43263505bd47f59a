#!/bin/bash
# One GPU-box visit of round 2: the whole GPU suite, smoke, the bench lines of both arms, the ncu launch list of the bench command and
# one `ncu --set full` capture per kernel family.  Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round2.sh r02m'          (add `quick` to skip the ncu passes)
tag=${1:-r02x}; quick=${2:-}; out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1; nproc >> $out/${tag}_gpu.txt
t0=$(date +%s)
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --durations=10 > $out/${tag}_pytest.log 2>&1; echo "pytest exit $? after $(( $(date +%s) - t0 )) s" >> $out/${tag}_pytest.log
tail -18 $out/${tag}_pytest.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $out/${tag}_smoke.txt
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench.err
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_ours.json 2>> $out/${tag}_bench.err; echo "bench exit $?"
tail -3 $out/${tag}_bench.err
if [ -z "$quick" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench_steps3.csv python bench.py --steps 3 --warmup 1 --no-packed --no-secondary --cpu-bscans 8 > $out/${tag}_bench_under_ncu.log 2>&1
for t in fused1024:oct_fused_kernel fused2048:oct_fused_kernel generic1664:oct_generic_kernel cufft1024:oct_pre_kernel cufft1024:oct_post_kernel u8_1024:oct_fused_kernel; do
  name=${t%%:*}; k=${t#*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $out/${tag}_ncu_${name}_${k} python tools/ncu_targets.py $name > $out/${tag}_ncu_${name}_${k}.log 2>&1
done
ls -la $out/${tag}_ncu_*.ncu-rep
fi
echo "total $(( $(date +%s) - t0 )) s"
