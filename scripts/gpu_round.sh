#!/bin/bash
# One GPU-box call: parity tests, the bench lines of every configuration, the ncu launch list and full captures of
# the dominant kernels.  Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh <tag>
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_gpu.txt 2>&1
nproc >> $out/${tag}_gpu.txt; grep -m1 "model name" /proc/cpuinfo >> $out/${tag}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -2 $out/${tag}_smoke.log
timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cat $out/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err; cat $out/${tag}_bench_ref.json
for c in c1 c2; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > $out/${tag}_bench_$c.json 2>> $out/${tag}_bench.err; echo "bench $c rc=$?"
done
for c in c4 c5; do   # crowded configurations: the oracle side of the parity sample is slow, check 2 streams
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 --parity-streams 2 > $out/${tag}_bench_$c.json 2>> $out/${tag}_bench.err; echo "bench $c rc=$?"
done
# launch list of the same command (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --skip-e2e --parity-streams 0 --cli-segments 0 > $out/${tag}_bench_under_ncu.log 2>&1
# full captures of one launch of each kernel at the bench's own size (ncu replays the launch ~40 times)
SEG=${2:-150}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sort_warp_kernel -s 1 -c 1 -f -o $out/${tag}_prof_sort \
    python bench.py --segments $SEG --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e --parity-streams 0 --cli-segments 0 > $out/${tag}_ncu_sort.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:softnms_kernel -s 1 -c 1 -f -o $out/${tag}_prof_nms \
    python bench.py --segments $SEG --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e --parity-streams 0 --cli-segments 0 > $out/${tag}_ncu_nms.log 2>&1
ls -la $out | tail -20
