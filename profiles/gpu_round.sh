#!/bin/bash
# One GPU session of a round: parity tests, bench (both arms), ncu launch lists and --set full captures of the three hot kernels.
#   gpurun --timeout 1500 -- 'bash profiles/gpu_round.sh r1p'         (outputs under gpurun_out/, summaries made by ncu_summary.py)
TAG=${1:-rX}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
python bench.py --steps 20 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
if [ "$2" != "noref" ]; then python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.json 2>> $O/bench_$TAG.err; fi
# every launch of the default command (tables, init, resets, the timed steps, e2e, world model, custom sizes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_full_$TAG.csv python bench.py --steps 5 --warmup 3 > /dev/null 2>&1
NCU="ncu --clock-control none --profile-from-start off"
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_step_$TAG.csv python bench.py --steps 3 --warmup 3 --leg step --cuda-profile step > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_pl_$TAG.csv python bench.py --steps 3 --warmup 3 --leg placements --cuda-profile placements > /dev/null 2>&1
$NCU --set full --import-source on -k regex:game_kernel -c 1 -f -o $O/prof_step_$TAG python bench.py --steps 3 --warmup 3 --leg step --cuda-profile step > $O/ncu_step_$TAG.log 2>&1
$NCU --set full --import-source on -k regex:enumerate_kernel -c 1 -f -o $O/prof_enum_$TAG python bench.py --steps 3 --warmup 3 --leg placements --cuda-profile placements > $O/ncu_enum_$TAG.log 2>&1
$NCU --set full --import-source on -k regex:game_kernel -c 1 -f -o $O/prof_expand_$TAG python bench.py --steps 3 --warmup 3 --leg placements --cuda-profile placements > $O/ncu_expand_$TAG.log 2>&1
tail -3 $O/pytest_gpu_$TAG.log; cat $O/bench_$TAG.json
