#!/bin/bash
# One gpurun call of the round: GPU test-suite, the bench line, effect-kernel launch lists.
# usage (on the GPU box, from the repo root): bash tools/gpu_round.sh <tag> [sections...]
tag=${1:-r02}; shift
sections=${@:-"tests bench fx"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
for s in $sections; do
case $s in
tests)
  timeout 1700 python -m pytest tests -m gpu -q --timeout 1500 -x > gpurun_out/${tag}_pytest_gpu.log 2>&1
  tail -5 gpurun_out/${tag}_pytest_gpu.log ;;
tests_fast)
  timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x --deselect tests/test_gpu_atsize.py > gpurun_out/${tag}_pytest_gpu_fast.log 2>&1
  tail -5 gpurun_out/${tag}_pytest_gpu_fast.log ;;
bench)
  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
  tail -c 3000 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err ;;
fx)
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_conv -c 60 --csv \
     --log-file gpurun_out/${tag}_launches_fx_conv.csv python tools/bench_effects.py --effect conv --voices 4096 --slots 32 --steps 4 > gpurun_out/${tag}_fx_conv_ncu.log 2>&1
  timeout 300 python tools/bench_effects.py --effect conv --voices 4096 --slots 32 > gpurun_out/${tag}_fx_conv.log 2>&1
  timeout 300 python tools/bench_effects.py --effect conv --voices 4096 --slots 16 >> gpurun_out/${tag}_fx_conv.log 2>&1
  tail -2 gpurun_out/${tag}_fx_conv.log; grep k_conv_mac gpurun_out/${tag}_launches_fx_conv.csv | tail -3 ;;
esac
done
