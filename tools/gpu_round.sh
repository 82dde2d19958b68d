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
  timeout 1700 python -m pytest tests -m gpu -q --timeout 1500 > gpurun_out/${tag}_pytest_gpu.log 2>&1
  grep -E '^(FAILED|ERROR)' gpurun_out/${tag}_pytest_gpu.log | head -30; tail -3 gpurun_out/${tag}_pytest_gpu.log ;;
tests_fast)
  timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x --deselect tests/test_gpu_atsize.py > gpurun_out/${tag}_pytest_gpu_fast.log 2>&1
  tail -5 gpurun_out/${tag}_pytest_gpu_fast.log ;;
bench)
  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
  tail -c 3000 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err ;;
launches)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
     --log-file gpurun_out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1
  grep -c k_ gpurun_out/${tag}_launches_bench.csv ;;
configs)
  for c in 1 2 3 "4a --voices 8192" "4b --voices 8192"; do timeout 600 python tools/bench_configs.py --config $c >> gpurun_out/${tag}_configs_n1.jsonl 2>> gpurun_out/${tag}_configs_n1.err; done
  timeout 900 python tools/bench_configs.py --config 5 --voices 131072 --slot-share 8 --steps 8 >> gpurun_out/${tag}_configs_n1.jsonl 2>> gpurun_out/${tag}_configs_n1.err
  python - <<PY
import json
for l in open('gpurun_out/${tag}_configs_n1.jsonl'):
    j=json.loads(l); print(j['config'], j['voices_total'], j['slots'], round(j['ms_per_update'],4), j.get('stage_us_rank0'))
PY
  ;;
multi)
  NG=$(nvidia-smi -L | wc -l)
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
  timeout 900 python -m pytest tests/test_gpu_shard.py -m gpu -q --timeout 600 > gpurun_out/${tag}_pytest_shard_n${NG}.log 2>&1
  tail -3 gpurun_out/${tag}_pytest_shard_n${NG}.log
  timeout 900 $TR --master-port 29511 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n${NG}.json 2> gpurun_out/${tag}_bench_n${NG}.err
  tail -c 2500 gpurun_out/${tag}_bench_n${NG}.json; tail -3 gpurun_out/${tag}_bench_n${NG}.err
  timeout 600 $TR --master-port 29512 bench.py --impl reference --gpus $NG --steps 20 --warmup 5 > gpurun_out/${tag}_bench_ref_n${NG}.json 2>> gpurun_out/${tag}_bench_n${NG}.err
  tail -c 600 gpurun_out/${tag}_bench_ref_n${NG}.json
  for c in "2" "2 --transport nccl" "3" "5 --voices $((131072*NG)) --slot-share $((8/NG)) --steps 8"; do
    timeout 900 $TR --master-port 29513 tools/bench_configs.py --config $c --gpus $NG >> gpurun_out/${tag}_configs_n${NG}.jsonl 2>> gpurun_out/${tag}_configs_n${NG}.err
  done
  python - <<PY
import json
for l in open('gpurun_out/${tag}_configs_n${NG}.jsonl'):
    if not l.startswith('{'): continue
    j=json.loads(l); print(j['config'], j['n_gpus'], j['voices_total'], j['slots'], j.get('transport'), round(j['ms_per_update'],4), j.get('stage_us_rank0'), j.get('collective'))
PY
  ;;
multi2)
  NG=$(nvidia-smi -L | wc -l)
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
  timeout 600 python -m pytest tests/test_gpu_shard.py -m gpu -q --timeout 500 > gpurun_out/${tag}_pytest_shard_n${NG}.log 2>&1
  tail -2 gpurun_out/${tag}_pytest_shard_n${NG}.log
  timeout 600 $TR --master-port 29511 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n${NG}.json 2> gpurun_out/${tag}_bench_n${NG}.err
  python - <<PY
import json
j=json.loads(open('gpurun_out/${tag}_bench_n${NG}.json').read().strip().split('\n')[-1]); print('N', j['n_gpus'], 'ms', j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'], 'reduce_us', j['collective']['reduce_us'], j['verification']['p2p'])
PY
  ;;
multi8)
  NG=$(nvidia-smi -L | wc -l)
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
  timeout 600 $TR --master-port 29511 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n${NG}.json 2> gpurun_out/${tag}_bench_n${NG}.err
  tail -c 1500 gpurun_out/${tag}_bench_n${NG}.json; tail -3 gpurun_out/${tag}_bench_n${NG}.err
  for c in "2 --voices $((4096*NG))" "3 --voices $((16384*NG/8)) " "5"; do
    timeout 900 $TR --master-port 29513 tools/bench_configs.py --config $c --gpus $NG >> gpurun_out/${tag}_configs_n${NG}.jsonl 2>> gpurun_out/${tag}_configs_n${NG}.err
  done
  python - <<PY
import json
for l in open('gpurun_out/${tag}_configs_n${NG}.jsonl'):
    if not l.startswith('{'): continue
    j=json.loads(l); print(j['config'], j['n_gpus'], j['voices_total'], j['slots'], j.get('transport'), round(j['ms_per_update'],4), j.get('stage_us_rank0'))
PY
  ;;
l3)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_send\|k_reduce\|k_reverb\|k_slot\|k_post\|k_filters -c 120 --csv \
     --log-file gpurun_out/${tag}_launches_cfg3.csv python tools/bench_configs.py --config 3 --steps 2 --warmup 2 > gpurun_out/${tag}_cfg3_ncu.log 2>&1
  python - <<PY
import csv,re,collections
rows=[r for r in csv.reader(open('gpurun_out/${tag}_launches_cfg3.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    n=re.sub(r'\(.*','',r[ki]).replace('void ','').replace('b200mix::','')
    agg.setdefault(n,[]).append(float(r[vi].replace(',','')))
for n,v in agg.items(): print('%-40s n=%3d last=%8.1f us'%(n[:40],len(v),v[-1]/1000))
PY
  ;;
rv)
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard.py tests/test_gpu_atsize.py -m gpu -q --timeout 800 -k "reverb or golden or config3 or chain or slot" > gpurun_out/${tag}_pytest_rv.log 2>&1
  tail -4 gpurun_out/${tag}_pytest_rv.log
  timeout 600 python tools/bench_configs.py --config 3 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('cfg3', round(j['ms_per_update'],4), j['stage_us_rank0'])" ;;
fx)
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_conv -c 60 --csv \
     --log-file gpurun_out/${tag}_launches_fx_conv.csv python tools/bench_effects.py --effect conv --voices 4096 --slots 32 --steps 4 > gpurun_out/${tag}_fx_conv_ncu.log 2>&1
  timeout 300 python tools/bench_effects.py --effect conv --voices 4096 --slots 32 > gpurun_out/${tag}_fx_conv.log 2>&1
  timeout 300 python tools/bench_effects.py --effect conv --voices 4096 --slots 16 >> gpurun_out/${tag}_fx_conv.log 2>&1
  tail -2 gpurun_out/${tag}_fx_conv.log; grep k_conv_mac gpurun_out/${tag}_launches_fx_conv.csv | tail -3 ;;
params)
  timeout 900 python -m pytest tests/test_gpu_params.py tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -s -k "sources_update or convolution or config2 or golden" > gpurun_out/${tag}_pytest_params.log 2>&1
  tail -8 gpurun_out/${tag}_pytest_params.log ;;
ncu_mix)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mix_voices -s 4 -c 1 -o gpurun_out/${tag}_prof_mix -f \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sustained > gpurun_out/${tag}_ncu_mix.log 2>&1
  tail -2 gpurun_out/${tag}_ncu_mix.log; ls -la gpurun_out/${tag}_prof_mix.ncu-rep ;;
ncu_conv)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_mac -s 4 -c 1 -o gpurun_out/${tag}_prof_conv_mac -f \
     python tools/bench_effects.py --effect conv --voices 4096 --slots 32 --steps 2 > gpurun_out/${tag}_ncu_conv.log 2>&1
  tail -2 gpurun_out/${tag}_ncu_conv.log ;;
ab4a)
  for env in "X=0" "B200MIX_PANMIX_SIMT=1" "B200MIX_MIX_GATHER=1" "B200MIX_PANMIX_SIMT=1 B200MIX_MIX_GATHER=1"; do
    echo "== $env" >> gpurun_out/${tag}_ab4a.log
    env $env timeout 600 python -m pytest tests/test_gpu_atsize.py -m gpu -q --timeout 600 -x -k "config4a" 2>&1 | grep -E "passed|failed|^E .*rms" >> gpurun_out/${tag}_ab4a.log
  done
  cat gpurun_out/${tag}_ab4a.log ;;
l4a)
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 40 --csv \
     --log-file gpurun_out/${tag}_launches_cfg4a.csv python tools/bench_configs.py --config 4a --voices 8192 --steps 2 --warmup 2 > gpurun_out/${tag}_cfg4a_ncu.log 2>&1
  grep -E "k_send|k_panmix|k_reduce|k_mix" gpurun_out/${tag}_launches_cfg4a.csv | awk -F'","' '{print $5, $9, $NF}' | tail -14 ;;
dropin)
  timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -q --timeout 800 > gpurun_out/${tag}_pytest_dropin.log 2>&1
  grep -E '^E .*b200mix|^E .*rms' gpurun_out/${tag}_pytest_dropin.log | cut -c1-300 | head -8; tail -3 gpurun_out/${tag}_pytest_dropin.log ;;
probe)
  ./tools/ubench/umma_probe > gpurun_out/${tag}_umma_probe.log 2>&1; cat gpurun_out/${tag}_umma_probe.log ;;
tc)
  ./tools/ubench/panmix_tc_test 1000 16 > gpurun_out/${tag}_panmix_tc.log 2>&1
  ./tools/ubench/panmix_tc_test 8192 16 >> gpurun_out/${tag}_panmix_tc.log 2>&1
  ./tools/ubench/panmix_tc_test 77 9 >> gpurun_out/${tag}_panmix_tc.log 2>&1
  cat gpurun_out/${tag}_panmix_tc.log
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_atsize.py -m gpu -q --timeout 600 -x -k "ambi3 or config4a or dry_mix" > gpurun_out/${tag}_pytest_tc.log 2>&1
  tail -5 gpurun_out/${tag}_pytest_tc.log
  for m in 0 1; do B200MIX_PANMIX_SIMT=$m timeout 300 python tools/bench_configs.py --config 4a --voices 8192 --steps 8 >> gpurun_out/${tag}_cfg4a.jsonl 2>> gpurun_out/${tag}_cfg4a.err; done
  cat gpurun_out/${tag}_cfg4a.jsonl ;;
esac
done
