#!/bin/bash
# One GPU-box session: parity tests, bench lines, ncu launch list, ncu --set full captures, GAE sweep.
# Usage (from the repo root, under gpurun):  bash tools/gpu_session.sh <tag> [what ...]
#   what: tests bench_c2 bench_c3 launches full_c2 full_c3 sweep dp   (default: all but full_c3 and dp)
set -u
TAG=${1:-r01}
shift || true
WHAT=${*:-tests bench_c2 launches full_c2 sweep bench_c3}
O=gpurun_out
mkdir -p $O
has() { [[ " $WHAT " == *" $1 "* ]]; }

if has gae; then
  timeout 600 python -m pytest tests -m gpu -x -q -k gae > $O/${TAG}_pytest_gae.log 2>&1
  tail -3 $O/${TAG}_pytest_gae.log
  timeout 600 python tools/gae_sweep.py --min-log2 22 --max-log2 28 --check > $O/${TAG}_gae_sweep.jsonl 2> $O/${TAG}_gae_sweep.err
  cut -c1-200 $O/${TAG}_gae_sweep.jsonl
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $O/${TAG}_full_gae \
      python tools/ncu_step.py --no-step --gae-log2 26 > $O/${TAG}_full_gae.log 2>&1
  tail -2 $O/${TAG}_full_gae.log
fi
if has tests; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1
  tail -3 $O/${TAG}_pytest_gpu.log
fi
if has bench_c2; then
  timeout 600 python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err
  cut -c1-600 $O/${TAG}_bench_c2.json
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_c2_reference.json 2>> $O/${TAG}_bench_c2.err
fi
if has launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file $O/${TAG}_launches_c2.csv python tools/ncu_step.py --workload c2 > $O/${TAG}_launches_c2.log 2>&1
  python tools/summarize_launches.py $O/${TAG}_launches_c2.csv > $O/${TAG}_launches_c2.md 2>&1
  cat $O/${TAG}_launches_c2.md
fi
if has full_c2; then
  timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $O/${TAG}_full_c2 \
      python tools/ncu_step.py --workload c2 --gae-log2 26 > $O/${TAG}_full_c2.log 2>&1
  tail -2 $O/${TAG}_full_c2.log
fi
if has sweep; then
  timeout 900 python tools/gae_sweep.py --max-log2 30 --check > $O/${TAG}_gae_sweep.jsonl 2> $O/${TAG}_gae_sweep.err
  cat $O/${TAG}_gae_sweep.jsonl | cut -c1-220
fi
if has bench_c3; then
  timeout 900 python bench.py --workload c3 --steps 5 --warmup 3 > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err
  cut -c1-600 $O/${TAG}_bench_c3.json
fi
if has launches_c3; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file $O/${TAG}_launches_c3.csv python tools/ncu_step.py --workload c3 > $O/${TAG}_launches_c3.log 2>&1
  python tools/summarize_launches.py $O/${TAG}_launches_c3.csv > $O/${TAG}_launches_c3.md 2>&1
  cat $O/${TAG}_launches_c3.md
fi
if has traffic_c3; then
  # DRAM traffic of EVERY GEMM launch of one c3 iteration (light sections only: ~230 launches)
  timeout 1500 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --clock-control none \
      --profile-from-start off -f -o $O/${TAG}_traffic_c3 -k regex:'rowgemm|wgrad|colsum|norm_clip' \
      python tools/ncu_step.py --workload c3 > $O/${TAG}_traffic_c3.log 2>&1
  tail -2 $O/${TAG}_traffic_c3.log
fi
if has gather; then
  timeout 300 python tools/gather_sweep.py > $O/${TAG}_gather_sweep.jsonl 2> $O/${TAG}_gather_sweep.err
  cat $O/${TAG}_gather_sweep.jsonl
fi
if has full_c3; then
  timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $O/${TAG}_full_c3 \
      -k regex:'rowgemm|wgrad' -c 12 python tools/ncu_step.py --workload c3 > $O/${TAG}_full_c3.log 2>&1
  tail -2 $O/${TAG}_full_c3.log
fi
if has dp; then
  # data-parallel check + bench lines with both gradient exchanges (run under `gpurun --gpus N`; N from the visible GPUs)
  N=$(nvidia-smi -L | wc -l)
  T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
  timeout 300 $T 29541 tests/dp_check.py > $O/${TAG}_dp_check_n$N.log 2>&1
  grep "dp_check\|replicated:" $O/${TAG}_dp_check_n$N.log
  for c in p2p nccl; do
    RLPPO_DP_COLLECTIVE=$c timeout 200 $T 29542 bench.py --gpus $N --steps 20 --warmup 3 \
        > $O/${TAG}_bench_c2_n${N}_$c.json 2> $O/${TAG}_bench_c2_n${N}_$c.err
    cut -c1-220 $O/${TAG}_bench_c2_n${N}_$c.json
  done
fi
ls -la $O | tail -20
