#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, then (optionally) the ncu launch list and a full capture.
#   tools/gpu_round.sh [tests] [bench] [launches] [full:<kernel-regex>]
# Everything is written under gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for what in "$@"; do
  case "$what" in
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
      echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
      tail -5 gpurun_out/pytest_gpu.log
      ;;
    smoke)
      timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
      echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
      ;;
    bench)
      timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
      echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
      ;;
    benchref)
      timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
      echo "benchref rc=$?"; cat gpurun_out/bench_ref.json
      ;;
    c5)
      timeout 900 python bench.py --config C5 --steps 10 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
      echo "c5 rc=$?"; cat gpurun_out/bench_c5.json; tail -3 gpurun_out/bench_c5.err
      ;;
    c5x2)
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
          bench.py --gpus 2 --config C5 --steps 10 > gpurun_out/bench_c5x2.json 2> gpurun_out/bench_c5x2.err
      echo "c5x2 rc=$?"; cat gpurun_out/bench_c5x2.json; tail -3 gpurun_out/bench_c5x2.err
      ;;
    c2x2)
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
          bench.py --gpus 2 > gpurun_out/bench_x2.json 2> gpurun_out/bench_x2.err
      echo "c2x2 rc=$?"; cat gpurun_out/bench_x2.json; tail -3 gpurun_out/bench_x2.err
      ;;
    scale:*)
      NG="${what#scale:}"
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 \
          bench.py --gpus $NG --steps 10 --cpu-images 64 > gpurun_out/bench_x$NG.json 2> gpurun_out/bench_x$NG.err
      echo "c2 x$NG rc=$?"; grep '^{' gpurun_out/bench_x$NG.json | cut -c1-400; tail -2 gpurun_out/bench_x$NG.err
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29514 \
          bench.py --gpus $NG --config C5 --steps 10 > gpurun_out/bench_c5x$NG.json 2> gpurun_out/bench_c5x$NG.err
      echo "c5 x$NG rc=$?"; grep '^{' gpurun_out/bench_c5x$NG.json | cut -c1-1300; tail -2 gpurun_out/bench_c5x$NG.err
      ;;
    launches)
      CMD="python bench.py --steps 2 --warmup 3 --e2e-steps 1 --cpu-images 8 --no-extras"
      timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
      timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
          --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
      echo "launches rc=$?"
      ;;
    full:*)
      KREGEX="${what#full:}"
      CMD="python bench.py --steps 1 --warmup 3 --e2e-steps 1 --cpu-images 8 --no-extras"
      timeout 600 $CMD > gpurun_out/plain_full.log 2>&1 &&
      timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:${KREGEX}" -s 3 -c 1 \
          -f -o "gpurun_out/prof_${KREGEX}" $CMD > gpurun_out/ncu_full_${KREGEX}.log 2>&1
      echo "full ${KREGEX} rc=$?"
      ;;
    fullc5:*)
      KREGEX="${what#fullc5:}"
      CMD="python bench.py --config C5 --steps 1 --warmup 3"
      timeout 600 $CMD > gpurun_out/plain_fullc5.log 2>&1 &&
      timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:${KREGEX}" -s 3 -c 1 \
          -f -o "gpurun_out/prof_${KREGEX}" $CMD > gpurun_out/ncu_full_${KREGEX}.log 2>&1
      echo "fullc5 ${KREGEX} rc=$?"
      ;;
    debug:*)
      timeout 600 python tools/gpu_debug.py "${what#debug:}" 2>&1 | tee -a gpurun_out/debug.log
      ;;
  esac
done
