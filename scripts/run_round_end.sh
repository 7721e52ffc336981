#!/bin/bash
# Round-end evidence run on one B200 (results under gpurun_out/, copied to profiles/ afterwards):
# sanitizers over the tcgen05 up-step tests, ncu of convt4_umma_kernel, the per-config bench lines.
set -u
prefix=${1:-r02b}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  log=gpurun_out/${prefix}_sanitizer_umma_${tool}.log
  sel='convt4_umma and (dims0 or dims4)'
  [ "$tool" = memcheck ] && sel='convt4_umma and (dims0 or dims3 or dims4) or stage_level_c_abi_registers'
  echo "# compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_reg_gpu.py -q -m gpu -x -k \"$sel\"" > "$log"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_reg_gpu.py -q -m gpu -x -k "$sel" 2>&1 | grep -v "^$" | tail -12 >> "$log"
  echo "exit ${PIPESTATUS[0]}" >> "$log"
done
# ncu: one full-set capture of the largest up-step launch (48 -> 16 at 40x96x96) and of 96 -> 32
OAI_BENCH_ONLY=umma timeout 600 ncu --set full --clock-control none --import-source on -k regex:convt4_umma -c 2 \
  -o gpurun_out/${prefix}_ncu_convt4_umma -f python scripts/bench_convt4.py 0 > gpurun_out/${prefix}_ncu_convt4_umma.log 2>&1
ncu -i gpurun_out/${prefix}_ncu_convt4_umma.ncu-rep --page raw --csv > gpurun_out/${prefix}_ncu_convt4_umma_raw.csv 2>/dev/null
for cfg in reg seg warp-sweep; do
  python bench.py --config $cfg --steps 10 --warmup 3 > gpurun_out/${prefix}_bench_$cfg.json 2> gpurun_out/${prefix}_bench_$cfg.err
done
python scripts/bench_convt4.py > gpurun_out/${prefix}_convt4_layers.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${prefix}_launches_step.csv python scripts/profile_step.py > gpurun_out/${prefix}_profile_step.log 2>&1
tail -n 3 gpurun_out/${prefix}_sanitizer_umma_memcheck.log gpurun_out/${prefix}_sanitizer_umma_racecheck.log
