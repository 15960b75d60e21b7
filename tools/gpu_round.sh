#!/bin/bash
# usage (on the GPU box, from the repo root): tools/gpu_round.sh TAG [full]
#   pytest -m gpu, the default bench line, the ncu launch list of a short bench run and -- with `full` -- one
#   `ncu --set full` capture of the tile kernel.  Everything lands under gpurun_out/ with TAG in the name;
#   tools/make_profiles.py turns the pulled files into the summaries committed under profiles/.
TAG=${1:-r02}
FULL=$2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_${TAG}.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_${TAG}.log
tail -3 gpurun_out/pytest_${TAG}.log
export HGP_BENCH_WATCHDOG=900
timeout 1000 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"
cut -c1-600 gpurun_out/bench_${TAG}.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err
echo "ref rc=$?"
OURS='regex:score_tiles|score_pairs|score_groups|score_blocks|snr_|lead_weights|hmm_scan|hard_resp|stats_|cholinv|chol_kernel|tri_inverse|pack_|whiten|tile_uniform'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-peak --no-cfg5 --no-fit > gpurun_out/ncu_${TAG}.log 2>&1
echo "launch list rc=$?"
if [ -n "$FULL" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_tiles_kernel -s 4 -c 1 \
      -o gpurun_out/prof_tiles_${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu --no-peak --no-cfg5 --no-fit \
      > gpurun_out/ncu_full_${TAG}.log 2>&1
  echo "ncu full rc=$?"
  ncu -i gpurun_out/prof_tiles_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_tiles_${TAG}_raw.csv 2>/dev/null
fi
