#!/bin/bash
# Round artefacts on the GPU box: bench line, reference arm, ncu launch list of one steady-state rollout, ncu full
# captures of the dominant kernels (single scene: k_layer, k_fourier_tc; batch of 32 scenes: k_attn, k_node).
# Outputs under gpurun_out/.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r1_v4}
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_target.py 1 0 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_layer" -s 8 -c 1 -o gpurun_out/${TAG}_layer python tools/profile_target.py 1 0 > gpurun_out/${TAG}_ncu_layer.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_fourier_tc" -s 8 -c 1 -o gpurun_out/${TAG}_fourier python tools/profile_target.py 1 0 > gpurun_out/${TAG}_ncu_fourier.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches_b32.csv python tools/profile_target.py 32 0 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_attn|k_node" -s 200 -c 6 -o gpurun_out/${TAG}_rows python tools/profile_target.py 32 0 > gpurun_out/${TAG}_ncu_rows.log 2>&1
for f in layer fourier rows; do
  ncu -i gpurun_out/${TAG}_$f.ncu-rep --page raw --csv > gpurun_out/${TAG}_${f}_raw.csv 2>/dev/null
done
ncu -i gpurun_out/${TAG}_rows.ncu-rep --page source --csv > gpurun_out/${TAG}_rows_src.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_fourier.ncu-rep --page source --csv > gpurun_out/${TAG}_fourier_src.csv 2>/dev/null
ls -la gpurun_out | tail -20
