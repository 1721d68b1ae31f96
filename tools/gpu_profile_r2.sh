#!/bin/bash
# Round-2 ncu artefacts (GPU box): launch lists (gpu__time_duration per launch) of one steady-state rollout of configs[1]
# and of the configs[2] shape, both with the insertion stage ON (plain launches: ncu serialises kernels, the WHILE / IF graph
# is replaced by the host-driven loop over the same enqueue functions), and --set full captures of the dominant kernels.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r2
NCU="ncu --clock-control none --profile-from-start off"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${T}_launches_c1.csv python tools/profile_target.py 1 0 1 > /dev/null 2>&1
timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${T}_launches_c2.csv python tools/profile_target.py 32 0 1 > /dev/null 2>&1
# the 18-layer launch of the motion stage (the roofline kernel of configs[1]): motion-only rollout, so that every k_layer
# launch after the prefill is one of them (with the insertion stage on, the skip count lands on single-row launches)
timeout 600 $NCU --set full --import-source on -k regex:"k_layer" -s 6 -c 4 -o gpurun_out/${T}_layer python tools/profile_target.py 1 0 0 > gpurun_out/${T}_ncu_layer.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:"k_attn|k_node_tc" -s 300 -c 12 -o gpurun_out/${T}_rows python tools/profile_target.py 32 0 1 > gpurun_out/${T}_ncu_rows.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:"k_fourier_tc" -s 4 -c 2 -o gpurun_out/${T}_fourier python tools/profile_target.py 32 0 1 > gpurun_out/${T}_ncu_fourier.log 2>&1
for f in layer rows fourier; do
  ncu -i gpurun_out/${T}_$f.ncu-rep --page raw --csv > gpurun_out/${T}_${f}_raw.csv 2>/dev/null
done
ls -la gpurun_out | tail -12
