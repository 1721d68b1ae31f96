#!/bin/bash
# Round artefacts on the GPU box: bench line, reference arm, ncu launch list of one steady-state rollout, ncu full
# captures of the dominant kernels.  Outputs under gpurun_out/.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r1_v2}
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
N=$(python tools/profile_target.py 1 0 | awk '/rollout 1:/ {print $3}')
echo "launches per rollout: $N"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $N -c $N --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_target.py 1 0 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_layer" -s 20 -c 1 -o gpurun_out/${TAG}_layer python tools/profile_target.py 1 0 > gpurun_out/${TAG}_ncu_layer.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fourier|k_embed_column" -s 36 -c 2 -o gpurun_out/${TAG}_fourier python tools/profile_target.py 1 0 > gpurun_out/${TAG}_ncu_fourier.log 2>&1
ls -la gpurun_out
