# k_node / k_attn timing with the row-tile path forced, 1 / 8 / 32 scenes (4 / 32 / 128 k_node CTAs)
for sc in 1 8 32; do
INFGEN_LAYER_PATH=rows python bench.py --scenes $sc --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_node_${sc}.json 2> gpurun_out/bench_node_${sc}.err
done
