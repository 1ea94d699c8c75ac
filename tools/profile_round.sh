#!/bin/bash
# Evidence run of one round on the GPU box (one GPU): the headline bench line, the ncu launch list of one training step,
# the per-shape GEMM traffic list, and ncu --set full captures of the GEMM, attention and LayerNorm kernels.  Everything
# lands in gpurun_out/; the summaries that are committed under profiles/ are produced from these files by
# tools/summarize_launches.py, tools/traffic_table.py and tools/ncu_summary.py.
#   usage: gpurun --timeout 1500 -- 'bash tools/profile_round.sh r02'
set -u
tag=${1:-rXX}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || exit 1
NCU="ncu --clock-control none"
ONE="python tools/step_once.py"
# (1) launch list of one whole step (every kernel, durations only)
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_${tag}.csv $ONE > gpurun_out/ncu_launches_${tag}.log 2>&1
# (2) DRAM bytes / tensor-pipe activity of every GEMM launch of the step + the shapes the library logged, in order
rm -f gpurun_out/gemm_log_${tag}.txt
UMD_GEMM_LOG=gpurun_out/gemm_log_${tag}.txt $NCU -k regex:gemm_kernel \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --csv --log-file gpurun_out/gemm_launches_${tag}.csv $ONE > gpurun_out/ncu_gemmlist_${tag}.log 2>&1
# (3) --set full captures (source view included) of a few launches of each hot kernel
$NCU --set full --import-source on -k regex:gemm_kernel --launch-skip 30 -c 8 -f -o gpurun_out/prof_gemm_${tag} $ONE > gpurun_out/ncu_gemm_${tag}.log 2>&1
$NCU --set full --import-source on -k regex:attn_ -c 12 -f -o gpurun_out/prof_attn_${tag} python tools/attn_bench.py 1 > gpurun_out/ncu_attn_${tag}.log 2>&1
$NCU --set full --import-source on -k regex:ln_mod --launch-skip 8 -c 4 -f -o gpurun_out/prof_ln_${tag} $ONE > gpurun_out/ncu_ln_${tag}.log 2>&1
$NCU --set full --import-source on -k regex:colsum_bf16 --launch-skip 2 -c 2 -f -o gpurun_out/prof_colsum_${tag} $ONE > gpurun_out/ncu_colsum_${tag}.log 2>&1
ls -la gpurun_out/*${tag}*
tail -c 600 gpurun_out/bench_${tag}.json
