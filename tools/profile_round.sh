#!/bin/bash
# Evidence run of one round on the GPU box (one GPU): the headline bench line, the ncu launch list of the same command,
# and ncu --set full captures of the GEMM, attention and LayerNorm kernels.  Everything lands in gpurun_out/; the
# summaries that are committed under profiles/ are produced from these files by tools/summarize_launches.py and
# tools/ncu_summary.py.   usage: gpurun --timeout 1500 -- 'bash tools/profile_round.sh r01c'
set -u
tag=${1:-rXX}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || exit 1
NCU="ncu --clock-control none"
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file gpurun_out/launches_${tag}.csv $B > gpurun_out/ncu_launches_${tag}.log 2>&1
$NCU --set full --import-source on -k regex:gemm_kernel --launch-skip 30 -c 8 -f -o gpurun_out/prof_gemm_${tag} $B > gpurun_out/ncu_gemm_${tag}.log 2>&1
$NCU --set full --import-source on -k regex:attn_ -c 12 -f -o gpurun_out/prof_attn_${tag} python tools/attn_bench.py 1 > gpurun_out/ncu_attn_${tag}.log 2>&1
$NCU --set full --import-source on -k regex:ln_mod --launch-skip 8 -c 4 -f -o gpurun_out/prof_ln_${tag} $B > gpurun_out/ncu_ln_${tag}.log 2>&1
ls -la gpurun_out/*${tag}*
tail -c 600 gpurun_out/bench_${tag}.json
