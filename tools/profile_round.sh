#!/bin/bash
# Evidence run of one round on the GPU box (one GPU): the headline bench line, the ncu launch list of one training step,
# the per-shape GEMM traffic list, and ncu --set full captures of the GEMM, attention, LayerNorm and column-sum kernels.
# The captures are summarised ON THE BOX (tools/ncu_summary.py) and the .ncu-rep files deleted: gpurun only brings back
# 64 MiB.  Everything lands in gpurun_out/; the summaries are then copied under profiles/.
#   usage: gpurun --timeout 1500 -- 'bash tools/profile_round.sh r02'
set -u
tag=${1:-rXX}
O=gpurun_out
mkdir -p $O
python bench.py --steps 10 --warmup 3 > $O/bench_${tag}.json 2> $O/bench_${tag}.err || exit 1
NCU="ncu --clock-control none"
ONE="python tools/step_once.py"
# (1) launch list of one whole step (every kernel, durations only)
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_${tag}.csv $ONE > $O/ncu_launches_${tag}.log 2>&1
python tools/summarize_launches.py $O/launches_${tag}.csv > $O/launches_${tag}.txt 2>&1
# (2) DRAM bytes / tensor-pipe activity of every GEMM launch of the step + the shapes the library logged, in order
rm -f $O/gemm_log_${tag}.txt
UMD_GEMM_LOG=$O/gemm_log_${tag}.txt $NCU -k regex:gemm_kernel \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --csv --log-file $O/gemm_launches_${tag}.csv $ONE > $O/ncu_gemmlist_${tag}.log 2>&1
python tools/traffic_table.py $O/gemm_launches_${tag}.csv $O/gemm_log_${tag}.txt $O/traffic_table_${tag}.json > $O/traffic_table_${tag}.txt 2>&1
# (3) --set full captures of a few launches of each hot kernel, summarised here
full() {  # name, kernel regex, extra ncu args, command...
  local name=$1 rx=$2 extra=$3; shift 3
  $NCU --set full --import-source on -k regex:$rx $extra -f -o $O/prof_${name}_${tag} "$@" > $O/ncu_${name}_${tag}.log 2>&1
  python tools/ncu_summary.py $O/prof_${name}_${tag}.ncu-rep --title "${name} ${tag}" --json $O/ncu_${name}_${tag}.json > $O/ncu_${name}_${tag}.txt 2>&1
  rm -f $O/prof_${name}_${tag}.ncu-rep
}
full gemm gemm_kernel "--launch-skip 30 -c 10" $ONE
full attn attn_ "-c 12" python tools/attn_bench.py 1
full ln ln_mod "--launch-skip 8 -c 4" $ONE
full colsum colsum_bf16 "--launch-skip 2 -c 2" $ONE
rm -f $O/gemm_launches_${tag}.csv.tmp
ls -la $O/*${tag}*
du -sh $O
tail -c 600 $O/bench_${tag}.json
