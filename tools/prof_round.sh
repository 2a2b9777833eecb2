# developer helper (GPU box): the per-round profiling passes of B200_PROFILING.md for the bench command.
#   $1 = tag (e.g. r1l).  Writes gpurun_out/<tag>_launches.csv and gpurun_out/<tag>_<kernel>.ncu-rep
TAG=${1:-cur}
# 1. launch list of one steady-state forward pass: skip the warm-up + capture launches, take two steps' worth
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 150 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
# 2. one --set full capture per dominant kernel family (launch indices inside the first eager forward pass)
for spec in "k_block_mma:6:blockmma_L38" "k_block_mma:11:blockmma_L61" "k_pw_tc:8:pwtc_L129" "k_yolo_filter:1:yolo_L130"; do
  k=${spec%%:*}; rest=${spec#*:}; s=${rest%%:*}; name=${rest#*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o gpurun_out/${TAG}_$name -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_$name.log 2>&1
  tail -1 gpurun_out/${TAG}_$name.log
done
ls -la gpurun_out | tail -12
