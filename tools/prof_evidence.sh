# developer helper (GPU box): the round's evidence set under gpurun_out/<tag>_* -- GPU suite, bench lines (own arm with the per-layer table, reference
# arm), ncu launch list of the bench command, one `ncu --set full` capture per dominant kernel with its raw-metric + source-level summaries.
#   usage: bash tools/prof_evidence.sh <tag>        (copy what is to be judged from gpurun_out/ into profiles/)
TAG=${1:-cur}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4) > gpurun_out/${TAG}_pytest_gpu.txt; cat gpurun_out/${TAG}_pytest_gpu.txt
timeout 600 python bench.py --layers > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_layers.txt; grep "rank 0" gpurun_out/${TAG}_layers.txt
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 300 gpurun_out/${TAG}_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 150 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_launches.csv "ncu launch list of: python bench.py --steps 4 --warmup 3 --no-cpu-baseline (launches 400-549)" > gpurun_out/${TAG}_launches_summary.txt
for spec in "k_block_mma:6:blockmma_L38" "k_block_mma:0:blockmma_L12" "k_block_mma:11:blockmma_L61" "k_block_reg_s2:0:blockreg_L9" "k_block_reg_s1:0:blockreg_L1" \
            "k_stem_block:0:stemblock_L0" "k_dw5s1_tma:2:dw5_L125" "k_pw_tc:6:pwtc_L129" "k_upsample:0:upsample_L123"; do
  k=${spec%%:*}; rest=${spec#*:}; s=${rest%%:*}; name=${rest#*:}
  bash tools/ncu_any.sh $k $s ${TAG}_ncu_$name > /dev/null 2>&1
  (python tools/ncu_raw.py gpurun_out/${TAG}_ncu_$name.ncu-rep; python tools/ncu_hot.py gpurun_out/${TAG}_ncu_$name.ncu-rep 12) > gpurun_out/${TAG}_ncu_$name.txt 2>&1
  head -3 gpurun_out/${TAG}_ncu_$name.txt
  rm -f gpurun_out/${TAG}_ncu_$name.ncu-rep          # gpurun copies back at most 64 MiB: the text summaries travel, the reports do not
done
rm -f gpurun_out/${TAG}_ncu_*.log
ls gpurun_out | head -60
