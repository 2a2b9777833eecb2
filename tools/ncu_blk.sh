# developer helper: ncu --set full capture of selected fused-block launches (launch indices in $LAUNCHES, default "0 9 19")
for k in ${LAUNCHES:-0 9 19}; do
  SKIP_PARITY=1 BATCH=256 timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_block_mma -s $k -c 1 -o gpurun_out/blk_${TAG:-cur}_$k -f python tools/blk_check.py > gpurun_out/ncu_blk_$k.log 2>&1
  tail -2 gpurun_out/ncu_blk_$k.log
done
