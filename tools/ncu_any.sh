# developer helper: ncu --set full of launch $2 (0-based) of kernels matching regex $1 in one blk_check forward; report tag $3
SKIP_PARITY=1 BATCH=256 timeout 200 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/$3 -f python tools/blk_check.py > gpurun_out/$3.log 2>&1
tail -2 gpurun_out/$3.log
