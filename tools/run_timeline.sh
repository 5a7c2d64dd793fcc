set -e
make -C jampr_plus_b200/csrc clean > /dev/null
make -C jampr_plus_b200/csrc -j16 EXTRA=-DJAMPR_TIMELINE > /dev/null 2>&1
python tools/timeline.py 3 > gpurun_out/r2_timeline_split.txt 2>&1 || true
cat gpurun_out/r2_timeline_split.txt
