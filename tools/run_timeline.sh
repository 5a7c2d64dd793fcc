set -e
# per-phase time lines of the two policy-step kernels (instrumented side build, see csrc/Makefile VARIANT)
make -C jampr_plus_b200/csrc -j16 VARIANT=_tl EXTRA=-DJAMPR_TIMELINE > /dev/null 2>&1
export JAMPR_B200_LIB=$PWD/jampr_plus_b200/lib/libjampr_b200_tl.so
python tools/timeline.py 3 > gpurun_out/r2_timeline_split.txt 2>&1 || true
python tools/dtimeline.py 30 > gpurun_out/r2_dtimeline.txt 2>&1 || true
cat gpurun_out/r2_timeline_split.txt gpurun_out/r2_dtimeline.txt
