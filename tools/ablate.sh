# needs a library built with: make -C jampr_plus_b200/csrc clean all EXTRA=-DJAMPR_ABLATE_BUILD
for a in 0 1 2 4 8 16 3 19 31; do
  JAMPR_TC_ABLATE=$a timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ablate=$a', 'launch_ms', round(d['roofline']['avg_launch_ms'],2), 'value', round(d['value']))
"
done
