for a in 32 8 40; do
  JAMPR_TC_ABLATE=$a timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ablate=$a', 'launch_ms', round(d['roofline']['avg_launch_ms'],2))
"
done
