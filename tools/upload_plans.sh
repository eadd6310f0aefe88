# e2e (pinned planes -> rows) stage marks of bench.py --quick under different upload chunk plans (sixteenths per chunk)
for plan in 16 8,8 10,6 12,4 8,5,3 6,5,3,2 9,7 7,5,4 11,5; do
  FRISK_UPLOAD_PLAN=$plan timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d['e2e']['stages']
print('$plan', round(d['e2e']['ms_per_step'],4), {k: round(v['max_ms'],3) for k,v in s.items() if isinstance(v,dict)})
"
done
