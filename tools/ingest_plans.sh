for plan in 4,2,32 3,2,32 4,1,32 4,2,64 6,2,32 4,3,32 2,2,32 4,2,16; do
  echo "plan $plan"; FRISK_INGEST_PLAN=$plan timeout 200 python tools/run_fasta_stages.py 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
for r in d['rows']: print(r['option'], r['uploaded/counted/finalised/scored/end/wall_ms'])
"
done
