export SCV_KV_BITS=32
run() { echo -n "$1: "; env $1 timeout 200 python bench.py --steps 8 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2))"; }
for rep in 1 2; do
run "SCV_ATT_UNR=4"
run "SCV_ATT_UNR=2"
done
