for n in 1 2 3 4; do SCV_SUBBATCHES=$n python bench.py --steps 3 --warmup 2 --cpu-rows 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('subbatches $n', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; done
