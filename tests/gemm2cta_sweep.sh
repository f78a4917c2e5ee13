# A/B of the CTA-pair projection kernel inside the real decode (not a test): default vs SCV_GEMM_2CTA=3 (N >= 768) vs =1
python tests/gemm_bench.py 4096 2>&1 | tail -7
for v in 0 3 1; do SCV_GEMM_2CTA=$v python bench.py --steps 3 --warmup 2 --cpu-rows 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('SCV_GEMM_2CTA=$v', round(d['value']), round(d['ms_per_step'],2), 'gemm ms', d['kernels']['gemm_tcgen05']['ms'])"; done
