for mb in 0 64; do
  echo "persist_mb $mb"
  MAMG_L2_PERSIST_MB=$mb python bench.py -n ${1:-128} --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['vcycle_ms'], {k:(v['ms'],v['alg_GBs']) for k,v in d['kernels'].items() if k in ('schwarz','gs','spmv')})"
done
