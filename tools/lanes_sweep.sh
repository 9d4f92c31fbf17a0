for cfg in "8 2" "16 2" "16 4" "4 4"; do set -- $cfg
  echo "LANES=$1 UNROLL=$2"
  MAMG_LANES=$1 MAMG_UNROLL=$2 python bench.py -n 199 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['vcycle_ms'], {k:(v['ms'],v['alg_GBs']) for k,v in d['kernels'].items() if k in ('gs','spmv','restrict','scale')}, d['gs_ms_by_level'][:6])"
done
