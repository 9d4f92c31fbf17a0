for v in "" $VARIANTS; do
  if [ -z "$v" ]; then lib=""; name=default; else lib=$PWD/metric_amg_examples_b200/libmamg_$v.so; name=$v; fi
  echo "== $name"
  MAMG_LIB=$lib python bench.py -n ${N:-128} --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['vcycle_ms'], {k:(v['ms'],v['alg_GBs']) for k,v in d['kernels'].items() if k in ('spmv','gs','schwarz','restrict','scale')})"
done
