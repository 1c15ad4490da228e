for i in 1 2 3; do
  for L in libA.so libb200lasso.so; do
    v=$(B200L_LIB=$PWD/convex_optimization_b200/$L python bench.py --no-cpu --eps 0 --steps 40 $EXTRA 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['e2e']['value'],1))")
    echo "$L $v"
  done
done
