export B200L_LIB=$PWD/_ab_old/libb200lasso_head.so
r=$(timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu --eps 0 --quick --no-parity --e2e-sweeps 1 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['ms_per_step'])")
echo "head: $r"
unset B200L_LIB
for d in 0 16777216 25165824 8388608 0 16777216; do
r=$(timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu --eps 0 --quick --no-parity --e2e-sweeps 1 --dbg $d 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['ms_per_step'])")
echo "new dbg=$d: $r"
done
