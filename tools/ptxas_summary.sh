#!/bin/bash
# registers / spills / stack of every fused-kernel instantiation (cross-compiles without a GPU)
cd "$(dirname "$0")/.." || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v \
     -I include convex_optimization_b200/csrc/b200lasso.cu -o /tmp/ptxas_probe.so 2>&1 |
python3 -c "
import re,sys
txt=sys.stdin.read()
for m in re.finditer(r\"Compiling entry function '(\S+)'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers\", txt):
    name=m.group(1)
    if 'lasso_fused' in name or len(sys.argv)>1:
        print('%-70s regs=%s stack=%s spill_st=%s spill_ld=%s'%(name[:70],m.group(5),m.group(2),m.group(3),m.group(4)))
for l in txt.splitlines():
    if 'error' in l or 'warning' in l: print(l)
" "$@"
