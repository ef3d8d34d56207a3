#!/bin/bash
# Compile csrc/pmg_apply.cu with -Xptxas -v and print registers / spills per sweep-kernel instantiation.
cd /root/repo/portable-multigrid_b200 || exit 1
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I../include -Icsrc -Ihost -Xptxas -v -c csrc/pmg_apply.cu -o build/pmg_apply.cu.o > /tmp/ptxas.log 2>&1
grep -i "error" /tmp/ptxas.log | head
python - <<'PY'
import re
txt=open('/tmp/ptxas.log').read()
for m in re.finditer(r"Compiling entry function '(\w+)'.*?\n(.*?)\n.*?Used (\d+) registers", txt, re.S):
    name=m.group(1)
    if 'sweep' not in name: continue
    nums=re.findall(r'ILi(\d+)E', name)
    spill=re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', m.group(0))
    print('P,BX,BY,SY,NT,MINB =', ','.join(nums), '| regs', m.group(3), '| stack/spill st/ld', spill.groups() if spill else None)
PY
