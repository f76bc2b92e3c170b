#!/bin/bash
cp grid2op_b200/libb200pf.so /tmp/inline.so
for v in inline ni; do
if [ $v == ni ]; then cp grid2op_b200/libb200pf_ni.so grid2op_b200/libb200pf.so; fi
echo "== $v"; timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --batch 148 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('b148', d['value'], d['ms_per_step'])"
done
echo "== pytest (ni)"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
