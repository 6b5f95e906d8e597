#!/bin/bash
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'hits', d['config']['hits_per_query'], 'e2e', d['e2e']['value'])"; }
run A=0
for v in "$@"; do run DRRT_B200_LIB=/root/repo/drrt_b200/build/var_$v/libdrrt.so; done
