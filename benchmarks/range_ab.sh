#!/bin/bash
# A/B timing of the headline range step: bash benchmarks/range_ab.sh [ENV=VALUE ...]  (one extra run per argument)
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'hits', d['config']['hits_per_query'], 'e2e', d['e2e']['value'])"; }
run A=0
for v in "$@"; do run $v; done
