#!/bin/bash
# experiment sweep for the CTA-per-query range kernel (debug knobs, not product paths)
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'hits', d['config']['hits_per_query'])"; }
run A=0
run DRRT_EXP_ALLOC=1
run DRRT_EXP_UNR=1
run DRRT_EXP_UNR=4
run DRRT_EXP_UNR=4 DRRT_EXP_ALLOC=1
run DRRT_EXP_XYSCALE=1.26
run DRRT_EXP_XYSCALE=1.59
run DRRT_EXP_XYSCALE=2.0
run DRRT_EXP_OCC=4
run DRRT_EXP_OCC=1
run DRRT_EXP_XYSCALE=1.59 DRRT_EXP_OCC=4
