#!/bin/bash
# look-ahead Cholesky: bit-identity against the plain order (both variants, both panel widths, 256 and 512 bit), tests, bench
mkdir -p gpurun_out
: > gpurun_out/w_chol.log
for noinv in 0 1; do for la in 0 1; do
  echo "== noinv=$noinv lookahead=$la" >> gpurun_out/w_chol.log
  CLRS_MP_CHOL_NOINV=$noinv CLRS_CHOL_LOOKAHEAD=$la timeout 120 python tools/gpu_chol_check.py 256 33 100 300 500 2>&1 | grep "^n=" >> gpurun_out/w_chol.log
  CLRS_MP_CHOL_NOINV=$noinv CLRS_CHOL_LOOKAHEAD=$la timeout 120 python tools/gpu_chol_check.py 512 200 640 2>&1 | grep "^n=" >> gpurun_out/w_chol.log
done; done
python - <<'PY'
import re
blocks = open('gpurun_out/w_chol.log').read().split('== ')[1:]
d = {}
for b in blocks:
    head, *lines = b.strip().splitlines()
    d[head] = {(m.group(1), m.group(2)): (m.group(3), m.group(4)) for m in (re.match(r'n=(\d+) prec=(\d+) sha=(\w+) wall_ms=([\d.]+)', l) for l in lines) if m}
for noinv in (0, 1):
    a, b = d[f'noinv={noinv} lookahead=0'], d[f'noinv={noinv} lookahead=1']
    for k in a:
        print('noinv', noinv, k, 'IDENTICAL' if a[k][0] == b.get(k, ('', ''))[0] else 'DIFFERENT', 'ms', a[k][1], '->', b.get(k, ('', '?'))[1])
PY
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) > gpurun_out/w_pytest.log 2>&1; tail -3 gpurun_out/w_pytest.log
for la in 0 1; do
CLRS_CHOL_LOOKAHEAD=$la timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-time-to-gap > gpurun_out/w_bench_la$la.json 2> gpurun_out/w_bench_la$la.err; echo "bench la=$la rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/w_bench_la$la.json'))
print('la=$la', round(d['ms_per_step'],3), 'cholS', d['phase_ms']['cholS'], 'Xinv', d['phase_ms']['Xinv'])
for k,v in d.get('configs',{}).items(): print('  ',k, round(v.get('ms_per_step',0),3), {a:b for a,b in v.get('phase_ms',{}).items() if a in ('cholS','cholQ','Xinv','LinvB','solve')})
PY
done
