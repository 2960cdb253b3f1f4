#!/bin/bash
# A/B of two library builds on the attention bench shapes: tools/attn_ab.sh OUTDIR lib_a.so lib_b.so ...
out=$1; shift
mkdir -p $out
for L in "$@"; do
  n=$(basename $L .so)
  echo "== $n"
  PT_B200_LIB=$L timeout 300 python tools/attn_probe.py full_ 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    n, _, j = l.partition(' ')
    try:
        d = json.loads(j)
        print(n, 'fwd %.1f us bwd %.1f us' % (d.get('fwd_us', 0), d.get('bwd_us', 0)), 'err o %.1e dq %.1e dk %.1e dv %.1e' % (d['err_o'], d['err_dq'], d['err_dk'], d['err_dv']))
    except Exception:
        print(l.strip()[:300])
" | tee $out/probe_$n.txt
done
