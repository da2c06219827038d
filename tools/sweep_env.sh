# usage: tools/sweep_env.sh VAR v1 v2 ... -- runs the channels-last bench once per value of an environment knob
VAR=$1; shift
for v in "$@"; do
  echo "== $VAR=$v"
  env $VAR=$v python bench.py --steps 10 --warmup 3 --layout nhwc --no-cpu-baseline --e2e-steps 0 --no-other-layout --no-pyramids --torch-cuda-steps 0 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('fwd %.3f ms %.0f GB/s | bwd %.3f ms %.0f GB/s | fwd+bwd %.3f ms frac %.3f' % (r['fwd']['ms'], r['fwd']['achieved'], r['bwd']['ms'], r['bwd']['achieved'], r['fwd_bwd']['ms'], r['fwd_bwd']['frac']))
    else: print(l.rstrip())
"
done
